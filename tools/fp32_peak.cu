// fp32_peak.cu -- measures the FP32 FMA issue ceiling of the GPU (scalar FFMA vs packed FFMA2) so the
// evaluation kernel's roofline has a measured denominator.  Build: nvcc -arch=sm_100a -O3 -o fp32_peak fp32_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>  // 0: scalar FFMA (3-reg), 1: FFMA2
__global__ void __launch_bounds__(1024) k(float* out, int iters, float s) {
    float2 acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    float a = s, b = s * 0.999f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) {
                acc[i].x = __fmaf_rn(a, acc[i].x, b);
                acc[i].y = __fmaf_rn(b, acc[i].y, a);
            } else {
                acc[i] = __ffma2_rn(make_float2(a, a), acc[i], make_float2(b, b));
            }
        }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(int threads, int blocks_per_sm, int sms, float* out) {
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms * blocks_per_sm, threads>>>(out, 100, 0.5f);
    cudaEventRecord(e0);
    k<MODE><<<sms * blocks_per_sm, threads>>>(out, iters, 0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double fma = (double)sms * blocks_per_sm * threads * iters * 32.0;
    printf("mode %s threads/SM %4d : %.2f TFLOP/s (%.1f FMA/clk/SM at 1965 MHz)\n", MODE ? "FFMA2" : "FFMA ", threads * blocks_per_sm,
           2 * fma / (ms * 1e-3) / 1e12, fma / (ms * 1e-3) / sms / 1.965e9);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    float* out;
    cudaMalloc(&out, sizeof(float) * 148 * 2048 * 2);
    for (int th : {128, 256, 512, 1024}) { run<0>(th, 1, p.multiProcessorCount, out); run<1>(th, 1, p.multiProcessorCount, out); }
    run<0>(1024, 2, p.multiProcessorCount, out);
    run<1>(1024, 2, p.multiProcessorCount, out);
    return 0;
}
