// Dependent-chain latencies of one warp on sm_100a: DFMA, DMUL+DADD, FFMA, IMAD, LDS, generic LD of a shared address, MUFU.RCP64H, f64 division.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/lat_probe tools/lat_probe.cu && tools/lat_probe
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
__global__ void k(long long* out, double x0, int i0, double* gsink) {
    __shared__ int chain[1024];
    __shared__ double dsh[64];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) chain[i] = (i * 37 + 11) & 1023;
    for (int i = threadIdx.x; i < 64; i += blockDim.x) dsh[i] = 1.0 / (i + 1);
    __syncthreads();
    long long t0, t1;
    double a = x0, b = x0 * 0.5;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = __fma_rn(a, b, x0);
    t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    double s = a;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = __dadd_rn(__dmul_rn(a, b), x0);
    t1 = clock64();
    if (threadIdx.x == 0) out[1] = t1 - t0;
    s += a;
    float f = (float)x0, g = f * 0.5f;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) f = __fmaf_rn(f, g, 1.0f);
    t1 = clock64();
    if (threadIdx.x == 0) out[2] = t1 - t0;
    s += f;
    unsigned u = i0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) u = u * 0xD2511F53u + 12345u;
    t1 = clock64();
    if (threadIdx.x == 0) out[3] = t1 - t0;
    s += u;
    int j = i0 & 1023;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) j = chain[j];
    t1 = clock64();
    if (threadIdx.x == 0) out[4] = t1 - t0;
    s += j;
    {
        const int* gen = chain;  // generic pointer to shared memory
        asm volatile("" : "+l"(gen));
        t0 = clock64();
#pragma unroll 16
        for (int i = 0; i < N; ++i) { int v; asm volatile("ld.b32 %0, [%1];" : "=r"(v) : "l"(gen + j)); j = v; }
        t1 = clock64();
        if (threadIdx.x == 0) out[5] = t1 - t0;
        s += j;
    }
    a = x0 + 3.0;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N / 8; ++i) a = x0 / a + 1.5;
    t1 = clock64();
    if (threadIdx.x == 0) out[6] = (t1 - t0) * 8;
    s += a;
    unsigned hi0 = u;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) hi0 = __umulhi(hi0, 0xCD9E8D57u) ^ 0x9E3779B9u;
    t1 = clock64();
    if (threadIdx.x == 0) out[7] = t1 - t0;
    s += hi0;
    // 4 independent DFMA chains (ILP)
    double c0 = x0, c1 = x0 + 1, c2 = x0 + 2, c3 = x0 + 3;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) { c0 = __fma_rn(c0, b, x0); c1 = __fma_rn(c1, b, x0); c2 = __fma_rn(c2, b, x0); c3 = __fma_rn(c3, b, x0); }
    t1 = clock64();
    if (threadIdx.x == 0) out[8] = t1 - t0;
    s += c0 + c1 + c2 + c3;
    // table lookup + markstein: the div_tab chain
    a = x0 + 7.0;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N / 8; ++i) { int bq = (i & 31) + 1; double y = dsh[bq]; double q = __dmul_rn(a, y); double r = __fma_rn(-q, (double)bq, a); a = __fma_rn(r, y, q) + 3.0; }
    t1 = clock64();
    if (threadIdx.x == 0) out[9] = (t1 - t0) * 8;
    s += a;
    if (s == 1.2345) *gsink = s;
}
int main() {
    long long* d; double* g;
    cudaMalloc(&d, 16 * 8); cudaMalloc(&g, 8);
    const char* names[] = {"DFMA", "DMUL+DADD", "FFMA", "IMAD", "LDS chain", "generic LD (shared)", "f64 div + add", "IMAD.HI+LOP", "4x DFMA ILP (per iter)", "div_tab + add"};
    for (int threads : {32, 128, 512}) {
        k<<<1, threads>>>(d, 1.000001, 5, g);
        k<<<1, threads>>>(d, 1.000001, 5, g);
        long long h[16];
        cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
        printf("threads %d:", threads);
        for (int i = 0; i < 10; ++i) printf("  %s %.1f", names[i], (double)h[i] / N);
        printf("\n");
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
