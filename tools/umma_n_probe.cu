// umma_n_probe.cu -- tcgen05.mma kind::i8 with narrow N (16 / 32) and accumulator windows at 16-column offsets (search_wg.cuh).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_n_probe tools/umma_n_probe.cu && tools/umma_n_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(2048u >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr) : "memory");
}
// D[128][N] = A[128][128] * B[n0 .. n0+N)[128]^T written at TMEM column dcol; status: 0 ok, 1 timeout
__global__ void __launch_bounds__(128, 1) k_probe(const int8_t* A, const int8_t* B, int32_t* D, int N, int n0, int dcol, int reps, long long* cyc, int* status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t* sA = smem;
    uint8_t* sB = smem + 16384;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 16384 / 16; i += blockDim.x) {
        reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(A)[i];
        reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(B)[i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid == 0) mbar_init(&bar, 1);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base_s;
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
    long long t0 = clock64();
    if (tid == 0) {
        for (int r = 0; r < reps; ++r)
            for (int k = 0; k < 4; ++k)
                umma_i8(tb + dcol, make_desc(smem_u32(sA) + k * 4096), make_desc(smem_u32(sB) + k * 4096 + n0 * 16), idesc, k > 0);
        umma_commit(&bar);
    }
    int spins = 0;
    bool ok = true;
    while (!mbar_try(&bar, 0)) if (++spins > (1 << 24)) { ok = false; break; }
    if (tid == 0) { cyc[0] = clock64() - t0; *status = ok ? 0 : 1; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (ok) {
        for (int c = 0; c < N; c += 16) {
            int32_t v[16];
            tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + dcol + c, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * 256 + c + j] = v[j];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "n"(512) : "memory");
}
int main() {
    std::vector<int8_t> a(128 * 128), b(128 * 128), ap(16384), bp(16384);
    srand(7);
    for (auto& x : a) x = (int8_t)(rand() % 256 - 128);
    for (auto& x : b) x = (int8_t)(rand() % 256 - 128);
    for (int r = 0; r < 128; ++r)
        for (int k = 0; k < 128; ++k) {
            ap[(k / 16) * 2048 + r * 16 + (k % 16)] = a[r * 128 + k];
            bp[(k / 16) * 2048 + r * 16 + (k % 16)] = b[r * 128 + k];
        }
    int8_t *dA, *dB; int32_t* dD; long long* dcyc; int* dst;
    CK(cudaMalloc(&dA, 16384)); CK(cudaMalloc(&dB, 16384)); CK(cudaMalloc(&dD, 128 * 256 * 4)); CK(cudaMalloc(&dcyc, 64)); CK(cudaMalloc(&dst, 4));
    CK(cudaMemcpy(dA, ap.data(), 16384, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, bp.data(), 16384, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    const int cases[][3] = {{64, 0, 0}, {32, 32, 0}, {32, 32, 32}, {16, 16, 0}, {16, 48, 16}, {16, 32, 48}, {16, 112, 32}, {16, 0, 224 + 16}, {16, 16, 96}, {8, 8, 8}};
    for (auto& c : cases) {
        const int N = c[0], n0 = c[1], dcol = c[2];
        for (int reps : {1, 200}) {
            CK(cudaMemset(dD, 0xFF, 128 * 256 * 4));
            k_probe<<<1, 128, 65536>>>(dA, dB, dD, N, n0, dcol, reps, dcyc, dst);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("N=%d n0=%d dcol=%d: kernel failed: %s\n", N, n0, dcol, cudaGetErrorString(e)); return 1; }
            std::vector<int32_t> d(128 * 256);
            long long cyc; int st;
            CK(cudaMemcpy(d.data(), dD, d.size() * 4, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost));
            int bad = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < N; ++n) {
                    int32_t s = 0;
                    for (int k = 0; k < 128; ++k) s += (int32_t)a[m * 128 + k] * (int32_t)b[(n0 + n) * 128 + k];
                    bad += d[m * 256 + n] != s;
                }
            if (reps == 1) printf("N=%3d n0=%3d dcol=%3d: status %d mismatches %d / %d", N, n0, dcol, st, bad, 128 * N);
            else printf("   | %d x 4 MMAs: %.1f cycles per MMA\n", reps, cyc / (4.0 * reps));
        }
    }
    return 0;
}
