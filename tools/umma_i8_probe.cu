// umma_i8_probe.cu -- bring-up probe for tcgen05.mma kind::i8 (int8 x int8 -> int32 in TMEM) on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_i8_probe tools/umma_i8_probe.cu && tools/umma_i8_probe
// Checks: (1) the no-swizzle K-major shared-memory descriptor ([k/16][row][16 B], SBO = 128 B, LBO = rows*16 B) and the
// instruction descriptor against a CPU product, exactly; (2) MMA issue rate; (3) tcgen05.ld read rate with 4 / 8 / 16 warps.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(
            smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version 1 (Blackwell)
    return d;                // base offset 0, lbo mode 0, layout type 0 = no swizzle
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

constexpr uint32_t IDESC_I8_128x128 = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

// A, B: [8][128][16] int8 (k-chunk, row, 16 k); D: [128][128] int32.  512 threads (16 warps).
__global__ void __launch_bounds__(512, 1) k_probe(const int8_t* A, const int8_t* B, int32_t* D, uint32_t lbo, uint32_t sbo, int reps, long long* cyc,
                                                 int32_t* sink) {
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
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core
    if (tid == 0) mbar_init(&bar, 1);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base_s;
    uint32_t phase = 0;

    // (1) correctness: D = A * B^T, 4 k-steps of 32
    if (tid == 0) {
        for (int k = 0; k < 4; ++k) {
            const uint64_t da = make_desc(smem_u32(sA) + k * 4096, lbo, sbo);
            const uint64_t db = make_desc(smem_u32(sB) + k * 4096, lbo, sbo);
            umma_i8(tb, da, db, IDESC_I8_128x128, k > 0);
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, phase); phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp < 4) {
        for (int c = 0; c < 4; ++c) {
            uint32_t v[32];
            tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + c * 32, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * 128 + c * 32 + j] = (int32_t)v[j];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();

    // (2) MMA rate: reps x (6 products x 4 k-steps) into 3 accumulators, one commit
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        t0 = clock64();
        for (int r = 0; r < reps; ++r)
            for (int pr = 0; pr < 6; ++pr)
                for (int k = 0; k < 4; ++k) {
                    const uint64_t da = make_desc(smem_u32(sA) + k * 4096, lbo, sbo);
                    const uint64_t db = make_desc(smem_u32(sB) + k * 4096, lbo, sbo);
                    umma_i8(tb + (pr % 3) * 128, da, db, IDESC_I8_128x128, 1);
                }
        umma_commit(&bar);
    }
    mbar_wait(&bar, phase); phase ^= 1;
    if (tid == 0) { t1 = clock64(); cyc[0] = t1 - t0; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    __syncthreads();

    // (3) tcgen05.ld rate: nw warps each read their 32 lanes x 384 columns, reps times
    int32_t acc = 0;
    for (int cfg = 0; cfg < 3; ++cfg) {
        const int nw = cfg == 0 ? 4 : (cfg == 1 ? 8 : 16);
        __syncthreads();
        const long long s0 = clock64();
        if (warp < nw) {
            // warps w, w+4, w+8, w+12 share lanes 32*(w%4); split the 384 columns between them
            const int share = nw / 4, part = warp / 4, ncol = 384 / share;
            for (int r = 0; r < reps; ++r) {
                for (int c = 0; c < ncol; c += 32) {
                    uint32_t v[32];
                    tmem_ld32(tb + ((uint32_t)((warp & 3) * 32) << 16) + part * ncol + c, v);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc += (int32_t)v[j];
                }
            }
        }
        __syncthreads();
        if (tid == 0) cyc[1 + cfg] = clock64() - s0;
    }
    sink[tid] = acc;
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
    std::vector<int32_t> ref(128 * 128);
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 128; ++n) {
            int32_t s = 0;
            for (int k = 0; k < 128; ++k) s += (int32_t)a[m * 128 + k] * (int32_t)b[n * 128 + k];
            ref[m * 128 + n] = s;
        }
    int8_t *dA, *dB;
    int32_t *dD, *dsink;
    long long* dcyc;
    CK(cudaMalloc(&dA, 16384)); CK(cudaMalloc(&dB, 16384)); CK(cudaMalloc(&dD, 65536)); CK(cudaMalloc(&dcyc, 64)); CK(cudaMalloc(&dsink, 4096));
    CK(cudaMemcpy(dA, ap.data(), 16384, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, bp.data(), 16384, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    const uint32_t cfgs[2][2] = {{2048, 128}, {128, 2048}};  // {LBO, SBO}: which field is the K-direction stride?
    for (int c = 0; c < 2; ++c) {
        CK(cudaMemset(dD, 0xFF, 65536));
        k_probe<<<1, 512, 65536>>>(dA, dB, dD, cfgs[c][0], cfgs[c][1], 100, dcyc, dsink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("cfg %d: kernel failed: %s\n", c, cudaGetErrorString(e)); return 1; }
        std::vector<int32_t> d(128 * 128);
        long long cyc[4];
        CK(cudaMemcpy(d.data(), dD, 65536, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(cyc, dcyc, 32, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int i = 0; i < 128 * 128; ++i) bad += d[i] != ref[i];
        printf("cfg LBO=%u SBO=%u: mismatches %d / 16384 (d[0]=%d ref[0]=%d d[129]=%d ref[129]=%d)\n", cfgs[c][0], cfgs[c][1], bad, d[0], ref[0], d[129],
               ref[129]);
        printf("  MMA: %lld cycles for 100 x 24 instr = %.1f cycles per 128x128x32 i8 MMA\n", cyc[0], cyc[0] / 2400.0);
        for (int k = 0; k < 3; ++k)
            printf("  LDTM %2d warps: %lld cycles for 100 x (128 lanes x 384 cols x 4 B) = %.1f B/cycle/SM\n", k == 0 ? 4 : (k == 1 ? 8 : 16), cyc[1 + k],
                   100.0 * 128 * 384 * 4 / cyc[1 + k]);
    }
    return 0;
}
