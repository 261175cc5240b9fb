// qmlp.cuh -- batched leaf evaluation with the H x H layers on the 5th-generation tensor cores (AZG_FLAG_EVAL_Q8).
//
// Same role as mlp.cuh (model.predict_V / predict_pi / sample_action of policies.py:154-160, :340-352, :656-669, batched
// over all leaves), different arithmetic for the hidden x hidden layers: every activation row and every weight row is
// scaled by a power of two to a 24-bit integer and split into three balanced signed base-256 digits; the six digit
// products of order <= 4 are int8 x int8 -> int32 matrix products (tcgen05.mma kind::i8, accumulators in TMEM).  Integer
// products are exact and order-independent, so the result is bit-reproducible by the CPU oracle (oracle/azg_oracle.h,
// "AZO_EVAL_Q8 contract") -- which a TF32 / BF16 split could not give -- and its error against an f64 evaluation is
// about 2x that of the FP32 path (tests/test_q8_eval.py).  Layer 0 (K = 3 or 4) and the heads stay FP32 FMA.
//
// One persistent CTA per SM, 512 threads, one tile of 128 rows at a time:
//   * thread (row r = 32*(warp%4) + lane, column quarter cq = warp/4) owns 32 columns of row r in every phase: TMEM lane r
//     is readable by warps with warp%4 == r/32 only, so the four threads of a row sit in four different warps;
//   * A (activation digits) and B (weight digits) live in shared memory in the no-swizzle K-major canonical layout
//     [k/16][row][16 B] (descriptor LBO = 2048 B between k-chunks, SBO = 128 B between 8-row groups; verified bit for bit
//     against a CPU product by tools/umma_i8_probe.cu); B is staged once per CTA with TMA bulk copies;
//   * per hidden layer: row maximum (4 partials through shared memory) -> quantise own 32 columns -> 6 x STS.128 ->
//     fence.proxy.async -> one thread issues 24 MMAs (6 products x 4 k-steps of 32) into three 128-column accumulators
//     (PA at columns 0..127, PB at 128..255, PC at 256..383) -> tcgen05.commit -> every thread reads its 3 x 32 accumulator
//     words with tcgen05.ld, rebuilds y, applies bias + activation.
// Bound: the epilogue (dequantise, activation, quantise) on the FMA/ALU pipes; the 24 MMAs of a tile-layer take ~1540
// cycles (64 cycles per 128x128x32 instruction, measured).
#pragma once
#include "mlp.cuh"

#define QMLP_THREADS 512
#define QMLP_PLANE 16384              // bytes of one 128 x 128 int8 digit plane
#define QMLP_IDESC ((2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24))  // s8 x s8 -> s32, K-major A and B, N = 128, M = 128

__host__ __device__ inline size_t qmlp_smem_bytes(int NL, int qfl_count, int PO_PAD) {
    return (size_t)NL * 3 * QMLP_PLANE + 3 * QMLP_PLANE + (size_t)qfl_count * 4 + 4 * 128 * 4 + (size_t)4 * PO_PAD * 128 * 4 + 1024;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {  // no swizzle, K-major, LBO 2048 B, SBO 128 B, descriptor version 1
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(2048u >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(QMLP_IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// scaling exponent of the Q8 contract: biased exponent of vmax * 1.004f, clamped to [32, 200]
__device__ __forceinline__ int q8_exponent(float vmax) {
    int e = (int)((__float_as_uint(__fmul_rn(vmax, 1.004f)) >> 23) & 0xFF);
    e = e < 32 ? 32 : e;
    return e > 200 ? 200 : e;
}

// quantise 16 consecutive activations of one row and store their three digit planes (16 B each) for k-chunk kc
__device__ __forceinline__ void q8_store16(const float* a, float sx, int8_t* sA, int kc, int r) {
    uint32_t whi[4], wmid[4], wlo[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        uint32_t t[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) t[i] = ((uint32_t)__float2int_rn(__fmul_rn(a[4 * g + i], sx)) + 0x8080u) ^ 0x8080u;
        const uint32_t lo01 = __byte_perm(t[0], t[1], 0x5140), lo23 = __byte_perm(t[2], t[3], 0x5140);
        const uint32_t hi01 = __byte_perm(t[0], t[1], 0x0062), hi23 = __byte_perm(t[2], t[3], 0x0062);
        wlo[g] = __byte_perm(lo01, lo23, 0x5410);
        wmid[g] = __byte_perm(lo01, lo23, 0x7632);
        whi[g] = __byte_perm(hi01, hi23, 0x5410);
    }
    int8_t* d = sA + kc * 2048 + r * 16;
    *reinterpret_cast<uint4*>(d) = make_uint4(whi[0], whi[1], whi[2], whi[3]);
    *reinterpret_cast<uint4*>(d + QMLP_PLANE) = make_uint4(wmid[0], wmid[1], wmid[2], wmid[3]);
    *reinterpret_cast<uint4*>(d + 2 * QMLP_PLANE) = make_uint4(wlo[0], wlo[1], wlo[2], wlo[3]);
}

template <int S, int ACT, int NL>
__global__ void __launch_bounds__(QMLP_THREADS, 1) k_qmlp(const MlpParams p) {
    constexpr int H = 128;
    extern __shared__ __align__(1024) uint8_t qsm_raw[];
    __shared__ __align__(8) uint64_t wbar, mbar;
    __shared__ uint32_t tmem_base_s;
    uint8_t* qsm = reinterpret_cast<uint8_t*>(((uintptr_t)qsm_raw + 1023) & ~(uintptr_t)1023);
    int8_t* sB = reinterpret_cast<int8_t*>(qsm);                 // [NL][3][8][128][16]
    int8_t* sA = sB + (size_t)NL * 3 * QMLP_PLANE;               // [3][8][128][16]
    float* fl = reinterpret_cast<float*>(sA + 3 * QMLP_PLANE);   // W0t[S][H], b0[H], NL x (cw, bias)[H], Wh[H][PO_PAD], bh[PO_PAD]
    float* pmax = fl + p.qfl_count;                              // [4][128]
    float* hsum = pmax + 4 * 128;                                // [4][PO_PAD][128]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int lg = warp & 3, cq = warp >> 2;
    const int r = lg * 32 + lane;

    // this CTA's contiguous row range; tiles of 128 rows, the last one partial
    const int per = (p.n + gridDim.x - 1) / gridDim.x;
    const int row_begin = blockIdx.x * per;
    const int row_end = min(row_begin + per, p.n);
    if (row_begin >= row_end) return;

    if (tid == 0) {
        mbar_init(&wbar, 1);
        mbar_init(&mbar, 1);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_base_s;
    if (tid == 0) {
        const uint32_t bytesB = (uint32_t)NL * 3 * QMLP_PLANE, bytesF = (uint32_t)p.qfl_count * 4u;
        mbar_expect_tx(&wbar, bytesB + bytesF);
        for (uint32_t o = 0; o < bytesB; o += 32768u) bulk_g2s(sB + o, p.qdigits + o, min(32768u, bytesB - o), &wbar);
        bulk_g2s(fl, p.qfl, bytesF, &wbar);
    }
    mbar_wait(&wbar, 0);

    const float* W0 = fl;
    const float* b0 = fl + S * H;
    const float2* cwb = reinterpret_cast<const float2*>(fl + S * H + H);  // [NL][H] (cw[j], bias[j])
    const float* Wh = fl + S * H + H + NL * 2 * H;
    const float* bh = Wh + H * p.PO_PAD;
    uint32_t mphase = 0;

    for (int tile0 = row_begin; tile0 < row_end; tile0 += 128) {
        const int nvalid = min(128, row_end - tile0);
        const bool wactive = lg * 32 < nvalid;  // warp-uniform: this warp's 32 rows contain at least one real row
        const int gr = tile0 + r;
        bool need = r < nvalid;
        int leafw = 0;
        double lr = 0.0;
        if (need && p.mode == 0) {
            if (p.variant == 1) {
                const uint4* cp = reinterpret_cast<const uint4*>(p.ctl + gr);
                const uint4 c0 = cp[0], c1 = cp[1];
                leafw = (int)c0.z;
                lr = __hiloint2double((int)c1.w, (int)c1.z);
            } else {
                leafw = p.leaf[gr];
            }
            need = (leafw & LEAF_EVAL) != 0;
        }
        float a[32];
        float pm = 0.0f;
        if (wactive) {
            // ---- layer 0 (S -> H) in FP32 FMA, own 32 columns
            float x[S];
#pragma unroll
            for (int s = 0; s < S; ++s) x[s] = 0.0f;
            if (need) {
                if (p.xstride == 4) {
                    const float4 v = *reinterpret_cast<const float4*>(p.X + (size_t)gr * 4);
                    x[0] = v.x; x[1] = v.y; x[2] = v.z;
                    if (S > 3) x[S - 1] = v.w;
                } else {
#pragma unroll
                    for (int s = 0; s < S; ++s) x[s] = p.X[(size_t)gr * p.xstride + s];
                }
            }
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                const int j = cq * 32 + i;
                float2 acc = *reinterpret_cast<const float2*>(b0 + j);
#pragma unroll
                for (int s = 0; s < S; ++s) acc = __ffma2_rn(make_float2(x[s], x[s]), *reinterpret_cast<const float2*>(W0 + s * H + j), acc);
                const float2 e = mlp_act2<ACT>(acc);
                a[i] = e.x; a[i + 1] = e.y;
                pm = fmaxf(pm, fmaxf(fabsf(e.x), fabsf(e.y)));
            }
        }
#pragma unroll 1
        for (int l = 0; l < NL; ++l) {
            // ---- row scale, quantise, store A
            pmax[cq * 128 + r] = pm;
            __syncthreads();
            const float m = fmaxf(fmaxf(pmax[r], pmax[128 + r]), fmaxf(pmax[256 + r], pmax[384 + r]));
            const int e = q8_exponent(m);
            const float sx = __uint_as_float((uint32_t)(276 - e) << 23);  // 2^(149-e)
            const float cx = __uint_as_float((uint32_t)(e - 22) << 23);   // 2^(e-149)
            if (wactive) {
                q8_store16(a, sx, sA, cq * 2, r);
                q8_store16(a + 16, sx, sA, cq * 2 + 1, r);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes of A -> visible to the tensor core
            tc_fence_before();
            __syncthreads();
            // ---- 6 digit products x 4 k-steps, one issuing thread
            if (tid == 0) {
                tc_fence_after();
                const uint32_t aH = smem_u32(sA), aM = aH + QMLP_PLANE, aL = aH + 2 * QMLP_PLANE;
                const uint32_t bH = smem_u32(sB) + (uint32_t)l * 3 * QMLP_PLANE, bM = bH + QMLP_PLANE, bL = bH + 2 * QMLP_PLANE;
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_i8(tb, umma_desc(aH + k * 4096), umma_desc(bH + k * 4096), k > 0);              // PA = xh*wh
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_i8(tb + 128, umma_desc(aH + k * 4096), umma_desc(bM + k * 4096), k > 0);        // PB = xh*wm
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_i8(tb + 128, umma_desc(aM + k * 4096), umma_desc(bH + k * 4096), 1);            //    + xm*wh
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_i8(tb + 256, umma_desc(aH + k * 4096), umma_desc(bL + k * 4096), k > 0);        // PC = xh*wl
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_i8(tb + 256, umma_desc(aM + k * 4096), umma_desc(bM + k * 4096), 1);            //    + xm*wm
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_i8(tb + 256, umma_desc(aL + k * 4096), umma_desc(bH + k * 4096), 1);            //    + xl*wh
                umma_commit(&mbar);
            }
            mbar_wait(&mbar, mphase);
            mphase ^= 1;
            tc_fence_after();
            // ---- accumulators -> y -> activation (own 32 columns)
            pm = 0.0f;
            if (wactive) {
                const float2* cb = cwb + l * H + cq * 32;
                const uint32_t ta = tb + ((uint32_t)(lg * 32) << 16) + cq * 32;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    int32_t pa[16], pb[16], pc[16];
                    tmem_ld16(ta + c * 16, pa);
                    tmem_ld16(ta + 128 + c * 16, pb);
                    tmem_ld16(ta + 256 + c * 16, pc);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        const float4 sb = *reinterpret_cast<const float4*>(cb + c * 16 + i);  // (cw, bias) of columns i, i+1
                        float u0 = __fmaf_rn(__int2float_rn(pa[i]), 256.0f, __int2float_rn(pb[i]));
                        float u1 = __fmaf_rn(__int2float_rn(pa[i + 1]), 256.0f, __int2float_rn(pb[i + 1]));
                        u0 = __fmaf_rn(u0, 256.0f, __int2float_rn(pc[i]));
                        u1 = __fmaf_rn(u1, 256.0f, __int2float_rn(pc[i + 1]));
                        const float2 y = make_float2(__fmaf_rn(u0, __fmul_rn(cx, sb.x), sb.y), __fmaf_rn(u1, __fmul_rn(cx, sb.z), sb.w));
                        const float2 ev = mlp_act2<ACT>(y);
                        a[c * 16 + i] = ev.x; a[c * 16 + i + 1] = ev.y;
                        pm = fmaxf(pm, fmaxf(fabsf(ev.x), fabsf(ev.y)));
                    }
                }
            }
            tc_fence_before();  // order these TMEM reads before the next layer's MMAs (issued after two more barriers)
        }
        // ---- heads: partial FMA chains over own 32 columns (column 0 = value_head, 1..P = dist_head)
        if (wactive) {
            for (int c4 = 0; c4 < p.PO_PAD / 4; ++c4) {
                float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float4 w4 = *reinterpret_cast<const float4*>(Wh + (cq * 32 + i) * p.PO_PAD + c4 * 4);
                    acc.x = __fmaf_rn(a[i], w4.x, acc.x);
                    acc.y = __fmaf_rn(a[i], w4.y, acc.y);
                    acc.z = __fmaf_rn(a[i], w4.z, acc.z);
                    acc.w = __fmaf_rn(a[i], w4.w, acc.w);
                }
                float* hs = hsum + (cq * p.PO_PAD + c4 * 4) * 128 + r;
                hs[0] = acc.x; hs[128] = acc.y; hs[256] = acc.z; hs[384] = acc.w;
            }
        }
        __syncthreads();
        if (cq == 0 && need) {
            float out[1 + 3 * AZG_MAX_K];
            const int stride = p.PO_PAD * 128;
            for (int c = 0; c <= p.P; ++c) {
                const float* hs = hsum + c * 128 + r;
                out[c] = __fadd_rn(__fadd_rn(__fadd_rn(hs[0], hs[stride]), __fadd_rn(hs[2 * stride], hs[3 * stride])), bh[c]);
            }
            mlp_finish_row(p, gr, leafw, lr, out[0], out + 1);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "n"(512) : "memory");
}
