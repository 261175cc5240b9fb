// qmlp.cuh -- building blocks of the tensor-core leaf evaluation (AZG_FLAG_EVAL_Q8); the kernel is in qmlp2.cuh.
//
// Same role as mlp.cuh (model.predict_V / predict_pi / sample_action of policies.py:154-160, :340-352, :656-669, batched
// over all leaves), different arithmetic for the hidden x hidden layers: every activation row and every weight row is
// scaled by a power of two to a 24-bit integer and split into three balanced signed base-256 digits; the six digit
// products of order <= 4 are int8 x int8 -> int32 matrix products (tcgen05.mma kind::i8, accumulators in TMEM).  Integer
// products are exact and order-independent, so the result is bit-reproducible by the CPU oracle (oracle/azg_oracle.h,
// "AZO_EVAL_Q8 contract") -- which a TF32 / BF16 split could not give -- and its error against an f64 evaluation is
// about 2x that of the FP32 path (tests/test_q8_eval.py).  Layer 0 (K = 3 or 4) and the heads stay FP32 FMA.
//
// A (activation digits) and B (weight digits) live in shared memory in the no-swizzle K-major canonical layout
// [k/16][row][16 B] (descriptor LBO = 2048 B between k-chunks, SBO = 128 B between 8-row groups; verified bit for bit
// against a CPU product by tools/umma_i8_probe.cu); one 128x128x32 kind::i8 instruction takes 64 cycles (measured).
#pragma once
#include "mlp.cuh"

#define QMLP_PLANE 16384              // bytes of one 128 x 128 int8 digit plane
#define QMLP_IDESC ((2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24))  // s8 x s8 -> s32, K-major A and B, N = 128, M = 128


// bounded spin: a protocol error traps (the launch fails with an error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    int spins = 0;
    while (!mbar_try_wait_hint(bar, parity, 20000u)) {
        if (++spins > (1 << 22)) {  // seconds
            printf("mbar_wait timeout: block %d thread %d barrier smem 0x%x parity %u\n", blockIdx.x, threadIdx.x, smem_u32(bar), parity);
            __trap();
        }
    }
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

// Conversions.  The XU pipe (F2I / I2F / FRND, 16 lanes per clock per SM) was 78 % busy in the first version of this kernel
// (profiles/r1c) and an all-FMA/ALU replacement costs twice the issue slots, so the work is split: the activation's rint and
// 2^n (mlp_act2) and the PA / PB conversions avoid the XU pipe, the quantisation and PC use it.
// q8_quant: rni_sat(x); the scaling guarantees |x| < 2^23, NaN quantises to 0 (cvt.rni.s32.f32).
__device__ __forceinline__ uint32_t q8_quant(float x) { return (uint32_t)__float2int_rn(x); }
// exact s32 -> f32 for |p| <= 2^22 (PA, PB): the integer is added into the mantissa of 1.5 * 2^23
__device__ __forceinline__ float q8_i2f_22(int32_t p) { return __fsub_rn(__uint_as_float((uint32_t)p + 0x4B400000u), 12582912.0f); }
__device__ __forceinline__ float q8_i2f_23(int32_t p) { return __int2float_rn(p); }

// quantise 16 consecutive activations of one row and store their three digit planes (16 B each) for k-chunk kc
__device__ __forceinline__ void q8_store16(const float* a, float sx, int8_t* sA, int kc, int r) {
    uint32_t whi[4], wmid[4], wlo[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        uint32_t t[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) t[i] = (q8_quant(__fmul_rn(a[4 * g + i], sx)) + 0x8080u) ^ 0x8080u;
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
