// mlp.cuh -- batched leaf evaluation: trunk MLP + value head + policy head in ONE kernel.
//
// Replaces the batch-1 model.predict_V / predict_pi / sample_action calls the reference makes per node
// (mcts.py:406-416, :619-623, :652 -> policies.py:154-160, :340-352, :656-669).  The trunk runs once per
// leaf (the reference runs it 2x per discrete node and again for every progressive-widening sample) and
// the post-processed policy head is cached with the node.
//
// Arithmetic: FP32 FMA, NOT tensor cores.  BASELINE.json's tolerance (Q/V within 1e-5, bit-exact visit
// counts) rules out TF32/BF16 products, and a fixed summation order -- acc = bias; acc = fma(w[j][k],
// x[k], acc) for k ascending -- makes every output bit-reproducible by the CPU oracle.  A register-tiled
// SGEMM has exactly that per-output order as long as k is the outer loop.
//
// Structure (one persistent CTA per SM, 512 threads for H = 128):
//   * all weights (138 KB for 3x128 ELU, 71 KB for 2x128 ReLU) are staged in shared memory once per CTA
//     with TMA bulk copies (cp.async.bulk + mbarrier) and reused for every pass of the CTA;
//   * rows are dealt to CTAs in units of 64 and processed in passes of 128 rows plus at most one 64-row
//     pass (half of the warps), so 65536 rows on 148 SMs cost 3.5 pass-times instead of the 4 that fixed
//     128-row tiles would;
//   * activations live in shared memory k-major ([H][128]); both operands of the register tile are read
//     with conflict-free LDS.128 and the K loop is software-pipelined (operands of k+1 are in flight while
//     the FMAs of k issue);
//   * activations are branch-free (ELU through the deterministic expm1 of detmath.cuh).
// Bound: FP32 FMA issue (2*H*H flop per row per hidden layer).
#pragma once
#include "common.cuh"
#include "detmath.cuh"

#define MLP_TM 128
#define MLP_UNIT 64
#define MLP_MAX_PO 28  // 1 + 3*AZG_MAX_K rounded up to a multiple of 4

struct MlpParams {
    const float* wpack;  // packed weights (engine.cu pack_weights): per layer Wt[K][H] k-major + b[H]; heads Wh[H][PO_PAD] + bh[PO_PAD]
    int32_t wcount;      // floats in wpack (multiple of 4)
    int32_t L, P, PO_PAD;
    int32_t n;           // rows to evaluate
    const float* X;      // [n][xstride]
    int32_t xstride;
    int32_t mode;        // 0: scatter into the tree tables of leaf[] rows; 1: dense outputs
    int32_t variant, A, K, R, HS;
    float ls_min, ls_max;
    const int32_t* leaf;
    DRow* drows;
    CRow* crows;
    float* chead;
    uint32_t* evals;     // per-tree evaluation counter
    double* leafR;       // continuous: leafR[t] += gamma_f32 * V  (first backup step, mcts.py:260-263)
    float gamma_f32;
    float* outV;
    float* outHead;
    int32_t head_dim;
};

// ---- TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP / SYNCS) --------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// ---- activations (branch-free) -------------------------------------------------------------------------
template <int ACT>
__device__ __forceinline__ float mlp_act(float v) {
    if (ACT == 0) return v > 0.0f ? v : 0.0f;  // ReLU (DiscretePolicy.yaml:8)
    // ELU alpha=1 (ContinuousPolicy.yaml:9): v > 0 ? v : expm1(v), expm1 = det::expm1f_ written with selects.
    // (n == 0 needs no special case: fma(p, 1, 0) == p; x < -17.5 is selected to -1 like the reference form.)
    const float xx = fmaxf(v, -20.0f);
    float n;
    const float r = det::exp_reduce(xx, n);
    const float pl = det::expm1_poly(r);
    const float t = __uint_as_float((uint32_t)((int32_t)n + 127) << 23);
    const float e = __fmaf_rn(pl, t, __fsub_rn(t, 1.0f));
    float res = v < -17.5f ? -1.0f : e;
    res = v > 0.0f ? v : res;
    return (v != v) ? v : res;
}

// post-processing of one row's raw head outputs (policies.py:275-297 softmax priors; :617-631 GMM params)
__device__ __forceinline__ void softmax_seq(const float* l, int n, float* p) {
    float m = l[0];
    for (int i = 1; i < n; ++i) m = l[i] > m ? l[i] : m;
    float s = 0.0f;
    for (int i = 0; i < n; ++i) {
        p[i] = det::expf_(__fsub_rn(l[i], m));
        s = __fadd_rn(s, p[i]);
    }
    for (int i = 0; i < n; ++i) p[i] = __fdiv_rn(p[i], s);
}

// Thread mapping of the hidden layers: a pass is 128 rows x H columns; each thread owns 4 rows x 8 columns,
// a warp owns a 32 x 32 square (8 row groups x 4 column groups), the CTA has 4 x (H/32) warps = 512 threads
// for H = 128, i.e. 4 warps per SM sub-partition to cover LDS latency and barrier skew (the first version
// ran 8x8 tiles on 256 threads = 2 warps per scheduler and idled 40 % of its issue slots).  Per k a warp
// reads 128 B of activations (LDS.128, 8 distinct lanes, rest broadcast) and 2 x 64 B of weights.
// The accumulators are float2 column pairs updated with the packed fma.rn.f32x2 (SASS FFMA2, scalar operand
// broadcast): half the issue slots of scalar FFMA for the same IEEE result per element, which leaves slots
// for the LDS / activation instructions.  k is the outer loop, so every output still sums in k order.
#define MLP_THREADS(H) (128 * ((H) / 32))

template <int H, int ACT, bool HALF>
__device__ __forceinline__ void hidden_layer(const float* __restrict__ Wt, const float* __restrict__ b, float* actb, int tid) {
    constexpr int TM = MLP_TM;
    const int lane = tid & 31, wid = tid >> 5;
    const int wr = wid / (H / 32), wc = wid % (H / 32);  // warp row block (32 rows) / column block (32 cols)
    const int row0 = wr * 32 + (lane & 7) * 4, col0 = wc * 32 + (lane >> 3) * 8;
    const bool active = !HALF || wr < 2;  // a 64-row pass uses warp rows 0 and 1 (warps 0..2*H/32-1: all 4 schedulers)
    float2 acc[4][4];                     // [row][column pair]
    if (active) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float2 bc = *reinterpret_cast<const float2*>(b + col0 + 2 * c);
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r][c] = bc;
        }
        const float* ap = actb + row0;
        const float* wp = Wt + col0;
        float4 a = *reinterpret_cast<const float4*>(ap);
        float4 w0 = *reinterpret_cast<const float4*>(wp);
        float4 w1 = *reinterpret_cast<const float4*>(wp + 4);
#pragma unroll 8
        for (int k = 0; k < H; ++k) {
            const int kn = k + 1 < H ? k + 1 : k;  // the last iteration re-reads row k (discarded)
            const float4 na = *reinterpret_cast<const float4*>(ap + kn * TM);
            const float4 nw0 = *reinterpret_cast<const float4*>(wp + kn * H);
            const float4 nw1 = *reinterpret_cast<const float4*>(wp + kn * H + 4);
            const float ar[4] = {a.x, a.y, a.z, a.w};
            const float2 wpair[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y), make_float2(w1.z, w1.w)};
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float2 av = make_float2(ar[r], ar[r]);
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = __ffma2_rn(av, wpair[c], acc[r][c]);
            }
            a = na; w0 = nw0; w1 = nw1;
        }
    }
    __syncthreads();  // everyone has finished reading this layer's input
    if (active) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float* d0 = actb + (col0 + 2 * c) * TM + row0;
            *reinterpret_cast<float4*>(d0) = make_float4(mlp_act<ACT>(acc[0][c].x), mlp_act<ACT>(acc[1][c].x),
                                                         mlp_act<ACT>(acc[2][c].x), mlp_act<ACT>(acc[3][c].x));
            *reinterpret_cast<float4*>(d0 + TM) = make_float4(mlp_act<ACT>(acc[0][c].y), mlp_act<ACT>(acc[1][c].y),
                                                              mlp_act<ACT>(acc[2][c].y), mlp_act<ACT>(acc[3][c].y));
        }
    }
    __syncthreads();
}

template <int H, int S, int ACT, bool HALF>
__device__ __forceinline__ void mlp_pass(const MlpParams& p, const float* w, float* actb, float* outs, int row0, int tid) {
    constexpr int TM = MLP_TM;
    constexpr int NT = MLP_THREADS(H);
    constexpr int NG = NT / TM;  // thread groups in the row-per-thread phases (H=128: 4, H=64: 2)
    constexpr int ROWS = HALF ? 64 : 128;
    const int row = tid % TM, grp = tid / TM;
    const int gr = row0 + row;
    int leafw = 0;
    bool need = row < ROWS && gr < p.n;
    if (need && p.mode == 0) {
        leafw = p.leaf[gr];
        need = (leafw & LEAF_EVAL) != 0;
    }
    // ---- layer 0: S -> H, one row per thread, H/NG outputs each
    if (row < ROWS) {
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
        const float* W0 = w;
        const float* b0 = w + S * H;
#pragma unroll 4
        for (int j = grp * (H / NG); j < (grp + 1) * (H / NG); ++j) {
            float acc = b0[j];
#pragma unroll
            for (int s = 0; s < S; ++s) acc = __fmaf_rn(W0[s * H + j], x[s], acc);
            actb[j * TM + row] = mlp_act<ACT>(acc);
        }
    }
    __syncthreads();
    // ---- hidden layers
    int off = S * H + H;
    for (int l = 1; l < p.L; ++l) {
        hidden_layer<H, ACT, HALF>(w + off, w + off + H * H, actb, tid);
        off += H * H + H;
    }
    // ---- heads: column 0 = value_head, columns 1..P = dist_head; one row per thread, 4 outputs per pass
    if (row < ROWS) {
        const float* Wh = w + off;
        const float* bh = Wh + H * p.PO_PAD;
        for (int c = grp; c < p.PO_PAD / 4; c += NG) {
            float4 acc = *reinterpret_cast<const float4*>(bh + c * 4);
#pragma unroll 8
            for (int k = 0; k < H; ++k) {
                const float a = actb[k * TM + row];
                const float4 w4 = *reinterpret_cast<const float4*>(Wh + k * p.PO_PAD + c * 4);
                acc.x = __fmaf_rn(w4.x, a, acc.x);
                acc.y = __fmaf_rn(w4.y, a, acc.y);
                acc.z = __fmaf_rn(w4.z, a, acc.z);
                acc.w = __fmaf_rn(w4.w, a, acc.w);
            }
            outs[(c * 4 + 0) * TM + row] = acc.x;
            outs[(c * 4 + 1) * TM + row] = acc.y;
            outs[(c * 4 + 2) * TM + row] = acc.z;
            outs[(c * 4 + 3) * TM + row] = acc.w;
        }
    }
    __syncthreads();
    // ---- post-processing + write-back, one row per thread (threads of group 0)
    if (grp == 0 && need) {
        float V = outs[row];
        float raw[3 * AZG_MAX_K], post[3 * AZG_MAX_K];
        for (int i = 0; i < p.P; ++i) raw[i] = outs[(1 + i) * TM + row];
        int npost;
        if (p.variant == 0) {
            softmax_seq(raw, p.A, post);
            npost = p.A;
        } else {
            const int K = p.K;
            for (int k = 0; k < K; ++k) {
                post[k] = raw[k];
                float ls = raw[K + k];
                ls = ls < p.ls_min ? p.ls_min : ls;
                ls = ls > p.ls_max ? p.ls_max : ls;
                post[K + k] = det::expf_(ls);
            }
            if (K > 1) softmax_seq(raw + 2 * K, K, post + 2 * K);
            else post[2] = 1.0f;
            npost = 3 * K;
        }
        if (p.mode == 1) {
            p.outV[gr] = V;
            for (int i = 0; i < npost; ++i) p.outHead[(size_t)gr * p.head_dim + i] = post[i];
        } else {
            const size_t ri = (size_t)gr * p.R + (leafw & LEAF_ROW_MASK);
            if (leafw & LEAF_TERMINAL) V = 0.0f;  // mcts.py:406-410, :619-623
            if (p.variant == 0) {
                DRow* d = p.drows + ri;
                d->V = V;
                *reinterpret_cast<float2*>(d->prior) = make_float2(post[0], post[1]);
            } else {
                p.crows[ri].V = V;
                p.leafR[gr] = p.leafR[gr] + (double)__fmul_rn(p.gamma_f32, V);
                float* h = p.chead + ri * p.HS;
                for (int i = 0; i < npost; ++i) h[i] = post[i];
            }
            p.evals[gr] += 1;
        }
    }
    __syncthreads();  // outs / actb are reused by the next pass
}

template <int H, int S, int ACT>
__global__ void __launch_bounds__(MLP_THREADS(H), 1) k_mlp(const MlpParams p) {
    constexpr int TM = MLP_TM;
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t wbar;
    float* w = smem;
    float* actb = smem + p.wcount;  // [H][TM]
    float* outs = actb + H * TM;    // [PO_PAD][TM]
    const int tid = threadIdx.x;

    // rows are dealt in units of 64: CTA b owns units [b*u, (b+1)*u)
    const int units = (p.n + MLP_UNIT - 1) / MLP_UNIT;
    const int per = (units + gridDim.x - 1) / gridDim.x;
    int unit = blockIdx.x * per;
    const int unit_end = min(unit + per, units);
    if (unit >= unit_end) return;

    // stage all weights with TMA bulk copies; every thread then waits on the mbarrier
    if (tid == 0) mbar_init(&wbar, 1);
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)p.wcount * 4u;
        mbar_expect_tx(&wbar, bytes);
        for (uint32_t o = 0; o < bytes; o += 32768u)
            bulk_g2s(reinterpret_cast<char*>(w) + o, reinterpret_cast<const char*>(p.wpack) + o, min(32768u, bytes - o), &wbar);
    }
    while (!mbar_try_wait(&wbar, 0)) {}

    while (unit < unit_end) {
        if (unit_end - unit >= 2) {
            mlp_pass<H, S, ACT, false>(p, w, actb, outs, unit * MLP_UNIT, tid);
            unit += 2;
        } else {
            mlp_pass<H, S, ACT, true>(p, w, actb, outs, unit * MLP_UNIT, tid);
            unit += 1;
        }
    }
}
