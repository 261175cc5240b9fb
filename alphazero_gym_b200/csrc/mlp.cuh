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
// Layout: one CTA owns a tile of TM = 128 leaves; all weights (138 KB for 3x128 ELU, 71 KB for 2x128 ReLU)
// are staged in shared memory once per CTA and reused for every tile the persistent CTA processes;
// activations live in shared memory k-major ([H][TM]) so both operands of the 8x8 register tile are read
// with conflict-free LDS.128.  Bound: FP32 FMA issue (2*H*H*TM flop per hidden layer per tile).
#pragma once
#include "common.cuh"
#include "detmath.cuh"

#define MLP_TM 128
#define MLP_MAX_PO 28  // 1 + 3*AZG_MAX_K rounded up to a multiple of 4

struct MlpParams {
    const float* wpack;  // packed weights (engine.cu pack_weights): per layer Wt[K][H] k-major + b[H]; heads Wh[H][PO_PAD] + bh[PO_PAD]
    int32_t wcount;      // floats in wpack (multiple of 4)
    int32_t S, L, P, PO_PAD;
    int32_t act;         // AZG_ACT_*
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

__device__ __forceinline__ float mlp_act(int act, float v) {
    if (act == 0) return v > 0.0f ? v : 0.0f;          // ReLU (DiscretePolicy.yaml:8)
    return v > 0.0f ? v : det::expm1f_(v);             // ELU alpha=1 (ContinuousPolicy.yaml:9)
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

template <int H>
__global__ void __launch_bounds__((MLP_TM / 8) * (H / 8), 1) k_mlp(const MlpParams p) {
    constexpr int TM = MLP_TM;
    constexpr int NT = (TM / 8) * (H / 8);
    constexpr int NG = NT / TM;  // thread groups in the row-per-thread phases (H=128: 2, H=64: 1)
    static_assert(NT % TM == 0 && NG >= 1, "thread mapping");
    extern __shared__ __align__(16) float smem[];
    float* w = smem;
    float* actb = smem + p.wcount;        // [H][TM]
    float* outs = actb + H * TM;          // [PO_PAD][TM]
    const int tid = threadIdx.x;

    for (int i = tid * 4; i < p.wcount; i += NT * 4)
        *reinterpret_cast<float4*>(w + i) = __ldg(reinterpret_cast<const float4*>(p.wpack + i));
    __syncthreads();

    const int ntiles = (p.n + TM - 1) / TM;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int row0 = tile * TM;
        const int row = tid % TM, grp = tid / TM;
        const int gr = row0 + row;
        int leafw = 0;
        bool need = gr < p.n;
        if (need && p.mode == 0) {
            leafw = p.leaf[gr];
            need = (leafw & LEAF_EVAL) != 0;
        }
        // ---- layer 0: S -> H, one row per thread, H/NG outputs each
        {
            float x[4] = {0.f, 0.f, 0.f, 0.f};
            if (need)
                for (int s = 0; s < p.S; ++s) x[s] = p.X[(size_t)gr * p.xstride + s];
            const float* W0 = w;
            const float* b0 = w + p.S * H;
            for (int j = grp * (H / NG); j < (grp + 1) * (H / NG); ++j) {
                float acc = b0[j];
                for (int s = 0; s < p.S; ++s) acc = __fmaf_rn(W0[s * H + j], x[s], acc);
                actb[j * TM + row] = mlp_act(p.act, acc);
            }
        }
        __syncthreads();
        // ---- hidden layers: [TM x H] = [TM x H] * [H x H], 8x8 register tile, k outermost
        int off = p.S * H + H;
        for (int l = 1; l < p.L; ++l) {
            const float* Wt = w + off;
            const float* b = Wt + H * H;
            off += H * H + H;
            const int tx = tid % 16, ty = tid / 16;  // rows {tx*4..+3, 64+tx*4..+3}, cols ty*8..+7
            float acc[8][8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float bc = b[ty * 8 + c];
#pragma unroll
                for (int r = 0; r < 8; ++r) acc[c][r] = bc;
            }
            const float* ap = actb + tx * 4;
            const float* wp = Wt + ty * 8;
#pragma unroll 2
            for (int k = 0; k < H; ++k) {
                const float4 a0 = *reinterpret_cast<const float4*>(ap + k * TM);
                const float4 a1 = *reinterpret_cast<const float4*>(ap + k * TM + 64);
                const float4 w0 = *reinterpret_cast<const float4*>(wp + k * H);
                const float4 w1 = *reinterpret_cast<const float4*>(wp + k * H + 4);
                const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int c = 0; c < 8; ++c)
#pragma unroll
                    for (int r = 0; r < 8; ++r) acc[c][r] = __fmaf_rn(ww[c], a[r], acc[c][r]);
            }
            __syncthreads();  // everyone has finished reading this layer's input
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float4 o0, o1;
                o0.x = mlp_act(p.act, acc[c][0]); o0.y = mlp_act(p.act, acc[c][1]);
                o0.z = mlp_act(p.act, acc[c][2]); o0.w = mlp_act(p.act, acc[c][3]);
                o1.x = mlp_act(p.act, acc[c][4]); o1.y = mlp_act(p.act, acc[c][5]);
                o1.z = mlp_act(p.act, acc[c][6]); o1.w = mlp_act(p.act, acc[c][7]);
                float* dst = actb + (ty * 8 + c) * TM + tx * 4;
                *reinterpret_cast<float4*>(dst) = o0;
                *reinterpret_cast<float4*>(dst + 64) = o1;
            }
            __syncthreads();
        }
        // ---- heads: column 0 = value_head, columns 1..P = dist_head; one row per thread, 4 outputs per pass
        {
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
        __syncthreads();  // outs / actb are reused by the next tile
    }
}
