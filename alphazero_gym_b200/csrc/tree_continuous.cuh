// tree_continuous.cuh -- backup + progressive-widening / UCT select + expansion for the continuous tree.
//
// Reference: MCTSContinuous.search (mcts.py:656-702), selectionUCT (:704-741), add_pw_action (:625-654),
// NodeContinuous.check_pw (states.py:252-275), epsilon_greedy (mcts.py:175-195), helpers.argmax
// (helpers.py:30-52), MCTS.backprop (mcts.py:241-267), DiagonalGMMPolicy.sample_action
// (policies.py:656-669) with SquashedNormal (distributions.py:50-63, :205-245).
//
// Mapping: a sub-warp of L lanes (8/16/32, picked on the host from the progressive-widening fan-out and
// the row count) owns one tree; a CTA owns 32 trees.  Each simulation step is three phases:
//   A (dense, thread i <-> tree i of the CTA): action noise for this step's possible widening insert
//     (f64 Box-Muller) -- at most ONE insert can happen per simulation, because an inserted edge is
//     unexpanded and ends the descent;
//   B (sub-warp): [backup of the previous simulation along its recorded path] then descent from the root.
//     Children are found by scanning the tree's byte-per-row parent array, which the sub-warp holds in
//     registers after ONE coalesced load; their 32 B rows are gathered one per lane, UCT is evaluated in
//     f64 and reduced with warp shuffles (max, ballot of winners, tie-break draw);
//   C (dense): env step + node creation for the selected edge, network input for the evaluation kernel.
// Phases A and C keep f64 transcendental work on full warps instead of one lane per sub-warp.
#pragma once
#include "common.cuh"
#include "env.cuh"

#define TREES_PER_CTA 32

__device__ __forceinline__ CRow load_crow(const CRow* p) {
    CRow r;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&r);
    d[0] = s[0]; d[1] = s[1];
    return r;
}
__device__ __forceinline__ void store_crow(CRow* p, const CRow& r) {
    const uint4* s = reinterpret_cast<const uint4*>(&r);
    uint4* d = reinterpret_cast<uint4*>(p);
    d[0] = s[0]; d[1] = s[1];
}

// stream 1: noise of the j-th widening insert of `tree`: component uniform + K standard normals
__device__ __forceinline__ void pw_noise(const TreeParams& p, int64_t tree, int j, float& u, float* z) {
    u = u32_to_unit(rng_block(p.seed, tree, 1, j, 0).x);
    for (int b = 0; 2 * b < p.K; ++b) {
        const u32x4 c = rng_block(p.seed, tree, 1, j, 1 + b);
        const double u1 = ((double)c.x + 0.5) * 2.3283064365386963e-10;
        const double u2 = ((double)c.y + 0.5) * 2.3283064365386963e-10;
        const double rad = sqrt(-2.0 * det::log_(u1));
        double sn, cs;
        det::sincos_(6.283185307179586 * u2, sn, cs);
        z[2 * b] = (float)(rad * cs);
        if (2 * b + 1 < p.K) z[2 * b + 1] = (float)(rad * sn);
    }
}

// MixtureSameFamily.sample with injected noise: component by inverse CDF, x = z*sigma + mu, a = bound*tanh(x)
__device__ __forceinline__ float sample_action(const TreeParams& p, const float* head, float u, const float* z) {
    const int K = p.K;
    int k = 0;
    if (K > 1) {
        float cum = head[2 * K];
        while (k < K - 1 && !(u < cum)) {
            ++k;
            cum = __fadd_rn(cum, head[2 * K + k]);
        }
    }
    const float x = __fadd_rn(__fmul_rn(z[k], head[K + k]), head[k]);
    return p.action_bound > 0.0f ? __fmul_rn(p.action_bound, det::tanhf_(x)) : x;
}

// MCTSContinuous.initialize_search (mcts.py:589-600): new root row, network input = obs(root)
__global__ void k_init_continuous(const TreeParams p) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.B) return;
    const double th = p.root_state[(size_t)t * 2], thdot = p.root_state[(size_t)t * 2 + 1];
    CRow row;
    row.W = 0.0; row.r = 0.0; row.V = 0.0f; row.action = 0.0f; row.n_e = 0;
    row.nn_flags = CROW_EXPANDED;
    if (p.use_tape) row.V = p.tapeV[(size_t)t * p.R];
    store_crow(p.crows + (size_t)t * p.R, row);
    p.cstate[(size_t)t * p.R] = make_double2(th, thdot);
    p.X[t] = env::pendulum_obs(th, thdot);
    p.leaf[t] = 0 | LEAF_EVAL;
    p.n_rows[t] = 1;
    p.draws[t] = 0;
    p.pw[t] = 0;
    p.depth[t] = 0;
    for (int k = 0; k < 4; ++k) p.ctr[(size_t)k * p.B + t] = 0;
}

// the add_pw_action(root) that precedes the rollout loop (mcts.py:673); runs after the root evaluation
__global__ void k_root_insert_continuous(const TreeParams p) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.B) return;
    float a;
    if (p.use_tape) {
        a = p.tapeA[(size_t)t * p.R + 1];
    } else {
        float u, z[AZG_MAX_K];
        pw_noise(p, p.tree_id0 + t, 0, u, z);
        a = sample_action(p, p.chead + (size_t)t * p.R * p.HS, u, z);
    }
    CRow row;
    row.W = 0.0; row.r = 0.0; row.V = 0.0f; row.action = a; row.n_e = 0; row.nn_flags = 0;
    store_crow(p.crows + (size_t)t * p.R + 1, row);
    p.cparent[(size_t)t * p.PSTRIDE + 1] = 0;
    p.n_rows[t] = 2;
    p.pw[t] = 1;
}

struct ExpandReq {  // phase B -> phase C hand-off (shared memory)
    int cur, sel;
    float action;
    int expand;  // 1: create the node behind `sel`; 0: the trace ended on an existing terminal node `cur`
};

template <int L, int PW, bool BACKUP, bool SELECT>
__global__ void __launch_bounds__(TREES_PER_CTA* L) k_step_continuous(const TreeParams p) {
    constexpr int ROWS_PER_LANE = 4 * PW;
    __shared__ float s_u[TREES_PER_CTA];
    __shared__ float s_z[TREES_PER_CTA][AZG_MAX_K];
    __shared__ uint8_t s_kid[TREES_PER_CTA][L];
    __shared__ ExpandReq s_req[TREES_PER_CTA];

    const int tid = threadIdx.x;
    const int sub = tid / L, lane = tid % L;
    const int t0 = blockIdx.x * TREES_PER_CTA;

    // ---- phase A: noise for this step's insert, one thread per tree -------------------------------
    if (SELECT && !p.use_tape && tid < TREES_PER_CTA && t0 + tid < p.B)
        pw_noise(p, p.tree_id0 + t0 + tid, p.pw[t0 + tid], s_u[tid], s_z[tid]);
    if (SELECT) {
        if (tid < TREES_PER_CTA) s_req[tid].expand = -1;
        __syncthreads();
    }

    // ---- phase B: one sub-warp per tree ---------------------------------------------------------------
    const int t = t0 + sub;
    if (t < p.B) {
        const unsigned smask = (L == 32) ? 0xFFFFFFFFu : (((1u << L) - 1u) << ((tid & 31) / L * L));
        CRow* rows = p.crows + (size_t)t * p.R;
        uint8_t* par = p.cparent + (size_t)t * p.PSTRIDE;
        uint8_t* path = p.path + (size_t)t * p.R;

        if (BACKUP) {
            // R = leaf.V; R = node.r + gamma*R up the recorded path; edge.n += 1; edge.W += R; parent.n += 1.
            // The first product is gamma(f32) * V(f32) as numpy (NEP 50) computes it; the rest is f64.
            const int d = p.depth[t];
            double Rv = 0.0;
            for (int base = ((d - 1) / L) * L; base >= 0; base -= L) {
                const int i = base + lane;
                CRow row;
                int ri = 0;
                const bool mine = i < d;
                if (mine) {
                    ri = path[i];
                    row = load_crow(rows + ri);
                }
                const int cnt = min(L, d - base);
                double myR = 0.0;
                for (int j = cnt - 1; j >= 0; --j) {
                    const double rj = __shfl_sync(smask, mine ? row.r : 0.0, j, L);
                    const float vj = __shfl_sync(smask, mine ? row.V : 0.0f, j, L);
                    if (base + j == d - 1) Rv = rj + (double)__fmul_rn(p.gamma_f32, vj);
                    else Rv = rj + p.gamma * Rv;
                    if (j == lane) myR = Rv;
                }
                if (mine) {
                    row.W = row.W + myR;
                    row.n_e += 1;
                    if (i < d - 1) row.nn_flags += 1;  // this row's node is the parent of the next path row
                    store_crow(rows + ri, row);
                }
            }
            if (lane == 0 && d > 0) rows[0].nn_flags += 1;  // root.n
            __syncwarp(smask);
        }

        if (SELECT) {
            const int64_t tree = p.tree_id0 + t;
            // the whole parent array of the tree: one coalesced load, kept in registers
            uint32_t pwds[PW];
            {
                const uint32_t* src = reinterpret_cast<const uint32_t*>(par) + lane * PW;
                if (PW == 1) pwds[0] = src[0];
                else if (PW == 2) { const uint2 v = *reinterpret_cast<const uint2*>(src); pwds[0] = v.x; pwds[1] = v.y; }
                else { const uint4 v = *reinterpret_cast<const uint4*>(src); pwds[0] = v.x; pwds[1] = v.y; pwds[2 % PW] = v.z; pwds[3 % PW] = v.w; }
            }
            int n_rows = p.n_rows[t], draws = p.draws[t], pwc = p.pw[t];
            const CRow root = load_crow(rows);
            int cur = 0;
            uint32_t cur_nn = root.nn_flags & CROW_NMASK;
            float cur_V = root.V;
            int depth = 0, sel = -1, expand = 0;
            float sel_action = 0.0f;
            uint32_t levels = 0, scanned = 0;
            bool nan = false, overflow = false;
            while (true) {
                // children of `cur`, in insertion (= row) order
                uint32_t m = 0;
#pragma unroll
                for (int w = 0; w < PW; ++w)
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if (((pwds[w] >> (8 * b)) & 0xFFu) == (uint32_t)cur) m |= 1u << (4 * w + b);
                const int cnt = __popc(m);
                int incl = cnt;
#pragma unroll
                for (int off = 1; off < L; off <<= 1) {
                    const int v = __shfl_up_sync(smask, incl, off, L);
                    if (lane >= off) incl += v;
                }
                const int C = __shfl_sync(smask, incl, L - 1, L);
                ++levels;
                if (p.pw_table[cur_nn] - C > 0) {
                    // progressive widening: sample a new action, append the edge, select it (mcts.py:725-727)
                    if (n_rows >= p.R || n_rows >= L * ROWS_PER_LANE) { overflow = true; break; }
                    float a;
                    if (p.use_tape) a = p.tapeA[(size_t)t * p.R + n_rows];
                    else a = sample_action(p, p.chead + ((size_t)t * p.R + cur) * p.HS, s_u[sub], s_z[sub]);
                    const int nr = n_rows++;
                    ++pwc;
                    if (lane == 0) {
                        CRow row;
                        row.W = 0.0; row.r = 0.0; row.V = 0.0f; row.action = a; row.n_e = 0; row.nn_flags = 0;
                        store_crow(rows + nr, row);
                        par[nr] = (uint8_t)cur;
                    }
                    sel = nr; sel_action = a; expand = 1;
                    if (lane == 0) path[depth] = (uint8_t)sel;
                    ++depth;
                    break;
                }
                if (C > L) { overflow = true; break; }
                // gather the children rows, one per lane
                {
                    int pos = incl - cnt;
                    uint32_t mm = m;
                    while (mm) {
                        const int b = __ffs(mm) - 1;
                        mm &= mm - 1;
                        s_kid[sub][pos++] = (uint8_t)(lane * ROWS_PER_LANE + b);
                    }
                }
                __syncwarp(smask);
                const bool has = lane < C;
                CRow c;
                c.W = 0.0; c.r = 0.0; c.V = 0.0f; c.action = 0.0f; c.n_e = 0; c.nn_flags = 0;
                int myrow = 0;
                if (has) {
                    myrow = s_kid[sub][lane];
                    c = load_crow(rows + myrow);
                }
                scanned += C;
                // UCT_j = Q_j + c_uct*(sqrt(node.n+1)/(n_j+1))   (mcts.py:731-732)
                const double sq = sqrt((double)(cur_nn + 1));
                const double Q = c.n_e > 0 ? c.W / (double)c.n_e : (double)cur_V;
                const double uct = Q + p.c_uct * (sq / (double)(c.n_e + 1));
                nan |= __any_sync(smask, has && (uct != uct)) != 0;
                int j;
                bool random_pick = false;
                if (p.epsilon != 0) {
                    const double x = (double)u32_to_unit(rng_select_u32(p, tree, draws++));
                    random_pick = x < p.epsilon;
                }
                if (random_pick) {
                    j = u32_to_index(rng_select_u32(p, tree, draws++), C);
                } else {
                    double mx = has ? uct : -CUDART_INF;
#pragma unroll
                    for (int off = L / 2; off > 0; off >>= 1) {
                        const double o = __shfl_xor_sync(smask, mx, off, L);
                        mx = o > mx ? o : mx;
                    }
                    const uint32_t win = (__ballot_sync(smask, has && uct == mx) >> ((tid & 31) / L * L)) & ((L == 32) ? 0xFFFFFFFFu : ((1u << L) - 1u));
                    const int nw = __popc(win);
                    // random.choice(winners): the draw is consumed even when there is a single winner
                    const int pick = nw > 1 ? u32_to_index(rng_select_u32(p, tree, draws), nw) : 0;
                    ++draws;
                    j = nw > 0 ? (int)__fns(win, 0, pick + 1) : 0;
                }
                sel = __shfl_sync(smask, myrow, j, L);
                const uint32_t nnf = __shfl_sync(smask, c.nn_flags, j, L);
                const float Vj = __shfl_sync(smask, c.V, j, L);
                sel_action = __shfl_sync(smask, c.action, j, L);
                if (lane == 0) path[depth] = (uint8_t)sel;
                ++depth;
                __syncwarp(smask);  // s_kid is rewritten at the next level
                if (!(nnf & CROW_EXPANDED)) { expand = 1; break; }
                cur = sel;
                cur_nn = nnf & CROW_NMASK;
                cur_V = Vj;
                if (nnf & CROW_TERMINAL) { expand = 0; break; }
            }
            if (lane == 0) {
                if (nan) atomicOr(p.err, ERR_NAN);
                if (overflow) atomicOr(p.err, ERR_CAPACITY);
                p.n_rows[t] = n_rows;
                p.draws[t] = draws;
                p.pw[t] = pwc;
                p.depth[t] = overflow ? 0 : depth;
                p.ctr[t] += levels;
                p.ctr[(size_t)p.B + t] += scanned;
                ExpandReq rq;
                rq.cur = cur; rq.sel = sel; rq.action = sel_action; rq.expand = overflow ? -1 : expand;
                s_req[sub] = rq;
            }
        }
    }

    // ---- phase C: expansion, one thread per tree -----------------------------------------------------
    if (SELECT) {
        __syncthreads();
        if (tid < TREES_PER_CTA && t0 + tid < p.B) {
            const int tt = t0 + tid;
            const ExpandReq rq = s_req[tid];
            if (rq.expand == 1) {
                const double2 s = p.cstate[(size_t)tt * p.R + rq.cur];
                double nth, nthdot, rew;
                const bool term = env::pendulum_step(s.x, s.y, rq.action, nth, nthdot, rew);
                CRow* row = p.crows + (size_t)tt * p.R + rq.sel;
                row->r = rew / AZG_PENDULUM_R_SCALE;  // mcts.py:687
                row->nn_flags = CROW_EXPANDED | (term ? CROW_TERMINAL : 0u);
                if (p.use_tape) row->V = term ? 0.0f : p.tapeV[(size_t)tt * p.R + rq.sel];
                p.cstate[(size_t)tt * p.R + rq.sel] = make_double2(nth, nthdot);
                p.X[tt] = env::pendulum_obs(nth, nthdot);
                p.leaf[tt] = rq.sel | LEAF_EVAL | (term ? LEAF_TERMINAL : 0);
            } else if (rq.expand == 0) {
                p.leaf[tt] = rq.cur;
                p.ctr[(size_t)2 * p.B + tt] += 1;
            } else {
                p.leaf[tt] = 0;
            }
        }
    }
}

// numpy pairwise summation for n <= 128 (np.sum in get_on_policy_value_target, mcts.py:111)
__device__ __forceinline__ double np_sum(const double* a, int n) {
    if (n < 8) {
        double r = 0.;
        for (int i = 0; i < n; ++i) r += a[i];
        return r;
    }
    double r[8];
    for (int k = 0; k < 8; ++k) r[k] = a[k];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
        for (int k = 0; k < 8; ++k) r[k] += a[i + k];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}

// MCTS.return_results (mcts.py:269-307): root children in insertion order.  One thread per tree; runs
// once per search, so the simple serial scan of the parent bytes is fine.
__global__ void k_results_continuous(const TreeParams p, int cmax, float* actions, int32_t* counts, double* Q, double* Vt,
                                     int32_t* nchild) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.B) return;
    const CRow* rows = p.crows + (size_t)t * p.R;
    const uint8_t* par = p.cparent + (size_t)t * p.PSTRIDE;
    const float Vroot = rows[0].V;
    const int n_rows = p.n_rows[t];
    int C = 0;
    double q[64];
    int32_t cn[64];
    for (int i = 1; i < n_rows && C < cmax && C < 64; ++i) {
        if (par[i] != 0) continue;
        const CRow c = load_crow(rows + i);
        q[C] = c.n_e > 0 ? c.W / (double)c.n_e : (double)Vroot;
        cn[C] = c.n_e;
        actions[(size_t)t * cmax + C] = c.action;
        counts[(size_t)t * cmax + C] = c.n_e;
        Q[(size_t)t * cmax + C] = q[C];
        ++C;
    }
    for (int i = C; i < cmax; ++i) {
        actions[(size_t)t * cmax + i] = 0.0f;
        counts[(size_t)t * cmax + i] = 0;
        Q[(size_t)t * cmax + i] = 0.0;
    }
    nchild[t] = C;
    double v = 0.0;
    if (C > 0) {
        if (p.v_target == 1) {
            long long tot = 0;
            for (int i = 0; i < C; ++i) tot += cn[i];
            for (int i = 0; i < C; ++i) q[i] = ((double)cn[i] / (double)tot) * q[i];
            v = np_sum(q, C);
        } else {
            v = q[0];
            for (int i = 1; i < C; ++i) v = q[i] > v ? q[i] : v;
        }
    }
    Vt[t] = v;
}
