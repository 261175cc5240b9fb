// tree_continuous.cuh -- backup + progressive-widening / UCT select + expansion for the continuous tree.
//
// Reference: MCTSContinuous.search (mcts.py:656-702), selectionUCT (:704-741), add_pw_action (:625-654),
// NodeContinuous.check_pw (states.py:252-275), epsilon_greedy (mcts.py:175-195), helpers.argmax
// (helpers.py:30-52), MCTS.backprop (mcts.py:241-267), DiagonalGMMPolicy.sample_action
// (policies.py:656-669) with SquashedNormal (distributions.py:50-63, :205-245).
//
// Mapping: ONE THREAD PER TREE, a warp advances 32 trees in lockstep.  The first version of this kernel
// gave each tree a 16-lane sub-warp with shuffle arg-max over children gathered one per lane; ncu showed
// it issue-bound (every lane re-executes the scalar control flow: ~480 warp instructions per simulated
// tree, profiles/r1a_ncu_step.csv) and barrier-bound on its dense f64 phases.  With a thread per tree the
// same work costs ~1/10 of the issue slots, the f64 env step / Box-Muller run on full warps with no
// barriers, and memory-level parallelism comes from 65536 independent threads instead of from lanes.
//
// Per level the thread holds the node's inline child list (sector 1 of its row, 2 x LDG.128), gathers the
// children's statistics sectors four at a time (independent LDG.128 pairs in flight), evaluates UCT in
// f64 in the reference's expression order and keeps (max, bit-set of winners) for the random tie-break.
#pragma once
#include "common.cuh"
#include "env.cuh"

struct CHot {  // sector 0 of a CRow
    double W, r;
    float V, action;
    int32_t n_e;
    uint32_t nn_flags;
};

__device__ __forceinline__ CHot load_hot(const CRow* p) {
    CHot h;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&h);
    d[0] = s[0]; d[1] = s[1];
    return h;
}
__device__ __forceinline__ void load_kids(const CRow* p, uint32_t kw[8]) {
    const uint4* s = reinterpret_cast<const uint4*>(p);
    const uint4 a = s[2], b = s[3];
    kw[0] = a.x; kw[1] = a.y; kw[2] = a.z; kw[3] = a.w;
    kw[4] = b.x; kw[5] = b.y; kw[6] = b.z; kw[7] = b.w;
}
__device__ __forceinline__ void store_new_row(CRow* p, double r, float V, float action, uint32_t nn_flags) {
    CHot h;
    h.W = 0.0; h.r = r; h.V = V; h.action = action; h.n_e = 0; h.nn_flags = nn_flags;
    const uint4* s = reinterpret_cast<const uint4*>(&h);
    uint4* d = reinterpret_cast<uint4*>(p);
    d[0] = s[0]; d[1] = s[1];
    d[2] = make_uint4(0, 0, 0, 0);
    d[3] = make_uint4(0, 0, 0, 0);
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

__device__ __forceinline__ float new_action(const TreeParams& p, int t, int node, int row, int j) {
    if (p.use_tape) return p.tapeA[(size_t)t * p.R + row];
    float u, z[AZG_MAX_K];
    pw_noise(p, p.tree_id0 + t, j, u, z);
    return sample_action(p, p.chead + ((size_t)t * p.R + node) * p.HS, u, z);
}

// MCTSContinuous.initialize_search (mcts.py:589-600): new root row, network input = obs(root)
__global__ void k_init_continuous(const TreeParams p) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.B) return;
    const double th = p.root_state[(size_t)t * 2], thdot = p.root_state[(size_t)t * 2 + 1];
    store_new_row(p.crows + (size_t)t * p.R, 0.0, p.use_tape ? p.tapeV[(size_t)t * p.R] : 0.0f, 0.0f, CROW_EXPANDED);
    p.cstate[(size_t)t * p.R] = make_double2(th, thdot);
    p.X[t] = env::pendulum_obs(th, thdot);
    p.leaf[t] = 0 | LEAF_EVAL;
    p.leafR[t] = 0.0;
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
    CRow* rows = p.crows + (size_t)t * p.R;
    const float a = new_action(p, t, 0, 1, 0);
    store_new_row(rows + 1, 0.0, 0.0f, a, 0);
    rows[0].kids[0] = 1;
    rows[0].nkids = 1;
    p.n_rows[t] = 2;
    p.pw[t] = 1;
}

#define KIND_INSERT 0    // progressive widening: create a new edge below `cur`, then its node
#define KIND_EXPAND 1    // an existing edge without a node (only the root's first edge when c_pw > 1)
#define KIND_TERMINAL 2  // the trace ended on an existing terminal node
#define KIND_ERROR 3

template <bool BACKUP, bool SELECT>
__global__ void __launch_bounds__(128) k_step_continuous(const TreeParams p) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.B) return;
    CRow* rows = p.crows + (size_t)t * p.R;
    uint8_t* path = p.path + (size_t)t * p.R;

    if (BACKUP) {
        // backprop (mcts.py:241-267) along the recorded path.  leafR already holds r_leaf + gamma*V_leaf
        // (the f32 product of NEP 50, added by the evaluation kernel's epilogue).
        const int d = p.depth[t];
        double Rv = p.leafR[t];
        for (int i = d - 1; i >= 0; --i) {
            CRow* row = rows + path[i];
            const CHot h = load_hot(row);
            if (i != d - 1) Rv = h.r + p.gamma * Rv;
            row->W = h.W + Rv;                                                      // Action.update (states.py:97-112)
            *reinterpret_cast<int2*>(&row->n_e) = make_int2(h.n_e + 1, (int)(h.nn_flags + (i < d - 1 ? 1u : 0u)));
        }
        if (d > 0) rows[0].nn_flags += 1;  // root.n
    }

    if (SELECT) {
        const int64_t tree = p.tree_id0 + t;
        int draws = p.draws[t], n_rows = p.n_rows[t], pwc = p.pw[t];
        int cur = 0, sel = -1, kind = KIND_ERROR, depth = 0;
        uint32_t kw[8];
        load_kids(rows, kw);
        const CHot root = load_hot(rows);
        uint32_t cur_nn = root.nn_flags & CROW_NMASK;
        float cur_V = root.V;
        int nk = kw[7] >> 24;
        float sel_action = 0.0f;
        double leaf_r = 0.0;
        uint32_t levels = 0, scanned = 0;
        bool nan = false;
        while (true) {
            ++levels;
            if (p.pw_table[cur_nn] - nk > 0) { kind = KIND_INSERT; break; }  // states.py:252-275
            // UCT_j = Q_j + c_uct*(sqrt(node.n+1)/(n_j+1))   (mcts.py:731-732), children in insertion order
            const double sq = sqrt((double)(cur_nn + 1));
            double best = -CUDART_INF;
            uint32_t win = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                if (w * 4 < nk) {
                    CHot c[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int j = w * 4 + i;
                        const int kid = j < nk ? (int)((kw[w] >> (8 * i)) & 0xFFu) : 0;
                        c[i] = load_hot(rows + kid);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int j = w * 4 + i;
                        if (j < nk) {
                            const int n = c[i].n_e;
                            const double Q = n > 0 ? c[i].W / (double)n : (double)cur_V;
                            const double u = Q + p.c_uct * (sq / (double)(n + 1));
                            nan |= (u != u);
                            if (u > best) { best = u; win = 1u << j; }
                            else if (u == best) win |= 1u << j;
                        }
                    }
                }
            }
            scanned += nk;
            int j;
            bool random_pick = false;
            if (p.epsilon != 0) {  // epsilon_greedy (mcts.py:175-195)
                const double x = (double)u32_to_unit(rng_select_u32(p, tree, draws++));
                random_pick = x < p.epsilon;
            }
            if (random_pick) {
                j = u32_to_index(rng_select_u32(p, tree, draws++), nk);
            } else {
                // random.choice(winners): the draw is consumed even when there is a single winner
                const int nw = __popc(win);
                const int pick = nw > 1 ? u32_to_index(rng_select_u32(p, tree, draws), nw) : 0;
                ++draws;
                j = nw > 0 ? (int)__fns(win, 0, pick + 1) : 0;
            }
            sel = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w)
                if ((j >> 2) == w) sel = (int)((kw[w] >> (8 * (j & 3))) & 0xFFu);
            const CHot sh = load_hot(rows + sel);
            path[depth++] = (uint8_t)sel;
            if (!(sh.nn_flags & CROW_EXPANDED)) { kind = KIND_EXPAND; sel_action = sh.action; break; }
            cur = sel;
            cur_nn = sh.nn_flags & CROW_NMASK;
            cur_V = sh.V;
            if (sh.nn_flags & CROW_TERMINAL) { kind = KIND_TERMINAL; leaf_r = sh.r; break; }
            load_kids(rows + cur, kw);
            nk = kw[7] >> 24;
        }
        if (nan) { atomicOr(p.err, ERR_NAN); kind = KIND_ERROR; }
        if (kind == KIND_INSERT && (n_rows >= p.R || nk >= CROW_MAX_KIDS)) { atomicOr(p.err, ERR_CAPACITY); kind = KIND_ERROR; }

        // ---- expansion: all lanes are reconverged here, so the f64 work below runs on full warps ----
        if (kind == KIND_INSERT) {
            // add_pw_action (mcts.py:625-654): the new edge is selected immediately (mcts.py:725-727)
            sel = n_rows++;
            sel_action = new_action(p, t, cur, sel, pwc++);
            rows[cur].kids[nk] = (uint8_t)sel;
            rows[cur].nkids = (uint8_t)(nk + 1);
            path[depth++] = (uint8_t)sel;
        }
        if (kind == KIND_INSERT || kind == KIND_EXPAND) {
            const double2 s = p.cstate[(size_t)t * p.R + cur];
            double nth, nthdot, rew;
            const bool term = env::pendulum_step(s.x, s.y, sel_action, nth, nthdot, rew);
            const double r = rew / AZG_PENDULUM_R_SCALE;  // mcts.py:687
            const uint32_t fl = CROW_EXPANDED | (term ? CROW_TERMINAL : 0u);
            float V = 0.0f;
            if (p.use_tape && !term) V = p.tapeV[(size_t)t * p.R + sel];
            if (kind == KIND_INSERT) {
                store_new_row(rows + sel, r, V, sel_action, fl);
            } else {
                rows[sel].r = r;
                rows[sel].V = V;
                rows[sel].nn_flags = fl;
            }
            p.cstate[(size_t)t * p.R + sel] = make_double2(nth, nthdot);
            p.X[t] = env::pendulum_obs(nth, nthdot);
            p.leaf[t] = sel | LEAF_EVAL | (term ? LEAF_TERMINAL : 0);
            // first backup step R = r + gamma*V: with tapes V is known here, otherwise the evaluation kernel adds it
            p.leafR[t] = p.use_tape ? r + (double)__fmul_rn(p.gamma_f32, V) : r;
        } else if (kind == KIND_TERMINAL) {
            p.leaf[t] = cur;
            p.leafR[t] = leaf_r + (double)__fmul_rn(p.gamma_f32, 0.0f);
            p.ctr[(size_t)2 * p.B + t] += 1;
        } else {
            p.leaf[t] = 0;
            depth = 0;
        }
        p.n_rows[t] = n_rows;
        p.draws[t] = draws;
        p.pw[t] = pwc;
        p.depth[t] = depth;
        p.ctr[t] += levels;
        p.ctr[(size_t)p.B + t] += scanned;
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

// MCTS.return_results (mcts.py:269-307): root children in insertion order, one thread per tree
__global__ void k_results_continuous(const TreeParams p, int cmax, float* actions, int32_t* counts, double* Q, double* Vt,
                                     int32_t* nchild) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.B) return;
    const CRow* rows = p.crows + (size_t)t * p.R;
    const float Vroot = rows[0].V;
    const int nk = min((int)rows[0].nkids, cmax);
    double q[CROW_MAX_KIDS + 1];
    int32_t cn[CROW_MAX_KIDS + 1];
    for (int j = 0; j < nk; ++j) {
        const CHot c = load_hot(rows + rows[0].kids[j]);
        q[j] = c.n_e > 0 ? c.W / (double)c.n_e : (double)Vroot;
        cn[j] = c.n_e;
        actions[(size_t)t * cmax + j] = c.action;
        counts[(size_t)t * cmax + j] = c.n_e;
        Q[(size_t)t * cmax + j] = q[j];
    }
    for (int j = nk; j < cmax; ++j) {
        actions[(size_t)t * cmax + j] = 0.0f;
        counts[(size_t)t * cmax + j] = 0;
        Q[(size_t)t * cmax + j] = 0.0;
    }
    nchild[t] = nk;
    double v = 0.0;
    if (nk > 0) {
        if (p.v_target == 1) {
            long long tot = 0;
            for (int j = 0; j < nk; ++j) tot += cn[j];
            for (int j = 0; j < nk; ++j) q[j] = ((double)cn[j] / (double)tot) * q[j];
            v = np_sum(q, nk);
        } else {
            v = q[0];
            for (int j = 1; j < nk; ++j) v = q[j] > v ? q[j] : v;
        }
    }
    Vt[t] = v;
}
