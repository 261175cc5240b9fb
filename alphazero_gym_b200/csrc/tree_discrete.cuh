// tree_discrete.cuh -- backup + PUCT select + expansion for the discrete (CartPole) tree.
//
// Reference: MCTSDiscrete.search (mcts.py:418-462), selectionUCT (:464-493), epsilon_greedy (:175-195),
// helpers.argmax (helpers.py:30-52), MCTS.backprop (mcts.py:241-267), Action.update (states.py:97-112),
// expansion (mcts.py:215-238).
//
// Mapping: fan-out is A = 2, so a warp-wide argmax would idle 30 lanes; instead one THREAD owns one tree
// and a warp advances 32 trees in lockstep.  A node visit is one 64 B row = two full 32 B sectors, read
// with 4 x LDG.128; the level loop is a chain of dependent row loads (parent -> chosen child).  The env
// step runs after the loop so that all 32 lanes execute the f64 dynamics together.
#pragma once
#include "common.cuh"
#include "env.cuh"

__device__ __forceinline__ DRow load_drow(const DRow* p) {
    DRow r;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&r);
    d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[3];
    return r;
}

// root row + first network input (MCTSDiscrete.initialize_search, mcts.py:364-383)
__device__ __forceinline__ void d_init(const TreeParams& p, int t) {
    const double* s = p.root_state + (size_t)t * 4;
    double* st = p.dstate + (size_t)t * p.R * 4;
    DRow row;
    row.W[0] = row.W[1] = 0.0;
    row.r = 0.0;
    row.n_e[0] = row.n_e[1] = 0;
    row.prior[0] = row.prior[1] = 0.0f;
    row.V = 0.0f;
    row.node_n = p.root_n_init ? p.root_n_init[t] : 0;
    row.child[0] = row.child[1] = DROW_NONE;
    row.parent = DROW_NONE;
    row.paction = 0;
    row.flags = 0;
    row.pad[0] = row.pad[1] = 0;
    if (p.use_tape) {
        row.V = p.tapeV[(size_t)t * p.R];
        row.prior[0] = p.tapeP[(size_t)t * p.R * 2];
        row.prior[1] = p.tapeP[(size_t)t * p.R * 2 + 1];
    }
    p.drows[(size_t)t * p.R] = row;
    st[0] = s[0]; st[1] = s[1]; st[2] = s[2]; st[3] = s[3];
    p.X[t] = make_float4((float)s[0], (float)s[1], (float)s[2], (float)s[3]);
    p.leaf[t] = 0 | LEAF_EVAL;
    p.n_rows[t] = 1;
    p.draws[t] = 0;
    if (p.rng_mt) mt_seed_dev(p.mt + (size_t)t * (MT_N + 1), __ldg(p.seedp) + (uint64_t)(tree_base(p) + t));  // random.seed(seed + tree)
    for (int k = 0; k < 4; ++k) p.ctr[(size_t)k * p.B + t] = 0;
}
__global__ void k_init_discrete(const TreeParams p) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < p.B) d_init(p, t);
}

// one simulation step of tree t: backup of the previous simulation (BACKUP), then descent + expansion of the next (SELECT).
// Shared by k_step_discrete (one launch per simulation) and the whole-search kernel (qmlp2.cuh, FUSED).
__device__ __forceinline__ void d_step(const TreeParams& p, const Tabs& tb, int t, const bool BACKUP, const bool SELECT) {
    DRow* rows = p.drows + (size_t)t * p.R;

    uint16_t* path = p.dpath + (size_t)t * p.R;
    TP_BEGIN();
    if (BACKUP) {
        // R = leaf.V; up the path: R = node.r + gamma*R; edge.n += 1; edge.W += R; parent.n += 1 (mcts.py:241-267).
        // The select step recorded the path (row, action) root -> leaf, so the row addresses are known up front and the rows are
        // fetched four at a time; following the parent links instead made every level a dependent L2 round trip, and CartPole
        // traces are 8 levels deep on average (tens in the deepest tree of a warp).
        const int leaf = p.leaf[t] & LEAF_ROW_MASK;
        const int d = p.ddepth[t];
        double Rv, r;
        {
            const DRow lr = load_drow(rows + leaf);
            Rv = (double)lr.V;
            r = lr.r;
        }
#pragma unroll 1
        for (int base = d - 1; base >= 0; base -= 4) {
            DRow q[4];
            int pe[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                pe[k] = base - k >= 0 ? (int)path[base - k] : 0;
                q[k] = load_drow(rows + (pe[k] >> 1));
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (base - k >= 0) {
                    DRow* pr = rows + (pe[k] >> 1);
                    const int pa = pe[k] & 1;
                    Rv = r + p.gamma * Rv;
                    pr->W[pa] = (pa ? q[k].W[1] : q[k].W[0]) + Rv;
                    pr->n_e[pa] = (pa ? q[k].n_e[1] : q[k].n_e[0]) + 1;
                    pr->node_n = q[k].node_n + 1;
                    r = q[k].r;
                }
            }
        }
    }

    TP_STAMP(1);
    if (SELECT) {
        const int64_t tree = tree_base(p) + t;
        int draws = p.draws[t];
        int cur = 0;
        DRow row = load_drow(rows);
        int a = -1;
        uint32_t levels = 0;
        uint32_t xr[4] = {0, 0, 0, 0};
        uint32_t* mt = p.rng_mt ? p.mt + (size_t)t * (MT_N + 1) : nullptr;
        int mti = p.rng_mt ? (int)mt[MT_N] : 0;
        bool nan = false;
        while (true) {
            // both children are LOADED while the scores are computed (32 registers): the level costs max(round trip, arithmetic)
            // instead of their sum, and the next level's children are requested the moment the choice is made
            const int c0 = row.child[0], c1 = row.child[1];
            uint4 k0[4], k1[4];
            {
                const uint4* s0 = reinterpret_cast<const uint4*>(rows + (c0 != DROW_NONE ? c0 : cur));
                const uint4* s1 = reinterpret_cast<const uint4*>(rows + (c1 != DROW_NONE ? c1 : cur));
#pragma unroll
                for (int q = 0; q < 4; ++q) { k0[q] = s0[q]; k1[q] = s1[q]; }
            }
            // UCT_a = Q_a + prior_a*c_uct*(sqrt(node.n+1)/(n_a+1))   (mcts.py:483-484)
            const double sq = sqrt_small(row.node_n + 1, tb.sq, tb.n);
            double u[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int n = row.n_e[i];
                const double Q = n > 0 ? div_small(row.W[i], n, tb.rcp, tb.n) : (double)row.V;
                const double pc = p.puct_f32 ? (double)__fmul_rn(row.prior[i], (float)p.c_uct) : (double)row.prior[i] * p.c_uct;
                u[i] = Q + pc * div_small(sq, n + 1, tb.rcp, tb.n);
            }
            nan |= (u[0] != u[0]) || (u[1] != u[1]);
#ifdef AZG_TREE_PROF
            if (u[0] + u[1] == 12345.678) atomicOr(p.err, 0);  // (profiling builds: the scores are complete before the stamp)
            TP_STAMP(5);
#endif
            bool random_pick = false;
            if (p.epsilon != 0 && p.rng_mt) {
                random_pick = mt_random_dev(mt, mti, draws) < p.epsilon;
            } else if (p.epsilon != 0) {  // epsilon_greedy: random.random() < eps -> random.randint(0, A-1)
                // every level consumes exactly two draws (random(), then randint or choice), so the random() of the next four
                // levels are draws + 0, 2, 4, 6: generated together (the generator was half of a level's dependent chain)
                const int k = (int)(levels & 3u);
                if (k == 0) rng_select_u32x4(__ldg(p.seedp), tree, draws, 2, xr);
                const uint32_t xk = k == 0 ? xr[0] : (k == 1 ? xr[1] : (k == 2 ? xr[2] : xr[3]));
                ++draws;
                const double x = (double)u32_to_unit(xk);
                random_pick = x < p.epsilon;
            }
            if (p.rng_mt) {  // randint(0, 1), or choice over the 1 or 2 winners (which also draws for a single winner)
                if (random_pick) a = mt_below_dev(mt, mti, draws, 2);
                else if (u[0] == u[1]) a = mt_below_dev(mt, mti, draws, 2);
                else { (void)mt_below_dev(mt, mti, draws, 1); a = u[1] > u[0] ? 1 : 0; }
            } else if (random_pick) {
                a = u32_to_index(rng_select_u32(p, tree, draws++), 2);
            } else {
                // argmax with random tie-break: one draw is consumed even for a single winner
                if (u[0] == u[1]) a = u32_to_index(rng_select_u32(p, tree, draws), 2);
                else a = u[1] > u[0] ? 1 : 0;
                ++draws;
            }
#ifdef AZG_TREE_PROF
            if (a == 77) atomicOr(p.err, 0);
            TP_STAMP(6);
            if ((threadIdx.x & 31) == 0 && p.prof) atomicAdd(p.prof + 7, 1ull);
#endif
            path[levels] = (uint16_t)((cur << 1) | a);
            ++levels;
            const int child = a ? c1 : c0;
            if (child == DROW_NONE) break;  // expansion
            cur = child;
            {
                uint4* d = reinterpret_cast<uint4*>(&row);
#pragma unroll
                for (int q = 0; q < 4; ++q) d[q] = a ? k1[q] : k0[q];
            }
            if (row.flags & ROW_TERMINAL) { a = -1; break; }  // trace ends on an existing terminal node
#ifdef AZG_TREE_PROF
            if (row.node_n == -77) atomicOr(p.err, 0);  // (the next row has arrived)
            TP_STAMP(3);
#endif
        }
        TP_STAMP(2);
        if (nan) atomicOr(p.err, ERR_NAN);
        if (p.rng_mt) mt[MT_N] = (uint32_t)mti;
        p.ddepth[t] = (int)levels;  // edges between the root and the leaf (the new node, or the existing terminal node the trace ended on)
        p.draws[t] = draws;
        p.ctr[t] += levels;
        p.ctr[(size_t)p.B + t] += levels * 2;
        if (a >= 0) {
            const int child = p.n_rows[t];
            if (child >= p.R) { atomicOr(p.err, ERR_CAPACITY); p.leaf[t] = cur; return; }
            p.n_rows[t] = child + 1;
            double* st = p.dstate + ((size_t)t * p.R + cur) * 4;
            const double s[4] = {st[0], st[1], st[2], st[3]};
            double o[4], rew;
            const bool term = env::cartpole_step(s, a, o, rew);
            double* so = p.dstate + ((size_t)t * p.R + child) * 4;
            so[0] = o[0]; so[1] = o[1]; so[2] = o[2]; so[3] = o[3];
            DRow nr;
            nr.W[0] = nr.W[1] = 0.0;
            nr.r = rew;
            nr.n_e[0] = nr.n_e[1] = 0;
            nr.prior[0] = nr.prior[1] = 0.0f;
            nr.V = 0.0f;
            nr.node_n = 0;
            nr.child[0] = nr.child[1] = DROW_NONE;
            nr.parent = (uint16_t)cur;
            nr.paction = (uint8_t)a;
            nr.flags = term ? ROW_TERMINAL : 0;
            nr.pad[0] = nr.pad[1] = 0;
            if (p.use_tape) {
                const size_t ti = (size_t)t * p.R + child;
                nr.V = term ? 0.0f : p.tapeV[ti];
                nr.prior[0] = p.tapeP[ti * 2];
                nr.prior[1] = p.tapeP[ti * 2 + 1];
            }
            rows[child] = nr;
            rows[cur].child[a] = (uint16_t)child;
            p.X[t] = make_float4((float)o[0], (float)o[1], (float)o[2], (float)o[3]);
            p.leaf[t] = child | LEAF_EVAL | (term ? LEAF_TERMINAL : 0);
        } else {
            p.leaf[t] = cur;
            p.ctr[(size_t)2 * p.B + t] += 1;
        }
        TP_STAMP(4);
    }
}

template <bool BACKUP, bool SELECT>
__global__ void __launch_bounds__(128) k_step_discrete(const TreeParams p) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const Tabs tb = {p.pw_table, p.rcp_tab, p.sqrt_tab, AZG_TAB};
    if (t < p.B) d_step(p, tb, t, BACKUP, SELECT);
}

// MCTS.return_results (mcts.py:269-307) for the discrete root
__global__ void k_results_discrete(const TreeParams p, int cmax, float* actions, int32_t* counts, double* Q, double* Vt,
                                   int32_t* nchild) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.B) return;
    const DRow row = load_drow(p.drows + (size_t)t * p.R);
    double q[2];
    for (int a = 0; a < 2; ++a) {
        const int n = row.n_e[a];
        q[a] = n > 0 ? row.W[a] / (double)n : (double)row.V;
        actions[(size_t)t * cmax + a] = (float)a;
        counts[(size_t)t * cmax + a] = n;
        Q[(size_t)t * cmax + a] = q[a];
    }
    nchild[t] = 2;
    double v;
    if (p.v_target == 1) {  // on_policy: np.sum((counts / np.sum(counts)) * Q), n < 8 -> sequential sum from 0.
        const double tot = (double)(row.n_e[0] + row.n_e[1]);
        v = 0.0;
        v += ((double)row.n_e[0] / tot) * q[0];
        v += ((double)row.n_e[1] / tot) * q[1];
    } else {
        v = q[1] > q[0] ? q[1] : q[0];
    }
    Vt[t] = v;
}
