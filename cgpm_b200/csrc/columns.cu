// Column (view) reassignment: State.transition_dims / _gibbs_transition_dim
// (src/crosscat/state.py:804-810,1050-1123).
//
//   aux_kernel      the auxiliary view of a column: alpha ~ grid, a CRP-prior
//                   partition of all rows (state.py:1078-1081 -> view.py:99-103 ->
//                   crp.py:71-81, a strictly sequential draw), then the column's
//                   marginal likelihood under it.  One warp per (chain, column):
//                   the first 32 tables live in registers as running prefix sums
//                   (one ballot per draw), later tables in a 32-ary count tree.
//   (stats_kernel, MARGINAL mode, gives the marginals under the existing views.)
//   scan_kernel     the sequential part: per column, CRP weights (crp.py:111-146)
//                   + Ci mask (state.py:1098-1102) + log_pflip draw + bookkeeping
//                   (state.py:1125-1154).  One warp per chain.  A move into the
//                   auxiliary view pauses the chain; the host installs the view
//                   (install_kernel) and resumes.
#include <cstdlib>
#include "cc_common.cuh"
#include <math_constants.h>

#define CC_MAXCAT 256

int cc_launch_stats(const cc_state* st, int n_work, const int32_t* d_chains, const int32_t* d_views,
                    int marginal, const int32_t* d_cols, const int32_t* d_coloff, const int32_t* d_colcnt,
                    int max_cols, const uint8_t* d_mask, cudaStream_t stream);

namespace {

__device__ __forceinline__ int warp_incl_scan_i(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

struct AuxParams {
    cc_state st;
    int n;
    const int32_t* chains;
    const int32_t* cols;
    const int32_t* aidx;     // injected grid index per work item, or NULL
    const double* u;         // injected uniforms [n][R-1], or NULL
    int write_zr;
    int prof;
    int smem_l1;
    uint64_t seed, counter;
    int32_t* arena;
    size_t worker_stride;    // int32 units
    int* queue;              // device work counter (dynamic assignment of pairs to warps)
};

__global__ void __launch_bounds__(128) aux_kernel(const AuxParams p) {
    const cc_state& st = p.st;
    const int lane = threadIdx.x & 31;
    const int worker = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nworkers = gridDim.x * (blockDim.x >> 5);
    const int R = st.R;
    const int n1 = (R + 31) >> 5, n2 = (n1 + 31) >> 5;
    int32_t* base = p.arena + (size_t)worker * p.worker_stride;
    int32_t* z = base;                      // [R] table of each row
    int32_t* Lv[4];
    Lv[0] = z + R;                          // [R]   counts of tables >= 32
    Lv[1] = Lv[0] + R;                      // [n1]
    Lv[2] = Lv[1] + n1;                     // [n2]
    Lv[3] = Lv[2] + n2;                     // [32]
    int32_t* start = Lv[3] + 32;            // [R+1]
    int32_t* cursor = start + (R + 1);      // [R]
    int32_t* grouped = cursor + R;          // [R]
    const int ntree = R + n1 + n2 + 32;
    // the upper levels of the count tree are kept in shared memory when they fit (p.smem_l1:
    // levels 1-3, else levels 2-3); the leaf level stays in the arena (L2)
    extern __shared__ int32_t s_tree[];
    {
        const int per_warp = (p.smem_l1 ? n1 : 0) + n2 + 32 + 128;
        int32_t* sw = s_tree + (size_t)(threadIdx.x >> 5) * per_warp;
        if (p.smem_l1) { Lv[1] = sw; sw += n1; }
        Lv[2] = sw;
        Lv[3] = sw + n2;
    }
    int32_t* zs = Lv[3] + 32;               // [128] table ids of the rows of one seating group (device stream only)
    const int nup = (p.smem_l1 ? n1 : 0) + n2 + 32;
    const Philox rng(p.seed);

    // pairs are handed out dynamically: their cost varies ~5x with the drawn alpha
    (void)nworkers;
    for (;;) {
        int w = 0;
        if (lane == 0) w = atomicAdd(p.queue, 1);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= p.n) break;
        const int s = p.chains[w], c = p.cols[w];
        // ---- alpha of the fresh view: rng.choice(grid) (view.py:99, dim.py:181-183)
        int ai;
        if (p.aidx) ai = p.aidx[w];
        else { const uint4 r4 = rng(p.counter, stream_id(4, s, c)); ai = (int)(((uint64_t)r4.x * CC_NGRID) >> 32); }
        const double alpha = st.row_alpha_grid[ai];
        const long long t0 = clock64();
        // ---- pass 1: CRP-prior seating of all rows (crp.py:71-81) ----------------------------
        int K = 1;              // tables so far
        int P = 1;              // this lane's inclusive prefix of the counts of tables 0..31 (table 0 holds row 0)
        int reg_total = 1, mem_total = 0;
        const bool member_draw = (p.u == nullptr);
        if (member_draw) {
            // Device stream: row i opens a table with probability alpha / (i + alpha), otherwise it joins the
            // table of a uniformly chosen earlier ROW -- the same law as choosing a table in proportion to its
            // size (crp.py:178-182), from the same single uniform t = u (i + alpha): t >= i opens a table,
            // else floor(t) is the row to join.  Nothing here depends on the table counts, so 128 rows are
            // seated at a time: rows whose chosen row lies before the group read its table (one gather), the few
            // that point into the group follow the chain through shared memory.  Table ids are in order of
            // appearance, exactly as in the sequential draw.  Lv[0] receives the table sizes afterwards.
            const uint64_t ustream = stream_id(3, s, c);
            if (lane == 0) z[0] = 0;
            __syncwarp();
            for (int i0 = 1; i0 < R; i0 += 128) {
                int par[4], zreg[4];
                bool act[4];
                int nnew_before = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int ii = i0 + q * 32 + lane;
                    act[q] = ii < R;
                    double uu = 0.0;
                    if (act[q]) { const uint4 r4 = rng(p.counter * 0x100000000ull + (uint64_t)ii, ustream); uu = u01_from_words(r4.x, r4.y); }
                    const double tt = uu * ((double)ii + alpha);
                    const bool isnew = act[q] && tt >= (double)ii;
                    par[q] = (int)fmin(tt, 2.0e9);
                    const unsigned nb = __ballot_sync(0xffffffffu, isnew);
                    zreg[q] = isnew ? K + nnew_before + __popc(nb & ((1u << lane) - 1u)) : -1;
                    nnew_before += __popc(nb);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (act[q] && zreg[q] < 0 && par[q] < i0) zreg[q] = z[par[q]];
#pragma unroll
                for (int q = 0; q < 4; ++q) zs[q * 32 + lane] = zreg[q];
                __syncwarp();
                for (;;) {
                    bool pend = false;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (act[q] && zreg[q] < 0) {
                            const int zp = zs[par[q] - i0];
                            if (zp >= 0) zreg[q] = zp; else pend = true;
                        }
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < 4; ++q) zs[q * 32 + lane] = zreg[q];
                    __syncwarp();
                    if (!__any_sync(0xffffffffu, pend)) break;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (act[q]) z[i0 + q * 32 + lane] = zreg[q];             // coalesced stores
                K += nnew_before;
                __syncwarp();
            }
            // table sizes (only the grouped statistics pass of categorical columns needs them)
            if (st.ctype[c] == CC_CATEGORICAL) for (int j = lane; j < K; j += 32) Lv[0][j] = 0;
            __syncwarp();
            for (int b0 = 0; st.ctype[c] == CC_CATEGORICAL && b0 < R; b0 += 32) {
                const int i = b0 + lane;
                const int k = (i < R) ? z[i] : -1;
                const unsigned peers = __match_any_sync(0xffffffffu, k);
                if (k >= 0 && (peers >> lane) == 1u) Lv[0][k] += __popc(peers);      // the highest peer adds for the group
                __syncwarp();
            }
        } else {
            for (int i = lane; i < (p.smem_l1 ? R : R + n1); i += 32) Lv[0][i] = 0;   // leaf (+ level 1 when it is in the arena)
            {
                int32_t* up = p.smem_l1 ? Lv[1] : Lv[2];
                for (int i = lane; i < nup; i += 32) up[i] = 0;
            }
            (void)ntree;
            __syncwarp();
            if (lane == 0) z[0] = 0;
            const uint64_t ustream = stream_id(3, s, c);
            // Rows are seated in batches of 32: every lane first turns its row's uniform into
            // t = u (i + alpha) (crp.py:178-182 normalised by N + alpha) and keeps floor(t) and
            // the "new table" bit t >= i, which do not depend on the seating; the sequential part
            // of a row is then one shuffle, one integer compare and one ballot (table counts are
            // integers, so prefix > t  <=>  prefix > floor(t)).
            for (int i0 = 1; i0 < R; i0 += 32) {
                const int ii = i0 + lane;
                double uu = 0.0;
                if (ii < R) {
                    if (p.u) uu = p.u[(size_t)w * (R - 1) + (ii - 1)];
                    else { const uint4 r4 = rng(p.counter * 0x100000000ull + (uint64_t)ii, ustream); uu = u01_from_words(r4.x, r4.y); }
                }
                const double tt = uu * ((double)ii + alpha);
                int packed = (int)fmin(tt, 2.0e9);                    // floor(t), t >= 0
                if (tt >= (double)ii) packed |= 0x80000000;           // new table
                int zreg = 0;
                const int nb = min(32, R - i0);
                for (int l = 0; l < nb; ++l) {
                    const int pk = __shfl_sync(0xffffffffu, packed, l);
                    const int tl = pk & 0x7fffffff;
                    int j;
                    // Tables >= 32 live in a 32-ary tree whose nodes store the INCLUSIVE PREFIX of their
                    // children's counts (entries past the last table hold the node total), so a search is
                    // one load + one ballot per level and an update is `entry[lane] += (lane >= child)`.
                    if (pk < 0) {
                        j = K;                                        // new table (index K)
                        if (K < 32) { if (lane >= K) P += 1; reg_total += 1; }
                        else {
                            const int jp = K - 32;
                            // a level comes into use when its first node fills up: everything seated so far
                            // belongs to its child 0
                            if (jp == 32) Lv[1][lane] = mem_total;
                            if (jp == 1024) Lv[2][lane] = mem_total;
                            if (jp == 32768) Lv[3][lane] = mem_total;
                            __syncwarp();
                            if (lane >= (jp & 31)) Lv[0][(jp & ~31) + lane] += 1;
                            if (jp >= 32 && lane >= ((jp >> 5) & 31)) Lv[1][((jp >> 5) & ~31) + lane] += 1;
                            if (jp >= 1024 && lane >= ((jp >> 10) & 31)) Lv[2][((jp >> 10) & ~31) + lane] += 1;
                            if (jp >= 32768 && lane >= (jp >> 15)) Lv[3][lane] += 1;
                            mem_total += 1;
                            __syncwarp();
                        }
                        K += 1;
                    } else if (tl < reg_total) {
                        const unsigned hit = __ballot_sync(0xffffffffu, P > tl);
                        j = __ffs(hit) - 1;
                        if (lane >= j) P += 1;
                        reg_total += 1;
                    } else {
                        int tr = tl - reg_total;
                        const int Kp = K - 32;
                        const int top = (Kp <= 32) ? 0 : (Kp <= 1024) ? 1 : (Kp <= 32768) ? 2 : 3;
                        int idx = 0;
                        for (int lev = top; lev >= 0; --lev) {
                            int32_t* node = Lv[lev] + idx * 32;
                            const int inc = node[lane];
                            unsigned hit = __ballot_sync(0xffffffffu, inc > tr);
                            if (!hit) hit = 0x80000000u;      // unreachable in exact arithmetic
                            const int h = __ffs(hit) - 1;
                            const int excl = __shfl_sync(0xffffffffu, inc, (h + 31) & 31);
                            if (h > 0) tr -= excl;
                            if (lane >= h) node[lane] = inc + 1;
                            idx = idx * 32 + h;
                        }
                        __syncwarp();
                        j = idx + 32;
                        mem_total += 1;
                    }
                    if (lane == l) zreg = j;
                }
                if (ii < R) z[ii] = zreg;                             // coalesced store of the batch
            }
        }
        __syncwarp();
        const long long t1 = clock64();
        if (member_draw && st.ctype[c] != CC_CATEGORICAL) {      // the sum types (normal, bernoulli, poisson, exponential, geometric)
            // ---- pass 2, normal column on the device stream: rows in their natural order, one 32-byte record
            //      (sum_x, sum_x_sq, count of present cells) per table.  The lowest lane of each group of rows
            //      that share a table loads the record, adds the group's cells in row order and stores it, so
            //      every table's sums are accumulated in row order exactly as before -- without the grouping
            //      pass, the table-size pass and the gathers through it (one random record per row instead of
            //      five random accesses).
            double* rec = reinterpret_cast<double*>(base + ((R + 7) & ~7));
            for (int e = lane; e < 4 * K; e += 32) rec[e] = 0.0;
            if (p.write_zr == 1 && lane == 0) st.aux_K[s] = K;
            if (p.write_zr == 2 && lane == 0) st.aux_K_all[(size_t)s * st.C + c] = K;
            __syncwarp();
            const double* xc = st.Xcm + (size_t)c * R;
            const int ctq = st.ctype[c];
            // second accumulated term: x*x for a normal column; gammaln(x+1) for poisson (a proposal weight: the device lgamma will do)
            auto sq = [&](double v) -> double { return (ctq == CC_NORMAL) ? __dmul_rn(v, v) : sum_second_term_ni(ctq, v, CUDART_NAN); };
            for (int b0 = 0; b0 < R; b0 += 32) {
                const int i = b0 + lane;
                const bool act = i < R;
                const int k = act ? z[i] : -1;
                const double x = act ? xc[i] : CUDART_NAN;
                if (act && p.write_zr == 1) st.aux_zr[(size_t)s * R + i] = k;
                if (act && p.write_zr == 2) st.aux_zr_all[((size_t)s * st.C + c) * R + i] = k;
                const unsigned peers = __match_any_sync(0xffffffffu, k);
                const bool leader = act && (__ffs(peers) - 1 == lane);
                double sx = 0.0, sxx = 0.0, cn = 0.0;
                if (leader) { sx = rec[4 * k]; sxx = rec[4 * k + 1]; cn = rec[4 * k + 2]; }
                if (__all_sync(0xffffffffu, !act || peers == (1u << lane))) {
                    if (leader && !isnan(x)) { cn += 1.0; sx = __dadd_rn(sx, x); sxx = __dadd_rn(sxx, sq(x)); }
                } else {
                    for (int l = 0; l < 32; ++l) {
                        const int kl = __shfl_sync(0xffffffffu, k, l);
                        const double xl = __shfl_sync(0xffffffffu, x, l);
                        if (leader && kl == k && !isnan(xl)) { cn += 1.0; sx = __dadd_rn(sx, xl); sxx = __dadd_rn(sxx, sq(xl)); }
                    }
                }
                if (leader) { rec[4 * k] = sx; rec[4 * k + 1] = sxx; rec[4 * k + 2] = cn; }
                __syncwarp();
            }
            const double* hyp = st.hyp + ((size_t)s * st.C + c) * 4;
            const double h0 = hyp[0], h1 = hyp[1], h2 = hyp[2], h3 = hyp[3];
            double lsum = 0.0;
            for (int j = lane; j < K; j += 32) lsum += sum_marginal(ctq, rec[4 * j + 2], rec[4 * j], rec[4 * j + 1], h0, h1, h2, h3);
            lsum = warp_sum(lsum);
            if (lane == 0) {
                st.colL[((size_t)s * st.C + c) * (st.Vcap + 1) + st.Vcap] = lsum;
                st.aux_alpha[(size_t)s * st.C + c] = alpha;
                if (p.prof) {
                    int32_t* q = st.seq + ((size_t)s * st.C + c) * 4;
                    q[0] = (int)((t1 - t0) >> 10); q[1] = (int)((clock64() - t1) >> 10); q[2] = K; q[3] = ai;
                }
            }
            __syncwarp();
            continue;
        }
        // ---- pass 2: ordered per-table statistics of column c, marginal ------------
        // counts -> start offsets
        {
            const int pm1 = __shfl_up_sync(0xffffffffu, P, 1);
            const int myc = (lane == 0) ? P : P - pm1;          // count of table `lane`
            int carry = 0;
            for (int b0 = 0; b0 < K; b0 += 32) {
                const int j = b0 + lane;
                int cn = 0;
                if (j < K) cn = member_draw ? Lv[0][j] : (b0 == 0) ? myc : (Lv[0][j - 32] - (((j - 32) & 31) ? Lv[0][j - 33] : 0));
                const int inc = warp_incl_scan_i(cn, lane);
                if (j < K) { start[j] = carry + inc - cn; cursor[j] = carry + inc - cn; }
                carry += __shfl_sync(0xffffffffu, inc, 31);
            }
            if (lane == 0) start[K] = carry;
            if (p.write_zr == 1 && lane == 0) st.aux_K[s] = K;
            if (p.write_zr == 2 && lane == 0) st.aux_K_all[(size_t)s * st.C + c] = K;
        }
        __syncwarp();
        for (int b0 = 0; b0 < R; b0 += 32) {
            const int i = b0 + lane;
            const bool act = i < R;
            const int k = act ? z[i] : -1;
            const unsigned peers = __match_any_sync(0xffffffffu, k);
            const int rank = __popc(peers & ((1u << lane) - 1u));
            int pos = 0;
            if (act) pos = cursor[k] + rank;
            __syncwarp();
            if (act && rank == __popc(peers) - 1) cursor[k] = pos + 1;
            if (act) grouped[pos] = i;
            if (act && p.write_zr == 1) st.aux_zr[(size_t)s * R + i] = k;
            if (act && p.write_zr == 2) st.aux_zr_all[((size_t)s * st.C + c) * R + i] = k;
            __syncwarp();
        }
        const double* hyp = st.hyp + ((size_t)s * st.C + c) * 4;
        const double h0 = hyp[0], h1 = hyp[1], h2 = hyp[2], h3 = hyp[3];
        const int ctype = st.ctype[c], kc = st.ncat[c];
        const double* xc = st.Xcm + (size_t)c * R;
        auto sq2 = [&](double v) -> double { return (ctype == CC_NORMAL) ? __dmul_rn(v, v) : sum_second_term_ni(ctype, v, CUDART_NAN); };
        double lsum = 0.0;
        const int BIG = 96;       // tables at least this large are summed by the whole warp
        for (int b0 = 0; b0 < K; b0 += 32) {
            const int j = b0 + lane;
            const int beg = (j < K) ? start[j] : 0, end = (j < K) ? start[j + 1] : 0;
            // large tables: the warp gathers 32 members at a time (coalesced for runs of consecutive
            // rows) and every lane replays the same sequential adds, so the order is unchanged
            unsigned big = __ballot_sync(0xffffffffu, end - beg >= BIG);
            while (big) {
                const int l0 = __ffs(big) - 1;
                big &= big - 1;
                const int bb = __shfl_sync(0xffffffffu, beg, l0), ee = __shfl_sync(0xffffffffu, end, l0);
                double N = 0.0, sx = 0.0, sxx = 0.0;
                if (ctype != CC_CATEGORICAL) {
                    for (int i0 = bb; i0 < ee; i0 += 32) {
                        const int i = i0 + lane;
                        const double xv = (i < ee) ? xc[grouped[i]] : CUDART_NAN;
                        const int cntm = min(32, ee - i0);
                        for (int l = 0; l < cntm; ++l) {
                            const double v = __shfl_sync(0xffffffffu, xv, l);
                            if (!isnan(v)) { N += 1.0; sx = __dadd_rn(sx, v); sxx = __dadd_rn(sxx, sq2(v)); }
                        }
                    }
                    if (lane == 0) lsum += sum_marginal(ctype, N, sx, sxx, h0, h1, h2, h3);
                } else {
                    uint32_t lc[CC_MAXCAT];
                    for (int q = 0; q < kc; ++q) lc[q] = 0;
                    for (int i = bb + lane; i < ee; i += 32) {
                        const double x0 = xc[grouped[i]];
                        if (!isnan(x0)) lc[(int)x0]++;
                    }
                    const double A = kc * h0;
                    double lg = 0.0;
                    int Ni = 0;
                    for (int q = 0; q < kc; ++q) {
                        const int tq = __reduce_add_sync(0xffffffffu, (int)lc[q]);
                        Ni += tq;
                        lg += lgamma((double)tq + h0);
                    }
                    if (lane == 0) lsum += lgamma(A) - lgamma(A + (double)Ni) + lg - kc * lgamma(h0);
                }
            }
            if (j >= K || end - beg >= BIG) continue;
            double N = 0.0, sx = 0.0, sxx = 0.0;
            if (ctype != CC_CATEGORICAL) {
                int i = beg;
                for (; i + 4 <= end; i += 4) {
                    const double x0 = xc[grouped[i]], x1 = xc[grouped[i + 1]], x2 = xc[grouped[i + 2]], x3 = xc[grouped[i + 3]];
                    if (!isnan(x0)) { N += 1.0; sx = __dadd_rn(sx, x0); sxx = __dadd_rn(sxx, sq2(x0)); }
                    if (!isnan(x1)) { N += 1.0; sx = __dadd_rn(sx, x1); sxx = __dadd_rn(sxx, sq2(x1)); }
                    if (!isnan(x2)) { N += 1.0; sx = __dadd_rn(sx, x2); sxx = __dadd_rn(sxx, sq2(x2)); }
                    if (!isnan(x3)) { N += 1.0; sx = __dadd_rn(sx, x3); sxx = __dadd_rn(sxx, sq2(x3)); }
                }
                for (; i < end; ++i) {
                    const double x0 = xc[grouped[i]];
                    if (!isnan(x0)) { N += 1.0; sx = __dadd_rn(sx, x0); sxx = __dadd_rn(sxx, sq2(x0)); }
                }
                lsum += sum_marginal(ctype, N, sx, sxx, h0, h1, h2, h3);
            } else {
                uint32_t lc[CC_MAXCAT];
                for (int q = 0; q < kc; ++q) lc[q] = 0;
                for (int i = beg; i < end; ++i) {
                    const double x0 = xc[grouped[i]];
                    if (!isnan(x0)) { N += 1.0; lc[(int)x0]++; }
                }
                const double A = kc * h0;
                double lg = 0.0;
                for (int q = 0; q < kc; ++q) lg += lgamma((double)lc[q] + h0);
                lsum += lgamma(A) - lgamma(A + N) + lg - kc * lgamma(h0);
            }
        }
        lsum = warp_sum(lsum);
        if (lane == 0) {
            st.colL[((size_t)s * st.C + c) * (st.Vcap + 1) + st.Vcap] = lsum;
            st.aux_alpha[(size_t)s * st.C + c] = alpha;
            if (p.prof) {
                int32_t* q = st.seq + ((size_t)s * st.C + c) * 4;
                q[0] = (int)((t1 - t0) >> 10); q[1] = (int)((clock64() - t1) >> 10); q[2] = K; q[3] = ai;
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
struct ScanParams {
    cc_state st;
    int n_chains;
    const int32_t* chains;
    const int32_t* cols;     // [n_chains][ncols]
    int ncols;
    int32_t* pos;            // [n_chains] in/out: next list position
    int32_t* pending;        // [n_chains] out: -1 finished, else the list position that moves to its auxiliary view
    const double* u;         // [n_chains][ncols] injected uniforms or NULL
    uint64_t seed, counter;
    const int32_t* ci_off;   // [C+1] independence constraints (CSR) or NULL
    const int32_t* ci_adj;
    double* trace;           // [n_chains][ncols][trace_stride] or NULL
    int trace_stride;
};

__global__ void __launch_bounds__(32) scan_kernel(const ScanParams p) {
    const cc_state& st = p.st;
    const int w = blockIdx.x, lane = threadIdx.x;
    const int s = p.chains[w];
    const int C = st.C, Vcap = st.Vcap;
    int32_t* zv = st.zv + (size_t)s * C;
    int32_t* nv = st.nv + (size_t)s * Vcap;
    int32_t* view_id = st.view_id + (size_t)s * Vcap;
    extern __shared__ double s_mem[];
    double* lp = s_mem;                                   // [Vcap+1]
    int* cand = reinterpret_cast<int*>(lp + Vcap + 1);    // [Vcap+1] view slot per candidate (-1 = aux)
    const double alpha = st.col_alpha[s];
    const Philox rng(p.seed);
    const int32_t* cols = p.cols + (size_t)w * p.ncols;
    int j = p.pos[w];
    for (; j < p.ncols; ++j) {
        const int c = cols[j];
        const int own = zv[c];
        const bool singleton = nv[own] == 1;
        // candidates: live views in ascending view id (crp.py:139), then the auxiliary view
        int V = 0;
        for (int v0 = 0; v0 < Vcap; v0 += 32) {
            const int v = v0 + lane;
            const bool livev = v < Vcap && nv[v] > 0;
            if (livev) {
                int rank = 0;
                const int myid = view_id[v];
                for (int q = 0; q < Vcap; ++q) if (nv[q] > 0 && view_id[q] < myid) ++rank;
                cand[rank] = v;
            }
            V += __popc(__ballot_sync(0xffffffffu, livev));
        }
        __syncwarp();
        const int ncand = V + (singleton ? 0 : 1);
        if (lane == 0 && !singleton) cand[V] = -1;
        __syncwarp();
        const double* Lrow = st.colL + ((size_t)s * C + c) * (Vcap + 1);
        for (int q = lane; q < ncand; q += 32) {
            const int v = cand[q];
            double wgt, ld;
            if (v < 0) { wgt = alpha; ld = Lrow[Vcap]; }
            else { ld = Lrow[v]; wgt = (v == own) ? (singleton ? alpha : (double)(nv[own] - 1)) : (double)nv[v]; }
            double val = ld + log(wgt);
            if (p.ci_off && v >= 0) {
                for (int e = p.ci_off[c]; e < p.ci_off[c + 1]; ++e)
                    if (zv[p.ci_adj[e]] == v) val = -INFINITY;              // state.py:1098-1102
            }
            lp[q] = val;
        }
        __syncwarp();
        int choice = 0;
        if (ncand > 1) {
            double uu;
            if (p.u) uu = p.u[(size_t)w * p.ncols + j];
            else { const uint4 r4 = rng(p.counter, stream_id(5, s, c)); uu = u01_from_words(r4.x, r4.y); }
            double mx = -INFINITY;
            for (int q = lane; q < ncand; q += 32) mx = fmax(mx, lp[q]);
            mx = warp_max(mx);
            double carry = 0.0;
            for (int b0 = 0; b0 < ncand; b0 += 32) {
                const int q = b0 + lane;
                const double e = (q < ncand) ? exp(lp[q] - mx) : 0.0;
                carry += __shfl_sync(0xffffffffu, warp_incl_scan(e, lane), 31);
            }
            const double tot = carry;
            carry = 0.0;
            choice = -1;
            for (int b0 = 0; b0 < ncand && choice < 0; b0 += 32) {
                const int q = b0 + lane;
                const double e = (q < ncand) ? exp(lp[q] - mx) : 0.0;
                const double inc = warp_incl_scan(e, lane) + carry;
                const unsigned hit = __ballot_sync(0xffffffffu, q < ncand && inc > uu * tot);
                if (hit) choice = b0 + __ffs(hit) - 1;
                carry = __shfl_sync(0xffffffffu, inc, 31);
            }
            if (choice < 0) {                 // NaN / all -inf: stay, flag
                if (lane == 0) st.err[s] = CC_ERR_ARG;
                for (int q = 0; q < ncand; ++q) if (cand[q] == own) choice = q;
            }
        }
        const int vnew = cand[choice];
        if (p.trace && lane == 0) {
            double* tr = p.trace + ((size_t)w * p.ncols + j) * p.trace_stride;
            int maxid = -1;
            for (int q = 0; q < Vcap; ++q) if (nv[q] > 0) maxid = max(maxid, view_id[q]);
            tr[0] = (double)c; tr[1] = (double)ncand;
            tr[2] = (double)(vnew < 0 ? maxid + 1 : view_id[vnew]);
            for (int q = 0; q < ncand && q + 3 < p.trace_stride; ++q) tr[3 + q] = lp[q];
        }
        if (vnew < 0) {                       // into the auxiliary view: host installs it
            if (lane == 0) { p.pending[w] = j; p.pos[w] = j; }
            return;
        }
        if (vnew != own && lane == 0) {       // state.py:1125-1145
            zv[c] = vnew;
            nv[vnew] += 1;
            nv[own] -= 1;
            if (nv[own] == 0) { view_id[own] = -1; st.nclu[sv_index(st, s, own)] = 0; }
        }
        __syncwarp();
    }
    if (lane == 0) { p.pending[w] = -1; p.pos[w] = p.ncols; }
}

// install an auxiliary partition (mode 1: aux_zr / aux_K of the chain, mode 2: the stored
// aux_zr_all / aux_K_all of the (chain, column)) as a new view and move the column into it
// (state.py:1110-1116, 1147-1154).  Table sizes are recounted from the partition.
__global__ void __launch_bounds__(256) install_kernel(const cc_state st, int n, const int32_t* chains,
                                                      const int32_t* cols, int mode) {
    const int w = blockIdx.x;
    if (w >= n) return;
    const int s = chains[w], c = cols[w];
    const int R = st.R, C = st.C, Vcap = st.Vcap, Kcap = st.Kcap;
    int32_t* nv = st.nv + (size_t)s * Vcap;
    int32_t* view_id = st.view_id + (size_t)s * Vcap;
    const int K = (mode == 2) ? st.aux_K_all[(size_t)s * C + c] : st.aux_K[s];
    const int32_t* src = (mode == 2) ? st.aux_zr_all + ((size_t)s * C + c) * R : st.aux_zr + (size_t)s * R;
    __shared__ int s_slot, s_newid;
    if (threadIdx.x == 0) {
        int slot = -1, maxid = -1;
        for (int v = 0; v < Vcap; ++v) {
            if (nv[v] > 0) maxid = max(maxid, view_id[v]);
            else if (slot < 0) slot = v;
        }
        if (K > Kcap - 1) slot = -1;
        s_slot = slot;
        s_newid = maxid + 1;
        if (slot < 0) st.err[s] = CC_ERR_CAPACITY;
    }
    __syncthreads();
    const int slot = s_slot;
    if (slot < 0) return;
    const size_t sv = sv_index(st, s, slot);
    for (int k = threadIdx.x; k < Kcap; k += blockDim.x) {
        st.live[sv * Kcap + k] = (k < K) ? k : -1;
        st.cid[sv * Kcap + k] = (k < K) ? k : 0;
        st.nk[sv * Kcap + k] = 0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < R; i += blockDim.x) {
        const int k = src[i];
        st.zr[sv * R + i] = k;
        st.order[sv * R + i] = i;
        atomicAdd(&st.nk[sv * Kcap + k], 1);
    }
    if (threadIdx.x == 0) {
        const int own = st.zv[(size_t)s * C + c];
        st.nclu[sv] = K;
        st.view_alpha[sv] = st.aux_alpha[(size_t)s * C + c];
        for (int g = 0; g < CC_NGRID; ++g) st.view_grid[sv * CC_NGRID + g] = st.row_alpha_grid[g];   // view.py:98-99
        view_id[slot] = s_newid;
        nv[slot] = 1;
        st.zv[(size_t)s * C + c] = slot;
        nv[own] -= 1;
        if (nv[own] == 0) { view_id[own] = -1; st.nclu[sv_index(st, s, own)] = 0; }
    }
}

template <typename T>
int to_device(T** out, const T* host, size_t n, cudaStream_t stream) {
    *out = nullptr;
    if (n == 0 || !host) return CC_OK;
    CC_CUDA(cudaMallocAsync(out, sizeof(T) * n, stream));
    CC_CUDA(cudaMemcpyAsync(*out, host, sizeof(T) * n, cudaMemcpyHostToDevice, stream));
    return CC_OK;
}

}  // namespace

extern "C" int cc_col_aux(const cc_state* st, int n, const int32_t* chains, const int32_t* cols,
                          const int32_t* aidx, const double* u_dev, int write_zr, uint64_t seed,
                          uint64_t counter, void* stream_) {
    if (!st || n < 0) { cc_set_error("cc_col_aux: bad arguments"); return CC_ERR_ARG; }
    if (n == 0) return CC_OK;
    if (write_zr == 2 && !st->aux_zr_all) { cc_set_error("cc_col_aux: write_zr=2 needs aux_zr_all"); return CC_ERR_ARG; }
    if (st->R > (1 << 20) + 32) { cc_set_error("cc_col_aux: R > 2^20 not supported by the count tree"); return CC_ERR_ARG; }
    cudaStream_t stream = (cudaStream_t)stream_;
    cc_prepare_pool();
    const int R = st->R;
    const size_t n1 = (R + 31) >> 5, n2 = (n1 + 31) >> 5;
    size_t per = (size_t)R + (R + n1 + n2 + 32) + (R + 1) + R + R;
    const size_t per_direct = (((size_t)R + 7) & ~(size_t)7) + 8 * (size_t)R;      // z + one 32-byte record per table
    if (per_direct > per) per = per_direct;
    per = (per + 7) & ~(size_t)7;
    const size_t avail = (size_t)st->arena_bytes / sizeof(int32_t);
    if (!st->arena || avail < per) { cc_set_error("cc_col_aux: arena too small (%lld bytes, need %lld per worker)", (long long)st->arena_bytes, (long long)(per * 4)); return CC_ERR_ARG; }
    int nworkers = (int)(avail / per);
    if (nworkers > n) nworkers = n;
    int dev = 0, nsm = 148;
    CC_CUDA(cudaGetDevice(&dev));
    CC_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    const int wpb = 4;
    int blocks = (nworkers + wpb - 1) / wpb;
    if (blocks > nsm * 8) blocks = nsm * 8;
    if (const char* e = getenv("CGPM_B200_AUX_BLOCKS")) { const int v = atoi(e); if (v > 0 && v < blocks) blocks = v; }
    if (blocks * wpb > nworkers) blocks = nworkers / wpb > 0 ? nworkers / wpb : 1;
    int tpb = (blocks == 1 && nworkers < wpb) ? nworkers * 32 : wpb * 32;
    AuxParams p;
    p.st = *st; p.n = n; p.aidx = nullptr; p.u = u_dev; p.write_zr = write_zr; p.seed = seed; p.counter = counter;
    p.arena = (int32_t*)st->arena; p.worker_stride = per;
    p.prof = getenv("CGPM_B200_AUX_PROF") ? 1 : 0;
    int32_t *dch = nullptr, *dco = nullptr, *dai = nullptr;
    if (to_device(&dch, chains, n, stream) || to_device(&dco, cols, n, stream) || to_device(&dai, aidx, aidx ? n : 0, stream)) return CC_ERR_CUDA;
    p.chains = dch; p.cols = dco; p.aidx = dai;
    int* dq = nullptr;
    CC_CUDA(cudaMallocAsync(&dq, sizeof(int), stream));
    CC_CUDA(cudaMemsetAsync(dq, 0, sizeof(int), stream));
    p.queue = dq;
    const int nw = tpb / 32;
    p.smem_l1 = ((size_t)nw * (n1 + n2 + 32 + 128) * 4 <= 56 * 1024) ? 1 : 0;
    const size_t smem = (size_t)nw * ((p.smem_l1 ? n1 : 0) + n2 + 32 + 128) * 4;
    CC_CUDA(cudaFuncSetAttribute(aux_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    aux_kernel<<<blocks, tpb, smem, stream>>>(p);
    CC_CUDA(cudaGetLastError());
    cudaFreeAsync(dch, stream); cudaFreeAsync(dco, stream); cudaFreeAsync(dq, stream);
    if (dai) cudaFreeAsync(dai, stream);
    return CC_OK;
}

extern "C" int cc_col_marginals(const cc_state* st, int n_work, const int32_t* chains, const int32_t* views,
                                const int32_t* cols_dev, const int32_t* coloff, const int32_t* colcnt,
                                int max_cols, void* stream_) {
    if (!st || n_work < 0) { cc_set_error("cc_col_marginals: bad arguments"); return CC_ERR_ARG; }
    if (n_work == 0) return CC_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    cc_prepare_pool();
    int32_t *dch = nullptr, *dvw = nullptr, *dof = nullptr, *dcn = nullptr;
    if (to_device(&dch, chains, n_work, stream) || to_device(&dvw, views, n_work, stream)) return CC_ERR_CUDA;
    if (cols_dev) {
        if (!coloff || !colcnt) { cc_set_error("cc_col_marginals: column lists need offsets and counts"); return CC_ERR_ARG; }
        if (to_device(&dof, coloff, n_work, stream) || to_device(&dcn, colcnt, n_work, stream)) return CC_ERR_CUDA;
    }
    const int rc = cc_launch_stats(st, n_work, dch, dvw, 1, cols_dev, dof, dcn, max_cols, nullptr, stream);
    cudaFreeAsync(dch, stream); cudaFreeAsync(dvw, stream);
    if (dof) cudaFreeAsync(dof, stream);
    if (dcn) cudaFreeAsync(dcn, stream);
    return rc;
}

extern "C" int cc_col_scan(const cc_state* st, int n_chains, const int32_t* chains_dev, const int32_t* cols_dev,
                           int ncols, int32_t* pos_dev, int32_t* pending_dev, const double* u_dev,
                           uint64_t seed, uint64_t counter, const int32_t* ci_off_dev, const int32_t* ci_adj_dev,
                           double* trace_dev, int trace_stride, void* stream_) {
    if (!st || n_chains < 0 || !chains_dev || !cols_dev || !pos_dev || !pending_dev) { cc_set_error("cc_col_scan: bad arguments"); return CC_ERR_ARG; }
    if (n_chains == 0) return CC_OK;
    ScanParams p;
    p.st = *st; p.n_chains = n_chains; p.chains = chains_dev; p.cols = cols_dev; p.ncols = ncols;
    p.pos = pos_dev; p.pending = pending_dev; p.u = u_dev; p.seed = seed; p.counter = counter;
    p.ci_off = ci_off_dev; p.ci_adj = ci_adj_dev; p.trace = trace_dev; p.trace_stride = trace_stride;
    const size_t smem = (size_t)(st->Vcap + 1) * (sizeof(double) + sizeof(int)) + 16;
    scan_kernel<<<n_chains, 32, smem, (cudaStream_t)stream_>>>(p);
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

extern "C" int cc_col_install(const cc_state* st, int n, const int32_t* chains, const int32_t* cols, int mode, void* stream_) {
    if (!st || n < 0) { cc_set_error("cc_col_install: bad arguments"); return CC_ERR_ARG; }
    if (n == 0) return CC_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    cc_prepare_pool();
    int32_t *dch = nullptr, *dco = nullptr;
    if (to_device(&dch, chains, n, stream) || to_device(&dco, cols, n, stream)) return CC_ERR_CUDA;
    if (mode == 2 && !st->aux_zr_all) { cc_set_error("cc_col_install: no stored partitions"); return CC_ERR_ARG; }
    install_kernel<<<n, 256, 0, stream>>>(*st, n, dch, dco, mode);
    CC_CUDA(cudaGetLastError());
    cudaFreeAsync(dch, stream); cudaFreeAsync(dco, stream);
    return CC_OK;
}
