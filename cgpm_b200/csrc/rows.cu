// Row-reassignment Gibbs sweep: View.transition_rows (src/mixtures/view.py:247-252)
// = for every row in a permuted order: _gibbs_transition_row (view.py:393-406)
//   -> Crp.gibbs_tables / gibbs_logps (src/primitives/crp.py:111-146)
//   -> _logpdf_row_gibbs / _logpdf_cell_gibbs (view.py:408-422)
//   -> gu.log_pflip (src/utils/general.py:123-139)
//   -> _migrate_row (view.py:424-429).
//
// One persistent CTA per (chain, view).  The result is always that of the strictly sequential
// scan of the reference; what runs in parallel is speculation:
//
//   * SPECULATIVE ROUNDS (views of up to 31 clusters and up to DL columns -- the steady state).
//     A row that stays where it is changes nothing (its own cluster's sums go through
//     `-= x; += x`, view.py:416-419, which almost always returns the same doubles).  So W
//     consecutive rows of the visiting order are scored AND drawn at once against the current
//     tables: a group of g = gk * gc lanes per row (gk >= number of candidate clusters, gc lanes
//     sharing the view's columns), 32 / g rows per warp, every warp of the CTA in parallel.  Then
//     the rows are committed in order: the first row of the window that moves (or whose sums
//     drift) is applied by the whole CTA, the rows behind it are dropped and scored again in the
//     next round against the updated tables.  A round costs one scoring + one draw + one CTA
//     barrier and retires min(W, rows up to the first mover) rows.  The window adapts to the
//     rate of moves: views whose rows rarely move (wide views hold most of a table's evidence)
//     run with many rows in flight and few lanes per row, views whose rows move at every other
//     step (one or two columns) with few rows in flight and all lanes of a group on one row.
//   * ONE ROW AT A TIME (more than 31 clusters, or very wide views): all consumer warps score
//     one row (groups of G lanes per candidate cluster), warp 0 draws, everybody applies.
//   * LARGE-K KERNEL (more clusters than the shared-memory candidate arrays hold, e.g. the
//     reference-faithful CRP-prior initialisation of a view with alpha ~ R): same per-row work
//     with every per-cluster array in global memory and all O(K) bookkeeping done by the whole
//     CTA (block-wide scan for the draw, parallel shift of the ascending-id list, parallel search
//     of the next free slot).  A sweep is handed from the first kernel to this one -- and, when
//     the cluster slots themselves run out, back to the host, which grows Kcap -- through
//     rows_prog: {next row position, rows re-seated so far, status}; it continues with the same
//     row, and the random stream is indexed by the row position, so nothing changes.
//
// A dedicated producer warp stages the rows of the next batches into a shared-memory ring (whole
// 16-byte-aligned rows by 1-D bulk copies -- the TMA engine -- or per-cell cp.async gathers for
// narrow / very wide views), together with the row ids, their current cluster and their uniform,
// behind full / empty mbarriers.  The per-(cluster, column) tables (sufficient statistics + the
// cached predictive constants) live in shared memory when they fit and in global memory (L2)
// otherwise; the sweep is compiled once per address space and migrates from the shared to the
// global variant if the number of clusters outgrows shared memory mid-sweep.
//
// Exactness: sufficient statistics are updated with the same IEEE operations, in the same order,
// as the reference (x*x is rounded before it is added, normal.py:69; scoring a row against its
// own cluster does `-= x` then `+= x`, view.py:416-419), so they are bit-identical for the same
// sequence of moves.  The predictive is evaluated in the algebraically equal Student-t form
// (see cc_common.cuh) in fp64.
#include <algorithm>
#include <cstdlib>
#include <vector>
#include "cc_common.cuh"
#include "rows_common.cuh"

// This file is compiled twice: as rows.cu (CC_ROWS_COUNT 0: the build for all-normal tables -- the kernel sits at
// the 128-register limit of two 256-thread CTAs per SM, and every other column type costs it registers and a stack
// frame) and through rows_counts.cu (CC_ROWS_COUNT 1: tables that hold categorical, bernoulli, poisson, exponential
// or geometric columns; the two-pass scoring below covers views of normal + categorical columns there).
#ifndef CC_ROWS_COUNT
#define CC_ROWS_COUNT 0
#endif
#if CC_ROWS_COUNT
#define CC_ROWS_ENTRY cc_transition_rows_counts
#else
#define CC_ROWS_ENTRY cc_transition_rows
extern "C" int cc_transition_rows_counts(const cc_state* st, int n_work, const cc_row_work* work, const int32_t* perm,
                                         const double* u, double* trace, int trace_stride, uint64_t seed, uint64_t counter,
                                         void* stream_);
#endif

// -DCC_ROWS_PHASE=1: thread 0 accumulates the cycles of each phase of a speculative round into st.seq[16..] (probe builds only)
#ifndef CC_ROWS_PHASE
#define CC_ROWS_PHASE 0
#endif
#if CC_ROWS_PHASE
#define PH_DECL long long ph_t = clock64(); long long ph0 = 0, ph1 = 0, ph2 = 0, ph3 = 0, ph4 = 0, ph5 = 0, ph6 = 0, ph7 = 0; int ph_rounds = 0, ph_moves = 0;
#define PH(v) { const long long t_ = clock64(); v += t_ - ph_t; ph_t = t_; }
#else
#define PH_DECL
#define PH(v)
#endif

namespace {

constexpr bool CTS = CC_ROWS_COUNT != 0;      // column types other than normal are compiled in

enum { ST_OK = 0, ST_HANDOVER = CC_ROWS_HANDOVER, ST_GROW = CC_ROWS_GROW, ST_MIGRATE = 3 };
enum { LV_SOLO = 9, SOLO_ROWS = 256 };      // window setting "one warp on its own" and the rows of one stretch
enum { LV_ONE = 10, ONE_ROWS = 256 };       // window setting "the whole CTA on one row" (wide views whose rows move) and its stretch

__device__ __forceinline__ int ceil_log2(int x) {          // smallest l with (1 << l) >= x, x >= 1
    return (x <= 1) ? 0 : 32 - __clz(x - 1);
}

// gammaln(x + 1) of a poisson cell as the host computed it (cc_state.Xaux), NaN when not available
__device__ __forceinline__ double row_aux(const Ctx& c, int col, int r) {
    const double* xa = c.st.Xaux;
    if (!xa || r < 0) return CUDART_NAN;
    const int a = c.st.auxcol[col];
    return (a < 0) ? CUDART_NAN : xa[(size_t)a * c.st.R + r];
}

// ---------------------------------------------------------------------------------------------
// Apply a move of row r from cluster slot z to zb (view.py:424-429), by all NT consumer threads:
// threads over (column, old | new cluster) update the two clusters' tables (the row was scored
// against its own cluster first, so the old cluster's sums go through -= x, += x, -= x); the CRP
// bookkeeping (crp.py:45-59) is done by one thread, or by the whole CTA in the large-K kernel.
// The caller synchronises afterwards.
template <int NT, bool GT, bool BIG>
__device__ __forceinline__ void apply_move(const Ctx& c, const Tab<GT>& T, int r, int z, int zb, bool newc, bool singleton,
                                           int ownpos, int K, int nkz, const double* xrow) {
    const int tid = threadIdx.x;
    const int D = c.D, Kcap = c.st.Kcap;
    const bool wide = c.wide != 0;
    const double* ghyp = c.st.hyp + (size_t)c.s * c.st.C * 4;
    const int pslot = T.prior();
    const int half = NT / 2;
    const bool is_old = tid < half;
    for (int j = tid % half; j < D; j += half) {
        if (is_old && singleton) continue;            // the emptied cluster is dropped
        const double x = xrow[c.bulk ? c.col[j] : j];
        const int k = is_old ? z : zb;
        const bool isnan_x = isnan(x);
        const int ct = !CTS ? CC_NORMAL : (wide ? c.st.ctype[c.col[j]] : c.ctype[j]);
        if (ct == CC_NORMAL) {
            const double m = hyper(c, wide, ghyp, 0, j), rr = hyper(c, wide, ghyp, 1, j), ss = hyper(c, wide, ghyp, 2, j),
                         nu = hyper(c, wide, ghyp, 3, j);
            // One instruction stream for both clusters (a branch per role would run the two logs one after the other):
            //   old cluster: scored against itself first (-= x, += x: view.py:416-419), then the row leaves: ((s - x) + x) - x;
            //   new cluster: s + x, a fresh cluster object starting from 0 (normal.py:51-53,64-70).
            // G of the new count follows from the stored neighbour by one step of the recurrence: the old cluster knows
            // G(N - 1) (it becomes G(N')) and needs G(N' - 1); the new one knows G(N) (it becomes G(N' - 1)) and needs G(N').
            const bool fresh = !is_old && newc;
            if (isnan_x && !fresh) continue;               // a missing cell changes nothing; a fresh cluster still gets its record
            const double N0 = fresh ? 0.0 : T.at(CC_F_N, k, j);
            const double sx0 = fresh ? 0.0 : T.at(CC_F_SX, k, j), sxx0 = fresh ? 0.0 : T.at(CC_F_SXX, k, j);
            const double Gknown = is_old ? T.at(CC_F_GM1, k, j) : T.at(CC_F_GN, fresh ? pslot : k, j);
            const double xx = __dmul_rn(x, x);
            const double t1 = is_old ? __dadd_rn(__dsub_rn(sx0, x), x) : sx0;
            const double t2 = is_old ? __dadd_rn(__dsub_rn(sxx0, xx), xx) : sxx0;
            const double N = isnan_x ? 0.0 : N0 + (is_old ? -1.0 : 1.0);
            const double sx = isnan_x ? 0.0 : __dadd_rn(t1, is_old ? -x : x);
            const double sxx = isnan_x ? 0.0 : __dadd_rn(t2, is_old ? -xx : xx);
            const double Gs = normal_G_step(Gknown, N - 1.0, rr, nu);
            const double gN = (is_old || isnan_x) ? Gknown : Gs;
            const double gM1 = is_old ? ((N >= 1.0) ? Gs : 0.0) : (isnan_x ? 0.0 : Gknown);
            const NormalCache cc = normal_cache_fast(N, sx, sxx, gN, m, rr, ss, nu);
            T.at(CC_F_N, k, j) = N; T.at(CC_F_SX, k, j) = sx; T.at(CC_F_SXX, k, j) = sxx;
            T.at(CC_F_MN, k, j) = cc.mn; T.at(CC_F_A, k, j) = cc.a; T.at(CC_F_CST, k, j) = cc.cst;
            T.at(CC_F_GN, k, j) = gN; T.at(CC_F_GM1, k, j) = gM1;
        } else if (CTS && ct != CC_CATEGORICAL) {
            if constexpr (CTS) {
                // sum types: N, sum_x and (poisson) sum_log_fact_x by the reference's -= / += (bernoulli.py:46-58,
                // poisson.py:49-63, exponential.py:47-58, geometric.py:47-58), then the cached constants
                const double h0 = hyper(c, wide, ghyp, 0, j), h1 = hyper(c, wide, ghyp, 1, j);
                const bool fresh = !is_old && newc;
                if (isnan_x && !fresh) continue;
                double N = fresh ? 0.0 : T.at(CC_F_N, k, j), sx = fresh ? 0.0 : T.at(CC_F_SX, k, j), sxx = fresh ? 0.0 : T.at(CC_F_SXX, k, j);
                if (!isnan_x) {
                    const double t2 = sum_second_term_ni(ct, x, row_aux(c, c.col[j], r));
                    if (is_old) {       // scored against its own cluster first (-= x, += x), then removed
                        N -= 1.0;
                        sx = __dsub_rn(__dadd_rn(__dsub_rn(sx, x), x), x);
                        sxx = __dsub_rn(__dadd_rn(__dsub_rn(sxx, t2), t2), t2);
                    } else { N += 1.0; sx = __dadd_rn(sx, x); sxx = __dadd_rn(sxx, t2); }
                }
                const NormalCache cc = count_cache_ni(ct, N, sx, h0, h1);
                T.at(CC_F_N, k, j) = N; T.at(CC_F_SX, k, j) = sx; T.at(CC_F_SXX, k, j) = sxx;
                T.at(CC_F_MN, k, j) = cc.mn; T.at(CC_F_A, k, j) = cc.a; T.at(CC_F_CST, k, j) = cc.cst;
            }
        } else {
            const double a = hyper(c, wide, ghyp, 0, j);
            const int kc = wide ? c.st.ncat[c.col[j]] : c.ncat[j];
            double* cn = T.counts(k, j);
            const bool fresh = !is_old && newc;
            double N = fresh ? 0.0 : T.at(CC_F_N, k, j);
            if (fresh) for (int q = 0; q < kc; ++q) cn[q] = 0.0;
            if (isnan_x && !fresh) continue;
            if (!isnan_x) {
                const int xi = (int)x;
                if (is_old) { cn[xi] -= 1.0; N -= 1.0; }
                else { cn[xi] += 1.0; N += 1.0; }
            }
            T.at(CC_F_N, k, j) = N;
            T.at(CC_F_CST, k, j) = log(N + a * kc);
        }
    }
    // ---- bookkeeping of the row CRP (crp.py:45-59) ----
    if (!BIG) {
        if (tid == NT - 1) {
            c.zr[r] = zb;
            const int nm = c.scal[3];
            c.moved[nm] = r;
            c.movedflag[r] = 1;
            c.scal[3] = nm + 1;
            c.nk[z] = nkz - 1;
            if (newc) {
                c.g_cid[zb] = c.g_cid[c.live[K - 1]] + 1;              // crp.py:142 max(counts)+1
                c.live[K] = zb;
                c.nk[zb] = 1;
                c.scal[0] = K + 1;
                int fh = zb + 1;
                while (fh < c.Ksm - 1 && c.nk[fh] != 0) ++fh;
                c.scal[2] = fh;
            } else {
                c.nk[zb] += 1;
            }
        }
        if (singleton && tid >= NT - 32) {
            // the emptied cluster leaves the ascending-id list; its slot is free again (last warp, lanes over positions)
            const int lane = tid & 31;
            for (int q0 = ownpos; q0 < K - 1; q0 += 32) {
                const int q = q0 + lane;
                const int nxt = (q < K - 1) ? c.live[q + 1] : -1;
                __syncwarp();
                if (q < K - 1) c.live[q] = nxt;
                __syncwarp();
            }
            if (tid == NT - 1) {
                c.live[K - 1] = -1;
                c.scal[0] = K - 1;
                if (z < c.scal[2]) c.scal[2] = z;
            }
        }
    } else {
        if (tid == 0) {
            c.zr[r] = zb;
            const int nm = c.scal[3];
            c.moved[nm] = r;
            c.movedflag[r] = 1;
            c.scal[3] = nm + 1;
            c.nk[z] = nkz - 1;
            if (newc) {
                c.g_cid[zb] = c.g_cid[c.live[K - 1]] + 1;
                c.live[K] = zb;
                c.nk[zb] = 1;
                c.scal[0] = K + 1;
                c.scal[40] = Kcap - 1;
            } else {
                c.nk[zb] += 1;
            }
        }
        if (newc) {
            // next free slot above zb: windows of 4 slots per thread until one is found
            consumer_sync<NT>();
            for (int b0 = zb + 1; b0 < Kcap - 1; b0 += NT * 4) {
                for (int t = 0; t < 4; ++t) {
                    const int q = b0 + t * NT + tid;
                    if (q < Kcap - 1 && c.nk[q] == 0) { atomicMin(&c.scal[40], q); break; }
                }
                consumer_sync<NT>();
                const int found = c.scal[40];
                consumer_sync<NT>();             // everybody has read the verdict before the next window may change it
                if (found < Kcap - 1) break;
            }
            if (tid == 0) c.scal[2] = c.scal[40];
        }
        if (singleton) {
            for (int q0 = ownpos; q0 < K - 1; q0 += NT) {
                const int q = q0 + tid;
                const int nxt = (q < K - 1) ? c.live[q + 1] : -1;
                consumer_sync<NT>();
                if (q < K - 1) c.live[q] = nxt;
                consumer_sync<NT>();
            }
            if (tid == 0) {
                c.live[K - 1] = -1;
                c.scal[0] = K - 1;
                if (z < c.scal[2]) c.scal[2] = z;
            }
        }
    }
}

// The row stays, but `-= x; += x` changed its cluster's sums in the last bit (view.py:416-419):
// store them and refresh the cached constants.  All NT consumer threads.
template <int NT, bool GT>
__device__ __forceinline__ void apply_drift(const Ctx& c, const Tab<GT>& T, int z, int r, const double* xrow) {
    const bool wide = c.wide != 0;
    const double* ghyp = c.st.hyp + (size_t)c.s * c.st.C * 4;
    for (int j = threadIdx.x; j < c.D; j += NT) {
        const int ct = !CTS ? CC_NORMAL : (wide ? c.st.ctype[c.col[j]] : c.ctype[j]);
        if (ct == CC_CATEGORICAL) continue;
        const double x = xrow[c.bulk ? c.col[j] : j];
        if (isnan(x)) continue;
        if (ct != CC_NORMAL) {
            if constexpr (CTS) {
                const double sx0 = T.at(CC_F_SX, z, j), sxx0 = T.at(CC_F_SXX, z, j);
                const double t2 = sum_second_term_ni(ct, x, row_aux(c, c.col[j], r));
                const double sx2 = __dadd_rn(__dsub_rn(sx0, x), x), sxx2 = __dadd_rn(__dsub_rn(sxx0, t2), t2);
                if (sx2 != sx0 || sxx2 != sxx0) {
                    T.at(CC_F_SX, z, j) = sx2;
                    T.at(CC_F_SXX, z, j) = sxx2;
                    const NormalCache c2 = count_cache_ni(ct, T.at(CC_F_N, z, j), sx2, hyper(c, wide, ghyp, 0, j), hyper(c, wide, ghyp, 1, j));
                    T.at(CC_F_MN, z, j) = c2.mn; T.at(CC_F_A, z, j) = c2.a; T.at(CC_F_CST, z, j) = c2.cst;
                }
            }
            continue;
        }
        const double sx0 = T.at(CC_F_SX, z, j), sxx0 = T.at(CC_F_SXX, z, j);
        const double xx = __dmul_rn(x, x);
        const double sx2 = __dadd_rn(__dsub_rn(sx0, x), x), sxx2 = __dadd_rn(__dsub_rn(sxx0, xx), xx);
        if (sx2 != sx0 || sxx2 != sxx0) {
            T.at(CC_F_SX, z, j) = sx2;
            T.at(CC_F_SXX, z, j) = sxx2;
            const NormalCache c2 = normal_cache_fast(T.at(CC_F_N, z, j), sx2, sxx2, T.at(CC_F_GN, z, j), hyper(c, wide, ghyp, 0, j),
                                                     hyper(c, wide, ghyp, 1, j), hyper(c, wide, ghyp, 2, j), hyper(c, wide, ghyp, 3, j));
            T.at(CC_F_MN, z, j) = c2.mn; T.at(CC_F_A, z, j) = c2.a; T.at(CC_F_CST, z, j) = c2.cst;
        }
    }
}

// One cell of a row against one candidate cluster (normal or categorical column j of the view).
// Sum-type cell (bernoulli, poisson, exponential, geometric): the cached form against any other cluster (a lone member's
// own cluster without the row is the empty cluster: ke), the statistics with the row removed against its own
// (view.py:416-419), reporting in `drift` whether -= x, += x returns different sums.
template <bool GT>
__device__ __forceinline__ double count_term(const Ctx& c, const Tab<GT>& T, bool wide, const double* ghyp, const double* xrow,
                                             int ct, int j, int k, int ke, bool own, bool dn, int r, int& drift) {
    const double x = xrow[c.bulk ? c.col[j] : j];
    if (isnan(x)) return 0.0;
    if (own) {
        const double sx0 = T.at(CC_F_SX, k, j), sxx0 = T.at(CC_F_SXX, k, j);
        const double t2 = sum_second_term_ni(ct, x, row_aux(c, c.col[j], r));
        const double sx1 = __dsub_rn(sx0, x), sxx1 = __dsub_rn(sxx0, t2);
        if (__dadd_rn(sx1, x) != sx0 || __dadd_rn(sxx1, t2) != sxx0) drift = 1;
        if (dn) return count_pred_raw_ni(ct, x, T.at(CC_F_N, k, j) - 1.0, sx1, hyper(c, wide, ghyp, 0, j), hyper(c, wide, ghyp, 1, j));
    }
    return count_pred_cached_ni(ct, x, T.at(CC_F_MN, ke, j), T.at(CC_F_A, ke, j), T.at(CC_F_CST, ke, j));
}

// CTS: the table has sum-type columns other than normal (compiled out otherwise: their lgamma calls cost the
// all-normal / categorical paths registers and a stack frame)
template <bool GT>
__device__ __forceinline__ double cell_term(const Ctx& c, const Tab<GT>& T, bool wide, bool allnorm, const double* ghyp,
                                            const double* xrow, int j, int k, int ke, bool own, bool dn, bool empty, int r, int& drift) {
    if (!CTS || allnorm) return normal_term<GT>(c, T, wide, ghyp, xrow, j, k, ke, own, dn, drift);
    const int ct = wide ? c.st.ctype[c.col[j]] : c.ctype[j];
    if (ct == CC_NORMAL) return normal_term<GT>(c, T, wide, ghyp, xrow, j, k, ke, own, dn, drift);
    if constexpr (CTS) {
        if (ct != CC_CATEGORICAL) return count_term<GT>(c, T, wide, ghyp, xrow, ct, j, k, ke, own, dn, r, drift);
    }
    return categorical_term<GT>(c, T, wide, ghyp, xrow, j, k, own, empty);
}

// ---------------------------------------------------------------------------------------------
// Scoring of views of normal and categorical columns with shared-memory tables, by column type: the view's normal
// columns come first (c.Dn of them), so a lane walks its normal cells and then its categorical cells in two
// straight-line loops -- no branch on the column type between two cells, which would keep their logs from
// overlapping (measured on configs[2]'s mixed views: ~600 cycles per cell with the branch, one log at a time).
struct ScoreEnv {
    const double *tab0, *cnt0, *hyp0;
    const int *colp, *ctop, *ncatp;
    int fsI, D, Dn, CTv;
    bool bulk;
};
__device__ __forceinline__ ScoreEnv score_env(const Ctx& c, bool allnorm) {
    ScoreEnv e;
    e.tab0 = c.tab; e.fsI = c.Kc * c.D; e.cnt0 = c.tab + CC_NFIELD * e.fsI; e.hyp0 = c.hyp;
    e.colp = c.col; e.ctop = c.cto; e.ncatp = c.ncat;
    e.D = c.D; e.Dn = (!CTS || allnorm) ? c.D : c.Dn; e.CTv = c.CTv; e.bulk = c.bulk != 0;
    return e;
}
// columns j0, j0 + st, ... of the row against cluster slot ke (any cluster but the row's own one with other members;
// emptyc: the empty cluster's record) -- view.py:408-422 through the cached forms
__device__ __forceinline__ double score_plain(const ScoreEnv& e, const double* xrow, int ke, bool emptyc, int j0, int st) {
    const double* pe = e.tab0 + ke * e.D;
    const double* hnu = e.hyp0 + 3 * e.D;
    const double* hr = e.hyp0 + e.D;
    const int fsI = e.fsI;
    int nodrift = 0;
    auto termN = [&](int jj) -> double {
        const double x = xrow[e.bulk ? e.colp[jj] : jj];
        return normal_core(pe + jj, pe + jj, fsI, x, hnu[jj], hr[jj], false, false, nodrift, [](double&, double&) {});
    };
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int j = j0;
    for (; j + 3 * st < e.Dn; j += 4 * st) { a0 += termN(j); a1 += termN(j + st); a2 += termN(j + 2 * st); a3 += termN(j + 3 * st); }
    if (j + st < e.Dn) { a0 += termN(j); a1 += termN(j + st); j += 2 * st; }
    if (j < e.Dn) a2 += termN(j);
    if constexpr (CTS) {
        if (e.Dn < e.D) {
            // categorical.py:144-148: log(alpha + count_x) - log(N + alpha k), the second log cached; the empty cluster's
            // term log(alpha) - log(alpha k) is cached in the hyper row 1 (selected, not branched to)
            const double* cntk = e.cnt0 + ke * e.CTv;
            const double* cstk = pe + CC_F_CST * fsI;
            auto termC = [&](int jj) -> double {
                const double x = xrow[e.bulk ? e.colp[jj] : jj];
                const bool nanx = isnan(x);
                const double cntx = cntk[e.ctop[jj] + (nanx ? 0 : (int)x)];
                const double lg = cc_log_pos(emptyc ? 1.0 : e.hyp0[jj] + cntx) - cstk[jj];
                return nanx ? 0.0 : (emptyc ? hr[jj] : lg);
            };
            j = e.Dn + j0;
            for (; j + 3 * st < e.D; j += 4 * st) { a0 += termC(j); a1 += termC(j + st); a2 += termC(j + 2 * st); a3 += termC(j + 3 * st); }
            if (j + st < e.D) { a0 += termC(j); a1 += termC(j + st); j += 2 * st; }
            if (j < e.D) a2 += termC(j);
        }
    }
    return (a0 + a1) + (a2 + a3);
}
// the same columns against the row's own cluster z WITHOUT the row (view.py:416-419), z not a lone member
__device__ __forceinline__ double score_own(const ScoreEnv& e, const double* xrow, int z, int j0, int st, int& drift) {
    const double* po = e.tab0 + z * e.D;
    const double* hnu = e.hyp0 + 3 * e.D;
    const double* hr = e.hyp0 + e.D;
    const int fsI = e.fsI;
    double b0 = 0.0, b1 = 0.0;
    int j = j0;
    for (; j + st < e.Dn; j += 2 * st) {
        bool ex0, ex1;
        const double x0 = xrow[e.bulk ? e.colp[j] : j], x1 = xrow[e.bulk ? e.colp[j + st] : j + st];
        double t0 = normal_own_fast(po + j, fsI, x0, hnu[j], hr[j], drift, ex0);
        double t1 = normal_own_fast(po + j + st, fsI, x1, hnu[j + st], hr[j + st], drift, ex1);
        if (ex0) t0 = normal_own_exact(po + j, fsI, x0, hnu[j], hr[j], e.hyp0[j], e.hyp0[2 * e.D + j]);
        if (ex1) t1 = normal_own_exact(po + j + st, fsI, x1, hnu[j + st], hr[j + st], e.hyp0[j + st], e.hyp0[2 * e.D + j + st]);
        b0 += t0; b1 += t1;
    }
    if (j < e.Dn) {
        bool ex0;
        const double x0 = xrow[e.bulk ? e.colp[j] : j];
        double t0 = normal_own_fast(po + j, fsI, x0, hnu[j], hr[j], drift, ex0);
        if (ex0) t0 = normal_own_exact(po + j, fsI, x0, hnu[j], hr[j], e.hyp0[j], e.hyp0[2 * e.D + j]);
        b0 += t0;
    }
    if constexpr (CTS) {
        // counts and N one lower (categorical.py:63-66,144-148)
        const double* cntz = e.cnt0 + z * e.CTv;
        for (j = e.Dn + j0; j < e.D; j += st) {
            const double x = xrow[e.bulk ? e.colp[j] : j];
            const bool nanx = isnan(x);
            const double a = e.hyp0[j];
            const double cntx = cntz[e.ctop[j] + (nanx ? 0 : (int)x)];
            const double t = cc_log_pos(nanx ? 1.0 : a + (cntx - 1.0)) - cc_log_pos((po[CC_F_N * fsI + j] - 1.0) + a * e.ncatp[j]);
            b1 += nanx ? 0.0 : t;
        }
    }
    return b0 + b1;
}
// a lone member is scored against the empty cluster; its cluster's sums still go through -= x, += x
__device__ __forceinline__ void score_drift_only(const ScoreEnv& e, const double* xrow, int z, int j0, int st, int& drift) {
    const double* po = e.tab0 + z * e.D;
    for (int j = j0; j < e.Dn; j += st) {
        const double x = xrow[e.bulk ? e.colp[j] : j];
        const double sx0 = po[CC_F_SX * e.fsI + j], sxx0 = po[CC_F_SXX * e.fsI + j];
        const double xx = __dmul_rn(x, x);
        if (!isnan(x) && (__dadd_rn(__dsub_rn(sx0, x), x) != sx0 || __dadd_rn(__dsub_rn(sxx0, xx), xx) != sxx0)) drift = 1;
    }
}

// One candidate of the one-row-at-a-time path (a real call: inlined next to the speculative rounds it costs the
// sweep its register budget -- the per-CTA context went to a 736-byte local-memory frame)
__device__ __noinline__ double score_candidate(const ScoreEnv e, const double* xrow, int k, int ke, bool emptyc, bool own, bool dn,
                                               int j0, int st, int* drift_out) {
    int drift = 0;
    double acc;
    if (dn) acc = score_own(e, xrow, k, j0, st, drift);
    else {
        acc = score_plain(e, xrow, ke, emptyc, j0, st);
        if (own) score_drift_only(e, xrow, k, j0, st, drift);
    }
    *drift_out = drift;
    return acc;
}

// ---------------------------------------------------------------------------------------------
// One pass over row positions [i_begin, n).  Returns the position it stopped at: n when finished
// (*status = ST_OK); otherwise the position of a row whose move cannot be applied here and the
// reason in *status -- ST_MIGRATE: the new cluster does not fit in the shared-memory tables (GT =
// false only; the caller copies the tables to global memory and continues with the same row);
// ST_HANDOVER: it does not fit in the shared-memory candidate arrays (the large-K kernel
// continues); ST_GROW: no free cluster slot below Kcap - 1 (the host grows Kcap and resumes).
// Rows before the returned position are done, the row at it is untouched.
template <int NT, bool GT, bool BIG>
__device__ int sweep(const Ctx& c, int i_begin, int* status, int& lvl) {
    const Tab<GT> T(c);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    const int D = c.D, n = c.n, Kcap = c.st.Kcap;
    const int pslot = T.prior();
    const bool wide = c.wide != 0;
    const bool allnorm = c.allnorm != 0;
    const double* ghyp = c.st.hyp + (size_t)c.s * c.st.C * 4;
    const int lgB = c.lgB, RBm = c.RB - 1;
    const int Wmax = (c.RB - c.B > 1) ? (c.RB - c.B) : 1;      // rows of the ring the window may span (one batch stays in flight)
    const int lgD = ceil_log2(D);
    // one row at a time: scoring groups of G lanes
    int G = 1;
    while (G < D && G < 32) G <<= 1;
    const int gidL = tid / G, glL = tid & (G - 1), ngroupsL = NT / G;
    *status = ST_OK;
    int start = i_begin;
    int round = 0;
    // thread 0: measurement epoch of the window setting (see below)
    int lvl_next = lvl, ep_rounds = -1, ep_rows = 0;
    long long ep_t0 = 0;
    // views whose staying rows do not end a window when their sums drift (see the commit below): the two-pass scoring path
    const bool softd = !GT && !BIG && !wide && (allnorm || (CTS && c.twotypes));
    // one warp on its own (see below): all-normal views of up to 16 columns with shared-memory tables
    const bool solo_ok = !GT && !BIG && allnorm && !wide && D <= 16;
    int solo_back = 2;                 // thread 0: the window setting to return to
    // the whole CTA on one row at a time (the code of the last section below) as one more window setting: views of 32
    // columns and more whose rows move at most steps retire one row per round whatever the window, and a round that
    // spreads the row's (candidate, column) cells over all consumer threads is several times shorter than one that
    // keeps them on the 32 lanes of a row group.  Stretches of ONE_ROWS rows, costed like the other settings.
    const bool one_ok = !BIG && !wide && D >= 32 && c.lvl0 < 0;
    int one_left = 0, one_n = 0, one_moves = 0, one_back = 2;      // (uniform over the CTA; one_back: thread 0)
    long long one_t0 = 0;
    if (tid == 0) { c.scal[44] = lvl; c.scal[45] = lvl; }
    consumer_sync<NT>();
    PH_DECL

    const int lgNB = 31 - __clz(c.NBUF);          // (a power of two)
    auto wait_row = [&](int i) {
        const int b = (i - c.base) >> lgB;
        mbar_wait(&c.bar_full[b & (c.NBUF - 1)], (b >> lgNB) & 1);
    };
    auto release_upto = [&](int pos) {          // thread 0: batches wholly before `pos` go back to the producer
        int rel = c.scal[42];
        const int upto = (pos >= n) ? ((n - c.base + c.B - 1) >> lgB) : ((pos - c.base) >> lgB);
        while (rel < upto) { mbar_arrive(&c.bar_empty[rel & (c.NBUF - 1)]); ++rel; }
        c.scal[42] = rel;
    };

    while (start < n) {
        const int K = c.scal[0];
        bool one_row = true;
        if constexpr (!BIG) one_row = (K + 1 > 32) || (D > c.DL);
        if (!one_row && one_left == 0) {
            // ===================== speculative round =====================
            const int lgk = ceil_log2(K + 1);
            const int lgcmax = min(5 - lgk, lgD);
            const int lvlmax = lgcmax + 3;
            lvl = (c.lvl0 >= 0) ? c.lvl0 : c.scal[44 + (round & 1)];
            if (lvl == LV_ONE && one_ok) {
                one_left = one_n = min((int)ONE_ROWS, n - start);
                one_moves = 0;
                if (tid == 0) one_t0 = clock64();
                continue;                                  // the rows of the stretch take the last section of the loop
            }
            if constexpr (!GT && !BIG) {
                if (lvl == LV_SOLO && solo_ok) {
                    // ===================== one warp on its own =====================
                    // Views whose rows move at every other step retire about one row per round whatever the window:
                    // what counts is the length of a round, and two CTA barriers, the publication of each warp's
                    // verdict and the bookkeeping of the window cost more than the arithmetic.  Warp 0 sweeps a
                    // stretch of rows alone -- 32 / gk rows per round (gk lanes = candidates per row), verdicts by
                    // ballot, a move applied by its own lanes -- while the other warps wait at the barrier below.
                    consumer_sync<NT>();                 // everybody has read K and the level of this round
                    if (warp == 0) {
                        const long long t0 = clock64();
                        int pos = start, code = 0, nmv = 0, rdy = 0;
                        auto release_crossing = [&](int before, int after) {       // lane 0: only when a batch was left behind
                            if ((((after - c.base) >> lgB) != ((before - c.base) >> lgB)) || after >= n) release_upto(after);
                        };
                        const int pend = min(n, start + SOLO_ROWS);
                        const int fsI = c.Kc * D;
                        const double* hyp0 = c.hyp;
                        while (pos < pend) {
                            const int Ks = c.scal[0];
                            if (Ks + 1 > 32) { code = 1; break; }
                            const int lgkS = ceil_log2(Ks + 1), gk = 1 << lgkS;
                            const int Ws = min(min(32 >> lgkS, Wmax), pend - pos);
                            const int rsS = lane >> lgkS, pcS = lane & (gk - 1);
                            const int iS = pos + rsS;
                            const bool arowS = rsS < Ws;
                            PH(ph0)
                            int r = -1, z = 0, nkz = 0, ncand = 0;
                            double uu = 0.0;
                            const double* xrow = c.x;
                            bool singleton = false;
                            if (arowS) {
                                if (iS >= rdy) { wait_row(iS); rdy = c.base + ((((iS - c.base) >> lgB) + 1) << lgB); }
                                const int q = (iS - c.base) & RBm;
                                r = c.rowidx[q];
                                z = c.zrow[q];
                                uu = c.u[q];
                                xrow = c.x + (size_t)q * c.xstride;
                                nkz = c.nk[z];
                                singleton = nkz == 1;
                                ncand = Ks + (singleton ? 0 : 1);
                            }
                            PH(ph1)
                            const bool cand = arowS && pcS < ncand;
                            const int k = (cand && pcS < Ks) ? c.live[pcS] : pslot;
                            const bool own = cand && pcS < Ks && k == z;
                            const int ke = (own && singleton) ? pslot : k;
                            const bool dn = own && !singleton;
                            int drift = 0;
                            double lpv = -INFINITY, wv = 0.0;
                            if (cand) {
                                const double* pe = c.tab + ke * D;
                                const double* po = c.tab + k * D;
                                double acc = 0.0, acc2 = 0.0;
                                int j = 0;
                                for (; j + 1 < D; j += 2) {      // two independent chains: the logs overlap
                                    const double x0 = xrow[c.bulk ? c.col[j] : j], x1 = xrow[c.bulk ? c.col[j + 1] : j + 1];
                                    acc += normal_core_bf(pe + j, po + j, fsI, x0, hyp0[3 * D + j], hyp0[D + j], own, dn, drift,
                                                       [&](double& m, double& ss) { m = hyp0[j]; ss = hyp0[2 * D + j]; });
                                    acc2 += normal_core_bf(pe + j + 1, po + j + 1, fsI, x1, hyp0[3 * D + j + 1], hyp0[D + j + 1], own, dn, drift,
                                                        [&](double& m, double& ss) { m = hyp0[j + 1]; ss = hyp0[2 * D + j + 1]; });
                                }
                                if (j < D) {
                                    const double x0 = xrow[c.bulk ? c.col[j] : j];
                                    acc += normal_core_bf(pe + j, po + j, fsI, x0, hyp0[3 * D + j], hyp0[D + j], own, dn, drift,
                                                       [&](double& m, double& ss) { m = hyp0[j]; ss = hyp0[2 * D + j]; });
                                }
                                lpv = acc + acc2;
                                wv = (pcS >= Ks) ? c.alpha : own ? (singleton ? c.alpha : (double)(nkz - 1)) : (double)c.nk[k];
                            }
                            PH(ph2)
                            // the draw (as below), per row segment of gk lanes
                            float mf = (float)lpv;
                            for (int o = 1; o < gk; o <<= 1) mf = fmaxf(mf, __shfl_xor_sync(0xffffffffu, mf, o));
                            const double dl = lpv - (double)mf;
                            double inc = (dl > -80.0) ? wv * exp(fmax(dl, -80.0)) : ((dl == dl) ? 0.0 : dl);
                            for (int t = 0; t < lgkS; ++t) {
                                const double v = __shfl_up_sync(0xffffffffu, inc, 1 << t);
                                if (pcS >= (1 << t)) inc += v;
                            }
                            const int seg0 = lane & ~(gk - 1);
                            const double tot = __shfl_sync(0xffffffffu, inc, seg0 + max(ncand, 1) - 1);
                            const unsigned segmask = (gk == 32) ? 0xffffffffu : ((1u << gk) - 1u);
                            const unsigned hit = (__ballot_sync(0xffffffffu, cand && inc > uu * tot) >> seg0) & segmask;
                            const unsigned ownb = (__ballot_sync(0xffffffffu, own) >> seg0) & segmask;
                            const unsigned drb = (__ballot_sync(0xffffffffu, own && drift != 0) >> seg0) & segmask;
                            const int ownpos = ownb ? (__ffs(ownb) - 1) : 0;
                            int choice = hit ? (__ffs(hit) - 1) : -1;
                            bool fail = false;
                            if (arowS && choice < 0) { choice = ownpos; fail = true; }
                            int nc = 0, dst = z;
                            if (arowS) {
                                nc = choice >= Ks;
                                dst = nc ? c.scal[2] : c.live[choice];
                            }
                            const bool flagged = arowS && (dst != z || drb != 0);
                            if (fail && pcS == 0) { c.scal[7] = CC_ERR_ARG; c.scal[12] = 1; c.scal[13] = iS; c.scal[14] = Ks; }
                            const unsigned fb = __ballot_sync(0xffffffffu, flagged && pcS == 0);
                            PH(ph3)
                            const int Lf = fb ? (__ffs(fb) - 1) : 0;
                            const int rstar = fb ? (Lf >> lgkS) : (Ws - 1);
                            if (c.trace && cand && rsS <= rstar) {
                                double* tr = c.trace + c.trace_off + (size_t)iS * c.trace_stride;
                                if (pcS == 0) {
                                    tr[0] = (double)r;
                                    tr[1] = (double)ncand;
                                    tr[2] = (double)((choice < Ks) ? c.g_cid[c.live[choice]] : (c.g_cid[c.live[Ks - 1]] + 1));
                                }
                                if (pcS + 3 < c.trace_stride) tr[3 + pcS] = lpv + log(wv);
                            }
                            if (!fb) {
                                if (lane == 0) release_crossing(pos, pos + Ws);
                                pos += Ws;
                                PH(ph5)
#if CC_ROWS_PHASE
                                ++ph_rounds;
#endif
                                continue;
                            }
                            const int zs = __shfl_sync(0xffffffffu, z, Lf), zb = __shfl_sync(0xffffffffu, dst, Lf);
                            const int newcI = __shfl_sync(0xffffffffu, nc, Lf), ownposS = __shfl_sync(0xffffffffu, ownpos, Lf);
                            const int moveI = __shfl_sync(0xffffffffu, (int)(dst != z), Lf), nkzS = __shfl_sync(0xffffffffu, nkz, Lf);
                            const int rS = __shfl_sync(0xffffffffu, r, Lf);
                            if (moveI && newcI && (zb >= Kcap - 1 || zb >= c.Ksm - 1 || zb >= c.Kc - 1)) {
                                if (lane == 0) release_crossing(pos, pos + rstar);
                                pos += rstar;                   // the rounds below take this row: grow / hand over / migrate
                                code = 1;
                                break;
                            }
                            const double* xs = c.x + (size_t)((pos + rstar - c.base) & RBm) * c.xstride;
                            PH(ph5)
                            if (moveI) { apply_move<32, false, false>(c, T, rS, zs, zb, newcI != 0, nkzS == 1, ownposS, Ks, nkzS, xs); ++nmv; }
                            else apply_drift<32, false>(c, T, zs, -1, xs);
                            __syncwarp();
                            PH(ph6)
                            if (lane == 0) release_crossing(pos, pos + rstar + 1);
                            pos += rstar + 1;
                            PH(ph7)
#if CC_ROWS_PHASE
                            ++ph_rounds; ++ph_moves;
#endif
                        }
                        if (lane == 0) {
                            // what the stretch cost per row decides whether the next one is swept the same way
                            float* cost = reinterpret_cast<float*>(c.scal + 46);
                            int* age = c.scal + 55;
                            const int done = max(pos - start, 1);
                            const float now = (float)(clock64() - t0) / (float)done;
                            reinterpret_cast<float*>(c.scal)[8] = now;
                            c.scal[9] = 0;
                            int next = LV_SOLO, fresh = -1;
                            for (int q = 0; q < 9; ++q) {
                                age[q] = min(age[q] + 4, 1000);
                                if (cost[q] > 0.f && age[q] <= 48 && (fresh < 0 || cost[q] < cost[fresh])) fresh = q;
                            }
                            if (code == 1) next = solo_back;
                            else if (fresh >= 0) { if (cost[fresh] < 0.97f * now) next = fresh; }
                            else if (8 * nmv < done) next = solo_back;          // rows stopped moving: measure the windows again
                            lvl_next = next;
                            ep_rounds = -1;
                            c.scal[44] = next; c.scal[45] = next;
                            c.scal[35] = pos;
                        }
                    }
                    consumer_sync<NT>();
                    start = c.scal[35];
                    ++round;
                    continue;
                }
            }
            if (lvl > lvlmax) lvl = lvlmax;
            const int lgc = min(lvl, lgcmax);
            const int nwact = max(1, NW >> (lvl - lgc));
            const int lgg = lgk + lgc, g = 1 << lgg, rpw = 32 >> lgg, gc = 1 << lgc;
            const int W = min(min(nwact * rpw, Wmax), n - start);
            const int rs = lane >> lgg, pc = (lane >> lgc) & ((1 << lgk) - 1), cc = lane & (gc - 1);
            const int ridx = warp * rpw + rs;
            const int i = start + ridx;
            const bool arow = ridx < W;
            int* pubw = c.pub + ((round & 1) * 16 + warp) * 8;
            // registers that outlive the barrier (trace of committed rows)
            double lpv = -INFINITY, wv = 0.0;
            int r = -1, z = 0, ncand = 0, choice = 0;
            bool cand = false;
            if (warp < nwact && warp * rpw < W) {
                const double* xrow = c.x;
                double uu = 0.0;
                int nkz = 0;
                bool singleton = false;
                PH(ph0)
                if (arow) {
                    wait_row(i);
                    const int q = (i - c.base) & RBm;
                    r = c.rowidx[q];
                    z = c.zrow[q];
                    uu = c.u[q];
                    xrow = c.x + (size_t)q * c.xstride;
                    nkz = c.nk[z];
                    singleton = nkz == 1;
                    ncand = K + (singleton ? 0 : 1);
                }
                PH(ph1)
                cand = arow && pc < ncand;
                const int k = (cand && pc < K) ? c.live[pc] : pslot;
                const bool own = cand && pc < K && k == z;
                // a singleton's own cluster without the row IS the empty cluster (exactly)
                const int ke = (own && singleton) ? pslot : k;
                const bool dn = own && !singleton;
                int drift = 0;
                double acc = 0.0, acc2 = 0.0;
                if (CTS && !GT && !wide && (allnorm || c.twotypes)) {
                    // The build with the other column types: the same two passes as below, each walking the view's normal
                    // and categorical columns in separate straight-line loops (score_plain / score_own above).
                    const ScoreEnv env = score_env(c, allnorm);
                    if (cand) acc = score_plain(env, xrow, ke, ke == pslot, cc, gc);
                    double accB = 0.0;
                    if (arow) {
                        if (!singleton) accB = score_own(env, xrow, z, lane & (g - 1), g, drift);
                        else score_drift_only(env, xrow, z, lane & (g - 1), g, drift);
                    }
                    for (int o = gc >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                    for (int o = (1 << ceil_log2(min(g, D))) >> 1; o > 0; o >>= 1) accB += __shfl_xor_sync(0xffffffffu, accB, o);
                    accB = __shfl_sync(0xffffffffu, accB, lane & ~(g - 1));
                    if (dn) acc = accB;
                } else if (!CTS && !GT && !wide && allnorm) {
                    // Views of normal and categorical columns, shared-memory tables.  Pass A: every (row, candidate) lane evaluates the plain
                    // Student-t term of its columns -- straight-line code, two independent chains per trip.  Pass B:
                    // the term of the row's own cluster WITHOUT the row (view.py:416-419) is a different formula; instead
                    // of letting one lane per group run it while the others wait at every column, all g lanes of the
                    // row group share the own cluster's columns and the sum replaces the own lane's result.
                    // (everything the loops touch is copied out of the context first: it stays in registers)
                    const int fsI = c.Kc * D;
                    const double* hyp0 = c.hyp;
                    const double* hnu = hyp0 + 3 * D;
                    const double* hr = hyp0 + D;
                    const double* tab0 = c.tab;
                    const double* cnt0 = tab0 + CC_NFIELD * fsI;
                    const int* colp = c.col;
                    const int* ctp = c.ctype;
                    const int* ctop = c.cto;
                    const int* ncatp = c.ncat;
                    const int CTv = c.CTv;
                    const bool bulk = c.bulk != 0;
                    int nodrift = 0;
                    if (cand) {
                        const double* pe = tab0 + ke * D;
                        const bool emptyc = ke == pslot;       // the empty cluster (a fresh one, or a lone member's own without it)
                        auto termA = [&](int jj) -> double {
                            const double x = xrow[bulk ? colp[jj] : jj];
                            if (!CTS || allnorm || ctp[jj] == CC_NORMAL)
                                return normal_core(pe + jj, pe + jj, fsI, x, hnu[jj], hr[jj], false, false, nodrift, [](double&, double&) {});
                            // categorical.py:144-148: log(alpha + count_x) - log(N + alpha k), the second log cached
                            if (isnan(x)) return 0.0;
                            const double a = hyp0[jj];
                            if (emptyc) return hr[jj];                       // cached log(alpha) - log(alpha k)
                            return cc_log_pos(a + (cnt0 + ke * CTv + ctop[jj])[(int)x]) - pe[CC_F_CST * fsI + jj];
                        };
                        int j = cc;
                        double acc3 = 0.0, acc4 = 0.0;
                        for (; j + 3 * gc < D; j += 4 * gc) {          // four independent chains per trip
                            acc += termA(j); acc2 += termA(j + gc); acc3 += termA(j + 2 * gc); acc4 += termA(j + 3 * gc);
                        }
                        for (; j < D; j += gc) acc += termA(j);
                        acc += acc2 + (acc3 + acc4);
                    }
                    // (pass B is issued before pass A's reduction: the two chains are independent and overlap)
                    double accB = 0.0;
                    if (arow) {
                        const double* po = tab0 + z * D;
                        for (int j = lane & (g - 1); j < D; j += g) {
                            const double x = xrow[bulk ? colp[j] : j];
                            if (CTS && !allnorm && ctp[j] != CC_NORMAL) {
                                // own cluster without the row: counts and N one lower (categorical.py:63-66,144-148)
                                if (!singleton && !isnan(x)) {
                                    const double a = hyp0[j];
                                    const double cntx = (cnt0 + z * CTv + ctop[j])[(int)x];
                                    accB += cc_log_pos(a + (cntx - 1.0)) - cc_log_pos((po[CC_F_N * fsI + j] - 1.0) + a * ncatp[j]);
                                }
                            } else if (!singleton) {
                                accB += normal_core(po + j, po + j, fsI, x, hnu[j], hr[j], true, true, drift,
                                                    [&](double& m, double& ss) { m = hyp0[j]; ss = hyp0[2 * D + j]; });
                            } else {
                                // a lone member is scored against the empty cluster (pass A); its sums still go through -= x, += x
                                const double sx0 = po[CC_F_SX * fsI + j], sxx0 = po[CC_F_SXX * fsI + j];
                                const double xx = __dmul_rn(x, x);
                                if (!isnan(x) && (__dadd_rn(__dsub_rn(sx0, x), x) != sx0 || __dadd_rn(__dsub_rn(sxx0, xx), xx) != sxx0)) drift = 1;
                            }
                        }
                    }
                    for (int o = gc >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                    // only the first min(g, D) lanes of a group hold a part of pass B; the sum goes to every lane from lane 0
                    for (int o = (1 << ceil_log2(min(g, D))) >> 1; o > 0; o >>= 1) accB += __shfl_xor_sync(0xffffffffu, accB, o);
                    accB = __shfl_sync(0xffffffffu, accB, lane & ~(g - 1));
                    if (dn) acc = accB;
                    // (drift stays per lane: the group's flag is collected by a ballot below)
                } else {
                if (cand) {
                    // two columns per trip: independent chains, so the two logs overlap
                    int j = cc;
                    for (; j + gc < D; j += 2 * gc) {
                        acc += cell_term<GT>(c, T, wide, allnorm, ghyp, xrow, j, k, ke, own, dn, pc >= K, CTS ? r : -1, drift);
                        acc2 += cell_term<GT>(c, T, wide, allnorm, ghyp, xrow, j + gc, k, ke, own, dn, pc >= K, CTS ? r : -1, drift);
                    }
                    if (j < D) acc += cell_term<GT>(c, T, wide, allnorm, ghyp, xrow, j, k, ke, own, dn, pc >= K, CTS ? r : -1, drift);
                    acc += acc2;
                }
                for (int o = gc >> 1; o > 0; o >>= 1) {
                    acc += __shfl_xor_sync(0xffffffffu, acc, o);
                    drift |= __shfl_xor_sync(0xffffffffu, drift, o);
                }
                if (!own) drift = 0;
                }
                if (cand) {
                    lpv = acc;
                    // crp.py:111-124: n_k, n_k - 1 for the row's own table, alpha for a fresh / re-used singleton table
                    wv = (pc >= K) ? c.alpha : own ? (singleton ? c.alpha : (double)(nkz - 1)) : (double)c.nk[k];
                }
                PH(ph2)
                // ---- the draw: first j with prefix_j > u * total over w_j exp(lp_j - max), per row group ----
                // any common reference point near the maximum will do (it cancels in inc > u * tot): the float-rounded maximum
                float mf = (float)lpv;
                for (int o = gc; o < g; o <<= 1) mf = fmaxf(mf, __shfl_xor_sync(0xffffffffu, mf, o));
                const double dl = lpv - (double)mf;
                // below e^-80 of the maximum a candidate cannot change a 53-bit prefix sum; keeping the
                // argument in range also keeps exp on its fast path
                double inc = (dl > -80.0) ? wv * exp(fmax(dl, -80.0)) : ((dl == dl) ? 0.0 : dl);
                for (int t = 0; t < lgk; ++t) {
                    const double v = __shfl_up_sync(0xffffffffu, inc, gc << t);
                    if (pc >= (1 << t)) inc += v;
                }
                const int seg0 = lane & ~(g - 1);
                const double tot = __shfl_sync(0xffffffffu, inc, seg0 + (max(ncand, 1) - 1) * gc);
                const unsigned segmask = (g == 32) ? 0xffffffffu : ((1u << g) - 1u);
                const unsigned hit = (__ballot_sync(0xffffffffu, cand && inc > uu * tot) >> seg0) & segmask;
                const unsigned ownb = (__ballot_sync(0xffffffffu, own) >> seg0) & segmask;
                const unsigned drb = (__ballot_sync(0xffffffffu, drift != 0) >> seg0) & segmask;
                const int ownpos = ownb ? ((__ffs(ownb) - 1) >> lgc) : 0;
                choice = hit ? ((__ffs(hit) - 1) >> lgc) : -1;
                bool fail = false;
                if (arow && choice < 0) { choice = ownpos; fail = true; }      // NaN / all -inf: keep the row, flag the chain
                int nc = 0, dst = z;
                if (arow) {
                    nc = choice >= K;
                    dst = nc ? c.scal[2] : c.live[choice];
                }
                // A row that stays but whose `-= x; += x` moved a sum by an ulp (view.py:416-419): in views of normal (and
                // categorical) columns it does not end the window -- the rows behind it were scored against sums that
                // differ in the last bit, far inside the tolerance of the conditionals -- and the sums are replayed
                // in row order when the window is committed (below).  Elsewhere it is treated like a mover.
                const bool drifted = arow && drb != 0;
                const bool flagged = arow && (dst != z || (!softd && drb != 0));
                const bool leader = (lane & (g - 1)) == 0;
                if (fail && leader) { c.scal[7] = CC_ERR_ARG; c.scal[12] = 1; c.scal[13] = i; c.scal[14] = K; }
                const unsigned fb = __ballot_sync(0xffffffffu, flagged && leader);
                const unsigned db = softd ? __ballot_sync(0xffffffffu, drifted && leader) : 0u;
                if (fb && lane == __ffs(fb) - 1) {
                    pubw[1] = ridx; pubw[2] = z; pubw[3] = dst; pubw[4] = nc; pubw[5] = ownpos;
                    pubw[6] = (dst != z) ? 1 : 2; pubw[7] = nkz;
                }
                // word 0: bit 0 = a flagged row (words 1..7), bits 8.. = 1 + the warp's first row with drifting sums
                if (lane == 0) pubw[0] = (fb ? 1 : 0) | (db ? ((warp * rpw + ((__ffs(db) - 1) >> lgg) + 1) << 8) : 0);
            } else if (lane == 0) pubw[0] = 0;
            PH(ph3)
            consumer_sync<NT>();
            PH(ph4)
            // ---- commit in order: the first flagged row of the window --------------------------------
            const int* pub0 = c.pub + (round & 1) * 16 * 8;
            int wstar = -1, dmin = 0x7fffffff;
            for (int w = 0; w < nwact; ++w) {
                const int v = pub0[w * 8];
                if ((v & 1) && wstar < 0) wstar = w;
                if (v >> 8) dmin = min(dmin, (v >> 8) - 1);
            }
            const int rstar = (wstar >= 0) ? pub0[wstar * 8 + 1] : (W - 1);
            if (c.trace && arow && cand && ridx <= rstar && cc == 0) {
                double* tr = c.trace + c.trace_off + (size_t)i * c.trace_stride;
                if (pc == 0) {
                    tr[0] = (double)r;
                    tr[1] = (double)ncand;
                    tr[2] = (double)((choice < K) ? c.g_cid[c.live[choice]] : (c.g_cid[c.live[K - 1]] + 1));
                }
                if (pc + 3 < c.trace_stride) tr[3 + pc] = lpv + log(wv);
            }
            if (tid == 0) {
                // ---- the window setting (lanes per row / active warps) follows the measured cost: cycles per retired
                //      row over epochs of EP rounds, hill-climbing over neighbouring levels (scal[46..54] cost per level,
                //      scal[55..63] epochs since it was measured).  Written two rounds ahead (a barrier lies between).
                constexpr int EP = 16;
                ep_rows += rstar + 1;
                if (++ep_rounds == 0) { ep_rows = 0; ep_t0 = clock64(); }          // the round after a decision still ran the old level
                else if (ep_rounds >= EP) {
                    float* cost = reinterpret_cast<float*>(c.scal + 46);
                    int* age = c.scal + 55;
                    const float now = (float)(clock64() - ep_t0) / (float)max(ep_rows, 1);
                    for (int q = 0; q < 9; ++q) age[q] = min(age[q] + 1, 1000);
                    cost[lvl] = now; age[lvl] = 0;
                    const int lo = lvl - 1, hi = lvl + 1;
                    int next = lvl;
                    if (hi <= lvlmax && (cost[hi] == 0.f || age[hi] > 48)) next = hi;
                    else if (lo >= 0 && (cost[lo] == 0.f || age[lo] > 48)) next = lo;
                    else {
                        float best = now * 0.97f;
                        if (lo >= 0 && cost[lo] < best) { best = cost[lo]; next = lo; }
                        if (hi <= lvlmax && cost[hi] < best) { best = cost[hi]; next = hi; }
                    }
                    if (solo_ok && c.lvl0 < 0) {
                        // fewer than three rows retired per round: try (or return to) one warp on its own
                        const float cs = reinterpret_cast<float*>(c.scal)[8];
                        const int as = c.scal[9] = min(c.scal[9] + 1, 1000);
                        if (cs > 0.f && as <= 48) { if (cs < 0.97f * now) { solo_back = lvl; next = LV_SOLO; } }
                        else if (ep_rows < 3 * EP) { solo_back = lvl; next = LV_SOLO; }
                    }
                    if (one_ok) {
                        // fewer than three rows retired per round: try (or return to) the whole CTA on one row
                        const float co = reinterpret_cast<float*>(c.scal)[36];
                        const int ao = c.scal[37] = min(c.scal[37] + 1, 1000);
                        if (co > 0.f && ao <= 48) { if (co < 0.97f * now) { one_back = lvl; next = LV_ONE; } }
                        else if (ep_rows < 3 * EP) { one_back = lvl; next = LV_ONE; }
                    }
                    lvl_next = next;
                    ep_rounds = -1;
                }
                c.scal[44 + (round & 1)] = lvl_next;
            }
            ++round;
#if CC_ROWS_PHASE
            ++ph_rounds;
#endif
            // ---- rows that stay: replay `-= x; += x` on their clusters' sums in row order, from the first row that
            //      reported a change (the rows before it leave the sums as they are).  One thread per (cluster, normal
            //      column); the cached constants follow the sums.
            bool replayed = false;
            if (softd && dmin < ((wstar >= 0) ? rstar : W)) {
                const int lim = (wstar >= 0) ? rstar : W;
                const int fsI = c.Kc * D;
                double* tab0 = c.tab;
                for (int p = tid; p < K * D; p += NT) {
                    const int kpos = p / D, j = p - kpos * D;
                    if (CTS && !allnorm && c.ctype[j] != CC_NORMAL) continue;
                    const int k = c.live[kpos];
                    double* cell = tab0 + k * D + j;
                    const double sx0 = cell[CC_F_SX * fsI], sxx0 = cell[CC_F_SXX * fsI];
                    double sx = sx0, sxx = sxx0;
                    for (int t = dmin; t < lim; ++t) {
                        const int q = (start + t - c.base) & RBm;
                        if (c.zrow[q] != k) continue;
                        const double x = c.x[(size_t)q * c.xstride + (c.bulk ? c.col[j] : j)];
                        if (isnan(x)) continue;
                        const double xx = __dmul_rn(x, x);
                        sx = __dadd_rn(__dsub_rn(sx, x), x);
                        sxx = __dadd_rn(__dsub_rn(sxx, xx), xx);
                    }
                    if (sx != sx0 || sxx != sxx0) {
                        cell[CC_F_SX * fsI] = sx;
                        cell[CC_F_SXX * fsI] = sxx;
                        const NormalCache c2 = normal_cache_fast(cell[CC_F_N * fsI], sx, sxx, cell[CC_F_GN * fsI], c.hyp[j], c.hyp[D + j],
                                                                 c.hyp[2 * D + j], c.hyp[3 * D + j]);
                        cell[CC_F_MN * fsI] = c2.mn; cell[CC_F_A * fsI] = c2.a; cell[CC_F_CST * fsI] = c2.cst;
                    }
                }
                replayed = true;
            }
            if (wstar < 0) {
                start += W;
                if (replayed) consumer_sync<NT>();          // the next round scores against the replayed sums
                if (tid == 0) release_upto(start);
                PH(ph5)
                continue;
            }
            if (replayed) consumer_sync<NT>();              // the mover's old cluster starts from the replayed sums
            const int* pw = pub0 + wstar * 8;
            const int istar = start + rstar;
            const int zs = pw[2], zb = pw[3], ownpos = pw[5], kind = pw[6], nkz = pw[7];
            const bool newc = pw[4] != 0;
            if (kind == 1 && newc) {
                if (zb >= Kcap - 1) { *status = ST_GROW; return istar; }
                if (zb >= c.Ksm - 1) { *status = ST_HANDOVER; return istar; }
                if (!GT && zb >= c.Kc - 1) { *status = ST_MIGRATE; return istar; }
            }
            const double* xs = c.x + (size_t)((istar - c.base) & RBm) * c.xstride;
            PH(ph5)
            if (kind == 1) apply_move<NT, GT, false>(c, T, c.rowidx[(istar - c.base) & RBm], zs, zb, newc, nkz == 1, ownpos, K, nkz, xs);
            else apply_drift<NT, GT>(c, T, zs, CTS ? c.rowidx[(istar - c.base) & RBm] : -1, xs);
            start = istar + 1;
            PH(ph6)
            consumer_sync<NT>();
            if (tid == 0) release_upto(start);       // after the barrier: the mover's cells were read from the ring
            PH(ph7)
#if CC_ROWS_PHASE
            ++ph_moves;
#endif
            continue;
        }

        // ===================== one row at a time =====================
        const int i = start;
        wait_row(i);
        const int q = (i - c.base) & RBm;
        const int r = c.rowidx[q];
        const int z = c.zrow[q];
        const double* xrow = c.x + (size_t)q * c.xstride;
        const int nkz = c.nk[z];
        const bool singleton = nkz == 1;
        const int ncand = K + (singleton ? 0 : 1);
        int* pubL = c.scal + 16 + (i & 1) * 8;
        // ---- step 1: per-candidate data log-likelihood ----
        for (int base = 0; base < ncand; base += ngroupsL) {      // trip count uniform per warp (full-mask shuffles)
            const int pc = base + gidL;
            const bool active = pc < ncand && gidL < ngroupsL;
            const int k = (active && pc < K) ? c.live[pc] : pslot;
            const bool own = active && (pc < K) && (k == z);
            const int ke = (own && singleton) ? pslot : k;
            const bool dn = own && !singleton;
            int drift = 0;
            double acc = 0.0, acc2 = 0.0;
            if (CTS && active && !GT && !wide && (allnorm || c.twotypes)) {
                // by column type, up to four cells in flight (score_plain / score_own above)
                acc = score_candidate(score_env(c, allnorm), xrow, k, ke, ke == pslot, own, dn, glL, G, &drift);
            } else if (active) {
                int j = glL;
                for (; j + G < D; j += 2 * G) {
                    acc += cell_term<GT>(c, T, wide, allnorm, ghyp, xrow, j, k, ke, own, dn, pc >= K, CTS ? r : -1, drift);
                    acc2 += cell_term<GT>(c, T, wide, allnorm, ghyp, xrow, j + G, k, ke, own, dn, pc >= K, CTS ? r : -1, drift);
                }
                if (j < D) acc += cell_term<GT>(c, T, wide, allnorm, ghyp, xrow, j, k, ke, own, dn, pc >= K, CTS ? r : -1, drift);
                acc += acc2;
            }
            for (int o = G >> 1; o > 0; o >>= 1) {
                acc += __shfl_xor_sync(0xffffffffu, acc, o);
                drift |= __shfl_xor_sync(0xffffffffu, drift, o);
            }
            if (glL == 0 && active) {
                c.lp[pc] = acc;
                c.w[pc] = (pc >= K) ? c.alpha : own ? (singleton ? c.alpha : (double)(nkz - 1)) : (double)c.nk[k];
                if (own) { pubL[0] = pc; pubL[1] = drift; }
            }
        }
        consumer_sync<NT>();
        // ---- step 2: the draw ----
        const double uu = c.u[q];
        if (!BIG) {
            if (tid < 32) {
                int choice = -1;
                if (ncand <= 32) {
                    const double lpv = (lane < ncand) ? c.lp[lane] : -INFINITY;
                    const double wv = (lane < ncand) ? c.w[lane] : 0.0;
                    float mf = (float)lpv;
                    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(mf) : "f"(mf));
                    const double dl = lpv - (double)mf;
                    double inc = (dl > -80.0) ? wv * exp(fmax(dl, -80.0)) : ((dl == dl) ? 0.0 : dl);
                    inc = warp_incl_scan(inc, lane);
                    const double tot = __shfl_sync(0xffffffffu, inc, ncand - 1);
                    const unsigned hit = __ballot_sync(0xffffffffu, lane < ncand && inc > uu * tot);
                    choice = hit ? (__ffs(hit) - 1) : -1;
                } else {
                    double mx = -INFINITY;
                    for (int pc = lane; pc < ncand; pc += 32) mx = fmax(mx, c.lp[pc]);
                    mx = warp_max(mx);
                    double carry = 0.0;
                    for (int base = 0; base < ncand; base += 32) {
                        const int pc = base + lane;
                        const double e = (pc < ncand) ? c.w[pc] * exp(c.lp[pc] - mx) : 0.0;
                        carry += __shfl_sync(0xffffffffu, warp_incl_scan(e, lane), 31);
                    }
                    const double tot = carry;
                    carry = 0.0;
                    for (int base = 0; base < ncand && choice < 0; base += 32) {
                        const int pc = base + lane;
                        const double e = (pc < ncand) ? c.w[pc] * exp(c.lp[pc] - mx) : 0.0;
                        const double inc = warp_incl_scan(e, lane) + carry;
                        const unsigned hit = __ballot_sync(0xffffffffu, pc < ncand && inc > uu * tot);
                        if (hit) choice = base + __ffs(hit) - 1;
                        carry = __shfl_sync(0xffffffffu, inc, 31);
                    }
                }
                if (choice < 0) {                       // NaN / all -inf: keep the row, flag the chain
                    choice = pubL[0];
                    if (lane == 0) { c.scal[7] = CC_ERR_ARG; c.scal[12] = 2; c.scal[13] = i; c.scal[14] = K; }
                }
                if (c.trace && lane == 0) {
                    double* tr = c.trace + c.trace_off + (size_t)i * c.trace_stride;
                    tr[0] = (double)r;
                    tr[1] = (double)ncand;
                    tr[2] = (double)((choice < K) ? c.g_cid[c.live[choice]] : (c.g_cid[c.live[K - 1]] + 1));
                    for (int pc = 0; pc < ncand && pc + 3 < c.trace_stride; ++pc) tr[3 + pc] = c.lp[pc] + log(c.w[pc]);
                }
                if (lane == 0) {
                    const int nc = choice >= K;
                    pubL[2] = choice; pubL[3] = nc ? c.scal[2] : c.live[choice]; pubL[4] = nc;
                }
            }
        } else {
            // block-wide: maximum, sum of w exp(lp - max) by contiguous chunks, first prefix above u * total
            __shared__ double s_red[32];
            __shared__ double s_tot;
            double mx = -INFINITY;
            for (int pc = tid; pc < ncand; pc += NT) mx = fmax(mx, c.lp[pc]);
            mx = warp_max(mx);
            if (lane == 0) s_red[warp] = mx;
            consumer_sync<NT>();
            mx = -INFINITY;
            for (int w2 = 0; w2 < NW; ++w2) mx = fmax(mx, s_red[w2]);
            if (c.trace) {
                double* tr = c.trace + c.trace_off + (size_t)i * c.trace_stride;
                for (int pc = tid; pc < ncand && pc + 3 < c.trace_stride; pc += NT) tr[3 + pc] = c.lp[pc] + log(c.w[pc]);
            }
            consumer_sync<NT>();
            const int cpt = (ncand + NT - 1) / NT;
            const int p0 = min(tid * cpt, ncand), p1 = min(p0 + cpt, ncand);
            double ls = 0.0;
            for (int pc = p0; pc < p1; ++pc) {
                const double e = c.w[pc] * exp(c.lp[pc] - mx);
                c.w[pc] = e;
                ls += e;
            }
            const double incw = warp_incl_scan(ls, lane);
            if (lane == 31) s_red[warp] = incw;
            if (tid == 0) c.scal[40] = 0x7fffffff;
            consumer_sync<NT>();
            double pre = incw - ls, tot = 0.0;
            for (int w2 = 0; w2 < NW; ++w2) { if (w2 < warp) pre += s_red[w2]; tot += s_red[w2]; }
            if (tid == 0) s_tot = tot;
            const double thr = uu * tot;
            double run = pre;
            for (int pc = p0; pc < p1; ++pc) {
                run += c.w[pc];
                if (run > thr) { atomicMin(&c.scal[40], pc); break; }
            }
            consumer_sync<NT>();
            if (tid == 0) {
                int choice = c.scal[40];
                if (choice == 0x7fffffff || !(s_tot == s_tot)) {
                    if (c.scal[7] == 0) {          // remember the first failure (debug probes read it from st.seq)
                        c.scal[12] = (choice == 0x7fffffff) ? 3 : 4; c.scal[13] = i; c.scal[14] = K;
                        int bad = -1;
                        for (int pc = 0; pc < ncand; ++pc) if (!(c.lp[pc] == c.lp[pc]) || !(c.w[pc] == c.w[pc])) { bad = pc; break; }
                        c.scal[15] = bad;
                        c.scal[32] = (bad >= 0 && bad < K) ? c.live[bad] : -1;
                        c.scal[33] = (int)(uu * 1e9);
                        c.scal[34] = (s_tot == s_tot) ? 1 : 0;
                    }
                    choice = pubL[0]; c.scal[7] = CC_ERR_ARG;
                }
                if (c.trace) {
                    double* tr = c.trace + c.trace_off + (size_t)i * c.trace_stride;
                    tr[0] = (double)r;
                    tr[1] = (double)ncand;
                    tr[2] = (double)((choice < K) ? c.g_cid[c.live[choice]] : (c.g_cid[c.live[K - 1]] + 1));
                }
                const int nc = choice >= K;
                pubL[2] = choice; pubL[3] = nc ? c.scal[2] : c.live[choice]; pubL[4] = nc;
            }
        }
        consumer_sync<NT>();
        // ---- step 3: apply ----
        const int ownpos = pubL[0];
        const int zb = pubL[3];
        const bool newc = pubL[4] != 0;
        const bool drifted = pubL[1] != 0;
        if (zb != z) {
            if (newc) {
                if (zb >= Kcap - 1) { *status = ST_GROW; return i; }
                if (!BIG && zb >= c.Ksm - 1) { *status = ST_HANDOVER; return i; }
                if (!GT && zb >= c.Kc - 1) { *status = ST_MIGRATE; return i; }
            }
            apply_move<NT, GT, BIG>(c, T, r, z, zb, newc, singleton, ownpos, K, nkz, xrow);
        } else if (drifted) {
            apply_drift<NT, GT>(c, T, z, r, xrow);
        }
        start = i + 1;
        if (zb != z || drifted) consumer_sync<NT>();
        if (tid == 0) release_upto(start);
        if (one_left > 0) {
            // a stretch chosen as a window setting: at its end thread 0 compares its cost per row with the measured windows
            one_moves += (zb != z);
            if (--one_left == 0 || start >= n) {
                one_left = 0;
                if (tid == 0) {
                    float* cost = reinterpret_cast<float*>(c.scal + 46);
                    int* age = c.scal + 55;
                    const float now = (float)(clock64() - one_t0) / (float)max(one_n, 1);
                    reinterpret_cast<float*>(c.scal)[36] = now;
                    c.scal[37] = 0;
                    int next = LV_ONE, fresh = -1;
                    for (int q = 0; q < 9; ++q) {
                        age[q] = min(age[q] + 4, 1000);
                        if (cost[q] > 0.f && age[q] <= 48 && (fresh < 0 || cost[q] < cost[fresh])) fresh = q;
                    }
                    if (K + 1 > 32) next = one_back;
                    else if (fresh >= 0) { if (cost[fresh] < 0.97f * now) next = fresh; }
                    else if (8 * one_moves < one_n) next = one_back;          // rows stopped moving: measure the windows again
                    lvl_next = next;
                    ep_rounds = -1;
                    c.scal[44] = next; c.scal[45] = next;
                }
                consumer_sync<NT>();
            }
        }
    }
#if CC_ROWS_PHASE
    if (tid == 0) {
        int32_t* q = c.st.seq + c.sv * c.st.R + 16;
        q[0] += (int)(ph0 >> 6); q[1] += (int)(ph1 >> 6); q[2] += (int)(ph2 >> 6); q[3] += (int)(ph3 >> 6); q[4] += (int)(ph4 >> 6);
        q[5] += (int)(ph5 >> 6); q[6] += (int)(ph6 >> 6); q[7] += (int)(ph7 >> 6); q[8] += ph_rounds; q[9] += ph_moves;
    }
#endif
    return n;
}

// BIG: the large-K kernel (candidate arrays in global memory, sweeps handed over through rows_prog)
template <int NT, int MINB, bool BIG>
__global__ void __launch_bounds__(NT + 32, MINB) rows_kernel(const RowsParams p) {
    const cc_state& st = p.st;
    const cc_row_work wk = p.work[blockIdx.x];
    const int s = wk.chain, v = wk.view, n = wk.n;
    const int R = st.R, C = st.C, Cp = st.Cp, Kcap = st.Kcap, CT = st.CT;
    const int tid = threadIdx.x;
    const bool producer = tid >= NT;
    const size_t sv = sv_index(st, s, v);
    int32_t* prog = st.rows_prog + sv * 4;
    // where the sweep starts
    int i0 = 0, nm0 = 0;
    if (BIG) {
        if (prog[2] != CC_ROWS_HANDOVER) return;
        i0 = prog[0]; nm0 = prog[1];
    } else if (wk.resume) {
        if (prog[2] == CC_ROWS_DONE) return;
        i0 = prog[0]; nm0 = prog[1];
    }

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const long long t_start = clock64();
    Ctx c;
    c.st = st; c.s = s; c.v = v; c.n = n; c.sv = sv; c.base = i0;
    c.Ksm = BIG ? Kcap : p.Ksm; c.DL = p.DL; c.lvl0 = p.lvl0;
    const int Ksm = c.Ksm;
    // ---- fixed shared region ----------------------------------------------------
    c.bar_full = reinterpret_cast<uint64_t*>(smem_raw);               // [8]
    c.bar_empty = c.bar_full + 8;                                     // [8]
    c.scal = reinterpret_cast<int*>(c.bar_empty + 8);                 // [64]
    c.pub = c.scal + 64;                                              // [2][16][8]
    int* after_pub = c.pub + 256;
    const int32_t* zv = st.zv + (size_t)s * C;
    int32_t* g_live = st.live + sv * Kcap;
    int32_t* g_nk = st.nk + sv * Kcap;
    c.g_cid = st.cid + sv * Kcap;
    if (BIG) {
        c.live = g_live; c.nk = g_nk;
        c.lp = p.big_lp + (size_t)blockIdx.x * 2 * (Kcap + 1);
        c.w = c.lp + (Kcap + 1);
    } else {
        c.live = after_pub;                                           // [Ksm]
        c.nk = c.live + Ksm;                                          // [Ksm]
        after_pub = c.nk + Ksm;                                       // 2 * Ksm ints: the doubles below stay 8-byte aligned
        c.lp = reinterpret_cast<double*>(after_pub);                  // [Ksm+1]
        c.w = c.lp + (Ksm + 1);                                       // [Ksm+1]
        after_pub = reinterpret_cast<int*>(c.w + (Ksm + 1));
    }
    __shared__ int s_D_tmp;
    if (tid == 0) {
        int D = 0;
        for (int q = 0; q < C; ++q) D += (zv[q] == v);
        s_D_tmp = D;
        for (int b = 0; b < 8; ++b) { mbar_init(&c.bar_full[b], 33); mbar_init(&c.bar_empty[b], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int D = s_D_tmp;
    c.D = D;
    c.wide = p.wide;
    // the staging ring of this view (rows_stage() on the host sizes the launch with the same rule)
    const RowsStage stg = rows_stage(Cp, C, D, p.bulk);
    c.bulk = stg.bulk; c.B = stg.B; c.lgB = 31 - __clz(stg.B); c.NBUF = stg.NBUF; c.RB = stg.B * stg.NBUF;
    const int RB = c.RB;
    c.hyp = reinterpret_cast<double*>(after_pub);                     // [4][D] (not in wide mode)
    c.u = c.hyp + (p.wide ? 0 : 4 * D);                               // [RB]
    c.col = reinterpret_cast<int*>(c.u + RB);                         // [D]
    c.cto = c.col + D;                                                // [D]   (these three not in wide mode)
    c.ctype = c.cto + (p.wide ? 0 : D);                               // [D]
    c.ncat = c.ctype + (p.wide ? 0 : D);                              // [D]
    c.rowidx = c.ncat + (p.wide ? 0 : D);                             // [RB]
    c.zrow = c.rowidx + RB;                                           // [RB]
    size_t off = (size_t)((unsigned char*)(c.zrow + RB) - smem_raw);
    off = (off + 15) & ~(size_t)15;
    c.xstride = c.bulk ? Cp : ((D + 1) & ~1);
    c.x = reinterpret_cast<double*>(smem_raw + off);                  // [RB][xstride]
    off += (size_t)RB * c.xstride * sizeof(double);
    c.tab = reinterpret_cast<double*>(smem_raw + off);
    const size_t tab_avail = (p.smem_bytes > (int)off) ? (size_t)p.smem_bytes - off : 0;

    if (tid == 0) {
        // the view's columns, normal ones first: lanes that walk the columns side by side then evaluate the same
        // type in the same trip (a warp pays for every branch any of its lanes takes)
        int j = 0, ct = 0, alln = 1, two = 1;
        for (int pass = 0; pass < 2; ++pass) {
            if (pass == 1) c.scal[1] = j;                  // the normal columns come first
            for (int q = 0; q < C; ++q)
                if (zv[q] == v && ((st.ctype[q] == CC_NORMAL) == (pass == 0))) {
                    c.col[j] = q;
                    if (st.ctype[q] != CC_NORMAL) alln = 0;
                    if (st.ctype[q] > CC_CATEGORICAL) two = 0;
                    if (!p.wide) {
                        c.cto[j] = ct;
                        c.ctype[j] = st.ctype[q];
                        c.ncat[j] = st.ncat[q];
                    }
                    ct += st.ncat[q];
                    ++j;
                }
        }
        c.scal[6] = ct;
        c.scal[11] = alln;
        c.scal[10] = two;
        c.scal[3] = nm0;
        c.scal[7] = 0;
        c.scal[42] = 0;
        c.scal[43] = 0;
        for (int q = 46; q < 64; ++q) c.scal[q] = 0;
        c.scal[8] = c.scal[9] = c.scal[36] = c.scal[37] = 0;      // measured cost / age of the solo and one-row settings
    }
    const int K0 = st.nclu[sv];
    c.K0 = K0;
    __syncthreads();                 // c.col[] (thread 0) is read by everybody below
    if (!BIG) {
        // the candidate arrays must hold every live slot and leave the empty-cluster position free
        __shared__ int s_fit;
        if (tid == 0) s_fit = 1;
        __syncthreads();
        bool bad = K0 + 1 > Ksm - 1;
        for (int i = tid; i < K0 && !bad; i += blockDim.x) if (g_live[i] >= Ksm - 1) bad = true;
        if (bad) s_fit = 0;
        __syncthreads();
        if (!s_fit) {
            if (tid == 0) { prog[0] = i0; prog[1] = nm0; prog[2] = CC_ROWS_HANDOVER; }
            return;
        }
        for (int i = tid; i < Ksm; i += blockDim.x) {
            c.nk[i] = 0;
            c.live[i] = (i < K0) ? g_live[i] : -1;
        }
        __syncthreads();
        for (int i = tid; i < K0; i += blockDim.x) c.nk[c.live[i]] = g_nk[c.live[i]];
    }
    for (int e = tid; !p.wide && e < 4 * D; e += blockDim.x) {
        const int f = e / D, j = e - f * D;
        double hv = st.hyp[((size_t)s * C + c.col[j]) * 4 + f];
        // a categorical column uses row 0 (alpha) only: row 1 caches its empty-cluster term log(alpha) - log(alpha k)
        // (categorical.py:144-148 with no counts)
        if (CTS && f == 1 && c.ctype[j] == CC_CATEGORICAL) {
            const double a = st.hyp[((size_t)s * C + c.col[j]) * 4];
            hv = log(a) - log(a * c.ncat[j]);
        }
        c.hyp[e] = hv;
    }
    __syncthreads();
    const int CTv = c.scal[6];
    c.CTv = CTv;
    c.allnorm = c.scal[11];
    c.twotypes = c.scal[10];
    c.Dn = c.scal[1];
    // shared-table capacity: Kc cluster slots (the empty-cluster record sits at slot Kc-1)
    const size_t per_cluster = (size_t)(CC_NFIELD * D + CTv) * sizeof(double);
    int Kc = (int)(tab_avail / (per_cluster ? per_cluster : 1));
    if (Kc > Ksm) Kc = Ksm;
    c.Kc = Kc;
    if (BIG) {
        // lowest free slot (parallel search)
        if (tid == 0) c.scal[40] = Kcap - 1;
        __syncthreads();
        for (int b0 = 0; b0 < Kcap - 1; b0 += blockDim.x * 4) {
            for (int t = 0; t < 4; ++t) {
                const int q = b0 + t * blockDim.x + tid;
                if (q < Kcap - 1 && c.nk[q] == 0) { atomicMin(&c.scal[40], q); break; }
            }
            __syncthreads();
            const int found = c.scal[40];
            __syncthreads();
            if (found < Kcap - 1) break;
        }
        if (tid == 0) { c.scal[0] = K0; c.scal[2] = c.scal[40]; c.scal[5] = 1; }
    } else if (tid == 0) {
        int maxslot = -1;
        for (int i = 0; i < K0; ++i) maxslot = max(maxslot, c.live[i]);
        int fh = 0;
        while (fh < Ksm - 1 && c.nk[fh] != 0) ++fh;
        c.scal[0] = K0;
        c.scal[2] = fh;
        c.scal[5] = (p.wide || Kc < 3 || maxslot + 2 > Kc - 1) ? 1 : 0;
    }
    __syncthreads();
    const bool start_global = c.scal[5] != 0;

    // load the shared tables -----------------------------------------------------------
    if (!start_global) {
        for (int e = tid; e < (K0 + 1) * D; e += blockDim.x) {
            const int pp = e / D, j = e - pp * D;
            const int gk = (pp < K0) ? c.live[pp] : (Kcap - 1);
            const int sk = (pp < K0) ? gk : (Kc - 1);
            const int q = c.col[j];
#pragma unroll
            for (int f = 0; f < CC_NFIELD; ++f)
                c.tab[(f * Kc + sk) * D + j] = stat_field(st, s, f)[(size_t)gk * C + q];
            if (c.ctype[j] == CC_CATEGORICAL && pp < K0) {
                const double* src = st.cnt + ((size_t)s * Kcap + gk) * CT + st.catoff[q];
                double* dst = c.tab + CC_NFIELD * Kc * D + sk * CTv + c.cto[j];
                for (int qq = 0; qq < c.ncat[j]; ++qq) dst[qq] = src[qq];
            }
        }
    }
    __syncthreads();

    c.alpha = st.view_alpha[sv];
    c.zr = st.zr + sv * R;
    c.moved = st.moved + sv * R;
    c.movedflag = st.movedflag + sv * R;
    c.trace = (p.trace && wk.trace_off >= 0) ? p.trace : nullptr;
    c.trace_off = wk.trace_off;
    c.trace_stride = p.trace_stride;
    const int32_t* order = st.order + sv * R;
    const Philox rng(p.seed);
    const int B = c.B;
    const int nbatch = (n - i0 + B - 1) / B;

    if (producer) {
        // =========================== producer warp ===========================
        const int lane = tid - NT;
        uint32_t bits = 1;
        while ((1u << bits) < (uint32_t)n && bits < 31) ++bits;
        const uint4 pkey = rng(p.counter, stream_id(2, s, v));
        const uint64_t ustream = stream_id(1, s, v);
        volatile int* stopflag = c.scal + 43;
        int b = 0;
        bool stopped = false;
        for (; b < nbatch; ++b) {
            const int buf = b & (c.NBUF - 1);
            if (b >= c.NBUF) {
                const uint32_t parity = ((b / c.NBUF) - 1) & 1;
                uint32_t done = 0;
                while (true) {
                    // suspended by the hardware until the phase completes or ~50 us pass (no polling instructions)
                    asm volatile(
                        "{\n"
                        ".reg .pred p;\n"
                        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
                        "selp.u32 %0, 1, 0, p;\n"
                        "}\n"
                        : "=r"(done)
                        : "r"(smem_u32(&c.bar_empty[buf])), "r"(parity), "r"(50000u)
                        : "memory");
                    if (done) break;
                    if (*stopflag) { stopped = true; break; }
                }
                if (stopped) break;
            }
            if (*stopflag) break;
            const int i = i0 + b * B + lane;
            const bool act = lane < B && i < n;
            const int qslot = buf * B + lane;
            int r = 0;
            if (act) {
                int idx;
                if (wk.perm_off >= 0) idx = p.perm[wk.perm_off + i];
                else idx = (int)keyed_perm((uint32_t)i, (uint32_t)n, bits, pkey);
                r = wk.perm_is_index ? order[idx] : idx;
                c.rowidx[qslot] = r;
                double uu;
                if (wk.u_off >= 0) uu = p.u[wk.u_off + i];
                else { const uint4 w4 = rng(p.counter * 0x100000000ull + (uint64_t)i, ustream); uu = u01_from_words(w4.x, w4.y); }
                c.u[qslot] = uu;
            }
            const int nrows = min(B, n - i0 - b * B);
            __syncwarp();
            if (lane == 0) mbar_arrive_expect_tx(&c.bar_full[buf], c.bulk ? (uint32_t)(nrows * Cp * 8) : 0u);
            __syncwarp();
            if (act) {
                cp_async_4(&c.zrow[qslot], &c.zr[r]);
                double* dst = c.x + (size_t)qslot * c.xstride;
                const double* src = st.Xrm + (size_t)r * Cp;
                if (c.bulk) bulk_g2s(dst, src, (uint32_t)(Cp * 8), &c.bar_full[buf]);
                else for (int j = 0; j < D; ++j) cp_async_8(&dst[j], &src[c.col[j]]);
            }
            cp_async_mbar_arrive_noinc(&c.bar_full[buf]);
        }
        // a sweep that stops early must not leave copies in flight into this CTA's shared memory
        if (b < nbatch || *stopflag) {
            const int lo = max(0, b - c.NBUF);
            for (int bb = lo; bb < b; ++bb) mbar_wait_sleep(&c.bar_full[bb & (c.NBUF - 1)], (bb / c.NBUF) & 1);
        }
        return;
    }

    // ============================== consumers ===============================
    int pos = i0, status = ST_OK;
    int lvl = 2;
    bool in_global = start_global;
    if (BIG) {
        pos = sweep<NT, true, true>(c, pos, &status, lvl);
    } else {
        if (!start_global) {
            pos = sweep<NT, false, false>(c, pos, &status, lvl);
            if (status == ST_MIGRATE) {
                // migrate the shared tables to global memory (L2) and continue there with the same row
                const int K = c.scal[0];
                for (int e = tid; e < K * D; e += NT) {
                    const int pp = e / D, j = e - pp * D;
                    const int k = c.live[pp], q = c.col[j];
#pragma unroll
                    for (int f = 0; f < CC_NFIELD; ++f)
                        stat_field(st, s, f)[(size_t)k * C + q] = c.tab[(f * Kc + k) * D + j];
                    if (c.ctype[j] == CC_CATEGORICAL) {
                        double* dst = st.cnt + ((size_t)s * Kcap + k) * CT + st.catoff[q];
                        const double* src = c.tab + CC_NFIELD * Kc * D + k * CTv + c.cto[j];
                        for (int qq = 0; qq < c.ncat[j]; ++qq) dst[qq] = src[qq];
                    }
                }
                __threadfence_block();
                consumer_sync<NT>();
                in_global = true;
                status = ST_OK;
                pos = sweep<NT, true, false>(c, pos, &status, lvl);
            }
        } else {
            pos = sweep<NT, true, false>(c, pos, &status, lvl);
        }
    }
    if (status != ST_OK && tid == 0) c.scal[43] = 1;       // tell the producer to stop and drain

    // ---- write back ---------------------------------------------------------------
    consumer_sync<NT>();
    const int K = c.scal[0];
    if (!in_global) {
        for (int e = tid; e < K * D; e += NT) {
            const int pp = e / D, j = e - pp * D;
            const int k = c.live[pp], q = c.col[j];
#pragma unroll
            for (int f = 0; f < CC_NFIELD; ++f)
                stat_field(st, s, f)[(size_t)k * C + q] = c.tab[(f * Kc + k) * D + j];
            if (c.ctype[j] == CC_CATEGORICAL) {
                double* dst = st.cnt + ((size_t)s * Kcap + k) * CT + st.catoff[q];
                const double* src = c.tab + CC_NFIELD * Kc * D + k * CTv + c.cto[j];
                for (int qq = 0; qq < c.ncat[j]; ++qq) dst[qq] = src[qq];
            }
        }
    }
    if (!BIG) {
        for (int i = tid; i < Kcap; i += NT) {
            g_live[i] = (i < K) ? c.live[i] : -1;
            g_nk[i] = (i < Ksm) ? c.nk[i] : 0;
        }
    }
    const int nmoved = c.scal[3];
    if (tid == 0) {
        st.nclu[sv] = K;
        if (c.scal[7]) st.err[s] = c.scal[7];
        prog[0] = pos; prog[1] = nmoved; prog[2] = status;
        prog[3] = (int)min((clock64() - t_start) >> 10, (long long)0x7fffffff) + ((BIG || wk.resume) ? prog[3] : 0);
        if (p.prof) {
            int32_t* q = st.seq + sv * R;
            q[0] = (int)((clock64() - t_start) >> 10); q[1] = D; q[2] = K; q[3] = nmoved;
            q[4] = (BIG ? 8 : 0) + (start_global ? 1 : (in_global ? 2 : 0));      // kernel; table mode: shared / global / migrated mid-sweep
            q[5] = lvl;
            if (c.scal[7]) { q[6] = c.scal[12]; q[7] = c.scal[13]; q[8] = c.scal[14]; q[9] = D; q[10] = c.scal[15]; q[11] = c.scal[32]; q[12] = c.scal[33]; q[13] = c.scal[34]; }      // where a draw failed
        }
    }
    if (status != ST_OK) return;
    // ---- member order of the view's CRP: re-seated rows move to the end, in the
    //      order they were re-seated (OrderedDict pop + re-insert, crp.py:52,55) -----
    if (nmoved > 0) {
        const int lane = tid & 31;
        int32_t* ord = st.order + sv * R;
        __shared__ int s_wsum[32];
        __shared__ int s_base;
        if (tid == 0) s_base = 0;
        consumer_sync<NT>();
        for (int i0_ = 0; i0_ < R; i0_ += NT) {
            const int i = i0_ + tid;
            int rr = -1, keep = 0;
            if (i < R) { rr = ord[i]; keep = !c.movedflag[rr]; }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            const int wpre = __popc(bal & ((1u << lane) - 1u));
            if (lane == 0) s_wsum[tid >> 5] = __popc(bal);
            consumer_sync<NT>();
            int wbase = 0;
            for (int w = 0; w < (tid >> 5); ++w) wbase += s_wsum[w];
            const int base = s_base;
            consumer_sync<NT>();
            if (keep) ord[base + wbase + wpre] = rr;
            if (tid == NT - 1) s_base = base + wbase + wpre + keep;
            consumer_sync<NT>();
        }
        const int base = s_base;
        for (int i = tid; i < nmoved; i += NT) {
            const int rr = c.moved[i];
            ord[base + i] = rr;
            c.movedflag[rr] = 0;
        }
    }
}

}  // namespace

extern "C" int CC_ROWS_ENTRY(const cc_state* st, int n_work, const cc_row_work* work,
                                  const int32_t* perm, const double* u, double* trace, int trace_stride,
                                  uint64_t seed, uint64_t counter, void* stream_) {
    if (!st || n_work < 0 || (n_work > 0 && !work)) { cc_set_error("cc_transition_rows: bad arguments"); return CC_ERR_ARG; }
    if (n_work == 0) return CC_OK;
    if (st->Kcap < 4) { cc_set_error("cc_transition_rows: Kcap must be at least 4, got %d", st->Kcap); return CC_ERR_ARG; }
    if (!st->rows_prog) { cc_set_error("cc_transition_rows: cc_state.rows_prog is not set"); return CC_ERR_ARG; }
    cudaStream_t stream = (cudaStream_t)stream_;
    cc_prepare_pool();
    for (int i = 0; i < n_work; ++i) {
        const cc_row_work& w = work[i];
        if (w.chain < 0 || w.chain >= st->S || w.view < 0 || w.view >= st->Vcap || w.n < 0 || w.n > st->R) {
            cc_set_error("cc_transition_rows: work item %d out of range", i);
            return CC_ERR_ARG;
        }
        if ((w.perm_off >= 0 && !perm) || (w.u_off >= 0 && !u)) {
            cc_set_error("cc_transition_rows: work item %d refers to a missing buffer", i);
            return CC_ERR_ARG;
        }
    }
    int dev = 0, nsm = 148, max_smem = 0;
    CC_CUDA(cudaGetDevice(&dev));
    CC_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    CC_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));

    // Work items are launched widest view first (the longest sweeps start first when there are more items than
    // resident CTAs).  Each CTA picks its own staging mode: whole rows by bulk copies when the view holds at
    // least a quarter of a short row, per-cell gathers otherwise (rows_stage).
    const int allow_bulk = getenv("CGPM_B200_ROWS_NO_BULK") ? 0 : 1;
    auto width = [&](const cc_row_work& w) -> int { return (w.n_cols > 0 && w.n_cols <= st->C) ? w.n_cols : st->C; };
    std::vector<cc_row_work> sorted(work, work + n_work);
    std::stable_sort(sorted.begin(), sorted.end(), [&](const cc_row_work& a, const cc_row_work& b) {
        if (a.cost_hint != b.cost_hint) return a.cost_hint > b.cost_hint;
        return width(a) > width(b);
    });
    cc_row_work* dwork = nullptr;
    double* big_lp = nullptr;
    CC_CUDA(cudaMallocAsync(&dwork, sizeof(cc_row_work) * n_work, stream));
    CC_CUDA(cudaMemcpyAsync(dwork, sorted.data(), sizeof(cc_row_work) * n_work, cudaMemcpyHostToDevice, stream));
    std::vector<int32_t> ctype_h(st->C);
    CC_CUDA(cudaMemcpyAsync(ctype_h.data(), st->ctype, sizeof(int32_t) * st->C, cudaMemcpyDeviceToHost, stream));
    CC_CUDA(cudaStreamSynchronize(stream));            // `sorted` is a local; n_work is small
#if !CC_ROWS_COUNT
    for (int c = 0; c < st->C; ++c)                    // any column that is not normal: the other build
        if (ctype_h[c] != CC_NORMAL) {
            CC_CUDA(cudaFreeAsync(dwork, stream));
            return cc_transition_rows_counts(st, n_work, work, perm, u, trace, trace_stride, seed, counter, stream_);
        }
#endif

    const int Ksm_env = getenv("CGPM_B200_ROWS_KSM") ? atoi(getenv("CGPM_B200_ROWS_KSM")) : 0;
    const int Ksm = std::max(4, std::min(st->Kcap, Ksm_env > 0 ? Ksm_env : 1024));
    int per_sm = std::max(1, std::min(2, (n_work + nsm - 1) / nsm));       // CTAs per SM wanted
    int rc = CC_OK;
    const int forced_smem = getenv("CGPM_B200_ROWS_SMEM_KB") ? std::max(0, atoi(getenv("CGPM_B200_ROWS_SMEM_KB"))) : 0;
    auto launch = [&](bool big) -> int {
        RowsParams p;
        p.st = *st; p.work = dwork; p.perm = perm; p.u = u; p.trace = trace; p.trace_stride = trace_stride;
        p.seed = seed; p.counter = counter;
        p.prof = getenv("CGPM_B200_ROWS_PROF") ? 1 : 0;
        p.Ksm = Ksm; p.big_lp = big_lp;
        p.DL = getenv("CGPM_B200_ROWS_DL") ? atoi(getenv("CGPM_B200_ROWS_DL")) : 128;
        p.lvl0 = getenv("CGPM_B200_ROWS_LVL") ? atoi(getenv("CGPM_B200_ROWS_LVL")) : -1;
        p.bulk = allow_bulk; p.B = 0; p.NBUF = 0;
        // widest view and largest per-view need (staging ring + per-column metadata) among the work items
        int dmax = 0;
        for (int i = 0; i < n_work; ++i) dmax = std::max(dmax, width(sorted[i]));
        p.wide = (dmax > 1024) ? 1 : 0;
        int vbytes = 0;
        for (int i = 0; i < n_work; ++i) {
            const int d = width(sorted[i]);
            const RowsStage sg = rows_stage(st->Cp, st->C, d, allow_bulk);
            const int RB = sg.B * sg.NBUF;
            vbytes = std::max(vbytes, RB * sg.stagebytes + RB * 16 + d * (p.wide ? 4 : 48));
        }
        const int kbytes = big ? 0 : (Ksm * 8 + 8 + (Ksm + 1) * 16);
        const int fixed = 128 + 256 + 1024 + kbytes + 64 + vbytes + 256;
        int budget;
        if (big) budget = std::min(max_smem, fixed + 64 * 1024);
        else {
            // Two CTAs per SM only if the widest view's tables still fit in half the shared memory with a dozen clusters
            // (+ the empty-cluster record and a slot to grow into): sweeping a wide view against tables in global memory
            // costs about four times the shared-memory sweep (measured on configs[2]), two waves of CTAs at most two.
            if (per_sm > 1 && !forced_smem) {
                const size_t per_cluster = ((size_t)CC_NFIELD * dmax + (size_t)st->CT * dmax / std::max(1, st->C)) * sizeof(double);
                if ((size_t)fixed + 14 * per_cluster > (size_t)((max_smem + 1024) / 2 - 1024 - 256)) per_sm = 1;
            }
            budget = (max_smem + 1024) / per_sm - 1024 - 256;          // 1 KB per-CTA system reservation
        }
        bool forced = false;
        if (forced_smem) { budget = fixed + forced_smem * 1024; forced = true; }   // tests: force small tables
        if (!forced && budget < fixed + 4096) budget = fixed + 4096;
        if (budget > max_smem) budget = max_smem;
        if (fixed > max_smem) { cc_set_error("cc_transition_rows: shared memory need %d exceeds %d", fixed, max_smem); return CC_ERR_ARG; }
        p.smem_bytes = budget;
        if (big) {
            CC_CUDA(cudaFuncSetAttribute(rows_kernel<480, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, budget));
            rows_kernel<480, 1, true><<<n_work, 512, budget, stream>>>(p);
        } else {
            // 7 consumer warps at 128 registers for all-normal tables; the build with the other column types needs more
            // registers per thread: 6 consumer warps (224 threads, 144 registers at two CTAs per SM)
            constexpr int NTC = CTS ? 192 : 224;
            CC_CUDA(cudaFuncSetAttribute(rows_kernel<NTC, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, budget));
            rows_kernel<NTC, 2, false><<<n_work, NTC + 32, budget, stream>>>(p);
        }
        CC_CUDA(cudaGetLastError());
        return CC_OK;
    };
    rc = launch(false);
    // sweeps the first kernel could not hold (or handed over mid-way) continue in the large-K kernel; its CTAs
    // return at once for every other work item
    if (rc == CC_OK && st->Kcap > Ksm) {
        CC_CUDA(cudaMallocAsync(&big_lp, sizeof(double) * 2 * (size_t)(st->Kcap + 1) * n_work, stream));
        rc = launch(true);
        if (big_lp) CC_CUDA(cudaFreeAsync(big_lp, stream));
    }
    if (rc != CC_OK) return rc;
    CC_CUDA(cudaFreeAsync(dwork, stream));
    return CC_OK;
}
