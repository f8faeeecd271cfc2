// Row-reassignment Gibbs sweep for WIDE views (more than 16 columns): the same per-row work as
// rows.cu's general kernel (same scoring routine, same draw, same table updates, so the same
// trajectories), with two rows in flight.  Wide views hold most of a table's evidence, so their
// rows almost never move; their cost is the scoring (a full warp or more per candidate cluster,
// two columns per lane, five shuffle rounds) followed by the draw.  Here warp 0 draws row i while
// the evaluator warps already score row i+1 against the current tables: if row i stays where it
// is and its own cluster's sums did not drift (the common case) those scores are exactly what the
// sequential scan would have computed, otherwise the move (or the drift) is applied first and
// row i+1 is scored again.  The result is always that of the strictly sequential scan
// (View.transition_rows, src/mixtures/view.py:247-252,393-429); a step costs max(score, draw)
// instead of score + draw.  For views whose rows move often the second scoring makes this
// slower than the plain sequence, which is why rows.cu keeps the sequential sweep for them.
#include <cstdlib>
#include "cc_common.cuh"
#include "rows_params.h"

namespace {

template <int NT>
__device__ __forceinline__ void consumer_sync() {
    asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
}

// cached predictive constants for the rows kernel: same formulas as normal_cache, with the two
// divisions replaced by reciprocals (the values feed log-densities only, never the statistics)
__device__ __forceinline__ NormalCache normal_cache_fast(double N, double sx, double sxx, double gN, double m,
                                                         double r, double s, double nu) {
    const double rn = r + N;
    const double irn = __drcp_rn(rn);
    const double mn = (r * m + sx) * irn;
    double sn = __dsub_rn(__dadd_rn(__dadd_rn(s, sxx), __dmul_rn(__dmul_rn(r, m), m)), __dmul_rn(__dmul_rn(rn, mn), mn));
    if (sn == 0.0) sn = s;
    NormalCache c;
    c.mn = mn;
    c.a = rn * __drcp_rn((rn + 1.0) * sn);
    c.cst = gN - 0.5 * ((sn > 0.0) ? cc_log_pos(sn) : log(sn));
    return c;
}

// bijection on [0, 2^bits) from add / odd-multiply / xorshift rounds, walked
// until the value falls below n: a keyed pseudo-random permutation of [0, n).
__device__ __forceinline__ uint32_t keyed_perm(uint32_t i, uint32_t n, uint32_t bits, uint4 key) {
    const uint32_t mask = (bits >= 32) ? 0xFFFFFFFFu : ((1u << bits) - 1u);
    const uint32_t sh = (bits + 1) >> 1;
    uint32_t x = i;
    do {
        x = (x + key.x) & mask; x = (x * 0x9E3779B1u) & mask; x ^= x >> sh;
        x = (x + key.y) & mask; x = (x * 0x85EBCA6Bu) & mask; x ^= x >> sh;
        x = (x + key.z) & mask; x = (x * 0xC2B2AE35u) & mask; x ^= x >> sh;
        x = (x + key.w) & mask; x = (x * 0x27D4EB2Fu) & mask; x ^= x >> sh;
    } while (x >= n);
    return x;
}

// Everything a sweep needs, resolved once per CTA.
struct Ctx {
    cc_state st;
    int s, v, n, D, K0, Kc, CTv, B, bulk, xstride, wide, allnorm;
    size_t sv;
    // shared memory
    uint64_t *bar_full, *bar_empty;
    int *scal;              // [32]: 0 K, 2 free head, 3 nmoved, 5 start in global, 6 CTv, 7 error, 8 all-normal view;
                            // 16 + 8 * (row parity) + {0 own position, 1 drift flag, 2 choice, 3 destination slot, 4 new-cluster flag}
    int *live, *nk;
    double* lp;             // [2][Kcap+1] candidate scores, by row parity
    double* w;              // [2][Kcap+1] CRP weight of each candidate (written with lp)
    int *col, *cto, *ctype, *ncat;
    double* hyp;            // [4][D]
    int *rowidx, *zrow;
    double *u, *x, *tab;
    // global
    int32_t *zr, *moved, *g_cid;
    uint8_t* movedflag;
    double alpha;
    double* trace;
    int trace_stride;
    long long trace_off;
};

// table access, specialised on the address space (GT = tables in global memory)
template <bool GT>
struct Tab {
    const Ctx& c;
    double* f0;        // field 0 of (cluster 0, column 0)
    size_t fs;         // distance between two fields of one (cluster, column) cell
    double* gcnt;
    __device__ __forceinline__ Tab(const Ctx& cx) : c(cx) {
        if (GT) {
            f0 = stat_field(cx.st, cx.s, 0);
            fs = (size_t)cx.st.Kcap * cx.st.C;
            gcnt = cx.st.cnt + (size_t)cx.s * cx.st.Kcap * cx.st.CT;
        } else {
            f0 = cx.tab;
            fs = (size_t)cx.Kc * cx.D;
            gcnt = nullptr;
        }
    }
    // field f of the cell is cell(k, j)[f * fs]
    __device__ __forceinline__ double* cell(int k, int j) const {
        if (GT) return f0 + (size_t)k * c.st.C + c.col[j];
        return f0 + k * c.D + j;
    }
    __device__ __forceinline__ double& at(int f, int k, int j) const { return cell(k, j)[f * fs]; }
    __device__ __forceinline__ double* counts(int k, int j) const {
        if (GT) return gcnt + (size_t)k * c.st.CT + c.st.catoff[c.col[j]];
        return c.tab + CC_NFIELD * c.Kc * c.D + k * c.CTv + c.cto[j];
    }
    __device__ __forceinline__ int prior() const { return GT ? (c.st.Kcap - 1) : (c.Kc - 1); }
};

__device__ __forceinline__ double hyper(const Ctx& c, bool wide, const double* ghyp, int f, int j) {
    return wide ? __ldg(&ghyp[c.col[j] * 4 + f]) : c.hyp[f * c.D + j];
}

// Normal cell of one row against one candidate cluster (rows kernel, step 1).
//   other cluster: cst - b log(1 + a (x-mn)^2)   (Student-t form, cached a, mn, cst)
//   own cluster  : the row is removed first (view.py:416-419).  With sn1 = sn - rn/(rn-1) (x-mn)^2 (the
//     Normal-Gamma downdate; rn, mn, sn, a = rn/((rn+1) sn) of the cluster WITH the row) the predictive
//     G(N-1) - b1 log sn + (b1-1/2) log sn1 becomes cst + G(N-1) - G(N) + (b1-1/2) log(1 - a (rn+1)/(rn-1)
//     (x-mn)^2): the same loads and the same single log as for any other cluster.  Heavy cancellation
//     (arg < 2^-20) takes the reference's posterior update of the decremented sums instead.  A singleton's
//     own cluster without the row IS the empty cluster (ke).  The group also replays `-= x; += x`
//     (view.py:416-419) and reports in `drift` whether the sums come back changed.
// p: the candidate's cell (field f at p[f * fs]); po: the cell of the row's own cluster (k), read only if `own`.
// ms(m, s): loads the two hyperparameters only the rare exact branch needs.
template <typename FS, typename MS>
__device__ __forceinline__ double normal_core(const double* p, const double* po, FS fs, double x, double nu, double rr,
                                              bool own, bool dn, int& drift, MS ms) {
    const double Nf = p[CC_F_N * fs];
    const double zz = x - p[CC_F_MN * fs];
    double sc = p[CC_F_A * fs];
    double bs = p[CC_F_CST * fs];
    double coef = -0.5 * (nu + Nf + 1.0);
    if (own) {
        const double sx0 = po[CC_F_SX * fs], sxx0 = po[CC_F_SXX * fs];
        const double xx = __dmul_rn(x, x);
        const double sx1 = __dsub_rn(sx0, x), sxx1 = __dsub_rn(sxx0, xx);
        if (!isnan(x) && (__dadd_rn(sx1, x) != sx0 || __dadd_rn(sxx1, xx) != sxx0)) drift = 1;
        if (dn) {
            const double rn = rr + Nf;
            sc = -sc * ((rn + 1.0) * __drcp_rn(rn - 1.0));
            coef = 0.5 * (nu + Nf - 1.0);
            bs += p[CC_F_GM1 * fs] - p[CC_F_GN * fs];
            if (!(fma(sc * zz, zz, 1.0) >= 0x1p-20)) {
                double m, ss;
                ms(m, ss);
                const double N1 = Nf - 1.0;
                const double rn1 = rr + N1;
                const double mn1 = (rr * m + sx1) * __drcp_rn(rn1);
                double sn1 = __dsub_rn(__dadd_rn(__dadd_rn(ss, sxx1), __dmul_rn(__dmul_rn(rr, m), m)),
                                       __dmul_rn(__dmul_rn(rn1, mn1), mn1));
                if (sn1 == 0.0) sn1 = ss;
                const double b1 = 0.5 * (nu + N1 + 1.0);
                bs = p[CC_F_GM1 * fs] - b1 * (2.0 * (p[CC_F_GN * fs] - p[CC_F_CST * fs]));
                const double t = bs + coef * log(sn1);
                return isnan(x) ? 0.0 : t;
            }
        }
    }
    const double t = bs + coef * cc_log_pos(fma(sc * zz, zz, 1.0));
    return isnan(x) ? 0.0 : t;                                   // dim.py:130-132
}

template <bool GT>
__device__ __forceinline__ double normal_term(const Ctx& c, const Tab<GT>& T, bool wide, const double* ghyp,
                                              const double* xrow, int j, int k, int ke, bool own, bool dn, int& drift) {
    const double x = xrow[c.bulk ? c.col[j] : j];
    return normal_core(T.cell(ke, j), T.cell(k, j), T.fs, x, hyper(c, wide, ghyp, 3, j), hyper(c, wide, ghyp, 1, j),
                       own, dn, drift, [&](double& m, double& ss) { m = hyper(c, wide, ghyp, 0, j); ss = hyper(c, wide, ghyp, 2, j); });
}

// Categorical cell: log((a + count_x) / (N + a k)), counts without the row for its own cluster.
template <bool GT>
__device__ __forceinline__ double categorical_term(const Ctx& c, const Tab<GT>& T, bool wide, const double* ghyp,
                                                   const double* xrow, int j, int k, bool own, bool empty) {
    const double x = xrow[c.bulk ? c.col[j] : j];
    if (isnan(x)) return 0.0;
    const double a = hyper(c, wide, ghyp, 0, j);
    const int xi = (int)x, kc = wide ? c.st.ncat[c.col[j]] : c.ncat[j];
    if (empty) return cc_log_pos(a) - cc_log_pos(a * kc);
    const double cntx = T.counts(k, j)[xi];
    if (!own) return cc_log_pos(a + cntx) - T.at(CC_F_CST, k, j);      // arguments are positive: branch-free log
    return cc_log_pos(a + (cntx - 1.0)) - cc_log_pos((T.at(CC_F_N, k, j) - 1.0) + a * kc);
}

// One pass over rows [i_begin, n).  Returns n when finished, or the index of the row whose
// move needs a cluster slot that does not fit in shared memory (GT = false only): the caller
// migrates the tables to global memory and continues there with the same row (resume = true:
// that row's draw is already published).
//
// Two rows are in flight.  While warp 0 draws the new cluster of row i from its finished
// scores, the evaluator warps already score row i+1 against the current tables.  If row i
// stays where it is and its own cluster's sums did not drift (the common case), those scores
// are exactly what the sequential scan would have computed and the next stage starts at once.
// Otherwise the move (or the drift) is applied first and row i+1 is scored again: the result
// is always that of the strictly sequential scan, only the latency of a step drops from
// evaluate + draw to max(evaluate, draw).
template <int NT, bool GT>
__device__ __forceinline__ int sweep(const Ctx& c, int i_begin, bool resume) {
    const Tab<GT> T(c);
    const int tid = threadIdx.x, lane = tid & 31;
    const int D = c.D, B = c.B, n = c.n, Kcap = c.st.Kcap;
    // evaluator threads (warps 1..): groups of G lanes, one candidate cluster per group
    int lgG = 0;
    while ((1 << lgG) < D && lgG < 5) ++lgG;
    const int G = 1 << lgG;
    const int et = tid - 32;
    const int gid = et >> lgG, gl = et & (G - 1), ngroups = (NT - 32) >> lgG;
    const int pslot = T.prior();
    const size_t fs = T.fs;
    // per-column hyperparameters / types: shared memory, or global memory for very wide views
    const bool wide = GT && c.wide;
    const bool allnorm = c.allnorm != 0;
    const double* ghyp = c.st.hyp + (size_t)c.s * c.st.C * 4;
    auto H = [&](int f, int j) -> double { return wide ? __ldg(&ghyp[c.col[j] * 4 + f]) : c.hyp[f * D + j]; };
    auto CTYPE = [&](int j) -> int { return wide ? c.st.ctype[c.col[j]] : c.ctype[j]; };
    auto NCAT = [&](int j) -> int { return wide ? c.st.ncat[c.col[j]] : c.ncat[j]; };
    const int lgB = 31 - __clz(B);                 // B is a power of two
    const int lpstride = Kcap + 1;
    // Fast path of the scoring (shared-memory tables, all-normal view, at most two columns per lane,
    // one candidate per group): whatever does not change from row to row lives in registers.
    const bool fastp = !GT && allnorm && D <= 2 * G;
    const int fsI = GT ? 0 : c.Kc * D;
    const int j0 = (gl < D) ? gl : 0, j1 = (gl + G < D) ? gl + G : j0;
    const bool on0 = gl < D, on1 = gl + G < D;
    int xo0 = 0, xo1 = 0;
    double nu0 = 0.0, r0 = 0.0, nu1 = 0.0, r1 = 0.0;
    if (fastp && tid >= 32) {
        xo0 = c.bulk ? c.col[j0] : j0; xo1 = c.bulk ? c.col[j1] : j1;
        nu0 = c.hyp[3 * D + j0]; r0 = c.hyp[D + j0];
        nu1 = c.hyp[3 * D + j1]; r1 = c.hyp[D + j1];
    }

    // ---- scores of row i against every candidate cluster (evaluator warps) -----------------
    auto eval_row = [&](int i) {
        const int b = i >> lgB, slot = i & (B - 1), buf = b & 1, pb = i & 1;
        if (slot == 0 || i == i_begin) mbar_wait(&c.bar_full[buf], (b >> 1) & 1);
        const int z = c.zrow[buf * 32 + slot];
        const double* xrow = c.x + (buf * B + slot) * c.xstride;
        const int K = c.scal[0];
        const int nkz = c.nk[z];
        const bool singleton = nkz == 1;
        const int ncand = K + (singleton ? 0 : 1);
        double* lpb = c.lp + pb * lpstride;
        double* wb = c.w + pb * lpstride;
        int* pub = c.scal + 16 + pb * 8;
        if (fastp && ncand <= ngroups) {
            const int pc = gid;
            const bool active = pc < ncand;
            const int k = (pc < K) ? c.live[pc] : pslot;
            const bool own = (pc < K) && (k == z);
            const int ke = (own && singleton) ? pslot : k;
            const bool dn = own && !singleton;
            int drift = 0;
            double acc = 0.0;
            if (active) {
                const double* pe = c.tab + ke * D, *po = c.tab + k * D;
                const double t0 = normal_core(pe + j0, po + j0, fsI, xrow[xo0], nu0, r0, own, dn, drift,
                                              [&](double& m, double& ss) { m = c.hyp[j0]; ss = c.hyp[2 * D + j0]; });
                if (on0) acc = t0;
                if (D > G) {
                    const double t1 = normal_core(pe + j1, po + j1, fsI, xrow[xo1], nu1, r1, own, dn, drift,
                                                  [&](double& m, double& ss) { m = c.hyp[j1]; ss = c.hyp[2 * D + j1]; });
                    if (on1) acc += t1;
                }
            }
            for (int o = G >> 1; o > 0; o >>= 1) {
                acc += __shfl_xor_sync(0xffffffffu, acc, o);
                drift |= __shfl_xor_sync(0xffffffffu, drift, o);
            }
            if (gl == 0 && active) {
                lpb[pc] = acc;
                wb[pc] = (pc >= K) ? c.alpha : own ? (singleton ? c.alpha : (double)(nkz - 1)) : (double)c.nk[k];
                if (own) { pub[0] = pc; pub[1] = drift; }
            }
            return;
        }
        for (int base = 0; base < ncand; base += ngroups) {      // trip count uniform per warp (full-mask shuffles)
            const int pc = base + gid;
            const bool active = pc < ncand;
            const int k = (active && pc < K) ? c.live[pc] : pslot;
            const bool own = active && (pc < K) && (k == z);
            // a singleton's own cluster without the row IS the empty cluster (exactly)
            const int ke = (own && singleton) ? pslot : k;
            const bool dn = own && !singleton;
            int drift = 0;
            double acc = 0.0;
            if (active) {
                if (allnorm && D > G) {
                    // two columns per lane and trip: independent chains, so the two logs overlap
                    for (int j = gl; j < D; j += 2 * G) {
                        const int j2 = j + G;
                        const double t1 = normal_term<GT>(c, T, wide, ghyp, xrow, j, k, ke, own, dn, drift);
                        const double t2 = normal_term<GT>(c, T, wide, ghyp, xrow, j2 < D ? j2 : j, k, ke, own, dn, drift);
                        acc += t1;
                        if (j2 < D) acc += t2;
                    }
                } else if (allnorm) {
                    if (gl < D) acc = normal_term<GT>(c, T, wide, ghyp, xrow, gl, k, ke, own, dn, drift);
                } else {
                    for (int j = gl; j < D; j += G) acc += (CTYPE(j) == CC_NORMAL) ? normal_term<GT>(c, T, wide, ghyp, xrow, j, k, ke, own, dn, drift)
                                                                    : categorical_term<GT>(c, T, wide, ghyp, xrow, j, k, own, pc >= K);
                }
            }
            for (int o = G >> 1; o > 0; o >>= 1) {
                acc += __shfl_xor_sync(0xffffffffu, acc, o);
                drift |= __shfl_xor_sync(0xffffffffu, drift, o);
            }
            if (gl == 0 && active) {
                lpb[pc] = acc;
                // crp.py:111-124: n_k, n_k - 1 for the row's own table, alpha for a fresh / re-used singleton table
                wb[pc] = (pc >= K) ? c.alpha : own ? (singleton ? c.alpha : (double)(nkz - 1)) : (double)c.nk[k];
                if (own) { pub[0] = pc; pub[1] = drift; }
            }
        }
    };

    // ---- the draw for row i (warp 0): first j with prefix_j > u * total over w_j exp(lp_j - max) ----
    auto draw_row = [&](int i, int r, int z, int K, int ncand, int buf, int slot) {
        const int pb = i & 1;
        const double* lpb = c.lp + pb * lpstride;
        const double* wb = c.w + pb * lpstride;
        int* pub = c.scal + 16 + pb * 8;
        const double uu = c.u[buf * 32 + slot];
        int choice = -1;
        if (ncand <= 32) {
            const double lpv = (lane < ncand) ? lpb[lane] : -INFINITY;
            const double wv = (lane < ncand) ? wb[lane] : 0.0;
            // any common reference point near the maximum will do (it cancels in inc > u * tot):
            // the float-rounded maximum costs one redux instead of five 64-bit shuffle rounds
            float mf = (float)lpv;
            asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(mf) : "f"(mf));
            const double dl = lpv - (double)mf;
            // below e^-80 of the maximum a candidate cannot change a 53-bit prefix sum; keeping the
            // argument in range also keeps exp on its fast path
            double inc = (dl > -80.0) ? wv * exp(fmax(dl, -80.0)) : ((dl == dl) ? 0.0 : dl);
            double t;
            t = __shfl_up_sync(0xffffffffu, inc, 1); if (lane >= 1) inc += t;
            t = __shfl_up_sync(0xffffffffu, inc, 2); if (lane >= 2) inc += t;
            if (ncand > 4) { t = __shfl_up_sync(0xffffffffu, inc, 4); if (lane >= 4) inc += t; }
            if (ncand > 8) { t = __shfl_up_sync(0xffffffffu, inc, 8); if (lane >= 8) inc += t; }
            if (ncand > 16) { t = __shfl_up_sync(0xffffffffu, inc, 16); if (lane >= 16) inc += t; }
            const double tot = __shfl_sync(0xffffffffu, inc, ncand - 1);
            const unsigned hit = __ballot_sync(0xffffffffu, lane < ncand && inc > uu * tot);
            choice = hit ? (__ffs(hit) - 1) : -1;
        } else {
            double mx = -INFINITY;
            for (int pc = lane; pc < ncand; pc += 32) mx = fmax(mx, lpb[pc]);
            mx = warp_max(mx);
            double carry = 0.0;
            for (int base = 0; base < ncand; base += 32) {
                const int pc = base + lane;
                const double e = (pc < ncand) ? wb[pc] * exp(lpb[pc] - mx) : 0.0;
                carry += __shfl_sync(0xffffffffu, warp_incl_scan(e, lane), 31);
            }
            const double tot = carry;
            carry = 0.0;
            for (int base = 0; base < ncand && choice < 0; base += 32) {
                const int pc = base + lane;
                const double e = (pc < ncand) ? wb[pc] * exp(lpb[pc] - mx) : 0.0;
                const double inc = warp_incl_scan(e, lane) + carry;
                const unsigned hit = __ballot_sync(0xffffffffu, pc < ncand && inc > uu * tot);
                if (hit) choice = base + __ffs(hit) - 1;
                carry = __shfl_sync(0xffffffffu, inc, 31);
            }
        }
        if (choice < 0) {                       // NaN / all -inf: keep the row, flag the chain
            choice = pub[0];
            if (lane == 0) c.scal[7] = CC_ERR_ARG;
        }
        if (c.trace && lane == 0) {
            double* tr = c.trace + c.trace_off + (size_t)i * c.trace_stride;
            tr[0] = (double)r;
            tr[1] = (double)ncand;
            tr[2] = (double)((choice < K) ? c.g_cid[c.live[choice]] : (c.g_cid[c.live[K - 1]] + 1));
            for (int pc = 0; pc < ncand && pc + 3 < c.trace_stride; ++pc) tr[3 + pc] = lpb[pc] + log(wb[pc]);
        }
        if (lane == 0) {
            // destination slot and "new cluster" flag, resolved here so that nobody reads the
            // live list / free head while the bookkeeping thread rewrites them
            int nc = choice >= K;
            int dst = nc ? c.scal[2] : c.live[choice];
            if (nc && dst >= Kcap - 1) { c.scal[7] = CC_ERR_CAPACITY; nc = 0; dst = z; }   // no free slot: keep the row
            pub[2] = choice; pub[3] = dst; pub[4] = nc;
        }
    };

    if (!resume && i_begin < n) {
        if (tid >= 32) eval_row(i_begin);
        consumer_sync<NT>();
    }
    // Speculation pays while rows rarely move: every move (or drift) costs a second scoring of the next row.
    // Each thread keeps the same exponential average of the re-done fraction (all see the same outcomes) and
    // the CTA falls back to score-after-draw, the plain sequential order, while more than ~1/4 of the recent
    // rows were re-done (the first sweeps after a random initialisation).
    int redo_avg = 0;                              // 65536 * fraction, time constant 16 rows
    for (int i = i_begin; i < n; ++i) {
        const bool speculate = redo_avg < 16384;
        const int b = i >> lgB, slot = i & (B - 1), buf = b & 1;
        if (slot == 0 || i == i_begin) mbar_wait(&c.bar_full[buf], (b >> 1) & 1);
        const int r = c.rowidx[buf * 32 + slot];
        const int z = c.zrow[buf * 32 + slot];
        const double* xrow = c.x + (buf * B + slot) * c.xstride;
        const int K = c.scal[0];
        const int nkz = c.nk[z];
        const bool singleton = nkz == 1;

        // ---- stage i: draw row i  ||  score row i+1 (speculatively) ------------------------
        if (!(resume && i == i_begin)) {
            if (tid < 32) draw_row(i, r, z, K, K + (singleton ? 0 : 1), buf, slot);
            else if (speculate && i + 1 < n) eval_row(i + 1);
            consumer_sync<NT>();
        }
        const int* pub = c.scal + 16 + (i & 1) * 8;
        const int ownpos = pub[0];
        const int zb = pub[3];
        const bool newc = pub[4] != 0;
        bool redo = false;
        if (zb != z) {
            // ---- apply the move (view.py:424-429) ------------------------------------------
            if (!GT && newc && zb >= c.Kc - 1) return i;     // needs the global-memory tables
            // the lower half of the CTA updates the old cluster, the upper half the new one (different
            // warps, so the two divergent update paths run side by side)
            const int half = NT / 2;
            const bool is_old = tid < half;
            for (int j = tid % half; j < D; j += half) {
                if (is_old && singleton) continue;            // the emptied cluster is dropped
                const double x = xrow[c.bulk ? c.col[j] : j];
                const int k = is_old ? z : zb;
                const bool isnan_x = isnan(x);
                if (CTYPE(j) == CC_NORMAL) {
                    const double m = H(0, j), rr = H(1, j), ss = H(2, j), nu = H(3, j);
                    double N, sx, sxx, gN, gM1;
                    if (is_old) {
                        if (isnan_x) continue;
                        // scored against its own cluster first (-= x, += x: view.py:416-419), then removed
                        const double xx = __dmul_rn(x, x);
                        N = T.at(CC_F_N, k, j) - 1.0;
                        sx = __dsub_rn(__dadd_rn(__dsub_rn(T.at(CC_F_SX, k, j), x), x), x);
                        sxx = __dsub_rn(__dadd_rn(__dsub_rn(T.at(CC_F_SXX, k, j), xx), xx), xx);
                        gN = T.at(CC_F_GM1, k, j);
                        gM1 = (N >= 1.0) ? normal_G(N - 1.0, rr, nu) : 0.0;
                    } else if (newc) {
                        // fresh cluster object: 0 + x, 0 + x*x (normal.py:51-53,64-70)
                        if (isnan_x) { N = 0.0; sx = 0.0; sxx = 0.0; gN = T.at(CC_F_GN, pslot, j); gM1 = 0.0; }
                        else {
                            N = 1.0; sx = __dadd_rn(0.0, x); sxx = __dadd_rn(0.0, __dmul_rn(x, x));
                            gM1 = T.at(CC_F_GN, pslot, j);
                            gN = normal_G(1.0, rr, nu);
                        }
                    } else {
                        if (isnan_x) continue;
                        N = T.at(CC_F_N, k, j) + 1.0;
                        sx = __dadd_rn(T.at(CC_F_SX, k, j), x);
                        sxx = __dadd_rn(T.at(CC_F_SXX, k, j), __dmul_rn(x, x));
                        gM1 = T.at(CC_F_GN, k, j);
                        gN = normal_G(N, rr, nu);
                    }
                    const NormalCache cc = normal_cache_fast(N, sx, sxx, gN, m, rr, ss, nu);
                    T.at(CC_F_N, k, j) = N; T.at(CC_F_SX, k, j) = sx; T.at(CC_F_SXX, k, j) = sxx;
                    T.at(CC_F_MN, k, j) = cc.mn; T.at(CC_F_A, k, j) = cc.a; T.at(CC_F_CST, k, j) = cc.cst;
                    T.at(CC_F_GN, k, j) = gN; T.at(CC_F_GM1, k, j) = gM1;
                } else {
                    const double a = H(0, j);
                    const int kc = NCAT(j);
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
            // bookkeeping of the row CRP (crp.py:45-59) by one thread, concurrently with the table updates
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
                    while (fh < Kcap - 1 && c.nk[fh] != 0) ++fh;
                    c.scal[2] = fh;
                } else {
                    c.nk[zb] += 1;
                }
                if (singleton) {
                    // the emptied cluster leaves the ascending-id list; its slot is free again
                    for (int q = ownpos; q < K - 1; ++q) c.live[q] = c.live[q + 1];
                    c.live[K - 1] = -1;
                    c.scal[0] = K - 1;
                    if (z < c.scal[2]) c.scal[2] = z;
                }
            }
            redo = true;
        } else if (pub[1]) {
            // ---- the row stays, but `-= x; += x` changed its cluster's sums in the last bit
            //      (view.py:416-419): store them and refresh the cached constants ------------
            for (int j = tid; j < D; j += NT) {
                if (CTYPE(j) != CC_NORMAL) continue;
                const double x = xrow[c.bulk ? c.col[j] : j];
                if (isnan(x)) continue;
                const double sx0 = T.at(CC_F_SX, z, j), sxx0 = T.at(CC_F_SXX, z, j);
                const double xx = __dmul_rn(x, x);
                const double sx2 = __dadd_rn(__dsub_rn(sx0, x), x), sxx2 = __dadd_rn(__dsub_rn(sxx0, xx), xx);
                if (sx2 != sx0 || sxx2 != sxx0) {
                    T.at(CC_F_SX, z, j) = sx2;
                    T.at(CC_F_SXX, z, j) = sxx2;
                    const NormalCache c2 = normal_cache_fast(T.at(CC_F_N, z, j), sx2, sxx2, T.at(CC_F_GN, z, j), H(0, j), H(1, j), H(2, j), H(3, j));
                    T.at(CC_F_MN, z, j) = c2.mn; T.at(CC_F_A, z, j) = c2.a; T.at(CC_F_CST, z, j) = c2.cst;
                }
            }
            redo = true;
        }
        redo_avg += (redo ? 4096 : 0) - (redo_avg >> 4);
        if (redo) consumer_sync<NT>();
        if ((redo || !speculate) && i + 1 < n) {
            if (tid >= 32) eval_row(i + 1);              // not scored yet, or the speculative scores are stale
            consumer_sync<NT>();
        }
        if ((slot == B - 1 || i == n - 1) && tid == 0) mbar_arrive(&c.bar_empty[buf]);
    }
    return n;
}

template <int NT, int MINB>
__global__ void __launch_bounds__(NT + 32, MINB) rows_pipe_kernel(const RowsParams p) {
    const cc_state& st = p.st;
    const cc_row_work wk = p.work[blockIdx.x];
    const int s = wk.chain, v = wk.view, n = wk.n;
    const int R = st.R, C = st.C, Cp = st.Cp, Kcap = st.Kcap, CT = st.CT;
    const int tid = threadIdx.x;
    const bool producer = tid >= NT;
    const int B = p.B;
    const size_t sv = sv_index(st, s, v);

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const long long t_start = clock64();
    Ctx c;
    c.st = st; c.s = s; c.v = v; c.n = n; c.B = B; c.bulk = p.bulk; c.sv = sv;
    // ---- fixed shared region ----------------------------------------------------
    c.bar_full = reinterpret_cast<uint64_t*>(smem_raw);               // [2]
    c.bar_empty = c.bar_full + 2;                                     // [2]
    c.scal = reinterpret_cast<int*>(c.bar_empty + 2);                 // [32]
    c.live = c.scal + 32;                                             // [Kcap]
    c.nk = c.live + Kcap;                                             // [Kcap]
    c.lp = reinterpret_cast<double*>(c.nk + Kcap);                    // [2][Kcap+1]
    const int32_t* zv = st.zv + (size_t)s * C;
    __shared__ int s_D_tmp;
    if (tid == 0) {
        int D = 0;
        for (int q = 0; q < C; ++q) D += (zv[q] == v);
        s_D_tmp = D;
        mbar_init(&c.bar_full[0], 33); mbar_init(&c.bar_full[1], 33);
        mbar_init(&c.bar_empty[0], 1); mbar_init(&c.bar_empty[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int D = s_D_tmp;
    c.D = D;
    c.wide = p.wide;
    c.w = c.lp + 2 * (Kcap + 1);                                      // [2][Kcap+1]
    c.hyp = c.w + 2 * (Kcap + 1);                                         // [4][D] (not in wide mode)
    c.u = c.hyp + (p.wide ? 0 : 4 * D);                               // [2][32]
    c.col = reinterpret_cast<int*>(c.u + 64);                         // [D]
    c.cto = c.col + D;                                                // [D]   (these three not in wide mode)
    c.ctype = c.cto + (p.wide ? 0 : D);                               // [D]
    c.ncat = c.ctype + (p.wide ? 0 : D);                              // [D]
    c.rowidx = c.ncat + (p.wide ? 0 : D);                             // [2][32]
    c.zrow = c.rowidx + 64;                                           // [2][32]
    size_t off = (size_t)((unsigned char*)(c.zrow + 64) - smem_raw);
    off = (off + 15) & ~(size_t)15;
    c.xstride = p.bulk ? Cp : ((D + 1) & ~1);
    c.x = reinterpret_cast<double*>(smem_raw + off);                  // [2][B][xstride]
    off += (size_t)2 * B * c.xstride * sizeof(double);
    c.tab = reinterpret_cast<double*>(smem_raw + off);
    const size_t tab_avail = (p.smem_bytes > (int)off) ? (size_t)p.smem_bytes - off : 0;

    if (tid == 0) {
        int j = 0, ct = 0, alln = 1;
        for (int q = 0; q < C; ++q)
            if (zv[q] == v) {
                c.col[j] = q;
                if (st.ctype[q] != CC_NORMAL) alln = 0;
                if (!p.wide) {
                    c.cto[j] = ct;
                    c.ctype[j] = st.ctype[q];
                    c.ncat[j] = st.ncat[q];
                }
                ct += st.ncat[q];
                ++j;
            }
        c.scal[6] = ct;
        c.scal[8] = alln;
        c.scal[3] = 0;
        c.scal[7] = 0;
    }
    const int K0 = st.nclu[sv];
    int32_t* g_live = st.live + sv * Kcap;
    int32_t* g_nk = st.nk + sv * Kcap;
    c.g_cid = st.cid + sv * Kcap;
    for (int i = tid; i < Kcap; i += blockDim.x) {
        c.nk[i] = 0;
        c.live[i] = (i < K0) ? g_live[i] : -1;
    }
    __syncthreads();
    for (int i = tid; i < K0; i += blockDim.x) c.nk[c.live[i]] = g_nk[c.live[i]];
    for (int e = tid; !p.wide && e < 4 * D; e += blockDim.x) {
        const int f = e / D, j = e - f * D;
        c.hyp[e] = st.hyp[((size_t)s * C + c.col[j]) * 4 + f];
    }
    __syncthreads();
    const int CTv = c.scal[6];
    c.CTv = CTv;
    c.allnorm = c.scal[8];
    c.K0 = K0;
    // shared-table capacity: Kc cluster slots (the empty-cluster record sits at slot Kc-1)
    const size_t per_cluster = (size_t)(CC_NFIELD * D + CTv) * sizeof(double);
    int Kc = (int)(tab_avail / (per_cluster ? per_cluster : 1));
    if (Kc > Kcap) Kc = Kcap;
    c.Kc = Kc;
    if (tid == 0) {
        int maxslot = -1;
        for (int i = 0; i < K0; ++i) maxslot = max(maxslot, c.live[i]);
        int fh = 0;
        while (fh < Kcap - 1 && c.nk[fh] != 0) ++fh;
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
    const int nbatch = (n + B - 1) / B;

    if (producer) {
        // =========================== producer warp ===========================
        const int lane = tid - NT;
        uint32_t bits = 1;
        while ((1u << bits) < (uint32_t)n && bits < 31) ++bits;
        const uint4 pkey = rng(p.counter, stream_id(2, s, v));
        const uint64_t ustream = stream_id(1, s, v);
        for (int b = 0; b < nbatch; ++b) {
            const int buf = b & 1;
            if (b >= 2) mbar_wait_sleep(&c.bar_empty[buf], ((b >> 1) - 1) & 1);
            const int i = b * B + lane;
            const bool act = lane < B && i < n;
            int r = 0;
            if (act) {
                int idx;
                if (wk.perm_off >= 0) idx = p.perm[wk.perm_off + i];
                else idx = (int)keyed_perm((uint32_t)i, (uint32_t)n, bits, pkey);
                r = wk.perm_is_index ? order[idx] : idx;
                c.rowidx[buf * 32 + lane] = r;
                double uu;
                if (wk.u_off >= 0) uu = p.u[wk.u_off + i];
                else { const uint4 w4 = rng(p.counter * 0x100000000ull + (uint64_t)i, ustream); uu = u01_from_words(w4.x, w4.y); }
                c.u[buf * 32 + lane] = uu;
            }
            const int nrows = min(B, n - b * B);
            __syncwarp();
            if (lane == 0) mbar_arrive_expect_tx(&c.bar_full[buf], p.bulk ? (uint32_t)(nrows * Cp * 8) : 0u);
            __syncwarp();
            if (act) {
                cp_async_4(&c.zrow[buf * 32 + lane], &c.zr[r]);
                double* dst = c.x + (size_t)(buf * B + lane) * c.xstride;
                const double* src = st.Xrm + (size_t)r * Cp;
                if (p.bulk) bulk_g2s(dst, src, (uint32_t)(Cp * 8), &c.bar_full[buf]);
                else for (int j = 0; j < D; ++j) cp_async_8(&dst[j], &src[c.col[j]]);
            }
            cp_async_mbar_arrive_noinc(&c.bar_full[buf]);
        }
        return;
    }

    // ============================== consumers ===============================
    int done = 0;
    if (!start_global) {
        done = sweep<NT, false>(c, 0, false);
        if (done < n) {
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
        }
    }
    const bool in_global = start_global || done < n;
    if (in_global) sweep<NT, true>(c, done, !start_global);

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
    for (int i = tid; i < Kcap; i += NT) {
        g_live[i] = (i < K) ? c.live[i] : -1;
        g_nk[i] = c.nk[i];
    }
    if (tid == 0) {
        st.nclu[sv] = K;
        if (c.scal[7]) st.err[s] = c.scal[7];
        if (p.prof) {
            int32_t* q = st.seq + sv * R;
            q[0] = (int)((clock64() - t_start) >> 10); q[1] = D; q[2] = K; q[3] = c.scal[3];
            q[4] = 4 + (start_global ? 1 : (in_global ? 2 : 0));      // two-rows-in-flight kernel; table mode: shared / global / migrated mid-sweep
        }
    }
    // ---- member order of the view's CRP: re-seated rows move to the end, in the
    //      order they were re-seated (OrderedDict pop + re-insert, crp.py:52,55) -----
    const int nmoved = c.scal[3];
    if (nmoved > 0) {
        const int lane = tid & 31;
        int32_t* ord = st.order + sv * R;
        __shared__ int s_wsum[32];
        __shared__ int s_base;
        if (tid == 0) s_base = 0;
        consumer_sync<NT>();
        for (int i0 = 0; i0 < R; i0 += NT) {
            const int i = i0 + tid;
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

// shared memory a CTA needs besides the tables (see the layout in rows_pipe_kernel)
int cc_rows_wide_fixed_smem(int Kcap, int dmax, int wide, int xbytes) {
    return 64 + 128 + Kcap * 8 + (Kcap + 2) * 32 + dmax * (wide ? 4 : 48) + 512 + 64 + 512 + xbytes;
}

int cc_launch_rows_wide(const RowsParams& p, int cnt, cudaStream_t stream) {
    CC_CUDA(cudaFuncSetAttribute(rows_pipe_kernel<224, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem_bytes));
    rows_pipe_kernel<224, 2><<<cnt, 256, p.smem_bytes, stream>>>(p);
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}
