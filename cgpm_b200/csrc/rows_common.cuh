// Pieces shared by the rows kernels (rows.cu): launch parameters, the per-CTA context, table
// access, the cell scoring routines.
#pragma once
#include "cc_common.cuh"

struct RowsParams {
    cc_state st;
    const cc_row_work* work;
    const int32_t* perm;
    const double* u;
    double* trace;
    int trace_stride;
    uint64_t seed, counter;
    int B, NBUF;    // (unused: every CTA sizes its own staging ring, see rows_stage)
    int bulk;       // 1: bulk row copies allowed (cp.async.bulk; each CTA decides from its view's width)
    int smem_bytes; // dynamic shared memory given to the CTA
    int wide;       // 1: very wide views -- per-column hypers / types stay in global memory, tables too
    int prof;       // debug: per-CTA cycles / D / K / moves into st.seq[(chain, view)][0..4]
    int Ksm;        // capacity of the candidate arrays (live, nk, lp, w) in shared memory; the large-K kernel
                    // keeps them in global memory instead
    int DL;         // views wider than this are swept one row at a time by the whole CTA
    int lvl0;       // debug: fixed speculation level (-1: adaptive)
    double* big_lp; // large-K kernel: [n_work][2][Kcap + 1] candidate scores / weights
};

// How one view's rows are staged: whole rows by bulk copies when the row is short and the view holds at least a
// quarter of it, per-cell gathers otherwise; NBUF (<= 8) batches of B (<= 32) rows within about 64 KB (32 KB for long rows; at least two
// batches of one row).  The ring bounds the number of rows a speculative round can have in flight.  Host (launch
// sizing) and device (per CTA) use the same rule.
struct RowsStage { int bulk, stagebytes, B, NBUF; };
__host__ __device__ inline RowsStage rows_stage(int Cp, int C, int D, int allow_bulk) {
    RowsStage r;
    const int rowbytes = Cp * 8;
    r.bulk = (allow_bulk && rowbytes <= 4096 && 4 * D >= C) ? 1 : 0;
    r.stagebytes = r.bulk ? rowbytes : ((D + 1) & ~1) * 8;
    r.B = 32; r.NBUF = 8;
    // long rows (1 KB and more) belong to wide views: their tables need the shared memory more than the ring does,
    // and a round over such a row is long enough for 16 rows in flight to cover the HBM latency
    const int lim = (r.stagebytes >= 1024) ? 32 * 1024 : 64 * 1024;
    while (r.NBUF > 4 && r.NBUF * r.B * r.stagebytes > lim) r.NBUF >>= 1;
    while (r.B > 1 && r.NBUF * r.B * r.stagebytes > lim) r.B >>= 1;
    while (r.NBUF > 2 && r.NBUF * r.B * r.stagebytes > lim) r.NBUF >>= 1;
    return r;
}

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
    int s, v, n, D, Dn, K0, Kc, CTv, B, lgB, NBUF, RB, bulk, xstride, wide, allnorm, twotypes, Ksm, DL, base, lvl0;   // Dn: normal columns (they come first)
    size_t sv;
    // shared memory
    uint64_t *bar_full, *bar_empty;   // [8] each
    int *scal;              // [64]: 0 K, 2 free head, 3 nmoved, 5 start in global, 6 CTv, 7 error, 11 all-normal view,
                            // 40 scratch of the parallel searches, 42 batches released, 43 stop flag for the producer;
                            // one row at a time: 16 + 8 * (row parity) + {0 own position, 1 drift flag, 2 choice,
                            // 3 destination slot, 4 new-cluster flag}
    int *pub;               // [2][16][8] by round parity and warp: first flagged row of the warp (speculative rounds)
    int *live, *nk;         // [Ksm] (shared) or [Kcap] (global, large-K kernel)
    double* lp;             // [Ksm+1] / [Kcap+1] candidate scores (one row at a time)
    double* w;              // same shape: CRP weight of each candidate (written with lp)
    int *col, *cto, *ctype, *ncat;
    double* hyp;            // [4][D]
    int *rowidx, *zrow;     // [RB]
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

// Normal cell of one row against one candidate cluster.
//   other cluster: cst - b log(1 + a (x-mn)^2)   (Student-t form, cached a, mn, cst)
//   own cluster  : the row is removed first (view.py:416-419).  With sn1 = sn - rn/(rn-1) (x-mn)^2 (the
//     Normal-Gamma downdate; rn, mn, sn, a = rn/((rn+1) sn) of the cluster WITH the row) the predictive
//     G(N-1) - b1 log sn + (b1-1/2) log sn1 becomes cst + G(N-1) - G(N) + (b1-1/2) log(1 - a (rn+1)/(rn-1)
//     (x-mn)^2): the same loads and the same single log as for any other cluster.  Heavy cancellation
//     (arg < 2^-20) takes the reference's posterior update of the decremented sums instead.  A singleton's
//     own cluster without the row IS the empty cluster (the caller passes that cell as p).  The `-= x; += x`
//     replay (view.py:416-419) is reported in `drift`: set if the sums come back changed.
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
    const double t = bs + coef * cc_log_pos(fma(sc * zz, zz, 1.0));     // argument in [2^-20, inf): no special cases
    return isnan(x) ? 0.0 : t;                                   // dim.py:130-132
}


// The same cell without a branch between the two formulas: every lane runs one instruction stream and selects its
// constants (a lane scoring the row's own cluster next to lanes scoring other clusters would otherwise make the warp
// run the two logs one after the other).  p: the candidate's cell -- for a lone member's own cluster the empty
// cluster's; po: the cell of the cluster whose sums go through `-= x; += x` (read by every lane, used if own).
template <typename FS, typename MS>
__device__ __forceinline__ double normal_core_bf(const double* p, const double* po, FS fs, double x, double nu, double rr,
                                                 bool own, bool dn, int& drift, MS ms) {
    const double Nf = p[CC_F_N * fs];
    const double zz = x - p[CC_F_MN * fs];
    const double a = p[CC_F_A * fs], cst = p[CC_F_CST * fs];
    const double dG = p[CC_F_GM1 * fs] - p[CC_F_GN * fs];
    const double sx0 = po[CC_F_SX * fs], sxx0 = po[CC_F_SXX * fs];
    const double xx = __dmul_rn(x, x);
    const double sx1 = __dsub_rn(sx0, x), sxx1 = __dsub_rn(sxx0, xx);
    const bool nanx = isnan(x);
    if (own && !nanx && (__dadd_rn(sx1, x) != sx0 || __dadd_rn(sxx1, xx) != sxx0)) drift = 1;
    const double rn = rr + Nf;
    const double sc = dn ? -a * ((rn + 1.0) * __drcp_rn(rn - 1.0)) : a;
    const double coef = dn ? 0.5 * (nu + Nf - 1.0) : -0.5 * (nu + Nf + 1.0);
    const double bs = dn ? cst + dG : cst;
    const double arg = fma(sc * zz, zz, 1.0);
    const bool exact = dn && !(arg >= 0x1p-20);       // heavy cancellation in the downdate (or NaN): the reference's posterior update
    double t = bs + coef * cc_log_pos(exact ? 1.0 : arg);
    if (exact) {
        double m, ss;
        ms(m, ss);
        const double N1 = Nf - 1.0;
        const double rn1 = rr + N1;
        const double mn1 = (rr * m + sx1) * __drcp_rn(rn1);
        double sn1 = __dsub_rn(__dadd_rn(__dadd_rn(ss, sxx1), __dmul_rn(__dmul_rn(rr, m), m)),
                               __dmul_rn(__dmul_rn(rn1, mn1), mn1));
        if (sn1 == 0.0) sn1 = ss;
        const double b1 = 0.5 * (nu + N1 + 1.0);
        t = (p[CC_F_GM1 * fs] - b1 * (2.0 * (p[CC_F_GN * fs] - cst))) + coef * log(sn1);
    }
    return nanx ? 0.0 : t;
}


// The row's own cluster without the row (the `dn` case of normal_core), split so that two cells can be in flight:
// the straight-line part reports heavy cancellation in `exact` instead of branching, and the caller then takes
// normal_own_exact (the reference's posterior update of the decremented sums) for that cell.
template <typename FS>
__device__ __forceinline__ double normal_own_fast(const double* p, FS fs, double x, double nu, double rr, int& drift, bool& exact) {
    const double Nf = p[CC_F_N * fs];
    const double zz = x - p[CC_F_MN * fs];
    const double a = p[CC_F_A * fs], cst = p[CC_F_CST * fs];
    const double dG = p[CC_F_GM1 * fs] - p[CC_F_GN * fs];
    const double sx0 = p[CC_F_SX * fs], sxx0 = p[CC_F_SXX * fs];
    const double xx = __dmul_rn(x, x);
    const double sx1 = __dsub_rn(sx0, x), sxx1 = __dsub_rn(sxx0, xx);
    const bool nanx = isnan(x);
    if (!nanx && (__dadd_rn(sx1, x) != sx0 || __dadd_rn(sxx1, xx) != sxx0)) drift = 1;
    const double rn = rr + Nf;
    const double sc = -a * ((rn + 1.0) * __drcp_rn(rn - 1.0));
    const double coef = 0.5 * (nu + Nf - 1.0);
    const double arg = fma(sc * zz, zz, 1.0);
    exact = !nanx && !(arg >= 0x1p-20);
    const double t = (cst + dG) + coef * cc_log_pos(exact ? 1.0 : arg);
    return nanx ? 0.0 : t;
}
template <typename FS>
__device__ __forceinline__ double normal_own_exact(const double* p, FS fs, double x, double nu, double rr, double m, double ss) {
    const double Nf = p[CC_F_N * fs];
    const double xx = __dmul_rn(x, x);
    const double sx1 = __dsub_rn(p[CC_F_SX * fs], x), sxx1 = __dsub_rn(p[CC_F_SXX * fs], xx);
    const double N1 = Nf - 1.0;
    const double rn1 = rr + N1;
    const double mn1 = (rr * m + sx1) * __drcp_rn(rn1);
    double sn1 = __dsub_rn(__dadd_rn(__dadd_rn(ss, sxx1), __dmul_rn(__dmul_rn(rr, m), m)), __dmul_rn(__dmul_rn(rn1, mn1), mn1));
    if (sn1 == 0.0) sn1 = ss;
    const double b1 = 0.5 * (nu + N1 + 1.0);
    const double bs = p[CC_F_GM1 * fs] - b1 * (2.0 * (p[CC_F_GN * fs] - p[CC_F_CST * fs]));
    return bs + 0.5 * (nu + Nf - 1.0) * log(sn1);
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

}  // namespace
