// Sufficient-statistic rebuild, column marginal matrix and logpdf_score.
//
//   group_kernel   stable grouping of a view's member order by cluster
//   stats_kernel   per (cluster, column): ordered sums -> tables (TABLE mode) or
//                  -> per-(column, view) marginal likelihood (MARGINAL mode)
//   score_kernel   State.logpdf_score
//
// References: View._bulk_incorporate (src/mixtures/view.py:483-494) sums each
// cluster's members in the order of the view's CRP `data` OrderedDict; this is
// reproduced exactly: a cluster's members are visited in `order`, one thread
// owns one (cluster, column) accumulator, and x*x is rounded before it is added
// (src/primitives/normal.py:64-70).  Dim.logpdf_score (src/mixtures/dim.py:121-122)
// and the marginals of normal.py:175-181 / categorical.py:150-155 give the
// column marginal matrix of State._gibbs_transition_dim (src/crosscat/state.py:1065-1072).
#include "cc_common.cuh"

#define CC_MAXCAT 256

namespace {

// ---------------------------------------------------------------------------
// One CTA per (chain, view): rows grouped by cluster slot, members of a cluster in `order`
// order (a stable counting sort).  The NW warps take contiguous chunks of the member order:
// each counts its chunk per cluster, the counts are scanned over (cluster, warp) on top of the
// clusters' start offsets, and each warp scatters its chunk from its own cursors -- so rows of
// a cluster keep their relative order within and across chunks.  grouped = st.moved (scratch);
// the start offsets go to `gstart` [n_work][Kcap+1].  Shared memory: NW * Kcap ints; when even one
// warp's counters do not fit (Kcap beyond ~24 000: the reference-faithful initialisation of a large
// table) the launch has one warp per CTA and the counters live in `gcnt` [n_work][Kcap] instead.
__global__ void group_kernel(const cc_state st, int n_work, const int32_t* __restrict__ chains,
                             const int32_t* __restrict__ views, int32_t* __restrict__ gstart, int32_t* gcnt) {
    const int w = blockIdx.x;
    if (w >= n_work) return;
    const int s = chains[w], v = views[w];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, NW = blockDim.x >> 5;
    const size_t sv = sv_index(st, s, v);
    const int R = st.R, Kcap = st.Kcap;
    const int32_t* order = st.order + sv * R;
    const int32_t* zr = st.zr + sv * R;
    const int32_t* nk = st.nk + sv * Kcap;
    int32_t* grouped = st.moved + sv * R;
    int32_t* start = gstart + (size_t)w * (Kcap + 1);
    extern __shared__ int s_cnt_sh[];                  // [NW][Kcap] chunk counts, then cursors
    int* s_cnt = gcnt ? gcnt + (size_t)w * Kcap : s_cnt_sh;
    int* mine = s_cnt + warp * Kcap;
    for (int k = lane; k < Kcap; k += 32) mine[k] = 0;
    // start offsets: exclusive scan of nk over slots (warp 0, serial chunks of 32)
    if (warp == 0) {
        int carry = 0;
        for (int base = 0; base < Kcap; base += 32) {
            const int k = base + lane;
            const int c = (k < Kcap) ? nk[k] : 0;
            int inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            if (k < Kcap) start[k] = carry + inc - c;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) start[Kcap] = carry;
    }
    __syncwarp();
    // this warp's chunk of the member order (a multiple of 32 rows)
    const int per = (((R + NW - 1) / NW) + 31) & ~31;
    const int lo = warp * per, hi = min(R, lo + per);
    for (int base = lo; base < hi; base += 32) {
        const int i = base + lane;
        const int k = (i < hi) ? zr[order[i]] : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, k);
        if (k >= 0 && (peers >> lane) == 1u) mine[k] += __popc(peers);      // the highest peer adds for the group
        __syncwarp();
    }
    __syncthreads();
    // cursors: start[k] + the counts of the earlier chunks
    for (int k = threadIdx.x; k < Kcap; k += blockDim.x) {
        int run = start[k];
        for (int q = 0; q < NW; ++q) { const int c = s_cnt[q * Kcap + k]; s_cnt[q * Kcap + k] = run; run += c; }
    }
    __syncthreads();
    // stable scatter of the chunk
    for (int base = lo; base < hi; base += 32) {
        const int i = base + lane;
        const bool act = i < hi;
        int r = 0, k = -1;
        if (act) { r = order[i]; k = zr[r]; }
        const unsigned peers = __match_any_sync(0xffffffffu, k);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        int pos = 0;
        if (act) pos = mine[k] + rank;
        __syncwarp();
        if (act && rank == __popc(peers) - 1) mine[k] = pos + 1;   // last peer advances the cursor
        if (act) grouped[pos] = r;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
struct StatsParams {
    cc_state st;
    int n_work;
    const int32_t* chains;
    const int32_t* views;
    const int32_t* gstart;   // [n_work][Kcap+1]
    int marginal;            // 0: TABLE mode (the view's own columns), 1: MARGINAL mode (listed columns)
    const int32_t* cols;     // MARGINAL: flat column lists (NULL = all columns for every work item)
    const int32_t* coloff;   // MARGINAL: [n_work] offset of the work item's list in `cols`
    const int32_t* colcnt;   // MARGINAL: [n_work] length of the list
    const uint8_t* mask;     // TABLE: optional [S][C]; only columns with mask != 0 are rebuilt
};

// grid: (n_work, column tiles of 32).  block: NW warps; warp w owns clusters
// pos = w, w+NW, ...; lane owns one column of the tile.
template <int NW>
__global__ void __launch_bounds__(NW * 32) stats_kernel(const StatsParams p) {
    const cc_state& st = p.st;
    const int w = blockIdx.x;
    const int s = p.chains[w], v = p.views[w];
    const size_t sv = sv_index(st, s, v);
    const int R = st.R, C = st.C, Cp = st.Cp, Kcap = st.Kcap, CT = st.CT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t* zv = st.zv + (size_t)s * C;

    // column of this lane
    __shared__ int s_cols[32];
    __shared__ int s_ncols;
    if (threadIdx.x == 0) {
        int cnt = 0;
        const int first = blockIdx.y * 32;
        if (p.marginal) {
            const int total = p.cols ? p.colcnt[w] : C;
            const int32_t* list = p.cols ? p.cols + p.coloff[w] : nullptr;
            for (int j = first; j < total && cnt < 32; ++j) s_cols[cnt++] = list ? list[j] : j;
        } else {
            int seen = 0;
            for (int c = 0; c < C && cnt < 32; ++c)
                if (zv[c] == v && (!p.mask || p.mask[(size_t)s * C + c])) { if (seen >= first) s_cols[cnt++] = c; ++seen; }
        }
        s_ncols = cnt;
    }
    __syncthreads();
    const int ncols = s_ncols;
    if (ncols == 0) return;
    const bool act = lane < ncols;
    const int c = act ? s_cols[lane] : 0;
    const int K = st.nclu[sv];
    const int32_t* live = st.live + sv * Kcap;
    const int32_t* grouped = st.moved + sv * R;
    const int32_t* start = p.gstart + (size_t)w * (Kcap + 1);
    const double* hyp = st.hyp + ((size_t)s * C + c) * 4;
    const int ctype = st.ctype[c];
    const int kc = st.ncat[c];
    const double h0 = hyp[0], h1 = hyp[1], h2 = hyp[2], h3 = hyp[3];
    const double* X = st.Xrm + c;

    if (p.marginal && __syncthreads_and(!act || ctype == CC_NORMAL)) {
        // MARGINAL mode, all-normal tile: the marginals are proposal weights (state.py:1065-1072), compared
        // with the reference to 1e-5, so a cluster's sums need not follow the member order bit by bit.  All
        // NW warps share every cluster: contiguous eighths of its member list, partial sums combined in warp
        // order (deterministic).  A view with one dominant cluster no longer waits for a single warp.
        __shared__ double s_p[NW][3][32];
        double tot = 0.0;
        for (int pos = 0; pos < K; ++pos) {
            const int k = live[pos];
            const int beg = start[k], end = start[k + 1];
            const int chunk = (((end - beg) + NW - 1) / NW + 7) & ~7;
            int i = beg + warp * chunk;
            const int e = min(end, i + chunk);
            double N = 0.0, sx = 0.0, sxx = 0.0;
            for (; i + 8 <= e; i += 8) {
                int rr[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) rr[q] = grouped[i + q];
                if (act) {
                    double xv[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) xv[q] = X[(size_t)rr[q] * Cp];
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if (!isnan(xv[q])) { N += 1.0; sx = __dadd_rn(sx, xv[q]); sxx = __dadd_rn(sxx, __dmul_rn(xv[q], xv[q])); }
                }
            }
            for (; i < e; ++i) {
                const int r0 = grouped[i];
                if (act) {
                    const double x0 = X[(size_t)r0 * Cp];
                    if (!isnan(x0)) { N += 1.0; sx = __dadd_rn(sx, x0); sxx = __dadd_rn(sxx, __dmul_rn(x0, x0)); }
                }
            }
            s_p[warp][0][lane] = N; s_p[warp][1][lane] = sx; s_p[warp][2][lane] = sxx;
            __syncthreads();
            if (warp == 0 && act) {
                double tN = 0.0, tsx = 0.0, tsxx = 0.0;
                for (int q = 0; q < NW; ++q) { tN += s_p[q][0][lane]; tsx += s_p[q][1][lane]; tsxx += s_p[q][2][lane]; }
                tot += normal_marginal(tN, tsx, tsxx, h0, h1, h2, h3);
            }
            __syncthreads();
        }
        if (warp == 0 && act) st.colL[((size_t)s * C + c) * (st.Vcap + 1) + v] = tot;
        return;
    }
    const bool sumt = ctype != CC_CATEGORICAL;       // ordered sums N, sum_x and a second term (x*x | gammaln(x+1) | none)
    const double* aux = (act && ctype == CC_POISSON && st.Xaux && st.auxcol[c] >= 0) ? st.Xaux + (size_t)st.auxcol[c] * R : nullptr;
    auto second = [&](double x, int r) -> double {
        return (ctype == CC_NORMAL) ? __dmul_rn(x, x) : sum_second_term_ni(ctype, x, aux ? aux[r] : CUDART_NAN);
    };
    double lacc = 0.0;     // MARGINAL: this warp's partial over its clusters
    for (int pos = warp; pos < K + (p.marginal ? 0 : 1); pos += NW) {
        // pos == K (TABLE mode): the prior (empty-cluster) record at slot Kcap-1
        const bool prior = pos >= K;
        const int k = prior ? (Kcap - 1) : live[pos];
        const int beg = prior ? 0 : start[k], end = prior ? 0 : start[k + 1];
        double N = 0.0, sx = 0.0, sxx = 0.0;
        uint32_t lc[CC_MAXCAT];
        if (act && ctype == CC_CATEGORICAL) for (int q = 0; q < kc; ++q) lc[q] = 0;
        int i = beg;
        // ordered accumulation: loads are batched 16 (then 8) deep (independent), the adds stay sequential
        if (ctype == CC_NORMAL) {
            for (; i + 16 <= end; i += 16) {
                int rr[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) rr[q] = grouped[i + q];
                if (act) {
                    double xv[16];
#pragma unroll
                    for (int q = 0; q < 16; ++q) xv[q] = X[(size_t)rr[q] * Cp];
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                        if (!isnan(xv[q])) { N += 1.0; sx = __dadd_rn(sx, xv[q]); sxx = __dadd_rn(sxx, __dmul_rn(xv[q], xv[q])); }
                }
            }
        }
        for (; i + 8 <= end; i += 8) {
            int rr[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) rr[q] = grouped[i + q];
            if (act) {
                double xv[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) xv[q] = X[(size_t)rr[q] * Cp];
                if (ctype == CC_NORMAL) {
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if (!isnan(xv[q])) { N += 1.0; sx = __dadd_rn(sx, xv[q]); sxx = __dadd_rn(sxx, __dmul_rn(xv[q], xv[q])); }
                } else if (sumt) {
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if (!isnan(xv[q])) { N += 1.0; sx = __dadd_rn(sx, xv[q]); sxx = __dadd_rn(sxx, second(xv[q], rr[q])); }
                } else {
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if (!isnan(xv[q])) { N += 1.0; lc[(int)xv[q]]++; }
                }
            }
        }
        for (; i < end; ++i) {
            const int r0 = grouped[i];
            if (act) {
                const double x0 = X[(size_t)r0 * Cp];
                if (!isnan(x0)) {
                    N += 1.0;
                    if (sumt) { sx = __dadd_rn(sx, x0); sxx = __dadd_rn(sxx, second(x0, r0)); }
                    else lc[(int)x0]++;
                }
            }
        }
        if (!act) continue;
        if (p.marginal) {
            double mg;
            if (sumt) mg = sum_marginal(ctype, N, sx, sxx, h0, h1, h2, h3);
            else {
                const double A = kc * h0;
                double lg = 0.0;
                for (int q = 0; q < kc; ++q) lg += lgamma((double)lc[q] + h0);
                mg = lgamma(A) - lgamma(A + N) + lg - kc * lgamma(h0);
            }
            lacc += mg;
        } else {
            const size_t e = (size_t)k * C + c;
            stat_field(st, s, CC_F_N)[e] = N;
            if (ctype == CC_NORMAL) {
                const double gN = normal_G(N, h1, h3);
                const double gM1 = (N >= 1.0) ? normal_G(N - 1.0, h1, h3) : 0.0;
                const NormalCache cc = normal_cache(N, sx, sxx, gN, h0, h1, h2, h3);
                stat_field(st, s, CC_F_SX)[e] = sx;
                stat_field(st, s, CC_F_SXX)[e] = sxx;
                stat_field(st, s, CC_F_MN)[e] = cc.mn;
                stat_field(st, s, CC_F_A)[e] = cc.a;
                stat_field(st, s, CC_F_CST)[e] = cc.cst;
                stat_field(st, s, CC_F_GN)[e] = gN;
                stat_field(st, s, CC_F_GM1)[e] = gM1;
            } else if (sumt) {
                const NormalCache cc = count_cache_ni(ctype, N, sx, h0, h1);
                stat_field(st, s, CC_F_SX)[e] = sx;
                stat_field(st, s, CC_F_SXX)[e] = sxx;
                stat_field(st, s, CC_F_MN)[e] = cc.mn;
                stat_field(st, s, CC_F_A)[e] = cc.a;
                stat_field(st, s, CC_F_CST)[e] = cc.cst;
                stat_field(st, s, CC_F_GN)[e] = 0.0;
                stat_field(st, s, CC_F_GM1)[e] = 0.0;
            } else {
                stat_field(st, s, CC_F_SX)[e] = 0.0;
                stat_field(st, s, CC_F_SXX)[e] = 0.0;
                stat_field(st, s, CC_F_MN)[e] = 0.0;
                stat_field(st, s, CC_F_A)[e] = 0.0;
                stat_field(st, s, CC_F_CST)[e] = log(N + h0 * kc);
                stat_field(st, s, CC_F_GN)[e] = 0.0;
                stat_field(st, s, CC_F_GM1)[e] = 0.0;
                double* cn = st.cnt + ((size_t)s * Kcap + k) * CT + st.catoff[c];
                for (int q = 0; q < kc; ++q) cn[q] = (double)lc[q];
            }
        }
    }
    if (p.marginal) {
        // deterministic cross-warp sum (warp order), one value per column
        __shared__ double s_part[NW][32];
        s_part[warp][lane] = lacc;
        __syncthreads();
        if (warp == 0 && act) {
            double tot = 0.0;
            for (int q = 0; q < NW; ++q) tot += s_part[q][lane];
            st.colL[((size_t)s * C + c) * (st.Vcap + 1) + v] = tot;
        }
    }
}

// ---------------------------------------------------------------------------
// State.logpdf_score (src/crosscat/state.py:384-387): one CTA per chain.
__device__ double crp_marginal_dev(int N, const int32_t* nk, const int32_t* live, int K, double alpha,
                                   int tid, int nt, double* s_red) {
    // crp.py:184-188: K log(alpha) + sum gammaln(n_k) + gammaln(alpha) - gammaln(N + alpha)
    double acc = 0.0;
    for (int i = tid; i < K; i += nt) acc += lgamma((double)nk[live ? live[i] : i]);
    s_red[tid] = acc;
    __syncthreads();
    for (int o = nt >> 1; o > 0; o >>= 1) { if (tid < o) s_red[tid] += s_red[tid + o]; __syncthreads(); }
    const double tot = s_red[0];
    __syncthreads();
    return K * log(alpha) + tot + lgamma(alpha) - lgamma(N + alpha);
}

template <int NT>
__global__ void __launch_bounds__(NT) score_kernel(const cc_state st, double* __restrict__ out) {
    const int s = blockIdx.x, tid = threadIdx.x;
    const int R = st.R, C = st.C, Kcap = st.Kcap, Vcap = st.Vcap, CT = st.CT;
    __shared__ double s_red[NT];
    __shared__ int s_vlist[1];
    (void)s_vlist;
    double total = 0.0;
    // column CRP: counts nv over live view slots
    {
        double acc = 0.0;
        int Kv = 0;
        for (int v = 0; v < Vcap; ++v) if (st.nv[(size_t)s * Vcap + v] > 0) ++Kv;
        for (int v = tid; v < Vcap; v += NT) { const int nvv = st.nv[(size_t)s * Vcap + v]; if (nvv > 0) acc += lgamma((double)nvv); }
        s_red[tid] = acc;
        __syncthreads();
        for (int o = NT >> 1; o > 0; o >>= 1) { if (tid < o) s_red[tid] += s_red[tid + o]; __syncthreads(); }
        const double a = st.col_alpha[s];
        total += Kv * log(a) + s_red[0] + lgamma(a) - lgamma(C + a);
        __syncthreads();
    }
    for (int v = 0; v < Vcap; ++v) {
        if (st.nv[(size_t)s * Vcap + v] <= 0) continue;
        const size_t sv = sv_index(st, s, v);
        const int K = st.nclu[sv];
        const int32_t* live = st.live + sv * Kcap;
        total += crp_marginal_dev(R, st.nk + sv * Kcap, live, K, st.view_alpha[sv], tid, NT, s_red);
        // sum over (cluster, column in view) marginals
        double acc = 0.0;
        for (int e = tid; e < K * C; e += NT) {
            const int pos = e / C, c = e - pos * C;
            if (st.zv[(size_t)s * C + c] != v) continue;
            const int k = live[pos];
            const double* hyp = st.hyp + ((size_t)s * C + c) * 4;
            const size_t idx = (size_t)k * C + c;
            const double N = stat_field(st, s, CC_F_N)[idx];
            if (st.ctype[c] != CC_CATEGORICAL)
                acc += sum_marginal(st.ctype[c], N, stat_field(st, s, CC_F_SX)[idx], stat_field(st, s, CC_F_SXX)[idx],
                                    hyp[0], hyp[1], hyp[2], hyp[3]);
            else
                acc += categorical_marginal(N, st.cnt + ((size_t)s * Kcap + k) * CT + st.catoff[c], st.ncat[c], hyp[0]);
        }
        s_red[tid] = acc;
        __syncthreads();
        for (int o = NT >> 1; o > 0; o >>= 1) { if (tid < o) s_red[tid] += s_red[tid + o]; __syncthreads(); }
        total += s_red[0];
        __syncthreads();
    }
    if (tid == 0) out[s] = total;
}

// cached constants only (statistics untouched): one thread per (cluster position | empty record, column of the view)
__global__ void refresh_kernel(const cc_state st, int n_work, const int32_t* __restrict__ chains,
                               const int32_t* __restrict__ views) {
    const int w = blockIdx.x;
    if (w >= n_work) return;
    const int s = chains[w], v = views[w];
    const size_t sv = sv_index(st, s, v);
    const int C = st.C, Kcap = st.Kcap, K = st.nclu[sv];
    const int32_t* live = st.live + sv * Kcap;
    for (int e = threadIdx.x; e < (K + 1) * C; e += blockDim.x) {
        const int pos = e / C, c = e - pos * C;
        if (st.zv[(size_t)s * C + c] != v) continue;
        const int k = (pos < K) ? live[pos] : (Kcap - 1);
        const size_t idx = (size_t)k * C + c;
        const double* hyp = st.hyp + ((size_t)s * C + c) * 4;
        if (pos >= K) {
            stat_field(st, s, CC_F_N)[idx] = 0.0; stat_field(st, s, CC_F_SX)[idx] = 0.0; stat_field(st, s, CC_F_SXX)[idx] = 0.0;
        }
        const double N = stat_field(st, s, CC_F_N)[idx];
        if (st.ctype[c] == CC_NORMAL) {
            const double gN = normal_G(N, hyp[1], hyp[3]);
            const double gM1 = (N >= 1.0) ? normal_G(N - 1.0, hyp[1], hyp[3]) : 0.0;
            const NormalCache cc = normal_cache(N, stat_field(st, s, CC_F_SX)[idx], stat_field(st, s, CC_F_SXX)[idx], gN,
                                                hyp[0], hyp[1], hyp[2], hyp[3]);
            stat_field(st, s, CC_F_MN)[idx] = cc.mn;
            stat_field(st, s, CC_F_A)[idx] = cc.a;
            stat_field(st, s, CC_F_CST)[idx] = cc.cst;
            stat_field(st, s, CC_F_GN)[idx] = gN;
            stat_field(st, s, CC_F_GM1)[idx] = gM1;
        } else if (st.ctype[c] != CC_CATEGORICAL) {
            const NormalCache cc = count_cache_ni(st.ctype[c], N, stat_field(st, s, CC_F_SX)[idx], hyp[0], hyp[1]);
            stat_field(st, s, CC_F_MN)[idx] = cc.mn;
            stat_field(st, s, CC_F_A)[idx] = cc.a;
            stat_field(st, s, CC_F_CST)[idx] = cc.cst;
        } else {
            stat_field(st, s, CC_F_CST)[idx] = log(N + hyp[0] * st.ncat[c]);
        }
    }
}

// Device form of the reference's debug invariants (View._check_partitions / State._check_partitions,
// src/mixtures/view.py:528-560, src/crosscat/state.py:1162-1184; enabled there by GPMCCDEBUG).
// One CTA per chain; out[s] = 0 or the code of the first violated invariant.
__global__ void __launch_bounds__(256) check_kernel(const cc_state st, int32_t* __restrict__ out, int32_t* gscr) {
    const int s = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int R = st.R, C = st.C, Vcap = st.Vcap, Kcap = st.Kcap;
    extern __shared__ int s_hist_sh[];              // [Kcap] + [Kcap] (in `gscr` [S][2 Kcap] when too large for shared memory)
    int* s_hist = gscr ? gscr + (size_t)s * 2 * Kcap : s_hist_sh;
    int* s_mark = s_hist + Kcap;
    __shared__ int s_bad, s_sum;
    if (tid == 0) { s_bad = 0; s_sum = 0; }
    __syncthreads();
    auto fail = [&](int code) { atomicCAS(&s_bad, 0, code); };
    const int32_t* zv = st.zv + (size_t)s * C;
    const int32_t* nv = st.nv + (size_t)s * Vcap;
    // 1. column partition: every column in a live view, nv is the histogram of zv, sum = C
    for (int v = tid; v < Vcap; v += nt) {
        int cnt = 0;
        for (int c = 0; c < C; ++c) cnt += (zv[c] == v);
        if (cnt != nv[v]) fail(1);
        if (nv[v] > 0 && st.view_id[(size_t)s * Vcap + v] < 0) fail(2);
        if (nv[v] < 0) fail(1);
    }
    for (int c = tid; c < C; c += nt) if (zv[c] < 0 || zv[c] >= Vcap) fail(3);
    if (!(st.col_alpha[s] > 0.0)) fail(4);
    __syncthreads();
    for (int v = 0; v < Vcap && !s_bad; ++v) {
        if (nv[v] <= 0) continue;
        const size_t sv = sv_index(st, s, v);
        const int K = st.nclu[sv];
        const int32_t* live = st.live + sv * Kcap;
        const int32_t* nk = st.nk + sv * Kcap;
        const int32_t* cid = st.cid + sv * Kcap;
        const int32_t* zr = st.zr + sv * R;
        const int32_t* ord = st.order + sv * R;
        uint8_t* flag = st.movedflag + sv * R;
        if (K < 1 || K > Kcap - 1) { if (tid == 0) fail(5); __syncthreads(); break; }
        if (!(st.view_alpha[sv] > 0.0) && tid == 0) fail(4);
        for (int k = tid; k < Kcap; k += nt) { s_hist[k] = 0; s_mark[k] = 0; }
        if (tid == 0) s_sum = 0;
        __syncthreads();
        // 2. live list: valid distinct slots, ascending cluster ids, positive sizes summing to R
        for (int i = tid; i < K; i += nt) {
            const int k = live[i];
            if (k < 0 || k >= Kcap - 1) { fail(6); continue; }
            if (atomicAdd(&s_mark[k], 1) != 0) fail(6);
            if (nk[k] < 1) fail(7);
            if (i > 0 && live[i - 1] >= 0 && live[i - 1] < Kcap && cid[live[i - 1]] >= cid[k]) fail(8);
            atomicAdd(&s_sum, nk[k]);
        }
        __syncthreads();
        if (s_sum != R && tid == 0) fail(9);
        // 3. every row sits in a live cluster and the sizes are the histogram of zr
        for (int r = tid; r < R; r += nt) {
            const int k = zr[r];
            if (k < 0 || k >= Kcap - 1 || s_mark[k] != 1) { fail(10); continue; }
            atomicAdd(&s_hist[k], 1);
        }
        __syncthreads();
        for (int i = tid; i < K; i += nt) { const int k = live[i]; if (k >= 0 && k < Kcap && s_hist[k] != nk[k]) fail(11); }
        // 4. the member order is a permutation of the rows (scratch flags must be clear between kernels)
        for (int i = tid; i < R; i += nt) {
            const int r = ord[i];
            if (r < 0 || r >= R) { fail(12); continue; }
            if (flag[r] != 0) fail(13);
        }
        __syncthreads();
        for (int i = tid; i < R; i += nt) { const int r = ord[i]; if (r >= 0 && r < R) flag[r] = 1; }
        __syncthreads();
        for (int r = tid; r < R; r += nt) if (flag[r] != 1) fail(12);
        __syncthreads();
        for (int r = tid; r < R; r += nt) flag[r] = 0;
        __syncthreads();
        // 5. law of conservation of rowids (view.py:552-560): N of a (cluster, column) = its non-NaN members
        for (int c = 0; c < C && !s_bad; ++c) {
            if (zv[c] != v) continue;
            for (int k = tid; k < Kcap; k += nt) s_hist[k] = 0;
            __syncthreads();
            for (int r = tid; r < R; r += nt) {
                const double x = st.Xrm[(size_t)r * st.Cp + c];
                const int k = zr[r];
                if (!isnan(x) && k >= 0 && k < Kcap) atomicAdd(&s_hist[k], 1);
            }
            __syncthreads();
            for (int i = tid; i < K; i += nt) {
                const int k = live[i];
                if (k >= 0 && k < Kcap && (double)s_hist[k] != stat_field(st, s, CC_F_N)[(size_t)k * C + c]) fail(14);
            }
            __syncthreads();
        }
        __syncthreads();
    }
    __syncthreads();
    if (tid == 0) out[s] = s_bad;
}

__global__ void list_views_kernel(const cc_state st, int32_t* chains, int32_t* views, int32_t* count) {
    // enumerate live (chain, view) pairs (single thread; S*Vcap is small)
    if (blockIdx.x || threadIdx.x) return;
    int n = 0;
    for (int s = 0; s < st.S; ++s)
        for (int v = 0; v < st.Vcap; ++v)
            if (st.nv[(size_t)s * st.Vcap + v] > 0) { chains[n] = s; views[n] = v; ++n; }
    *count = n;
}

}  // namespace

// Host-side launcher shared by cc_rebuild and the column kernel.
int cc_launch_stats(const cc_state* st, int n_work, const int32_t* d_chains, const int32_t* d_views,
                    int marginal, const int32_t* d_cols, const int32_t* d_coloff, const int32_t* d_colcnt,
                    int max_cols, const uint8_t* d_mask, cudaStream_t stream) {
    if (n_work == 0) return CC_OK;
    int32_t* gstart = nullptr;
    CC_CUDA(cudaMallocAsync(&gstart, sizeof(int32_t) * (size_t)n_work * (st->Kcap + 1), stream));
    // as many warps per (chain, view) as 96 KB of per-warp cluster counters allow (8 at Kcap <= 3072)
    const int gmax = 96 * 1024;
    int gw = gmax / ((int)sizeof(int) * st->Kcap);
    int32_t* gcnt = nullptr;
    if (gw < 1) {
        CC_CUDA(cudaMallocAsync(&gcnt, sizeof(int32_t) * (size_t)n_work * st->Kcap, stream));
        group_kernel<<<n_work, 32, 0, stream>>>(*st, n_work, d_chains, d_views, gstart, gcnt);
    } else {
        gw = gw > 8 ? 8 : gw;
        const size_t gbytes = sizeof(int) * (size_t)st->Kcap * gw;
        if (gbytes > 48 * 1024) CC_CUDA(cudaFuncSetAttribute(group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, gmax));
        group_kernel<<<n_work, 32 * gw, gbytes, stream>>>(*st, n_work, d_chains, d_views, gstart, nullptr);
    }
    CC_CUDA(cudaGetLastError());
    if (gcnt) CC_CUDA(cudaFreeAsync(gcnt, stream));
    StatsParams p;
    p.st = *st; p.n_work = n_work; p.chains = d_chains; p.views = d_views; p.gstart = gstart;
    p.marginal = marginal; p.cols = d_cols; p.coloff = d_coloff; p.colcnt = d_colcnt; p.mask = d_mask;
    const int maxcols = marginal ? (d_cols ? max_cols : st->C) : st->C;
    dim3 grid(n_work, (maxcols + 31) / 32);
    stats_kernel<8><<<grid, 8 * 32, 0, stream>>>(p);
    CC_CUDA(cudaGetLastError());
    CC_CUDA(cudaFreeAsync(gstart, stream));
    return CC_OK;
}

extern "C" int cc_rebuild(const cc_state* st, int n_work, const int32_t* chains, const int32_t* views,
                          const uint8_t* d_mask, void* stream_) {
    if (!st) { cc_set_error("cc_rebuild: null state"); return CC_ERR_ARG; }
    cudaStream_t stream = (cudaStream_t)stream_;
    cc_prepare_pool();
    int32_t *dch = nullptr, *dvw = nullptr, *dcount = nullptr;
    int rc = CC_OK;
    if (views == nullptr) {
        const size_t cap = (size_t)st->S * st->Vcap;
        CC_CUDA(cudaMallocAsync(&dch, sizeof(int32_t) * cap, stream));
        CC_CUDA(cudaMallocAsync(&dvw, sizeof(int32_t) * cap, stream));
        CC_CUDA(cudaMallocAsync(&dcount, sizeof(int32_t), stream));
        list_views_kernel<<<1, 1, 0, stream>>>(*st, dch, dvw, dcount);
        CC_CUDA(cudaGetLastError());
        CC_CUDA(cudaMemcpyAsync(&n_work, dcount, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
        CC_CUDA(cudaStreamSynchronize(stream));
    } else {
        if (n_work < 0 || !chains) { cc_set_error("cc_rebuild: bad work list"); return CC_ERR_ARG; }
        if (n_work == 0) return CC_OK;
        CC_CUDA(cudaMallocAsync(&dch, sizeof(int32_t) * n_work, stream));
        CC_CUDA(cudaMallocAsync(&dvw, sizeof(int32_t) * n_work, stream));
        CC_CUDA(cudaMemcpyAsync(dch, chains, sizeof(int32_t) * n_work, cudaMemcpyHostToDevice, stream));
        CC_CUDA(cudaMemcpyAsync(dvw, views, sizeof(int32_t) * n_work, cudaMemcpyHostToDevice, stream));
    }
    rc = cc_launch_stats(st, n_work, dch, dvw, 0, nullptr, nullptr, nullptr, 0, d_mask, stream);
    if (dch) cudaFreeAsync(dch, stream);
    if (dvw) cudaFreeAsync(dvw, stream);
    if (dcount) cudaFreeAsync(dcount, stream);
    return rc;
}

extern "C" int cc_refresh_tables(const cc_state* st, int n_work, const int32_t* chains, const int32_t* views,
                                 void* stream_) {
    if (!st || n_work < 0 || (n_work > 0 && (!chains || !views))) { cc_set_error("cc_refresh_tables: bad arguments"); return CC_ERR_ARG; }
    if (n_work == 0) return CC_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    cc_prepare_pool();
    int32_t *dch = nullptr, *dvw = nullptr;
    CC_CUDA(cudaMallocAsync(&dch, sizeof(int32_t) * n_work, stream));
    CC_CUDA(cudaMallocAsync(&dvw, sizeof(int32_t) * n_work, stream));
    CC_CUDA(cudaMemcpyAsync(dch, chains, sizeof(int32_t) * n_work, cudaMemcpyHostToDevice, stream));
    CC_CUDA(cudaMemcpyAsync(dvw, views, sizeof(int32_t) * n_work, cudaMemcpyHostToDevice, stream));
    refresh_kernel<<<n_work, 256, 0, stream>>>(*st, n_work, dch, dvw);
    CC_CUDA(cudaGetLastError());
    cudaFreeAsync(dch, stream);
    cudaFreeAsync(dvw, stream);
    return CC_OK;
}

extern "C" int cc_check_state(const cc_state* st, int32_t* out_dev, void* stream_) {
    if (!st || !out_dev) { cc_set_error("cc_check_state: null argument"); return CC_ERR_ARG; }
    if (st->S == 0) return CC_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t bytes = sizeof(int) * 2 * (size_t)st->Kcap;
    if (bytes > 40 * 1024) {
        int32_t* gscr = nullptr;
        CC_CUDA(cudaMallocAsync(&gscr, bytes * st->S, stream));
        check_kernel<<<st->S, 256, 0, stream>>>(*st, out_dev, gscr);
        CC_CUDA(cudaGetLastError());
        CC_CUDA(cudaFreeAsync(gscr, stream));
    } else {
        check_kernel<<<st->S, 256, bytes, stream>>>(*st, out_dev, nullptr);
        CC_CUDA(cudaGetLastError());
    }
    return CC_OK;
}

extern "C" int cc_logpdf_score(const cc_state* st, double* out, void* stream_) {
    if (!st || !out) { cc_set_error("cc_logpdf_score: null argument"); return CC_ERR_ARG; }
    score_kernel<256><<<st->S, 256, 0, (cudaStream_t)stream_>>>(*st, out);
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}
