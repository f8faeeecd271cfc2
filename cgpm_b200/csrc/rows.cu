// Row-reassignment Gibbs sweep: View.transition_rows (src/mixtures/view.py:247-252)
// = for every row in a permuted order: _gibbs_transition_row (view.py:393-406)
//   -> Crp.gibbs_tables / gibbs_logps (src/primitives/crp.py:111-146)
//   -> _logpdf_row_gibbs / _logpdf_cell_gibbs (view.py:408-422)
//   -> gu.log_pflip (src/utils/general.py:123-139)
//   -> _migrate_row (view.py:424-429).
//
// One persistent CTA per (chain, view).  The scan over rows is strictly
// sequential, exactly as in the reference; parallelism inside a step is over
// (candidate cluster x column).  Warp roles:
//   * NT consumer threads: groups of G lanes evaluate one candidate cluster
//     each (lanes over the view's columns, shuffle-reduce), then every warp
//     redundantly does the max / exp / inclusive-scan / inverse-CDF draw over
//     the candidates (no second barrier), then threads over columns apply the
//     move to the sufficient statistics.
//   * one producer warp: stages the rows of the next batch into shared memory
//     (whole 16-byte-aligned rows by 1-D bulk copies -- the TMA engine -- or
//     per-cell cp.async gathers when rows are too wide), together with the
//     row ids and their current cluster, behind a full/empty mbarrier pair.
// The per-(cluster, column) tables (sufficient statistics + the cached
// predictive constants) live in shared memory when they fit and in global
// memory (L2) otherwise; the kernel migrates them from shared to global if the
// number of clusters outgrows shared memory mid-sweep.
//
// Exactness: sufficient statistics are updated with the same IEEE operations,
// in the same order, as the reference (x*x is rounded before it is added,
// normal.py:69; scoring a row against its own cluster does `-= x` then
// `+= x`, view.py:416-419), so they are bit-identical for the same sequence of
// moves.  The predictive is evaluated in the algebraically equal Student-t form
// (see cc_common.cuh) in fp64.
#include <cstdlib>
#include "cc_common.cuh"

namespace {

struct RowsParams {
    cc_state st;
    const cc_row_work* work;
    const int32_t* perm;
    const double* u;
    double* trace;
    int trace_stride;
    uint64_t seed, counter;
    int B;          // rows per staged batch (<= 32)
    int bulk;       // 1: stage whole rows with cp.async.bulk
    int smem_bytes; // dynamic shared memory given to the CTA
};

struct Tab {            // table accessor (shared or global)
    double* f[CC_NFIELD];
    double* cnt;
    int kstride;        // elements between clusters in a field
    int cstride;        // elements between clusters in cnt
    int global;         // 0: shared-memory tables
};

template <int NT>
__device__ __forceinline__ void consumer_sync_t() {
    if (NT == 32) __syncwarp();                       // single consumer warp: no block barrier at all
    else asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
}
#define consumer_sync(nt) consumer_sync_t<NT>()

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
    c.cst = gN - 0.5 * log(sn);
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

template <int NT, int MINB>
__global__ void __launch_bounds__(NT + 32, MINB) rows_kernel(const RowsParams p) {
    const cc_state& st = p.st;
    const cc_row_work wk = p.work[blockIdx.x];
    const int s = wk.chain, v = wk.view, n = wk.n;
    const int R = st.R, C = st.C, Cp = st.Cp, Kcap = st.Kcap, CT = st.CT;
    const int tid = threadIdx.x;
    const bool producer = tid >= NT;
    const int B = p.B;
    const size_t sv = sv_index(st, s, v);

    extern __shared__ __align__(16) unsigned char smem_raw[];
    // ---- fixed shared region --------------------------------------------------
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem_raw);      // [2]
    uint64_t* bar_empty = bar_full + 2;                               // [2]
    int* s_scal = reinterpret_cast<int*>(bar_empty + 2);              // [16] scalars
    // s_scal: 0 K, 1 D, 2 free head, 3 nmoved, 4 own position, 5 global mode, 6 CTv, 7 error
    int* s_live = s_scal + 16;                                        // [Kcap]
    int* s_nk = s_live + Kcap;                                        // [Kcap]
    double* s_lp = reinterpret_cast<double*>(s_nk + Kcap + ((2 * Kcap) & 1)); // [Kcap+1] (8-byte aligned)
    // ---- view columns -----------------------------------------------------------
    // count the view's columns (all threads, redundantly cheap) ---------------
    const int32_t* zv = st.zv + (size_t)s * C;
    __shared__ int s_D_tmp;
    if (tid == 0) {
        int D = 0;
        for (int c = 0; c < C; ++c) D += (zv[c] == v);
        s_D_tmp = D;
        mbar_init(&bar_full[0], 33); mbar_init(&bar_full[1], 33);
        mbar_init(&bar_empty[0], 1); mbar_init(&bar_empty[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int D = s_D_tmp;
    double* after_lp = s_lp + (Kcap + 1);
    int* s_col = reinterpret_cast<int*>(after_lp);                    // [D]
    int* s_cto = s_col + D;                                           // [D] local count offsets
    int* s_rowidx = s_cto + D;                                        // [2][32]
    int* s_zrow = s_rowidx + 64;                                      // [2][32]
    size_t off = (size_t)((unsigned char*)(s_zrow + 64) - smem_raw);
    off = (off + 15) & ~(size_t)15;
    double* s_u = reinterpret_cast<double*>(smem_raw + off);          // [2][32] the row's uniform
    off += 64 * sizeof(double);
    const int xstride = p.bulk ? Cp : ((D + 1) & ~1);
    double* s_x = reinterpret_cast<double*>(smem_raw + off);          // [2][B][xstride]
    off += (size_t)2 * B * xstride * sizeof(double);
    double* s_tab = reinterpret_cast<double*>(smem_raw + off);
    const size_t tab_avail = (p.smem_bytes > (int)off) ? (size_t)p.smem_bytes - off : 0;

    if (tid == 0) {
        int j = 0, ct = 0;
        for (int c = 0; c < C; ++c)
            if (zv[c] == v) {
                s_col[j] = c;
                s_cto[j] = ct;
                ct += st.ncat[c];
                ++j;
            }
        s_scal[6] = ct;
        s_scal[3] = 0;
        s_scal[7] = 0;
    }
    // live list / counts ----------------------------------------------------------
    const int K0 = st.nclu[sv];
    int32_t* g_live = st.live + sv * Kcap;
    int32_t* g_nk = st.nk + sv * Kcap;
    int32_t* g_cid = st.cid + sv * Kcap;
    for (int i = tid; i < Kcap; i += blockDim.x) {
        s_nk[i] = 0;
        s_live[i] = (i < K0) ? g_live[i] : -1;
    }
    __syncthreads();
    for (int i = tid; i < K0; i += blockDim.x) s_nk[s_live[i]] = g_nk[s_live[i]];
    __syncthreads();
    const int CTv = s_scal[6];
    // shared-table capacity: Kc cluster slots (the prior record sits at slot Kc-1)
    const size_t per_cluster = (size_t)(CC_NFIELD * D + CTv) * sizeof(double);
    int Kc = (int)(tab_avail / (per_cluster ? per_cluster : 1));
    if (Kc > Kcap) Kc = Kcap;
    int maxslot = -1;
    if (tid == 0) {
        for (int i = 0; i < K0; ++i) maxslot = max(maxslot, s_live[i]);
        int fh = 0;
        while (fh < Kcap - 1 && s_nk[fh] != 0) ++fh;
        s_scal[0] = K0;
        s_scal[1] = D;
        s_scal[2] = fh;
        s_scal[5] = (Kc < 3 || maxslot + 2 > Kc - 1) ? 1 : 0;
    }
    __syncthreads();

    Tab T;
    auto set_tab = [&](int global) {
        T.global = global;
        if (global) {
            for (int f = 0; f < CC_NFIELD; ++f) T.f[f] = stat_field(st, s, f);
            T.cnt = st.cnt + (size_t)s * Kcap * CT;
            T.kstride = C;
            T.cstride = CT;
        } else {
            for (int f = 0; f < CC_NFIELD; ++f) T.f[f] = s_tab + (size_t)f * Kc * D;
            T.cnt = s_tab + (size_t)CC_NFIELD * Kc * D;
            T.kstride = D;
            T.cstride = CTv;
        }
    };
    set_tab(s_scal[5]);
    // slot index of the prior (empty-cluster) record in the active table
    auto prior_slot = [&]() { return T.global ? (Kcap - 1) : (Kc - 1); };
    auto jidx = [&](int j) { return T.global ? s_col[j] : j; };
    auto cidx = [&](int j) { return T.global ? st.catoff[s_col[j]] : s_cto[j]; };

    // load shared tables ---------------------------------------------------------
    if (!T.global) {
        for (int e = tid; e < (K0 + 1) * D; e += blockDim.x) {
            const int pp = e / D, j = e - pp * D;
            const int gk = (pp < K0) ? s_live[pp] : (Kcap - 1);
            const int sk = (pp < K0) ? gk : (Kc - 1);
            const int c = s_col[j];
#pragma unroll
            for (int f = 0; f < CC_NFIELD; ++f)
                T.f[f][(size_t)sk * D + j] = stat_field(st, s, f)[(size_t)gk * C + c];
            if (st.ctype[c] == CC_CATEGORICAL && pp < K0) {
                const double* src = st.cnt + ((size_t)s * Kcap + gk) * CT + st.catoff[c];
                double* dst = T.cnt + (size_t)sk * CTv + s_cto[j];
                for (int q = 0; q < st.ncat[c]; ++q) dst[q] = src[q];
            }
        }
    }
    __syncthreads();

    const double alpha = st.view_alpha[sv];
    const double* hyp = st.hyp + (size_t)s * C * 4;
    int32_t* zr = st.zr + sv * R;
    const int32_t* order = st.order + sv * R;
    int32_t* moved = st.moved + sv * R;
    uint8_t* movedflag = st.movedflag + sv * R;
    const Philox rng(p.seed);
    const int nbatch = (n + B - 1) / B;

    if (producer) {
        // =========================== producer warp ===========================
        const int lane = tid - NT;
        uint32_t bits = 1;
        while ((1u << bits) < (uint32_t)n && bits < 31) ++bits;
        const uint4 pkey = rng(p.counter, stream_id(2, s, v));
        for (int b = 0; b < nbatch; ++b) {
            const int buf = b & 1;
            if (b >= 2) mbar_wait(&bar_empty[buf], ((b >> 1) - 1) & 1);
            const int i = b * B + lane;
            const bool act = lane < B && i < n;
            int r = 0;
            if (act) {
                int idx;
                if (wk.perm_off >= 0) idx = p.perm[wk.perm_off + i];
                else idx = (int)keyed_perm((uint32_t)i, (uint32_t)n, bits, pkey);
                r = wk.perm_is_index ? order[idx] : idx;
                s_rowidx[buf * 32 + lane] = r;
                double uu;
                if (wk.u_off >= 0) uu = p.u[wk.u_off + i];
                else { const uint4 w4 = rng(p.counter * 0x100000000ull + (uint64_t)i, stream_id(1, s, v)); uu = u01_from_words(w4.x, w4.y); }
                s_u[buf * 32 + lane] = uu;
            }
            const int nrows = min(B, n - b * B);
            __syncwarp();
            if (lane == 0) mbar_arrive_expect_tx(&bar_full[buf], p.bulk ? (uint32_t)(nrows * Cp * 8) : 0u);
            __syncwarp();
            if (act) {
                cp_async_4(&s_zrow[buf * 32 + lane], &zr[r]);
                double* dst = s_x + (size_t)(buf * B + lane) * xstride;
                const double* src = st.Xrm + (size_t)r * Cp;
                if (p.bulk) bulk_g2s(dst, src, (uint32_t)(Cp * 8), &bar_full[buf]);
                else for (int j = 0; j < D; ++j) cp_async_8(&dst[j], &src[s_col[j]]);
            }
            cp_async_mbar_arrive_noinc(&bar_full[buf]);
        }
        return;
    }

    // ============================== consumers ===============================
    const int lane = tid & 31;
    int G = 1;
    while (G < D && G < 32) G <<= 1;
    const int gid = tid / G, gl = tid & (G - 1), ngroups = NT / G;

    for (int i = 0; i < n; ++i) {
        const int b = i / B, slot = i - b * B, buf = b & 1;
        if (slot == 0) mbar_wait(&bar_full[buf], (b >> 1) & 1);
        const int r = s_rowidx[buf * 32 + slot];
        const int z = s_zrow[buf * 32 + slot];
        const double* xrow = s_x + (size_t)(buf * B + slot) * xstride;
        const int K = s_scal[0];
        const int nkz = s_nk[z];
        const bool singleton = nkz == 1;
        const int ncand = K + (singleton ? 0 : 1);
        const int pslot = prior_slot();

        // ---- per-candidate data log-likelihood ----------------------------------
        for (int base = 0; base < ncand; base += ngroups) {      // trip count uniform per warp (full-mask shuffles)
            const int pc = base + gid;
            const bool active = pc < ncand;
            const int k = (active && pc < K) ? s_live[pc] : pslot;
            const bool own = active && (pc < K) && (k == z);
            double acc = 0.0;
            for (int j = gl; active && j < D; j += G) {
                const int c = s_col[j];
                const double x = xrow[p.bulk ? c : j];
                if (isnan(x)) continue;                                  // dim.py:130-132
                const size_t e = (size_t)k * T.kstride + jidx(j);
                if (st.ctype[c] == CC_NORMAL) {
                    const double nu = __ldg(&hyp[c * 4 + 3]);
                    if (!own) {
                        acc += normal_predictive_cached(x, T.f[CC_F_MN][e], T.f[CC_F_A][e], T.f[CC_F_CST][e],
                                                        T.f[CC_F_N][e], nu);
                    } else {
                        // view.py:416-419: unincorporate, score, incorporate
                        const double m = __ldg(&hyp[c * 4 + 0]), rr = __ldg(&hyp[c * 4 + 1]), ss = __ldg(&hyp[c * 4 + 2]);
                        const double N0 = T.f[CC_F_N][e], sx0 = T.f[CC_F_SX][e], sxx0 = T.f[CC_F_SXX][e];
                        const double xx = __dmul_rn(x, x);
                        const double N1 = N0 - 1.0, sx1 = __dsub_rn(sx0, x), sxx1 = __dsub_rn(sxx0, xx);
                        const NormalCache c1 = normal_cache_fast(N1, sx1, sxx1, T.f[CC_F_GM1][e], m, rr, ss, nu);
                        acc += normal_predictive_cached(x, c1.mn, c1.a, c1.cst, N1, nu);
                        const double sx2 = __dadd_rn(sx1, x), sxx2 = __dadd_rn(sxx1, xx);
                        if (sx2 != sx0 || sxx2 != sxx0) {
                            T.f[CC_F_SX][e] = sx2;
                            T.f[CC_F_SXX][e] = sxx2;
                            const NormalCache c2 = normal_cache_fast(N0, sx2, sxx2, T.f[CC_F_GN][e], m, rr, ss, nu);
                            T.f[CC_F_MN][e] = c2.mn; T.f[CC_F_A][e] = c2.a; T.f[CC_F_CST][e] = c2.cst;
                        }
                    }
                } else {
                    const double a = __ldg(&hyp[c * 4 + 0]);
                    const int xi = (int)x;
                    const int kc = st.ncat[c];
                    if (pc >= K) {
                        acc += log(a) - log(a * kc);                     // empty cluster
                    } else {
                        const double cntx = T.cnt[(size_t)k * T.cstride + cidx(j) + xi];
                        if (!own) acc += log(a + cntx) - T.f[CC_F_CST][e];
                        else acc += log(a + (cntx - 1.0)) - log((T.f[CC_F_N][e] - 1.0) + a * kc);
                    }
                }
            }
            for (int o = G >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (gl == 0 && active) {
                s_lp[pc] = acc;
                if (own) s_scal[4] = pc;
            }
        }
        consumer_sync(NT);

        // ---- draw (every warp redundantly; identical arithmetic) ------------------
        const double uu = s_u[buf * 32 + slot];
        double mx = -INFINITY;
        for (int pc = lane; pc < ncand; pc += 32) mx = fmax(mx, s_lp[pc]);
        mx = warp_max(mx);
        auto weight = [&](int pc) -> double {
            if (pc >= K) return alpha;
            const int k = s_live[pc];
            if (k == z) return singleton ? alpha : (double)(nkz - 1);
            return (double)s_nk[k];
        };
        int choice = -1;
        if (ncand <= 32) {
            const double e = (lane < ncand) ? weight(lane) * exp(s_lp[lane] - mx) : 0.0;
            const double inc = warp_incl_scan(e, lane);
            const double tot = __shfl_sync(0xffffffffu, inc, 31);
            const unsigned hit = __ballot_sync(0xffffffffu, lane < ncand && inc > uu * tot);
            choice = hit ? (__ffs(hit) - 1) : -1;
        } else {
            double carry = 0.0;
            for (int base = 0; base < ncand; base += 32) {
                const int pc = base + lane;
                const double e = (pc < ncand) ? weight(pc) * exp(s_lp[pc] - mx) : 0.0;
                carry += __shfl_sync(0xffffffffu, warp_incl_scan(e, lane), 31);
            }
            const double tot = carry;
            carry = 0.0;
            for (int base = 0; base < ncand && choice < 0; base += 32) {
                const int pc = base + lane;
                const double e = (pc < ncand) ? weight(pc) * exp(s_lp[pc] - mx) : 0.0;
                const double inc = warp_incl_scan(e, lane) + carry;
                const unsigned hit = __ballot_sync(0xffffffffu, pc < ncand && inc > uu * tot);
                if (hit) choice = base + __ffs(hit) - 1;
                carry = __shfl_sync(0xffffffffu, inc, 31);
            }
        }
        if (choice < 0) {                       // NaN / all -inf: keep the row, flag the chain
            choice = s_scal[4];
            if (tid == 0) s_scal[7] = CC_ERR_ARG;
        }
        if (p.trace && wk.trace_off >= 0 && tid == 0) {
            double* tr = p.trace + wk.trace_off + (size_t)i * p.trace_stride;
            tr[0] = (double)r;
            tr[1] = (double)ncand;
            tr[2] = (double)((choice < K) ? g_cid[s_live[choice]] : (g_cid[s_live[K - 1]] + 1));
            for (int pc = 0; pc < ncand && pc + 3 < p.trace_stride; ++pc) tr[3 + pc] = s_lp[pc] + log(weight(pc));
        }

        // ---- apply the move (view.py:424-429) -------------------------------------
        bool newc = choice >= K;
        int zb = newc ? s_scal[2] : s_live[choice];
        const int ownpos = s_scal[4];
        if (newc && zb >= Kcap - 1) {            // no free cluster slot: keep the row, flag the chain
            if (tid == 0) s_scal[7] = CC_ERR_CAPACITY;
            newc = false;
            zb = z;
        }
        const bool move = zb != z;
        if (move && newc && !T.global && zb >= Kc - 1) {
            // the new slot does not fit in shared memory: migrate the tables to
            // global memory (L2) and continue there
            for (int e = tid; e < K * D; e += NT) {
                const int pp = e / D, j = e - pp * D;
                const int k = s_live[pp], c = s_col[j];
#pragma unroll
                for (int f = 0; f < CC_NFIELD; ++f)
                    stat_field(st, s, f)[(size_t)k * C + c] = T.f[f][(size_t)k * D + j];
                if (st.ctype[c] == CC_CATEGORICAL) {
                    double* dst = st.cnt + ((size_t)s * Kcap + k) * CT + st.catoff[c];
                    const double* src = T.cnt + (size_t)k * CTv + s_cto[j];
                    for (int q = 0; q < st.ncat[c]; ++q) dst[q] = src[q];
                }
            }
            consumer_sync(NT);
            set_tab(1);
        }
        if (move) {
            const bool del_old = singleton;
            for (int t = tid; t < 2 * D; t += NT) {
                const int j = (t < D) ? t : t - D;
                const bool is_old = t < D;
                if (is_old && del_old) continue;
                const int c = s_col[j];
                const double x = xrow[p.bulk ? c : j];
                const int k = is_old ? z : zb;
                const size_t e = (size_t)k * T.kstride + jidx(j);
                const bool isnan_x = isnan(x);
                if (st.ctype[c] == CC_NORMAL) {
                    const double m = __ldg(&hyp[c * 4 + 0]), rr = __ldg(&hyp[c * 4 + 1]), ss = __ldg(&hyp[c * 4 + 2]),
                                 nu = __ldg(&hyp[c * 4 + 3]);
                    double N, sx, sxx, gN, gM1;
                    if (is_old) {
                        if (isnan_x) continue;
                        const double xx = __dmul_rn(x, x);
                        N = T.f[CC_F_N][e] - 1.0;
                        sx = __dsub_rn(T.f[CC_F_SX][e], x);
                        sxx = __dsub_rn(T.f[CC_F_SXX][e], xx);
                        gN = T.f[CC_F_GM1][e];
                        gM1 = (N >= 1.0) ? normal_G(N - 1.0, rr, nu) : 0.0;
                    } else if (newc) {
                        // fresh cluster object: 0 + x, 0 + x*x (normal.py:51-53,64-70)
                        const size_t ep = (size_t)pslot * T.kstride + jidx(j);
                        if (isnan_x) { N = 0.0; sx = 0.0; sxx = 0.0; gN = T.f[CC_F_GN][ep]; gM1 = 0.0; }
                        else {
                            N = 1.0; sx = __dadd_rn(0.0, x); sxx = __dadd_rn(0.0, __dmul_rn(x, x));
                            gM1 = T.f[CC_F_GN][ep];
                            gN = normal_G(1.0, rr, nu);
                        }
                    } else {
                        if (isnan_x) continue;
                        N = T.f[CC_F_N][e] + 1.0;
                        sx = __dadd_rn(T.f[CC_F_SX][e], x);
                        sxx = __dadd_rn(T.f[CC_F_SXX][e], __dmul_rn(x, x));
                        gM1 = T.f[CC_F_GN][e];
                        gN = normal_G(N, rr, nu);
                    }
                    const NormalCache cc = normal_cache_fast(N, sx, sxx, gN, m, rr, ss, nu);
                    T.f[CC_F_N][e] = N; T.f[CC_F_SX][e] = sx; T.f[CC_F_SXX][e] = sxx;
                    T.f[CC_F_MN][e] = cc.mn; T.f[CC_F_A][e] = cc.a; T.f[CC_F_CST][e] = cc.cst;
                    T.f[CC_F_GN][e] = gN; T.f[CC_F_GM1][e] = gM1;
                } else {
                    const double a = __ldg(&hyp[c * 4 + 0]);
                    const int kc = st.ncat[c];
                    double* cn = T.cnt + (size_t)k * T.cstride + cidx(j);
                    const bool fresh = !is_old && newc;
                    double N = fresh ? 0.0 : T.f[CC_F_N][e];
                    if (fresh) for (int q = 0; q < kc; ++q) cn[q] = 0.0;
                    if (isnan_x && !fresh) continue;
                    if (!isnan_x) {
                        const int xi = (int)x;
                        if (is_old) { cn[xi] -= 1.0; N -= 1.0; }
                        else { cn[xi] += 1.0; N += 1.0; }
                    }
                    T.f[CC_F_N][e] = N;
                    T.f[CC_F_CST][e] = log(N + a * kc);
                }
            }
        }
        consumer_sync(NT);           // every warp has finished its draw; tables are updated
        if (move) {
            // bookkeeping of the row CRP (crp.py:45-59) by one thread
            if (tid == 0) {
                zr[r] = zb;
                const int nm = s_scal[3];
                moved[nm] = r;
                movedflag[r] = 1;
                s_scal[3] = nm + 1;
                s_nk[z] = nkz - 1;
                if (newc) {
                    g_cid[zb] = g_cid[s_live[K - 1]] + 1;                  // crp.py:142 max(counts)+1
                    s_live[K] = zb;
                    s_nk[zb] = 1;
                    s_scal[0] = K + 1;
                    int fh = zb + 1;
                    while (fh < Kcap - 1 && s_nk[fh] != 0) ++fh;
                    s_scal[2] = fh;
                } else {
                    s_nk[zb] += 1;
                }
                if (singleton) {
                    // the emptied cluster leaves the ascending-id list; its slot is free again
                    for (int q = ownpos; q < K - 1; ++q) s_live[q] = s_live[q + 1];
                    s_live[K - 1] = -1;
                    s_scal[0] = K - 1;
                    if (z < s_scal[2]) s_scal[2] = z;
                }
            }
            consumer_sync(NT);
        }
        if ((slot == B - 1 || i == n - 1) && tid == 0) mbar_arrive(&bar_empty[buf]);
    }

    // ---- write back ---------------------------------------------------------------
    consumer_sync(NT);
    const int K = s_scal[0];
    if (!T.global) {
        for (int e = tid; e < K * D; e += NT) {
            const int pp = e / D, j = e - pp * D;
            const int k = s_live[pp], c = s_col[j];
#pragma unroll
            for (int f = 0; f < CC_NFIELD; ++f)
                stat_field(st, s, f)[(size_t)k * C + c] = T.f[f][(size_t)k * D + j];
            if (st.ctype[c] == CC_CATEGORICAL) {
                double* dst = st.cnt + ((size_t)s * Kcap + k) * CT + st.catoff[c];
                const double* src = T.cnt + (size_t)k * CTv + s_cto[j];
                for (int q = 0; q < st.ncat[c]; ++q) dst[q] = src[q];
            }
        }
    }
    for (int i = tid; i < Kcap; i += NT) {
        g_live[i] = (i < K) ? s_live[i] : -1;
        g_nk[i] = s_nk[i];
    }
    if (tid == 0) {
        st.nclu[sv] = K;
        if (s_scal[7]) st.err[s] = s_scal[7];
    }
    // ---- member order of the view's CRP: re-seated rows move to the end, in the
    //      order they were re-seated (OrderedDict pop + re-insert, crp.py:52,55) -----
    const int nmoved = s_scal[3];
    if (nmoved > 0) {
        int32_t* ord = st.order + sv * R;
        __shared__ int s_wsum[32];
        __shared__ int s_base;
        if (tid == 0) s_base = 0;
        consumer_sync(NT);
        for (int i0 = 0; i0 < R; i0 += NT) {
            const int i = i0 + tid;
            int rr = -1, keep = 0;
            if (i < R) { rr = ord[i]; keep = !movedflag[rr]; }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            const int wpre = __popc(bal & ((1u << lane) - 1u));
            if (lane == 0) s_wsum[tid >> 5] = __popc(bal);
            consumer_sync(NT);
            int wbase = 0;
            for (int w = 0; w < (tid >> 5); ++w) wbase += s_wsum[w];
            const int base = s_base;
            consumer_sync(NT);
            if (keep) ord[base + wbase + wpre] = rr;
            if (tid == NT - 1) s_base = base + wbase + wpre + keep;
            consumer_sync(NT);
        }
        const int base = s_base;
        for (int i = tid; i < nmoved; i += NT) {
            const int rr = moved[i];
            ord[base + i] = rr;
            movedflag[rr] = 0;
        }
    }
}

}  // namespace

extern "C" int cc_transition_rows(const cc_state* st, int n_work, const cc_row_work* work,
                                  const int32_t* perm, const double* u, double* trace, int trace_stride,
                                  uint64_t seed, uint64_t counter, void* stream_) {
    if (!st || n_work < 0 || (n_work > 0 && !work)) { cc_set_error("cc_transition_rows: bad arguments"); return CC_ERR_ARG; }
    if (n_work == 0) return CC_OK;
    if (st->Kcap > 4096 || st->Kcap < 4) { cc_set_error("cc_transition_rows: Kcap must be in [4, 4096], got %d", st->Kcap); return CC_ERR_ARG; }
    cudaStream_t stream = (cudaStream_t)stream_;
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
    cc_row_work* dwork = nullptr;
    CC_CUDA(cudaMallocAsync(&dwork, sizeof(cc_row_work) * n_work, stream));
    CC_CUDA(cudaMemcpyAsync(dwork, work, sizeof(cc_row_work) * n_work, cudaMemcpyHostToDevice, stream));

    int dev = 0, nsm = 148, max_smem = 0;
    CC_CUDA(cudaGetDevice(&dev));
    CC_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    CC_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    RowsParams p;
    p.st = *st; p.work = dwork; p.perm = perm; p.u = u; p.trace = trace; p.trace_stride = trace_stride;
    p.seed = seed; p.counter = counter;
    // rows are staged whole (bulk copy) when two batches of them stay small
    const int rowbytes = st->Cp * 8;
    p.bulk = (rowbytes <= 4096) ? 1 : 0;
    p.B = 32;
    if (p.bulk) while (p.B > 4 && 2 * p.B * rowbytes > 32 * 1024) p.B >>= 1;
    // consumer threads: one warp (no block barriers) for narrow views, four warps otherwise;
    // override with CGPM_B200_ROWS_NT=32|128
    int nt = 128;
    if (const char* e = getenv("CGPM_B200_ROWS_NT")) nt = atoi(e) == 32 ? 32 : 128;
    // CTAs per SM wanted: enough to co-schedule all work items
    const int maxb = (nt == 32) ? 8 : 3;
    int per_sm = (n_work + nsm - 1) / nsm;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > maxb) per_sm = maxb;
    int budget = (max_smem + 1024) / per_sm - 1024 - 256;      // 1 KB per-CTA system reservation
    const int xbytes = 2 * p.B * (p.bulk ? st->Cp : ((st->C + 1) & ~1)) * 8;
    const int fixed = 64 + 64 + st->Kcap * 8 + (st->Kcap + 2) * 8 + st->C * 8 + 512 + 64 + 512 + xbytes;
    if (budget < fixed + 4096) budget = fixed + 4096;
    if (budget > max_smem) budget = max_smem;
    if (fixed > max_smem) { cc_set_error("cc_transition_rows: shared memory need %d exceeds %d", fixed, max_smem); return CC_ERR_ARG; }
    p.smem_bytes = budget;
    if (nt == 32) {
        CC_CUDA(cudaFuncSetAttribute(rows_kernel<32, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, budget));
        rows_kernel<32, 8><<<n_work, 64, budget, stream>>>(p);
    } else {
        CC_CUDA(cudaFuncSetAttribute(rows_kernel<128, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, budget));
        rows_kernel<128, 3><<<n_work, 160, budget, stream>>>(p);
    }
    CC_CUDA(cudaGetLastError());
    CC_CUDA(cudaFreeAsync(dwork, stream));
    return CC_OK;
}
