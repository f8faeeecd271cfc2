// Row-reassignment Gibbs sweep: View.transition_rows (src/mixtures/view.py:247-252)
// = for every row in a permuted order: _gibbs_transition_row (view.py:393-406)
//   -> Crp.gibbs_tables / gibbs_logps (src/primitives/crp.py:111-146)
//   -> _logpdf_row_gibbs / _logpdf_cell_gibbs (view.py:408-422)
//   -> gu.log_pflip (src/utils/general.py:123-139)
//   -> _migrate_row (view.py:424-429).
//
// One persistent CTA per (chain, view).  The scan over rows is strictly
// sequential, exactly as in the reference; parallelism inside a step is over
// (candidate cluster x column).  Three kernels share the per-row arithmetic
// (rows_common.cuh) and are launched side by side by cc_transition_rows:
//   rows_kernel<128, 3, false>   views of up to 16 columns: the sequential sweep below
//   rows_kernel<224, 2, true>    wider views: two rows in flight (rows_sweep_pipe.cuh)
//   rows_narrow_kernel           views of one or two normal columns: one warp, lane = candidate
// Warp roles of rows_kernel:
//   * NT consumer threads.  Step 1: groups of G lanes evaluate one candidate
//     cluster each (lanes over the view's columns, shuffle-reduce).  Step 2:
//     warp 0 does the max / w*exp / inclusive-scan / inverse-CDF draw over the
//     candidates while the group that scored the row's own cluster replays the
//     reference's re-incorporation.  Step 3 (only if the row moves): threads
//     over columns update the two clusters' tables, one thread the CRP.
//   * one producer warp: stages the rows of the next batch into shared memory
//     (whole 16-byte-aligned rows by 1-D bulk copies -- the TMA engine -- or
//     per-cell cp.async gathers when rows are too wide), together with the row
//     ids, their current cluster and their uniform, behind a full/empty
//     mbarrier pair.
// The per-(cluster, column) tables (sufficient statistics + the cached
// predictive constants) live in shared memory when they fit and in global
// memory (L2) otherwise; the sweep is compiled once per address space and
// migrates from the shared to the global variant if the number of clusters
// outgrows shared memory mid-sweep.
//
// Exactness: sufficient statistics are updated with the same IEEE operations,
// in the same order, as the reference (x*x is rounded before it is added,
// normal.py:69; scoring a row against its own cluster does `-= x` then
// `+= x`, view.py:416-419), so they are bit-identical for the same sequence of
// moves.  The predictive is evaluated in the algebraically equal Student-t form
// (see cc_common.cuh) in fp64.
#include <algorithm>
#include <cstdlib>
#include <vector>
#include "cc_common.cuh"
#include "rows_common.cuh"
#include "rows_sweep_pipe.cuh"

namespace {


// One pass over rows [i_begin, n).  Returns n when finished, or the index of the row whose
// move needs a cluster slot that does not fit in shared memory (GT = false only): the caller
// migrates the tables to global memory and continues there with the same row.
template <int NT, bool GT>
__device__ int sweep_seq(const Ctx& c, int i_begin, bool resume) {
    const Tab<GT> T(c);
    const int tid = threadIdx.x, lane = tid & 31;
    const int D = c.D, B = c.B, n = c.n, Kcap = c.st.Kcap;
    int G = 1;
    while (G < D && G < 32) G <<= 1;
    const int gid = tid / G, gl = tid & (G - 1), ngroups = NT / G;
    const int pslot = T.prior();
    // per-column hyperparameters / types: shared memory, or global memory for very wide views
    const double* ghyp = c.st.hyp + (size_t)c.s * c.st.C * 4;
    const bool wide = c.wide != 0;
    const bool allnorm = c.allnorm != 0;
    auto H = [&](int f, int j) -> double { return c.wide ? __ldg(&ghyp[c.col[j] * 4 + f]) : c.hyp[f * D + j]; };
    auto CTYPE = [&](int j) -> int { return c.wide ? c.st.ctype[c.col[j]] : c.ctype[j]; };
    auto NCAT = [&](int j) -> int { return c.wide ? c.st.ncat[c.col[j]] : c.ncat[j]; };

    const int lgB = 31 - __clz(B);                 // B is a power of two
    const int gmask = ngroups - 1;                 // highest group index (ngroups is 7, 14, ... for NT = 224)
    for (int i = i_begin; i < n; ++i) {
        const int b = i >> lgB, slot = i & (B - 1), buf = b & 1;
        if (slot == 0 || i == i_begin) mbar_wait(&c.bar_full[buf], (b >> 1) & 1);
        const int r = c.rowidx[buf * 32 + slot];
        const int z = c.zrow[buf * 32 + slot];
        const double* xrow = c.x + (buf * B + slot) * c.xstride;
        const int K = c.scal[0];
        const int nkz = c.nk[z];
        const bool singleton = nkz == 1;
        const int ncand = K + (singleton ? 0 : 1);

        // a row handed over by the shared-memory sweep was already scored and drawn (steps 1-2)
        const bool skip12 = resume && i == i_begin;
        if (!skip12) {
        // ---- step 1: per-candidate data log-likelihood --------------------------------
        for (int base = 0; base < ncand; base += ngroups) {      // trip count uniform per warp (full-mask shuffles)
            const int pc = base + (gmask - gid);      // low positions -> high groups: keeps warp 0 free for the draw
            const bool active = pc < ncand;
            const int k = (active && pc < K) ? c.live[pc] : pslot;
            const bool own = active && (pc < K) && (k == z);
            // a singleton's own cluster without the row IS the empty cluster (exactly)
            const int ke = (own && singleton) ? pslot : k;
            const bool dn = own && !singleton;
            int drift = 0;               // (the `-= x; += x` replay itself is step 2 here)
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
                    for (int j = gl; j < D; j += G)
                        acc += (CTYPE(j) == CC_NORMAL) ? normal_term<GT>(c, T, wide, ghyp, xrow, j, k, ke, own, dn, drift)
                                                       : categorical_term<GT>(c, T, wide, ghyp, xrow, j, k, own, pc >= K);
                }
            }
            for (int o = G >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (gl == 0 && active) {
                c.lp[pc] = acc;
                // crp.py:111-124: n_k, n_k - 1 for the row's own table, alpha for a fresh / re-used singleton table
                c.w[pc] = (pc >= K) ? c.alpha : own ? (singleton ? c.alpha : (double)(nkz - 1)) : (double)c.nk[k];
                if (own) c.scal[4] = pc;
            }
        }
        consumer_sync<NT>();

        // ---- step 2: warp 0 draws; the own cluster's group re-incorporates the row -----
        if (tid < 32) {
            const double uu = c.u[buf * 32 + slot];
            auto weight = [&](int pc) -> double { return c.w[pc]; };
            int choice = -1;
            if (ncand <= 32) {
                const double lpv = (lane < ncand) ? c.lp[lane] : -INFINITY;
                const double wv = (lane < ncand) ? c.w[lane] : 0.0;
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
                for (int pc = lane; pc < ncand; pc += 32) mx = fmax(mx, c.lp[pc]);
                mx = warp_max(mx);
                double carry = 0.0;
                for (int base = 0; base < ncand; base += 32) {
                    const int pc = base + lane;
                    const double e = (pc < ncand) ? weight(pc) * exp(c.lp[pc] - mx) : 0.0;
                    carry += __shfl_sync(0xffffffffu, warp_incl_scan(e, lane), 31);
                }
                const double tot = carry;
                carry = 0.0;
                for (int base = 0; base < ncand && choice < 0; base += 32) {
                    const int pc = base + lane;
                    const double e = (pc < ncand) ? weight(pc) * exp(c.lp[pc] - mx) : 0.0;
                    const double inc = warp_incl_scan(e, lane) + carry;
                    const unsigned hit = __ballot_sync(0xffffffffu, pc < ncand && inc > uu * tot);
                    if (hit) choice = base + __ffs(hit) - 1;
                    carry = __shfl_sync(0xffffffffu, inc, 31);
                }
            }
            if (choice < 0) {                       // NaN / all -inf: keep the row, flag the chain
                choice = c.scal[4];
                if (lane == 0) c.scal[7] = CC_ERR_ARG;
            }
            if (c.trace && lane == 0) {
                double* tr = c.trace + c.trace_off + (size_t)i * c.trace_stride;
                tr[0] = (double)r;
                tr[1] = (double)ncand;
                tr[2] = (double)((choice < K) ? c.g_cid[c.live[choice]] : (c.g_cid[c.live[K - 1]] + 1));
                for (int pc = 0; pc < ncand && pc + 3 < c.trace_stride; ++pc) tr[3 + pc] = c.lp[pc] + log(weight(pc));
            }
            if (lane == 0) {
                // destination slot and "new cluster" flag, resolved here so that step 3 never reads
                // the live list / free head while the bookkeeping thread is already rewriting them
                int nc = choice >= K;
                int dst = nc ? c.scal[2] : c.live[choice];
                if (nc && dst >= Kcap - 1) { c.scal[7] = CC_ERR_CAPACITY; nc = 0; dst = z; }   // no free slot: keep the row
                c.scal[8] = choice; c.scal[9] = dst; c.scal[10] = nc;
            }
        }
        if (gid == gmask - ((NT & (NT - 1)) ? (c.scal[4] % ngroups) : (c.scal[4] & gmask))) {
            // view.py:419: incorporate the row back; the sums may differ from before in the last bit
            for (int j = gl; j < D; j += G) {
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
        }
        consumer_sync<NT>();
        }   // !skip12

        // ---- step 3: apply the move (view.py:424-429) -----------------------------------
        const int ownpos = c.scal[4];
        const int zb = c.scal[9];
        const bool newc = c.scal[10] != 0;
        if (zb != z) {
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
                        N = T.at(CC_F_N, k, j) - 1.0;
                        sx = __dsub_rn(T.at(CC_F_SX, k, j), x);
                        sxx = __dsub_rn(T.at(CC_F_SXX, k, j), __dmul_rn(x, x));
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
            consumer_sync<NT>();
        }
        if ((slot == B - 1 || i == n - 1) && tid == 0) mbar_arrive(&c.bar_empty[buf]);
    }
    return n;
}

// PIPE: two rows in flight (rows_sweep_pipe.cuh) instead of the sequential sweep
template <int NT, int MINB, bool PIPE>
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
    const long long t_start = clock64();
    Ctx c;
    c.st = st; c.s = s; c.v = v; c.n = n; c.B = B; c.bulk = p.bulk; c.sv = sv;
    // ---- fixed shared region ----------------------------------------------------
    c.bar_full = reinterpret_cast<uint64_t*>(smem_raw);               // [2]
    c.bar_empty = c.bar_full + 2;                                     // [2]
    c.scal = reinterpret_cast<int*>(c.bar_empty + 2);                 // [32]
    c.live = c.scal + 32;                                             // [Kcap]
    c.nk = c.live + Kcap;                                             // [Kcap]
    c.lp = reinterpret_cast<double*>(c.nk + Kcap);                    // [Kcap+1], twice with two rows in flight
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
    c.w = c.lp + (PIPE ? 2 : 1) * (Kcap + 1);                         // same shape
    c.hyp = c.w + (PIPE ? 2 : 1) * (Kcap + 1);                                         // [4][D] (not in wide mode)
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
        c.scal[11] = alln;
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
    c.allnorm = c.scal[11];
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
        done = PIPE ? sweep_pipe<NT, false>(c, 0, false) : sweep_seq<NT, false>(c, 0, false);
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
    if (in_global) { if (PIPE) sweep_pipe<NT, true>(c, done, !start_global); else sweep_seq<NT, true>(c, done, !start_global); }

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
            q[4] = (PIPE ? 4 : 0) + (start_global ? 1 : (in_global ? 2 : 0));      // sweep variant; table mode: shared / global / migrated mid-sweep
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

// ---------------------------------------------------------------------------------------------
// Narrow views (one or two normal columns): one warp per (chain, view), no CTA barriers at all.
// Such views carry little evidence per row, so a large share of their rows move at every sweep
// and the cost of a step is pure latency.  Here a lane owns one candidate cluster (its score and
// CRP weight never leave its registers), the draw is a handful of warp collectives, the row batch
// is prefetched into registers one batch ahead, and the whole step is one short loop that stays
// in the instruction cache.  The arithmetic, the order of the operations on the sufficient
// statistics and the stream use are those of the general kernel above (same trajectories).
// Shared memory: live[Kcap], nk[Kcap], cid[Kcap] (int); tab[8][Kcap][D]; lp[Kcap+1], w[Kcap+1]
// (only used when a row has more than 32 candidates).
__global__ void __launch_bounds__(32) rows_narrow_kernel(const RowsParams p) {
    const cc_state& st = p.st;
    const cc_row_work wk = p.work[blockIdx.x];
    const int s = wk.chain, v = wk.view, n = wk.n;
    const int R = st.R, C = st.C, Cp = st.Cp, Kcap = st.Kcap;
    const int lane = threadIdx.x;
    const size_t sv = sv_index(st, s, v);
    const long long t_start = clock64();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int* live = reinterpret_cast<int*>(smem_raw);
    int* nk = live + Kcap;
    int* cid = nk + Kcap;
    double* tab = reinterpret_cast<double*>(cid + Kcap + (Kcap & 1));
    // ---- the view's columns ----
    const int32_t* zv = st.zv + (size_t)s * C;
    int col0 = -1, col1 = -1, D = 0, alln = 1;
    for (int q = 0; q < C; ++q)
        if (zv[q] == v) {
            if (D == 0) col0 = q; else if (D == 1) col1 = q;
            if (st.ctype[q] != CC_NORMAL) alln = 0;
            ++D;
        }
    if (D < 1 || D > 2 || !alln) {          // the caller's hint was wrong: nothing is touched
        if (lane == 0) st.err[s] = CC_ERR_ARG;
        return;
    }
    if (D == 1) col1 = col0;
    const bool two = D == 2;
    double* lpS = tab + (size_t)CC_NFIELD * Kcap * D;
    double* wS = lpS + (Kcap + 1);
    const int fs = Kcap * D;
    const int pslot = Kcap - 1;
    const double* gh = st.hyp + (size_t)s * C * 4;
    const double m0 = gh[col0 * 4 + 0], r0 = gh[col0 * 4 + 1], s0 = gh[col0 * 4 + 2], nu0 = gh[col0 * 4 + 3];
    const double m1 = gh[col1 * 4 + 0], r1 = gh[col1 * 4 + 1], s1 = gh[col1 * 4 + 2], nu1 = gh[col1 * 4 + 3];

    int K = st.nclu[sv];
    int32_t* g_live = st.live + sv * Kcap;
    int32_t* g_nk = st.nk + sv * Kcap;
    int32_t* g_cid = st.cid + sv * Kcap;
    for (int i = lane; i < Kcap; i += 32) { nk[i] = 0; live[i] = (i < K) ? g_live[i] : -1; cid[i] = g_cid[i]; }
    __syncwarp();
    for (int i = lane; i < K; i += 32) nk[live[i]] = g_nk[live[i]];
    for (int e = lane; e < (K + 1) * D; e += 32) {
        const int pp = e / D, d = e - pp * D;
        const int k = (pp < K) ? live[pp] : pslot;
        const int q = d ? col1 : col0;
#pragma unroll
        for (int f = 0; f < CC_NFIELD; ++f) tab[f * fs + k * D + d] = stat_field(st, s, f)[(size_t)k * C + q];
    }
    __syncwarp();
    int fh = 0;                               // lowest free slot
    while (fh < Kcap - 1 && nk[fh] != 0) ++fh;
    int nmoved = 0, err = 0;
    const double alpha = st.view_alpha[sv];
    int32_t* zr = st.zr + sv * R;
    int32_t* moved = st.moved + sv * R;
    uint8_t* movedflag = st.movedflag + sv * R;
    double* trace = (p.trace && wk.trace_off >= 0) ? p.trace : nullptr;
    const int32_t* order = st.order + sv * R;
    const Philox rng(p.seed);
    uint32_t bits = 1;
    while ((1u << bits) < (uint32_t)n && bits < 31) ++bits;
    const uint4 pkey = rng(p.counter, stream_id(2, s, v));
    const uint64_t ustream = stream_id(1, s, v);

    // lane l of a batch holds row (batch * 32 + l): id, current cluster, uniform, cells
    int nr = 0, nz = 0;
    double nu_ = 0.0, nx0 = 0.0, nx1 = 0.0;
    auto fetch = [&](int i) {
        if (i < n) {
            int idx;
            if (wk.perm_off >= 0) idx = p.perm[wk.perm_off + i];
            else idx = (int)keyed_perm((uint32_t)i, (uint32_t)n, bits, pkey);
            nr = wk.perm_is_index ? order[idx] : idx;
            if (wk.u_off >= 0) nu_ = p.u[wk.u_off + i];
            else { const uint4 w4 = rng(p.counter * 0x100000000ull + (uint64_t)i, ustream); nu_ = u01_from_words(w4.x, w4.y); }
            nz = zr[nr];
            const double* src = st.Xrm + (size_t)nr * Cp;
            nx0 = src[col0];
            nx1 = src[col1];
        }
    };
    fetch(lane);
    for (int b0 = 0; b0 < n; b0 += 32) {
        const int br = nr, bz = nz;
        const double bu = nu_, bx0 = nx0, bx1 = nx1;
        fetch(b0 + 32 + lane);                                     // in flight while this batch is processed
        const int cnt = min(32, n - b0);
        for (int t = 0; t < cnt; ++t) {
            const int i = b0 + t;
            const int r = __shfl_sync(0xffffffffu, br, t);
            const int z = __shfl_sync(0xffffffffu, bz, t);
            const double uu = __shfl_sync(0xffffffffu, bu, t);
            const double x0 = __shfl_sync(0xffffffffu, bx0, t);
            const double x1 = __shfl_sync(0xffffffffu, bx1, t);
            const int nkz = nk[z];
            const bool singleton = nkz == 1;
            const int ncand = K + (singleton ? 0 : 1);
            // ---- scores: lane = candidate position (32 at a time) ----
            double lpv = -INFINITY, wv = 0.0;
            int ownpos = -1, drift = 0;
            for (int base = 0; base < ncand; base += 32) {
                const int pc = base + lane;
                const bool active = pc < ncand;
                const int k = (pc < K) ? live[pc] : pslot;
                const bool own = (pc < K) && (k == z);
                const int ke = (own && singleton) ? pslot : k;
                const bool dn = own && !singleton;
                double acc = 0.0;
                int dr = 0;
                if (active) {
                    acc = normal_core(tab + ke * D, tab + k * D, fs, x0, nu0, r0, own, dn, dr,
                                      [&](double& m, double& ss) { m = m0; ss = s0; });
                    if (two) acc += normal_core(tab + ke * D + 1, tab + k * D + 1, fs, x1, nu1, r1, own, dn, dr,
                                                [&](double& m, double& ss) { m = m1; ss = s1; });
                }
                // crp.py:111-124: n_k, n_k - 1 for the row's own table, alpha for a fresh / re-used singleton table
                const double wc = (pc >= K) ? alpha : own ? (singleton ? alpha : (double)(nkz - 1)) : (double)nk[k];
                const unsigned ob = __ballot_sync(0xffffffffu, own);
                if (ob) { ownpos = base + __ffs(ob) - 1; drift = __shfl_sync(0xffffffffu, dr, __ffs(ob) - 1); }
                if (ncand <= 32) { if (active) { lpv = acc; wv = wc; } }
                else if (active) { lpS[pc] = acc; wS[pc] = wc; }
            }
            // ---- the draw: first j with prefix_j > u * total over w_j exp(lp_j - max) ----
            int choice = -1;
            if (ncand <= 32) {
                float mf = (float)lpv;
                asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(mf) : "f"(mf));
                const double dl = lpv - (double)mf;
                double inc = (dl > -80.0) ? wv * exp(fmax(dl, -80.0)) : ((dl == dl) ? 0.0 : dl);
                double tt;
                tt = __shfl_up_sync(0xffffffffu, inc, 1); if (lane >= 1) inc += tt;
                tt = __shfl_up_sync(0xffffffffu, inc, 2); if (lane >= 2) inc += tt;
                if (ncand > 4) { tt = __shfl_up_sync(0xffffffffu, inc, 4); if (lane >= 4) inc += tt; }
                if (ncand > 8) { tt = __shfl_up_sync(0xffffffffu, inc, 8); if (lane >= 8) inc += tt; }
                if (ncand > 16) { tt = __shfl_up_sync(0xffffffffu, inc, 16); if (lane >= 16) inc += tt; }
                const double tot = __shfl_sync(0xffffffffu, inc, ncand - 1);
                const unsigned hit = __ballot_sync(0xffffffffu, lane < ncand && inc > uu * tot);
                choice = hit ? (__ffs(hit) - 1) : -1;
            } else {
                __syncwarp();
                double mx = -INFINITY;
                for (int pc = lane; pc < ncand; pc += 32) mx = fmax(mx, lpS[pc]);
                mx = warp_max(mx);
                double carry = 0.0;
                for (int base = 0; base < ncand; base += 32) {
                    const int pc = base + lane;
                    const double e = (pc < ncand) ? wS[pc] * exp(lpS[pc] - mx) : 0.0;
                    carry += __shfl_sync(0xffffffffu, warp_incl_scan(e, lane), 31);
                }
                const double tot = carry;
                carry = 0.0;
                for (int base = 0; base < ncand && choice < 0; base += 32) {
                    const int pc = base + lane;
                    const double e = (pc < ncand) ? wS[pc] * exp(lpS[pc] - mx) : 0.0;
                    const double inc = warp_incl_scan(e, lane) + carry;
                    const unsigned hit = __ballot_sync(0xffffffffu, pc < ncand && inc > uu * tot);
                    if (hit) choice = base + __ffs(hit) - 1;
                    carry = __shfl_sync(0xffffffffu, inc, 31);
                }
            }
            if (choice < 0) { choice = ownpos; err = CC_ERR_ARG; }      // NaN / all -inf: keep the row, flag the chain
            if (trace) {
                double* tr = trace + wk.trace_off + (size_t)i * p.trace_stride;
                if (lane == 0) {
                    tr[0] = (double)r;
                    tr[1] = (double)ncand;
                    tr[2] = (double)((choice < K) ? cid[live[choice]] : (cid[live[K - 1]] + 1));
                }
                if (ncand <= 32) { if (lane < ncand && lane + 3 < p.trace_stride) tr[3 + lane] = lpv + log(wv); }
                else for (int pc = lane; pc < ncand && pc + 3 < p.trace_stride; pc += 32) tr[3 + pc] = lpS[pc] + log(wS[pc]);
            }
            bool newc = choice >= K;
            int zb = newc ? fh : live[choice];
            if (newc && zb >= Kcap - 1) { err = CC_ERR_CAPACITY; newc = false; zb = z; }     // no free slot: keep the row
            if (zb != z) {
                // ---- apply the move (view.py:424-429): lane = (column, old / new cluster) ----
                const int d = lane >> 1;
                const bool is_old = (lane & 1) == 0;
                const bool fresh = !is_old && newc;
                const double x = d ? x1 : x0;
                const bool on = lane < 2 * D && !(is_old && singleton) && (!isnan(x) || fresh);
                if (on) {
                    const int k = is_old ? z : zb;
                    double* pc_ = tab + k * D + d;
                    const double m = d ? m1 : m0, rr = d ? r1 : r0, ss = d ? s1 : s0, nu = d ? nu1 : nu0;
                    const double gP = tab[CC_F_GN * fs + pslot * D + d];
                    const double N0 = fresh ? 0.0 : pc_[CC_F_N * fs];
                    const double sx0 = fresh ? 0.0 : pc_[CC_F_SX * fs], sxx0 = fresh ? 0.0 : pc_[CC_F_SXX * fs];
                    const double gn0 = fresh ? gP : pc_[CC_F_GN * fs], gm0 = pc_[CC_F_GM1 * fs];
                    const double xx = __dmul_rn(x, x);
                    double N, sx, sxx, gN, gM1;
                    const bool nanx = isnan(x);
                    // G of the changed count: G(N-2) for the cluster that loses the row, G(N+1) for the one that gains it
                    const double garg = is_old ? N0 - 2.0 : N0 + 1.0;
                    const double g = normal_G(fmax(garg, 0.0), rr, nu);
                    if (is_old) {
                        // scored against its own cluster first (-= x, += x: view.py:416-419), then removed
                        N = N0 - 1.0;
                        sx = __dsub_rn(__dadd_rn(__dsub_rn(sx0, x), x), x);
                        sxx = __dsub_rn(__dadd_rn(__dsub_rn(sxx0, xx), xx), xx);
                        gN = gm0;
                        gM1 = (N >= 1.0) ? g : 0.0;
                    } else if (nanx) {          // fresh cluster object, missing cell (normal.py:51-53)
                        N = 0.0; sx = 0.0; sxx = 0.0; gN = gP; gM1 = 0.0;
                    } else {                    // 0 + x, 0 + x*x for a fresh cluster object (normal.py:64-70)
                        N = N0 + 1.0;
                        sx = __dadd_rn(sx0, x);
                        sxx = __dadd_rn(sxx0, xx);
                        gM1 = gn0;
                        gN = g;
                    }
                    const NormalCache cc = normal_cache_fast(N, sx, sxx, gN, m, rr, ss, nu);
                    pc_[CC_F_N * fs] = N; pc_[CC_F_SX * fs] = sx; pc_[CC_F_SXX * fs] = sxx;
                    pc_[CC_F_MN * fs] = cc.mn; pc_[CC_F_A * fs] = cc.a; pc_[CC_F_CST * fs] = cc.cst;
                    pc_[CC_F_GN * fs] = gN; pc_[CC_F_GM1 * fs] = gM1;
                }
                // ---- bookkeeping of the row CRP (crp.py:45-59) ----
                if (lane == 0) {
                    zr[r] = zb;
                    moved[nmoved] = r;
                    movedflag[r] = 1;
                    nk[z] = nkz - 1;
                    if (newc) {
                        cid[zb] = cid[live[K - 1]] + 1;              // crp.py:142 max(counts)+1
                        live[K] = zb;
                        nk[zb] = 1;
                    } else {
                        nk[zb] += 1;
                    }
                }
                ++nmoved;
                __syncwarp();
                if (newc) {
                    ++K;
                    ++fh;
                    while (fh < Kcap - 1 && nk[fh] != 0) ++fh;
                }
                if (singleton) {
                    // the emptied cluster leaves the ascending-id list; its slot is free again
                    for (int q0 = ownpos; q0 < K - 1; q0 += 32) {
                        const int q = q0 + lane;
                        const int nxt = (q < K - 1) ? live[q + 1] : -1;
                        __syncwarp();
                        if (q < K - 1) live[q] = nxt;
                        __syncwarp();
                    }
                    if (lane == 0) live[K - 1] = -1;
                    --K;
                    if (z < fh) fh = z;
                }
                __syncwarp();
            } else if (drift) {
                // the row stays, but `-= x; += x` changed its cluster's sums in the last bit (view.py:416-419)
                const double x = lane ? x1 : x0;
                if (lane < D && !isnan(x)) {
                    double* pc_ = tab + z * D + lane;
                    const double sx0 = pc_[CC_F_SX * fs], sxx0 = pc_[CC_F_SXX * fs];
                    const double xx = __dmul_rn(x, x);
                    const double sx2 = __dadd_rn(__dsub_rn(sx0, x), x), sxx2 = __dadd_rn(__dsub_rn(sxx0, xx), xx);
                    if (sx2 != sx0 || sxx2 != sxx0) {
                        pc_[CC_F_SX * fs] = sx2;
                        pc_[CC_F_SXX * fs] = sxx2;
                        const NormalCache c2 = normal_cache_fast(pc_[CC_F_N * fs], sx2, sxx2, pc_[CC_F_GN * fs], lane ? m1 : m0,
                                                                 lane ? r1 : r0, lane ? s1 : s0, lane ? nu1 : nu0);
                        pc_[CC_F_MN * fs] = c2.mn; pc_[CC_F_A * fs] = c2.a; pc_[CC_F_CST * fs] = c2.cst;
                    }
                }
                __syncwarp();
            }
        }
    }

    // ---- write back ----
    __syncwarp();
    for (int e = lane; e < K * D; e += 32) {
        const int pp = e / D, d = e - pp * D;
        const int k = live[pp], q = d ? col1 : col0;
#pragma unroll
        for (int f = 0; f < CC_NFIELD; ++f) stat_field(st, s, f)[(size_t)k * C + q] = tab[f * fs + k * D + d];
    }
    for (int i = lane; i < Kcap; i += 32) {
        g_live[i] = (i < K) ? live[i] : -1;
        g_nk[i] = nk[i];
        g_cid[i] = cid[i];
    }
    if (lane == 0) {
        st.nclu[sv] = K;
        if (err) st.err[s] = err;
        if (p.prof) {
            int32_t* q = st.seq + sv * R;
            q[0] = (int)((clock64() - t_start) >> 10); q[1] = D; q[2] = K; q[3] = nmoved; q[4] = 3;    // mode 3: single-warp kernel
        }
    }
    // ---- member order of the view's CRP: re-seated rows move to the end, in the order they
    //      were re-seated (OrderedDict pop + re-insert, crp.py:52,55) ----
    if (nmoved > 0) {
        __threadfence_block();
        int32_t* ord = st.order + sv * R;
        int base = 0;
        for (int i0 = 0; i0 < R; i0 += 128) {
            int rr[4], keep[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { const int i = i0 + q * 32 + lane; rr[q] = (i < R) ? ord[i] : -1; }
#pragma unroll
            for (int q = 0; q < 4; ++q) keep[q] = (rr[q] >= 0) ? !movedflag[rr[q]] : 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const unsigned bal = __ballot_sync(0xffffffffu, keep[q]);
                if (keep[q]) ord[base + __popc(bal & ((1u << lane) - 1u))] = rr[q];
                base += __popc(bal);
            }
        }
        __syncwarp();
        for (int i = lane; i < nmoved; i += 32) {
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
    // Work items are split three ways; the launches run concurrently (the extra ones on internal
    // streams ordered after / before `stream` by events):
    //   * views of one or two normal columns (the caller's n_cols / n_normal hints) -> one warp each
    //     (rows_narrow_kernel)
    //   * views wider than 12 columns -> 7 consumer warps (each candidate cluster needs a full warp or more;
    //     with the producer warp that is 256 threads, i.e. 128 registers each at two CTAs per SM)
    //   * the others -> 4 consumer warps
    const int Kcap_ = st->Kcap;
    auto narrow_bytes = [&](int d) { return 12 * Kcap_ + 16 + 64 * Kcap_ * d + 16 * (Kcap_ + 1); };
    const bool env_forced = getenv("CGPM_B200_ROWS_NT") != nullptr;
    const bool narrow_ok = !env_forced && !getenv("CGPM_B200_ROWS_NO_NARROW") && !getenv("CGPM_B200_ROWS_SMEM_KB");
    int wide_min = 12;          // views wider than this get the 7-warp, two-rows-in-flight kernel (measured: 16 -> 12 is 2 % faster, 8 much slower)
    if (const char* e = getenv("CGPM_B200_ROWS_WIDE_MIN")) { const int v = atoi(e); if (v >= 2) wide_min = v; }
    auto klass = [&](const cc_row_work& w) -> int {
        if (narrow_ok && w.n_cols >= 1 && w.n_cols <= 2 && w.n_normal == w.n_cols && narrow_bytes(w.n_cols) <= 64 * 1024) return 0;
        if (!env_forced && w.n_cols > wide_min) return 1;
        return 2;
    };
    std::vector<cc_row_work> sorted(work, work + n_work);
    std::stable_sort(sorted.begin(), sorted.end(), [&](const cc_row_work& a, const cc_row_work& b) { return klass(a) < klass(b); });
    int n_narrow = 0, n_wide = 0, dmax_narrow = 0;
    for (int i = 0; i < n_work; ++i) {
        const int k = klass(sorted[i]);
        n_narrow += k == 0; n_wide += k == 1;
        if (k == 0 && sorted[i].n_cols > dmax_narrow) dmax_narrow = sorted[i].n_cols;
    }
    const int n_med = n_work - n_narrow - n_wide;
    cc_row_work* dwork = nullptr;
    CC_CUDA(cudaMallocAsync(&dwork, sizeof(cc_row_work) * n_work, stream));
    CC_CUDA(cudaMemcpyAsync(dwork, sorted.data(), sizeof(cc_row_work) * n_work, cudaMemcpyHostToDevice, stream));
    CC_CUDA(cudaStreamSynchronize(stream));            // `sorted` is a local; n_work is small

    int dev = 0, nsm = 148, max_smem = 0;
    CC_CUDA(cudaGetDevice(&dev));
    CC_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    CC_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (dev < 0 || dev >= 64) { cc_set_error("cc_transition_rows: device ordinal %d not supported", dev); return CC_ERR_ARG; }
    static cudaStream_t side[64][2] = {{nullptr, nullptr}};
    static cudaEvent_t ev_fork[64] = {nullptr}, ev_join[64][2] = {{nullptr, nullptr}};
    if (!side[dev][0]) {
        for (int q = 0; q < 2; ++q) {
            CC_CUDA(cudaStreamCreateWithFlags(&side[dev][q], cudaStreamNonBlocking));
            CC_CUDA(cudaEventCreateWithFlags(&ev_join[dev][q], cudaEventDisableTiming));
        }
        CC_CUDA(cudaEventCreateWithFlags(&ev_fork[dev], cudaEventDisableTiming));
    }
    // shared memory: the single-warp CTAs take what they need; the rest of each SM is divided among
    // the multi-warp CTAs expected on it
    const int narrow_need = n_narrow ? narrow_bytes(dmax_narrow) : 0;
    const int reserve = ((n_narrow + nsm - 1) / nsm) * (narrow_need + 1024);
    const int n_general = n_wide + n_med;
    int env_nt = 0;
    if (const char* e = getenv("CGPM_B200_ROWS_NT")) { const int v = atoi(e); env_nt = (v == 256) ? v : 128; }

    auto launch = [&](const cc_row_work* hw, const cc_row_work* dw, int cnt, int nt, int total_ctas, cudaStream_t strm) -> int {
        RowsParams p;
        p.st = *st; p.work = dw; p.perm = perm; p.u = u; p.trace = trace; p.trace_stride = trace_stride;
        p.seed = seed; p.counter = counter;
        p.prof = getenv("CGPM_B200_ROWS_PROF") ? 1 : 0;
        // widest view among the work items (callers that do not know pass 0 -> assume all columns)
        int dmax = 0;
        for (int i = 0; i < cnt; ++i) {
            const int d = hw[i].n_cols;
            if (d <= 0 || d > st->C) { dmax = st->C; break; }
            if (d > dmax) dmax = d;
        }
        // rows are staged whole (bulk copy) when they are short; otherwise only the view's cells
        // are gathered.  Two batches of B rows (a power of two) must stay within 32 KB.
        const int rowbytes = st->Cp * 8;
        p.bulk = (rowbytes <= 4096) ? 1 : 0;
        const int stagebytes = p.bulk ? rowbytes : ((dmax + 1) & ~1) * 8;
        p.wide = (dmax > 1024) ? 1 : 0;
        p.B = 32;
        while (p.B > 1 && 2 * p.B * stagebytes > 32 * 1024) p.B >>= 1;
        // CTAs per SM wanted: enough to co-schedule all work items of both launches
        const int maxb = (nt == 256) ? 2 : 3;
        int per_sm = (total_ctas + nsm - 1) / nsm;
        if (per_sm < 1) per_sm = 1;
        if (per_sm > maxb) per_sm = maxb;
        int avail = max_smem + 1024 - reserve;
        if (avail < (max_smem + 1024) / 2) avail = (max_smem + 1024) / 2;
        int budget = avail / per_sm - 1024 - 256;      // 1 KB per-CTA system reservation
        const int xbytes = 2 * p.B * stagebytes;
        const bool pipe = nt == 256 && !getenv("CGPM_B200_ROWS_NO_PIPE") && !env_forced;       // wide views: two rows in flight
        const int fixed = 64 + 128 + st->Kcap * 8 + (st->Kcap + 2) * (pipe ? 32 : 16) + dmax * (p.wide ? 4 : 48) + 512 + 64 + 512 + xbytes;
        bool forced = false;
        if (const char* e = getenv("CGPM_B200_ROWS_SMEM_KB")) { const int kb = atoi(e); if (kb > 0) { budget = fixed + kb * 1024; forced = true; } }   // tests: force small tables
        if (!forced && budget < fixed + 4096) budget = fixed + 4096;
        if (budget > max_smem) budget = max_smem;
        if (fixed > max_smem) { cc_set_error("cc_transition_rows: shared memory need %d exceeds %d", fixed, max_smem); return CC_ERR_ARG; }
        p.smem_bytes = budget;
        if (pipe) {
            CC_CUDA(cudaFuncSetAttribute(rows_kernel<224, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, budget));
            rows_kernel<224, 2, true><<<cnt, 256, budget, strm>>>(p);
        } else if (nt == 256) {
            CC_CUDA(cudaFuncSetAttribute(rows_kernel<224, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, budget));
            rows_kernel<224, 2, false><<<cnt, 256, budget, strm>>>(p);
        } else {
            CC_CUDA(cudaFuncSetAttribute(rows_kernel<128, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, budget));
            rows_kernel<128, 3, false><<<cnt, 160, budget, strm>>>(p);
        }
        CC_CUDA(cudaGetLastError());
        return CC_OK;
    };
    int rc = CC_OK;
    const int n_launch = (n_narrow > 0) + (n_wide > 0) + (n_med > 0);
    if (n_launch > 1) CC_CUDA(cudaEventRecord(ev_fork[dev], stream));
    int used_side = 0, launched = 0;
    auto pick_stream = [&]() -> cudaStream_t {
        if (++launched == n_launch) return stream;          // the last launch goes on the caller's stream
        cudaStream_t sd = side[dev][used_side++];
        cudaStreamWaitEvent(sd, ev_fork[dev], 0);
        return sd;
    };
    if (n_wide > 0) rc = launch(sorted.data() + n_narrow, dwork + n_narrow, n_wide, 256, n_general, pick_stream());
    if (n_narrow > 0 && rc == CC_OK) {
        RowsParams p;
        p.st = *st; p.work = dwork; p.perm = perm; p.u = u; p.trace = trace; p.trace_stride = trace_stride;
        p.seed = seed; p.counter = counter;
        p.prof = getenv("CGPM_B200_ROWS_PROF") ? 1 : 0;
        p.B = 32; p.bulk = 0; p.smem_bytes = narrow_need; p.wide = 0;
        CC_CUDA(cudaFuncSetAttribute(rows_narrow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, narrow_need));
        rows_narrow_kernel<<<n_narrow, 32, narrow_need, pick_stream()>>>(p);
        CC_CUDA(cudaGetLastError());
    }
    if (n_med > 0 && rc == CC_OK)
        rc = launch(sorted.data() + n_narrow + n_wide, dwork + n_narrow + n_wide, n_med, env_nt ? env_nt : 128, n_general, pick_stream());
    for (int q = 0; q < used_side; ++q) {
        CC_CUDA(cudaEventRecord(ev_join[dev][q], side[dev][q]));
        CC_CUDA(cudaStreamWaitEvent(stream, ev_join[dev][q], 0));
    }
    if (rc != CC_OK) return rc;
    CC_CUDA(cudaFreeAsync(dwork, stream));
    return CC_OK;
}
