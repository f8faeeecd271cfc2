// Row-reassignment Gibbs sweep for WIDE views (more than 16 columns), included by rows.cu: the same per-row work as
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
#pragma once
#include "rows_common.cuh"

namespace {

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
__device__ __forceinline__ int sweep_pipe(const Ctx& c, int i_begin, bool resume) {
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

}  // namespace
