// A host with no Python: loads a table and an initial CrossCat state into plain cudaMalloc'ed arrays and runs
// inference and queries through the C ABI only (include/cgpm_b200.h) -- the call sequence a C / C++ embedder of
// the reference's State.transition_lovecat precedent would use (src/crosscat/state.py:812-831,
// src/crosscat/lovecat.py:377-391):  cc_rebuild -> cc_transition(kernel list, n steps) -> cc_check_state ->
// cc_logpdf_score -> cc_logpdf_bulk.   Built and run by tests/test_gpu_c_driver.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "cgpm_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s\n", cudaGetErrorString(e_), #x); return 2; } } while (0)
#define CC(x) do { int rc_ = (x); if (rc_ != CC_OK) { printf("%s failed (%d): %s\n", #x, rc_, cc_last_error()); return 3; } } while (0)

template <typename T> static T* dalloc(size_t n, int fill_byte = 0) {
    T* p = nullptr;
    if (cudaMalloc(&p, sizeof(T) * (n ? n : 1)) != cudaSuccess) { printf("cudaMalloc failed\n"); exit(2); }
    cudaMemset(p, fill_byte, sizeof(T) * (n ? n : 1));
    return p;
}
template <typename T> static void up(T* d, const std::vector<T>& h) { cudaMemcpy(d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice); }
template <typename T> static std::vector<T> down(const T* d, size_t n) { std::vector<T> h(n); cudaMemcpy(h.data(), d, sizeof(T) * n, cudaMemcpyDeviceToHost); return h; }

static uint64_t rs = 88172645463325252ull;
static double unif() { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (double)(rs >> 11) / 9007199254740992.0; }
static double gauss() { return sqrt(-2.0 * log(unif() + 1e-300)) * cos(6.283185307179586 * unif()); }
static void logspace(double a, double b, double* out) { for (int g = 0; g < CC_NGRID; ++g) out[g] = exp(log(a) + (log(b) - log(a)) * g / (CC_NGRID - 1)); }

int main() {
    const int R = 3000, C = 6, S = 3, Vcap = C, Kcap = R + 1, Cp = (C + 1) & ~1;
    // two independent views of three columns, three well separated clusters each
    std::vector<double> X((size_t)R * Cp, NAN), Xc((size_t)C * R);
    for (int r = 0; r < R; ++r) {
        const int za = (int)(unif() * 3), zb = (int)(unif() * 3);
        for (int c = 0; c < C; ++c) { const double x = 6.0 * (c < 3 ? za : zb) + gauss(); X[(size_t)r * Cp + c] = x; Xc[(size_t)c * R + r] = x; }
    }
    cc_state st = {};
    st.R = R; st.C = C; st.Cp = Cp; st.S = S; st.Vcap = Vcap; st.Kcap = Kcap; st.CT = 1;
    double* dX = dalloc<double>(X.size()); up(dX, X); st.Xrm = dX;
    double* dXc = dalloc<double>(Xc.size()); up(dXc, Xc); st.Xcm = dXc;
    st.ctype = dalloc<int32_t>(C); st.ncat = dalloc<int32_t>(C); st.catoff = dalloc<int32_t>(C);     // all CC_NORMAL (0)
    // hyper grids (normal.py:128-139, crp.py:148-152)
    std::vector<double> grids((size_t)C * 4 * CC_NGRID), rowg(CC_NGRID), colg(CC_NGRID);
    for (int c = 0; c < C; ++c) {
        double mn = 1e300, mx = -1e300, mean = 0, ssq = 0;
        for (int r = 0; r < R; ++r) { const double x = Xc[(size_t)c * R + r]; mn = fmin(mn, x); mx = fmax(mx, x); mean += x / R; }
        for (int r = 0; r < R; ++r) { const double d = Xc[(size_t)c * R + r] - mean; ssq += d * d; }
        double* g = &grids[(size_t)c * 4 * CC_NGRID];
        for (int q = 0; q < CC_NGRID; ++q) g[q] = mn + (mx + 5 - mn) * q / (CC_NGRID - 1);
        logspace(1.0 / (R + 1), R + 1, g + CC_NGRID); logspace((ssq + 1) / 100, ssq + 1, g + 2 * CC_NGRID); logspace(1.0, R + 1, g + 3 * CC_NGRID);
    }
    logspace(1.0 / R, R, rowg.data()); logspace(1.0 / C, C, colg.data());
    double* dg = dalloc<double>(grids.size()); up(dg, grids); st.grids = dg;
    double* drg = dalloc<double>(CC_NGRID); up(drg, rowg); st.row_alpha_grid = drg;
    double* dcg = dalloc<double>(CC_NGRID); up(dcg, colg); st.col_alpha_grid = dcg;
    // initial state of every chain: two arbitrary views (even / odd columns) of eight arbitrary clusters (row mod 8) --
    // what State(X, Zv={c: c % 2}, Zrv={v: [r % 8 ...]}) would load in the reference
    const int K0 = 8;
    std::vector<double> hyp((size_t)S * C * 4), ones_s(S, 1.0), ones_sv((size_t)S * Vcap, 1.0), vgrid((size_t)S * Vcap * CC_NGRID);
    for (int s = 0; s < S; ++s) for (int c = 0; c < C; ++c) { double* h = &hyp[((size_t)s * C + c) * 4]; h[0] = grids[(size_t)c * 4 * CC_NGRID + 10]; h[1] = 1; h[2] = grids[((size_t)c * 4 + 2) * CC_NGRID + 5]; h[3] = 2; }
    for (size_t i = 0; i < (size_t)S * Vcap; ++i) for (int g = 0; g < CC_NGRID; ++g) vgrid[i * CC_NGRID + g] = rowg[g];
    st.hyp = dalloc<double>(hyp.size()); up(st.hyp, hyp);
    std::vector<int32_t> zv0((size_t)S * C), zr0((size_t)S * Vcap * R, 0), cid0((size_t)S * Vcap * Kcap, 0);
    std::vector<int32_t> nv((size_t)S * Vcap, 0), vid((size_t)S * Vcap, -1), nclu((size_t)S * Vcap, 0), live((size_t)S * Vcap * Kcap, -1), nk((size_t)S * Vcap * Kcap, 0), order((size_t)S * Vcap * R, 0);
    for (int s = 0; s < S; ++s) {
        for (int c = 0; c < C; ++c) zv0[(size_t)s * C + c] = c % 2;
        for (int v = 0; v < 2; ++v) {
            const size_t sv = (size_t)s * Vcap + v;
            nv[sv] = C / 2; vid[sv] = v; nclu[sv] = K0;
            for (int k = 0; k < K0; ++k) { live[sv * Kcap + k] = k; cid0[sv * Kcap + k] = k; }
            for (int r = 0; r < R; ++r) { const int k = (r + s + v) % K0; zr0[sv * R + r] = k; nk[sv * Kcap + k] += 1; order[sv * R + r] = r; }
        }
    }
    st.zv = dalloc<int32_t>(zv0.size()); up(st.zv, zv0);
    st.nv = dalloc<int32_t>(nv.size()); up(st.nv, nv);
    st.view_id = dalloc<int32_t>(vid.size()); up(st.view_id, vid);
    st.col_alpha = dalloc<double>(S); up(st.col_alpha, ones_s);
    st.view_alpha = dalloc<double>(ones_sv.size()); up(st.view_alpha, ones_sv);
    st.view_grid = dalloc<double>(vgrid.size()); up(st.view_grid, vgrid);
    st.nclu = dalloc<int32_t>(nclu.size()); up(st.nclu, nclu);
    st.live = dalloc<int32_t>(live.size()); up(st.live, live);
    st.cid = dalloc<int32_t>(cid0.size()); up(st.cid, cid0);
    st.nk = dalloc<int32_t>(nk.size()); up(st.nk, nk);
    st.zr = dalloc<int32_t>(zr0.size()); up(st.zr, zr0);
    st.order = dalloc<int32_t>(order.size()); up(st.order, order);
    st.stat = dalloc<double>((size_t)S * CC_NFIELD * Kcap * C);
    st.cnt = dalloc<double>((size_t)S * Kcap * st.CT);
    st.err = dalloc<int32_t>(S);
    st.rows_prog = dalloc<int32_t>((size_t)S * Vcap * 4);
    st.seq = dalloc<int32_t>((size_t)S * Vcap * R); st.moved = dalloc<int32_t>((size_t)S * Vcap * R);
    st.movedflag = dalloc<uint8_t>((size_t)S * Vcap * R);
    st.colL = dalloc<double>((size_t)S * C * (Vcap + 1));
    st.aux_zr = dalloc<int32_t>((size_t)S * R); st.aux_K = dalloc<int32_t>(S);
    st.aux_alpha = dalloc<double>((size_t)S * C);
    st.aux_zr_all = dalloc<int32_t>((size_t)S * C * R); st.aux_K_all = dalloc<int32_t>((size_t)S * C);
    const size_t n1 = (R + 31) >> 5, n2 = (n1 + 31) >> 5;
    size_t per = (size_t)R + (R + n1 + n2 + 32) + (R + 1) + R + R;
    if ((((size_t)R + 7) & ~(size_t)7) + 8 * (size_t)R > per) per = (((size_t)R + 7) & ~(size_t)7) + 8 * (size_t)R;
    per = (per + 7) & ~(size_t)7;
    st.arena_bytes = (int64_t)(per * 4 * S * C);
    st.arena = dalloc<int32_t>(per * S * C);

    printf("abi %d\n", cc_abi_version());
    CC(cc_rebuild(&st, 0, nullptr, nullptr, nullptr, nullptr));
    std::vector<double> sc0(S), sc1(S);
    double* dsc = dalloc<double>(S);
    CC(cc_logpdf_score(&st, dsc, nullptr));
    CK(cudaDeviceSynchronize());
    sc0 = down(dsc, S);
    const int32_t chains[3] = {0, 1, 2};
    const int32_t kernels[5] = {CC_K_ROWS, CC_K_COLUMNS, CC_K_VIEW_ALPHAS, CC_K_ALPHA, CC_K_COLUMN_HYPERS};
    uint64_t counter = 1;
    CC(cc_transition(&st, S, chains, 60, 5, kernels, nullptr, nullptr, 20260101ull, &counter, nullptr, nullptr, nullptr));
    CK(cudaDeviceSynchronize());
    int32_t* dchk = dalloc<int32_t>(S);
    CC(cc_check_state(&st, dchk, nullptr));
    CC(cc_logpdf_score(&st, dsc, nullptr));
    CK(cudaDeviceSynchronize());
    sc1 = down(dsc, S);
    const std::vector<int32_t> chk = down(dchk, S), err = down(st.err, S), nvh = down(st.nv, (size_t)S * Vcap), zvh = down(st.zv, (size_t)S * C),
                               ncl = down(st.nclu, (size_t)S * Vcap);
    int ok = 1, recovered = 0;
    std::vector<int> good(S, 0);
    for (int s = 0; s < S; ++s) {
        int views = 0, kmin = 1 << 30;
        for (int v = 0; v < Vcap; ++v) if (nvh[(size_t)s * Vcap + v] > 0) { ++views; if (ncl[(size_t)s * Vcap + v] < kmin) kmin = ncl[(size_t)s * Vcap + v]; }
        const bool split = zvh[(size_t)s * C] == zvh[(size_t)s * C + 1] && zvh[(size_t)s * C + 1] == zvh[(size_t)s * C + 2] &&
                           zvh[(size_t)s * C + 3] == zvh[(size_t)s * C + 4] && zvh[(size_t)s * C + 4] == zvh[(size_t)s * C + 5] &&
                           zvh[(size_t)s * C] != zvh[(size_t)s * C + 3];
        printf("chain %d: invariants %d err %d views %d min clusters %d score %.3f -> %.3f, columns {0,1,2} | {3,4,5} recovered: %d\n",
               s, chk[s], err[s], views, kmin, sc0[s], sc1[s], (int)split);
        ok &= chk[s] == 0 && err[s] == 0 && sc1[s] > sc0[s];           // consistent partitions, a far better score than the arbitrary start
        good[s] = split && kmin >= 2;
        recovered += good[s];
    }
    ok &= recovered >= 1;       // the generating structure (two views of three clusters) found from the arbitrary start
                                // (Gibbs chains that merged everything into one cluster split again only slowly, as in the reference)
    // a query batch: p(x0 | x1) for three hypothetical rows, every chain
    const int Q = 3;
    std::vector<int32_t> rowid(Q, -1), qoff = {0, 2, 4, 6}, qcol = {0, 1, 0, 1, 0, 1};
    std::vector<double> qval = {0.0, 0.0, 6.0, 6.0, 0.0, 12.0};
    std::vector<uint8_t> qrole = {0, 1, 0, 1, 0, 1};
    int32_t *drow = dalloc<int32_t>(Q), *doff = dalloc<int32_t>(Q + 1), *dcol = dalloc<int32_t>(6), *dflag = dalloc<int32_t>((size_t)S * Q);
    double *dval = dalloc<double>(6), *dout = dalloc<double>((size_t)S * Q);
    uint8_t* drole = dalloc<uint8_t>(6);
    up(drow, rowid); up(doff, qoff); up(dcol, qcol); up(dval, qval); up(drole, qrole);
    CC(cc_logpdf_bulk(&st, S, chains, Q, drow, doff, dcol, dval, drole, dout, dflag, nullptr));
    CK(cudaDeviceSynchronize());
    const std::vector<double> lp = down(dout, (size_t)S * Q);
    for (int s = 0; s < S; ++s) {
        printf("chain %d: log p(x0=0|x1=0) %.3f  log p(x0=6|x1=6) %.3f  log p(x0=0|x1=12) %.3f\n", s, lp[s * Q], lp[s * Q + 1], lp[s * Q + 2]);
        if (good[s]) ok &= lp[s * Q] > lp[s * Q + 2] + 3 && lp[s * Q + 1] > lp[s * Q + 2] + 3;      // matching clusters are far more likely
    }
    printf(ok ? "C driver OK\n" : "C driver FAILED\n");
    return ok ? 0 : 1;
}
