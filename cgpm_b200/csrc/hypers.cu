// Grid-Gibbs on hyperparameters: Dim.transition_hypers (src/mixtures/dim.py:152-174)
// used for 'column_hypers' (src/crosscat/state.py:781-786), the column-CRP
// 'alpha' (state.py:763-765) and 'view_alphas' (state.py:767-772,
// src/mixtures/view.py:231-233: two transitions per view).
//
// For each hyper (in a shuffled order) the score of each of the 30 grid values
// is the sum over the column's clusters of the marginal likelihood
// (normal.py:175-181, categorical.py:150-155, crp.py:184-188) with that value
// substituted; one value is drawn by log_pflip with a uniform grid prior.
// Sufficient statistics are not touched; the cached predictive constants of
// the column (and its empty-cluster record) are refreshed at the end.
#include "cc_common.cuh"

namespace {

struct HyperParams {
    cc_state st;
    int n;
    const int32_t* chains;
    const int32_t* cols;
    const int32_t* hord;     // [n][4] hyper visiting order (-1 padded) or NULL: device shuffle
    const double* u;         // [n][4] uniforms (indexed by visiting position) or NULL
    uint64_t seed, counter;
};

// draw an index from 30 log-weights held in shared memory (single warp)
__device__ int draw_grid(const double* lp, double uu, int lane) {
    double v = (lane < CC_NGRID) ? lp[lane] : -INFINITY;
    const double mx = warp_max(v);
    const double e = (lane < CC_NGRID) ? exp(v - mx) : 0.0;
    const double inc = warp_incl_scan(e, lane);
    const double tot = __shfl_sync(0xffffffffu, inc, 31);
    const unsigned hit = __ballot_sync(0xffffffffu, lane < CC_NGRID && inc > uu * tot);
    return hit ? (__ffs(hit) - 1) : -1;
}

template <int NW>
__global__ void __launch_bounds__(NW * 32) col_hyper_kernel(const HyperParams p) {
    const cc_state& st = p.st;
    const int w = blockIdx.x;
    const int s = p.chains[w], c = p.cols[w];
    const int C = st.C, Kcap = st.Kcap, CT = st.CT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int v = st.zv[(size_t)s * C + c];
    const size_t sv = sv_index(st, s, v);
    const int K = st.nclu[sv];
    const int32_t* live = st.live + sv * Kcap;
    const int ctype = st.ctype[c], kc = st.ncat[c];
    const double* grid = st.grids + (size_t)c * 4 * CC_NGRID;
    double* hyp = st.hyp + ((size_t)s * C + c) * 4;
    __shared__ double s_lp[32];
    __shared__ double s_h[4];
    __shared__ int s_ord[4];
    const int nh = (ctype == CC_NORMAL) ? 4 : (ctype == CC_CATEGORICAL) ? 1 : 2;      // m r s nu | alpha | (alpha beta) or (a b)
    const Philox rng(p.seed);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) s_h[i] = hyp[i];
        if (p.hord) for (int i = 0; i < 4; ++i) s_ord[i] = p.hord[(size_t)w * 4 + i];
        else {
            // Fisher-Yates over the hyper names (dim.py:154-155)
            for (int i = 0; i < 4; ++i) s_ord[i] = (i < nh) ? i : -1;
            const uint4 r4 = rng(p.counter, stream_id(6, s, c));
            const uint32_t rw[3] = {r4.x, r4.y, r4.z};
            for (int i = nh - 1; i >= 1; --i) {
                const int j = (int)(((uint64_t)rw[i - 1] * (uint32_t)(i + 1)) >> 32);
                const int t = s_ord[i]; s_ord[i] = s_ord[j]; s_ord[j] = t;
            }
        }
    }
    __syncthreads();
    const double* fN = stat_field(st, s, CC_F_N);
    const double* fSX = stat_field(st, s, CC_F_SX);
    const double* fSXX = stat_field(st, s, CC_F_SXX);
    for (int step = 0; step < nh; ++step) {
        const int h = s_ord[step];
        if (h < 0) break;
        // scores of the grid values: warp per grid point, lanes over clusters
        for (int g = warp; g < CC_NGRID; g += NW) {
            double hh[4] = {s_h[0], s_h[1], s_h[2], s_h[3]};
            hh[h] = grid[h * CC_NGRID + g];
            double acc = 0.0;
            for (int pos = lane; pos < K; pos += 32) {
                const int k = live[pos];
                const size_t e = (size_t)k * C + c;
                if (ctype != CC_CATEGORICAL) acc += sum_marginal(ctype, fN[e], fSX[e], fSXX[e], hh[0], hh[1], hh[2], hh[3]);
                else acc += categorical_marginal(fN[e], st.cnt + ((size_t)s * Kcap + k) * CT + st.catoff[c], kc, hh[0]);
            }
            acc = warp_sum(acc);
            // a grid shorter than CC_NGRID is padded with NaN (poisson.py:122-124): no mass there
            if (lane == 0) s_lp[g] = (hh[h] == hh[h]) ? acc : -INFINITY;
        }
        __syncthreads();
        if (warp == 0) {
            double uu;
            if (p.u) uu = p.u[(size_t)w * 4 + step];
            else { const uint4 r4 = rng(p.counter * 8 + 1 + step, stream_id(6, s, c)); uu = u01_from_words(r4.x, r4.y); }
            const int idx = draw_grid(s_lp, uu, lane);
            if (lane == 0) {
                if (idx < 0) st.err[s] = CC_ERR_ARG;
                else s_h[h] = grid[h * CC_NGRID + idx];
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) for (int i = 0; i < 4; ++i) hyp[i] = s_h[i];
    // refresh the cached predictive constants of the column (clusters + empty record)
    const double h0 = s_h[0], h1 = s_h[1], h2 = s_h[2], h3 = s_h[3];
    for (int pos = threadIdx.x; pos < K + 1; pos += NW * 32) {
        const int k = (pos < K) ? live[pos] : (Kcap - 1);
        const size_t e = (size_t)k * C + c;
        const double N = (pos < K) ? fN[e] : 0.0;
        if (ctype == CC_NORMAL) {
            const double sx = (pos < K) ? fSX[e] : 0.0, sxx = (pos < K) ? fSXX[e] : 0.0;
            const double gN = normal_G(N, h1, h3);
            const double gM1 = (N >= 1.0) ? normal_G(N - 1.0, h1, h3) : 0.0;
            const NormalCache cc = normal_cache(N, sx, sxx, gN, h0, h1, h2, h3);
            stat_field(st, s, CC_F_MN)[e] = cc.mn;
            stat_field(st, s, CC_F_A)[e] = cc.a;
            stat_field(st, s, CC_F_CST)[e] = cc.cst;
            stat_field(st, s, CC_F_GN)[e] = gN;
            stat_field(st, s, CC_F_GM1)[e] = gM1;
            if (pos >= K) { stat_field(st, s, CC_F_N)[e] = 0.0; stat_field(st, s, CC_F_SX)[e] = 0.0; stat_field(st, s, CC_F_SXX)[e] = 0.0; }
        } else if (ctype != CC_CATEGORICAL) {
            const double sx = (pos < K) ? fSX[e] : 0.0;
            const NormalCache cc = count_cache_ni(ctype, N, sx, h0, h1);
            stat_field(st, s, CC_F_MN)[e] = cc.mn;
            stat_field(st, s, CC_F_A)[e] = cc.a;
            stat_field(st, s, CC_F_CST)[e] = cc.cst;
            if (pos >= K) { stat_field(st, s, CC_F_N)[e] = 0.0; stat_field(st, s, CC_F_SX)[e] = 0.0; stat_field(st, s, CC_F_SXX)[e] = 0.0; }
        } else {
            stat_field(st, s, CC_F_CST)[e] = log(N + h0 * kc);
            if (pos >= K) stat_field(st, s, CC_F_N)[e] = 0.0;
        }
    }
}

// crp.py:184-188 on a grid; one warp per item.  kind 0: column CRP of a chain,
// kind 1: row CRP of (chain, view).
struct AlphaParams {
    cc_state st;
    int n;
    const int32_t* chains;
    const int32_t* views;    // NULL: column CRP
    const double* u;         // [n][ndraw] or NULL
    int ndraw;               // 1 (state.py:763-765) or 2 (view.py:231-233); the last draw is kept
    uint64_t seed, counter;
};

__global__ void __launch_bounds__(32) alpha_kernel(const AlphaParams p) {
    const cc_state& st = p.st;
    const int w = blockIdx.x, lane = threadIdx.x;
    const int s = p.chains[w];
    __shared__ double s_lp[32];
    double sumlg = 0.0;
    int K = 0, N = 0;
    const double* grid;
    if (!p.views) {
        const int32_t* nv = st.nv + (size_t)s * st.Vcap;
        for (int v = lane; v < st.Vcap; v += 32) if (nv[v] > 0) { sumlg += lgamma((double)nv[v]); ++K; }
        N = st.C;
        grid = st.col_alpha_grid;
    } else {
        const size_t sv = sv_index(st, s, p.views[w]);
        const int Kv = st.nclu[sv];
        const int32_t* live = st.live + sv * st.Kcap;
        const int32_t* nk = st.nk + sv * st.Kcap;
        for (int i = lane; i < Kv; i += 32) { sumlg += lgamma((double)nk[live[i]]); ++K; }
        N = st.R;
        grid = st.view_grid + sv * CC_NGRID;
    }
    sumlg = warp_sum(sumlg);
    K = __reduce_add_sync(0xffffffffu, K);
    if (lane < CC_NGRID) {
        const double a = grid[lane];
        s_lp[lane] = K * log(a) + sumlg + lgamma(a) - lgamma(N + a);
    }
    __syncwarp();
    const Philox rng(p.seed);
    double uu;
    const int last = p.ndraw - 1;
    if (p.u) uu = p.u[(size_t)w * p.ndraw + last];
    else { const uint4 r4 = rng(p.counter * 4 + last, stream_id(7, s, p.views ? (uint32_t)p.views[w] + 1 : 0)); uu = u01_from_words(r4.x, r4.y); }
    const int idx = draw_grid(s_lp, uu, lane);
    if (lane == 0) {
        if (idx < 0) st.err[s] = CC_ERR_ARG;
        else if (!p.views) st.col_alpha[s] = grid[idx];
        else st.view_alpha[sv_index(st, s, p.views[w])] = grid[idx];
    }
}

template <typename T>
int h2d(T** out, const T* host, size_t n, cudaStream_t stream) {
    *out = nullptr;
    if (n == 0 || !host) return CC_OK;
    CC_CUDA(cudaMallocAsync(out, sizeof(T) * n, stream));
    CC_CUDA(cudaMemcpyAsync(*out, host, sizeof(T) * n, cudaMemcpyHostToDevice, stream));
    return CC_OK;
}

}  // namespace

extern "C" int cc_transition_col_hypers(const cc_state* st, int n, const int32_t* chains, const int32_t* cols,
                                        const int32_t* hord, const double* u, uint64_t seed, uint64_t counter,
                                        void* stream_) {
    if (!st || n < 0) { cc_set_error("cc_transition_col_hypers: bad arguments"); return CC_ERR_ARG; }
    if (n == 0) return CC_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    cc_prepare_pool();
    HyperParams p;
    p.st = *st; p.n = n; p.seed = seed; p.counter = counter;
    int32_t *dch = nullptr, *dco = nullptr, *dho = nullptr;
    double* du = nullptr;
    if (h2d(&dch, chains, n, stream) || h2d(&dco, cols, n, stream) || h2d(&dho, hord, hord ? (size_t)n * 4 : 0, stream) ||
        h2d(&du, u, u ? (size_t)n * 4 : 0, stream)) return CC_ERR_CUDA;
    p.chains = dch; p.cols = dco; p.hord = dho; p.u = du;
    col_hyper_kernel<8><<<n, 256, 0, stream>>>(p);
    CC_CUDA(cudaGetLastError());
    cudaFreeAsync(dch, stream); cudaFreeAsync(dco, stream);
    if (dho) cudaFreeAsync(dho, stream);
    if (du) cudaFreeAsync(du, stream);
    return CC_OK;
}

extern "C" int cc_transition_alphas(const cc_state* st, int n, const int32_t* chains, const int32_t* views,
                                    const double* u, int ndraw, uint64_t seed, uint64_t counter, void* stream_) {
    if (!st || n < 0 || ndraw < 1) { cc_set_error("cc_transition_alphas: bad arguments"); return CC_ERR_ARG; }
    if (n == 0) return CC_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    cc_prepare_pool();
    AlphaParams p;
    p.st = *st; p.n = n; p.ndraw = ndraw; p.seed = seed; p.counter = counter;
    int32_t *dch = nullptr, *dvw = nullptr;
    double* du = nullptr;
    if (h2d(&dch, chains, n, stream) || h2d(&dvw, views, views ? n : 0, stream) || h2d(&du, u, u ? (size_t)n * ndraw : 0, stream)) return CC_ERR_CUDA;
    p.chains = dch; p.views = dvw; p.u = du;
    alpha_kernel<<<n, 32, 0, stream>>>(p);
    CC_CUDA(cudaGetLastError());
    cudaFreeAsync(dch, stream);
    if (dvw) cudaFreeAsync(dvw, stream);
    if (du) cudaFreeAsync(du, stream);
    return CC_OK;
}
