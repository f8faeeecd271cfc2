// Batched predictive queries on non-composite CrossCat states:
//   logpdf    sampling.state_logpdf / view_logpdf / _logpdf_row  (src/crosscat/sampling.py:37-49,70-82,102-107)
//   simulate  sampling.state_simulate / view_simulate / _simulate_row (sampling.py:52-67,85-99,110-116),
//             Normal.simulate / sample_parameters (src/primitives/normal.py:85-94,202-206),
//             Categorical.simulate (src/primitives/categorical.py:76-82)
// Queries are shared by all listed chains (Engine.logpdf_bulk / simulate_bulk,
// src/crosscat/engine.py:202-210,241-251); cells are stored CSR: cell e of
// query q has column qcol[e], value qval[e] and role qrole[e] (0 target, 1 constraint).
#include "cc_common.cuh"

namespace {

// predictive log density of value x for column c under cluster slot k (or the
// empty cluster when k < 0), from the cached tables (dim.py:127-132)
__device__ double cell_logpdf(const cc_state& st, int s, int c, int k, double x) {
    if (isnan(x)) return 0.0;
    const double* hyp = st.hyp + ((size_t)s * st.C + c) * 4;
    const int slot = (k < 0) ? st.Kcap - 1 : k;
    const size_t e = (size_t)slot * st.C + c;
    if (st.ctype[c] == CC_NORMAL)
        return normal_predictive_cached(x, stat_field(st, s, CC_F_MN)[e], stat_field(st, s, CC_F_A)[e],
                                        stat_field(st, s, CC_F_CST)[e], (k < 0) ? 0.0 : stat_field(st, s, CC_F_N)[e], hyp[3]);
    if (st.ctype[c] != CC_CATEGORICAL)       // bernoulli.py:62-68, poisson.py:68-74, exponential.py:60-66, geometric.py:60-66
        return count_pred_cached_ni(st.ctype[c], x, stat_field(st, s, CC_F_MN)[e], stat_field(st, s, CC_F_A)[e], stat_field(st, s, CC_F_CST)[e]);
    const int kc = st.ncat[c];
    if (!(x == floor(x) && x >= 0 && x < kc)) return -INFINITY;            // categorical.py:71-72
    const double a = hyp[0];
    if (k < 0) return log(a) - log(a * kc);
    const double cnt = st.cnt[((size_t)s * st.Kcap + k) * st.CT + st.catoff[c] + (int)x];
    return log(a + cnt) - stat_field(st, s, CC_F_CST)[e];
}

struct QueryParams {
    cc_state st;
    int n_chains, Q;
    const int32_t* chains;
    const int32_t* rowid;     // [Q] observed rowid or -1
    const int32_t* qoff;      // [Q+1]
    const int32_t* qcol;
    const double* qval;
    const uint8_t* qrole;
    double* out;              // [n_chains][Q]
    int32_t* flag;            // [n_chains][Q] 1 = zero-density constraints (sampling.py:78-79)
};

// one warp per (chain, query); lanes over candidate clusters
__global__ void __launch_bounds__(128) logpdf_kernel(const QueryParams p) {
    const cc_state& st = p.st;
    const int lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gw >= (long long)p.n_chains * p.Q) return;
    const int ci = (int)(gw / p.Q), q = (int)(gw % p.Q);
    const int s = p.chains[ci];
    const int beg = p.qoff[q], end = p.qoff[q + 1];
    const int rid = p.rowid[q];
    const int32_t* zv = st.zv + (size_t)s * st.C;
    double total = 0.0;
    int bad = 0;
    for (int a = beg; a < end; ++a) {
        if (p.qrole[a] != 0) continue;                       // views are those of the targets
        const int v = zv[p.qcol[a]];
        bool seen = false;
        for (int b = beg; b < a; ++b) if (p.qrole[b] == 0 && zv[p.qcol[b]] == v) seen = true;
        if (seen) continue;
        const size_t sv = sv_index(st, s, v);
        if (rid >= 0) {
            // observed row: joint density of the targets in its own cluster
            const int k = st.zr[sv * st.R + rid];
            double acc = 0.0;
            for (int b = beg + lane; b < end; b += 32)
                if (p.qrole[b] == 0 && zv[p.qcol[b]] == v) acc += cell_logpdf(st, s, p.qcol[b], k, p.qval[b]);
            total += warp_sum(acc);
            continue;
        }
        const int K = st.nclu[sv];
        const int32_t* live = st.live + sv * st.Kcap;
        const int32_t* nk = st.nk + sv * st.Kcap;
        const double alpha = st.view_alpha[sv];
        const double lden = log((double)st.R + alpha);
        // pass 1: log-normaliser of the cluster posterior given the constraints
        double mx = -INFINITY;
        for (int p0 = 0; p0 <= K; p0 += 32) {
            const int pc = p0 + lane;
            if (pc <= K) {
                const int k = (pc < K) ? live[pc] : -1;
                double lc = log((pc < K) ? (double)nk[k] : alpha) - lden;    // crp.py:178-182
                for (int b = beg; b < end; ++b)
                    if (p.qrole[b] == 1 && zv[p.qcol[b]] == v) lc += cell_logpdf(st, s, p.qcol[b], k, p.qval[b]);
                mx = fmax(mx, lc);
            }
        }
        mx = warp_max(mx);
        if (mx == -INFINITY) { bad = 1; continue; }
        double se = 0.0, mt = -INFINITY;
        // pass 2: sum exp(lc - mx), and the max of lc + lt for the outer logsumexp
        for (int p0 = 0; p0 <= K; p0 += 32) {
            const int pc = p0 + lane;
            if (pc <= K) {
                const int k = (pc < K) ? live[pc] : -1;
                double lc = log((pc < K) ? (double)nk[k] : alpha) - lden, lt = 0.0;
                for (int b = beg; b < end; ++b) {
                    if (zv[p.qcol[b]] != v) continue;
                    const double l = cell_logpdf(st, s, p.qcol[b], k, p.qval[b]);
                    if (p.qrole[b] == 1) lc += l; else lt += l;
                }
                se += exp(lc - mx);
                mt = fmax(mt, lc + lt);
            }
        }
        const double Z = mx + log(warp_sum(se));
        mt = warp_max(mt);
        if (mt == -INFINITY) { total += -INFINITY; continue; }
        double st2 = 0.0;
        for (int p0 = 0; p0 <= K; p0 += 32) {
            const int pc = p0 + lane;
            if (pc <= K) {
                const int k = (pc < K) ? live[pc] : -1;
                double lc = log((pc < K) ? (double)nk[k] : alpha) - lden, lt = 0.0;
                for (int b = beg; b < end; ++b) {
                    if (zv[p.qcol[b]] != v) continue;
                    const double l = cell_logpdf(st, s, p.qcol[b], k, p.qval[b]);
                    if (p.qrole[b] == 1) lc += l; else lt += l;
                }
                st2 += exp(lc + lt - mt);
            }
        }
        total += mt + log(warp_sum(st2)) - Z;
    }
    if (lane == 0) {
        p.out[(size_t)ci * p.Q + q] = total;
        p.flag[(size_t)ci * p.Q + q] = bad;
    }
}

// --------------------------------------------------------------------------- simulate
struct Rng {                      // per-thread Philox stream
    Philox ph;
    uint64_t ctr, sid;
    uint4 buf;
    int have;
    __device__ Rng(uint64_t seed, uint64_t c0, uint64_t s) : ph(seed), ctr(c0), sid(s), have(0) {}
    __device__ double uniform() {
        if (have == 0) { buf = ph(ctr++, sid); have = 2; }
        const double u = (have == 2) ? u01_from_words(buf.x, buf.y) : u01_from_words(buf.z, buf.w);
        --have;
        return u;
    }
    __device__ double normal() {  // Box-Muller
        double u1 = uniform(), u2 = uniform();
        if (u1 < 1e-300) u1 = 1e-300;
        return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    }
    __device__ double poisson(double lam) {   // Knuth's product of uniforms in chunks of rate <= 30 (sums of Poissons are Poisson)
        double x = 0.0;
        while (lam > 0.0) {
            const double l = fmin(lam, 30.0);
            lam -= l;
            const double lim = exp(-l);
            double prod = uniform();
            while (prod > lim) { x += 1.0; prod *= uniform(); }
        }
        return x;
    }
    __device__ double gamma(double a) {   // Marsaglia-Tsang (2000), shape a, scale 1
        if (a < 1.0) {
            const double u = uniform();
            return gamma(a + 1.0) * pow(u > 0 ? u : 1e-300, 1.0 / a);
        }
        const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
        for (;;) {
            double x, v;
            do { x = normal(); v = 1.0 + c * x; } while (v <= 0.0);
            v = v * v * v;
            const double u = uniform();
            if (u < 1.0 - 0.0331 * x * x * x * x) return d * v;
            if (log(u > 0 ? u : 1e-300) < 0.5 * x * x + d * (1.0 - v + log(v))) return d * v;
        }
    }
};

struct SimParams {
    cc_state st;
    int n_chains, Q;
    const int32_t* chains;
    const int32_t* rowid;     // [Q]
    const int32_t* qoff;      // [Q+1] cells (targets role 0 with value ignored, constraints role 1)
    const int32_t* qcol;
    const double* qval;
    const uint8_t* qrole;
    const int32_t* soff;      // [Q+1] sample offsets (N per query)
    int total_samples, max_targets;
    double* out;              // [2][n_chains][total_samples][max_targets]: values, then cluster slots, in target order
    int32_t* flag;            // [n_chains][Q]
    uint64_t seed, counter;
};

// one thread per (chain, query, sample)
__global__ void __launch_bounds__(128) simulate_kernel(const SimParams p) {
    const cc_state& st = p.st;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)p.n_chains * p.total_samples) return;
    const int ci = (int)(gid / p.total_samples), smp = (int)(gid % p.total_samples);
    // query of this sample (binary search in soff)
    int lo = 0, hi = p.Q;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (p.soff[mid] <= smp) lo = mid; else hi = mid; }
    const int q = lo;
    const int s = p.chains[ci];
    const int beg = p.qoff[q], end = p.qoff[q + 1];
    const int rid = p.rowid[q];
    const int32_t* zv = st.zv + (size_t)s * st.C;
    Rng rng(p.seed, p.counter << 20, ((uint64_t)8 << 56) | ((uint64_t)(uint32_t)ci << 32) | (uint32_t)smp);
    double* out = p.out + ((size_t)ci * p.total_samples + smp) * p.max_targets;
    // second plane of the output: the cluster (slot, -1 = fresh cluster) each target was drawn from
    double* kplane = p.out + ((size_t)(p.n_chains + ci) * p.total_samples + smp) * p.max_targets;
    int tpos = 0;
    for (int a = beg; a < end; ++a) {
        if (p.qrole[a] != 0) continue;
        const int c = p.qcol[a];
        const int v = zv[c];
        const size_t sv = sv_index(st, s, v);
        // cluster of this sample in view v: drawn once per view; an earlier target of the
        // same view left it in the cluster plane of the output
        int k = -2;
        {
            int tprev = 0;
            for (int b = beg; b < a; ++b)
                if (p.qrole[b] == 0) { if (zv[p.qcol[b]] == v && k == -2) k = (int)__double_as_longlong(kplane[tprev]); ++tprev; }
        }
        if (k == -2) {
            if (rid >= 0) k = st.zr[sv * st.R + rid];
            else {
                const int K = st.nclu[sv];
                const int32_t* live = st.live + sv * st.Kcap;
                const int32_t* nk = st.nk + sv * st.Kcap;
                const double alpha = st.view_alpha[sv];
                // weights nk * exp(constraint logpdf) (the common 1/(N+alpha) cancels)
                double mx = -INFINITY;
                for (int pc = 0; pc <= K; ++pc) {
                    const int kk = (pc < K) ? live[pc] : -1;
                    double lc = log((pc < K) ? (double)nk[kk] : alpha);
                    for (int b = beg; b < end; ++b)
                        if (p.qrole[b] == 1 && zv[p.qcol[b]] == v) lc += cell_logpdf(st, s, p.qcol[b], kk, p.qval[b]);
                    mx = fmax(mx, lc);
                }
                if (mx == -INFINITY) { p.flag[(size_t)ci * p.Q + q] = 1; k = -1; }
                else {
                    double tot = 0.0;
                    for (int pc = 0; pc <= K; ++pc) {
                        const int kk = (pc < K) ? live[pc] : -1;
                        double lc = log((pc < K) ? (double)nk[kk] : alpha);
                        for (int b = beg; b < end; ++b)
                            if (p.qrole[b] == 1 && zv[p.qcol[b]] == v) lc += cell_logpdf(st, s, p.qcol[b], kk, p.qval[b]);
                        tot += exp(lc - mx);
                    }
                    const double thr = rng.uniform() * tot;
                    double run = 0.0;
                    k = -1;
                    for (int pc = 0; pc <= K; ++pc) {
                        const int kk = (pc < K) ? live[pc] : -1;
                        double lc = log((pc < K) ? (double)nk[kk] : alpha);
                        for (int b = beg; b < end; ++b)
                            if (p.qrole[b] == 1 && zv[p.qcol[b]] == v) lc += cell_logpdf(st, s, p.qcol[b], kk, p.qval[b]);
                        run += exp(lc - mx);
                        if (run > thr) { k = kk; break; }
                    }
                }
            }
        }
        // sample the cell given the cluster
        const double* hyp = st.hyp + ((size_t)s * st.C + c) * 4;
        const int slot = (k < 0) ? st.Kcap - 1 : k;
        const size_t e = (size_t)slot * st.C + c;
        double x;
        if (st.ctype[c] == CC_NORMAL) {
            const double N = (k < 0) ? 0.0 : stat_field(st, s, CC_F_N)[e];
            const double sx = (k < 0) ? 0.0 : stat_field(st, s, CC_F_SX)[e];
            const double sxx = (k < 0) ? 0.0 : stat_field(st, s, CC_F_SXX)[e];
            double mn, rn, sn, nun;
            normal_posterior(N, sx, sxx, hyp[0], hyp[1], hyp[2], hyp[3], mn, rn, sn, nun);
            const double rho = rng.gamma(nun * 0.5) * (2.0 / sn);          // normal.py:202-206
            const double mu = mn + rng.normal() / sqrt(rho * rn);
            x = mu + rng.normal() / sqrt(rho);                             // normal.py:93
        } else if (st.ctype[c] != CC_CATEGORICAL) {
            const int ct = st.ctype[c];
            const double N = (k < 0) ? 0.0 : stat_field(st, s, CC_F_N)[e];
            const double sx = (k < 0) ? 0.0 : stat_field(st, s, CC_F_SX)[e];
            if (ct == CC_BERNOULLI) {                                      // bernoulli.py:70-80
                x = (rng.uniform() * (N + hyp[0] + hyp[1]) < sx + hyp[0]) ? 1.0 : 0.0;
            } else if (ct == CC_POISSON) {                                 // poisson.py:76-84: negative_binomial(an, bn / (bn + 1))
                const double an = hyp[0] + sx, bn = hyp[1] + N;
                x = rng.poisson(rng.gamma(an) / bn);                       //   = Poisson(Gamma(an, scale (1 - p) / p)), p = bn / (bn + 1)
            } else if (ct == CC_EXPONENTIAL) {                             // exponential.py:68-77
                const double mu = rng.gamma(hyp[0] + N) / (hyp[1] + sx);
                double u = rng.uniform();
                x = -log(u > 0 ? u : 1e-300) / mu;
            } else {                                                       // geometric.py:68-75: geometric(pn) - 1, pn ~ Beta(an, bn)
                const double ga = rng.gamma(hyp[0] + N), gb = rng.gamma(hyp[1] + sx);
                const double pn = ga / (ga + gb);
                double u = rng.uniform();
                x = (pn >= 1.0) ? 0.0 : floor(log(u > 0 ? u : 1e-300) / log1p(-pn));
            }
        } else {
            const int kc = st.ncat[c];
            const double a = hyp[0];
            const double* cn = st.cnt + ((size_t)s * st.Kcap + slot) * st.CT + st.catoff[c];
            double tot = 0.0;
            for (int j = 0; j < kc; ++j) tot += ((k < 0) ? 0.0 : cn[j]) + a;
            const double thr = rng.uniform() * tot;
            double run = 0.0;
            int pick = kc - 1;
            for (int j = 0; j < kc; ++j) { run += ((k < 0) ? 0.0 : cn[j]) + a; if (run > thr) { pick = j; break; } }
            x = (double)pick;
        }
        kplane[tpos] = __longlong_as_double((long long)k);
        out[tpos] = x;
        ++tpos;
    }
}

template <typename T>
int q_h2d(T** out, const T* host, size_t n, cudaStream_t stream) {
    *out = nullptr;
    if (n == 0 || !host) return CC_OK;
    CC_CUDA(cudaMallocAsync(out, sizeof(T) * n, stream));
    CC_CUDA(cudaMemcpyAsync(*out, host, sizeof(T) * n, cudaMemcpyHostToDevice, stream));
    return CC_OK;
}

}  // namespace

extern "C" int cc_logpdf_bulk(const cc_state* st, int n_chains, const int32_t* chains, int Q,
                              const int32_t* rowid_dev, const int32_t* qoff_dev, const int32_t* qcol_dev,
                              const double* qval_dev, const uint8_t* qrole_dev, double* out_dev,
                              int32_t* flag_dev, void* stream_) {
    if (!st || n_chains < 0 || Q < 0) { cc_set_error("cc_logpdf_bulk: bad arguments"); return CC_ERR_ARG; }
    if (n_chains == 0 || Q == 0) return CC_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    cc_prepare_pool();
    QueryParams p;
    p.st = *st; p.n_chains = n_chains; p.Q = Q; p.rowid = rowid_dev; p.qoff = qoff_dev; p.qcol = qcol_dev;
    p.qval = qval_dev; p.qrole = qrole_dev; p.out = out_dev; p.flag = flag_dev;
    int32_t* dch = nullptr;
    if (q_h2d(&dch, chains, n_chains, stream)) return CC_ERR_CUDA;
    p.chains = dch;
    const long long warps = (long long)n_chains * Q;
    const long long blocks = (warps + 3) / 4;
    if (blocks > 0x7fffffffLL) { cc_set_error("cc_logpdf_bulk: too many (chain, query) pairs for one launch"); return CC_ERR_ARG; }
    logpdf_kernel<<<(unsigned)blocks, 128, 0, stream>>>(p);
    CC_CUDA(cudaGetLastError());
    cudaFreeAsync(dch, stream);
    return CC_OK;
}

extern "C" int cc_simulate_bulk(const cc_state* st, int n_chains, const int32_t* chains, int Q,
                                const int32_t* rowid_dev, const int32_t* qoff_dev, const int32_t* qcol_dev,
                                const double* qval_dev, const uint8_t* qrole_dev, const int32_t* soff_dev,
                                int total_samples, int max_targets, double* out_dev, int32_t* flag_dev,
                                uint64_t seed, uint64_t counter, void* stream_) {
    if (!st || n_chains < 0 || Q < 0 || max_targets < 1) { cc_set_error("cc_simulate_bulk: bad arguments"); return CC_ERR_ARG; }
    if (n_chains == 0 || Q == 0 || total_samples == 0) return CC_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    cc_prepare_pool();
    SimParams p;
    p.st = *st; p.n_chains = n_chains; p.Q = Q; p.rowid = rowid_dev; p.qoff = qoff_dev; p.qcol = qcol_dev;
    p.qval = qval_dev; p.qrole = qrole_dev; p.soff = soff_dev; p.total_samples = total_samples;
    p.max_targets = max_targets; p.out = out_dev; p.flag = flag_dev; p.seed = seed; p.counter = counter;
    int32_t* dch = nullptr;
    if (q_h2d(&dch, chains, n_chains, stream)) return CC_ERR_CUDA;
    p.chains = dch;
    const long long threads = (long long)n_chains * total_samples;
    const long long blocks = (threads + 127) / 128;
    if (blocks > 0x7fffffffLL) { cc_set_error("cc_simulate_bulk: too many samples for one launch"); return CC_ERR_ARG; }
    simulate_kernel<<<(unsigned)blocks, 128, 0, stream>>>(p);
    CC_CUDA(cudaGetLastError());
    cudaFreeAsync(dch, stream);
    return CC_OK;
}
