// Shared device helpers for the CrossCat kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <math_constants.h>
#include "../../include/cgpm_b200.h"

#define CC_LOGPI   1.1447298858494001741434273513531
#define CC_LOG2    0.69314718055994530941723212145818
#define CC_LOG2PI  1.8378770664093454835606594728112

void cc_set_error(const char* fmt, ...);
int cc_check_cuda(cudaError_t e, const char* what);
void cc_prepare_pool();
#define CC_CUDA(x) do { if (cc_check_cuda((x), #x)) return CC_ERR_CUDA; } while (0)

// ---------------------------------------------------------------------------
// index helpers
__device__ __forceinline__ size_t sv_index(const cc_state& st, int s, int v) {
    return (size_t)s * st.Vcap + v;
}
__device__ __forceinline__ double* stat_field(const cc_state& st, int s, int f) {
    return st.stat + ((size_t)s * CC_NFIELD + f) * (size_t)st.Kcap * st.C;
}

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter-based generator for the device
// stream.  u01 maps two 32-bit words to a double exactly like numpy's legacy
// random_sample: (a>>5, b>>6) -> (a*2^26+b)/2^53, so the device stream has the
// same support as the injected host stream.
struct Philox {
    uint32_t k0, k1;
    __device__ __forceinline__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
    __device__ __forceinline__ uint4 operator()(uint64_t ctr_lo, uint64_t ctr_hi) const {
        uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32);
        uint32_t c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            uint32_t n0 = hi1 ^ c1 ^ a, n1 = lo1, n2 = hi0 ^ c3 ^ b, n3 = lo0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};
__device__ __forceinline__ double u01_from_words(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}
// stream ids (high counter word): purpose | chain | view/column
__device__ __forceinline__ uint64_t stream_id(uint32_t purpose, uint32_t chain, uint32_t sub) {
    return ((uint64_t)purpose << 56) | ((uint64_t)(chain & 0xFFFFFFu) << 32) | sub;
}

// ---------------------------------------------------------------------------
// Normal with Normal-Gamma prior (src/primitives/normal.py).  The posterior
// update keeps the reference's operand order and never contracts a*b+c into an
// FMA, so (mn, sn) are the same doubles the reference computes from the same
// sufficient statistics (normal.py:183-191).
__device__ __forceinline__ void normal_posterior(double N, double sx, double sxx, double m, double r,
                                                 double s, double nu, double& mn, double& rn,
                                                 double& sn, double& nun) {
    rn = __dadd_rn(r, N);
    nun = __dadd_rn(nu, N);
    mn = __ddiv_rn(__dadd_rn(__dmul_rn(r, m), sx), rn);
    double t = __dadd_rn(__dadd_rn(s, sxx), __dmul_rn(__dmul_rn(r, m), m));
    sn = __dsub_rn(t, __dmul_rn(__dmul_rn(rn, mn), mn));
    if (sn == 0.0) sn = s;                                   // normal.py:189-190
}
// normal.py:193-200
__device__ __forceinline__ double normal_log_Z(double r, double s, double nu) {
    return ((nu + 1.0) * 0.5) * CC_LOG2 + 0.5 * CC_LOGPI - 0.5 * log(r) - (nu * 0.5) * log(s) + lgamma(nu * 0.5);
}
// normal.py:175-181
__device__ __forceinline__ double normal_marginal(double N, double sx, double sxx, double m, double r,
                                                  double s, double nu) {
    double mn, rn, sn, nun;
    normal_posterior(N, sx, sxx, m, r, s, nu, mn, rn, sn, nun);
    return -(N * 0.5) * CC_LOG2PI + normal_log_Z(rn, sn, nun) - normal_log_Z(r, s, nu);
}
// log(x) for positive, normal, finite x, without a branch: fdlibm's __ieee754_log kernel (argument reduced
// to [sqrt(1/2), sqrt(2)), log(1+f) = f - f^2/2 + s (f^2/2 + R(s^2)), s = f / (2 + f)) with the division
// replaced by a correctly rounded reciprocal.  Within 1 ulp of the library log over [2^-20, 2^40]
// (scripts/ubench/logtest.cu); 228 instead of 276 dependent cycles on B200, and -- unlike two library
// calls, whose special-case branches keep them apart -- two independent evaluations overlap (366 cycles
// for both instead of 546).  The callers' arguments are 1 + a z^2 >= 1, a ratio in [2^-20, 1], a posterior
// scale or a ratio of positive counts.
__device__ __forceinline__ double cc_log_pos(double x) {
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                 Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                 Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    int hx = __double2hiint(x);
    const int lx = __double2loint(x);
    int k = (hx >> 20) - 1023;
    hx &= 0x000fffff;
    const int i = (hx + 0x95f64) & 0x100000;          // mantissa above sqrt(2): halve it
    k += i >> 20;
    const double m = __hiloint2double(hx | (i ^ 0x3ff00000), lx);
    const double f = m - 1.0;
    const double s = f * __drcp_rn(2.0 + f);
    const double dk = (double)k;
    const double z = s * s, w = z * z;
    const double t1 = w * fma(w, fma(w, Lg6, Lg4), Lg2);
    const double t2 = z * fma(w, fma(w, fma(w, Lg7, Lg5), Lg3), Lg1);
    const double R = t2 + t1;
    const double hfsq = 0.5 * f * f;
    return dk * ln2_hi - ((hfsq - fma(s, hfsq + R, dk * ln2_lo)) - f);
}
// G(N): the N-only part of the predictive Z(N+1)-Z(N)-log(2pi)/2 (normal.py:165-173):
//   predictive(x) = G(N) - log(sn)/2 - ((nun+1)/2) log1p(rn (x-mn)^2 / ((rn+1) sn))
// lgamma(h + 1/2) - lgamma(h) is evaluated directly: shift h up to >= 16 with
// Gamma(h+1) = h Gamma(h), then the asymptotic series
//   1/2 log h - 1/(8h) + 1/(192 h^3) - 1/(640 h^5) + 17/(14336 h^7) - 31/(18432 h^9) + 691/(180224 h^11)
// (coefficients (2^{1-k} - 2) B_k / (k (k-1)), k = 2,4,..,12; the next term is < 3e-18 at h = 16).
// Relative error <= 1e-15 against 40-digit arithmetic for h in [1e-2, 3e6]; the reference's
// difference of two libm lgamma values is itself only good to ~1e-9 at N ~ 1e6.  All logs of G
// are folded into one.
__device__ __forceinline__ double normal_G(double N, double r, double nu) {
    double h = 0.5 * (nu + N);
    const double rn = r + N;
    double num = rn + 1.0, den = rn;            // argument of the single log: h' (den/num)^2 rn/(rn+1)
    double pn = 1.0, pd = 1.0;
    while (h < 16.0) { pn *= (h + 0.5); pd *= h; h += 1.0; }
    num *= pn * pn;
    den *= pd * pd * h;
    const double ih = __drcp_rn(h), ih2 = ih * ih;
    const double poly = ih * (-0.125 + ih2 * (1.0 / 192.0 + ih2 * (-1.0 / 640.0 + ih2 * (17.0 / 14336.0 +
                        ih2 * (-31.0 / 18432.0 + ih2 * (691.0 / 180224.0))))));
    return 0.5 * cc_log_pos(den * __drcp_rn(num)) + poly - 0.5 * CC_LOGPI;
}
// G one count up or down by the recurrence lgamma(h + 1) - lgamma(h + 1/2) = log h - (lgamma(h + 1/2) - lgamma(h)):
//   G(N + 1) = -G(N) + 1/2 log(h^2 rn / (rn + 2)) - log(pi),   h = (nu + N) / 2, rn = r + N
// (one log instead of the series; read downwards it gives G(N - 1) from G(N)).  Used when a row moves; the
// rebuilds recompute G from the series, so rounding does not accumulate beyond the moves of one run of sweeps.
__device__ __forceinline__ double normal_G_step(double G_other, double N_lower, double r, double nu) {
    const double h = 0.5 * (nu + N_lower), rn = r + N_lower;
    return 0.5 * cc_log_pos(h * h * rn * __drcp_rn(rn + 2.0)) - CC_LOGPI - G_other;
}
struct NormalCache { double mn, a, cst; };
__device__ __forceinline__ NormalCache normal_cache(double N, double sx, double sxx, double gN, double m,
                                                    double r, double s, double nu) {
    double mn, rn, sn, nun;
    normal_posterior(N, sx, sxx, m, r, s, nu, mn, rn, sn, nun);
    NormalCache c;
    c.mn = mn;
    c.a = rn / ((rn + 1.0) * sn);
    c.cst = gN - 0.5 * log(sn);
    return c;
}
__device__ __forceinline__ double normal_predictive_cached(double x, double mn, double a, double cst,
                                                           double N, double nu) {
    double z = x - mn;
    return cst - 0.5 * (nu + N + 1.0) * log1p(a * z * z);
}

// Categorical with symmetric Dirichlet prior (src/primitives/categorical.py:144-155)
__device__ __forceinline__ double categorical_marginal(double N, const double* counts, int k, double alpha) {
    double A = k * alpha;
    double lg = 0.0;
    for (int j = 0; j < k; ++j) lg += lgamma(counts[j] + alpha);
    return lgamma(A) - lgamma(A + N) + lg - k * lgamma(alpha);
}

// ---------------------------------------------------------------------------
// The other collapsed primitives ("sum types": bernoulli, poisson, exponential, geometric).  Statistics: N,
// sum_x (CC_F_SX) and for poisson sum_log_fact_x (CC_F_SXX); two hyperparameters h0, h1 = (alpha, beta) or (a, b).
__device__ __forceinline__ bool cc_is_sum_type(int ct) { return ct != CC_CATEGORICAL; }       // ordered-sum statistics
__device__ __forceinline__ bool cc_is_count_type(int ct) { return ct >= CC_BERNOULLI; }
// the values a primitive's incorporate accepts / its logpdf gives mass to (bernoulli.py:49, poisson.py:52,
// exponential.py:50, geometric.py:50)
__device__ __forceinline__ bool count_valid(int ct, double x) {
    if (ct == CC_BERNOULLI) return x == 0.0 || x == 1.0;
    if (ct == CC_EXPONENTIAL) return x >= 0.0;
    return x >= 0.0 && x == floor(x);
}
__device__ __forceinline__ double cc_betaln(double a, double b) { return lgamma(a) + lgamma(b) - lgamma(a + b); }
// logpdf_score of one cluster: bernoulli.py:153-155, poisson.py:155-160, exponential.py:148-153, geometric.py:148-153
__device__ __forceinline__ double count_marginal(int ct, double N, double sx, double sxx, double h0, double h1) {
    if (ct == CC_BERNOULLI) return cc_betaln(sx + h0, N - sx + h1) - cc_betaln(h0, h1);
    if (ct == CC_GEOMETRIC) return cc_betaln(h0 + N, h1 + sx) - cc_betaln(h0, h1);
    const double an = (ct == CC_POISSON) ? h0 + sx : h0 + N, bn = (ct == CC_POISSON) ? h1 + N : h1 + sx;
    const double z = (lgamma(an) - an * log(bn)) - (lgamma(h0) - h0 * log(h1));
    return (ct == CC_POISSON) ? z - sxx : z;
}
// Out-of-line copies: the lgamma calls of these primitives need a stack frame and many registers, which must not leak
// into the kernels' normal / categorical hot paths (they run only for bernoulli / poisson / exponential / geometric cells).
static __device__ __noinline__ double count_marginal_ni(int ct, double N, double sx, double sxx, double h0, double h1) {
    return count_marginal(ct, N, sx, sxx, h0, h1);
}
// marginal of any sum type (the normal included)
__device__ __forceinline__ double sum_marginal(int ct, double N, double sx, double sxx, double h0, double h1, double h2, double h3) {
    return (ct == CC_NORMAL) ? normal_marginal(N, sx, sxx, h0, h1, h2, h3) : count_marginal_ni(ct, N, sx, sxx, h0, h1);
}
// the second accumulated term of a sum type: x*x (normal.py:69), gammaln(x+1) (poisson.py:56), nothing otherwise.
// lf: the host's gammaln(x + 1) of the cell when it is known (Xaux), else NaN -> the device lgamma
__device__ __forceinline__ double sum_second_term(int ct, double x, double lf) {
    if (ct == CC_NORMAL) return __dmul_rn(x, x);
    if (ct == CC_POISSON) return (lf == lf) ? lf : lgamma(x + 1.0);
    return 0.0;
}
__device__ __forceinline__ double cell_aux(const cc_state& st, int c, int r) {
    if (!st.Xaux || r < 0) return CUDART_NAN;
    const int a = st.auxcol[c];
    return (a < 0) ? CUDART_NAN : st.Xaux[(size_t)a * st.R + r];
}
// cached predictive constants (CC_F_MN, CC_F_A, CC_F_CST) of a sum-type cluster, so that the predictive of x is
//   bernoulli    (x == 1 ? MN : A) - CST             MN = log(sx + alpha), A = log(N - sx + beta), CST = log(N + alpha + beta)
//   poisson      lgamma(A + x) - (A + x) MN + CST - lgamma(x + 1)      A = a + sx, MN = log(b + N + 1), CST = A log(b + N) - lgamma(A)
//   exponential  CST - A log(MN + x)                 A = a + N + 1, MN = b + sx, CST = log(a + N) + (a + N) log(MN)
//   geometric    CST + lgamma(MN + x) - lgamma(A + x + 1)     MN = b + sx, A = a + N + MN, CST = log(a + N) - lgamma(MN) + lgamma(A)
// (bernoulli.py:141-151, poisson.py:147-153, exponential.py:140-146, geometric.py:140-146 with Z(M) - Z(N) written out)
__device__ __forceinline__ NormalCache count_cache(int ct, double N, double sx, double h0, double h1) {
    NormalCache c;
    if (ct == CC_BERNOULLI) { c.mn = log(sx + h0); c.a = log(N - sx + h1); c.cst = log(N + h0 + h1); }
    else if (ct == CC_POISSON) { c.a = h0 + sx; c.mn = log(h1 + N + 1.0); c.cst = c.a * log(h1 + N) - lgamma(c.a); }
    else if (ct == CC_EXPONENTIAL) { const double an = h0 + N; c.mn = h1 + sx; c.a = an + 1.0; c.cst = log(an) + an * log(c.mn); }
    else { const double an = h0 + N; c.mn = h1 + sx; c.a = an + c.mn; c.cst = log(an) - lgamma(c.mn) + lgamma(c.a); }
    return c;
}
__device__ __forceinline__ double count_pred_cached(int ct, double x, double mn, double a, double cst) {
    if (!count_valid(ct, x)) return -INFINITY;
    if (ct == CC_BERNOULLI) return ((x == 1.0) ? mn : a) - cst;
    if (ct == CC_POISSON) return lgamma(a + x) - (a + x) * mn + cst - lgamma(x + 1.0);
    if (ct == CC_EXPONENTIAL) return cst - a * log(mn + x);
    return cst + lgamma(mn + x) - lgamma(a + x + 1.0);
}
// the same predictive straight from the statistics (a row scored against its own cluster without itself)
__device__ __forceinline__ double count_pred_raw(int ct, double x, double N, double sx, double h0, double h1) {
    const NormalCache c = count_cache(ct, N, sx, h0, h1);
    return count_pred_cached(ct, x, c.mn, c.a, c.cst);
}

static __device__ __noinline__ double count_pred_cached_ni(int ct, double x, double mn, double a, double cst) { return count_pred_cached(ct, x, mn, a, cst); }
static __device__ __noinline__ double count_pred_raw_ni(int ct, double x, double N, double sx, double h0, double h1) { return count_pred_raw(ct, x, N, sx, h0, h1); }
static __device__ __noinline__ NormalCache count_cache_ni(int ct, double N, double sx, double h0, double h1) { return count_cache(ct, N, sx, h0, h1); }
static __device__ __noinline__ double sum_second_term_ni(int ct, double x, double lf) { return sum_second_term(ct, x, lf); }

// ---------------------------------------------------------------------------
// warp helpers
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_incl_scan(double v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// ---------------------------------------------------------------------------
// mbarrier / async-copy PTX (sm_90+; SASS: SYNCS.*, UBLKCP, LDGSTS)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{ .reg .b64 t; mbarrier.arrive.shared::cta.b64 t, [%0]; }" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{ .reg .b64 t; mbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1; }" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// same, for a thread that can afford to sleep between probes (the producer warp)
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(2000);
    }
}
// 1-D bulk copy global -> shared, completion on an mbarrier (TMA engine)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void cp_async_4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// the mbarrier receives one (pre-counted) arrival when all prior cp.async of this thread landed
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
