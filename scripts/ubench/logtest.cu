// Branch-free double log for positive normal arguments (fdlibm's __ieee754_log kernel with the
// division replaced by a reciprocal): latency of one and of two interleaved evaluations, and
// the worst error against the CUDA library log over a sweep of arguments.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__device__ __forceinline__ double log_pn(double x) {
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
template <int OP> __global__ void lat(double* out, double seed, long long* cyc) {
    double x = seed + threadIdx.x * 1e-3, y = seed * 1.7 + threadIdx.x * 1e-3;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) {
        if (OP == 0) x = log(x) + 3.0;
        if (OP == 1) x = log_pn(x) + 3.0;
        if (OP == 2) { const double a = log(x), b = log(y); x = a + 3.0; y = b + 3.1; }
        if (OP == 3) { const double a = log_pn(x), b = log_pn(y); x = a + 3.0; y = b + 3.1; }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + y;
    if (threadIdx.x == 0) cyc[OP] = t1 - t0;
}
__global__ void acc(double* worst) {
    double w = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (1 << 24); i += gridDim.x * blockDim.x) {
        const double x = exp2(-20.0 + 60.0 * (double)i / (double)(1 << 24)) * (1.0 + 1e-3 * (i & 1023));
        const double a = log(x), b = log_pn(x);
        const double e = fabs(a - b) / fmax(fabs(a), 1e-300);
        const double ulp = fabs(a - b) / (fabs(a) * 2.220446049250313e-16 + 1e-320);
        if (fabs(a) > 1e-3) w = fmax(w, ulp);
        (void)e;
    }
    atomicMax((unsigned long long*)worst, (unsigned long long)__double_as_longlong(w));
}
int main() {
    double *out, *worst; long long* cyc; cudaMalloc(&out, 4096); cudaMallocManaged(&cyc, 64); cudaMallocManaged(&worst, 8); *worst = 0;
    for (int rep = 0; rep < 2; ++rep) { lat<0><<<1, 32>>>(out, 0.7, cyc); lat<1><<<1, 32>>>(out, 0.7, cyc); lat<2><<<1, 32>>>(out, 0.7, cyc); lat<3><<<1, 32>>>(out, 0.7, cyc); cudaDeviceSynchronize(); }
    acc<<<256, 256>>>(worst); cudaDeviceSynchronize();
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    printf("log: %.1f  log_pn: %.1f  two logs: %.1f  two log_pn: %.1f cycles/iter; worst |log - log_pn| = %.2f ulp of log\n",
           cyc[0] / 256.0, cyc[1] / 256.0, cyc[2] / 256.0, cyc[3] / 256.0, *worst);
    return 0;
}
