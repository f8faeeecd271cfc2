// Isolated copy of the rows kernel's fast-path scoring step: how long does one dependent
// pass (smem loads -> Student-t argument -> log -> store) take for a single warp?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int iters, int variant) {
    __shared__ double tab[8 * 16]; __shared__ int live[16], nk[16], zrow[64]; __shared__ double xr[64], lp[34];
    const int tid = threadIdx.x;
    for (int i = tid; i < 128; i += 32) tab[i] = 1.0 + 0.01 * i;
    if (tid < 16) { live[tid] = tid; nk[tid] = 100 + tid; }
    for (int i = tid; i < 64; i += 32) { zrow[i] = i & 3; xr[i] = 0.1 * i; }
    __syncwarp();
    const int Kc = 16, D = 1, fs = Kc * D;
    const double nu = 2.0, rr = 1.5;
    long long t0 = clock64();
    double sink = 0.0;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
        const int slot = i & 63;
        const int z = zrow[slot];
        const int K = 5 + (i & 1);
        const int nkz = nk[z];
        const bool singleton = nkz == 1;
        const int ncand = K + (singleton ? 0 : 1);
        const int pc = tid;
        const bool active = pc < ncand;
        const int k = (pc < K) ? live[pc] : Kc - 1;
        const bool own = (pc < K) && (k == z);
        double acc = 0.0;
        if (active) {
            const double x = xr[slot];
            const double* p = tab + k;
            const double Nf = p[0 * fs], zz = x - p[3 * fs];
            double sc = p[4 * fs], bs = p[5 * fs], coef = -0.5 * (nu + Nf + 1.0);
            if (own && variant) {
                const double rn = rr + Nf;
                sc = -sc * ((rn + 1.0) * __drcp_rn(rn - 1.0)) * 1e-3;
                coef = 0.5 * (nu + Nf - 1.0);
                bs += p[7 * fs] - p[6 * fs];
            }
            acc = bs + coef * log(fma(sc * zz, zz, 1.0));
        }
        if (active) lp[pc] = acc;
        __syncwarp();
        sink += lp[(i + 1) % 5];     // consumer of the stored scores (keeps the iterations dependent)
        zrow[slot] = (z + (sink > 1e300)) & 3;
    }
    long long t1 = clock64();
    out[tid] = sink;
    if (tid == 0) cyc[variant] = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 4096); cudaMallocManaged(&cyc, 64);
    for (int rep = 0; rep < 2; ++rep) { k<<<1, 32>>>(out, cyc, 4096, 0); k<<<1, 32>>>(out, cyc, 4096, 1); cudaDeviceSynchronize(); }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    printf("scoring pass, other clusters only: %.1f cycles/row;  with own-cluster branch: %.1f cycles/row\n", cyc[0] / 4096.0, cyc[1] / 4096.0);
    return 0;
}
