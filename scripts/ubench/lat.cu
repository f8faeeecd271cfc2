// Dependent-chain latencies of the instructions on the rows kernel's critical path (one warp).
#include <cstdio>
#include <cuda_runtime.h>
#define N 256
template <int OP> __global__ void k(double* out, double seed, long long* cyc) {
    double x = seed + threadIdx.x * 1e-3, y = seed * 0.5;
    __shared__ double sm[64]; __shared__ int si[64];
    sm[threadIdx.x] = x; si[threadIdx.x] = (threadIdx.x + 1) & 31; __syncwarp();
    int idx = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
        if (OP == 0) x = fma(x, y, 1e-9);
        if (OP == 1) x = log(x) + 3.0;
        if (OP == 2) x = exp(x) - 0.5 * x - 0.9;   // stays near 0.1..1
        if (OP == 3) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1e-9;
        if (OP == 4) x = __drcp_rn(x) + 0.5;
        if (OP == 5) { idx = si[idx]; }
        if (OP == 6) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, 1));
        if (OP == 7) x = 1.0 / x + 0.5;
        if (OP == 8) { float f = (float)x; asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(f) : "f"(f)); x = (double)f + 1e-9; }
        if (OP == 9) x = __dadd_rn(x, y);
        if (OP == 10) { asm volatile("bar.sync 1, 32;" ::: "memory"); x += sm[(threadIdx.x + i) & 31]; }
        if (OP == 11) { __threadfence_block(); x += 1.0; }
        if (OP == 12) x = lgamma(x) + 2.5;
        if (OP == 13) { float f = (float)x; f = __expf(f) * 0.3f + 0.1f; x = (double)f; }
        if (OP == 14) { unsigned b = __ballot_sync(0xffffffffu, x > 0.3); x += (double)(__ffs(b)) * 1e-9; }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + idx;
    if (threadIdx.x == 0) cyc[OP] = (t1 - t0);
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 4096); cudaMallocManaged(&cyc, 256);
    const char* names[] = {"dfma", "log", "exp", "shfl.f64+dadd", "__drcp_rn+dadd", "lds chase", "shfl.f64+fmax", "div+dadd", "cvt+redux.max.f32+cvt+dadd", "dadd", "bar.sync(1 warp)+lds+dadd", "membar.cta+dadd", "lgamma", "cvt+__expf+ffma+cvt", "ballot+ffs+i2d+dfma"};
    for (int rep = 0; rep < 2; ++rep) {
        k<0><<<1, 32>>>(out, 0.7, cyc); k<1><<<1, 32>>>(out, 0.7, cyc); k<2><<<1, 32>>>(out, 0.3, cyc); k<3><<<1, 32>>>(out, 0.7, cyc);
        k<4><<<1, 32>>>(out, 0.7, cyc); k<5><<<1, 32>>>(out, 0.7, cyc); k<6><<<1, 32>>>(out, 0.7, cyc); k<7><<<1, 32>>>(out, 0.7, cyc);
        k<8><<<1, 32>>>(out, 0.7, cyc); k<9><<<1, 32>>>(out, 0.7, cyc); k<10><<<1, 32>>>(out, 0.7, cyc); k<11><<<1, 32>>>(out, 0.7, cyc);
        k<12><<<1, 32>>>(out, 0.7, cyc); k<13><<<1, 32>>>(out, 0.7, cyc); k<14><<<1, 32>>>(out, 0.7, cyc);
        cudaDeviceSynchronize();
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    for (int i = 0; i < 15; ++i) printf("%-32s %7.1f cycles/iter\n", names[i], cyc[i] / (double)N);
    return 0;
}
