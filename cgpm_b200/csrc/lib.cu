// Error reporting and small utilities shared by the C-ABI entry points.
#include <cstdarg>
#include <cstdio>
#include "cc_common.cuh"

static thread_local char g_err[512] = "";

void cc_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cc_check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    cc_set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return 1;
}

extern "C" const char* cc_last_error(void) { return g_err; }
extern "C" int cc_abi_version(void) { return CC_ABI_VERSION; }

// Keep the stream-ordered allocator's memory across synchronisation points: with the default
// release threshold (0) every cudaMallocAsync after a sync goes back to the driver.
void cc_prepare_pool() {
    static int done[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    done[dev] = 1;
}
