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
