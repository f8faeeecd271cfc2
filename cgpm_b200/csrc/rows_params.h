// Launch parameters shared by the rows kernels (rows.cu, rows_wide.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/cgpm_b200.h"

struct RowsParams {
    cc_state st;
    const cc_row_work* work;
    const int32_t* perm;
    const double* u;
    double* trace;
    int trace_stride;
    uint64_t seed, counter;
    int B;          // rows per staged batch (power of two, <= 32)
    int bulk;       // 1: stage whole rows with cp.async.bulk
    int smem_bytes; // dynamic shared memory given to the CTA
    int wide;       // 1: very wide views -- per-column hypers / types stay in global memory, tables too
    int prof;       // debug: per-CTA cycles / D / K / moves into st.seq[(chain, view)][0..3]
};

// rows_wide.cu: two-rows-in-flight sweep for views wider than 16 columns
int cc_rows_wide_fixed_smem(int Kcap, int dmax, int wide, int xbytes);
int cc_launch_rows_wide(const RowsParams& p, int cnt, cudaStream_t stream);
