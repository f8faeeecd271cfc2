// The rows kernels for tables that hold columns other than normal ones (categorical, bernoulli, poisson,
// exponential, geometric): rows.cu compiled with those types enabled (entry cc_transition_rows_counts, reached from
// cc_transition_rows).
#define CC_ROWS_COUNT 1
#include "rows.cu"
