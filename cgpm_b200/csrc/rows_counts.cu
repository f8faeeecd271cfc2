// The rows kernels for tables that hold bernoulli / poisson / exponential / geometric columns: rows.cu compiled with
// those primitives enabled (entry cc_transition_rows_counts, reached from cc_transition_rows).
#define CC_ROWS_COUNT 1
#include "rows.cu"
