/* cgpm_b200.h -- C ABI of the B200 CrossCat Gibbs hot path.
 *
 * Drop-in boundary.  The reference (probcomp/cgpm, pure Python) has exactly one
 * precedent for handing `State.transition` to native code:
 * `State.transition_lovecat` (src/crosscat/state.py:812-831) which exports the
 * state to dense arrays (src/crosscat/lovecat.py:73-223), calls
 * `LocalEngine.analyze(M_c, T, X_L, X_D, kernel_list, n_steps, ...)`
 * (lovecat.py:377-391) and writes the partitions back (lovecat.py:226-290).
 * The entry points below are that interface for a GPU: dense chain state in
 * device memory + one call per CrossCat kernel.  No torch types appear here;
 * every pointer inside `cc_state` is a plain CUDA device pointer owned by the
 * caller (the Python host allocates them as torch tensors, a C++ host would
 * use cudaMalloc).  All functions return 0 on success, CC_ERR_* otherwise;
 * `cc_last_error()` gives the message.  They enqueue on `stream` (a
 * cudaStream_t passed as void*) and do not synchronise unless stated.
 *
 * Array layout (S chains, R rows, C columns, Vcap view slots per chain, Kcap
 * cluster slots per view, CT = sum of categories over categorical columns):
 * see `cc_state` below.  "slot" = dense index; the reference's sparse ids
 * (view ids, cluster ids: crp.py:58-59,142) are carried in `view_id`/`cid`.
 */
#ifndef CGPM_B200_H
#define CGPM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CC_ABI_VERSION 4
#define CC_NGRID 30          /* hyper-grid points, dim.py:176 n_grid=30        */
#define CC_NFIELD 8          /* per (cluster, column) statistic fields          */

/* column types (config.py:25-39): the two named by the scope contract, and the other collapsed primitives of
 * SURVEY.md 8(f) rank 3 ("sum types": their statistics are N, sum_x and, for poisson, sum_log_fact_x) */
#define CC_NORMAL 0
#define CC_CATEGORICAL 1
#define CC_BERNOULLI 2       /* src/primitives/bernoulli.py    hypers alpha, beta */
#define CC_POISSON 3         /* src/primitives/poisson.py      hypers a, b        */
#define CC_EXPONENTIAL 4     /* src/primitives/exponential.py  hypers a, b        */
#define CC_GEOMETRIC 5       /* src/primitives/geometric.py    hypers a, b        */

/* statistic fields of `stat[S][CC_NFIELD][Kcap][C]`                         */
#define CC_F_N    0   /* count of non-NaN members   (normal.py:51, categorical.py:47) */
#define CC_F_SX   1   /* sum_x                      (normal.py:52; x_sum / sum_x of the other sum types)  */
#define CC_F_SXX  2   /* sum_x_sq                   (normal.py:53); poisson: sum_log_fact_x (poisson.py:56) */
#define CC_F_MN   3   /* cached posterior mean      (normal.py:186); sum types: see cc_common.cuh count_cache */
#define CC_F_A    4   /* cached rn/((rn+1) sn)                      */
#define CC_F_CST  5   /* cached constant: normal G(N) - log(sn)/2 ; categorical log(N + alpha k) */
#define CC_F_GN   6   /* cached G(N)   = lgamma((nu+N+1)/2)-lgamma((nu+N)/2)+log((r+N)/(r+N+1))/2-log(pi)/2 */
#define CC_F_GM1  7   /* cached G(N-1)                               */

/* rows_prog[..][2] */
#define CC_ROWS_DONE      0
#define CC_ROWS_HANDOVER  1
#define CC_ROWS_GROW      2

#define CC_OK            0
#define CC_ERR_ARG       1   /* invalid argument -> Python ValueError           */
#define CC_ERR_CUDA      2   /* CUDA failure     -> Python RuntimeError         */
#define CC_ERR_CAPACITY  3   /* Vcap/Kcap exceeded in some chain                */

typedef struct cc_state {
    int32_t R, C, Cp, S, Vcap, Kcap, CT, _pad;
    /* ---- the table (shared by all chains; read-only to kernels) ---------- */
    const double*  Xrm;            /* [R][Cp] row-major, NaN = missing; Cp = C rounded up to even (16-byte rows) */
    const double*  Xcm;            /* [C][R]  column-major copy                  */
    const int32_t* ctype;          /* [C]   CC_NORMAL / CC_CATEGORICAL           */
    const int32_t* ncat;           /* [C]   k of categorical(k), else 0          */
    const int32_t* catoff;         /* [C]   offset of the column's counts in `cnt` rows */
    const double*  grids;          /* [C][4][CC_NGRID] normal: m,r,s,nu ; categorical: alpha in row 0 ; other types: their
                                      two hypers in rows 0, 1.  A grid shorter than CC_NGRID (poisson a: poisson.py:122-124)
                                      is padded with NaN: such points have no mass */
    const double*  Xaux;           /* [n_aux][R] per-cell constants some column types add to their statistics, column-major:
                                      poisson gammaln(x + 1) as the host computes it (poisson.py:56), or NULL */
    const int32_t* auxcol;         /* [C] row of Xaux belonging to the column, -1 = none (NULL when Xaux is NULL) */
    const double*  row_alpha_grid; /* [CC_NGRID] crp.py:148-152 with n = R       */
    const double*  col_alpha_grid; /* [CC_NGRID] with n = C                      */
    /* ---- per chain -------------------------------------------------------- */
    double*  hyp;        /* [S][C][4]   m,r,s,nu | alpha,-,-,-                    */
    int32_t* zv;         /* [S][C]      view slot of each column                  */
    int32_t* nv;         /* [S][Vcap]   number of columns per view slot (0 = free)*/
    int32_t* view_id;    /* [S][Vcap]   reference view id, -1 = free slot         */
    double*  col_alpha;  /* [S]         column-CRP alpha                          */
    double*  view_alpha; /* [S][Vcap]   row-CRP alpha per view                    */
    double*  view_grid;  /* [S][Vcap][CC_NGRID] alpha grid of each view's row CRP: fixed when the view is
                            created from the row count at that time (view.py:98-99) */
    int32_t* nclu;       /* [S][Vcap]   live clusters K                           */
    int32_t* live;       /* [S][Vcap][Kcap] cluster slots in ascending cluster-id order */
    int32_t* cid;        /* [S][Vcap][Kcap] reference cluster id of a slot        */
    int32_t* nk;         /* [S][Vcap][Kcap] members per cluster slot              */
    int32_t* zr;         /* [S][Vcap][R]    cluster slot of each row              */
    int32_t* order;      /* [S][Vcap][R]    member order of the view's CRP (crp.py:45-59) */
    double*  stat;       /* [S][CC_NFIELD][Kcap][C]                               */
    double*  cnt;        /* [S][Kcap][CT]   categorical counts                    */
    int32_t* err;        /* [S]  sticky per-chain error code (CC_ERR_CAPACITY)    */
    int32_t* rows_prog;  /* [S][Vcap][4]  progress of the view's rows sweep: {next row position, rows re-seated so
                            far, status, kilocycles the sweep took}.  status CC_ROWS_DONE after a complete sweep; CC_ROWS_HANDOVER while a
                            sweep is passed from the shared-memory kernel to the large-K kernel inside one
                            cc_transition_rows call; CC_ROWS_GROW when the sweep stopped in front of a move that
                            needs a cluster slot beyond Kcap: the caller re-allocates the per-cluster arrays with a
                            larger Kcap and calls cc_transition_rows again with `resume` = 1 (same seed, counter
                            and injected buffers) -- the sweep continues with the same row, nothing is lost */
    /* ---- scratch (caller-allocated, contents undefined between calls) ----- */
    int32_t* seq;        /* [S][Vcap][R]  spare scratch (per-CTA / per-pair cycle counters of the debug probes) */
    int32_t* moved;      /* [S][Vcap][R]  rows re-seated in the current sweep, in order */
    uint8_t* movedflag;  /* [S][Vcap][R]                                          */
    double*  colL;       /* [S][C][Vcap+1] column marginal matrix (last = auxiliary view) */
    int32_t* aux_zr;     /* [S][R]   auxiliary-view partition (one column step)   */
    int32_t* aux_K;      /* [S]      number of tables                             */
    double*  aux_alpha;  /* [S][C]   alpha drawn for each column's auxiliary view */
    int32_t* aux_zr_all; /* [S][C][R] every column's auxiliary partition, or NULL when it does not fit */
    int32_t* aux_K_all;  /* [S][C]   number of tables of each stored partition    */
    void*    arena;      /* per-worker scratch for the auxiliary CRP simulation   */
    int64_t  arena_bytes;
} cc_state;

/* One unit of work of the rows kernel: a (chain, view slot) pair and where its
 * injected randomness lives.  perm/u offsets index `perm`/`u` of
 * cc_transition_rows (element offsets), or -1 to use the device generator. */
typedef struct cc_row_work {
    int32_t chain, view;
    int32_t n;            /* number of rows to visit                             */
    int32_t perm_is_index;/* 1: perm[i] indexes `order` (rows=None, view.py:249-250); 0: perm[i] is a rowid */
    int64_t perm_off;     /* into `perm`, or -1: device pseudo-random permutation of `order` */
    int64_t u_off;        /* into `u`,    or -1: Philox uniforms                  */
    int64_t trace_off;    /* into `trace`, or -1                                  */
    int32_t n_cols;       /* number of columns of the view if the caller knows it (sizes shared memory), else 0 */
    int32_t n_normal;     /* how many of them are normal columns, if the caller knows it, else 0 */
    int32_t resume;       /* 0: a fresh sweep; 1: continue the sweep recorded in rows_prog (items whose status is
                             CC_ROWS_DONE are skipped) */
    int32_t cost_hint;    /* expected cost of the sweep in any unit (e.g. rows_prog[..][3] of the view's last sweep), or
                             0: work items are started in descending order of (cost_hint, n_cols) so that the longest
                             sweeps never start last when there are more items than resident CTAs */
} cc_row_work;

const char* cc_last_error(void);
int cc_abi_version(void);

/* Recompute every cached field of `stat` and the counts from (X, zr, order,
 * hyp) for the listed (chain, view slot) pairs, summing each cluster in the
 * view's member order -- the device form of View._bulk_incorporate
 * (src/mixtures/view.py:483-494).  views == NULL: all live views of all chains
 * (synchronises the stream once to size the launch). */
int cc_rebuild(const cc_state* st, int n_work, const int32_t* chains /*host*/,
               const int32_t* views /*host*/, const uint8_t* mask /*device [S][C] or NULL:
               rebuild only these columns, state.py:1117-1121 re-incorporates only the
               visited column*/, void* stream);

/* Recompute only the cached predictive constants (and the empty-cluster record) of the
 * listed views' columns from their sufficient statistics and hypers, leaving the sums
 * untouched: used after incremental edits (State.incorporate / unincorporate / force_cell,
 * src/crosscat/state.py:255-315) that update the statistics with the reference's own `+=`. */
int cc_refresh_tables(const cc_state* st, int n_work, const int32_t* chains /*host*/,
                      const int32_t* views /*host*/, void* stream);

/* Debug invariants of the partitions (View._check_partitions / State._check_partitions,
 * src/mixtures/view.py:528-560, src/crosscat/state.py:1162-1184, enabled in the reference by
 * the GPMCCDEBUG environment variable): out_dev[S] = 0, or the code of the first violated
 * invariant of that chain. */
int cc_check_state(const cc_state* st, int32_t* out_dev, void* stream);

/* State.logpdf_score (src/crosscat/state.py:384-387): column-CRP marginal +
 * per view [row-CRP marginal + sum of cluster marginals].  out: device [S]. */
int cc_logpdf_score(const cc_state* st, double* out, void* stream);

/* View.transition_rows for every work item (src/mixtures/view.py:247-252,
 * 393-429).  `perm`, `u`, `trace` are device buffers (may be NULL when no work
 * item refers to them).  seed/counter key the device generator.  A sweep that
 * runs out of cluster slots stops with rows_prog status CC_ROWS_GROW (see
 * cc_state.rows_prog); callers whose Kcap is below R + 1 check it after the call. */
int cc_transition_rows(const cc_state* st, int n_work, const cc_row_work* work /*host*/,
                       const int32_t* perm, const double* u, double* trace,
                       int trace_stride, uint64_t seed, uint64_t counter,
                       void* stream);

/* ---- State.transition_dims (src/crosscat/state.py:804-810,1050-1154) --------
 * The column sweep is exposed as its four stages; the host runs
 *   cc_col_aux -> cc_col_marginals -> { cc_col_scan -> [cc_col_aux(write_zr) ->
 *   cc_col_install -> cc_col_marginals(new view)] }* -> cc_rebuild(mask)
 * (cgpm_b200/core.py: transition_columns).  Arguments named *_dev are device
 * pointers; the others are host arrays copied on `stream`. */

/* Auxiliary view of each listed (chain, column): alpha = row_alpha_grid[aidx]
 * (view.py:99; aidx == NULL: device generator), CRP-prior partition of all rows
 * (view.py:100-103, crp.py:71-81; uniforms u_dev[n][R-1] or device generator),
 * the column's marginal under it -> colL[chain][col][Vcap], aux_alpha.  With
 * write_zr = 1 the partition is kept in aux_zr/aux_K of the chain (one column
 * per chain per call), with write_zr = 2 in aux_zr_all/aux_K_all.  Needs st->arena. */
int cc_col_aux(const cc_state* st, int n, const int32_t* chains, const int32_t* cols,
               const int32_t* aidx, const double* u_dev, int write_zr, uint64_t seed,
               uint64_t counter, void* stream);

/* Marginal likelihood of columns under existing views (state.py:1065-1072 ->
 * view.py:483-494 -> dim.py:121-122): for work item i, every column of its list
 * cols_dev[coloff[i] .. coloff[i]+colcnt[i]) (cols_dev == NULL: all columns)
 * under view views[i] of chain chains[i] -> colL[chain][col][view]. */
int cc_col_marginals(const cc_state* st, int n_work, const int32_t* chains, const int32_t* views,
                     const int32_t* cols_dev, const int32_t* coloff, const int32_t* colcnt,
                     int max_cols, void* stream);

/* The sequential scan (state.py:1087-1121): for chain chains_dev[i], columns
 * cols_dev[i][pos_dev[i] .. ncols) in order.  A move into the auxiliary view
 * stops the chain: pending_dev[i] = list position (else -1 when finished).
 * u_dev[n_chains][ncols]: injected uniforms, or NULL for the device generator.
 * ci_off/ci_adj: CSR of the independence constraints Ci (state.py:1098-1102). */
int cc_col_scan(const cc_state* st, int n_chains, const int32_t* chains_dev, const int32_t* cols_dev,
                int ncols, int32_t* pos_dev, int32_t* pending_dev, const double* u_dev,
                uint64_t seed, uint64_t counter, const int32_t* ci_off_dev, const int32_t* ci_adj_dev,
                double* trace_dev, int trace_stride, void* stream);

/* Install aux_zr/aux_K of each listed chain (or the stored aux_zr_all/aux_K_all of the pair) as a new view and move the
 * column into it (state.py:1110-1116,1125-1154). */
int cc_col_install(const cc_state* st, int n, const int32_t* chains, const int32_t* cols,
                   int mode /* 1: aux_zr of the chain, 2: aux_zr_all of the (chain, column) */, void* stream);

/* ---- grid Gibbs on hyperparameters (src/mixtures/dim.py:152-174) ------------ */

/* 'column_hypers' (state.py:781-786) for each listed (chain, column).  hord[n][4]:
 * hyper indices (0..3 = m,r,s,nu | 0 = alpha) in visiting order, -1 padded --
 * the host's rng.shuffle of the names (dim.py:154-155) -- or NULL for a device
 * shuffle; u[n][4]: one uniform per visited hyper, or NULL.  Host arrays. */
int cc_transition_col_hypers(const cc_state* st, int n, const int32_t* chains, const int32_t* cols,
                             const int32_t* hord, const double* u, uint64_t seed, uint64_t counter,
                             void* stream);

/* 'alpha' (views == NULL, ndraw = 1; state.py:763-765) and 'view_alphas'
 * (views[n], ndraw = 2; state.py:767-772, view.py:231-233).  u[n][ndraw] host
 * uniforms or NULL.  The grid scores do not depend on the current alpha, so the
 * value kept is the one selected by the last uniform. */
int cc_transition_alphas(const cc_state* st, int n, const int32_t* chains, const int32_t* views,
                         const double* u, int ndraw, uint64_t seed, uint64_t counter, void* stream);

/* ---- the whole kernel schedule behind one call ---------------------------------
 * The reference's native precedent takes the kernel list and the number of steps in ONE call:
 * LocalEngine.analyze(M_c, T, X_L, X_D, seed, kernel_list, n_steps, ...) (src/crosscat/lovecat.py:377-391,
 * reached from State.transition_lovecat, src/crosscat/state.py:812-831).  cc_transition is that call for the
 * device-resident state: n_iters iterations of the kernels listed (in the caller's order, State.transition /
 * _transition_generic, state.py:727-761,866-905) over the listed chains, all rows and all columns, on the device
 * random stream keyed by (seed, *counter); *counter is advanced once per kernel call.  A C or C++ host needs
 * nothing else to run inference: cc_rebuild once after loading a state, then cc_transition, then the queries. */
#define CC_K_ALPHA          0   /* 'alpha'          state.py:763-765 */
#define CC_K_VIEW_ALPHAS    1   /* 'view_alphas'    state.py:767-772 */
#define CC_K_COLUMN_HYPERS  2   /* 'column_hypers'  state.py:781-786 */
#define CC_K_ROWS           3   /* 'rows'           state.py:795-802 */
#define CC_K_COLUMNS        4   /* 'columns'        state.py:804-810; needs cc_state.arena (cc_col_aux) */

/* Capacity is the caller's memory.  When a rows sweep needs more cluster slots, or a view birth needs more cluster
 * or view slots, cc_transition synchronises the stream and calls `grow`: it must re-allocate the per-cluster /
 * per-view arrays with at least the requested capacities (contents carried over: slots keep their numbers, the
 * empty-cluster record of `stat` / `cnt` moves to the new last slot, the auxiliary column of colL to the new
 * last position) and update the cc_state passed to cc_transition in place; return 0 on success.  The schedule
 * then continues exactly where it stopped.  With grow == NULL running out of capacity returns CC_ERR_CAPACITY
 * (never needed with Kcap = R + 1 and Vcap = C). */
typedef int (*cc_grow_fn)(void* user, int32_t min_kcap, int32_t min_vcap);

int cc_transition(cc_state* st, int n_chains, const int32_t* chains /*host*/, int n_iters, int n_kernels,
                  const int32_t* kernels /*host, CC_K_* */, const int32_t* ci_off_dev, const int32_t* ci_adj_dev
                  /* CSR of the independence constraints or NULL, as in cc_col_scan */, uint64_t seed,
                  uint64_t* counter, cc_grow_fn grow, void* user, void* stream);

/* ---- batched queries (src/crosscat/sampling.py:37-116) ------------------------
 * Queries are CSR: cells qoff[q]..qoff[q+1] of query q have column qcol, value
 * qval and role qrole (0 target, 1 constraint); rowid[q] is an observed rowid
 * or -1 for a hypothetical row.  Every listed chain answers every query
 * (Engine.logpdf_bulk / simulate_bulk, src/crosscat/engine.py:202-210,241-251). */

/* out_dev[n_chains][Q] = State.logpdf; flag_dev[n_chains][Q] = 1 where every
 * cluster gives the constraints zero density (ValueError, sampling.py:78-79). */
int cc_logpdf_bulk(const cc_state* st, int n_chains, const int32_t* chains /*host*/, int Q,
                   const int32_t* rowid_dev, const int32_t* qoff_dev, const int32_t* qcol_dev,
                   const double* qval_dev, const uint8_t* qrole_dev, double* out_dev,
                   int32_t* flag_dev, void* stream);

/* soff_dev[Q+1]: sample offsets (N samples per query).  out_dev is
 * [2][n_chains][total_samples][max_targets]: plane 0 the sampled values in the
 * query's target order, plane 1 the cluster slot each was drawn from. */
int cc_simulate_bulk(const cc_state* st, int n_chains, const int32_t* chains /*host*/, int Q,
                     const int32_t* rowid_dev, const int32_t* qoff_dev, const int32_t* qcol_dev,
                     const double* qval_dev, const uint8_t* qrole_dev, const int32_t* soff_dev,
                     int total_samples, int max_targets, double* out_dev, int32_t* flag_dev,
                     uint64_t seed, uint64_t counter, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CGPM_B200_H */
