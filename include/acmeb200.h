/*
 * acmeb200.h -- C ABI of the B200-native batched DK-method runtime.
 *
 * The reference (HSU-ANT/ACME.jl) has no FFI seam: it is pure Julia and its hot
 * path is the method
 *     run!(runner::ModelRunner{<:DiscreteModel,false}, y, u)   src/ACME.jl:658-664
 * which calls step! (src/ACME.jl:666-715) per sample, which in turn calls
 *     solve(model.solvers[idx], p)                               src/ACME.jl:687
 * on a HomotopySolver{CachingSolver{SimpleSolver}} (src/solvers.jl:207-236,
 * 268-296, 347-396) backed by a KDTree (src/kdtree.jl:192-234).
 * This header is what a Julia `ccall` binding for that path would bind; the
 * binding itself (julia/ACMEB200.jl) is shown in INTEGRATION.md.
 *
 * Conventions
 *  - every matrix is Float64, column-major, exactly the Julia `Matrix{Float64}`
 *    memory of the corresponding DiscreteModel field (src/ACME.jl:118-137);
 *  - an `acmeb200_array` may be shared by all instances (stride 0) or hold one
 *    copy per instance (`stride` = distance in doubles between instances);
 *  - the non-linear element closures of the reference (src/elements.jl) cannot
 *    cross a C ABI, so each sub-problem carries an element table instead
 *    (kind + q offset + parameters), in the order of CircuitNLFunc
 *    (src/circuit.jl:68-86);
 *  - all functions return 0 on success or a negative ACMEB200_E* code; the
 *    message is available from acmeb200_last_error() (thread local);
 *  - numerical failure is not an error: it is reported per instance in the
 *    status words (mirrors the @warn / error() of src/ACME.jl:688-694).
 *  - threading: one host thread per model handle at a time; the library never
 *    calls back into the host runtime.
 */
#ifndef ACMEB200_H
#define ACMEB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACMEB200_ABI_VERSION 1

/* error codes */
#define ACMEB200_OK           0
#define ACMEB200_EINVAL      -1  /* bad argument / inconsistent descriptor  */
#define ACMEB200_ECUDA       -2  /* CUDA runtime error (see last_error)      */
#define ACMEB200_ENOMEM      -3
#define ACMEB200_EUNSUPPORTED -4 /* dimensions beyond what the kernels hold  */
#define ACMEB200_ENODEVICE   -5  /* no CUDA device: there is no CPU fallback */

/* element kinds and their parameter layouts (all Float64)
 *  DIODE       src/elements.jl:236-245  {is, eta}                               nq=2 nn=1
 *  BJT         src/elements.jl:309-406  {ise, isc, eta_e, eta_c, beta_f, beta_r,
 *                                        ile, ilc, eta_el, eta_cl, vaf, var, ikf, ikr}
 *                                                                               nq=4 nn=2
 *  POT         src/elements.jl:20-31    {r}                                     nq=5 nn=2
 *  MOSFET      src/elements.jl:436-481  {polarity, lambda, n_vt, n_alpha,
 *                                        vt[4], alpha[4]}                       nq=3 nn=1
 *  OPAMP_TANH  src/elements.jl:536-551  {gain, scale}                           nq=2 nn=1
 *  JA          src/elements.jl:104-135  {Ms, a, alpha, c, k}                    nq=4 nn=1
 *  TEST_QUAD   test/runtests.jl:207-219 (res = q1^2 - 1 + q2), no parameters    nq=2 nn=1
 */
#define ACMEB200_ELEM_DIODE       1
#define ACMEB200_ELEM_BJT         2
#define ACMEB200_ELEM_POT         3
#define ACMEB200_ELEM_MOSFET      4
#define ACMEB200_ELEM_OPAMP_TANH  5
#define ACMEB200_ELEM_JA          6
#define ACMEB200_ELEM_TEST_QUAD 100

/* solver selection; the reference default is HOMOTOPY_CACHING
 * (src/ACME.jl:150).  With HOMOTOPY_CACHING every instance keeps the reference's
 * learning solution cache on the device: insertion when a solve needed > 5
 * iterations (src/solvers.jl:374-386), the k-d tree of src/kdtree.jl rebuilt on the
 * reference's schedule (src/solvers.jl:387-394), start point = nearest of {origin,
 * newest entries, tree} (src/solvers.jl:347-371) -- the same start points, hence the
 * same iteration counts, as the reference; the physical capacity is bounded
 * (cache_capacity).  Alternatively a frozen, host-built k-d tree can be supplied per
 * sub-problem (acmeb200_cache), which is then used read-only by all instances. */
#define ACMEB200_SOLVER_SIMPLE            0 /* SimpleSolver                          */
#define ACMEB200_SOLVER_HOMOTOPY          1 /* HomotopySolver{SimpleSolver}          */
#define ACMEB200_SOLVER_HOMOTOPY_CACHING  2 /* HomotopySolver{CachingSolver{Simple}} */

/* per-instance status bits (src/ACME.jl:688-694) */
#define ACMEB200_STATUS_NOT_CONVERGED 1u /* "Failed to converge while solving non-linear equation." */
#define ACMEB200_STATUS_NONFINITE     2u /* "... got non-finite result." -- instance halted        */

/* run flags */
#define ACMEB200_U_DEVICE 1u /* U is a device pointer (else host)  */
#define ACMEB200_Y_DEVICE 2u /* Y is a device pointer (else host)  */
/* Sample-major ("instance-fastest") streams, SURVEY.md section 8(d) "I/O layout": U is
 * (nu, B, N) and Y is (ny, B, N) column-major, i.e. sample n of instance b is the nu
 * values at U + n*u_stride + b*nu; u_stride / y_stride are then the distances in doubles
 * between consecutive SAMPLES (>= nu*B / ny*B; y_stride == 0 -> ny*B; u_stride == 0 still
 * means one nu x N input shared by all instances).  Every time step of a warp's 32
 * instances is one contiguous segment, so the tiles of the thread-per-instance kernels
 * move as 256-byte rows.  Only models run by those kernels (specialised or generic) accept
 * the flag -- the lane-parallel kernels (warp / sub-warp group per instance) gain nothing
 * from it and return ACMEB200_EUNSUPPORTED; results are bit-identical to the default layout. */
#define ACMEB200_SAMPLE_MAJOR 4u

#define ACMEB200_HIST_BINS 32 /* iteration histogram: bins 1..31, last bin = 32 and more */

typedef struct acmeb200_array {
    const double *ptr;
    int64_t       stride; /* doubles between consecutive instances; 0 = shared */
} acmeb200_array;

typedef struct acmeb200_elem {
    int32_t kind;         /* ACMEB200_ELEM_*                                   */
    int32_t q_offset;     /* first q index of this element within the sub      */
    int32_t param_offset; /* first parameter within the sub's parameter vector */
    int32_t nparam;
} acmeb200_elem;

/* frozen solution cache of one sub-problem: the fields of KDTree
 * (src/kdtree.jl:4-9) + CachingSolver.zs (src/solvers.jl:321-323);
 * indices are 1-based as in the reference.  n_points == 0: no cache. */
typedef struct acmeb200_cache {
    int32_t        n_points;
    int32_t        n_columns; /* columns of ps/zs (>= every index in ps_idx)   */
    const int32_t *cut_dim;   /* n_points-1                                    */
    const double  *cut_val;   /* n_points-1                                    */
    const int32_t *ps_idx;    /* n_points                                      */
    const double  *ps;        /* np x n_columns, column-major                  */
    const double  *zs;        /* nn x n_columns, column-major                  */
} acmeb200_cache;

typedef struct acmeb200_sub_desc {
    int32_t nn, nq, np, nelem;
    acmeb200_array dq;      /* np x nx        model.dqs[i]      ACME.jl:124 */
    acmeb200_array eq;      /* np x nu        model.eqs[i]      ACME.jl:125 */
    acmeb200_array fqprev;  /* np x nn_total  model.fqprevs[i]  ACME.jl:126 */
    acmeb200_array pexp;    /* nq x np        model.pexps[i]    ACME.jl:123 */
    acmeb200_array q0;      /* nq             model.q0s[i]      ACME.jl:128 */
    acmeb200_array fq;      /* nq x nn        model.fqs[i]      ACME.jl:127 */
    acmeb200_array init_z;  /* nn   initial solution at p = 0   ACME.jl:259 */
    const acmeb200_elem *elems; /* nelem entries                              */
    acmeb200_array params;  /* nparams element parameters (per instance for sweeps) */
    int32_t nparams;
    int32_t reserved;
    acmeb200_cache cache;
} acmeb200_sub_desc;

typedef struct acmeb200_model_desc {
    int32_t abi_version; /* ACMEB200_ABI_VERSION */
    int32_t nx, nu, ny, nsub;
    int32_t solver;      /* ACMEB200_SOLVER_*                                  */
    int32_t maxiter;     /* 0 -> 500 (src/solvers.jl:207)                      */
    int32_t cache_capacity; /* stored solutions per instance the learning CachingSolver can hold (0 = automatic) */
    double  tol;         /* 0 -> 1e-10 (src/solvers.jl:175)                    */
    acmeb200_array a;    /* nx x nx        ACME.jl:119 */
    acmeb200_array b;    /* nx x nu        ACME.jl:120 */
    acmeb200_array c;    /* nx x nn_total  ACME.jl:121 */
    acmeb200_array x0;   /* nx             ACME.jl:122 */
    acmeb200_array dy;   /* ny x nx        ACME.jl:129 */
    acmeb200_array ey;   /* ny x nu        ACME.jl:130 */
    acmeb200_array fy;   /* ny x nn_total  ACME.jl:131 */
    acmeb200_array y0;   /* ny             ACME.jl:132 */
    const acmeb200_sub_desc *subs; /* nsub entries, in solve order            */
} acmeb200_model_desc;

typedef struct acmeb200_stats {
    uint64_t samples;                        /* instance-samples processed since reset      */
    uint64_t solves;                         /* sub-problem solves                          */
    uint64_t newton_iters;                   /* function evaluations (needediterations)     */
    uint64_t homotopy_solves;                /* solves that entered the homotopy loop       */
    uint64_t not_converged;                  /* solves that failed (finite)                 */
    uint64_t iter_hist[ACMEB200_HIST_BINS];  /* per-solve iteration histogram               */
} acmeb200_stats;

typedef struct acmeb200_model acmeb200_model; /* opaque; owns all device memory */

/* Devices: a model lives on the CUDA device that is current when it is created; hosts that do not link the CUDA
 * runtime themselves choose it here.  Several models on several devices may be driven by one process (one host thread
 * per model at a time). */
int acmeb200_device_count(int32_t *count_out);
int acmeb200_set_device(int32_t device);
int acmeb200_get_device(acmeb200_model *m, int32_t *device_out);

/* Create the device-resident model for instances [first_instance,
 * first_instance + n_instances) of the descriptor, on the current CUDA device.
 * Descriptor arrays are host pointers and are copied; nothing is retained.
 * Per-instance state (x = 0, extrapolation origin at (p = 0, init_z):
 * src/solvers.jl:164-178, ACME.jl:145) is initialised on the device. */
int acmeb200_model_create(const acmeb200_model_desc *desc, int64_t first_instance,
                          int64_t n_instances, acmeb200_model **out);
void acmeb200_model_destroy(acmeb200_model *m);

/* run!(runner, y, u) for every instance (src/ACME.jl:658-664).
 * U: (nu, N, B) column-major, i.e. instance b's block is the reference's nu x N
 * matrix at U + b*u_stride (u_stride == 0: one input shared by all instances).
 * Y: (ny, N, B), instance b at Y + b*y_stride (y_stride == 0 -> ny*N).
 * State (x, extrapolation origins) persists across calls (src/ACME.jl:561-562).
 * `stream` is a cudaStream_t (NULL = default stream).  With device pointers the
 * call is asynchronous on `stream`; with host pointers it stages through pinned
 * buffers and returns when Y is complete. */
int acmeb200_run(acmeb200_model *m, const double *U, int64_t u_stride, double *Y,
                 int64_t y_stride, int64_t n_samples, uint32_t flags, void *stream);

/* The batch over ALL the GPUs of the box from one call ("library drives all GPUs itself"):
 * instances never interact, so the n_instances of the descriptor are split into contiguous shards
 * [g*B/G, (g+1)*B/G) (the first B mod G shards one longer), one acmeb200_model per device.  acmeb200_multi_run is
 * acmeb200_run with HOST streams for the whole batch: every shard runs its own host-buffer pipeline from its own host
 * thread, concurrently; it returns when all of Y is complete.  n_gpus <= 0: every visible device.  Per-shard state
 * and statistics are reached through acmeb200_multi_model (shard g, its first instance and instance count). */
typedef struct acmeb200_multi acmeb200_multi;
int acmeb200_multi_create(const acmeb200_model_desc *desc, int64_t n_instances, int32_t n_gpus, acmeb200_multi **out);
void acmeb200_multi_destroy(acmeb200_multi *mm);
int acmeb200_multi_shards(const acmeb200_multi *mm);
acmeb200_model *acmeb200_multi_model(acmeb200_multi *mm, int32_t shard, int64_t *first_out, int64_t *count_out);
int acmeb200_multi_run(acmeb200_multi *mm, const double *U, int64_t u_stride, double *Y, int64_t y_stride,
                       int64_t n_samples, uint32_t flags);

/* model.x (nx x B, column-major) */
int acmeb200_get_state(acmeb200_model *m, double *x_host);
int acmeb200_set_state(acmeb200_model *m, const double *x_host, int64_t stride);
/* x = 0 and extrapolation origins back to (0, init_z); clears stats/status */
int acmeb200_reset(acmeb200_model *m);

/* The complete mutable state of the model -- what deepcopy(model) of the reference carries besides the matrices:
 * x (src/ACME.jl:145), every solver's extrapolation origin last_p / last_z / last_LU / last_Jp
 * (src/solvers.jl:155-158), the CachingSolver's stored solutions, tree and counters (src/solvers.jl:321-325),
 * status words, statistics and the running sample count -- as one opaque blob.  A blob restores into a model created
 * from the same descriptor and instance count with the same kernel selected (on any device, in any process):
 * create -> run(chunk 1) -> get -> destroy -> create -> set -> run(chunk 2) equals one run, bit for bit.
 * acmeb200_solver_state_size returns the bytes needed (negative: error). */
int64_t acmeb200_solver_state_size(acmeb200_model *m);
int acmeb200_get_solver_state(acmeb200_model *m, void *buf, int64_t bytes);
int acmeb200_set_solver_state(acmeb200_model *m, const void *buf, int64_t bytes);
/* get_extrapolation_origin(solver) (src/solvers.jl:199): origin (p, z) of sub-problem `sub` for every instance,
 * p_host np x B and z_host nn x B column-major; either may be NULL */
int acmeb200_get_extrapolation_origin(acmeb200_model *m, int32_t sub, double *p_host, double *z_host);

/* status words (one uint32 per instance) and the first failing sample index
 * (int64 per instance, -1 if none); either pointer may be NULL */
int acmeb200_get_status(acmeb200_model *m, uint32_t *status_host, int64_t *first_fail_host);
int acmeb200_get_stats(acmeb200_model *m, acmeb200_stats *out);
/* number of solutions the CachingSolver of sub-problem `sub` holds per instance
 * (num_ps of /root/reference/src/solvers.jl:321-323; 0 when the model has no store);
 * *capacity_out (may be NULL) receives the physical per-instance capacity: the reference's
 * arrays double for ever (solvers.jl:376-382), the device store stops accepting solutions
 * when it is full (flag ACMEB200_CACHE_FULL of acmeb200_get_cache_info) */
int acmeb200_get_cache_sizes(acmeb200_model *m, int32_t sub, int32_t *sizes_host, int32_t *capacity_out);

/* the CachingSolver bookkeeping of sub-problem `sub`, 8 int32 per instance: num_ps, new_count,
 * new_count_limit (solvers.jl:321-325), the capacity the reference's doubling arrays would have,
 * the number of points in the current tree, flags (ACMEB200_CACHE_*), 2 reserved */
#define ACMEB200_CACHE_FROZEN        1 /* host-built store of the descriptor, read-only          */
#define ACMEB200_CACHE_FULL          2 /* a solution could not be stored: capacity reached       */
#define ACMEB200_CACHE_HEAP_OVERFLOW 4 /* a search dropped an alternative (never seen in tests)  */
int acmeb200_get_cache_info(acmeb200_model *m, int32_t sub, int32_t *info_host);

/* Shapes compiled outside the library.  The thread-per-instance kernel is a template over the model's dimensions and
 * element sequence; the library instantiates it for the BASELINE circuits.  For any other small single-sub-problem
 * model the host generates the one-line instantiation from csrc/shape_plugin.cu.in, builds it with nvcc into its own
 * shared object (acme_jl_b200.specialise does both) and hands the plugin's entry (the pointer returned by
 * the plugin's exported function `shape_entry` with the acmeb200_ prefix) to the library; later models of that shape run on the specialised kernel. */
int acmeb200_register_tpi(const void *entry);

/* kernel selection: 0 = automatic, 1 = force the generic thread-per-instance
 * kernel, 2 = force the cooperative (lanes-per-instance) kernel, 3 = force the
 * warp-per-instance kernel with the LU rows in registers; solver state is reset */
int acmeb200_set_kernel(acmeb200_model *m, int32_t mode);
const char *acmeb200_kernel_name(const acmeb200_model *m);
/* number of kernels launched by this model since creation */
int64_t acmeb200_launch_count(const acmeb200_model *m);

/* KDTree(p, Np) (src/kdtree.jl:11-73), host side: builds the tree arrays of an acmeb200_cache over the first
 * `n_points` of `ps` (np x n_columns, column-major) exactly as the reference's constructor does -- including its
 * sort over ALL n_columns columns (kdtree.jl:37).  cut_dim / cut_val receive n_points-1 entries, ps_idx n_points
 * (1-based, like the reference).  This is the same code the device runs when a learning CachingSolver rebuilds its
 * tree (solvers.jl:390); a host that has collected solutions (ps, zs) uses it to freeze them into a cache. */
int acmeb200_kdtree_build(int32_t np, int32_t n_columns, int32_t n_points, const double *ps,
                          int32_t *cut_dim, double *cut_val, int32_t *ps_idx);
/* indnearest(tree, p, alts) (src/kdtree.jl:192-234) with alts initialised by init!(alts, best_dist, best_pidx):
 * the 1-based column of the tree point nearest to each of the n_queries columns of `queries` (np x n_queries) that
 * is strictly nearer than best_dist (INFINITY: plain nearest neighbour), else best_pidx.  Host side, same code as
 * the device search. */
int acmeb200_kdtree_indnearest(int32_t np, int32_t n_columns, int32_t n_points, const int32_t *cut_dim,
                               const double *cut_val, const int32_t *ps_idx, const double *ps,
                               int32_t n_queries, const double *queries, double best_dist, int32_t best_pidx,
                               int32_t *nearest_out);

/* Element Jacobians Jq = d(res)/dq of sub-problem `sub` for the whole batch (CircuitNLFunc, src/circuit.jl:10-17, as used
 * by linearize, src/ACME.jl:520-546 / get_extrapolation_jacobian, src/solvers.jl:407-414): q_host holds one q per instance,
 * [nq][count] with the instance index fastest; jq_host receives [nq][nn][count] (column-major nn x nq per instance,
 * instance index fastest).  Uses each instance's own element constants; the solver state is not touched. */
int acmeb200_eval_jq(acmeb200_model *m, int32_t sub, const double *q_host, double *jq_host);

/* diagnostic: measured FP64 FMA throughput of the current device in TFLOP/s
 * (2 flops per DFMA), the denominator of the FP64-pipe roofline in bench.py */
int acmeb200_measure_fp64_peak(double *tflops_out);

/* diagnostic: the device's exp (csrc/elements.cuh, the one the diode and BJT laws use) of n host values */
int acmeb200_diag_exp(const double *x_host, double *out_host, int64_t n);

const char *acmeb200_last_error(void);
int acmeb200_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ACMEB200_H */
