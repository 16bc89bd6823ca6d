// Device-side model description shared by the host code and all kernels.
#pragma once
#include <cstdint>
#include "../../include/acmeb200.h"

namespace acme {

constexpr int MAX_SUBS = 8;
constexpr int MAX_ELEMS = 48;
constexpr int MAX_ROWS = 32;  // nn per sub-problem supported by the cooperative kernel
constexpr int MAX_TOTAL_ROWS = 64;

// One row of the block-diagonal element Jacobian Jq as a short list of terms
// coef * e_q, coef either a variable entry jv[j] of the element evaluation or a
// constant (+-1).  Derived on the host by probing the elements' row<R>() functions, so
// the cooperative kernel forms J = Jq*fq and Jp = Jq*pexp without a per-entry switch
// over element kinds.
struct RowProg {
    unsigned char n;
    unsigned char q[5];   // q index within the sub
    signed char jv[5];    // index into the sub's jv scratch, or -1 for a constant
    float c[5];           // the constant when jv < 0
};

struct DevElem {
    int kind;
    int q_off;  // first q of this element within its sub (CircuitNLFunc order, circuit.jl:76)
    int c_off;  // first derived constant within the model's constant table
    int row;    // first residual row within its sub
    int j_off;  // first variable Jacobian entry within the sub's jv scratch
};

struct DevSub {
    int nn, nq, np, nelem, elem0, zoff, njv, o_initz;
    // offsets (in doubles) into the matrix blob, all column-major
    int o_dq, o_eq, o_fqprev, o_pexp, o_q0, o_fq;
    // generic-kernel workspace rows that persist across run calls
    int w_lastp, w_lastz, w_lastJp, w_LU[2], w_ipiv[2], w_sel;
    // solution store of the CachingSolver (kdcache.cuh: the reference's k-d tree + doubling arrays, solvers.jl:319-396,
    // kdtree.jl): one instance-contiguous block per instance (kd_stride doubles apart; 0 = one frozen, host-built
    // store shared by all instances, acmeb200_cache), scratch for the rebuilds, physical capacity in columns
    double* kd_base;
    int64_t kd_stride;
    double* kd_scr;
    int64_t kd_sstride;
    int kd_cap;     // 0: no store (solver without cache, or np beyond what the search holds)
    int kd_frozen;  // 1: the host-built store of the descriptor (read-only)
    // Leaf mirror of the specialised kernels: the points of the current tree in leaf order, laid out for the kernel's
    // lanes to scan (thread-per-instance: doubles, [(leaf*np + d)*kd_mld + instance]; warp-per-instance: see kernel_rows.cuh).
    // What the reference's tree search returns is the nearest tree point if it is strictly nearer than the seed; scanning
    // every leaf with all lanes finds it at streaming speed where the search is a chain of dependent loads -- the search
    // itself only runs when two leaves tie for the minimum (then its visiting order decides).
    double* kd_mir;
    int64_t kd_mld;   // leading dimension of the mirror (instances), 1 for the shared mirror of a frozen store
    int kd_mshared;   // 1: one mirror for all instances
};

struct DevModel {
    int nx, nu, ny, nsub, nnt, solver, maxiter, nelem_total;
    double tol;
    int o_a, o_b, o_c, o_x0, o_dy, o_ey, o_fy, o_y0, blob_len, nconst, ninitz;
    // generic-kernel workspace rows
    int w_x, w_u, w_zall, w_xnew, w_p, w_pfull, w_q, w_res, w_jv, w_z, w_tmp, w_startp, w_pa, w_cp, w_rows;
    DevSub subs[MAX_SUBS];
    DevElem elems[MAX_ELEMS];
    unsigned char row_elem[MAX_SUBS][MAX_ROWS];  // residual row -> element index within the sub
    RowProg rows[MAX_TOTAL_ROWS];                // sparse rows of Jq, indexed by sub.zoff + row
};

// device statistics block (mirrors acmeb200_stats, all 64-bit counters)
struct DevStats {
    unsigned long long samples, solves, newton_iters, homotopy_solves, not_converged;
    unsigned long long iter_hist[ACMEB200_HIST_BINS];
};

struct RunArgs {
    const double* blob;      // matrices: shared (blob_stride 0) or one blob per instance
    int64_t blob_stride;
    const double* consts;    // derived element constants, [nconst][ld]
    const double* initz;     // initial solutions, [ninitz][ld]
    double* ws;              // generic workspace / kernel-private state, [rows][ld]
    int64_t ld;              // instances in this model (leading dimension of the SoA arrays)
    const double* U;         // (nu, N, B) for the instances of this launch
    int64_t u_stride;
    double* Y;
    int64_t y_stride;
    int64_t N;               // samples in this call
    int64_t n_done;          // samples processed by earlier calls (for first_fail)
    int64_t inst0;           // first instance of this launch (index into the SoA arrays)
    int64_t ninst;           // instances in this launch
    uint32_t* status;
    long long* first_fail;
    DevStats* stats;
    int init;                // 1: (re)initialise the solver state, process no samples
    int smaj;                // 1: sample-major streams (ACMEB200_SAMPLE_MAJOR): u_stride / y_stride are sample pitches
};

}  // namespace acme
