// Cooperative kernel: a sub-warp group of L lanes owns one instance.
//
// For models whose non-linear system is too large for one thread's registers
// (superover: nn = 13, nq = 29, np = 11 -- a 13x13 LU per Newton iteration) and for
// batches too small to fill the GPU with one thread per instance (8 192 instances
// over 8 GPUs = 7 per SM).  Runtime dimensions, any number of sub-problems, every
// element kind.  Each group keeps its instance's complete solver state -- the same
// rows the generic kernel keeps in global memory (devmodel.h: w_*) -- in shared
// memory; rows of every matrix-vector product, the LU row updates and the element
// evaluations are spread over the lanes; pivot search and residual norms are
// warp-shuffle reductions inside the group.  The shared model matrices are staged
// into shared memory once per CTA by ONE TMA bulk copy (cp.async.bulk + mbarrier).
//
// Control flow inside a group is uniform by construction (every decision uses
// group-reduced values), so __syncwarp(group mask) is the only synchronisation.
// Follows step! (/root/reference/src/ACME.jl:666-715), solve(::SimpleSolver)
// (src/solvers.jl:207-236), LinearSolver (src/solvers.jl:46-132) and
// solve(::HomotopySolver) (src/solvers.jl:268-296) statement by statement, exactly
// like kernel_generic.cuh; the two share the persistent state layout.
#pragma once
#include "devmodel.h"
#include "elements.cuh"
#include "kernel_generic.cuh"
#include "tma.cuh"

namespace acme {

// Dimension policy: CoopDyn reads every dimension, blob offset and workspace row from the
// DevModel at run time; CoopStatic<...> fixes them at compile time (single sub-problem), which
// turns the index arithmetic of the runtime-dimension code into immediates and lets the loops
// unroll.  The static layout mirrors the host's packing order (acmeb200.cu) and is verified against
// the DevModel before such an instantiation is selected.
#define SD(stat, dyn) (P::STATIC ? (P::stat) : (dyn))
struct CoopDyn {
    static constexpr bool STATIC = false;
    static constexpr int ZERO = 0, ONE = 1, NX = 0, NU = 0, NY = 0, NN = 0, NQ = 0, NP = 0, NE = 0;
    static constexpr int O_A = 0, O_B = 0, O_C = 0, O_X0 = 0, O_DY = 0, O_EY = 0, O_FY = 0, O_Y0 = 0, O_DQ = 0, O_EQ = 0,
                         O_FQPREV = 0, O_PEXP = 0, O_Q0 = 0, O_FQ = 0, BLOB_LEN = 0;
    static constexpr int W_X = 0, W_U = 0, W_ZALL = 0, W_XNEW = 0, W_P = 0, W_PFULL = 0, W_Q = 0, W_RES = 0, W_JV = 0, W_Z = 0,
                         W_TMP = 0, W_STARTP = 0, W_PA = 0, W_CP = 0, W_LASTP = 0, W_LASTZ = 0, W_LASTJP = 0, W_LU0 = 0,
                         W_LU1 = 0, W_IPIV0 = 0, W_IPIV1 = 0, W_SEL = 0, W_ROWS = 0;
};
template <int NX_, int NU_, int NY_, int NN_, int NQ_, int NP_, int NE_, int NJV_>
struct CoopStatic {
    static constexpr bool STATIC = true;
    static constexpr int ZERO = 0, ONE = 1, NX = NX_, NU = NU_, NY = NY_, NN = NN_, NQ = NQ_, NP = NP_, NE = NE_, NJV = NJV_;
    static constexpr int O_A = 0, O_B = O_A + NX * NX, O_C = O_B + NX * NU, O_X0 = O_C + NX * NN, O_DY = O_X0 + NX,
                         O_EY = O_DY + NY * NX, O_FY = O_EY + NY * NU, O_Y0 = O_FY + NY * NN, O_DQ = O_Y0 + NY,
                         O_EQ = O_DQ + NP * NX, O_FQPREV = O_EQ + NP * NU, O_PEXP = O_FQPREV + NP * NN,
                         O_Q0 = O_PEXP + NQ * NP, O_FQ = O_Q0 + NQ, BLOB_LEN = O_FQ + NQ * NN;
    static constexpr int W_X = 0, W_U = W_X + NX, W_ZALL = W_U + NU, W_XNEW = W_ZALL + NN, W_P = W_XNEW + NX,
                         W_PFULL = W_P + NP, W_Q = W_PFULL + NQ, W_RES = W_Q + NQ, W_JV = W_RES + NN, W_Z = W_JV + NJV,
                         W_TMP = W_Z + NN, W_STARTP = W_TMP + NN, W_PA = W_STARTP + NP, W_CP = W_PA + NP,
                         W_LASTP = W_CP + NP, W_LASTZ = W_LASTP + NP, W_LASTJP = W_LASTZ + NN, W_LU0 = W_LASTJP + NN * NP,
                         W_LU1 = W_LU0 + NN * NN, W_IPIV0 = W_LU1 + NN * NN, W_IPIV1 = W_IPIV0 + NN, W_SEL = W_IPIV1 + NN,
                         W_ROWS = W_SEL + 1;
    // does a DevModel have exactly this shape and layout?
    static bool matches(const DevModel& m) {
        if (m.nsub != 1 || m.nx != NX || m.nu != NU || m.ny != NY) return false;
        const DevSub& s = m.subs[0];
        return s.nn == NN && s.nq == NQ && s.np == NP && s.nelem == NE && s.njv == NJV && s.zoff == 0 && s.elem0 == 0 &&
               s.o_initz == 0 && m.o_a == O_A && m.o_b == O_B && m.o_c == O_C && m.o_x0 == O_X0 && m.o_dy == O_DY &&
               m.o_ey == O_EY && m.o_fy == O_FY && m.o_y0 == O_Y0 && s.o_dq == O_DQ && s.o_eq == O_EQ &&
               s.o_fqprev == O_FQPREV && s.o_pexp == O_PEXP && s.o_q0 == O_Q0 && s.o_fq == O_FQ && m.blob_len == BLOB_LEN &&
               m.w_x == W_X && m.w_u == W_U && m.w_zall == W_ZALL && m.w_xnew == W_XNEW && m.w_p == W_P &&
               m.w_pfull == W_PFULL && m.w_q == W_Q && m.w_res == W_RES && m.w_jv == W_JV && m.w_z == W_Z &&
               m.w_tmp == W_TMP && m.w_startp == W_STARTP && m.w_pa == W_PA && m.w_cp == W_CP && s.w_lastp == W_LASTP &&
               s.w_lastz == W_LASTZ && s.w_lastJp == W_LASTJP && s.w_LU[0] == W_LU0 && s.w_LU[1] == W_LU1 &&
               s.w_ipiv[0] == W_IPIV0 && s.w_ipiv[1] == W_IPIV1 && s.w_sel == W_SEL && m.w_rows == W_ROWS;
    }
};

template <int L>
struct CCtx {
    const DevModel& m;
    const double* blob;    // shared matrices, in shared memory
    double* w;             // this group's rows, in shared memory
    const double* consts;  // this instance's element constants, in shared memory
    const double* initz;   // global, already offset by the instance
    int64_t ld;
    int lane;              // 0..L-1 within the group
    unsigned gmask;        // lanes of this group within the warp
    __device__ __forceinline__ double& W(int row) const { return w[row]; }
    __device__ __forceinline__ double mat(int off, int ldm, int i, int j) const { return blob[off + j * ldm + i]; }
    __device__ __forceinline__ void sync() const { __syncwarp(gmask); }
    __device__ __forceinline__ double iz(int k) const { return initz[(int64_t)k * ld]; }
};

template <int L>
__device__ __forceinline__ double group_max(const CCtx<L>& g, double v) {
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) {
        const double t = __shfl_xor_sync(g.gmask, v, o, L);
        v = t > v ? t : v;
    }
    return v;
}

// set_p!: pfull = q0 + pexp*p   (ACME.jl:237-243)
template <int L, class P>
__device__ inline void c_set_p(const CCtx<L>& g, const DevSub& s, int prow) {
    for (int i = g.lane; i < SD(NQ, s.nq); i += L) {
        double acc = g.mat(SD(O_Q0, s.o_q0), SD(NQ, s.nq), i, 0);
        for (int j = 0; j < SD(NP, s.np); j++) acc = fma(g.mat(SD(O_PEXP, s.o_pexp), SD(NQ, s.nq), i, j), g.W(prow + j), acc);
        g.W(SD(W_PFULL, g.m.w_pfull) + i) = acc;
    }
    g.sync();
}

// evaluate!  (ACME.jl:178-188): returns max|res| (NaN-propagating), J into rows Jrow
template <int L, class P>
__device__ inline double c_evaluate(const CCtx<L>& g, const DevSub& s, int si, int zrow, int Jrow, bool& Jfinite) {
    const DevModel& m = g.m;
    for (int i = g.lane; i < SD(NQ, s.nq); i += L) {
        double acc = g.W(SD(W_PFULL, m.w_pfull) + i);
        for (int j = 0; j < SD(NN, s.nn); j++) acc = fma(g.mat(SD(O_FQ, s.o_fq), SD(NQ, s.nq), i, j), g.W(zrow + j), acc);
        g.W(SD(W_Q, m.w_q) + i) = acc;
    }
    g.sync();
    double rmax = 0.0;
    bool bad = false, nan = false;
    for (int e = g.lane; e < SD(NE, s.nelem); e += L) {  // one element per lane
        const DevElem& el = m.elems[SD(ZERO, s.elem0) + e];
        double res[2], jv[4];
        elem_eval(el.kind, g.consts + el.c_off, &g.W(SD(W_Q, m.w_q) + el.q_off), res, jv);
        const int nn = elem_nn(el.kind), nj = elem_nj(el.kind);
        for (int k = 0; k < nj; k++) {
            g.W(SD(W_JV, m.w_jv) + el.j_off + k) = jv[k];
            if (!isfinite(jv[k])) bad = true;
        }
        for (int r = 0; r < nn; r++) {
            g.W(SD(W_RES, m.w_res) + el.row + r) = res[r];
            const double a = fabs(res[r]);
            if (a != a) nan = true;
            if (a > rmax) rmax = a;
        }
    }
    g.sync();
    const int nn2 = SD(NN, s.nn) * SD(NN, s.nn);
    for (int idx = g.lane; idx < nn2; idx += L) {  // J = Jq*fq, one entry per lane and pass
        const int r = idx % SD(NN, s.nn), c = idx / SD(NN, s.nn);
        const RowProg& rp = m.rows[SD(ZERO, s.zoff) + r];
        double v = 0.0;
        for (int t = 0; t < rp.n; t++) {
            const double coef = rp.jv[t] >= 0 ? g.W(SD(W_JV, m.w_jv) + rp.jv[t]) : (double)rp.c[t];
            v = fma(coef, g.mat(SD(O_FQ, s.o_fq), SD(NQ, s.nq), rp.q[t], c), v);
        }
        g.W(Jrow + idx) = v;
        if (!isfinite(v)) bad = true;
    }
    rmax = group_max<L>(g, rmax);
    const unsigned any_nan = __ballot_sync(g.gmask, nan) & g.gmask;
    const unsigned any_bad = __ballot_sync(g.gmask, bad) & g.gmask;
    Jfinite = any_bad == 0;
    g.sync();
    return any_nan ? NAN : rmax;
}

// calc_Jp!: Jp = Jq*pexp   (ACME.jl:246-251)
template <int L, class P>
__device__ inline void c_calc_Jp(const CCtx<L>& g, const DevSub& s, int si, int Jprow) {
    const DevModel& m = g.m;
    const int n = SD(NN, s.nn) * SD(NP, s.np);
    for (int idx = g.lane; idx < n; idx += L) {
        const int r = idx % SD(NN, s.nn), c = idx / SD(NN, s.nn);
        const RowProg& rp = m.rows[SD(ZERO, s.zoff) + r];
        double v = 0.0;
        for (int t = 0; t < rp.n; t++) {
            const double coef = rp.jv[t] >= 0 ? g.W(SD(W_JV, m.w_jv) + rp.jv[t]) : (double)rp.c[t];
            v = fma(coef, g.mat(SD(O_PEXP, s.o_pexp), SD(NQ, s.nq), rp.q[t], c), v);
        }
        g.W(Jprow + idx) = v;
    }
    g.sync();
}

// setlhs!  (solvers.jl:46-96): rows over lanes, pivot search by group reduction
template <int L, class P>
__device__ inline bool c_lu(const CCtx<L>& g, int n, int A, int piv) {
    for (int k = 0; k < n; k++) {  // not unrolled on purpose: 13x the code of three inlined call sites thrashes the I-cache (measured -15%)
        // first strict maximum of |A[i][k]|, i >= k
        double best = 0.0;
        int bi = k;
        for (int i = k + g.lane; i < n; i += L) {
            const double a = fabs(g.W(A + k * n + i));
            if (a > best) { best = a; bi = i; }
        }
#pragma unroll
        for (int o = L / 2; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(g.gmask, best, o, L);
            const int oi = __shfl_xor_sync(g.gmask, bi, o, L);
            if (ob > best || (ob == best && ob > 0.0 && oi < bi)) { best = ob; bi = oi; }
        }
        const int kp = best > 0.0 ? bi : k;
        if (g.lane == 0) g.W(piv + k) = (double)kp;
        if (g.W(A + k * n + kp) == 0.0) return false;  // uniform: every lane reads the same entry
        if (kp != k) {
            g.sync();
            for (int j = g.lane; j < n; j += L) {
                const double t = g.W(A + j * n + k);
                g.W(A + j * n + k) = g.W(A + j * n + kp);
                g.W(A + j * n + kp) = t;
            }
        }
        g.sync();
        const double inv = 1.0 / g.W(A + k * n + k);
        for (int i = k + 1 + g.lane; i < n; i += L) {
            const double l = g.W(A + k * n + i) * inv;
            g.W(A + k * n + i) = l;
            for (int j = k + 1; j < n; j++)  // not fused, like the reference's scalar LU (exact zero pivots)
                g.W(A + j * n + i) = __dsub_rn(g.W(A + j * n + i), __dmul_rn(l, g.W(A + j * n + k)));
        }
        g.sync();
        if (g.lane == 0) g.W(A + k * n + k) = inv;
    }
    g.sync();
    return true;
}

// solve!  (solvers.jl:98-132)
template <int L, class P>
__device__ inline void c_lusolve(const CCtx<L>& g, int n, int A, int piv, int xr) {
    g.sync();
    if (g.lane == 0)
        for (int i = 0; i < n; i++) {
            const int p = (int)g.W(piv + i);
            const double t = g.W(xr + i);
            g.W(xr + i) = g.W(xr + p);
            g.W(xr + p) = t;
        }
    for (int j = 0; j < n; j++) {
        g.sync();
        const double xj = g.W(xr + j);
        for (int i = j + 1 + g.lane; i < n; i += L)
            g.W(xr + i) = __dsub_rn(g.W(xr + i), __dmul_rn(g.W(A + j * n + i), xj));
    }
    for (int j = n - 1; j >= 0; j--) {
        g.sync();
        const double xj = g.W(A + j * n + j) * g.W(xr + j);
        g.sync();
        if (g.lane == 0) g.W(xr + j) = xj;
        for (int i = g.lane; i < j; i += L)
            g.W(xr + i) = __dsub_rn(g.W(xr + i), __dmul_rn(g.W(A + j * n + i), xj));
    }
    g.sync();
}

// set_extrapolation_origin(solver, p, z)  (solvers.jl:183-196)
template <int L, class P>
__device__ inline void c_set_origin(const CCtx<L>& g, const DevSub& s, int si, int prow, int zrow) {
    const int sel = (int)g.W(SD(W_SEL, s.w_sel));
    bool Jfin;
    c_set_p<L, P>(g, s, prow);
    c_evaluate<L, P>(g, s, si, zrow, (P::STATIC ? ((sel) ? P::W_LU1 : P::W_LU0) : s.w_LU[sel]), Jfin);
    c_lu<L, P>(g, SD(NN, s.nn), (P::STATIC ? ((sel) ? P::W_LU1 : P::W_LU0) : s.w_LU[sel]), (P::STATIC ? ((sel) ? P::W_IPIV1 : P::W_IPIV0) : s.w_ipiv[sel]));
    c_calc_Jp<L, P>(g, s, si, SD(W_LASTJP, s.w_lastJp));
    for (int i = g.lane; i < SD(NP, s.np); i += L) g.W(SD(W_LASTP, s.w_lastp) + i) = g.W(prow + i);
    for (int i = g.lane; i < SD(NN, s.nn); i += L) g.W(SD(W_LASTZ, s.w_lastz) + i) = g.W(zrow + i);
    g.sync();
}

// solve(::SimpleSolver, p)  (solvers.jl:207-236)
template <int L, class P>
__device__ inline GSolveResult c_simple_solve(const CCtx<L>& g, const DevSub& s, int si, int prow) {
    const DevModel& m = g.m;
    const int nn = SD(NN, s.nn), np = SD(NP, s.np);
    const int sel = (int)g.W(SD(W_SEL, s.w_sel));
    c_set_p<L, P>(g, s, prow);
    for (int i = g.lane; i < nn; i += L) {
        double acc = 0.0;
        for (int j = 0; j < np; j++)
            acc = fma(g.W(SD(W_LASTJP, s.w_lastJp) + j * nn + i), g.W(prow + j) - g.W(SD(W_LASTP, s.w_lastp) + j), acc);
        g.W(SD(W_TMP, m.w_tmp) + i) = acc;
    }
    c_lusolve<L, P>(g, nn, (P::STATIC ? ((sel) ? P::W_LU1 : P::W_LU0) : s.w_LU[sel]), (P::STATIC ? ((sel) ? P::W_IPIV1 : P::W_IPIV0) : s.w_ipiv[sel]), SD(W_TMP, m.w_tmp));
    for (int i = g.lane; i < nn; i += L) g.W(SD(W_Z, m.w_z) + i) = g.W(SD(W_LASTZ, s.w_lastz) + i) - g.W(SD(W_TMP, m.w_tmp) + i);
    g.sync();
    const int cur = 1 - sel;
    GSolveResult r{false, 0};
    for (r.iters = 1; r.iters <= m.maxiter; r.iters++) {
        bool Jfin;
        double resmax = c_evaluate<L, P>(g, s, si, SD(W_Z, m.w_z), (P::STATIC ? ((cur) ? P::W_LU1 : P::W_LU0) : s.w_LU[cur]), Jfin);
        if (nn == 0) resmax = 0.0;
        if (!isfinite(resmax) || !Jfin) {
            r.converged = resmax < m.tol;
            return r;
        }
        if (!c_lu<L, P>(g, nn, (P::STATIC ? ((cur) ? P::W_LU1 : P::W_LU0) : s.w_LU[cur]), (P::STATIC ? ((cur) ? P::W_IPIV1 : P::W_IPIV0) : s.w_ipiv[cur]))) {
            r.converged = resmax < m.tol;
            return r;
        }
        if (resmax < m.tol) { r.converged = true; break; }
        for (int i = g.lane; i < nn; i += L) g.W(SD(W_TMP, m.w_tmp) + i) = g.W(SD(W_RES, m.w_res) + i);
        c_lusolve<L, P>(g, nn, (P::STATIC ? ((cur) ? P::W_LU1 : P::W_LU0) : s.w_LU[cur]), (P::STATIC ? ((cur) ? P::W_IPIV1 : P::W_IPIV0) : s.w_ipiv[cur]), SD(W_TMP, m.w_tmp));
        for (int i = g.lane; i < nn; i += L) g.W(SD(W_Z, m.w_z) + i) -= g.W(SD(W_TMP, m.w_tmp) + i);
        g.sync();
    }
    if (r.iters > m.maxiter) r.iters = m.maxiter;
    if (r.converged) {
        c_calc_Jp<L, P>(g, s, si, SD(W_LASTJP, s.w_lastJp));
        if (g.lane == 0) g.W(SD(W_SEL, s.w_sel)) = (double)cur;
        for (int i = g.lane; i < np; i += L) g.W(SD(W_LASTP, s.w_lastp) + i) = g.W(prow + i);
        for (int i = g.lane; i < nn; i += L) g.W(SD(W_LASTZ, s.w_lastz) + i) = g.W(SD(W_Z, m.w_z) + i);
        g.sync();
    }
    return r;
}

// solve(::CachingSolver, p)  (solvers.jl:347-396) on the instance's solution store (kdcache.cuh: the reference's k-d
// tree, newest-entries scan and rebuild schedule).  The search, the store and the rebuilds are serial code run by the
// group's first lane (this kernel is the fallback for shapes without a specialised instantiation); the chosen column
// is broadcast and the group re-origins / solves together.
template <int L, class P>
__device__ inline GSolveResult c_base_solve(const CCtx<L>& g, const DevSub& s, int si, int prow, int64_t inst) {
    const DevModel& m = g.m;
    const bool caching = m.solver == ACMEB200_SOLVER_HOMOTOPY_CACHING && s.kd_cap > 0;
    if (caching) {
        int idx = 0;
        if (g.lane == 0) {
            double best = 0.0;
            for (int i = 0; i < SD(NP, s.np); i++) {
                const double d = g.W(prow + i) - g.W(SD(W_LASTP, s.w_lastp) + i);
                best = kd_add(best, kd_mul(d, d));
            }
            const KdStore c = g_store(s, inst);
            int ovf = 0;
            idx = kd_lookup(c, [&](int i) { return g.W(prow + i); }, best, &ovf);
            if (ovf) c.hdr[KD_H_FLAGS] |= KD_F_HEAP_OVERFLOW;
        }
        idx = __shfl_sync(g.gmask, idx, 0, L);
        if (idx != 0) {  // uniform within the group
            const KdStore c = g_store(s, inst);
            g.sync();
            for (int i = g.lane; i < SD(NP, s.np); i += L) g.W(SD(W_CP, m.w_cp) + i) = c.P(i, idx);
            for (int i = g.lane; i < SD(NN, s.nn); i += L) g.W(SD(W_Z, m.w_z) + i) = c.Z(i, idx);
            g.sync();
            c_set_origin<L, P>(g, s, si, SD(W_CP, m.w_cp), SD(W_Z, m.w_z));
        }
    }
    const GSolveResult r = c_simple_solve<L, P>(g, s, si, prow);
    if (caching && !s.kd_frozen) {  // solvers.jl:374-394
        g.sync();
        if (g.lane == 0) {
            KdStore c = g_store(s, inst);
            kd_after_solve(c, r.iters > 5 && r.converged, [&](int i) { return g.W(prow + i); }, [&](int i) { return g.W(SD(W_Z, m.w_z) + i); });
        }
        __threadfence_block();
        g.sync();
    }
    return r;
}

// solve(::HomotopySolver, p)  (solvers.jl:268-296)
template <int L, class P>
__device__ inline GSolveResult c_solve(const CCtx<L>& g, const DevSub& s, int si, bool& used_homotopy, int64_t inst) {
    const DevModel& m = g.m;
    GSolveResult r = c_base_solve<L, P>(g, s, si, SD(W_P, m.w_p), inst);
    used_homotopy = false;
    if (m.solver == ACMEB200_SOLVER_SIMPLE || r.converged) return r;
    used_homotopy = true;
    int iters = r.iters;
    double a = 0.5, best_a = 0.0;
    g.sync();
    for (int i = g.lane; i < SD(NP, s.np); i += L) g.W(SD(W_STARTP, m.w_startp) + i) = g.W(SD(W_LASTP, s.w_lastp) + i);
    g.sync();
    while (best_a < 1) {
        for (int i = g.lane; i < SD(NP, s.np); i += L) {
            double pa = g.W(SD(W_STARTP, m.w_startp) + i);
            pa *= (1 - a);
            pa += a * g.W(SD(W_P, m.w_p) + i);
            g.W(SD(W_PA, m.w_pa) + i) = pa;
        }
        g.sync();
        r = c_base_solve<L, P>(g, s, si, SD(W_PA, m.w_pa), inst);
        iters += r.iters;
        if (r.converged) {
            best_a = a;
            a = 1.0;
        } else {
            const double new_a = (a + best_a) / 2;
            if (!(best_a < new_a && new_a < a)) break;
            a = new_a;
        }
    }
    r.iters = iters;
    return r;
}

constexpr int COOP_TPB = 128;

// shared memory per CTA: [mbarrier 16 B][blob][groups x (w_rows + nconst) doubles]
__host__ __device__ inline size_t coop_blob_doubles(const DevModel& m) { return (size_t)((m.blob_len + 1) & ~1); }
__host__ __device__ inline size_t coop_group_doubles(const DevModel& m) {
    return (size_t)((m.w_rows + m.nconst + 1) & ~1) + ACMEB200_HIST_BINS / 2;  // + 32 uint32 histogram bins
}
template <int L>
__host__ __device__ inline size_t coop_smem_bytes(const DevModel& m) {
    return 16 + 8 * (coop_blob_doubles(m) + (COOP_TPB / L) * coop_group_doubles(m));
}

template <int L, class P>
__global__ void __launch_bounds__(COOP_TPB) k_coop(const __grid_constant__ DevModel m, const RunArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* blob_s = reinterpret_cast<double*>(smem_raw + 16);
    const uint32_t bar = smem_u32(smem_raw);
    const size_t nblob = coop_blob_doubles(m), ngrp = coop_group_doubles(m);
    constexpr int GPC = COOP_TPB / L;
    const int grp = threadIdx.x / L, lane = threadIdx.x % L;
    double* w = blob_s + nblob + (size_t)grp * ngrp;
    double* consts_s = w + SD(W_ROWS, m.w_rows);
    unsigned int* hist_s = reinterpret_cast<unsigned int*>(w + ((SD(W_ROWS, m.w_rows) + m.nconst + 1) & ~1));

    // ---- stage the shared model matrices with one TMA bulk copy
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init_fence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)(nblob * 8);
        mbar_arrive_expect_tx(bar, bytes);
        bulk_g2s(smem_u32(blob_s), a.blob, bytes, bar);
    }
    const int64_t t = (int64_t)blockIdx.x * GPC + grp;  // launch-local instance of this group
    const bool active = t < a.ninst;
    const int64_t inst = a.inst0 + (active ? t : 0);
    const unsigned gmask = (L == 32) ? 0xffffffffu : (((1u << L) - 1u) << ((threadIdx.x & 31) / L * L));
    // per-instance constants and persistent state rows -> shared memory
    for (int k = lane; k < m.nconst; k += L) consts_s[k] = a.consts[(int64_t)k * a.ld + inst];
    for (int r = lane; r < SD(W_ROWS, m.w_rows); r += L) w[r] = a.ws[(int64_t)r * a.ld + inst];
    for (int b = lane; b < ACMEB200_HIST_BINS; b += L) hist_s[b] = 0u;
    mbar_wait(bar, 0);
    __syncwarp(gmask);

    CCtx<L> g{m, blob_s, w, consts_s, a.initz + inst, a.ld, lane, gmask};
    const double* u = a.U + t * a.u_stride;
    double* y = a.Y + t * a.y_stride;
    unsigned long long st_samples = 0, st_solves = 0, st_iters = 0, st_hom = 0, st_nc = 0;
    uint32_t status = active ? a.status[inst] : ACMEB200_STATUS_NONFINITE;
    int64_t n = 0;
    if (!(status & ACMEB200_STATUS_NONFINITE)) {
        for (; n < a.N; n++) {
            // ---- step!  (ACME.jl:666-715)
            for (int k = lane; k < SD(NU, m.nu); k += L) g.W(SD(W_U, m.w_u) + k) = __ldg(u + n * SD(NU, m.nu) + k);
            for (int k = lane; k < SD(NN, m.nnt); k += L) g.W(SD(W_ZALL, m.w_zall) + k) = 0.0;
            g.sync();
            bool fatal = false;
            for (int si = 0; si < SD(ONE, m.nsub); si++) {
                const DevSub& s = m.subs[si];
                for (int i = lane; i < SD(NP, s.np); i += L) {
                    double acc = 0.0;
                    for (int j = 0; j < SD(NX, m.nx); j++) acc = fma(g.mat(SD(O_DQ, s.o_dq), SD(NP, s.np), i, j), g.W(SD(W_X, m.w_x) + j), acc);
                    for (int j = 0; j < SD(NU, m.nu); j++) acc = fma(g.mat(SD(O_EQ, s.o_eq), SD(NP, s.np), i, j), g.W(SD(W_U, m.w_u) + j), acc);
                    if (si > 0)
                        for (int j = 0; j < SD(NN, m.nnt); j++) acc = fma(g.mat(SD(O_FQPREV, s.o_fqprev), SD(NP, s.np), i, j), g.W(SD(W_ZALL, m.w_zall) + j), acc);
                    g.W(SD(W_P, m.w_p) + i) = acc;
                }
                g.sync();
                bool used_h;
                const GSolveResult r = c_solve<L, P>(g, s, si, used_h, inst);
                if (lane == 0) {
                    st_solves++;
                    st_iters += (unsigned)r.iters;
                    st_hom += used_h ? 1u : 0u;
                    int bin = r.iters < 1 ? 1 : r.iters;
                    if (bin > ACMEB200_HIST_BINS) bin = ACMEB200_HIST_BINS;
                    hist_s[bin - 1] += 1u;
                }
                if (!r.converged) {
                    if (lane == 0 && a.first_fail[inst] < 0) a.first_fail[inst] = a.n_done + n;
                    bool fin = true;
                    for (int i = 0; i < SD(NN, s.nn); i++) fin = fin && isfinite(g.W(SD(W_Z, m.w_z) + i));
                    if (fin) {
                        status |= ACMEB200_STATUS_NOT_CONVERGED;
                        st_nc++;
                    } else {
                        status |= ACMEB200_STATUS_NONFINITE;
                        fatal = true;
                        break;
                    }
                }
                for (int i = lane; i < SD(NN, s.nn); i += L) g.W(SD(W_ZALL, m.w_zall) + SD(ZERO, s.zoff) + i) = g.W(SD(W_Z, m.w_z) + i);
                g.sync();
            }
            if (fatal) break;
            for (int i = lane; i < SD(NY, m.ny); i += L) {
                double acc = g.mat(SD(O_Y0, m.o_y0), SD(NY, m.ny), i, 0);
                for (int j = 0; j < SD(NX, m.nx); j++) acc = fma(g.mat(SD(O_DY, m.o_dy), SD(NY, m.ny), i, j), g.W(SD(W_X, m.w_x) + j), acc);
                for (int j = 0; j < SD(NU, m.nu); j++) acc = fma(g.mat(SD(O_EY, m.o_ey), SD(NY, m.ny), i, j), g.W(SD(W_U, m.w_u) + j), acc);
                for (int j = 0; j < SD(NN, m.nnt); j++) acc = fma(g.mat(SD(O_FY, m.o_fy), SD(NY, m.ny), i, j), g.W(SD(W_ZALL, m.w_zall) + j), acc);
                y[n * SD(NY, m.ny) + i] = acc;
            }
            for (int i = lane; i < SD(NX, m.nx); i += L) {
                double acc = g.mat(SD(O_X0, m.o_x0), SD(NX, m.nx), i, 0);
                for (int j = 0; j < SD(NX, m.nx); j++) acc = fma(g.mat(SD(O_A, m.o_a), SD(NX, m.nx), i, j), g.W(SD(W_X, m.w_x) + j), acc);
                for (int j = 0; j < SD(NU, m.nu); j++) acc = fma(g.mat(SD(O_B, m.o_b), SD(NX, m.nx), i, j), g.W(SD(W_U, m.w_u) + j), acc);
                for (int j = 0; j < SD(NN, m.nnt); j++) acc = fma(g.mat(SD(O_C, m.o_c), SD(NX, m.nx), i, j), g.W(SD(W_ZALL, m.w_zall) + j), acc);
                g.W(SD(W_XNEW, m.w_xnew) + i) = acc;
            }
            g.sync();
            for (int i = lane; i < SD(NX, m.nx); i += L) g.W(SD(W_X, m.w_x) + i) = g.W(SD(W_XNEW, m.w_xnew) + i);
            g.sync();
            st_samples++;
        }
    }
    if (active) {
        for (; n < a.N; n++)  // the reference throws here (ACME.jl:692); mark the rest
            for (int i = lane; i < SD(NY, m.ny); i += L) y[n * SD(NY, m.ny) + i] = NAN;
        g.sync();
        for (int r = lane; r < SD(W_ROWS, m.w_rows); r += L) a.ws[(int64_t)r * a.ld + inst] = w[r];
        if (lane == 0) {
            a.status[inst] = status;
            if (st_samples) atomicAdd(&a.stats->samples, st_samples);
            if (st_solves) atomicAdd(&a.stats->solves, st_solves);
            if (st_iters) atomicAdd(&a.stats->newton_iters, st_iters);
            if (st_hom) atomicAdd(&a.stats->homotopy_solves, st_hom);
            if (st_nc) atomicAdd(&a.stats->not_converged, st_nc);
            for (int b = 0; b < ACMEB200_HIST_BINS; b++)
                if (hist_s[b]) atomicAdd(&a.stats->iter_hist[b], (unsigned long long)hist_s[b]);
        }
    }
}

}  // namespace acme
