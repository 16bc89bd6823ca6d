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
#include "kernel_tpi.cuh"  // TMA / mbarrier helpers

namespace acme {

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
template <int L>
__device__ inline void c_set_p(const CCtx<L>& g, const DevSub& s, int prow) {
    for (int i = g.lane; i < s.nq; i += L) {
        double acc = g.mat(s.o_q0, s.nq, i, 0);
        for (int j = 0; j < s.np; j++) acc = fma(g.mat(s.o_pexp, s.nq, i, j), g.W(prow + j), acc);
        g.W(g.m.w_pfull + i) = acc;
    }
    g.sync();
}

// evaluate!  (ACME.jl:178-188): returns max|res| (NaN-propagating), J into rows Jrow
template <int L>
__device__ inline double c_evaluate(const CCtx<L>& g, const DevSub& s, int si, int zrow, int Jrow, bool& Jfinite) {
    const DevModel& m = g.m;
    for (int i = g.lane; i < s.nq; i += L) {
        double acc = g.W(m.w_pfull + i);
        for (int j = 0; j < s.nn; j++) acc = fma(g.mat(s.o_fq, s.nq, i, j), g.W(zrow + j), acc);
        g.W(m.w_q + i) = acc;
    }
    g.sync();
    double rmax = 0.0;
    bool bad = false, nan = false;
    for (int e = g.lane; e < s.nelem; e += L) {  // one element per lane
        const DevElem& el = m.elems[s.elem0 + e];
        double res[2], jv[4];
        elem_eval(el.kind, g.consts + el.c_off, &g.W(m.w_q + el.q_off), res, jv);
        const int nn = elem_nn(el.kind), nj = elem_nj(el.kind);
        for (int k = 0; k < nj; k++) {
            g.W(m.w_jv + el.j_off + k) = jv[k];
            if (!isfinite(jv[k])) bad = true;
        }
        for (int r = 0; r < nn; r++) {
            g.W(m.w_res + el.row + r) = res[r];
            const double a = fabs(res[r]);
            if (a != a) nan = true;
            if (a > rmax) rmax = a;
        }
    }
    g.sync();
    const int nn2 = s.nn * s.nn;
    for (int idx = g.lane; idx < nn2; idx += L) {  // J = Jq*fq, one entry per lane and pass
        const int r = idx % s.nn, c = idx / s.nn;
        const RowProg& rp = m.rows[s.zoff + r];
        double v = 0.0;
        for (int t = 0; t < rp.n; t++) {
            const double coef = rp.jv[t] >= 0 ? g.W(m.w_jv + rp.jv[t]) : (double)rp.c[t];
            v = fma(coef, g.mat(s.o_fq, s.nq, rp.q[t], c), v);
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
template <int L>
__device__ inline void c_calc_Jp(const CCtx<L>& g, const DevSub& s, int si, int Jprow) {
    const DevModel& m = g.m;
    const int n = s.nn * s.np;
    for (int idx = g.lane; idx < n; idx += L) {
        const int r = idx % s.nn, c = idx / s.nn;
        const RowProg& rp = m.rows[s.zoff + r];
        double v = 0.0;
        for (int t = 0; t < rp.n; t++) {
            const double coef = rp.jv[t] >= 0 ? g.W(m.w_jv + rp.jv[t]) : (double)rp.c[t];
            v = fma(coef, g.mat(s.o_pexp, s.nq, rp.q[t], c), v);
        }
        g.W(Jprow + idx) = v;
    }
    g.sync();
}

// setlhs!  (solvers.jl:46-96): rows over lanes, pivot search by group reduction
template <int L>
__device__ inline bool c_lu(const CCtx<L>& g, int n, int A, int piv) {
    for (int k = 0; k < n; k++) {
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
template <int L>
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
template <int L>
__device__ inline void c_set_origin(const CCtx<L>& g, const DevSub& s, int si, int prow, int zrow) {
    const int sel = (int)g.W(s.w_sel);
    bool Jfin;
    c_set_p<L>(g, s, prow);
    c_evaluate<L>(g, s, si, zrow, s.w_LU[sel], Jfin);
    c_lu<L>(g, s.nn, s.w_LU[sel], s.w_ipiv[sel]);
    c_calc_Jp<L>(g, s, si, s.w_lastJp);
    for (int i = g.lane; i < s.np; i += L) g.W(s.w_lastp + i) = g.W(prow + i);
    for (int i = g.lane; i < s.nn; i += L) g.W(s.w_lastz + i) = g.W(zrow + i);
    g.sync();
}

// solve(::SimpleSolver, p)  (solvers.jl:207-236)
template <int L>
__device__ inline GSolveResult c_simple_solve(const CCtx<L>& g, const DevSub& s, int si, int prow) {
    const DevModel& m = g.m;
    const int nn = s.nn, np = s.np;
    const int sel = (int)g.W(s.w_sel);
    c_set_p<L>(g, s, prow);
    for (int i = g.lane; i < nn; i += L) {
        double acc = 0.0;
        for (int j = 0; j < np; j++)
            acc = fma(g.W(s.w_lastJp + j * nn + i), g.W(prow + j) - g.W(s.w_lastp + j), acc);
        g.W(m.w_tmp + i) = acc;
    }
    c_lusolve<L>(g, nn, s.w_LU[sel], s.w_ipiv[sel], m.w_tmp);
    for (int i = g.lane; i < nn; i += L) g.W(m.w_z + i) = g.W(s.w_lastz + i) - g.W(m.w_tmp + i);
    g.sync();
    const int cur = 1 - sel;
    GSolveResult r{false, 0};
    for (r.iters = 1; r.iters <= m.maxiter; r.iters++) {
        bool Jfin;
        double resmax = c_evaluate<L>(g, s, si, m.w_z, s.w_LU[cur], Jfin);
        if (nn == 0) resmax = 0.0;
        if (!isfinite(resmax) || !Jfin) {
            r.converged = resmax < m.tol;
            return r;
        }
        if (!c_lu<L>(g, nn, s.w_LU[cur], s.w_ipiv[cur])) {
            r.converged = resmax < m.tol;
            return r;
        }
        if (resmax < m.tol) { r.converged = true; break; }
        for (int i = g.lane; i < nn; i += L) g.W(m.w_tmp + i) = g.W(m.w_res + i);
        c_lusolve<L>(g, nn, s.w_LU[cur], s.w_ipiv[cur], m.w_tmp);
        for (int i = g.lane; i < nn; i += L) g.W(m.w_z + i) -= g.W(m.w_tmp + i);
        g.sync();
    }
    if (r.iters > m.maxiter) r.iters = m.maxiter;
    if (r.converged) {
        c_calc_Jp<L>(g, s, si, s.w_lastJp);
        if (g.lane == 0) g.W(s.w_sel) = (double)cur;
        for (int i = g.lane; i < np; i += L) g.W(s.w_lastp + i) = g.W(prow + i);
        for (int i = g.lane; i < nn; i += L) g.W(s.w_lastz + i) = g.W(m.w_z + i);
        g.sync();
    }
    return r;
}

// solve(::CachingSolver, p)  (solvers.jl:347-396) with a DYNAMIC per-instance cache: the start
// point is the nearest of {current origin, every stored solution}; solutions that needed more
// than 5 iterations are appended (solvers.jl:374-386).  The reference finds the nearest stored
// point with a k-d tree plus a linear scan of the newest entries; here all stored points are
// scanned, the lanes of the group taking one point each -- the same exact nearest neighbour (up to
// distance ties), no tree to rebuild.  Capacity is fixed (dyn_cap); once full nothing is added.
template <int L>
__device__ inline GSolveResult c_base_solve(const CCtx<L>& g, const DevSub& s, int si, int prow, int64_t inst) {
    const DevModel& m = g.m;
    const bool caching = m.solver == ACMEB200_SOLVER_HOMOTOPY_CACHING && s.dyn_cap > 0;
    double* cps = nullptr;
    double* czs = nullptr;
    int n = 0;
    if (caching) {
        cps = s.dyn_ps + inst * (int64_t)s.np * s.dyn_cap;
        czs = s.dyn_zs + inst * (int64_t)s.nn * s.dyn_cap;
        n = s.dyn_n[inst];
        double best = 0.0;
        for (int i = 0; i < s.np; i++) {
            const double d = g.W(prow + i) - g.W(s.w_lastp + i);
            best = fma(d, d, best);
        }
        double lbest = best;
        int lidx = -1;
        for (int idx = g.lane; idx < n; idx += L) {
            double d2 = 0.0;
            for (int d = 0; d < s.np; d++) {
                const double df = cps[(int64_t)d * s.dyn_cap + idx] - g.W(prow + d);
                d2 = fma(df, df, d2);
            }
            if (d2 < lbest) { lbest = d2; lidx = idx; }
        }
#pragma unroll
        for (int o = L / 2; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(g.gmask, lbest, o, L);
            const int oi = __shfl_xor_sync(g.gmask, lidx, o, L);
            // ties: the current origin (-1) wins over stored points (the reference replaces it only on
            // a strictly smaller distance), among stored points the older one wins
            if (ob < lbest || (ob == lbest && (oi < 0 ? lidx >= 0 : (lidx >= 0 && oi < lidx)))) { lbest = ob; lidx = oi; }
        }
        if (lidx >= 0) {  // uniform within the group
            g.sync();
            for (int i = g.lane; i < s.np; i += L) g.W(m.w_cp + i) = cps[(int64_t)i * s.dyn_cap + lidx];
            for (int i = g.lane; i < s.nn; i += L) g.W(m.w_z + i) = czs[(int64_t)i * s.dyn_cap + lidx];
            g.sync();
            c_set_origin<L>(g, s, si, m.w_cp, m.w_z);
        }
    }
    const GSolveResult r = c_simple_solve<L>(g, s, si, prow);
    if (caching && r.iters > 5 && r.converged && n < s.dyn_cap) {
        for (int i = g.lane; i < s.np; i += L) cps[(int64_t)i * s.dyn_cap + n] = g.W(prow + i);
        for (int i = g.lane; i < s.nn; i += L) czs[(int64_t)i * s.dyn_cap + n] = g.W(m.w_z + i);
        g.sync();
        if (g.lane == 0) s.dyn_n[inst] = n + 1;
        __threadfence_block();
        g.sync();
    }
    return r;
}

// solve(::HomotopySolver, p)  (solvers.jl:268-296)
template <int L>
__device__ inline GSolveResult c_solve(const CCtx<L>& g, const DevSub& s, int si, bool& used_homotopy, int64_t inst) {
    const DevModel& m = g.m;
    GSolveResult r = c_base_solve<L>(g, s, si, m.w_p, inst);
    used_homotopy = false;
    if (m.solver == ACMEB200_SOLVER_SIMPLE || r.converged) return r;
    used_homotopy = true;
    int iters = r.iters;
    double a = 0.5, best_a = 0.0;
    g.sync();
    for (int i = g.lane; i < s.np; i += L) g.W(m.w_startp + i) = g.W(s.w_lastp + i);
    g.sync();
    while (best_a < 1) {
        for (int i = g.lane; i < s.np; i += L) {
            double pa = g.W(m.w_startp + i);
            pa *= (1 - a);
            pa += a * g.W(m.w_p + i);
            g.W(m.w_pa + i) = pa;
        }
        g.sync();
        r = c_base_solve<L>(g, s, si, m.w_pa, inst);
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

template <int L>
__global__ void __launch_bounds__(COOP_TPB) k_coop(const __grid_constant__ DevModel m, const RunArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* blob_s = reinterpret_cast<double*>(smem_raw + 16);
    const uint32_t bar = smem_u32(smem_raw);
    const size_t nblob = coop_blob_doubles(m), ngrp = coop_group_doubles(m);
    constexpr int GPC = COOP_TPB / L;
    const int grp = threadIdx.x / L, lane = threadIdx.x % L;
    double* w = blob_s + nblob + (size_t)grp * ngrp;
    double* consts_s = w + m.w_rows;
    unsigned int* hist_s = reinterpret_cast<unsigned int*>(w + ((m.w_rows + m.nconst + 1) & ~1));

    // ---- stage the shared model matrices with one TMA bulk copy
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
    for (int r = lane; r < m.w_rows; r += L) w[r] = a.ws[(int64_t)r * a.ld + inst];
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
            for (int k = lane; k < m.nu; k += L) g.W(m.w_u + k) = __ldg(u + n * m.nu + k);
            for (int k = lane; k < m.nnt; k += L) g.W(m.w_zall + k) = 0.0;
            g.sync();
            bool fatal = false;
            for (int si = 0; si < m.nsub; si++) {
                const DevSub& s = m.subs[si];
                for (int i = lane; i < s.np; i += L) {
                    double acc = 0.0;
                    for (int j = 0; j < m.nx; j++) acc = fma(g.mat(s.o_dq, s.np, i, j), g.W(m.w_x + j), acc);
                    for (int j = 0; j < m.nu; j++) acc = fma(g.mat(s.o_eq, s.np, i, j), g.W(m.w_u + j), acc);
                    if (si > 0)
                        for (int j = 0; j < m.nnt; j++) acc = fma(g.mat(s.o_fqprev, s.np, i, j), g.W(m.w_zall + j), acc);
                    g.W(m.w_p + i) = acc;
                }
                g.sync();
                bool used_h;
                const GSolveResult r = c_solve<L>(g, s, si, used_h, inst);
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
                    for (int i = 0; i < s.nn; i++) fin = fin && isfinite(g.W(m.w_z + i));
                    if (fin) {
                        status |= ACMEB200_STATUS_NOT_CONVERGED;
                        st_nc++;
                    } else {
                        status |= ACMEB200_STATUS_NONFINITE;
                        fatal = true;
                        break;
                    }
                }
                for (int i = lane; i < s.nn; i += L) g.W(m.w_zall + s.zoff + i) = g.W(m.w_z + i);
                g.sync();
            }
            if (fatal) break;
            for (int i = lane; i < m.ny; i += L) {
                double acc = g.mat(m.o_y0, m.ny, i, 0);
                for (int j = 0; j < m.nx; j++) acc = fma(g.mat(m.o_dy, m.ny, i, j), g.W(m.w_x + j), acc);
                for (int j = 0; j < m.nu; j++) acc = fma(g.mat(m.o_ey, m.ny, i, j), g.W(m.w_u + j), acc);
                for (int j = 0; j < m.nnt; j++) acc = fma(g.mat(m.o_fy, m.ny, i, j), g.W(m.w_zall + j), acc);
                y[n * m.ny + i] = acc;
            }
            for (int i = lane; i < m.nx; i += L) {
                double acc = g.mat(m.o_x0, m.nx, i, 0);
                for (int j = 0; j < m.nx; j++) acc = fma(g.mat(m.o_a, m.nx, i, j), g.W(m.w_x + j), acc);
                for (int j = 0; j < m.nu; j++) acc = fma(g.mat(m.o_b, m.nx, i, j), g.W(m.w_u + j), acc);
                for (int j = 0; j < m.nnt; j++) acc = fma(g.mat(m.o_c, m.nx, i, j), g.W(m.w_zall + j), acc);
                g.W(m.w_xnew + i) = acc;
            }
            g.sync();
            for (int i = lane; i < m.nx; i += L) g.W(m.w_x + i) = g.W(m.w_xnew + i);
            g.sync();
            st_samples++;
        }
    }
    if (active) {
        for (; n < a.N; n++)  // the reference throws here (ACME.jl:692); mark the rest
            for (int i = lane; i < m.ny; i += L) y[n * m.ny + i] = NAN;
        g.sync();
        for (int r = lane; r < m.w_rows; r += L) a.ws[(int64_t)r * a.ld + inst] = w[r];
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
