// Generic (runtime-dimension) kernel: one thread per instance, every vector and
// matrix of the instance's solver state lives in a global workspace laid out
// [row][instance] so that warp accesses coalesce.  It handles any model the ABI
// can describe (several sub-problems, per-instance matrices, frozen caches) and
// is the fallback for shapes that have no specialised instantiation.
//
// Follows step! (/root/reference/src/ACME.jl:666-715) and the solver stack
// (src/solvers.jl:207-236, 268-296, 347-373) statement by statement.
#pragma once
#include "devmodel.h"
#include "elements.cuh"
#include "kdcache.cuh"

namespace acme {

struct GCtx {
    const DevModel& m;
    const double* blob;
    const double* consts;  // already offset by the instance
    const double* initz;   // already offset by the instance
    double* ws;            // already offset by the instance
    int64_t ld;
    int64_t inst;          // instance index (per-instance solution cache)
    __device__ __forceinline__ double iz(int k) const { return initz[(int64_t)k * ld]; }
    __device__ __forceinline__ double& w(int row) const { return ws[(int64_t)row * ld]; }
    __device__ __forceinline__ double c(int k) const { return consts[(int64_t)k * ld]; }
    __device__ __forceinline__ double mat(int off, int ldm, int i, int j) const {
        return __ldg(blob + off + (int64_t)j * ldm + i);
    }
};

// (row r of element block) . m(k)
template <class F>
__device__ __forceinline__ double elem_row(int kind, int r, const double* jv, F m) {
    switch (kind) {
        case Diode::KIND: return Diode::row<0>(jv, m);
        case Bjt::KIND: return r == 0 ? Bjt::row<0>(jv, m) : Bjt::row<1>(jv, m);
        case Pot::KIND: return r == 0 ? Pot::row<0>(jv, m) : Pot::row<1>(jv, m);
        case Mosfet::KIND: return Mosfet::row<0>(jv, m);
        case OpampTanh::KIND: return OpampTanh::row<0>(jv, m);
        case JilesAtherton::KIND: return JilesAtherton::row<0>(jv, m);
        case TestQuad::KIND: return TestQuad::row<0>(jv, m);
    }
    return NAN;
}

__device__ __forceinline__ void elem_eval(int kind, const double* C, const double* q, double* res,
                                          double* jv) {
    switch (kind) {
        case Diode::KIND: Diode::eval(C, q, res, jv, ACME_EXPC); break;
        case Bjt::KIND: Bjt::eval(C, q, res, jv, ACME_EXPC); break;
        case Pot::KIND: Pot::eval(C, q, res, jv); break;
        case Mosfet::KIND: Mosfet::eval(C, q, res, jv); break;
        case OpampTanh::KIND: OpampTanh::eval(C, q, res, jv); break;
        case JilesAtherton::KIND: JilesAtherton::eval(C, q, res, jv); break;
        case TestQuad::KIND: TestQuad::eval(C, q, res, jv); break;
        default: res[0] = NAN;
    }
}

// set_p!: pfull = q0 + pexp*p   (ACME.jl:237-243); p is read from workspace row prow
__device__ inline void g_set_p(const GCtx& g, const DevSub& s, int prow) {
    for (int i = 0; i < s.nq; i++) {
        double acc = g.mat(s.o_q0, s.nq, i, 0);
        for (int j = 0; j < s.np; j++) acc = fma(g.mat(s.o_pexp, s.nq, i, j), g.w(prow + j), acc);
        g.w(g.m.w_pfull + i) = acc;
    }
}

// evaluate!: q = pfull + fq*z, element laws, J = Jq*fq  (ACME.jl:178-188, circuit.jl:10-17).
// J goes to workspace rows Jrow (nn x nn, column-major); returns max|res| (NaN-propagating)
// and whether J is finite.
__device__ inline double g_evaluate(const GCtx& g, const DevSub& s, int zrow, int Jrow, bool& Jfinite) {
    const DevModel& m = g.m;
    for (int i = 0; i < s.nq; i++) {
        double acc = g.w(m.w_pfull + i);
        for (int j = 0; j < s.nn; j++) acc = fma(g.mat(s.o_fq, s.nq, i, j), g.w(zrow + j), acc);
        g.w(m.w_q + i) = acc;
    }
    double resmax = 0.0;
    bool nan = false;
    Jfinite = true;
    for (int e = 0; e < s.nelem; e++) {
        const DevElem& el = m.elems[s.elem0 + e];
        double C[20], q[5], res[2], jv[4];
        const int nc = elem_nc(el.kind), nq = elem_nq(el.kind), nn = elem_nn(el.kind), nj = elem_nj(el.kind);
        for (int k = 0; k < nc; k++) C[k] = g.c(el.c_off + k);
        for (int k = 0; k < nq; k++) q[k] = g.w(m.w_q + el.q_off + k);
        elem_eval(el.kind, C, q, res, jv);
        for (int k = 0; k < nj; k++) g.w(m.w_jv + el.j_off + k) = jv[k];
        for (int r = 0; r < nn; r++) {
            g.w(m.w_res + el.row + r) = res[r];
            const double a = fabs(res[r]);
            if (a != a) nan = true;
            if (a > resmax) resmax = a;
            for (int c = 0; c < s.nn; c++) {
                const double v = elem_row(el.kind, r, jv, [&](int k) { return g.mat(s.o_fq, s.nq, el.q_off + k, c); });
                g.w(Jrow + c * s.nn + el.row + r) = v;
                if (!isfinite(v)) Jfinite = false;
            }
        }
    }
    return nan ? NAN : resmax;
}

// calc_Jp!: Jp = Jq*pexp into workspace rows Jprow  (ACME.jl:246-251)
__device__ inline void g_calc_Jp(const GCtx& g, const DevSub& s, int Jprow) {
    const DevModel& m = g.m;
    for (int e = 0; e < s.nelem; e++) {
        const DevElem& el = m.elems[s.elem0 + e];
        double jv[4];
        const int nn = elem_nn(el.kind), nj = elem_nj(el.kind);
        for (int k = 0; k < nj; k++) jv[k] = g.w(m.w_jv + el.j_off + k);
        for (int r = 0; r < nn; r++)
            for (int c = 0; c < s.np; c++)
                g.w(Jprow + c * s.nn + el.row + r) =
                    elem_row(el.kind, r, jv, [&](int k) { return g.mat(s.o_pexp, s.nq, el.q_off + k, c); });
    }
}

// setlhs!: in-place LU with partial pivoting, inverses on the diagonal (solvers.jl:46-96)
__device__ inline bool g_lu(const GCtx& g, int n, int A, int piv) {
    for (int k = 0; k < n; k++) {
        int kp = k;
        double amax = 0.0;
        for (int i = k; i < n; i++) {
            const double absi = fabs(g.w(A + k * n + i));
            if (absi > amax) { kp = i; amax = absi; }
        }
        g.w(piv + k) = (double)kp;
        if (g.w(A + k * n + kp) != 0.0) {
            if (k != kp)
                for (int i = 0; i < n; i++) {
                    const double t = g.w(A + i * n + k);
                    g.w(A + i * n + k) = g.w(A + i * n + kp);
                    g.w(A + i * n + kp) = t;
                }
            const double inv = 1.0 / g.w(A + k * n + k);
            g.w(A + k * n + k) = inv;
            for (int i = k + 1; i < n; i++) g.w(A + k * n + i) *= inv;
        } else {
            return false;
        }
        for (int j = k + 1; j < n; j++) {
            const double akj = g.w(A + j * n + k);
            for (int i = k + 1; i < n; i++)  // not fused, like the reference's scalar LU (exact zero pivots)
                g.w(A + j * n + i) = __dsub_rn(g.w(A + j * n + i), __dmul_rn(g.w(A + k * n + i), akj));
        }
    }
    return true;
}

// solve!: x (workspace rows xr) <- LU \ x   (solvers.jl:98-132)
__device__ inline void g_lusolve(const GCtx& g, int n, int A, int piv, int xr) {
    for (int i = 0; i < n; i++) {
        const int p = (int)g.w(piv + i);
        const double t = g.w(xr + i);
        g.w(xr + i) = g.w(xr + p);
        g.w(xr + p) = t;
    }
    for (int j = 0; j < n; j++) {
        const double xj = g.w(xr + j);
        for (int i = j + 1; i < n; i++) g.w(xr + i) = __dsub_rn(g.w(xr + i), __dmul_rn(g.w(A + j * n + i), xj));
    }
    for (int j = n - 1; j >= 0; j--) {
        const double xj = g.w(A + j * n + j) * g.w(xr + j);
        g.w(xr + j) = xj;
        for (int i = 0; i < j; i++) g.w(xr + i) = __dsub_rn(g.w(xr + i), __dmul_rn(g.w(A + j * n + i), xj));
    }
}

struct GSolveResult {
    bool converged;
    int iters;
};

// set_extrapolation_origin(solver, p, z): evaluate at (p, z), LU, Jp, store as origin
// (solvers.jl:183-196).  p in rows prow, z in rows zrow.
__device__ inline void g_set_origin(const GCtx& g, const DevSub& s, int prow, int zrow) {
    const int sel = (int)g.w(s.w_sel);
    bool Jfin;
    g_set_p(g, s, prow);
    g_evaluate(g, s, zrow, s.w_LU[sel], Jfin);
    g_lu(g, s.nn, s.w_LU[sel], s.w_ipiv[sel]);
    g_calc_Jp(g, s, s.w_lastJp);
    for (int i = 0; i < s.np; i++) g.w(s.w_lastp + i) = g.w(prow + i);
    for (int i = 0; i < s.nn; i++) g.w(s.w_lastz + i) = g.w(zrow + i);
}

// solve(::SimpleSolver, p)  (solvers.jl:207-236); p in rows prow, result z in rows m.w_z
__device__ inline GSolveResult g_simple_solve(const GCtx& g, const DevSub& s, int prow) {
    const DevModel& m = g.m;
    const int nn = s.nn, np = s.np;
    int sel = (int)g.w(s.w_sel);  // which LU buffer holds the origin's factors
    g_set_p(g, s, prow);
    // z = last_z - last_LU \ (last_Jp*(p - last_p))
    for (int i = 0; i < nn; i++) {
        double acc = 0.0;
        for (int j = 0; j < np; j++)
            acc = fma(g.w(s.w_lastJp + j * nn + i), g.w(prow + j) - g.w(s.w_lastp + j), acc);
        g.w(m.w_tmp + i) = acc;
    }
    g_lusolve(g, nn, s.w_LU[sel], s.w_ipiv[sel], m.w_tmp);
    for (int i = 0; i < nn; i++) g.w(m.w_z + i) = g.w(s.w_lastz + i) - g.w(m.w_tmp + i);

    const int cur = 1 - sel;  // Newton factorises into the other buffer
    GSolveResult r{false, 0};
    double resmax = 0.0;
    for (r.iters = 1; r.iters <= m.maxiter; r.iters++) {
        bool Jfin;
        resmax = g_evaluate(g, s, m.w_z, s.w_LU[cur], Jfin);
        if (nn == 0) resmax = 0.0;
        if (!isfinite(resmax) || !Jfin) {
            r.converged = resmax < m.tol;  // hasconverged() only looks at resmaxabs (solvers.jl:203)
            return r;
        }
        if (!g_lu(g, nn, s.w_LU[cur], s.w_ipiv[cur])) {
            r.converged = resmax < m.tol;  // hasconverged() only looks at resmaxabs
            return r;
        }
        if (resmax < m.tol) { r.converged = true; break; }
        for (int i = 0; i < nn; i++) g.w(m.w_tmp + i) = g.w(m.w_res + i);
        g_lusolve(g, nn, s.w_LU[cur], s.w_ipiv[cur], m.w_tmp);
        for (int i = 0; i < nn; i++) g.w(m.w_z + i) -= g.w(m.w_tmp + i);
    }
    if (r.iters > m.maxiter) r.iters = m.maxiter;
    if (r.converged) {
        g_calc_Jp(g, s, s.w_lastJp);
        g.w(s.w_sel) = (double)cur;  // the fresh factors become the origin's
        for (int i = 0; i < np; i++) g.w(s.w_lastp + i) = g.w(prow + i);
        for (int i = 0; i < nn; i++) g.w(s.w_lastz + i) = g.w(m.w_z + i);
    }
    return r;
}

// this instance's solution store of sub-problem s (kdcache.cuh); a frozen host-built store is shared (stride 0)
__device__ __forceinline__ KdStore g_store(const DevSub& s, int64_t inst) {
    return KdStore::at(s.kd_base + inst * s.kd_stride, s.kd_scr ? s.kd_scr + inst * s.kd_sstride : nullptr, s.np, s.nn, s.kd_cap);
}

// solve(::CachingSolver, p) (solvers.jl:347-396) on the device store: start from the nearest of {origin, newest stored
// solutions, k-d tree}, solve, store the solution if it was expensive, rebuild the tree on the reference's schedule
__device__ inline GSolveResult g_base_solve(const GCtx& g, const DevSub& s, int prow) {
    const DevModel& m = g.m;
    if (m.solver == ACMEB200_SOLVER_HOMOTOPY_CACHING) {
        double best = 0.0;
        for (int i = 0; i < s.np; i++) {
            const double d = g.w(prow + i) - g.w(s.w_lastp + i);
            best = kd_add(best, kd_mul(d, d));
        }
        if (s.kd_cap > 0) {
            KdStore c = g_store(s, g.inst);
            int ovf = 0;
            const int idx = kd_lookup(c, [&](int i) { return g.w(prow + i); }, best, &ovf);
            if (ovf) c.hdr[KD_H_FLAGS] |= KD_F_HEAP_OVERFLOW;
            if (idx != 0) {
                for (int i = 0; i < s.np; i++) g.w(m.w_cp + i) = c.P(i, idx);
                for (int i = 0; i < s.nn; i++) g.w(m.w_z + i) = c.Z(i, idx);
                g_set_origin(g, s, m.w_cp, m.w_z);
            }
            const GSolveResult r = g_simple_solve(g, s, prow);
            kd_after_solve(c, r.iters > 5 && r.converged, [&](int i) { return g.w(prow + i); }, [&](int i) { return g.w(m.w_z + i); });
            return r;
        }
        // no store (np beyond the search's delta vector): the cache holds only (p = 0, z = init_z)  (solvers.jl:327-333)
        double d0 = 0.0;
        for (int i = 0; i < s.np; i++) d0 = kd_add(d0, kd_mul(g.w(prow + i), g.w(prow + i)));
        if (d0 < best) {
            for (int i = 0; i < s.np; i++) g.w(m.w_cp + i) = 0.0;
            for (int i = 0; i < s.nn; i++) g.w(m.w_z + i) = g.iz(s.o_initz + i);
            g_set_origin(g, s, m.w_cp, m.w_z);
        }
    }
    return g_simple_solve(g, s, prow);
}

// solve(::HomotopySolver, p)  (solvers.jl:268-296); p in rows m.w_p, z left in rows m.w_z
__device__ inline GSolveResult g_solve(const GCtx& g, const DevSub& s, bool& used_homotopy) {
    const DevModel& m = g.m;
    GSolveResult r = g_base_solve(g, s, m.w_p);
    used_homotopy = false;
    if (m.solver == ACMEB200_SOLVER_SIMPLE || r.converged) return r;
    used_homotopy = true;
    int iters = r.iters;
    double a = 0.5, best_a = 0.0;
    for (int i = 0; i < s.np; i++) g.w(m.w_startp + i) = g.w(s.w_lastp + i);
    while (best_a < 1) {
        for (int i = 0; i < s.np; i++) {
            double pa = g.w(m.w_startp + i);
            pa *= (1 - a);
            pa += a * g.w(m.w_p + i);
            g.w(m.w_pa + i) = pa;
        }
        r = g_base_solve(g, s, m.w_pa);
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

struct LocalStats {
    unsigned long long samples = 0, solves = 0, iters = 0, homotopy = 0, notconv = 0;
};

#ifdef ACME_GENERIC_KERNEL_TU  // the kernel itself is compiled into one translation unit (acmeb200.cu)
__global__ void __launch_bounds__(128) k_generic(const __grid_constant__ DevModel m, const RunArgs a) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.ninst) return;
    const int64_t inst = a.inst0 + t;
    GCtx g{m, a.blob + inst * a.blob_stride, a.consts + inst, a.initz + inst, a.ws + inst, a.ld, inst};

    if (a.init) {
        // DiscreteModel ctor: x = 0 (ACME.jl:145); solver ctor: origin at (0, init_z) (solvers.jl:176)
        for (int i = 0; i < m.nx; i++) g.w(m.w_x + i) = 0.0;
        for (int si = 0; si < m.nsub; si++) {
            const DevSub& s = m.subs[si];
            g.w(s.w_sel) = 0.0;
            for (int i = 0; i < s.np; i++) g.w(m.w_cp + i) = 0.0;
            for (int i = 0; i < s.nn; i++) g.w(m.w_z + i) = g.iz(s.o_initz + i);
            g_set_origin(g, s, m.w_cp, m.w_z);
            if (s.kd_cap > 0 && !s.kd_frozen) {  // CachingSolver ctor: the store holds (0, init_z)  (solvers.jl:327-333); memory zeroed by the host
                KdStore c = g_store(s, inst);
                kd_init(c, [&](int i) { return g.iz(s.o_initz + i); });
            }
        }
        a.status[inst] = 0;
        a.first_fail[inst] = -1;
        return;
    }

    // streams: instance-major (instance pitch u_stride, sample pitch nu) or, with ACMEB200_SAMPLE_MAJOR, sample-major
    // (instance pitch nu, sample pitch u_stride; then consecutive threads read consecutive words); u_stride == 0 is
    // one shared nu x N input in both layouts
    const int64_t u_is = a.smaj && a.u_stride != 0 ? m.nu : a.u_stride, u_ss = a.smaj && a.u_stride != 0 ? a.u_stride : m.nu;
    const int64_t y_is = a.smaj ? m.ny : a.y_stride, y_ss = a.smaj ? a.y_stride : m.ny;
    const double* u = a.U + t * u_is;
    double* y = a.Y + t * y_is;
    LocalStats st;
    unsigned int hist_lo[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // bins 1..8; the rest go straight to global
    uint32_t status = a.status[inst];
    int64_t n = 0;
    if (!(status & ACMEB200_STATUS_NONFINITE)) {
        for (; n < a.N; n++) {
            // ---- step!  (ACME.jl:666-715)
            for (int k = 0; k < m.nu; k++) g.w(m.w_u + k) = __ldg(u + n * u_ss + k);
            for (int k = 0; k < m.nnt; k++) g.w(m.w_zall + k) = 0.0;
            bool fatal = false;
            for (int si = 0; si < m.nsub; si++) {
                const DevSub& s = m.subs[si];
                for (int i = 0; i < s.np; i++) {
                    double acc = 0.0;
                    for (int j = 0; j < m.nx; j++) acc = fma(g.mat(s.o_dq, s.np, i, j), g.w(m.w_x + j), acc);
                    for (int j = 0; j < m.nu; j++) acc = fma(g.mat(s.o_eq, s.np, i, j), g.w(m.w_u + j), acc);
                    if (si > 0)
                        for (int j = 0; j < m.nnt; j++)
                            acc = fma(g.mat(s.o_fqprev, s.np, i, j), g.w(m.w_zall + j), acc);
                    g.w(m.w_p + i) = acc;
                }
                bool used_h;
                const GSolveResult r = g_solve(g, s, used_h);
                st.solves++;
                st.iters += (unsigned)r.iters;
                st.homotopy += used_h ? 1u : 0u;
                {
                    int bin = r.iters < 1 ? 1 : r.iters;
                    if (bin > ACMEB200_HIST_BINS) bin = ACMEB200_HIST_BINS;
                    if (bin <= 8) hist_lo[bin - 1]++;
                    else atomicAdd(&a.stats->iter_hist[bin - 1], 1ull);
                }
                if (!r.converged) {
                    if (a.first_fail[inst] < 0) a.first_fail[inst] = a.n_done + n;
                    bool fin = true;
                    for (int i = 0; i < s.nn; i++) fin = fin && isfinite(g.w(m.w_z + i));
                    if (fin) {
                        status |= ACMEB200_STATUS_NOT_CONVERGED;
                        st.notconv++;
                    } else {
                        status |= ACMEB200_STATUS_NONFINITE;
                        fatal = true;
                        break;
                    }
                }
                for (int i = 0; i < s.nn; i++) g.w(m.w_zall + s.zoff + i) = g.w(m.w_z + i);
            }
            if (fatal) break;
            for (int i = 0; i < m.ny; i++) {
                double acc = g.mat(m.o_y0, m.ny, i, 0);
                for (int j = 0; j < m.nx; j++) acc = fma(g.mat(m.o_dy, m.ny, i, j), g.w(m.w_x + j), acc);
                for (int j = 0; j < m.nu; j++) acc = fma(g.mat(m.o_ey, m.ny, i, j), g.w(m.w_u + j), acc);
                for (int j = 0; j < m.nnt; j++) acc = fma(g.mat(m.o_fy, m.ny, i, j), g.w(m.w_zall + j), acc);
                y[n * y_ss + i] = acc;
            }
            for (int i = 0; i < m.nx; i++) {
                double acc = g.mat(m.o_x0, m.nx, i, 0);
                for (int j = 0; j < m.nx; j++) acc = fma(g.mat(m.o_a, m.nx, i, j), g.w(m.w_x + j), acc);
                for (int j = 0; j < m.nu; j++) acc = fma(g.mat(m.o_b, m.nx, i, j), g.w(m.w_u + j), acc);
                for (int j = 0; j < m.nnt; j++) acc = fma(g.mat(m.o_c, m.nx, i, j), g.w(m.w_zall + j), acc);
                g.w(m.w_xnew + i) = acc;
            }
            for (int i = 0; i < m.nx; i++) g.w(m.w_x + i) = g.w(m.w_xnew + i);
            st.samples++;
        }
    }
    for (; n < a.N; n++)  // the reference throws here (ACME.jl:692); mark the rest
        for (int i = 0; i < m.ny; i++) y[n * y_ss + i] = NAN;
    a.status[inst] = status;
    if (st.samples) atomicAdd(&a.stats->samples, st.samples);
    if (st.solves) atomicAdd(&a.stats->solves, st.solves);
    if (st.iters) atomicAdd(&a.stats->newton_iters, st.iters);
    if (st.homotopy) atomicAdd(&a.stats->homotopy_solves, st.homotopy);
    if (st.notconv) atomicAdd(&a.stats->not_converged, st.notconv);
    for (int b = 0; b < 8; b++)
        if (hist_lo[b]) atomicAdd(&a.stats->iter_hist[b], (unsigned long long)hist_lo[b]);
}

#endif  // ACME_GENERIC_KERNEL_TU

}  // namespace acme
