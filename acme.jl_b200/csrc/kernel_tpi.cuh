// Thread-per-instance kernels with compile-time dimensions and element sequence.
//
// One thread owns one model instance for the whole call: its state vector x,
// the extrapolation origin (last_p, last_z) and the extrapolation matrix
// Mx = J^-1 * Jp live in registers; the shared model matrices arrive as a
// __grid_constant__ kernel parameter, so with fully unrolled loops every matrix
// element is a constant-bank operand of a DFMA.  The (nu, N, B) / (ny, N, B)
// streams are instance-slowest (each instance is a contiguous reference-layout
// block), so each warp stages [32 instances x T samples] tiles through shared
// memory with coalesced 128-byte row segments.
//
// Per sample this is step! (/root/reference/src/ACME.jl:666-715) with the
// SimpleSolver Newton loop (src/solvers.jl:207-236) inlined; the homotopy
// fallback (solvers.jl:268-296) and the cache lookup (solvers.jl:347-371) are
// cold, out-of-line paths working on a spilled copy of the state.
//
// Deviation from the reference's data layout (not its arithmetic): instead of
// keeping (last_LU, last_Jp) and solving last_LU \ (last_Jp*dp) every sample
// (solvers.jl:209-215) the kernel stores Mx = last_LU \ last_Jp once per
// converged solve; the start vector z0 = last_z - Mx*dp is the same up to
// rounding.
#pragma once
#include "devmodel.h"
#include "tma.cuh"
#include "elements.cuh"
#include "kernel_generic.cuh"  // shares elements.cuh / kdcache.cuh
#include "kdcache_warp.cuh"    // warp-cooperative tree rebuild

namespace acme {

// ACME_TPI_PROF (tuning builds only, tools/build_variants.py): per-warp cycle counts and event counters of the
// nonlinear sample loop in a device array, read back with acmeb200_diag_tpi_prof (tools/tail_prof.py).
#ifdef ACME_TPI_PROF
#define TPI_PROF(...) __VA_ARGS__
constexpr int TPI_PROF_WARPS = 8192, TPI_PROF_REC = 8;
static __device__ unsigned long long g_tpi_prof[TPI_PROF_WARPS * TPI_PROF_REC];  // one per translation unit: profile with ACMEB200_TPI_WIDE=0
struct TpiProf { unsigned reo, tie, scan, stores, nontriv, small; };
#else
#define TPI_PROF(...)
#endif

template <class... Es> struct EList {};

template <int ROW, int QOFF, int COFF, int JOFF, class F>
__device__ __forceinline__ void for_each_elem(EList<>, F&&) {}
template <int ROW, int QOFF, int COFF, int JOFF, class F, class E, class... R>
__device__ __forceinline__ void for_each_elem(EList<E, R...>, F&& f) {
    f(E{}, IC<ROW>{}, IC<QOFF>{}, IC<COFF>{}, IC<JOFF>{});
    for_each_elem<ROW + E::NN, QOFF + E::NQ, COFF + E::NC, JOFF + E::NJ>(EList<R...>{}, f);
}

__host__ __device__ constexpr int dim1(int n) { return n > 0 ? n : 1; }

template <int NX_, int NU_, int NY_, int NP_, class... Es>
struct TpiCfg {
    static constexpr int NX = NX_, NU = NU_, NY = NY_, NP = NP_;
    static constexpr int NE = sizeof...(Es);
    static constexpr int NN = (Es::NN + ... + 0);
    static constexpr int NQ = (Es::NQ + ... + 0);
    static constexpr int NC = (Es::NC + ... + 0);
    static constexpr int NJ = (Es::NJ + ... + 0);
    using Elems = EList<Es...>;
    static constexpr int kinds[dim1(NE)] = {Es::KIND...};
    // rows of the kernel-private state in RunArgs::ws
    static constexpr int S_X = 0, S_LP = NX, S_LZ = NX + NP, S_MX = NX + NP + NN, S_ROWS = NX + NP + NN + NN * NP;
};

// shared model matrices, column-major (field for field ACME.jl:119-132)
template <class C>
struct TpiMats {
    double a[dim1(C::NX * C::NX)], b[dim1(C::NX * C::NU)], c[dim1(C::NX * C::NN)], x0[dim1(C::NX)];
    double dq[dim1(C::NP * C::NX)], eq[dim1(C::NP * C::NU)], pexp[dim1(C::NQ * C::NP)], q0[dim1(C::NQ)],
        fq[dim1(C::NQ * C::NN)];
    double dy[dim1(C::NY * C::NX)], ey[dim1(C::NY * C::NU)], fy[dim1(C::NY * C::NN)], y0[dim1(C::NY)];
};

struct SolverCfg {
    double tol;
    int maxiter;
    int solver;
    ExpTable ek;  // exp() constants in parameter space (uniform-register operands)
};

template <class C>
struct TpiState {  // what persists from sample to sample
    double x[dim1(C::NX)], lp[dim1(C::NP)], lz[dim1(C::NN)], Mx[dim1(C::NN * C::NP)];
};

// ---- small dense LU in registers: setlhs!/solve! (solvers.jl:46-132), fully unrolled.
// Deliberately NOT fused (no FMA): the reference's LU is plain scalar Julia, and an exactly
// singular Jacobian has to give an exactly zero pivot so that Newton bails out into the homotopy
// (with FMA the pivot becomes rounding noise and Newton wanders for hundreds of iterations). ----
template <int N>
__device__ __forceinline__ bool lu_reg(double (&A)[dim1(N * N)], int (&piv)[dim1(N)]) {
    bool ok = true;
    static_for<0, N>([&](auto kk) {
        constexpr int k = decltype(kk)::value;
        // first strict maximum of |A[i][k]|, i >= k (solvers.jl:60-68).  Starting from the diagonal
        // entry instead of 0.0 only differs for NaN entries, which the caller has already rejected.
        int kp = k;
        double amax = fabs(A[k * N + k]);
        static_for<k + 1, N>([&](auto ii) {
            constexpr int i = decltype(ii)::value;
            const double absi = fabs(A[k * N + i]);
            if (absi > amax) { kp = i; amax = absi; }
        });
        piv[k] = kp;
        static_for<k + 1, N>([&](auto ii) {
            constexpr int i = decltype(ii)::value;
            if (kp == i) {
                static_for<0, N>([&](auto cc) {
                    constexpr int c = decltype(cc)::value;
                    const double t = A[c * N + k];
                    A[c * N + k] = A[c * N + i];
                    A[c * N + i] = t;
                });
            }
        });
        if (A[k * N + k] == 0.0) ok = false;  // exactly singular: the reference bails out (solvers.jl:84-86)
        const double inv = 1.0 / A[k * N + k];
        A[k * N + k] = inv;
        static_for<k + 1, N>([&](auto ii) { A[k * N + decltype(ii)::value] *= inv; });
        static_for<k + 1, N>([&](auto jj) {
            constexpr int j = decltype(jj)::value;
            const double akj = A[j * N + k];
            static_for<k + 1, N>([&](auto ii) {
                constexpr int i = decltype(ii)::value;
                A[j * N + i] = __dsub_rn(A[j * N + i], __dmul_rn(A[k * N + i], akj));
            });
        });
    });
    return ok;
}

template <int N>
__device__ __forceinline__ void lu_solve_reg(const double (&A)[dim1(N * N)], const int (&piv)[dim1(N)],
                                             double (&x)[dim1(N)]) {
    static_for<0, N>([&](auto kk) {
        constexpr int k = decltype(kk)::value;
        static_for<k + 1, N>([&](auto ii) {
            constexpr int i = decltype(ii)::value;
            if (piv[k] == i) { const double t = x[k]; x[k] = x[i]; x[i] = t; }
        });
    });
    static_for<0, N>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        static_for<j + 1, N>([&](auto ii) {
            constexpr int i = decltype(ii)::value;
            x[i] = __dsub_rn(x[i], __dmul_rn(A[j * N + i], x[j]));
        });
    });
    static_for<0, N>([&](auto jr) {
        constexpr int j = N - 1 - decltype(jr)::value;
        x[j] = A[j * N + j] * x[j];
        static_for<0, j>([&](auto ii) {
            constexpr int i = decltype(ii)::value;
            x[i] = __dsub_rn(x[i], __dmul_rn(A[j * N + i], x[j]));
        });
    });
}

// ---- model closures (ACME.jl:176-194, 236-252), compile-time shapes ----
template <class C, class M>
__device__ __forceinline__ void tpi_set_p(const M& m, const double (&p)[dim1(C::NP)], double (&pfull)[dim1(C::NQ)]) {
    static_for<0, C::NQ>([&](auto ii) {
        constexpr int i = decltype(ii)::value;
        double acc = m.q0[i];
        static_for<0, C::NP>([&](auto jj) {
            constexpr int j = decltype(jj)::value;
            acc = fma(m.pexp[j * C::NQ + i], p[j], acc);
        });
        pfull[i] = acc;
    });
}

// evaluate!: returns max|res| (NaN if any residual is NaN); J column-major; jv kept for calc_Jp
template <class C, class M>
__device__ __forceinline__ double tpi_evaluate(const M& m, const double (&Cn)[dim1(C::NC)],
                                               const double (&pfull)[dim1(C::NQ)], const double (&z)[dim1(C::NN)],
                                               double (&res)[dim1(C::NN)], double (&jv)[dim1(C::NJ)],
                                               double (&J)[dim1(C::NN * C::NN)], bool& Jfinite, const SolverCfg& sc) {
    double q[dim1(C::NQ)];
    static_for<0, C::NQ>([&](auto ii) {
        constexpr int i = decltype(ii)::value;
        double acc = pfull[i];
        static_for<0, C::NN>([&](auto jj) {
            constexpr int j = decltype(jj)::value;
            acc = fma(m.fq[j * C::NQ + i], z[j], acc);
        });
        q[i] = acc;
    });
    // max|res| plus ONE finiteness test for everything the reference checks
    // (isfinite(resmaxabs) && all(isfinite, J), solvers.jl:220): the sum of the
    // absolute residuals and variable Jacobian entries is finite iff all of them are
    // (J = Jq*fq is a finite-coefficient combination of those entries).
    double resmax = 0.0, rsum = 0.0, jsum = 0.0;
    for_each_elem<0, 0, 0, 0>(typename C::Elems{}, [&](auto e, auto row, auto qoff, auto coff, auto joff) {
        using E = decltype(e);
        constexpr int ROW = decltype(row)::value, QOFF = decltype(qoff)::value, COFF = decltype(coff)::value,
                      JOFF = decltype(joff)::value;
        E::eval(&Cn[COFF], &q[QOFF], &res[ROW], &jv[JOFF], sc.ek.k);
        static_for<0, E::NJ>([&](auto kk) { jsum += fabs(jv[JOFF + decltype(kk)::value]); });
        static_for<0, E::NN>([&](auto rr) {
            constexpr int r = decltype(rr)::value;
            const double ar = fabs(res[ROW + r]);
            rsum += ar;
            resmax = ar > resmax ? ar : resmax;
            static_for<0, C::NN>([&](auto cc) {
                constexpr int c = decltype(cc)::value;
                J[c * C::NN + ROW + r] =
                    E::template row<r>(&jv[JOFF], [&](int k) { return m.fq[c * C::NQ + QOFF + k]; });
            });
        });
    });
    const double dmax = 1.7976931348623157e308;
    Jfinite = (rsum + jsum) <= dmax;        // false for NaN and Inf
    if (!Jfinite && !(rsum <= dmax)) return NAN;  // non-finite residual: hasconverged() is false
    return resmax;
}

// Mx = LU \ (Jq*pexp)  (calc_Jp! ACME.jl:246-251 followed by the solve of solvers.jl:213)
template <class C, class M>
__device__ __forceinline__ void tpi_update_Mx(const M& m, const double (&jv)[dim1(C::NJ)],
                                              const double (&LU)[dim1(C::NN * C::NN)],
                                              const int (&piv)[dim1(C::NN)], double (&Mx)[dim1(C::NN * C::NP)]) {
    static_for<0, C::NP>([&](auto cc) {
        constexpr int c = decltype(cc)::value;
        double col[dim1(C::NN)];
        for_each_elem<0, 0, 0, 0>(typename C::Elems{}, [&](auto e, auto row, auto qoff, auto, auto joff) {
            using E = decltype(e);
            constexpr int ROW = decltype(row)::value, QOFF = decltype(qoff)::value, JOFF = decltype(joff)::value;
            static_for<0, E::NN>([&](auto rr) {
                constexpr int r = decltype(rr)::value;
                col[ROW + r] = E::template row<r>(&jv[JOFF], [&](int k) { return m.pexp[c * C::NQ + QOFF + k]; });
            });
        });
        lu_solve_reg<C::NN>(LU, piv, col);
        static_for<0, C::NN>([&](auto ii) { Mx[c * C::NN + decltype(ii)::value] = col[decltype(ii)::value]; });
    });
}

// set_extrapolation_origin(solver, p, z)  (solvers.jl:183-196)
template <class C, class M>
__device__ __forceinline__ void tpi_set_origin(const M& m, const double (&Cn)[dim1(C::NC)], TpiState<C>& S,
                                               const double (&p)[dim1(C::NP)], const double (&z)[dim1(C::NN)],
                                               const SolverCfg& sc) {
    double pfull[dim1(C::NQ)], res[dim1(C::NN)], jv[dim1(C::NJ)], J[dim1(C::NN * C::NN)];
    int piv[dim1(C::NN)];
    bool Jfin;
    tpi_set_p<C>(m, p, pfull);
    tpi_evaluate<C>(m, Cn, pfull, z, res, jv, J, Jfin, sc);
    lu_reg<C::NN>(J, piv);
    tpi_update_Mx<C>(m, jv, J, piv, S.Mx);
    static_for<0, C::NP>([&](auto i) { S.lp[decltype(i)::value] = p[decltype(i)::value]; });
    static_for<0, C::NN>([&](auto i) { S.lz[decltype(i)::value] = z[decltype(i)::value]; });
}

// solve(::SimpleSolver, p)  (solvers.jl:207-236).  With `reorigin` the solve is preceded by
// set_extrapolation_origin(solver, pc, zc) (solvers.jl:183-196) for a cached solution (pc, zc):
// that is what the CachingSolver does when a stored point is nearer to p than the current origin
// (solvers.jl:347-371).  The origin evaluation runs as "iteration 0" of the same loop so that the
// element laws and the LU exist once in the instruction stream.  pc/zc are strided (SoA cache).
template <class C, class M>
__device__ __forceinline__ bool tpi_simple_solve(const M& m, const double (&Cn)[dim1(C::NC)], TpiState<C>& S,
                                                 const double (&p)[dim1(C::NP)], double (&z)[dim1(C::NN)],
                                                 const SolverCfg& sc, int& iters, bool reorigin = false,
                                                 const double* pc = nullptr, const double* zc = nullptr,
                                                 int64_t cstride = 0) {
    double pfull[dim1(C::NQ)], res[dim1(C::NN)], jv[dim1(C::NJ)], J[dim1(C::NN * C::NN)];
    int piv[dim1(C::NN)];
    auto start_from_origin = [&]() {
        tpi_set_p<C>(m, p, pfull);
        static_for<0, C::NN>([&](auto ii) {
            constexpr int i = decltype(ii)::value;
            double acc = S.lz[i];
            static_for<0, C::NP>([&](auto jj) {
                constexpr int j = decltype(jj)::value;
                acc = fma(-S.Mx[j * C::NN + i], p[j] - S.lp[j], acc);
            });
            z[i] = acc;
        });
    };
    if (reorigin) {
        double pcv[dim1(C::NP)];
        static_for<0, C::NP>([&](auto i) { pcv[decltype(i)::value] = pc[(int64_t)decltype(i)::value * cstride]; });
        static_for<0, C::NN>([&](auto i) { z[decltype(i)::value] = zc[(int64_t)decltype(i)::value * cstride]; });
        tpi_set_p<C>(m, pcv, pfull);
        static_for<0, C::NP>([&](auto i) { S.lp[decltype(i)::value] = pcv[decltype(i)::value]; });  // origin p, final below
    } else {
        start_from_origin();
    }
    bool converged = false;
    for (iters = reorigin ? 0 : 1; iters <= sc.maxiter; iters++) {
        bool Jfin;
        const double resmax = tpi_evaluate<C>(m, Cn, pfull, z, res, jv, J, Jfin, sc);
        if (iters == 0) {
            lu_reg<C::NN>(J, piv);
            tpi_update_Mx<C>(m, jv, J, piv, S.Mx);
            static_for<0, C::NN>([&](auto i) { S.lz[decltype(i)::value] = z[decltype(i)::value]; });
            start_from_origin();
            continue;
        }
        // one exit test after the factorisation: non-finite residual/Jacobian (solvers.jl:220) or an
        // exactly singular Jacobian (solvers.jl:223) end the solve; hasconverged() is resmax < tol
        const bool lu_ok = lu_reg<C::NN>(J, piv);
        if (!(Jfin && lu_ok)) return resmax < sc.tol;
        if (resmax < sc.tol) { converged = true; break; }
        lu_solve_reg<C::NN>(J, piv, res);
        static_for<0, C::NN>([&](auto i) { z[decltype(i)::value] -= res[decltype(i)::value]; });
    }
    if (iters > sc.maxiter) iters = sc.maxiter;
    if (converged) {
        tpi_update_Mx<C>(m, jv, J, piv, S.Mx);
        static_for<0, C::NP>([&](auto i) { S.lp[decltype(i)::value] = p[decltype(i)::value]; });
        static_for<0, C::NN>([&](auto i) { S.lz[decltype(i)::value] = z[decltype(i)::value]; });
    }
    return converged;
}

// ---- the CachingSolver's solution store (solvers.jl:319-396; kdcache.cuh: the reference's k-d tree on the device).
// The searches, the stores and the tree rebuilds are out-of-line: a thread whose store holds nothing but the initial
// solution (0, init_z) -- every instance of config 2 for most of its run -- decides in registers and touches no memory.
template <class C>
__device__ __forceinline__ KdStore tpi_store(const DevSub& c, int64_t inst) {
    return KdStore::at(c.kd_base + inst * c.kd_stride, c.kd_scr ? c.kd_scr + inst * c.kd_sstride : nullptr, C::NP, C::NN, c.kd_cap);
}
// leaf mirror (devmodel.h): leaf `leaf` (0-based), dimension d of this instance
template <class C>
__device__ __forceinline__ double* tpi_mir(const DevSub& c, int64_t inst, int leaf, int d) {
    return c.kd_mir + ((int64_t)leaf * C::NP + d) * c.kd_mld + (c.kd_mshared ? 0 : inst);
}
// after a (re)build by one thread: the tree's points in leaf order
template <class C>
__device__ inline void tpi_mirror_fill(const DevSub& cache, const KdStore& c, int64_t inst) {
    if (!cache.kd_mir || cache.kd_mshared) return;
    const int n = c.hdr[KD_H_TREEN];
    for (int leaf = 0; leaf < n; leaf++) {
        const int col = c.psidx(leaf + 1);
        for (int d = 0; d < C::NP; d++) *tpi_mir<C>(cache, inst, leaf, d) = c.P(d, col);
    }
}
// after a store into column `col`: a spare (zero) column that the tree holds as a leaf has become a real point
// (kdtree.jl:37 lets spare columns into the tree; the reference's leaf then sees the new point through tree.ps)
template <class C>
__device__ inline void tpi_mirror_fix(const DevSub& cache, const KdStore& c, int64_t inst, int col) {
    if (!cache.kd_mir || cache.kd_mshared || col <= 0) return;
    const int n = c.hdr[KD_H_TREEN];
    for (int leaf = 0; leaf < n; leaf++)
        if (c.psidx(leaf + 1) == col)
            for (int d = 0; d < C::NP; d++) *tpi_mir<C>(cache, inst, leaf, d) = c.P(d, col);
}

// register copy of what the hot path needs to know about the store
struct TpiKd {
    int newc;   // new_count: > 0 means every solve counts down to the next rebuild (solvers.jl:387-389)
    bool triv;  // learning store holding only column 1 = (0, init_z), tree of one leaf, nothing new
    bool on;    // the model has a store at all
    bool due;   // the tree is due for a rebuild (solvers.jl:390): done by the whole warp at the end of the sample
    int treen;  // points in the current tree
    int num;    // stored solutions
    int limit;  // new_count_limit: counted down in this register while new_count > 0 (solvers.jl:387-389), written back to
                // the store's header before anything else reads it (tpi_kd_flush)
    bool sm;    // the lookup data of this (small) store is in the warp's shared memory (tpi_kd_sm_fill)
};
// A small learning store -- a tree of at most TPI_KT leaves and at most TPI_KN new solutions, what the stiff corner of a
// parameter sweep holds (config 2: four solutions per instance) -- keeps what the start-point choice reads in the warp's
// shared memory: the p-coordinates of the tree's points in leaf order with their columns, then those of the newest
// entries, [slot][dimension][lane].  The choice then costs no global load (the chosen column's (p, z) still come from
// the store).  Measured on config 2 (tools/tail_prof.py): the kernel ends with its slowest warps, and those are the
// warps of the stiff corner, whose lanes choose among four solutions every sample and re-origin on a third of them;
// their per-instance store blocks (one cache line per lane and access) do not stay in L1.  Filled whenever the
// register copy of the header is (re)read (tpi_kd_reload).  Tried and measured slower: (z) in shared memory as well
// (+25 % shared memory per warp: 18.2 against 21.0 Gsamples/s), and a memo of the extrapolation matrix per stored
// solution (saves the origin evaluation, costs a round trip to L2: 18.9 against 19.7).
// ACME_TPI_SMKD: 0 off, 1 every shape, 2 (default) shapes with one parameter -- with more the tiles' share of L1 is
// worth more to the large stores of config 5 than this is (measured: birdie -16 % with it on).
#ifndef ACME_TPI_SMKD
#define ACME_TPI_SMKD 2
#endif
constexpr int TPI_KT = 4, TPI_KN = 4;
template <class C> __host__ __device__ constexpr bool tpi_smkd() { return C::NN > 0 && (ACME_TPI_SMKD == 1 || (ACME_TPI_SMKD == 2 && C::NP == 1)); }
// the register copy of new_count_limit back into the store's header
template <class C>
__device__ __forceinline__ void tpi_kd_flush(const DevSub& c, int64_t inst, const TpiKd& kd) {
    if (kd.on && !c.kd_frozen) reinterpret_cast<int*>(c.kd_base + inst * c.kd_stride)[KD_H_LIMIT] = kd.limit;
}
template <class C>
__device__ __forceinline__ TpiKd tpi_kd_state(const DevSub& c, int64_t inst) {
    TpiKd k{0, false, false, false, 0, 0, 0, false};
    if (c.kd_cap > 0) {
        const int* h = reinterpret_cast<const int*>(c.kd_base + inst * c.kd_stride);
        k.on = true;
        k.newc = h[KD_H_NEW];
        k.treen = h[KD_H_TREEN];
        k.num = h[KD_H_NUM];
        k.limit = h[KD_H_LIMIT];
        k.triv = !c.kd_frozen && h[KD_H_NUM] == 1 && h[KD_H_NEW] == 0 && h[KD_H_TREEN] == 1;
    }
    return k;
}
// header -> registers, small stores -> shared memory (smd: this lane's first double, smi: this lane's first int)
template <class C>
__device__ __noinline__ TpiKd tpi_kd_reload(const DevSub* cachep, int64_t inst, double* smd, int* smi) {
    const DevSub& cache = *cachep;
    TpiKd kd = tpi_kd_state<C>(cache, inst);
    if (tpi_smkd<C>() && kd.on && !kd.triv && !cache.kd_frozen && kd.treen <= TPI_KT && kd.newc <= TPI_KN && kd.treen >= 1) {
        const KdStore c = tpi_store<C>(cache, inst);
        for (int leaf = 0; leaf < kd.treen; leaf++) {
            const int col = c.psidx(leaf + 1);
            smi[leaf * 32] = col;
            for (int d = 0; d < C::NP; d++) smd[(leaf * C::NP + d) * 32] = c.P(d, col);
        }
        for (int i = 0; i < kd.newc; i++)
            for (int d = 0; d < C::NP; d++) smd[((TPI_KT + i) * C::NP + d) * 32] = c.P(d, kd.num - kd.newc + 1 + i);
        kd.sm = true;
    }
    return kd;
}
template <class C>
__device__ __noinline__ int tpi_kd_lookup(const DevSub* cache, int64_t inst, const double* p, double best) {
    const KdStore c = tpi_store<C>(*cache, inst);
    int ovf = 0;
    const int idx = kd_lookup(c, [&](int i) { return p[i]; }, best, &ovf);
    if (ovf) c.hdr[KD_H_FLAGS] |= KD_F_HEAP_OVERFLOW;
    return idx;
}
// solvers.jl:374-389; returns whether the tree is due for a rebuild
template <class C>
__device__ __noinline__ bool tpi_kd_after(const DevSub* cache, int64_t inst, const double* p, const double* z, int store) {
    KdStore c = tpi_store<C>(*cache, inst);
    int stored = 0;
    const bool due = kd_store_step(c, store != 0, [&](int i) { return p[i]; }, [&](int i) { return z[i]; }, &stored);
    if (stored) tpi_mirror_fix<C>(*cache, c, inst, stored);
    return due;
}
// squared distance, accumulated unfused like the reference's loops (solvers.jl:349-352)
// indnearest (kdtree.jl:192-234) written out for a tree of one or two leaves and nothing new in the store -- what an
// instance of an easy circuit holds once it has stored its first solution (config 2: a thread that calls the general
// search every sample holds its whole warp back, and the kernel ends with its slowest warp).  Same decisions in the
// same order as the general search: near leaf first, the far leaf only if its bound is below the best distance both
// when it is enqueued and after the near leaf was seen.
template <class C>
__device__ __forceinline__ int tpi_kd_small(const DevSub& cache, int64_t inst, const double (&p)[dim1(C::NP)], double best, int treen) {
    const double* const cst = cache.kd_base + inst * cache.kd_stride;
    const KdNode* const nd = reinterpret_cast<const KdNode*>(cst + KD_HDR_INTS / 2);
    const double* const cols = cst + KD_HDR_INTS / 2 + 2 * (int64_t)cache.kd_cap;
    const int cap = cache.kd_cap;
    auto dist = [&](int col) {
        const double* const pc = cols + (int64_t)((col <= cap ? col : 1) - 1) * (C::NP + C::NN);
        double acc = 0.0;
        static_for<0, C::NP>([&](auto ii) {
            constexpr int i = decltype(ii)::value;
            const double d = p[i] - (col <= cap ? pc[i] : 0.0);
            acc = __dadd_rn(acc, __dmul_rn(d, d));
        });
        return acc;
    };
    int bidx = 0;
    if (treen <= 1) {
        if (treen == 1) {
            const int col = (int)((unsigned)nd[0].ti >> 8);
            if (dist(col) < best) bidx = col;
        }
        return bidx;
    }
    const KdNode n0 = nd[0], n1 = nd[1];
    const int dim = (n0.ti & 0xff) - 1;
    double pd = p[0];
    static_for<1, C::NP>([&](auto ii) { pd = dim == decltype(ii)::value ? p[decltype(ii)::value] : pd; });
    const double dcut = pd - n0.cut_val;
    const double far_norm = __dmul_rn(dcut, dcut);  // (0 - 0*0) + dcut^2
    bool far = far_norm < best;                      // enqueue!  (kdtree.jl:207-210)
    const bool left = pd <= n0.cut_val;
    const int cnear = (int)((unsigned)(left ? n0.ti : n1.ti) >> 8), cfar = (int)((unsigned)(left ? n1.ti : n0.ti) >> 8);
    const double dn = dist(cnear);
    if (dn < best) { best = dn; bidx = cnear; far = far && far_norm < best; }  // update_best_dist! prunes (kdtree.jl:177-187)
    if (far) {
        const double df = dist(cfar);
        if (df < best) bidx = cfar;
    }
    return bidx;
}

// KDTree(ps, num_ps) for the store of instance `inst`, by the whole warp (every lane calls this with the same inst)
template <class C>
__device__ __noinline__ void tpi_kd_rebuild(const DevSub* cache, int64_t inst, int lane) {
    KdStore c = tpi_store<C>(*cache, inst);
    __syncwarp();
    const int num = c.hdr[KD_H_NUM], cap_ref = (c.hdr[KD_H_FLAGS] & KD_F_FULL) ? num : c.hdr[KD_H_CAPREF];  // kdcache.cuh kd_after_solve
    kd_build_warp(c, num, num, cap_ref, lane);  // spare columns, physical or not, are zeros: the virtual ones of kd_build
    if (lane == 0) kd_rebuilt(c);
    if (cache->kd_mir && !cache->kd_mshared)  // the leaf mirror, lanes over leaves
        for (int leaf = lane; leaf < num; leaf += 32) {
            const int col = c.psidx(leaf + 1);
            for (int d = 0; d < C::NP; d++) *tpi_mir<C>(*cache, inst, leaf, d) = c.P(d, col);
        }
    __threadfence_block();
    __syncwarp();
}
template <class C>
__device__ __forceinline__ double tpi_dist2(const double (&p)[dim1(C::NP)], const double (&q)[dim1(C::NP)]) {
    double acc = 0.0;
    static_for<0, C::NP>([&](auto ii) {
        constexpr int i = decltype(ii)::value;
        const double d = p[i] - q[i];
        acc = __dadd_rn(acc, __dmul_rn(d, d));
    });
    return acc;
}

// everything the cold paths need, spilled to local memory on purpose
template <class C>
struct TpiCold {
    TpiState<C> S;
    double Cn[dim1(C::NC)];
    double p[dim1(C::NP)];
    double z[dim1(C::NN)];
    int64_t inst, ld;
    int iters;
    int used_homotopy;
};

// solve(::CachingSolver, p) (solvers.jl:347-396) on the spilled state
template <class C, class M>
__device__ __forceinline__ bool tpi_base_solve_cold(const M& m, TpiCold<C>& k, const double (&p)[dim1(C::NP)],
                                                    const SolverCfg& sc, const DevSub& cache, int& iters) {
    if (sc.solver == ACMEB200_SOLVER_HOMOTOPY_CACHING && cache.kd_cap > 0) {
        const double best = tpi_dist2<C>(p, k.S.lp);
        KdStore c = tpi_store<C>(cache, k.inst);
        int ovf = 0;
        const int idx = kd_lookup(c, [&](int i) { return p[i]; }, best, &ovf);
        if (ovf) c.hdr[KD_H_FLAGS] |= KD_F_HEAP_OVERFLOW;
        if (idx != 0) {
            double cp[dim1(C::NP)], cz[dim1(C::NN)];
            for (int i = 0; i < C::NP; i++) cp[i] = c.P(i, idx);
            for (int i = 0; i < C::NN; i++) cz[i] = c.Z(i, idx);
            tpi_set_origin<C>(m, k.Cn, k.S, cp, cz, sc);
        }
        const bool conv = tpi_simple_solve<C>(m, k.Cn, k.S, p, k.z, sc, iters);
        int stored = 0;
        const bool rebuilt = kd_after_solve(c, conv && iters > 5, [&](int i) { return p[i]; }, [&](int i) { return k.z[i]; }, &stored);
        if (rebuilt) tpi_mirror_fill<C>(cache, c, k.inst);
        else if (stored) tpi_mirror_fix<C>(cache, c, k.inst, stored);
        return conv;
    }
    return tpi_simple_solve<C>(m, k.Cn, k.S, p, k.z, sc, iters);
}

// cold entry: mode 0 = cache lookup + base solve, then homotopy if needed
// (solve(::HomotopySolver, p), solvers.jl:268-296); mode 1 = the base solve has
// already failed in the hot path, go straight to the homotopy.
template <class C, class M>
__device__ __noinline__ bool tpi_cold_solve(const M* mp, TpiCold<C>* kp, const SolverCfg* scp, const DevSub* cache,
                                            int mode) {
    const SolverCfg& sc = *scp;
    const M& m = *mp;
    TpiCold<C>& k = *kp;
    bool conv = false;
    int total = k.iters;
    if (mode == 0) {
        int it;
        conv = tpi_base_solve_cold<C>(m, k, k.p, sc, *cache, it);
        total = it;
    }
    k.used_homotopy = 0;
    if (!conv && sc.solver != ACMEB200_SOLVER_SIMPLE) {
        k.used_homotopy = 1;
        double a = 0.5, best_a = 0.0;
        double start_p[dim1(C::NP)], pa[dim1(C::NP)];
        for (int i = 0; i < C::NP; i++) start_p[i] = k.S.lp[i];
        while (best_a < 1) {
            for (int i = 0; i < C::NP; i++) {
                double v = start_p[i];
                v *= (1 - a);
                v += a * k.p[i];
                pa[i] = v;
            }
            int it;
            conv = tpi_base_solve_cold<C>(m, k, pa, sc, *cache, it);
            total += it;
            if (conv) {
                best_a = a;
                a = 1.0;
            } else {
                const double new_a = (a + best_a) / 2;
                if (!(best_a < new_a && new_a < a)) break;
                a = new_a;
            }
        }
    }
    k.iters = total;
    return conv;
}

#ifndef ACME_TPI_TPB
#define ACME_TPI_TPB 64
#endif
#ifndef ACME_TPI_MINB
#define ACME_TPI_MINB 8
#endif
// CTAs per SM of the "wide" instantiation of the non-linear kernels: a batch that needs at most this many CTAs per SM
// runs a build compiled for 255 registers (no spills in the sample loop; the 128-register build that lets all 1024 CTAs
// of config 2 be resident at once makes ~170 local-memory loads per sample in the BJT shapes).  Measured: birdie
// B = 32768 +21 %, the diode clipper at B = 32768 +20 %.
#ifndef ACME_TPI_WIDE_MINB
#define ACME_TPI_WIDE_MINB 4
#endif
#ifndef ACME_TPI_T
#define ACME_TPI_T 8
#endif
#ifndef ACME_TPI_STAGES
#define ACME_TPI_STAGES 2
#endif
#ifndef ACME_TPI_OSTAGES
#define ACME_TPI_OSTAGES 2
#endif
constexpr int TPI_T = ACME_TPI_T;      // samples per staged tile
constexpr int TPI_STAGES = ACME_TPI_STAGES;  // input tiles in flight per warp (power of two): the linear
                                             // kernel needs ~50 KB of loads in flight per SM to cover HBM latency
constexpr int TPI_OSTAGES = ACME_TPI_OSTAGES;  // output tiles per warp (1 or 2)
static_assert((TPI_STAGES & (TPI_STAGES - 1)) == 0 && TPI_STAGES >= 2, "stages: power of two");
static_assert(TPI_OSTAGES == 1 || TPI_OSTAGES == 2, "output stages");
constexpr int TPI_TPB = ACME_TPI_TPB;  // threads per block (2 warps): small CTAs balance 148 SMs

// Tensor maps of the launch's input and output streams (built by the host, launch_tpi):
// rank 2, dim0 = samples*channels of one instance (contiguous), dim1 = instances, box = one
// warp tile {TPI_T*channels, 32}.  *_ok = 0: stream not 16-byte aligned, synchronous path.
struct alignas(64) TpiMaps {
    CUtensorMap u, y;
    int in_ok, out_ok;
};

// per-warp shared memory: 2 input tiles and 2 output tiles (TMA double buffering both ways;
// dense box layout, row = instance), 8 histogram counters per lane, 2 mbarriers.
// TMA swizzle of a tile whose rows are `row_bytes` long: 16-byte chunks of a row are XOR-permuted
// with the row index so that the lanes of a warp -- each reading ITS row at the same column --
// spread over the shared-memory banks (dense 64-byte rows read column-wise are a 16-way bank
// conflict).  mask m: byte address bits [4, 4+log2(m+1)) ^= bits [7, ...)  (CU_TENSOR_MAP_SWIZZLE_32B/64B/128B).
__host__ __device__ constexpr int tpi_swizzle_mask(int row_bytes) {
    return row_bytes == 128 ? 7 : (row_bytes == 64 ? 3 : (row_bytes == 32 ? 1 : 0));
}

// per-warp shared memory: TPI_STAGES input tiles and 2 output tiles (TMA multi-buffering;
// box layout, row = instance, swizzled), 8 histogram counters per lane, one mbarrier per input stage.
//
// Sample-major streams (SMAJ, ACMEB200_SAMPLE_MAJOR): the tile is transposed -- one row per SAMPLE holding the
// 32 instances of the warp (box {32*channels, TPI_T} of a (channels*B, N) tensor).  Lanes read consecutive
// words of a row, which is conflict-free without a swizzle, and every box row is one contiguous
// 256*channels-byte segment of HBM instead of 32 segments of TPI_T*channels*8 bytes.
template <class C, bool SMAJ = false>
struct TpiSmem {
    static constexpr int IROW = (SMAJ ? 32 : TPI_T) * C::NU * 8;  // bytes per tile row, input
    static constexpr int OROW = (SMAJ ? 32 : TPI_T) * C::NY * 8;
    static constexpr int ISW = tpi_swizzle_mask(IROW), OSW = tpi_swizzle_mask(OROW);
    static constexpr int IN_TX = 32 * TPI_T * C::NU * 8;  // bytes one TMA tile load delivers
    static constexpr int IN_BYTES = (IN_TX + 1023) / 1024 * 1024, OUT_BYTES = (32 * TPI_T * C::NY * 8 + 1023) / 1024 * 1024;
    static constexpr int OUT_OFF = TPI_STAGES * IN_BYTES;
    static constexpr int HIST_OFF = OUT_OFF + TPI_OSTAGES * OUT_BYTES;
    static constexpr int BAR_OFF = HIST_OFF + 32 * 8 * 4;
    static constexpr int KD_OFF = (BAR_OFF + 8 * TPI_STAGES + 15) / 16 * 16;  // small solution stores (tpi_kd_reload)
    static constexpr int KDI_OFF = KD_OFF + (tpi_smkd<C>() ? 32 * (TPI_KT + TPI_KN) * C::NP * 8 : 0);
    static constexpr int KD_END = KDI_OFF + (tpi_smkd<C>() ? 32 * TPI_KT * 4 : 0);
    static constexpr int PER_WARP = (KD_END + 1023) / 1024 * 1024;  // swizzled tiles want 1024-byte alignment
    // Byte offset of (row, byte o within the row) inside a tile = row*ROW + (o ^ xor_term(row)):
    // swizzled rows are at most 128 bytes long, so address bits >= 7 depend on the row alone and
    // the XOR term is a per-lane constant (computed once, outside the sample loop).
    __device__ static __forceinline__ int in_xor(int row) { return (((row * IROW) >> 7) & ISW) << 4; }
    __device__ static __forceinline__ int out_xor(int row) { return (((row * OROW) >> 7) & OSW) << 4; }
};

template <class C>
constexpr size_t tpi_smem_bytes() {
    static_assert(TpiSmem<C, true>::PER_WARP == TpiSmem<C, false>::PER_WARP, "both tile orientations have the same footprint");
    return (size_t)(TPI_TPB / 32) * TpiSmem<C>::PER_WARP;
}

// p = dq*x + eq*u   (ACME.jl:678-686)
template <class C, class M>
__device__ __forceinline__ void tpi_calc_p(const M& m, const TpiState<C>& S, const double (&u)[dim1(C::NU)],
                                           double (&p)[dim1(C::NP)]) {
    constexpr int NX = C::NX, NU = C::NU, NP = C::NP;
    static_for<0, NP>([&](auto ii) {
        constexpr int i = decltype(ii)::value;
        double acc = 0.0;
        static_for<0, NX>([&](auto jj) { acc = fma(m.dq[decltype(jj)::value * NP + i], S.x[decltype(jj)::value], acc); });
        static_for<0, NU>([&](auto jj) { acc = fma(m.eq[decltype(jj)::value * NP + i], u[decltype(jj)::value], acc); });
        p[i] = acc;
    });
}

// y = y0 + dy*x + ey*u + fy*z (x before the update) and x = x0 + a*x + b*u + c*z  (ACME.jl:699-714)
template <class C, class M>
__device__ __forceinline__ void tpi_output_update(const M& m, TpiState<C>& S, const double (&u)[dim1(C::NU)],
                                                  const double (&zall)[dim1(C::NN)], double (&y)[dim1(C::NY)]) {
    constexpr int NX = C::NX, NU = C::NU, NY = C::NY, NN = C::NN;
    static_for<0, NY>([&](auto ii) {
        constexpr int i = decltype(ii)::value;
        double acc = m.y0[i];
        static_for<0, NX>([&](auto jj) { acc = fma(m.dy[decltype(jj)::value * NY + i], S.x[decltype(jj)::value], acc); });
        static_for<0, NU>([&](auto jj) { acc = fma(m.ey[decltype(jj)::value * NY + i], u[decltype(jj)::value], acc); });
        static_for<0, NN>([&](auto jj) { acc = fma(m.fy[decltype(jj)::value * NY + i], zall[decltype(jj)::value], acc); });
        y[i] = acc;
    });
    double xn[dim1(NX)];
    static_for<0, NX>([&](auto ii) {
        constexpr int i = decltype(ii)::value;
        double acc = m.x0[i];
        static_for<0, NX>([&](auto jj) { acc = fma(m.a[decltype(jj)::value * NX + i], S.x[decltype(jj)::value], acc); });
        static_for<0, NU>([&](auto jj) { acc = fma(m.b[decltype(jj)::value * NX + i], u[decltype(jj)::value], acc); });
        static_for<0, NN>([&](auto jj) { acc = fma(m.c[decltype(jj)::value * NX + i], zall[decltype(jj)::value], acc); });
        xn[i] = acc;
    });
    static_for<0, NX>([&](auto i) { S.x[decltype(i)::value] = xn[decltype(i)::value]; });
}

// Hot step: the common case of step! (ACME.jl:666-715) -- the origin is the nearest start point
// and plain Newton converges.  Returns the iteration count (>= 1; 0 for linear models) when the
// sample was completed, or -(iters+1) when the cold path has to take this sample over (cache
// lookup says another start point is nearer, or Newton failed); the state is untouched then.
template <class C, class M>
__device__ __forceinline__ int tpi_step_hot(const M& m, const double (&Cn)[dim1(C::NC)], TpiState<C>& S,
                                            const double (&u)[dim1(C::NU)], double (&y)[dim1(C::NY)],
                                            const SolverCfg& sc, const DevSub& cache, int64_t inst, TpiKd& kd,
                                            const double* smd, const int* smi TPI_PROF(, TpiProf& pr)) {
    constexpr int NN = C::NN, NP = C::NP;
    double zall[dim1(NN)];
    int iters = 0;
    if constexpr (NN > 0) {
        double p[dim1(NP)];
        tpi_calc_p<C>(m, S, u, p);
        const bool caching = sc.solver == ACMEB200_SOLVER_HOMOTOPY_CACHING && kd.on;
        int col = 0;  // stored solution (1-based column) to re-origin at, 0: keep the origin
        if (caching) {
            const double best = tpi_dist2<C>(p, S.lp);
            if (kd.triv) {  // the store holds only (0, init_z): its squared distance is |p|^2 -- no memory touched
                double d0 = 0.0;
                static_for<0, NP>([&](auto ii) { d0 = __dadd_rn(d0, __dmul_rn(p[decltype(ii)::value], p[decltype(ii)::value])); });
                col = d0 < best ? 1 : 0;
            } else if (tpi_smkd<C>() && kd.sm) {
                TPI_PROF(pr.nontriv++; pr.small++;)
                // the same choice as below from shared memory: newest entries in index order (solvers.jl:354-363), then the
                // nearest tree point if it is strictly nearer (kdtree.jl:192-234); two leaves tying -> the serial search
                constexpr int F = NP;
                double bseed = best;
#pragma unroll
                for (int i = 0; i < TPI_KN; i++)
                    if (i < kd.newc) {
                        double acc = 0.0;
                        static_for<0, NP>([&](auto ii) {
                            constexpr int k = decltype(ii)::value;
                            const double d = smd[((TPI_KT + i) * F + k) * 32] - p[k];
                            acc = __dadd_rn(acc, __dmul_rn(d, d));
                        });
                        if (acc < bseed) { bseed = acc; col = kd.num - kd.newc + 1 + i; }
                    }
                double bd = __longlong_as_double(0x7ff0000000000000ll);
                int bleaf = 0;
                bool tie = false;
#pragma unroll
                for (int l = 0; l < TPI_KT; l++)
                    if (l < kd.treen) {
                        double acc = 0.0;
                        static_for<0, NP>([&](auto ii) {
                            constexpr int k = decltype(ii)::value;
                            const double d = p[k] - smd[(l * F + k) * 32];
                            acc = __dadd_rn(acc, __dmul_rn(d, d));
                        });
                        tie = acc < bd ? false : (acc == bd ? true : tie);
                        if (acc < bd) { bd = acc; bleaf = l; }
                    }
                if (bd < bseed) {
                    if (!tie) {
                        col = smi[bleaf * 32];
                    } else {
                        double pl[dim1(NP)];
                        static_for<0, NP>([&](auto ii) { pl[decltype(ii)::value] = p[decltype(ii)::value]; });
                        col = tpi_kd_lookup<C>(&cache, inst, pl, best);
                        TPI_PROF(pr.tie++;)
                    }
                }
            } else {
                TPI_PROF(pr.nontriv++;)
                // the new_count newest stored solutions, first minimum in index order (solvers.jl:354-363) ...
                double bseed = best;
                if (kd.newc > 0) {
                    const double* const cols = cache.kd_base + inst * cache.kd_stride + KD_HDR_INTS / 2 + 2 * (int64_t)cache.kd_cap;
#pragma unroll 1
                    for (int i = kd.num - kd.newc + 1; i <= kd.num; i++) {
                        const double* const pc = cols + (int64_t)(i - 1) * (NP + NN);
                        double acc = 0.0;
                        static_for<0, NP>([&](auto ii) {
                            constexpr int k = decltype(ii)::value;
                            const double d = pc[k] - p[k];
                            acc = __dadd_rn(acc, __dmul_rn(d, d));
                        });
                        if (acc < bseed) { bseed = acc; col = i; }
                    }
                }
                // ... then the tree, seeded with that (kdtree.jl:93-100, 192-234)
                if (kd.treen <= 2) {
                    TPI_PROF(pr.small++;)
                    const int tcol = tpi_kd_small<C>(cache, inst, p, bseed, kd.treen);
                    col = tcol ? tcol : col;
                } else if (cache.kd_mir) {
                    TPI_PROF(pr.scan++;)
                    // every leaf of the tree, several per pass (independent loads in flight together), lanes in step: the
                    // nearest tree point, taken if it is strictly nearer than the seed (what indnearest returns)
                    double bd = __longlong_as_double(0x7ff0000000000000ll);
                    int bleaf = 0;
                    bool tie = false;
                    const int nt = kd.treen;
                    constexpr int UNR = NP <= 2 ? 8 : 4;
                    // addresses by pointer steps (the mirror is [(leaf*np + d)*ld + instance]): the scan is 40 % of the
                    // instructions of config 5, and a third of those were 64-bit index arithmetic
                    const int64_t ls = cache.kd_mld;
                    const double* mp = cache.kd_mir + (cache.kd_mshared ? 0 : inst);  // leaf l0, dimension 0
                    auto dist_at = [&](const double* q) {
                        double acc = 0.0;
                        static_for<0, NP>([&](auto ii) {
                            constexpr int i = decltype(ii)::value;
                            const double d = p[i] - *q;
                            q += ls;
                            acc = i == 0 ? __dmul_rn(d, d) : __dadd_rn(acc, __dmul_rn(d, d));  // 0 + d*d = d*d exactly
                        });
                        return acc;
                    };
                    int l0 = 0;
#pragma unroll 1
                    for (; l0 + UNR <= nt; l0 += UNR) {  // full blocks
                        double d2[UNR];
#pragma unroll
                        for (int k = 0; k < UNR; k++) { d2[k] = dist_at(mp); mp += NP * ls; }
#pragma unroll
                        for (int k = 0; k < UNR; k++) {
                            tie = d2[k] < bd ? false : (d2[k] == bd ? true : tie);
                            if (d2[k] < bd) { bd = d2[k]; bleaf = l0 + k; }
                        }
                    }
                    if (l0 < nt) {  // the last, partial block: surplus slots re-read the last leaf and are ignored
                        double d2[UNR];
#pragma unroll
                        for (int k = 0; k < UNR; k++) { d2[k] = dist_at(mp); if (l0 + k + 1 < nt) mp += NP * ls; }
#pragma unroll
                        for (int k = 0; k < UNR; k++)
                            if (l0 + k < nt) {
                                tie = d2[k] < bd ? false : (d2[k] == bd ? true : tie);
                                if (d2[k] < bd) { bd = d2[k]; bleaf = l0 + k; }
                            }
                    }
                    if (bd < bseed) {
                        if (!tie) {
                            col = tpi_store<C>(cache, inst).psidx(bleaf + 1);
                        } else {  // two leaves tie for the minimum: the order of the reference's search decides
                            double pl[dim1(NP)];
                            static_for<0, NP>([&](auto ii) { pl[decltype(ii)::value] = p[decltype(ii)::value]; });
                            col = tpi_kd_lookup<C>(&cache, inst, pl, best);
                            TPI_PROF(pr.tie++;)
                        }
                    }
                } else {
                    double pl[dim1(NP)];
                    static_for<0, NP>([&](auto ii) { pl[decltype(ii)::value] = p[decltype(ii)::value]; });
                    col = tpi_kd_lookup<C>(&cache, inst, pl, best);
                }
            }
        }
        const double* const cst = cache.kd_base + inst * cache.kd_stride;  // this instance's store (kdcache.cuh layout)
        const int cap = cache.kd_cap;
        const int64_t e = col > 0 && col <= cap ? col - 1 : 0;
        const double* const pc = cst + KD_HDR_INTS / 2 + 2 * (int64_t)cap + e * (NP + NN);  // KdStore::col(col): p then z
        const double* const zc = pc + NP;
        TPI_PROF(if (col > 0) pr.reo++;)
        if (caching && col > cap) return -1;  // a virtual zero column as start point (kdcache.cuh): the cold path handles it
        const bool conv = tpi_simple_solve<C>(m, Cn, S, p, zall, sc, iters, col > 0, pc, zc, 1);
        if (caching && !cache.kd_frozen) {  // solvers.jl:374-394
            if (conv && iters > 5) {  // store (p, z), out of line
                double pl[dim1(NP)], zl[dim1(NN)];
                static_for<0, NP>([&](auto ii) { pl[decltype(ii)::value] = p[decltype(ii)::value]; });
                static_for<0, NN>([&](auto ii) { zl[decltype(ii)::value] = zall[decltype(ii)::value]; });
                tpi_kd_flush<C>(cache, inst, kd);
                TPI_PROF(pr.stores++;)
                const bool due = tpi_kd_after<C>(&cache, inst, pl, zl, 1);
                kd = tpi_kd_reload<C>(&cache, inst, const_cast<double*>(smd), const_cast<int*>(smi));
                kd.due = due;
            } else if (kd.newc > 0) {  // count down to the rebuild in registers
                kd.limit -= 1;
                if (kd.newc > kd.limit) { tpi_kd_flush<C>(cache, inst, kd); kd.due = true; }
            }
        }
        if (!conv) {
            if (sc.solver != ACMEB200_SOLVER_SIMPLE) return -(iters + 1);
            return -(iters + 1) - (1 << 20);  // SimpleSolver only: no homotopy, the failure is final
        }
    }
    tpi_output_update<C>(m, S, u, zall, y);
    return iters;
}

// Cold step: everything else (frozen-cache start points, homotopy, failures).  `code` is the
// negative value the hot step returned.  Out of line on purpose: its register needs must not
// shape the hot loop.
template <class C, class M>
__device__ __noinline__ int tpi_step_cold(const M* mp, const double* Cn_, TpiState<C>* Sp, const double* u_,
                                          double* y_, const SolverCfg* scp, const DevSub* cachep, const RunArgs* ap,
                                          int64_t inst, int n, int code) {
    constexpr int NN = C::NN, NP = C::NP, NU = C::NU, NY = C::NY;
    const M& m = *mp;
    const SolverCfg& sc = *scp;
    const RunArgs& a = *ap;
    TpiCold<C> k;
    k.S = *Sp;
    for (int i = 0; i < C::NC; i++) k.Cn[i] = Cn_[i];
    double u[dim1(NU)], y[dim1(NY)];
    for (int i = 0; i < NU; i++) u[i] = u_[i];
    tpi_calc_p<C>(m, k.S, u, k.p);
    for (int i = 0; i < NN; i++) k.z[i] = 0.0;
    k.inst = inst;
    k.ld = a.ld;
    const bool final_fail = code <= -(1 << 20);
    const int failed_iters = final_fail ? (-(code + (1 << 20)) - 1) : (-code - 1);
    bool conv = false;
    k.used_homotopy = 0;
    if (final_fail) {
        k.iters = failed_iters;
        // recompute the last iterate for the finite/non-finite distinction
        int it;
        conv = tpi_simple_solve<C>(m, k.Cn, k.S, k.p, k.z, sc, it);
        k.iters = it;
    } else {
        k.iters = failed_iters;
        conv = tpi_cold_solve<C>(&m, &k, &sc, cachep, failed_iters > 0 ? 1 : 0);
    }
    const int iters = k.iters;
    bool finite = true;
    if (!conv) for (int i = 0; i < NN; i++) finite = finite && isfinite(k.z[i]);
    if (k.used_homotopy) atomicAdd(&a.stats->homotopy_solves, 1ull);
    if (iters > 8) {
        const int bin = iters > ACMEB200_HIST_BINS ? ACMEB200_HIST_BINS : iters;
        atomicAdd(&a.stats->iter_hist[bin - 1], 1ull);
        atomicAdd(&a.stats->newton_iters, (unsigned long long)iters);
    }
    if (!conv) {
        if (a.first_fail[inst] < 0) a.first_fail[inst] = a.n_done + n;
        if (finite) {
            a.status[inst] |= ACMEB200_STATUS_NOT_CONVERGED;
            atomicAdd(&a.stats->not_converged, 1ull);
        } else {
            a.status[inst] |= ACMEB200_STATUS_NONFINITE;
            return -1;  // instance halts (the reference throws, ACME.jl:692)
        }
    }
    tpi_output_update<C>(m, k.S, u, k.z, y);
    *Sp = k.S;
    for (int i = 0; i < NY; i++) y_[i] = y[i];
    return iters;
}

// (RunArgs is a __grid_constant__ too: its address goes to the cold path, and a plain parameter would be copied to the
// stack and read from there -- measured +2.7 % on config 2.)
template <class C, bool PERINST, bool SMAJ, bool WIDE = false>
__global__ void __launch_bounds__(TPI_TPB, WIDE ? ACME_TPI_WIDE_MINB : ACME_TPI_MINB) k_tpi(const __grid_constant__ TpiMats<C> Msh, const __grid_constant__ RunArgs a,
                                                 const __grid_constant__ SolverCfg sc, const __grid_constant__ DevSub cache,
                                                 const __grid_constant__ TpiMaps maps) {
    constexpr int NX = C::NX, NU = C::NU, NY = C::NY, NN = C::NN, NP = C::NP;
    constexpr int T = TPI_T;
    using SM = TpiSmem<C, SMAJ>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* wsm = smem_raw + (size_t)warp * SM::PER_WARP;

    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // launch-local instance
    const bool active = t < a.ninst;
    const int64_t inst = a.inst0 + (active ? t : 0);

    // matrices: shared constant-bank copy, or this instance's own copy in registers
    TpiMats<C> Mown;
    if constexpr (PERINST) {
        const double* src = a.blob + inst * a.blob_stride;
        double* dst = reinterpret_cast<double*>(&Mown);
        for (int i = 0; i < (int)(sizeof(TpiMats<C>) / sizeof(double)); i++) dst[i] = 0.0;
        // blob order: a b c x0 dy ey fy y0 | dq eq fqprev pexp q0 fq (true, unpadded lengths)
        int o = 0;
        auto take = [&](double* f, int n) { for (int i = 0; i < n; i++) f[i] = __ldg(src + o + i); o += n; };
        take(Mown.a, NX * NX); take(Mown.b, NX * NU); take(Mown.c, NX * NN); take(Mown.x0, NX);
        take(Mown.dy, NY * NX); take(Mown.ey, NY * NU); take(Mown.fy, NY * NN); take(Mown.y0, NY);
        take(Mown.dq, NP * NX); take(Mown.eq, NP * NU);
        o += NP * NN;  // fqprev (single sub: unused)
        take(Mown.pexp, C::NQ * NP); take(Mown.q0, C::NQ); take(Mown.fq, C::NQ * NN);
    }
    const TpiMats<C>& m = PERINST ? Mown : Msh;

    double Cn[dim1(C::NC)];
    static_for<0, C::NC>([&](auto i) { Cn[decltype(i)::value] = a.consts[(int64_t)decltype(i)::value * a.ld + inst]; });
    TpiState<C> S;
    double* st = a.ws + inst;
    auto store_state = [&]() {
        static_for<0, NX>([&](auto i) { st[(int64_t)(C::S_X + decltype(i)::value) * a.ld] = S.x[decltype(i)::value]; });
        static_for<0, NP>([&](auto i) { st[(int64_t)(C::S_LP + decltype(i)::value) * a.ld] = S.lp[decltype(i)::value]; });
        static_for<0, NN>([&](auto i) { st[(int64_t)(C::S_LZ + decltype(i)::value) * a.ld] = S.lz[decltype(i)::value]; });
        static_for<0, NN * NP>([&](auto i) { st[(int64_t)(C::S_MX + decltype(i)::value) * a.ld] = S.Mx[decltype(i)::value]; });
    };

    if (a.init) {
        if (!active) return;
        static_for<0, NX>([&](auto i) { S.x[decltype(i)::value] = 0.0; });
        if constexpr (NN > 0) {
            double p0[dim1(NP)], z0[dim1(NN)];
            static_for<0, NP>([&](auto i) { p0[decltype(i)::value] = 0.0; });
            static_for<0, NN>([&](auto i) { z0[decltype(i)::value] = a.initz[(int64_t)decltype(i)::value * a.ld + inst]; });
            tpi_set_origin<C>(m, Cn, S, p0, z0, sc);
            if (cache.kd_cap > 0 && !cache.kd_frozen) {  // CachingSolver ctor: the store holds (0, init_z)  (solvers.jl:327-333); memory zeroed by the host
                KdStore c = tpi_store<C>(cache, inst);
                kd_init(c, [&](int i) { return z0[i]; });
            }
        }
        store_state();
        a.status[inst] = 0;
        a.first_fail[inst] = -1;
        return;
    }

    static_for<0, NX>([&](auto i) { S.x[decltype(i)::value] = st[(int64_t)(C::S_X + decltype(i)::value) * a.ld]; });
    static_for<0, NP>([&](auto i) { S.lp[decltype(i)::value] = st[(int64_t)(C::S_LP + decltype(i)::value) * a.ld]; });
    static_for<0, NN>([&](auto i) { S.lz[decltype(i)::value] = st[(int64_t)(C::S_LZ + decltype(i)::value) * a.ld]; });
    static_for<0, NN * NP>([&](auto i) { S.Mx[decltype(i)::value] = st[(int64_t)(C::S_MX + decltype(i)::value) * a.ld]; });

    TpiKd kd{0, false, false, false, 0, 0, 0, false};
    TPI_PROF(TpiProf pr{0, 0, 0, 0, 0, 0}; unsigned pr_cold = 0, pr_reb = 0; long long pr_t_cold = 0; const long long pr_t0 = clock64();)
    double* const kd_smd = reinterpret_cast<double*>(wsm + SM::KD_OFF) + lane;
    int* const kd_smi = reinterpret_cast<int*>(wsm + SM::KDI_OFF) + lane;
    if constexpr (NN > 0)
        if (sc.solver == ACMEB200_SOLVER_HOMOTOPY_CACHING && active) kd = tpi_kd_reload<C>(&cache, inst, kd_smd, kd_smi);
    bool dead = !active || (a.status[inst] & ACMEB200_STATUS_NONFINITE);
    int dead_at = dead ? 0 : -1;  // sample index (this call) at which the instance halted, -1 = alive
    const bool shared_u = (a.u_stride == 0) || NU == 0;
    const bool tma_in = !shared_u && maps.in_ok;   // the host checked alignment and built the maps
    const bool tma_out = NY > 0 && maps.out_ok;
    const int t32 = (int)t;                 // launch-local instance (ninst < 2^31)
    const int w0 = t32 - lane;              // first instance of this warp = tile row coordinate
    const bool wact = w0 < (int)a.ninst;    // warp has at least one instance (lane 0 is active)
    const int n_samp = (int)a.N;            // N < 2^27 (checked by the host)
    const int n_tiles = (n_samp + T - 1) / T;
    // first value of sample n of this lane's instance (the synchronous paths)
    auto urow = [&](int n) {
        if constexpr (SMAJ) return a.U + (int64_t)n * a.u_stride + (int64_t)t32 * NU;
        else return a.U + (shared_u ? 0 : (int64_t)t32 * a.u_stride) + (int64_t)n * NU;
    };
    auto yrow = [&](int n) {
        if constexpr (SMAJ) return a.Y + (int64_t)n * a.y_stride + (int64_t)t32 * NY;
        else return a.Y + (int64_t)t32 * a.y_stride + (int64_t)n * NY;
    };
    unsigned int* const hist_s = reinterpret_cast<unsigned int*>(wsm + SM::HIST_OFF) + lane;  // [bin*32]
    const uint32_t bar0 = smem_u32(wsm + SM::BAR_OFF);  // bar1 = bar0 + 8
    const int ixr = SM::in_xor(lane), oxr = SM::out_xor(lane);  // swizzle terms of this lane's rows

#pragma unroll
    for (int b = 0; b < 8; b++) hist_s[b * 32] = 0u;
    if (tma_in) {
        if (lane == 0)
            for (int st_ = 0; st_ < TPI_STAGES; st_++) mbar_init(bar0 + 8 * st_, 1);
        mbar_init_fence();
        __syncwarp();
        // prologue: the first TPI_STAGES tiles in flight
        if (lane == 0 && wact)
            for (int k = 0; k < TPI_STAGES && k < n_tiles; k++) {
                mbar_arrive_expect_tx(bar0 + 8 * k, SM::IN_TX);
                if constexpr (SMAJ) tma_load_2d(smem_u32(wsm + k * SM::IN_BYTES), &maps.u, w0 * NU, k * T, bar0 + 8 * k);
                else tma_load_2d(smem_u32(wsm + k * SM::IN_BYTES), &maps.u, k * T * NU, w0, bar0 + 8 * k);
            }
    }

#pragma unroll 1
    for (int k = 0; k < n_tiles; k++) {
        const int n0 = k * T;
        const int cnt = (n_samp - n0) < T ? (n_samp - n0) : T;
        const int buf = k & (TPI_STAGES - 1);
        unsigned char* const in_tile = wsm + buf * SM::IN_BYTES;
        unsigned char* const out_tile = wsm + SM::OUT_OFF + (k & (TPI_OSTAGES - 1)) * SM::OUT_BYTES;
        unsigned char* const in_row = in_tile + lane * SM::IROW;
        unsigned char* const out_row = out_tile + lane * SM::OROW;
        // value c = tt*channels + q of this lane's instance within the tile
        auto in_at = [&](int c) -> double& {
            if constexpr (SMAJ) return *reinterpret_cast<double*>(in_tile + ((c / dim1(NU)) * 32 * NU + lane * NU + c % dim1(NU)) * 8);
            else return *reinterpret_cast<double*>(in_row + ((c * 8) ^ ixr));
        };
        auto out_at = [&](int c) -> double& {
            if constexpr (SMAJ) return *reinterpret_cast<double*>(out_tile + ((c / dim1(NY)) * 32 * NY + lane * NY + c % dim1(NY)) * 8);
            else return *reinterpret_cast<double*>(out_row + ((c * 8) ^ oxr));
        };
        if (!shared_u) {
            if (tma_in) {
                if (wact) mbar_wait(bar0 + 8 * buf, (uint32_t)((k / TPI_STAGES) & 1));
            } else {
                // synchronous path (unaligned streams): own-row loads
                __syncwarp();
                if (active) {
                    if constexpr (SMAJ) {
                        for (int ts = 0; ts < cnt; ts++)
                            for (int q = 0; q < NU; q++) in_at(ts * NU + q) = __ldcs(urow(n0 + ts) + q);
                    } else {
                        for (int c = 0; c < cnt * NU; c++) in_at(c) = __ldcs(urow(n0) + c);
                    }
                }
            }
        }
        if (tma_out && k >= TPI_OSTAGES) {  // the store of tile k-TPI_OSTAGES has drained this output buffer
            if (lane == 0) { if (TPI_OSTAGES == 2) bulk_wait_read1(); else bulk_wait_read0(); }
            __syncwarp();
        }
        int tt = 0;
        if constexpr (NN == 0 && (SMAJ || ((SM::IROW % 16 == 0) && (SM::OROW % 16 == 0)))) {
            // ---- linear model: the whole tile row through registers.  128-bit loads/stores of the
            //      swizzled rows are bank-conflict free; the per-sample work is a handful of DFMAs
            //      (ACME.jl:699-714 with nn = 0), so this path is what makes the kernel HBM-bound.
            if (!dead) {
                double ub[dim1(T * NU)], yb[dim1(T * NY)];
                if (shared_u) {
                    static_for<0, T * NU>([&](auto cc) {
                        constexpr int c = decltype(cc)::value;
                        ub[c] = c < cnt * NU ? __ldg(a.U + (int64_t)n0 * NU + c) : 0.0;
                    });
                } else if constexpr (SMAJ) {
                    static_for<0, T * NU>([&](auto cc) { ub[decltype(cc)::value] = in_at(decltype(cc)::value); });
                } else {
                    static_for<0, SM::IROW / 16>([&](auto cc) {
                        constexpr int c = decltype(cc)::value;
                        const double2 v = *reinterpret_cast<const double2*>(in_row + ((c * 16) ^ ixr));
                        ub[2 * c] = v.x;
                        ub[2 * c + 1] = v.y;
                    });
                }
                static_for<0, T>([&](auto tc) {
                    constexpr int ti = decltype(tc)::value;
                    double u[dim1(NU)], y[dim1(NY)], znone[1] = {0.0};
                    static_for<0, NU>([&](auto kk) { u[decltype(kk)::value] = ub[ti * NU + decltype(kk)::value]; });
                    static_for<0, NY>([&](auto kk) { y[decltype(kk)::value] = 0.0; });
                    if (ti < cnt) tpi_output_update<C>(m, S, u, znone, y);  // warp-uniform: only the last tile is short
                    static_for<0, NY>([&](auto kk) { yb[ti * NY + decltype(kk)::value] = y[decltype(kk)::value]; });
                });
                if constexpr (SMAJ) {
                    static_for<0, T * NY>([&](auto cc) { out_at(decltype(cc)::value) = yb[decltype(cc)::value]; });
                } else {
                    static_for<0, SM::OROW / 16>([&](auto cc) {
                        constexpr int c = decltype(cc)::value;
                        *reinterpret_cast<double2*>(out_row + ((c * 16) ^ oxr)) = make_double2(yb[2 * c], yb[2 * c + 1]);
                    });
                }
                tt = cnt;
            }
        } else {
            // Every lane walks the tile in step (the trip counts are warp-uniform; halted and surplus lanes idle), so
            // that the warp is converged wherever it has to act together: a tree rebuild of one lane's store is done by
            // all 32 lanes (kdcache_warp.cuh).  The hot loop has no calls; it is left by the whole warp when some lane
            // needs the cold path or a rebuild.
            int code = 0;
            while (tt < cnt) {
    #pragma unroll 1
                for (; tt < cnt; tt++) {
                    code = 0;
                    if (!dead) {
                        double u[dim1(NU)], y[dim1(NY)];
                        static_for<0, NU>([&](auto kk) {
                            constexpr int q = decltype(kk)::value;
                            u[q] = shared_u ? __ldg(a.U + (int64_t)(n0 + tt) * NU + q) : in_at(tt * NU + q);
                        });
                        code = tpi_step_hot<C>(m, Cn, S, u, y, sc, cache, inst, kd, kd_smd, kd_smi TPI_PROF(, pr));
                        if (code >= 0) {
                            if (NN > 0) {
                                if (code <= 8) hist_s[(code - 1) * 32] += 1u;
                                else atomicAdd(&a.stats->iter_hist[(code > ACMEB200_HIST_BINS ? ACMEB200_HIST_BINS : code) - 1], 1ull),
                                     atomicAdd(&a.stats->newton_iters, (unsigned long long)code);
                            }
                            static_for<0, NY>([&](auto kk) { out_at(tt * NY + decltype(kk)::value) = y[decltype(kk)::value]; });
                        }
                    } else {
                        static_for<0, NY>([&](auto kk) { out_at(tt * NY + decltype(kk)::value) = NAN; });  // the reference throws (ACME.jl:692)
                    }
                    if (__any_sync(0xffffffffu, code < 0 || kd.due)) break;
                }
                if (tt < cnt) {  // warp-uniform
                    TPI_PROF(const long long pr_tc = clock64();)
                    if (code < 0) {
                        TPI_PROF(pr_cold++;)
                        // this lane's sample tt needs the cold path
                        double u[dim1(NU)], y[dim1(NY)];
                        static_for<0, NU>([&](auto kk) {
                            constexpr int q = decltype(kk)::value;
                            u[q] = shared_u ? __ldg(a.U + (int64_t)(n0 + tt) * NU + q) : in_at(tt * NU + q);
                        });
                        if constexpr (NN > 0) tpi_kd_flush<C>(cache, inst, kd);
                        const int it = tpi_step_cold<C>(&m, Cn, &S, u, y, &sc, &cache, &a, inst, n0 + tt, code);
                        if constexpr (NN > 0)
                            if (kd.on) kd = tpi_kd_reload<C>(&cache, inst, kd_smd, kd_smi);  // the cold solves may have stored solutions / rebuilt the tree
                        if (it < 0) {
                            dead = true; dead_at = n0 + tt;
                            static_for<0, NY>([&](auto kk) { out_at(tt * NY + decltype(kk)::value) = NAN; });
                        } else {
                            if (it >= 1 && it <= 8) hist_s[(it - 1) * 32] += 1u;
                            static_for<0, NY>([&](auto kk) { out_at(tt * NY + decltype(kk)::value) = y[decltype(kk)::value]; });
                        }
                    }
                    if constexpr (NN > 0) {
                        // KDTree(ps, num_ps) (solvers.jl:390) for every lane whose store is due, one store at a time, all lanes
                        unsigned due = __ballot_sync(0xffffffffu, kd.due);
                        while (due) {
                            const int src = __ffs(due) - 1;
                            due &= due - 1;
                            const int64_t isrc = __shfl_sync(0xffffffffu, (long long)inst, src);
                            tpi_kd_rebuild<C>(&cache, isrc, lane);
                            TPI_PROF(if (lane == src) pr_reb++;)
                            if (lane == src) { kd = tpi_kd_reload<C>(&cache, inst, kd_smd, kd_smi); kd.due = false; }
                        }
                    }
                    tt++;
                    TPI_PROF(pr_t_cold += clock64() - pr_tc;)
                }
            }
        }
        for (; tt < cnt; tt++)  // halted instance: the reference throws (ACME.jl:692); mark the rest
            static_for<0, NY>([&](auto kk) { out_at(tt * NY + decltype(kk)::value) = NAN; });
        // ---- output tile: one TMA store per warp (rows of inactive lanes and the columns past N
        //      lie outside the tensor and are clipped), or own-row stores
        if (NY > 0) {
            if (tma_out) fence_async_smem();  // generic-proxy writes -> visible to the async proxy
            else if (active) {
                if constexpr (SMAJ) {
                    for (int ts = 0; ts < cnt; ts++)
                        for (int q = 0; q < NY; q++) __stcs(yrow(n0 + ts) + q, out_at(ts * NY + q));
                } else {
                    for (int c = 0; c < cnt * NY; c++) __stcs(yrow(n0) + c, out_at(c));
                }
            }
        }
        if (tma_in || tma_out) __syncwarp();  // all lanes are done with in_tile / have written out_tile
        if (lane == 0 && wact) {
            if (tma_out) {
                if constexpr (SMAJ) tma_store_2d(&maps.y, w0 * NY, n0, smem_u32(out_tile));
                else tma_store_2d(&maps.y, n0 * NY, w0, smem_u32(out_tile));
                bulk_commit();
            }
            // ---- refill this input buffer with tile k+TPI_STAGES
            if (tma_in && k + TPI_STAGES < n_tiles) {
                mbar_arrive_expect_tx(bar0 + 8 * buf, SM::IN_TX);
                if constexpr (SMAJ) tma_load_2d(smem_u32(in_tile), &maps.u, w0 * NU, (k + TPI_STAGES) * T, bar0 + 8 * buf);
                else tma_load_2d(smem_u32(in_tile), &maps.u, (k + TPI_STAGES) * T * NU, w0, bar0 + 8 * buf);
            }
        }
    }
    if (tma_out && lane == 0) bulk_wait_all();
    __syncwarp();

    if (active) store_state();
    if constexpr (NN > 0)
        if (active) tpi_kd_flush<C>(cache, inst, kd);
#ifdef ACME_TPI_PROF
    if constexpr (NN > 0) {
        const long long pr_t1 = clock64();
        const int64_t gw = t / 32;
        const unsigned r_reo = __reduce_add_sync(0xffffffffu, pr.reo), r_tie = __reduce_add_sync(0xffffffffu, pr.tie),
                       r_scan = __reduce_add_sync(0xffffffffu, pr.scan), r_small = __reduce_add_sync(0xffffffffu, pr.small),
                       r_nt = __reduce_add_sync(0xffffffffu, pr.nontriv), r_st = __reduce_add_sync(0xffffffffu, pr.stores),
                       r_cold = __reduce_add_sync(0xffffffffu, pr_cold), r_reb = __reduce_add_sync(0xffffffffu, pr_reb);
        const unsigned r_mx = __reduce_max_sync(0xffffffffu, pr.nontriv);
        if (lane == 0 && gw < TPI_PROF_WARPS) {
            unsigned long long* r = g_tpi_prof + gw * TPI_PROF_REC;
            r[0] = (unsigned long long)(pr_t1 - pr_t0);
            r[1] = (unsigned long long)pr_t_cold;
            r[2] = ((unsigned long long)r_reo << 32) | r_tie;
            r[3] = ((unsigned long long)r_scan << 32) | r_small;
            r[4] = ((unsigned long long)r_nt << 32) | r_mx;
            r[5] = ((unsigned long long)r_cold << 32) | r_reb;
            r[6] = r_st;
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            r[7] = smid;
        }
    }
#endif
    // warp-reduce the counters, one set of atomics per warp
    const unsigned full_mask = 0xffffffffu;
    unsigned hsum[8];
    unsigned long long it_sum = 0;
#pragma unroll
    for (int b = 0; b < 8; b++) {
        hsum[b] = __reduce_add_sync(full_mask, hist_s[b * 32]);
        it_sum += (unsigned long long)hsum[b] * (b + 1);
    }
    const unsigned n_alive = active ? (unsigned)(dead_at < 0 ? n_samp : dead_at) : 0u;  // < 2^27
    const unsigned na = __reduce_add_sync(full_mask, n_alive);
    if (lane == 0) {
        if (na) {
            atomicAdd(&a.stats->samples, (unsigned long long)na);
            if (NN > 0) atomicAdd(&a.stats->solves, (unsigned long long)na);
        }
        if (it_sum) atomicAdd(&a.stats->newton_iters, it_sum);
#pragma unroll
        for (int b = 0; b < 8; b++)
            if (hsum[b]) atomicAdd(&a.stats->iter_hist[b], (unsigned long long)hsum[b]);
    }
}

}  // namespace acme
