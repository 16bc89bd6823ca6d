// Rows-in-registers kernel, grouped: G = 1 or 2 instances per warp, run in lockstep.
//
// Same data layout and arithmetic as kernel_rows.cuh (read its header first): lane r of an
// instance's group owns row r of the Newton system in registers, the LU factorisation with partial
// pivoting relabels rows instead of moving them, the pivot row travels through shared memory, the
// right-hand side rides along as an extra column.  With G = 2 a warp carries TWO instances, one per
// half-warp (nn, np, nx, nelem <= 16; a lane owns ceil(nq/16) components of q), and every warp
// instruction works for both: per instance the instruction count -- what bounds this
// latency-dominated kernel -- halves.
//
// Lockstep rules:
//  * Sub-warp masks are emulated by a slow loop on sm_100a (tools/lat_warp2.cu: a 16-lane ballot
//    costs 147 cycles, the full-warp one 23), so EVERY collective uses the full mask: shuffles with
//    width 16 and a per-group source lane, one REDUX per group with the other group's lanes
//    contributing the neutral element, ballots from which each group reads its own 16 bits.
//  * The two instances are at different places of the solver stack (one still iterating, one
//    converged; one re-originating from its cache, one in a homotopy step): every piece of the
//    solver is a block that the WARP enters when any group needs it (`__any_sync`, warp-uniform
//    branches) and whose effects are committed per lane under the group's own predicate.  The
//    per-group solver state (phase, iteration count, homotopy parameters) lives in registers,
//    uniform within a group.  Both groups start each sample together.
//  * Lane-conditional work is branch-free (clamped indices, predicated stores): a divergent branch
//    the compiler does not bracket with a reconvergence barrier splits the warp for good
//    (__syncwarp synchronises but does not re-merge), after which every collective takes its slow
//    divergent path.
//
// Solver stack, statement by statement: HomotopySolver (/root/reference/src/solvers.jl:268-296)
// { CachingSolver, learning per-instance cache (solvers.jl:347-396) { SimpleSolver (solvers.jl:207-236)
// with LinearSolver (solvers.jl:46-132) } } inside step! (src/ACME.jl:666-715).
#pragma once
#include "devmodel.h"
#include "elements.cuh"
#include "kernel_coop.cuh"     // CoopStatic: compile-time shape + the shared state layout
#include "kernel_generic.cuh"  // elem_eval
#include "kernel_rows.cuh"     // RowsSmem, Pack8, rcp_nobranch, rcp_slow, lds_vec, RowsProg
#include "tma.cuh"

namespace acme {

enum { RG_NEWTARGET = 0, RG_SOLVING = 1, RG_FINISHED = 2 };

template <int G>
struct RowsGrp {
    static constexpr int L = 32 / G;
    // per-group reduction of a value over the group's lanes, full-mask REDUX per group
    __device__ static __forceinline__ unsigned gmax(unsigned v, int grp) {
        if constexpr (G == 1) return __reduce_max_sync(ROWS_FULL, v);
        const unsigned m0 = __reduce_max_sync(ROWS_FULL, grp == 0 ? v : 0u);
        const unsigned m1 = __reduce_max_sync(ROWS_FULL, grp == 1 ? v : 0u);
        return grp ? m1 : m0;
    }
    __device__ static __forceinline__ unsigned gmin(unsigned v, int grp) {
        if constexpr (G == 1) return __reduce_min_sync(ROWS_FULL, v);
        const unsigned m0 = __reduce_min_sync(ROWS_FULL, grp == 0 ? v : 0xffffffffu);
        const unsigned m1 = __reduce_min_sync(ROWS_FULL, grp == 1 ? v : 0xffffffffu);
        return grp ? m1 : m0;
    }
    // does the predicate hold on every lane of this lane's group?
    __device__ static __forceinline__ bool gall(bool p, int grp) {
        const unsigned bal = __ballot_sync(ROWS_FULL, p);
        if constexpr (G == 1) return bal == ROWS_FULL;
        const unsigned gm = 0xffffu << (grp * 16);
        return (bal & gm) == gm;
    }
    __device__ static __forceinline__ double shfl(double v, int src) {
        const int hi = __shfl_sync(ROWS_FULL, __double2hiint(v), src, L), lo = __shfl_sync(ROWS_FULL, __double2loint(v), src, L);
        return __hiloint2double(hi, lo);
    }
};

// shared-memory rows of the pivot exchange per group: 2 buffers, a spare row for the lanes that do
// not own the pivot row, and a row of zeros that inactive lanes read
template <int NN>
struct RowsgProw {
    static constexpr int PITCH = rows_even(NN) + 2, DOUBLES = 4 * PITCH;
};

// setlhs! on rows-in-lanes (see rows_lu in kernel_rows.cuh), for the lanes with `active`; the other
// lanes (a group that is not factorising in this pass) keep A, b, pos, src, kpv.
template <int NN, int G>
__device__ __forceinline__ bool rowsg_lu(double (&A)[NN], double& b, int& pos, Pack8<NN>& src, Pack8<NN>& kpv, int gl,
                                         int grp, bool active, double* prow) {
    using GR = RowsGrp<G>;
    constexpr int NNP = rows_even(NN), PITCH = RowsgProw<NN>::PITCH;
    int npos = gl;
    Pack8<NN> nsrc, nkp;
    nsrc.clear();
    nkp.clear();
    bool ok = true;  // group-uniform
    const double* const rd_base = prow + (active ? 0 : 3 * PITCH);  // inactive lanes read zeros
    static_for<0, NN>([&](auto kk) {
        constexpr int k = decltype(kk)::value;
        constexpr int J0 = (k + 1) & ~1;  // first (even) column of the pivot row that is still needed
        const double a = A[k];
        const bool cand = active && gl < NN && npos >= k;
        const double inv_own = rcp_nobranch(a);  // speculative: overlaps the pivot search
        const bool usable = cand && a == a;      // abs(NaN) > amax is false: NaN never wins
        const unsigned hi = usable ? ((unsigned)__double2hiint(a) & 0x7fffffffu) : 0u;
        const unsigned lo = usable ? (unsigned)__double2loint(a) : 0u;
        const unsigned mh = GR::gmax(hi, grp);
        const bool c1 = cand && hi == mh;
        const unsigned ml = GR::gmax(c1 ? lo : 0u, grp);
        const bool c2 = c1 && lo == ml;
        // first strict maximum in position order = smallest position among the maxima
        const unsigned mk = GR::gmin(c2 ? (((unsigned)npos << 5) | (unsigned)gl) : 0xffffffffu, grp);
        const int kp = (int)(mk >> 5) & 31, s = (int)(mk & 31u);
        const bool mine = active && gl == s;
        // the pivot row (columns > k and the right-hand side) -> shared memory
        double* const dst = prow + (mine ? (k & 1) * PITCH : 2 * PITCH);
        static_for<J0 / 2, NNP / 2>([&](auto ii) {
            constexpr int j = 2 * decltype(ii)::value;
            reinterpret_cast<double2*>(dst)[j / 2] = make_double2(A[j], j + 1 < NN ? A[j + 1 < NN ? j + 1 : j] : 0.0);
        });
        dst[NNP] = b;
        // the pivot's magnitude is the search's maximum (mh:ml): its zero test (solvers.jl:70) and the
        // range test of the fast reciprocal need no further exchange
        double inv = GR::shfl(inv_own, s);
        const unsigned ex = mh >> 20;
        const bool special = active && (ex < 23u || ex > 2023u);
        if (__any_sync(ROWS_FULL, special)) {  // warp-uniform, practically never
            const double sl = rcp_slow(GR::shfl(a, s));
            inv = special ? sl : inv;
        }
        __syncwarp();
        double pr[NNP];
        static_for<J0 / 2, NNP / 2>([&](auto ii) {
            constexpr int j = 2 * decltype(ii)::value;
            const double2 v = reinterpret_cast<const double2*>(rd_base + (active ? (k & 1) * PITCH : 0))[j / 2];
            pr[j] = v.x;
            pr[j + 1] = v.y;
        });
        const double pb = rd_base[(active ? (k & 1) * PITCH : 0) + NNP];
        nkp.w[k >> 2] |= (ok ? (unsigned)kp : 0u) << ((k & 3) * 8);  // ipiv[k] is written before the zero test (solvers.jl:69)
        ok = ok && (mh | ml) != 0u;
        nsrc.w[k >> 2] |= (ok ? (unsigned)s : 0u) << ((k & 3) * 8);
        // the reference's row interchange, as a relabelling
        npos = !ok ? npos : (npos == k ? kp : (mine ? k : npos));
        const bool below = ok && cand && !mine;  // rows under the pivot row
        const double l = below ? a * inv : 0.0;
        A[k] = ok && mine ? inv : (below ? l : a);  // inverse pivot on the diagonal (solvers.jl:80)
        static_for<k + 1, NN>([&](auto jj) {
            constexpr int j = decltype(jj)::value;
            A[j] = __dsub_rn(A[j], __dmul_rn(l, pr[j]));  // not fused: exact zero pivots
        });
        b = __dsub_rn(b, __dmul_rn(l, pb));
    });
    pos = active ? npos : pos;
#pragma unroll
    for (int i = 0; i < (NN + 3) / 4; i++) {
        src.w[i] = active ? nsrc.w[i] : src.w[i];
        kpv.w[i] = active ? nkp.w[i] : kpv.w[i];
    }
    return ok;
}

// solve! on rows-in-lanes for the lanes with `need`; `fwd` lanes also run the forward substitution
// (the block is entered when any lane of the warp needs it)
template <int NN, int G>
__device__ __forceinline__ double rowsg_lusolve(const double (&A)[NN], int pos, const Pack8<NN>& src, double b, bool need,
                                                bool fwd, bool any_fwd, int gl) {
    using GR = RowsGrp<G>;
    const bool row = need && gl < NN;
    if (any_fwd) {  // warp-uniform
        const bool frow = row && fwd;
        static_for<0, NN - 1>([&](auto jj) {
            constexpr int j = decltype(jj)::value;
            const double xj = GR::shfl(b, (int)src.template get<j>());
            const double t = __dsub_rn(b, __dmul_rn(A[j], xj));
            b = (frow && pos > j) ? t : b;
        });
    }
    static_for<0, NN>([&](auto jr) {
        constexpr int j = NN - 1 - decltype(jr)::value;
        const double sc = A[j] * b;
        b = (row && pos == j) ? sc : b;
        if constexpr (j > 0) {
            const double xj = GR::shfl(b, (int)src.template get<j>());
            const double t = __dsub_rn(b, __dmul_rn(A[j], xj));
            b = (row && pos < j) ? t : b;
        }
    });
    return b;
}

template <class S, int G>
__host__ __device__ inline size_t rowsg_group_doubles(int nconst) {
    return (size_t)RowsSmem<S>::CONSTS + rows_even(nconst) + RowsgProw<S::NN>::DOUBLES;
}
template <class S, int G>
__host__ __device__ inline size_t rowsg_smem_bytes(int warps, int nconst) {
    return 16 + 8 * ((size_t)RowsSmem<S>::CTA_DOUBLES + (size_t)warps * G * rowsg_group_doubles<S, G>(nconst));
}

template <class S, int WARPS, int G, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) k_rowsg(const __grid_constant__ DevModel m, const RunArgs a) {
    using SM = RowsSmem<S>;
    using GR = RowsGrp<G>;
    constexpr int L = 32 / G;
    constexpr int NX = S::NX, NU = S::NU, NY = S::NY, NN = S::NN, NQ = S::NQ, NP = S::NP, NE = S::NE;
    constexpr int NNP = SM::NNP, NPP = SM::NPP;
    constexpr int QPL = (NQ + L - 1) / L;  // components of q per lane
    static_assert(NN <= L && NP <= L && NX <= L && NE <= L && NU <= L && NY <= L, "one lane per row");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* const cta = reinterpret_cast<double*>(smem_raw + 16);
    const uint32_t bar = smem_u32(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane & (L - 1), grp = lane / L;
    const int gdoubles = (int)rowsg_group_doubles<S, G>(m.nconst);
    double* const w = cta + SM::CTA_DOUBLES + (size_t)(warp * G + grp) * gdoubles;
    double* const prow = w + SM::CONSTS + rows_even(m.nconst);
    const double* const blob = cta;
    const double* const fqt = cta + SM::FQT;
    const double* const pexpt = cta + SM::PEXPT;

    // ---- shared model matrices: one TMA bulk copy per CTA, then q-major copies of fq and pexp
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)(SM::BLOB * 8);
        mbar_arrive_expect_tx(bar, bytes);
        bulk_g2s(smem_u32(cta), a.blob, bytes, bar);
    }
    mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < NQ * NNP; i += WARPS * 32) {
        const int q = i / NNP, c = i % NNP;
        cta[SM::FQT + i] = c < NN ? blob[S::O_FQ + c * NQ + q] : 0.0;
    }
    for (int i = threadIdx.x; i < NQ * NPP; i += WARPS * 32) {
        const int q = i / NPP, c = i % NPP;
        cta[SM::PEXPT + i] = c < NP ? blob[S::O_PEXP + c * NQ + q] : 0.0;
    }
    __syncthreads();

    const int64_t t0 = ((int64_t)blockIdx.x * WARPS + warp) * G;  // first launch-local instance of this warp
    if (t0 >= a.ninst) return;
    const bool act = t0 + grp < a.ninst;     // a trailing group without an instance shadows group 0 (nothing stored)
    const int64_t t = act ? t0 + grp : t0;
    const int64_t inst = a.inst0 + t;
    const int64_t ld = a.ld;
    double* const ws = a.ws + inst;
    auto WS = [&](int row) -> double& { return ws[(int64_t)row * ld]; };

    // this lane's rows for loads (clamped) and for multi-word stores (lanes without a row: spare row)
    int lq[QPL];
#pragma unroll
    for (int i = 0; i < QPL; i++) lq[i] = gl + i * L < NQ ? gl + i * L : NQ - 1;
    const int lp = gl < NP ? gl : NP - 1, lr = gl < NN ? gl : NN - 1;
    const int lx = gl < NX ? gl : (NX > 0 ? NX - 1 : 0), ly = gl < NY ? gl : (NY > 0 ? NY - 1 : 0);
    const int lu = gl < NU ? gl : (NU > 0 ? NU - 1 : 0);

    // ---- per-instance constants and persistent state -> shared memory / registers
    for (int k0 = 0; k0 < gdoubles; k0 += L)
        if (k0 + gl < gdoubles) w[k0 + gl] = 0.0;
    __syncwarp();
    for (int k0 = 0; k0 < m.nconst; k0 += L) {
        const int k = k0 + gl < m.nconst ? k0 + gl : m.nconst - 1;
        const double v = a.consts[(int64_t)k * ld + inst];
        if (k0 + gl < m.nconst) w[SM::CONSTS + k] = v;
    }
    {
        const double xv = NX > 0 ? WS(S::W_X + lx) : 0.0, pv = WS(S::W_LASTP + lp), zv = WS(S::W_LASTZ + lr);
        if (gl < NX) w[SM::X + gl] = xv;
        if (gl < NP) w[SM::LASTP + gl] = pv;
        if (gl < NN) w[SM::LASTZ + gl] = zv;
    }
    const int sel = (int)WS(S::W_SEL);
    const int LUB = sel ? S::W_LU1 : S::W_LU0, IPB = sel ? S::W_IPIV1 : S::W_IPIV0;
    {
        const int lw0 = gl < NN ? gl : NN;
        for (int j = 0; j < NN; j++) w[SM::LUO + lw0 * NNP + j] = WS(LUB + j * NN + lr);
        for (int j = 0; j < NP; j++) w[SM::JPO + lw0 * NPP + j] = WS(S::W_LASTJP + j * NN + lr);
    }
    // origin LU bookkeeping: rows sit at their positions; b must be permuted like solve! does
    int o_pos = gl, o_orig = gl;
    Pack8<NN> o_src, o_kp;
    o_src.clear();
    o_kp.clear();
    static_for<0, NN>([&](auto kk) {
        constexpr int k = decltype(kk)::value;
        const int kp = (int)WS(IPB + k);
        o_src.template put<k>((unsigned)k);
        o_kp.template put<k>((unsigned)kp);
        const int t1 = __shfl_sync(ROWS_FULL, o_orig, kp & (L - 1), L), t2 = __shfl_sync(ROWS_FULL, o_orig, k, L);
        o_orig = gl == k ? t1 : (gl == kp ? t2 : o_orig);
    });

    // element of this lane
    const DevElem& el = m.elems[gl < NE ? gl : 0];
    const int e_kind = gl < NE ? el.kind : -1, e_q = el.q_off, e_c = el.c_off, e_row = el.row, e_j = el.j_off;
    const int e_nn = gl < NE ? elem_nn(el.kind) : 0, e_nj = gl < NE ? elem_nj(el.kind) : 0;
    // row program of this lane
    RowsProg rp;
    rp.qi[0] = rp.qi[1] = rp.ji[0] = rp.ji[1] = 0u;
    {
        const RowProg& g = m.rows[lr];
        rp.n = gl < NN ? g.n : 0;
#pragma unroll
        for (int i = 0; i < ROWS_MAXT; i++) {
            const bool on = i < rp.n;
            rp.qi[i >> 2] |= (on ? (unsigned)g.q[i] : 0u) << ((i & 3) * 8);
            rp.ji[i >> 2] |= (on && g.jv[i] >= 0 ? (unsigned)(g.jv[i] + 1) : 0u) << ((i & 3) * 8);
            rp.c[i] = on && g.jv[i] < 0 ? g.c[i] : 0.f;
        }
    }
    const int maxterms = (int)__reduce_max_sync(ROWS_FULL, (unsigned)rp.n);
    // this lane's rows of fq (components of q = pfull + fq*z), constant for the whole run
    double fqrow[QPL][NN];
#pragma unroll
    for (int i = 0; i < QPL; i++)
        static_for<0, NN>([&](auto jj) { fqrow[i][decltype(jj)::value] = gl + i * L < NQ ? fqt[lq[i] * NNP + decltype(jj)::value] : 0.0; });

    const bool caching = m.solver == ACMEB200_SOLVER_HOMOTOPY_CACHING && m.subs[0].dyn_cap > 0;
    const int cap = m.subs[0].dyn_cap;
    double* const cps = caching ? m.subs[0].dyn_ps + inst * (int64_t)NP * cap : nullptr;
    double* const czs = caching ? m.subs[0].dyn_zs + inst * (int64_t)NN * cap : nullptr;
    int ncache = caching ? m.subs[0].dyn_n[inst] : 0;
    const double tol = m.tol;
    const int maxiter = m.maxiter;
    unsigned int* const hist_s = reinterpret_cast<unsigned int*>(w + SM::HIST);
    __syncwarp();

    // (row of Jq) * (q-major matrix T, row pitch PITCH doubles, 16-byte aligned rows) into out[0..COLS)
    auto row_times = [&](const double* T, auto colsC, auto pitchC, auto& out) {
        constexpr int COLS = decltype(colsC)::value, PITCH = decltype(pitchC)::value;
        static_for<0, COLS>([&](auto cc) { out[decltype(cc)::value] = 0.0; });
#pragma unroll
        for (int tI = 0; tI < ROWS_MAXT; tI++) {
            if (tI < maxterms) {  // warp-uniform
                const unsigned q = (rp.qi[tI >> 2] >> ((tI & 3) * 8)) & 0xffu;
                const unsigned j1 = (rp.ji[tI >> 2] >> ((tI & 3) * 8)) & 0xffu;
                const double jvv = w[SM::JV + (j1 ? j1 - 1 : 0)];
                const double coef = j1 ? jvv : (double)rp.c[tI];  // 0 beyond this row's terms
                double f[PITCH];
                lds_vec<PITCH>(T + q * PITCH, f);
                static_for<0, COLS>([&](auto cc) {
                    constexpr int c = decltype(cc)::value;
                    out[c] = fma(coef, f[c], out[c]);
                });
            }
        }
    };

    const double* const u = a.U + t * a.u_stride;
    double* const y = a.Y + t * a.y_stride;
    unsigned long long st_samples = 0, st_solves = 0, st_iters = 0, st_hom = 0, st_nc = 0;
    uint32_t status = a.status[inst];
    const int64_t N = a.N;
    bool alive = !(status & ACMEB200_STATUS_NONFINITE);  // group-uniform
    int64_t n_end = alive ? N : 0;                       // samples this instance produced
    double unext = (NU > 0 && N > 0) ? __ldg(u + lu) : 0.0;
    for (int64_t n = 0; n < N; n++) {
        if (!__any_sync(ROWS_FULL, alive)) break;
        // ---- step!  (ACME.jl:666-715)
        if (gl < NU) w[SM::U + gl] = unext;
        if (NU > 0) unext = __ldg(u + (n + 1 < N ? n + 1 : n) * NU + lu);  // prefetch the next sample
        __syncwarp();
        {  // p = dq*x + eq*u
            double acc = 0.0;
            static_for<0, NX>([&](auto jj) { constexpr int j = decltype(jj)::value; acc = fma(blob[S::O_DQ + j * NP + lp], w[SM::X + j], acc); });
            static_for<0, NU>([&](auto jj) { constexpr int j = decltype(jj)::value; acc = fma(blob[S::O_EQ + j * NP + lp], w[SM::U + j], acc); });
            if (gl < NP) w[SM::P + gl] = acc;
        }
        __syncwarp();

        // ---- solve(::HomotopySolver, p)  (solvers.jl:268-296), per group
        bool hom = false, conv = false, used_h = false;
        double ha = 1.0, best_a = 0.0;
        int total = 0, iters = 0;
        int stage = alive ? RG_NEWTARGET : RG_FINISHED;
        int phase = ROWS_PH_START;
        double pfull[QPL];
#pragma unroll
        for (int i = 0; i < QPL; i++) pfull[i] = 0.0;
        while (__any_sync(ROWS_FULL, stage != RG_FINISHED)) {
            const double* const ptar = w + (hom ? SM::PA : SM::P);
            // ---- (a) a new target p for the base solver: solve(::CachingSolver, p) picks the start point
            //          (solvers.jl:347-373)
            const bool fresh = stage == RG_NEWTARGET;
            if (__any_sync(ROWS_FULL, fresh)) {
                phase = fresh ? ROWS_PH_START : phase;
                iters = fresh ? 0 : iters;
                conv = fresh ? false : conv;
                if (caching) {
                    double best = 0.0;
                    static_for<0, NP>([&](auto ii) {
                        constexpr int i = decltype(ii)::value;
                        const double d = ptar[i] - w[SM::LASTP + i];
                        best = fma(d, d, best);
                    });
                    double lb = __longlong_as_double(0x7ff0000000000000ll);
                    int li = 0x7fffffff;
                    const int nscan = fresh ? ncache : 0;
                    for (int rd = 0; __any_sync(ROWS_FULL, rd * L < nscan); rd++) {
                        const int idx = rd * L + gl;
                        const int ic = idx < nscan ? idx : (nscan > 0 ? nscan - 1 : 0);
                        double d2 = 0.0;
                        static_for<0, NP>([&](auto dd) {
                            constexpr int d = decltype(dd)::value;
                            const double df = cps[(int64_t)d * cap + ic] - ptar[d];
                            d2 = fma(df, df, d2);
                        });
                        const bool better = idx < nscan && d2 < lb;
                        lb = better ? d2 : lb;
                        li = better ? idx : li;
                    }
                    // exact minimum over the group (distances are non-negative: bit patterns order like values);
                    // ties: the older stored point wins, and the current origin wins over stored points
                    const unsigned bh = (unsigned)__double2hiint(lb), bl = (unsigned)__double2loint(lb);
                    const unsigned mh = GR::gmin(bh, grp);
                    const bool c1 = bh == mh;
                    const unsigned ml = GR::gmin(c1 ? bl : 0xffffffffu, grp);
                    const bool c2 = c1 && bl == ml;
                    const int mi = (int)GR::gmin(c2 ? (unsigned)li : 0x7fffffffu, grp);
                    const double mb = __hiloint2double((int)mh, (int)ml);
                    const bool take = fresh && mi != 0x7fffffff && mb < best;  // group-uniform
                    if (__any_sync(ROWS_FULL, take)) {
                        const int mic = take ? mi : 0;
                        const double cpv = cps[(int64_t)lp * cap + mic], czv = czs[(int64_t)lr * cap + mic];
                        if (take && gl < NP) w[SM::CP + gl] = cpv;
                        if (take && gl < NN) w[SM::Z + gl] = czv;
                        __syncwarp();
                    }
                    phase = take ? ROWS_PH_ORIGIN : phase;
                }
                stage = fresh ? RG_SOLVING : stage;
            }
            const bool solving = stage == RG_SOLVING;

            // ---- (b) set_p!: pfull = q0 + pexp*p  (ACME.jl:237-243), p = target or cached point
            const bool do_setp = solving && phase != ROWS_PH_NEWTON;
            if (__any_sync(ROWS_FULL, do_setp)) {
                const double* const pp = phase == ROWS_PH_START ? ptar : w + SM::CP;
#pragma unroll
                for (int i = 0; i < QPL; i++) {
                    double acc = blob[S::O_Q0 + lq[i]];
                    static_for<0, NP>([&](auto jj) { constexpr int j = decltype(jj)::value; acc = fma(pexpt[lq[i] * NPP + j], pp[j], acc); });
                    pfull[i] = do_setp ? acc : pfull[i];
                }
            }

            double A[NN];
            double rhs = 0.0;
            int pos = gl;
            Pack8<NN> src, kpv;
            src.clear();
            kpv.clear();
            // ---- (c) start of a SimpleSolver solve: z0 = last_z - last_LU \ (last_Jp*(p - last_p))  (solvers.jl:209-215)
            const bool do_start = solving && phase == ROWS_PH_START;
            if (__any_sync(ROWS_FULL, do_start)) {
                const double dpv = ptar[lp] - w[SM::LASTP + lp];
                if (do_start && gl < NP) w[SM::DP + gl] = dpv;
                __syncwarp();
                double r0 = 0.0;
                static_for<0, NP>([&](auto jj) { constexpr int j = decltype(jj)::value; r0 = fma(w[SM::JPO + lr * NPP + j], w[SM::DP + j], r0); });
                static_for<0, NN>([&](auto jj) { A[decltype(jj)::value] = w[SM::LUO + lr * NNP + decltype(jj)::value]; });
                pos = o_pos;
                src = o_src;
                rhs = GR::shfl(r0, o_orig);
            } else {
                static_for<0, NN>([&](auto jj) { A[decltype(jj)::value] = 0.0; });
            }

            // ---- (d) evaluate! (ACME.jl:178-188) and setlhs! for the groups in a Newton / origin pass
            const bool do_eval = solving && phase != ROWS_PH_START;
            bool all_fin = true, all_small = false, all_jfin = true, ok = true;
            if (__any_sync(ROWS_FULL, do_eval)) {
                iters += (do_eval && phase == ROWS_PH_NEWTON) ? 1 : 0;
                {
                    double zr[NNP];
                    lds_vec<NNP>(w + SM::Z, zr);
#pragma unroll
                    for (int i = 0; i < QPL; i++) {
                        double acc = pfull[i];
                        static_for<0, NN>([&](auto jj) { constexpr int j = decltype(jj)::value; acc = fma(fqrow[i][j], zr[j], acc); });
                        if (do_eval && gl + i * L < NQ) w[SM::Q + gl + i * L] = acc;
                    }
                }
                __syncwarp();
                if (do_eval && gl < NE) {
                    double res[2], jv[4];
                    elem_eval(e_kind, w + SM::CONSTS + e_c, w + SM::Q + e_q, res, jv);
                    for (int k = 0; k < e_nj; k++) w[SM::JV + e_j + k] = jv[k];
                    for (int r = 0; r < e_nn; r++) w[SM::RES + e_row + r] = res[r];
                }
                __syncwarp();
                const double re = w[SM::RES + lr];
                const double ar = fabs(re);
                const bool fin = gl >= NN || ar <= 1.7976931348623157e308;
                const bool small = gl >= NN || ar < tol;
                double Ae[NN];
                row_times(fqt, IC<NN>{}, IC<NNP>{}, Ae);  // lanes without a row have an empty program: zeros
                unsigned mx = 0u;
                static_for<0, NN>([&](auto jj) {
                    const unsigned h = (unsigned)__double2hiint(Ae[decltype(jj)::value]) & 0x7fffffffu;
                    mx = h > mx ? h : mx;
                });
                const bool jfin = mx < 0x7ff00000u;
                all_fin = GR::gall(fin, grp);
                all_small = GR::gall(small, grp);
                all_jfin = GR::gall(jfin, grp);
                static_for<0, NN>([&](auto jj) { A[decltype(jj)::value] = do_eval ? Ae[decltype(jj)::value] : A[decltype(jj)::value]; });
                rhs = do_eval ? re : rhs;
                // the reference tests finiteness before factorising (solvers.jl:220); the factorisation of a
                // non-finite matrix is simply not used here
                ok = rowsg_lu<NN, G>(A, rhs, pos, src, kpv, gl, grp, do_eval, prow);
            }
            const bool ev_origin = do_eval && phase == ROWS_PH_ORIGIN;
            const bool ev_newton = do_eval && phase == ROWS_PH_NEWTON;
            const bool failed = ev_newton && (!(all_fin && all_jfin) || !ok);   // solvers.jl:220-225
            const bool done_conv = ev_newton && !failed && all_small;            // solvers.jl:226
            // ---- (e) (p, z) with this factorisation becomes the extrapolation origin (solvers.jl:190-196, 231-234)
            const bool to_origin = ev_origin || done_conv;
            if (__any_sync(ROWS_FULL, to_origin)) {
                const double* const psrc = phase == ROWS_PH_ORIGIN ? w + SM::CP : ptar;
                double jp[NP];
                row_times(pexpt, IC<NP>{}, IC<NPP>{}, jp);  // calc_Jp!: Jp = Jq*pexp  (ACME.jl:246-251)
                const int lw = (to_origin && gl < NN) ? gl : NN;
                static_for<0, NP>([&](auto jj) { w[SM::JPO + lw * NPP + decltype(jj)::value] = jp[decltype(jj)::value]; });
                static_for<0, NN>([&](auto jj) { w[SM::LUO + lw * NNP + decltype(jj)::value] = A[decltype(jj)::value]; });
                const double zv = w[SM::Z + lr], pv = psrc[lp];
                if (to_origin && gl < NN) w[SM::LASTZ + gl] = zv;
                if (to_origin && gl < NP) w[SM::LASTP + gl] = pv;
                o_pos = to_origin ? pos : o_pos;
                o_orig = to_origin ? gl : o_orig;
#pragma unroll
                for (int i = 0; i < (NN + 3) / 4; i++) {
                    o_src.w[i] = to_origin ? src.w[i] : o_src.w[i];
                    o_kp.w[i] = to_origin ? kpv.w[i] : o_kp.w[i];
                }
                __syncwarp();
            }
            conv = failed ? (all_fin && all_small) : (done_conv ? true : conv);
            bool solve_done = failed || done_conv;
            const bool need_ls = do_start || (ev_newton && !failed && !done_conv);
            // ---- (f) solve!: forward (start pass only) and back substitution, z update
            if (__any_sync(ROWS_FULL, need_ls)) {
                const bool fwd = do_start;
                const double xs = rowsg_lusolve<NN, G>(A, pos, src, rhs, need_ls, fwd, __any_sync(ROWS_FULL, fwd), gl);
                const int zp = (need_ls && gl < NN) ? pos : 0;
                const double base = phase == ROWS_PH_START ? w[SM::LASTZ + zp] : w[SM::Z + zp];
                if (need_ls && gl < NN) w[SM::Z + zp] = base - xs;
                __syncwarp();
            }
            const bool hitmax = need_ls && phase == ROWS_PH_NEWTON && iters >= maxiter;  // 500 updates without convergence
            solve_done = solve_done || hitmax;
            phase = ev_origin ? ROWS_PH_START : ((need_ls && !hitmax) ? ROWS_PH_NEWTON : phase);

            // ---- (g) a base solve has ended: cache bookkeeping (solvers.jl:374-386), homotopy control (solvers.jl:270-295)
            if (__any_sync(ROWS_FULL, solve_done)) {
                total += solve_done ? iters : 0;
                const bool append = solve_done && caching && iters > 5 && conv && ncache < cap;
                if (__any_sync(ROWS_FULL, append)) {
                    const double pv = ptar[lp], zv = w[SM::Z + lr];
                    const int slot = append ? ncache : 0;
                    if (append && act && gl < NP) cps[(int64_t)gl * cap + slot] = pv;
                    if (append && act && gl < NN) czs[(int64_t)gl * cap + slot] = zv;
                    ncache += append ? 1 : 0;
                    __threadfence_block();
                    __syncwarp();
                }
                const bool first = solve_done && !hom;
                bool finish = first && (conv || m.solver == ACMEB200_SOLVER_SIMPLE);
                const bool start_h = first && !finish;
                const bool cont = solve_done && hom;
                const bool c_conv = cont && conv, c_fail = cont && !conv;
                const double new_a = (ha + best_a) / 2;
                const bool give_up = c_fail && !(best_a < new_a && new_a < ha);  // no float between best_a and a
                const bool reached = c_conv && !(ha < 1);                          // best_a becomes a = 1: done
                best_a = c_conv ? ha : (start_h ? 0.0 : best_a);
                ha = c_conv ? 1.0 : ((c_fail && !give_up) ? new_a : (start_h ? 0.5 : ha));
                finish = finish || give_up || reached;
                if (__any_sync(ROWS_FULL, start_h)) {
                    const double sp = w[SM::LASTP + lp];
                    if (start_h && gl < NP) w[SM::STARTP + gl] = sp;
                }
                hom = hom || start_h;
                used_h = used_h || start_h;
                const bool retarget = solve_done && !finish;
                if (__any_sync(ROWS_FULL, retarget)) {
                    __syncwarp();
                    double pa = w[SM::STARTP + lp];
                    pa *= (1 - ha);
                    pa += ha * w[SM::P + lp];
                    if (retarget && gl < NP) w[SM::PA + gl] = pa;
                    __syncwarp();
                }
                stage = finish ? RG_FINISHED : (retarget ? RG_NEWTARGET : stage);
            }
        }

        // ---- bookkeeping of step! (ACME.jl:688-694)
        if (alive) {
            st_solves++;
            st_iters += (unsigned)total;
            st_hom += used_h ? 1u : 0u;
        }
        {
            int bin = total < 1 ? 1 : total;
            if (bin > ACMEB200_HIST_BINS) bin = ACMEB200_HIST_BINS;
            if (alive && gl == 0) hist_s[bin - 1] += 1u;
        }
        const bool trouble = alive && !conv;
        if (__any_sync(ROWS_FULL, trouble)) {
            if (trouble && act && gl == 0 && a.first_fail[inst] < 0) a.first_fail[inst] = a.n_done + n;
            const bool zfin = gl >= NN || isfinite(w[SM::Z + lr]);
            const bool all_zfin = GR::gall(zfin, grp);
            const bool warn = trouble && all_zfin, fatal = trouble && !all_zfin;
            status |= warn ? ACMEB200_STATUS_NOT_CONVERGED : 0u;
            status |= fatal ? ACMEB200_STATUS_NONFINITE : 0u;
            st_nc += warn ? 1u : 0u;
            n_end = fatal ? n : n_end;  // the reference throws here (ACME.jl:692)
            alive = alive && !fatal;
        }
        // ---- y = y0 + dy*x + ey*u + fy*z ; x = x0 + a*x + b*u + c*z  (ACME.jl:699-714)
        double yv, xn;
        {
            double acc = blob[S::O_Y0 + ly];
            static_for<0, NX>([&](auto jj) { constexpr int j = decltype(jj)::value; acc = fma(blob[S::O_DY + j * NY + ly], w[SM::X + j], acc); });
            static_for<0, NU>([&](auto jj) { constexpr int j = decltype(jj)::value; acc = fma(blob[S::O_EY + j * NY + ly], w[SM::U + j], acc); });
            static_for<0, NN>([&](auto jj) { constexpr int j = decltype(jj)::value; acc = fma(blob[S::O_FY + j * NY + ly], w[SM::Z + j], acc); });
            yv = acc;
        }
        {
            double acc = blob[S::O_X0 + lx];
            static_for<0, NX>([&](auto jj) { constexpr int j = decltype(jj)::value; acc = fma(blob[S::O_A + j * NX + lx], w[SM::X + j], acc); });
            static_for<0, NU>([&](auto jj) { constexpr int j = decltype(jj)::value; acc = fma(blob[S::O_B + j * NX + lx], w[SM::U + j], acc); });
            static_for<0, NN>([&](auto jj) { constexpr int j = decltype(jj)::value; acc = fma(blob[S::O_C + j * NX + lx], w[SM::Z + j], acc); });
            xn = acc;
        }
        __syncwarp();
        if (alive && gl < NX) w[SM::X + gl] = xn;
        if (alive && act && gl < NY) y[n * NY + gl] = yv;
        st_samples += alive ? 1u : 0u;
    }
    if (act && gl < NY)
        for (int64_t n = n_end; n < N; n++) y[n * NY + gl] = NAN;  // samples after a fatal failure
    __syncwarp();

    // ---- persistent state back to the generic layout: rows at their positions + ipiv
    if (act) {
        if (gl < NX) WS(S::W_X + gl) = w[SM::X + gl];
        if (gl < NP) WS(S::W_LASTP + gl) = w[SM::LASTP + gl];
        if (gl < NN) {
            WS(S::W_LASTZ + gl) = w[SM::LASTZ + gl];
            for (int j = 0; j < NN; j++) WS(LUB + j * NN + o_pos) = w[SM::LUO + gl * NNP + j];
            for (int j = 0; j < NP; j++) WS(S::W_LASTJP + j * NN + gl) = w[SM::JPO + gl * NPP + j];
        }
        if (gl == 0) {
            static_for<0, NN>([&](auto kk) { WS(IPB + decltype(kk)::value) = (double)o_kp.template get<decltype(kk)::value>(); });
            if (caching) m.subs[0].dyn_n[inst] = ncache;
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
