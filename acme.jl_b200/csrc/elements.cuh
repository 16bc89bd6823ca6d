// Non-linear element laws as __device__ functions.
//
// The reference stores these as Julia closures q -> (res, J)
// (/root/reference/src/elements.jl:25-30 pot, :107-129 Jiles-Atherton,
// :238-244 diode, :323-401 BJT, :453-479 MOSFET, :540-546 tanh op-amp).
// Here every element type E provides
//   NN, NQ        rows / columns of its Jacobian block
//   NPAR          raw parameters (layout in include/acmeb200.h)
//   NC            per-instance constants derived once from the parameters
//                 (exactly the sub-expressions the reference's closure captures)
//   NJ            number of Jacobian entries that are not compile-time constants
//   prep(P, C)    parameters -> constants
//   eval(C, q, res, jv)   residuals and the variable Jacobian entries
//   row<R>(jv, m)         (row R of the Jacobian block) . (m(0..NQ-1)), used for
//                         J = Jq*fq (ACME.jl:186) and Jp = Jq*pexp (ACME.jl:249)
//                         without multiplying by the structural zeros / +-1
#pragma once
#include <cmath>
#include "../../include/acmeb200.h"

namespace acme {

#define ACME_DI __host__ __device__ __forceinline__

// exp(double) for the element laws, built for the latency-bound Newton loops: the FP64 pipe issues
// one warp instruction per 2 cycles and a dependent DFMA costs 8, so both the NUMBER of FP64
// instructions and the depth of their chain matter.  Table-driven:
//   x = (64 m + j) ln2/64 + r,  |r| <= ln2/128,   exp(x) = 2^m * 2^(j/64) * (1 + r + r^2/2 + ... + r^5/120)
// Cody-Waite reduction (2^52+2^51 rounding trick, ln2/64 split in two for the FMA), a 64-entry table
// of correctly rounded 2^(j/64) read through the read-only path (512 B, L1-resident), a degree-5
// polynomial in Estrin form (3 dependent levels; truncation error r^6/720 < 0.17 ulp), exponent
// splice in two steps (exact for normal results, one rounding for subnormal ones, overflow to +Inf
// in the multiply), no branch.  12 FP64 instructions instead of the library's ~30, <= 1 ulp from
// the correctly rounded result (tools/exp_test.py measures the distance to the library exp).
#ifdef __CUDACC__
static __device__ const double ACME_EXP_T64[64] = {
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0};
// 64/ln2, 2^52+2^51, -ln2_hi/64, -ln2_lo/64, 1/2, 1/6, 1/24, 1/120
#define ACME_EXP_CONSTANTS                                                                              \
    0x1.71547652b82fep+6, 0x1.8000000000000p+52, -0x1.62e42fefa39efp-7, -0x1.abc9e3b39803fp-62, 0x1.0p-1, \
        0x1.5555555555555p-3, 0x1.5555555555555p-5, 0x1.1111111111111p-7, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0
static __constant__ double ACME_EXPC[16] = {ACME_EXP_CONSTANTS};

__device__ __forceinline__ double acme_exp(double x, const double* __restrict__ K) {
    // Branch-free so that the exps of neighbouring elements interleave in one basic block
    double t = fma(x, K[0], K[1]);
    const int i = __double2loint(t);  // 64 m + j
    t = t - K[1];
    double r = fma(t, K[2], x);
    r = fma(t, K[3], r);
    const double tj = __ldg(&ACME_EXP_T64[i & 63]);
    const double r2 = r * r;
    const double a = fma(r, K[5], K[4]);  // 1/2 + r/6
    const double b = fma(r, K[7], K[6]);  // 1/24 + r/120
    const double q = fma(b, r2, a);
    const double p = fma(q, r2, r);       // exp(r) - 1
    const double v = fma(tj, p, tj);      // 2^(j/64) exp(r), in [1, 2)
    // scale by 2^m in two steps (m = k + (m - k)): exact for normal results, one rounding for
    // subnormal ones, overflow to +Inf happens naturally in the multiply
    const int m = i >> 6;
    const int k = m >> 1;
    const double s1 = __hiloint2double(__double2hiint(v) + (k << 20), __double2loint(v));
    const double sc = __hiloint2double((0x3ff + (m - k)) << 20, 0);
    double res = s1 * sc;
    const int hx = __double2hiint(x) & 0x7fffffff;
    // |x| >= 745 (or NaN): i is meaningless; exp is +Inf / 0 / NaN there
    // +Inf for large x, NaN for NaN (NaN*Inf), 0 for very negative x -- no branch
    const double far = x < 0 ? 0.0 : x * __longlong_as_double(0x7ff0000000000000ll);
    res = hx >= 0x40874800 ? far : res;
    return res;
}
#endif
// host passes of the __host__ __device__ element code never evaluate laws (the host only runs prep)
#ifdef __CUDA_ARCH__
#define ACME_EXPD(x) acme_exp(x, K)
#else
#define ACME_EXPD(x) exp(x)
#endif
// the table as a plain struct so kernels can also receive it as a __grid_constant__ parameter
// (param-space constants are eligible for uniform-register operands)
struct ExpTable { double k[16]; };
inline ExpTable make_exp_table() {
    return ExpTable{{0x1.71547652b82fep+6, 0x1.8000000000000p+52, -0x1.62e42fefa39efp-7, -0x1.abc9e3b39803fp-62, 0x1.0p-1,
                     0x1.5555555555555p-3, 0x1.5555555555555p-5, 0x1.1111111111111p-7, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0}};
}

struct Diode {  // elements.jl:236-245
    static constexpr int KIND = ACMEB200_ELEM_DIODE, NN = 1, NQ = 2, NPAR = 2, NC = 3, NJ = 1;
    ACME_DI static void prep(const double* P, double* C) {
        const double is = P[0], eta = P[1];
        C[0] = is;
        C[1] = 1 / (25e-3 * eta);
        C[2] = is / (25e-3 * eta);
    }
    ACME_DI static void eval(const double* C, const double* q, double* res, double* jv, const double* K = nullptr) {
        const double ex = ACME_EXPD(q[0] * C[1]);
        res[0] = C[0] * (ex - 1) - q[1];
        jv[0] = C[2] * ex;
    }
    template <int R, class F> ACME_DI static double row(const double* jv, F m) {
        return fma(jv[0], m(0), -m(1));
    }
};

struct Pot {  // elements.jl:20-31
    static constexpr int KIND = ACMEB200_ELEM_POT, NN = 2, NQ = 5, NPAR = 1, NC = 1, NJ = 4;
    ACME_DI static void prep(const double* P, double* C) { C[0] = P[0]; }
    ACME_DI static void eval(const double* C, const double* q, double* res, double* jv, const double* K = nullptr) {
        const double r = C[0];
        const double v1 = q[0], v2 = q[1], i1 = q[2], i2 = q[3], pos = q[4];
        res[0] = v1 - r * pos * i1;
        res[1] = v2 - r * (1 - pos) * i2;
        jv[0] = -r * pos;
        jv[1] = -r * i1;
        jv[2] = -r * (1 - pos);
        jv[3] = -r * i2;
    }
    template <int R, class F> ACME_DI static double row(const double* jv, F m) {
        if constexpr (R == 0) return fma(jv[1], m(4), fma(jv[0], m(2), m(0)));
        return fma(jv[3], m(4), fma(jv[2], m(3), m(1)));
    }
};

struct OpampTanh {  // elements.jl:536-551
    static constexpr int KIND = ACMEB200_ELEM_OPAMP_TANH, NN = 1, NQ = 2, NPAR = 2, NC = 3, NJ = 1;
    ACME_DI static void prep(const double* P, double* C) {
        C[0] = P[0];         // gain
        C[1] = P[1];         // scale
        C[2] = P[0] / P[1];  // gain/scale
    }
    ACME_DI static void eval(const double* C, const double* q, double* res, double* jv, const double* K = nullptr) {
        const double vs = q[0] * C[2];
        const double ch = cosh(vs);
        res[0] = tanh(vs) * C[1] - q[1];
        jv[0] = C[0] / (ch * ch);
    }
    template <int R, class F> ACME_DI static double row(const double* jv, F m) {
        return fma(jv[0], m(0), -m(1));
    }
};

struct TestQuad {  // test/runtests.jl:207-219
    static constexpr int KIND = ACMEB200_ELEM_TEST_QUAD, NN = 1, NQ = 2, NPAR = 0, NC = 1, NJ = 1;
    ACME_DI static void prep(const double*, double* C) { C[0] = 0; }
    ACME_DI static void eval(const double*, const double* q, double* res, double* jv, const double* K = nullptr) {
        res[0] = q[0] * q[0] - 1 + q[1];
        jv[0] = 2 * q[0];
    }
    template <int R, class F> ACME_DI static double row(const double* jv, F m) {
        return fma(jv[0], m(0), m(1));
    }
};

struct Bjt {  // elements.jl:309-406
    static constexpr int KIND = ACMEB200_ELEM_BJT, NN = 2, NQ = 4, NPAR = 14, NC = 20, NJ = 4;
    // flags packed in C[19]: bit0 early, bit1 knee, bit2 ile!=0, bit3 own exp for the
    // emitter leakage, bit4 ilc!=0, bit5 own exp for the collector leakage
    ACME_DI static void prep(const double* P, double* C) {
        const double ise = P[0], isc = P[1], ne = P[2], nc = P[3], bf = P[4], br = P[5];
        const double ile = P[6], ilc = P[7], nel = P[8], ncl = P[9];
        const double vaf = P[10], var = P[11], ikf = P[12], ikr = P[13];
        C[0] = 1 / (25e-3 * ne);
        C[1] = 1 / (25e-3 * nc);
        C[2] = bf / (1 + bf) * ise;
        C[3] = br / (1 + br) * isc;
        C[4] = bf / (1 + bf) * ise / (25e-3 * ne);
        C[5] = br / (1 + br) * isc / (25e-3 * nc);
        C[6] = 1 / bf;
        C[7] = 1 / br;
        C[8] = ile;
        C[9] = ilc;
        C[10] = 1 / (25e-3 * nel);
        C[11] = 1 / (25e-3 * ncl);
        C[12] = ile / (25e-3 * ne);
        C[13] = ilc / (25e-3 * nc);
        C[14] = 1 / var;
        C[15] = 1 / vaf;
        C[16] = 1 / ikf;
        C[17] = 1 / ikr;
        C[18] = 0;
        int flags = 0;
        if (!(var == INFINITY && vaf == INFINITY)) flags |= 1;
        if (!(ikf == INFINITY && ikr == INFINITY)) flags |= 2;
        if (ile != 0) flags |= 4;
        if (nel != ne) flags |= 8;
        if (ilc != 0) flags |= 16;
        if (ncl != nc) flags |= 32;
        C[19] = (double)flags;
    }
    ACME_DI static void eval(const double* C, const double* q, double* res, double* jv, const double* K = nullptr) {
        const int flags = (int)C[19];
        const double vE = q[0], vC = q[1], iE = q[2], iC = q[3];
        const double expE = ACME_EXPD(vE * C[0]);
        const double expC = ACME_EXPD(vC * C[1]);
        const double i_f = C[2] * (expE - 1);
        const double i_r = C[3] * (expC - 1);
        const double di_f1 = C[4] * expE;
        const double di_r2 = C[5] * expC;
        double i_cc, di_cc1, di_cc2;
        if ((flags & 3) == 0) {
            i_cc = i_f - i_r;
            di_cc1 = di_f1;
            di_cc2 = -di_r2;
        } else if ((flags & 3) == 1) {
            const double q1 = 1 - vE * C[14] - vC * C[15];
            i_cc = q1 * (i_f - i_r);
            di_cc1 = (-C[14]) * (i_f - i_r) + q1 * di_f1;
            di_cc2 = (-C[15]) * (i_f - i_r) - q1 * di_r2;
        } else if ((flags & 3) == 2) {
            const double q2 = i_f * C[16] + i_r * C[17];
            const double qden = 1 + sqrt(1 + 4 * q2);
            const double qfact = 2 / qden;
            i_cc = qfact * (i_f - i_r);
            const double dq21 = di_f1 * C[16], dq22 = di_r2 * C[17];
            const double dqf1 = -4 * dq21 / (qden - 1) / (qden * qden);
            const double dqf2 = -4 * dq22 / (qden - 1) / (qden * qden);
            di_cc1 = dqf1 * (i_f - i_r) + qfact * di_f1;
            di_cc2 = dqf2 * (i_f - i_r) - qfact * di_r2;
        } else {
            const double q1 = 1 - vE * C[14] - vC * C[15];
            const double q2 = i_f * C[16] + i_r * C[17];
            const double qden = 1 + sqrt(1 + 4 * q2);
            const double qfact = 2 * q1 / qden;
            i_cc = qfact * (i_f - i_r);
            const double dq21 = di_f1 * C[16], dq22 = di_r2 * C[17];
            const double dqf1 = (2 * (-C[14]) * qden - q1 * 4 * dq21 / (qden - 1)) / (qden * qden);
            const double dqf2 = (2 * (-C[15]) * qden - q1 * 4 * dq22 / (qden - 1)) / (qden * qden);
            di_cc1 = dqf1 * (i_f - i_r) + qfact * di_f1;
            di_cc2 = dqf2 * (i_f - i_r) - qfact * di_r2;
        }
        double iBE = C[6] * i_f, diBE1 = C[6] * di_f1;
        if (flags & 4) {
            const double expEl = (flags & 8) ? ACME_EXPD(vE * C[10]) : expE;
            iBE += C[8] * (expEl - 1);
            diBE1 += C[12] * expEl;
        }
        double iBC = C[7] * i_r, diBC2 = C[7] * di_r2;
        if (flags & 16) {
            const double expCl = (flags & 32) ? ACME_EXPD(vC * C[11]) : expC;
            iBC += C[9] * (expCl - 1);
            diBC2 += C[13] * expCl;
        }
        res[0] = i_cc + iBE - iE;
        res[1] = -i_cc + iBC - iC;
        jv[0] = di_cc1 + diBE1;
        jv[1] = di_cc2;
        jv[2] = -di_cc1;
        jv[3] = -di_cc2 + diBC2;
    }
    template <int R, class F> ACME_DI static double row(const double* jv, F m) {
        if constexpr (R == 0) return fma(jv[1], m(1), fma(jv[0], m(0), -m(2)));
        return fma(jv[3], m(1), fma(jv[2], m(0), -m(3)));
    }
};

struct Mosfet {  // elements.jl:436-481
    static constexpr int KIND = ACMEB200_ELEM_MOSFET, NN = 1, NQ = 3, NPAR = 12, NC = 12, NJ = 2;
    ACME_DI static void prep(const double* P, double* C) {
        for (int i = 0; i < 12; i++) C[i] = P[i];
    }
    ACME_DI static double poly(double x, const double* c, int n) {  // Horner (Base.evalpoly)
        double acc = c[n - 1];
        for (int i = n - 2; i >= 0; i--) acc = fma(x, acc, c[i]);
        return acc;
    }
    ACME_DI static void eval(const double* C, const double* q, double* res, double* jv, const double* K = nullptr) {
        const double pol = C[0], lam = C[1];
        const int nvt = (int)C[2], nal = (int)C[3];
        const double *vt = C + 4, *al = C + 8;
        double dvt[3] = {0, 0, 0}, dal[3] = {0, 0, 0};
        for (int k = 1; k < 4; k++) {
            if (k < nvt) dvt[k - 1] = vt[k] * k;
            if (k < nal) dal[k - 1] = al[k] * k;
        }
        const double vgs = q[0], vds = q[1], id = q[2];
        const double a_ = poly(pol * vgs, al, nal);
        const double da = nal > 1 ? poly(pol * vgs, dal, nal - 1) : 0;
        const double vt_ = poly(pol * vgs, vt, nvt);
        const double dvt_ = nvt > 1 ? poly(pol * vgs, dvt, nvt - 1) : 0;
        const double lam_ = vds >= 0 ? lam : 0.0;
        if (vgs <= vt_) {
            res[0] = -id;
            jv[0] = 0.0;
            jv[1] = 0.0;
        } else if (vds <= vgs - vt_) {
            res[0] = a_ * (vgs - vt_ - 0.5 * vds) * vds * (1 + lam_ * vds) - id;
            jv[0] = a_ * (1 - dvt_) * vds * (1 + lam_ * vds) +
                    da * (vgs - vt_ - 0.5 * vds) * vds * (1 + lam_ * vds);
            jv[1] = a_ * (vgs - vt_ + vds * (2 * lam_ * (vgs - vt_ - 0.75 * vds) - 1));
        } else {
            const double d = vgs - vt_;
            res[0] = (a_ / 2) * (d * d) * (1 + lam_ * vds) - id;
            jv[0] = a_ * d * (1 - dvt_) * (1 + lam_ * vds) + da / 2 * (d * d) * (1 + lam_ * vds);
            jv[1] = lam_ * a_ / 2 * (d * d);
        }
    }
    template <int R, class F> ACME_DI static double row(const double* jv, F m) {
        return fma(jv[1], m(1), fma(jv[0], m(0), -m(2)));
    }
};

struct JilesAtherton {  // elements.jl:104-135
    static constexpr int KIND = ACMEB200_ELEM_JA, NN = 1, NQ = 4, NPAR = 5, NC = 5, NJ = 4;
    ACME_DI static void prep(const double* P, double* C) {
        for (int i = 0; i < 5; i++) C[i] = P[i];
    }
    ACME_DI static double sgn(double x) { return (double)((x > 0) - (x < 0)); }
    ACME_DI static void eval(const double* C, const double* q, double* res, double* jv, const double* K = nullptr) {
        const double Ms = C[0], a = C[1], alpha = C[2], c = C[3], k = C[4];
        const double q1 = q[0], q2 = q[1], q3 = q[2], q4 = q[3];
        const double coth_q1 = 1 / tanh(q1);
        const double a_q1 = fabs(q1);
        const double L = a_q1 < 1e-4 ? q1 / 3 : coth_q1 - 1 / q1;
        const double Ld = a_q1 < 1e-4 ? 1.0 / 3 : 1 / (q1 * q1) - coth_q1 * coth_q1 + 1;
        const double Ld2 = a_q1 < 1e-3 ? -2.0 / 15 * q1
                                        : 2 * coth_q1 * (coth_q1 * coth_q1 - 1) - 2 / (q1 * q1 * q1);
        const double delta = q3 > 0 ? 1.0 : -1.0;
        const double Man = Ms * L;
        const double dM = sgn(q3) == sgn(Man - q2) ? 1.0 : 0.0;
        const double den = delta * (k * (1 - c)) - alpha * (Man - q2);
        const double s = 1e-4 / Ms;
        res[0] = s * ((1 - c) * dM * (Man - q2) / den * q3 + (c * Ms / a) * (q3 + alpha * q4) * Ld - q4);
        jv[0] = s * (((1 - c) * (1 - c) * k * Ms) * dM * Ld * delta / (den * den) * q3 +
                     (c * Ms / a) * (q3 + alpha * q4) * Ld2);
        jv[1] = s * -((1 - c) * (1 - c)) * k * dM * delta / (den * den) * q3;
        jv[2] = s * ((1 - c) * dM * (Man - q2) / den + (c * Ms / a) * Ld);
        jv[3] = s * ((c * Ms / a * alpha) * Ld - 1);
    }
    template <int R, class F> ACME_DI static double row(const double* jv, F m) {
        return fma(jv[3], m(3), fma(jv[2], m(2), fma(jv[1], m(1), jv[0] * m(0))));
    }
};

// runtime-kind dispatch helpers (generic kernel / host-side tables)
#define ACME_FOR_EACH_ELEM(X) X(Diode) X(Bjt) X(Pot) X(Mosfet) X(OpampTanh) X(JilesAtherton) X(TestQuad)

__host__ __device__ inline int elem_nn(int kind) {
    switch (kind) {
#define X(E) case E::KIND: return E::NN;
        ACME_FOR_EACH_ELEM(X)
#undef X
    }
    return -1;
}
__host__ __device__ inline int elem_nq(int kind) {
    switch (kind) {
#define X(E) case E::KIND: return E::NQ;
        ACME_FOR_EACH_ELEM(X)
#undef X
    }
    return -1;
}
__host__ __device__ inline int elem_npar(int kind) {
    switch (kind) {
#define X(E) case E::KIND: return E::NPAR;
        ACME_FOR_EACH_ELEM(X)
#undef X
    }
    return -1;
}
__host__ __device__ inline int elem_nc(int kind) {
    switch (kind) {
#define X(E) case E::KIND: return E::NC;
        ACME_FOR_EACH_ELEM(X)
#undef X
    }
    return -1;
}
__host__ __device__ inline int elem_nj(int kind) {
    switch (kind) {
#define X(E) case E::KIND: return E::NJ;
        ACME_FOR_EACH_ELEM(X)
#undef X
    }
    return -1;
}

}  // namespace acme
