/*
 * acme_oracle.c -- CPU restatement of the reference's run!/step!/solver path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this file's library.
 * The shipped path is the CUDA library under acme.jl_b200/csrc and never calls
 * into here.
 *
 * The reference (HSU-ANT/ACME.jl, pure Julia) cannot be executed in this image
 * (no Julia toolchain), so this is a function-by-function restatement, in
 * scalar IEEE double with the reference's operation order, of
 *     step!/run!                 src/ACME.jl:650-715
 *     LinearSolver               src/solvers.jl:38-137
 *     SimpleSolver               src/solvers.jl:151-236
 *     HomotopySolver             src/solvers.jl:247-302
 *     CachingSolver              src/solvers.jl:319-405
 *     KDTree / Alts / indnearest src/kdtree.jl:4-234
 *     model closures             src/ACME.jl:176-194, 236-252
 *     CircuitNLFunc              src/circuit.jl:6-20, 68-86
 *     element laws               src/elements.jl:25-30, 107-129, 238-244,
 *                                323-401, 453-479, 540-546
 * Parity status: PINNED against the reference's golden vectors G1
 * (docs/src/gettingstarted.md:106-113) and G2 (docs/src/ug.md:107-114) and the
 * known-answer tests K1-K9 of test/runtests.jl (see tests/test_oracle_*.py);
 * the waveforms of sallenkey/birdie/superover are NOT pinned by the reference's
 * own tests (they say "TODO: further validate y").
 *
 * Build: gcc -O3 -ffp-contract=off -pthread -shared -fPIC (see oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#include "../include/acmeb200.h"

#define CM(a, ld, i, j) ((a)[(size_t)(j) * (size_t)(ld) + (size_t)(i)]) /* column-major */

/* ------------------------------------------------------------------ LinearSolver */
typedef struct {
    int n;
    double *factors; /* n x n */
    int *ipiv;       /* 0-based */
} LinearSolver;

static void ls_init(LinearSolver *s, int n) {
    s->n = n;
    s->factors = (double *)calloc((size_t)n * n + 1, sizeof(double));
    s->ipiv = (int *)calloc((size_t)n + 1, sizeof(int));
}
static void ls_free(LinearSolver *s) { free(s->factors); free(s->ipiv); }

/* setlhs!  src/solvers.jl:46-96 */
static int ls_setlhs(LinearSolver *s, const double *A) {
    const int n = s->n;
    double *f = s->factors;
    memcpy(f, A, sizeof(double) * (size_t)n * n);
    for (int k = 0; k < n; k++) {
        int kp = k;
        double amax = 0.0;
        for (int i = k; i < n; i++) {
            double absi = fabs(CM(f, n, i, k));
            if (absi > amax) { kp = i; amax = absi; }
        }
        s->ipiv[k] = kp;
        if (CM(f, n, kp, k) != 0.0) {
            if (k != kp) {
                for (int i = 0; i < n; i++) {
                    double tmp = CM(f, n, k, i);
                    CM(f, n, k, i) = CM(f, n, kp, i);
                    CM(f, n, kp, i) = tmp;
                }
            }
            double fkkinv = CM(f, n, k, k) = 1.0 / CM(f, n, k, k);
            for (int i = k + 1; i < n; i++) CM(f, n, i, k) *= fkkinv;
        } else {
            return 0;
        }
        for (int j = k + 1; j < n; j++)
            for (int i = k + 1; i < n; i++)
                CM(f, n, i, j) -= CM(f, n, i, k) * CM(f, n, k, j);
    }
    return 1;
}

/* solve!  src/solvers.jl:98-132 (x may alias b) */
static void ls_solve(const LinearSolver *s, double *x, const double *b) {
    const int n = s->n;
    const double *f = s->factors;
    if (x != b) memcpy(x, b, sizeof(double) * (size_t)n);
    for (int i = 0; i < n; i++) {
        double t = x[i]; x[i] = x[s->ipiv[i]]; x[s->ipiv[i]] = t;
    }
    for (int j = 0; j < n; j++) {
        double xj = x[j];
        for (int i = j + 1; i < n; i++) x[i] -= CM(f, n, i, j) * xj;
    }
    for (int j = n - 1; j >= 0; j--) {
        double xj = x[j] = CM(f, n, j, j) * x[j];
        for (int i = 0; i < j; i++) x[i] -= CM(f, n, i, j) * xj;
    }
}

static void ls_copy(LinearSolver *d, const LinearSolver *s) {
    memcpy(d->factors, s->factors, sizeof(double) * (size_t)s->n * s->n);
    memcpy(d->ipiv, s->ipiv, sizeof(int) * (size_t)s->n);
}

/* exported for the K1 unit test (test/runtests.jl:23-41) */
int oracle_linsolve(int n, const double *A, const double *b, double *x) {
    LinearSolver s;
    ls_init(&s, n);
    int ok = ls_setlhs(&s, A);
    if (ok) ls_solve(&s, x, b);
    ls_free(&s);
    return ok;
}

/* ------------------------------------------------------------------ BLAS-like helpers
 * gemv 'N': y = A*x + beta*y, accumulated column by column like reference BLAS */
static void gemv(int m, int n, const double *A, const double *x, double beta, double *y) {
    if (beta == 0.0) for (int i = 0; i < m; i++) y[i] = 0.0;
    for (int j = 0; j < n; j++) {
        double t = x[j];
        for (int i = 0; i < m; i++) y[i] += CM(A, m, i, j) * t;
    }
}
/* C = A*B, A m x k, B k x n */
static void gemm(int m, int n, int k, const double *A, const double *B, double *C) {
    for (int j = 0; j < n; j++) {
        for (int i = 0; i < m; i++) CM(C, m, i, j) = 0.0;
        for (int l = 0; l < k; l++) {
            double t = CM(B, k, l, j);
            for (int i = 0; i < m; i++) CM(C, m, i, j) += CM(A, m, i, l) * t;
        }
    }
}

/* ------------------------------------------------------------------ element laws */
static double evalpoly(double x, const double *c, int n) { /* Horner, Base.evalpoly */
    double acc = c[n - 1];
    for (int i = n - 2; i >= 0; i--) acc = x * acc + c[i];
    return acc;
}
static double sgn(double x) { return (x > 0) - (x < 0); }

/* Evaluates one element: q (nq) -> res (nn), J (nn x nq, written into the
 * block-diagonal Jq at (row0, q_off)); Jq is nn_sub x nq_sub column-major. */
static void eval_element(int kind, const double *P, const double *q, double *res,
                         double *Jq, int ldj, int row0, int qoff) {
#define JQ(r, c) CM(Jq, ldj, row0 + (r), qoff + (c))
    switch (kind) {
    case ACMEB200_ELEM_DIODE: { /* src/elements.jl:238-244 */
        double is = P[0], eta = P[1];
        double v = q[0], i = q[1];
        double ex = exp(v * (1 / (25e-3 * eta)));
        res[0] = is * (ex - 1) - i;
        JQ(0, 0) = is / (25e-3 * eta) * ex;
        JQ(0, 1) = -1;
    } break;
    case ACMEB200_ELEM_POT: { /* src/elements.jl:25-30 */
        double r = P[0];
        double v1 = q[0], v2 = q[1], i1 = q[2], i2 = q[3], pos = q[4];
        res[0] = v1 - r * pos * i1;
        res[1] = v2 - r * (1 - pos) * i2;
        JQ(0, 0) = 1; JQ(0, 1) = 0; JQ(0, 2) = -r * pos; JQ(0, 3) = 0; JQ(0, 4) = -r * i1;
        JQ(1, 0) = 0; JQ(1, 1) = 1; JQ(1, 2) = 0; JQ(1, 3) = -r * (1 - pos); JQ(1, 4) = -r * i2;
    } break;
    case ACMEB200_ELEM_OPAMP_TANH: { /* src/elements.jl:540-546 */
        double gain = P[0], scale = P[1];
        double vi = q[0], vo = q[1];
        double vs = vi * (gain / scale);
        double ch = cosh(vs);
        res[0] = tanh(vs) * scale - vo;
        JQ(0, 0) = gain / (ch * ch);
        JQ(0, 1) = -1;
    } break;
    case ACMEB200_ELEM_BJT: { /* src/elements.jl:323-401 */
        double ise = P[0], isc = P[1], ne = P[2], nc = P[3], bf = P[4], br = P[5];
        double ile = P[6], ilc = P[7], nel = P[8], ncl = P[9];
        double vaf = P[10], var = P[11], ikf = P[12], ikr = P[13];
        double vE = q[0], vC = q[1], iE = q[2], iC = q[3];
        double expE = exp(vE * (1 / (25e-3 * ne)));
        double expC = exp(vC * (1 / (25e-3 * nc)));
        double i_f = (bf / (1 + bf) * ise) * (expE - 1);
        double i_r = (br / (1 + br) * isc) * (expC - 1);
        double di_f1 = (bf / (1 + bf) * ise / (25e-3 * ne)) * expE;
        double di_r2 = (br / (1 + br) * isc / (25e-3 * nc)) * expC;
        double i_cc, di_cc1, di_cc2;
        int early = !(var == INFINITY && vaf == INFINITY);
        int knee = !(ikf == INFINITY && ikr == INFINITY);
        if (!early && !knee) {
            i_cc = i_f - i_r; di_cc1 = di_f1; di_cc2 = -di_r2;
        } else if (early && !knee) {
            double q1 = 1 - vE * (1 / var) - vC * (1 / vaf);
            i_cc = q1 * (i_f - i_r);
            di_cc1 = (-1 / var) * (i_f - i_r) + q1 * di_f1;
            di_cc2 = (-1 / vaf) * (i_f - i_r) - q1 * di_r2;
        } else if (!early && knee) {
            double q2 = i_f * (1 / ikf) + i_r * (1 / ikr);
            double qden = 1 + sqrt(1 + 4 * q2);
            double qfact = 2 / qden;
            i_cc = qfact * (i_f - i_r);
            double dq21 = di_f1 * (1 / ikf), dq22 = di_r2 * (1 / ikr);
            double dqf1 = -4 * dq21 / (qden - 1) / (qden * qden);
            double dqf2 = -4 * dq22 / (qden - 1) / (qden * qden);
            di_cc1 = dqf1 * (i_f - i_r) + qfact * di_f1;
            di_cc2 = dqf2 * (i_f - i_r) - qfact * di_r2;
        } else {
            double q1 = 1 - vE * (1 / var) - vC * (1 / vaf);
            double q2 = i_f * (1 / ikf) + i_r * (1 / ikr);
            double qden = 1 + sqrt(1 + 4 * q2);
            double qfact = 2 * q1 / qden;
            i_cc = qfact * (i_f - i_r);
            double dq11 = -1 / var, dq12 = -1 / vaf;
            double dq21 = di_f1 * (1 / ikf), dq22 = di_r2 * (1 / ikr);
            double dqf1 = (2 * dq11 * qden - q1 * 4 * dq21 / (qden - 1)) / (qden * qden);
            double dqf2 = (2 * dq12 * qden - q1 * 4 * dq22 / (qden - 1)) / (qden * qden);
            di_cc1 = dqf1 * (i_f - i_r) + qfact * di_f1;
            di_cc2 = dqf2 * (i_f - i_r) - qfact * di_r2;
        }
        double iBE = (1 / bf) * i_f, diBE1 = (1 / bf) * di_f1;
        if (ile != 0) {
            double expEl = (nel != ne) ? exp(vE * (1 / (25e-3 * nel))) : expE;
            iBE += ile * (expEl - 1);
            diBE1 += (ile / (25e-3 * ne)) * expEl;
        }
        double iBC = (1 / br) * i_r, diBC2 = (1 / br) * di_r2;
        if (ilc != 0) {
            double expCl = (ncl != nc) ? exp(vC * (1 / (25e-3 * ncl))) : expC;
            iBC += ilc * (expCl - 1);
            diBC2 += (ilc / (25e-3 * nc)) * expCl;
        }
        res[0] = i_cc + iBE - iE;
        res[1] = -i_cc + iBC - iC;
        JQ(0, 0) = di_cc1 + diBE1; JQ(0, 1) = di_cc2; JQ(0, 2) = -1.0; JQ(0, 3) = 0.0;
        JQ(1, 0) = -di_cc1; JQ(1, 1) = -di_cc2 + diBC2; JQ(1, 2) = 0.0; JQ(1, 3) = -1.0;
    } break;
    case ACMEB200_ELEM_MOSFET: { /* src/elements.jl:453-479 */
        double pol = P[0], lam = P[1];
        int nvt = (int)P[2], nal = (int)P[3];
        const double *vt = P + 4, *al = P + 8;
        double dvt[3], dal[3];
        for (int k = 1; k < nvt; k++) dvt[k - 1] = vt[k] * k;
        for (int k = 1; k < nal; k++) dal[k - 1] = al[k] * k;
        double vgs = q[0], vds = q[1], id = q[2];
        double a_ = evalpoly(pol * vgs, al, nal);
        double da = nal > 1 ? evalpoly(pol * vgs, dal, nal - 1) : 0;
        double vt_ = evalpoly(pol * vgs, vt, nvt);
        double dvt_ = nvt > 1 ? evalpoly(pol * vgs, dvt, nvt - 1) : 0;
        double lam_ = vds >= 0 ? lam : 0.0;
        if (vgs <= vt_) {
            res[0] = -id;
            JQ(0, 0) = 0.0; JQ(0, 1) = 0.0; JQ(0, 2) = -1.0;
        } else if (vds <= vgs - vt_) {
            res[0] = a_ * (vgs - vt_ - 0.5 * vds) * vds * (1 + lam_ * vds) - id;
            JQ(0, 0) = a_ * (1 - dvt_) * vds * (1 + lam_ * vds)
                     + da * (vgs - vt_ - 0.5 * vds) * vds * (1 + lam_ * vds);
            JQ(0, 1) = a_ * (vgs - vt_ + vds * (2 * lam_ * (vgs - vt_ - 0.75 * vds) - 1));
            JQ(0, 2) = -1.0;
        } else {
            double d = vgs - vt_;
            res[0] = (a_ / 2) * (d * d) * (1 + lam_ * vds) - id;
            JQ(0, 0) = a_ * d * (1 - dvt_) * (1 + lam_ * vds) + da / 2 * (d * d) * (1 + lam_ * vds);
            JQ(0, 1) = lam_ * a_ / 2 * (d * d);
            JQ(0, 2) = -1.0;
        }
    } break;
    case ACMEB200_ELEM_JA: { /* src/elements.jl:107-129 */
        double Ms = P[0], a = P[1], alpha = P[2], c = P[3], k = P[4];
        double q1 = q[0], q2 = q[1], q3 = q[2], q4 = q[3];
        double coth_q1 = 1 / tanh(q1);
        double a_q1 = fabs(q1);
        double L = a_q1 < 1e-4 ? q1 / 3 : coth_q1 - 1 / q1;
        double Ld = a_q1 < 1e-4 ? 1.0 / 3 : 1 / (q1 * q1) - coth_q1 * coth_q1 + 1;
        double Ld2 = a_q1 < 1e-3 ? -2.0 / 15 * q1
                                 : 2 * coth_q1 * (coth_q1 * coth_q1 - 1) - 2 / (q1 * q1 * q1);
        double delta = q3 > 0 ? 1.0 : -1.0;
        double Man = Ms * L;
        double dM = sgn(q3) == sgn(Man - q2) ? 1.0 : 0.0;
        double den = delta * (k * (1 - c)) - alpha * (Man - q2);
        double s = 1e-4 / Ms;
        res[0] = s * ((1 - c) * dM * (Man - q2) / den * q3 + (c * Ms / a) * (q3 + alpha * q4) * Ld - q4);
        JQ(0, 0) = s * (((1 - c) * (1 - c) * k * Ms) * dM * Ld * delta / (den * den) * q3
                        + (c * Ms / a) * (q3 + alpha * q4) * Ld2);
        JQ(0, 1) = s * -((1 - c) * (1 - c)) * k * dM * delta / (den * den) * q3;
        JQ(0, 2) = s * ((1 - c) * dM * (Man - q2) / den + (c * Ms / a) * Ld);
        JQ(0, 3) = s * ((c * Ms / a * alpha) * Ld - 1);
    } break;
    case ACMEB200_ELEM_TEST_QUAD: { /* test/runtests.jl:207-219: z^2 - 1 + p */
        res[0] = q[0] * q[0] - 1 + q[1];
        JQ(0, 0) = 2 * q[0];
        JQ(0, 1) = 1;
    } break;
    default:
        res[0] = NAN;
    }
#undef JQ
}

static int elem_nn(int kind) {
    switch (kind) {
    case ACMEB200_ELEM_BJT: case ACMEB200_ELEM_POT: return 2;
    default: return 1;
    }
}
/* ------------------------------------------------------------------ KDTree (src/kdtree.jl) */
typedef struct {
    int np;       /* point dimension */
    int n_points; /* Np used to build */
    int *cut_dim; /* n_points-1, 1-based dims */
    double *cut_val;
    int *ps_idx;  /* n_points, 1-based columns */
    double *ps;   /* np x cap */
    int cap;      /* columns of ps */
} KDTree;

typedef struct { int idx; double *delta; double delta_norm; } AltEntry;
typedef struct {
    AltEntry *entries;
    int n_alloc;
    int np;
    double best_dist;
    int best_pidx;
    int number_valid;
} Alts;

static void kdtree_free_index(KDTree *t) {
    free(t->cut_dim); free(t->cut_val); free(t->ps_idx);
    t->cut_dim = NULL; t->cut_val = NULL; t->ps_idx = NULL;
}

static int calc_cut_idx(int min_idx, int max_idx) { /* src/kdtree.jl:12-20, 1-based */
    int N = max_idx - min_idx + 1;
    int e = 0;
    while ((1 << (e + 1)) <= N - 1) e++;
    int N2 = 1 << e;
    if (3 * (N2 / 2) <= N) return min_idx + N2 - 1;
    return min_idx + N - N2 / 2 - 1;
}

typedef struct { double v; int i; } SortItem;
static void stable_sort_items(SortItem *a, int n, SortItem *tmp) { /* merge sort = stable, like sortperm */
    if (n < 2) return;
    int h = n / 2;
    stable_sort_items(a, h, tmp);
    stable_sort_items(a + h, n - h, tmp);
    int i = 0, j = h, k = 0;
    while (i < h && j < n) tmp[k++] = (a[j].v < a[i].v) ? a[j++] : a[i++];
    while (i < h) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, sizeof(SortItem) * (size_t)n);
}

/* argmax(vec(var(p[:, cols], dims=2))): first maximum, NaN counts as maximal
 * (Julia's argmax); cols are 1-based column ids */
static int argmax_var(const KDTree *t, const int *cols, int n) {
    int best = 1;
    double bestv = 0;
    for (int d = 0; d < t->np; d++) {
        double mean = 0;
        for (int k = 0; k < n; k++) mean += CM(t->ps, t->np, d, cols[k] - 1);
        mean /= n;
        double ss = 0;
        for (int k = 0; k < n; k++) {
            double dv = CM(t->ps, t->np, d, cols[k] - 1) - mean;
            ss += dv * dv;
        }
        double v = ss / (n - 1);
        if (d == 0) { best = 1; bestv = v; continue; }
        if (isnan(bestv)) continue;
        if (isnan(v) || v > bestv) { best = d + 1; bestv = v; }
    }
    return best;
}

/* KDTree(p, Np)  src/kdtree.jl:11-73.  NOTE (faithful quirk): the initial
 * sortperm runs over ALL `cap` columns of p (`p[dim,:]`, kdtree.jl:37), not
 * only the first Np, so zero-filled spare capacity columns can enter the tree. */
static void kdtree_build(KDTree *t, int Np) {
    kdtree_free_index(t);
    t->n_points = Np;
    if (Np == 0) return;
    t->ps_idx = (int *)calloc((size_t)Np, sizeof(int));
    if (Np == 1) { t->ps_idx[0] = 1; return; }
    int nin = Np - 1;
    int *min_idx = (int *)calloc((size_t)nin + 1, sizeof(int));
    int *max_idx = (int *)calloc((size_t)nin + 1, sizeof(int));
    int *cut_idx = (int *)calloc((size_t)nin + 1, sizeof(int));
    t->cut_dim = (int *)calloc((size_t)nin, sizeof(int));
    t->cut_val = (double *)calloc((size_t)nin, sizeof(double));
    int cap = t->cap;
    int *p_idx = (int *)calloc((size_t)cap + 1, sizeof(int)); /* 1-based content, 0-based storage */
    SortItem *items = (SortItem *)calloc((size_t)cap + 1, sizeof(SortItem));
    SortItem *tmp = (SortItem *)calloc((size_t)cap + 1, sizeof(SortItem));
    int *cols = (int *)calloc((size_t)cap + 1, sizeof(int));

    for (int k = 0; k < Np; k++) cols[k] = k + 1;
    int dim = argmax_var(t, cols, Np);
    for (int k = 0; k < cap; k++) { items[k].v = CM(t->ps, t->np, dim - 1, k); items[k].i = k + 1; }
    stable_sort_items(items, cap, tmp);
    for (int k = 0; k < cap; k++) p_idx[k] = items[k].i;

    min_idx[1] = 1; max_idx[1] = Np;
    cut_idx[1] = calc_cut_idx(1, Np);
    t->cut_dim[0] = dim;
    t->cut_val[0] = (CM(t->ps, t->np, dim - 1, p_idx[cut_idx[1] - 1] - 1)
                   + CM(t->ps, t->np, dim - 1, p_idx[cut_idx[1]] - 1)) / 2;
    for (int n = 2; n <= Np - 1; n++) {
        int parent = n / 2;
        if (n % 2 == 0) { min_idx[n] = min_idx[parent]; max_idx[n] = cut_idx[parent]; }
        else { min_idx[n] = cut_idx[parent] + 1; max_idx[n] = max_idx[parent]; }
        int cnt = max_idx[n] - min_idx[n] + 1;
        for (int k = 0; k < cnt; k++) cols[k] = p_idx[min_idx[n] - 1 + k];
        dim = argmax_var(t, cols, cnt);
        for (int k = 0; k < cnt; k++) { items[k].v = CM(t->ps, t->np, dim - 1, cols[k] - 1); items[k].i = cols[k]; }
        stable_sort_items(items, cnt, tmp);
        for (int k = 0; k < cnt; k++) p_idx[min_idx[n] - 1 + k] = items[k].i;
        cut_idx[n] = calc_cut_idx(min_idx[n], max_idx[n]);
        t->cut_dim[n - 1] = dim;
        t->cut_val[n - 1] = (CM(t->ps, t->np, dim - 1, p_idx[cut_idx[n] - 1] - 1)
                           + CM(t->ps, t->np, dim - 1, p_idx[cut_idx[n]] - 1)) / 2;
    }
    for (int n = 1; n <= Np; n++) {
        int parent = (n + Np - 1) / 2;
        if ((n + Np) % 2 == 1) t->ps_idx[n - 1] = p_idx[min_idx[parent] - 1];
        else t->ps_idx[n - 1] = p_idx[max_idx[parent] - 1];
    }
    free(min_idx); free(max_idx); free(cut_idx); free(p_idx); free(items); free(tmp); free(cols);
}

static void alts_init_alloc(Alts *a, int np) {
    a->np = np;
    a->n_alloc = 1;
    a->entries = (AltEntry *)calloc(1, sizeof(AltEntry));
    a->entries[0].delta = (double *)calloc((size_t)np + 1, sizeof(double));
    a->entries[0].idx = 1;
    a->best_dist = INFINITY; a->best_pidx = 0; a->number_valid = 1;
}
static void alts_free(Alts *a) {
    for (int i = 0; i < a->n_alloc; i++) free(a->entries[i].delta);
    free(a->entries);
}
/* init!  src/kdtree.jl:93-100 */
static void alts_init(Alts *a, double best_dist, int best_pidx) {
    a->number_valid = 1;
    a->entries[0].idx = 1;
    memset(a->entries[0].delta, 0, sizeof(double) * (size_t)a->np);
    a->entries[0].delta_norm = 0;
    a->best_dist = best_dist;
    a->best_pidx = best_pidx;
}
static void alt_swap(AltEntry *x, AltEntry *y) { AltEntry t = *x; *x = *y; *y = t; }
/* entries are 1-based in the reference: E(i) */
#define E(i) (a->entries[(i) - 1])
static void siftup(Alts *a, int i) { /* src/kdtree.jl:102-113 */
    int parent = i / 2;
    while (i > 1 && E(i).delta_norm < E(parent).delta_norm) {
        alt_swap(&E(i), &E(parent));
        i = parent; parent = i / 2;
    }
}
static void siftdown(Alts *a, int i) { /* src/kdtree.jl:115-135 */
    int N = a->number_valid;
    for (;;) {
        int mn = i;
        if (2 * i <= N && E(2 * i).delta_norm < E(mn).delta_norm) mn = 2 * i;
        if (2 * i + 1 <= N && E(2 * i + 1).delta_norm < E(mn).delta_norm) mn = 2 * i + 1;
        if (mn == i) break;
        alt_swap(&E(i), &E(mn));
        i = mn;
    }
}
static void alts_deleteat(Alts *a, int i) { /* src/kdtree.jl:140-150 */
    alt_swap(&E(i), &E(a->number_valid));
    a->number_valid -= 1;
    if (i <= a->number_valid) {
        if (i == 1 || E(i).delta_norm > E(i / 2).delta_norm) siftdown(a, i);
        else siftup(a, i);
    }
}
static void alts_enqueue(Alts *a, int new_idx, const double *ref_delta, int upd_dim /*0-based*/,
                         double upd_val, double new_norm) { /* src/kdtree.jl:158-175 */
    if (a->number_valid == a->n_alloc) {
        a->entries = (AltEntry *)realloc(a->entries, sizeof(AltEntry) * (size_t)(a->n_alloc + 1));
        a->entries[a->n_alloc].delta = (double *)calloc((size_t)a->np + 1, sizeof(double));
        a->n_alloc += 1;
    }
    AltEntry *e = &a->entries[a->number_valid];
    e->idx = new_idx;
    memmove(e->delta, ref_delta, sizeof(double) * (size_t)a->np);
    e->delta[upd_dim] = upd_val;
    e->delta_norm = new_norm;
    if (e->delta_norm < a->best_dist) {
        a->number_valid += 1;
        siftup(a, a->number_valid);
    }
}
static void alts_update_best(Alts *a, double dist, int p_idx) { /* src/kdtree.jl:177-187 */
    if (dist < a->best_dist) {
        a->best_dist = dist;
        a->best_pidx = p_idx;
        for (int i = a->number_valid; i >= 1; i--)
            if (E(i).delta_norm >= a->best_dist) alts_deleteat(a, i);
    }
}
#undef E

/* diagnostics (not part of the reference; compiled in with -DORACLE_KD_PROFILE only: shared counters would make the
 * threads of the CPU baseline fight over one cache line): how much work the tree search does */
#ifdef ORACLE_KD_PROFILE
#define KDPROF(x) x
#else
#define KDPROF(x)
#endif
static unsigned long long g_kd_reorigin, g_kd_queries, g_kd_leaves, g_kd_nodes, g_kd_maxheap, g_kd_rebuilds, g_kd_newscan, g_kd_maxleaves;
void oracle_kd_counters(unsigned long long *out, int reset) {
    out[0] = g_kd_queries; out[1] = g_kd_leaves; out[2] = g_kd_nodes; out[3] = g_kd_maxheap; out[4] = g_kd_rebuilds; out[5] = g_kd_newscan; out[6] = g_kd_maxleaves; out[7] = g_kd_reorigin;
    if (reset) g_kd_reorigin = g_kd_queries = g_kd_leaves = g_kd_nodes = g_kd_maxheap = g_kd_rebuilds = g_kd_newscan = g_kd_maxleaves = 0;
}

/* indnearest  src/kdtree.jl:192-234 (max_leaves = typemax) */
static int kd_indnearest(const KDTree *t, const double *p, Alts *a, double *delta_scratch) {
    int ncut = t->n_points - 1;
    if (ncut < 0) ncut = 0;
    KDPROF(g_kd_queries++;)
    unsigned long long leaves_here = 0;
    while (a->number_valid > 0) {
        KDPROF(if ((unsigned long long)a->number_valid > g_kd_maxheap) g_kd_maxheap = (unsigned long long)a->number_valid;)
        /* dequeue!: copy the best entry out (its storage is recycled by enqueue) */
        int idx = a->entries[0].idx;
        double delta_norm = a->entries[0].delta_norm;
        memcpy(delta_scratch, a->entries[0].delta, sizeof(double) * (size_t)a->np);
        alts_deleteat(a, 1);
        if (t->n_points == 0) break;
        while (idx <= ncut) {
            KDPROF(g_kd_nodes++;)
            int dim = t->cut_dim[idx - 1] - 1;
            double dcut = p[dim] - t->cut_val[idx - 1];
            double new_norm = delta_norm - delta_scratch[dim] * delta_scratch[dim] + dcut * dcut;
            if (new_norm < a->best_dist) {
                int new_idx = p[dim] <= t->cut_val[idx - 1] ? 2 * idx + 1 : 2 * idx;
                alts_enqueue(a, new_idx, delta_scratch, dim, dcut, new_norm);
            }
            if (p[dim] <= t->cut_val[idx - 1]) idx *= 2; else idx = 2 * idx + 1;
        }
        idx -= ncut;
        int p_idx = t->ps_idx[idx - 1];
        double dist = 0;
        for (int i = 0; i < t->np; i++) {
            double d = p[i] - CM(t->ps, t->np, i, p_idx - 1);
            dist += d * d;
        }
        alts_update_best(a, dist, p_idx);
        KDPROF(g_kd_leaves++;) leaves_here++;
    }
    KDPROF(if (leaves_here > g_kd_maxleaves) g_kd_maxleaves = leaves_here;)
    (void)leaves_here;
    return a->best_pidx;
}

/* exported for the K2 unit test (test/runtests.jl:186-205): builds a tree over
 * ps (np x n) and returns the nearest index (1-based) for each query */
int oracle_kdtree_query(int np, int n, const double *ps, int nq, const double *queries, int *out_idx) {
    KDTree t; memset(&t, 0, sizeof t);
    t.np = np; t.cap = n;
    t.ps = (double *)malloc(sizeof(double) * (size_t)np * n + 8);
    memcpy(t.ps, ps, sizeof(double) * (size_t)np * n);
    kdtree_build(&t, n);
    Alts a; alts_init_alloc(&a, np);
    double *scratch = (double *)calloc((size_t)np + 1, sizeof(double));
    for (int k = 0; k < nq; k++) {
        alts_init(&a, INFINITY, 0);
        out_idx[k] = kd_indnearest(&t, queries + (size_t)k * np, &a, scratch);
    }
    free(scratch); alts_free(&a); kdtree_free_index(&t); free(t.ps);
    return 0;
}
/* exports the tree arrays so a frozen cache can be handed to the device library */
int oracle_kdtree_build(int np, int n_cols, int n_points, const double *ps, int32_t *cut_dim,
                        double *cut_val, int32_t *ps_idx) {
    KDTree t; memset(&t, 0, sizeof t);
    t.np = np; t.cap = n_cols;
    t.ps = (double *)ps;
    kdtree_build(&t, n_points);
    for (int i = 0; i < n_points - 1; i++) { cut_dim[i] = t.cut_dim[i]; cut_val[i] = t.cut_val[i]; }
    for (int i = 0; i < n_points; i++) ps_idx[i] = t.ps_idx[i];
    kdtree_free_index(&t);
    return 0;
}

/* ------------------------------------------------------------------ model + solvers */
typedef struct {
    int nn, nq, np, nelem, nparams;
    int nn_total, nx, nu;
    /* shared or per-instance matrices are resolved per instance into these */
    const double *dq, *eq, *fqprev, *pexp, *q0, *fq, *params;
    const acmeb200_elem *elems;
} SubView;

typedef struct {
    SubView v;
    /* ParametricNonLinEq  src/solvers.jl:6-36 */
    double *res, *Jp, *J, *pfull, *Jq, *q;
    /* SimpleSolver        src/solvers.jl:151-179 */
    double *z, *last_z, *last_p, *last_Jp, *tmp_nn, *tmp_np;
    LinearSolver linsolver, last_linsolver;
    int iters;
    double resmaxabs, tol;
    int maxiter;
    /* HomotopySolver      src/solvers.jl:247-260 */
    double *start_p, *pa;
    int h_iters;
    int used_homotopy;
    /* CachingSolver       src/solvers.jl:319-339 */
    KDTree tree;
    double *zs; /* nn x cap */
    int num_ps, new_count, new_count_limit;
    Alts alts;
    double *p_tmp, *delta_scratch;
} SubSolver;

static void set_p(SubSolver *s, const double *p) { /* src/ACME.jl:237-243 */
    memcpy(s->pfull, s->v.q0, sizeof(double) * (size_t)s->v.nq);
    gemv(s->v.nq, s->v.np, s->v.pexp, p, 1.0, s->pfull);
}
static void calc_Jp(SubSolver *s) { /* src/ACME.jl:246-251 */
    gemm(s->v.nn, s->v.np, s->v.nq, s->Jq, s->v.pexp, s->Jp);
}
static void evaluate(SubSolver *s, const double *z) { /* src/ACME.jl:178-188, circuit.jl:10-17 */
    const SubView *v = &s->v;
    memcpy(s->q, s->pfull, sizeof(double) * (size_t)v->nq);
    gemv(v->nq, v->nn, v->fq, z, 1.0, s->q);
    memset(s->Jq, 0, sizeof(double) * (size_t)v->nn * v->nq);
    int row = 0;
    for (int e = 0; e < v->nelem; e++) {
        const acmeb200_elem *el = &v->elems[e];
        eval_element(el->kind, v->params + el->param_offset, s->q + el->q_offset, s->res + row,
                     s->Jq, v->nn, row, el->q_offset);
        row += elem_nn(el->kind);
    }
    gemm(v->nn, v->nn, v->nq, s->Jq, v->fq, s->J);
}
static int all_finite(const double *a, int n) {
    for (int i = 0; i < n; i++) if (!isfinite(a[i])) return 0;
    return 1;
}
static double maximum_abs(const double *a, int n) { /* maximum(abs, res): NaN propagates */
    double m = 0;
    for (int i = 0; i < n; i++) {
        double v = fabs(a[i]);
        if (isnan(v)) return NAN;
        if (v > m) m = v;
    }
    return m;
}
/* set_extrapolation_origin(solver, p, z)  src/solvers.jl:183-196 */
static void simple_set_origin_eval(SubSolver *s, const double *p, const double *z) {
    set_p(s, p);
    evaluate(s, z);
    ls_setlhs(&s->linsolver, s->J);
    calc_Jp(s);
    ls_copy(&s->last_linsolver, &s->linsolver);
    memcpy(s->last_Jp, s->Jp, sizeof(double) * (size_t)s->v.nn * s->v.np);
    memmove(s->last_p, p, sizeof(double) * (size_t)s->v.np);
    memmove(s->last_z, z, sizeof(double) * (size_t)s->v.nn);
}
static int simple_converged(const SubSolver *s) { return s->resmaxabs < s->tol; }

/* solve(::SimpleSolver, p)  src/solvers.jl:207-236 */
static double *simple_solve(SubSolver *s, const double *p) {
    const int nn = s->v.nn, np = s->v.np;
    set_p(s, p);
    for (int i = 0; i < np; i++) s->tmp_np[i] = p[i];
    for (int i = 0; i < np; i++) s->tmp_np[i] += -1.0 * s->last_p[i];
    gemv(nn, np, s->last_Jp, s->tmp_np, 0.0, s->tmp_nn);
    ls_solve(&s->last_linsolver, s->tmp_nn, s->tmp_nn);
    memcpy(s->z, s->last_z, sizeof(double) * (size_t)nn);
    for (int i = 0; i < nn; i++) s->z[i] += -1.0 * s->tmp_nn[i];

    for (s->iters = 1; s->iters <= s->maxiter; s->iters++) {
        evaluate(s, s->z);
        s->resmaxabs = nn == 0 ? 0.0 : maximum_abs(s->res, nn);
        if (!isfinite(s->resmaxabs) || !all_finite(s->J, nn * nn)) return s->z;
        if (!ls_setlhs(&s->linsolver, s->J)) return s->z;
        if (simple_converged(s)) break;
        ls_solve(&s->linsolver, s->tmp_nn, s->res);
        for (int i = 0; i < nn; i++) s->z[i] += -1.0 * s->tmp_nn[i];
    }
    if (s->iters > s->maxiter) s->iters = s->maxiter; /* Julia leaves the last loop value */
    if (simple_converged(s)) {
        calc_Jp(s);
        ls_copy(&s->last_linsolver, &s->linsolver);
        memcpy(s->last_Jp, s->Jp, sizeof(double) * (size_t)nn * np);
        memmove(s->last_p, p, sizeof(double) * (size_t)np);
        memcpy(s->last_z, s->z, sizeof(double) * (size_t)nn);
    }
    return s->z;
}

/* solve(::CachingSolver, p)  src/solvers.jl:347-396 */
static double *caching_solve(SubSolver *s, const double *p) {
    const int nn = s->v.nn, np = s->v.np;
    double best_diff = 0.0;
    for (int i = 0; i < np; i++) { double d = p[i] - s->last_p[i]; best_diff += d * d; }
    int idx = 0;
    for (int i = s->num_ps - s->new_count + 1; i <= s->num_ps; i++) {
        double diff = 0.;
        for (int j = 0; j < np; j++) { double d = CM(s->tree.ps, np, j, i - 1) - p[j]; diff += d * d; }
        if (diff < best_diff) { best_diff = diff; idx = i; }
    }
    KDPROF(g_kd_newscan += (unsigned long long)s->new_count;)
    alts_init(&s->alts, best_diff, idx);
    idx = kd_indnearest(&s->tree, p, &s->alts, s->delta_scratch);
    KDPROF(if (idx != 0) g_kd_reorigin++;)
    if (idx != 0)
        simple_set_origin_eval(s, &CM(s->tree.ps, np, 0, idx - 1), &CM(s->zs, nn, 0, idx - 1));

    double *z = simple_solve(s, p);
    if (s->iters > 5 && simple_converged(s)) {
        s->num_ps += 1;
        if (s->num_ps > s->tree.cap) {
            int ncap = 2 * s->num_ps;
            double *nps = (double *)calloc((size_t)np * ncap + 1, sizeof(double));
            double *nzs = (double *)calloc((size_t)nn * ncap + 1, sizeof(double));
            memcpy(nps, s->tree.ps, sizeof(double) * (size_t)np * s->tree.cap);
            memcpy(nzs, s->zs, sizeof(double) * (size_t)nn * s->tree.cap);
            free(s->tree.ps); free(s->zs);
            s->tree.ps = nps; s->zs = nzs; s->tree.cap = ncap;
        }
        for (int j = 0; j < np; j++) CM(s->tree.ps, np, j, s->num_ps - 1) = p[j];
        for (int j = 0; j < nn; j++) CM(s->zs, nn, j, s->num_ps - 1) = z[j];
        s->new_count += 1;
    }
    if (s->new_count > 0) s->new_count_limit -= 1;
    if (s->new_count > s->new_count_limit) {
        kdtree_build(&s->tree, s->num_ps);
        KDPROF(g_kd_rebuilds++;)
        s->new_count = 0;
        s->new_count_limit = 2 * s->tree.cap;
    }
    return z;
}

static double *base_solve(SubSolver *s, const double *p, int solver) {
    return solver == ACMEB200_SOLVER_HOMOTOPY_CACHING ? caching_solve(s, p) : simple_solve(s, p);
}

/* solve(::HomotopySolver, p)  src/solvers.jl:268-296 */
static double *homotopy_solve(SubSolver *s, const double *p, int solver) {
    const int np = s->v.np;
    double *z = base_solve(s, p, solver);
    s->h_iters = s->iters;
    s->used_homotopy = 0;
    if (solver == ACMEB200_SOLVER_SIMPLE) return z;
    if (!simple_converged(s)) {
        s->used_homotopy = 1;
        double a = 0.5, best_a = 0.0;
        memcpy(s->start_p, s->last_p, sizeof(double) * (size_t)np);
        while (best_a < 1) {
            for (int i = 0; i < np; i++) s->pa[i] = s->start_p[i];
            for (int i = 0; i < np; i++) s->pa[i] *= (1 - a);
            for (int i = 0; i < np; i++) s->pa[i] += a * p[i];
            z = base_solve(s, s->pa, solver);
            s->h_iters += s->iters;
            if (simple_converged(s)) {
                best_a = a;
                a = 1.0;
            } else {
                double new_a = (a + best_a) / 2;
                if (!(best_a < new_a && new_a < a)) break;
                a = new_a;
            }
        }
    }
    return z;
}

/* ------------------------------------------------------------------ instances */
typedef struct {
    const double *a, *b, *c, *x0, *dy, *ey, *fy, *y0;
    double *x, *xnew, *ucur, *ycur, *z, *p;
    SubSolver *subs;
    uint32_t status;
    int64_t first_fail;
} Instance;

typedef struct oracle_model {
    int nx, nu, ny, nsub, nn_total, solver;
    int64_t B;
    /* owned copies of every descriptor array (for [first, first+count)) */
    double **owned; int n_owned, cap_owned;
    acmeb200_elem **elem_tables;
    Instance *inst;
    int *sub_nn, *sub_np;
    acmeb200_stats stats;
    int64_t samples_done;
} oracle_model;

static double *own_array(oracle_model *m, acmeb200_array a, int64_t len, int64_t first, int64_t count,
                         int64_t *stride_out) {
    int64_t n = a.stride ? count : 1;
    double *d = (double *)calloc((size_t)(len * n) + 1, sizeof(double));
    if (a.ptr && len > 0) {
        if (a.stride) {
            for (int64_t b = 0; b < count; b++)
                memcpy(d + b * len, a.ptr + (first + b) * a.stride, sizeof(double) * (size_t)len);
        } else {
            memcpy(d, a.ptr, sizeof(double) * (size_t)len);
        }
    }
    *stride_out = a.stride ? len : 0;
    if (m->n_owned == m->cap_owned) {
        m->cap_owned = m->cap_owned ? 2 * m->cap_owned : 64;
        m->owned = (double **)realloc(m->owned, sizeof(double *) * (size_t)m->cap_owned);
    }
    m->owned[m->n_owned++] = d;
    return d;
}

static double *dalloc(int n) { return (double *)calloc((size_t)n + 1, sizeof(double)); }

static void sub_init_state(SubSolver *s, const double *init_z, int solver) {
    const int nn = s->v.nn, np = s->v.np;
    /* SimpleSolver ctor: set_extrapolation_origin(solver, zeros(np), init_z)  solvers.jl:176 */
    memset(s->p_tmp, 0, sizeof(double) * (size_t)np);
    simple_set_origin_eval(s, s->p_tmp, init_z);
    s->iters = 0; s->resmaxabs = 0.0;
    /* CachingSolver ctor  src/solvers.jl:327-333 */
    free(s->tree.ps); free(s->zs); kdtree_free_index(&s->tree);
    s->tree.np = np; s->tree.cap = 1;
    s->tree.ps = dalloc(np);
    s->zs = dalloc(nn);
    memcpy(s->zs, init_z, sizeof(double) * (size_t)nn);
    kdtree_build(&s->tree, 1);
    s->num_ps = 1; s->new_count = 0; s->new_count_limit = 2;
    (void)solver;
}

oracle_model *oracle_create(const acmeb200_model_desc *d, int64_t first, int64_t count) {
    oracle_model *m = (oracle_model *)calloc(1, sizeof(oracle_model));
    m->nx = d->nx; m->nu = d->nu; m->ny = d->ny; m->nsub = d->nsub; m->B = count;
    m->solver = d->solver;
    int nn_total = 0;
    for (int i = 0; i < d->nsub; i++) nn_total += d->subs[i].nn;
    m->nn_total = nn_total;
    const int nx = d->nx, nu = d->nu, ny = d->ny;
    double tol = d->tol > 0 ? d->tol : 1e-10;
    int maxiter = d->maxiter > 0 ? d->maxiter : 500;
    int64_t s_a, s_b, s_c, s_x0, s_dy, s_ey, s_fy, s_y0;
    double *A = own_array(m, d->a, (int64_t)nx * nx, first, count, &s_a);
    double *Bm = own_array(m, d->b, (int64_t)nx * nu, first, count, &s_b);
    double *C = own_array(m, d->c, (int64_t)nx * nn_total, first, count, &s_c);
    double *X0 = own_array(m, d->x0, nx, first, count, &s_x0);
    double *DY = own_array(m, d->dy, (int64_t)ny * nx, first, count, &s_dy);
    double *EY = own_array(m, d->ey, (int64_t)ny * nu, first, count, &s_ey);
    double *FY = own_array(m, d->fy, (int64_t)ny * nn_total, first, count, &s_fy);
    double *Y0 = own_array(m, d->y0, ny, first, count, &s_y0);
    m->inst = (Instance *)calloc((size_t)count + 1, sizeof(Instance));
    m->elem_tables = (acmeb200_elem **)calloc((size_t)d->nsub + 1, sizeof(acmeb200_elem *));
    typedef struct { double *p[8]; int64_t s[8]; } SubArrays;
    SubArrays *sa = (SubArrays *)calloc((size_t)d->nsub + 1, sizeof(SubArrays));
    for (int i = 0; i < d->nsub; i++) {
        const acmeb200_sub_desc *sd = &d->subs[i];
        m->elem_tables[i] = (acmeb200_elem *)calloc((size_t)sd->nelem + 1, sizeof(acmeb200_elem));
        memcpy(m->elem_tables[i], sd->elems, sizeof(acmeb200_elem) * (size_t)sd->nelem);
        sa[i].p[0] = own_array(m, sd->dq, (int64_t)sd->np * nx, first, count, &sa[i].s[0]);
        sa[i].p[1] = own_array(m, sd->eq, (int64_t)sd->np * nu, first, count, &sa[i].s[1]);
        sa[i].p[2] = own_array(m, sd->fqprev, (int64_t)sd->np * nn_total, first, count, &sa[i].s[2]);
        sa[i].p[3] = own_array(m, sd->pexp, (int64_t)sd->nq * sd->np, first, count, &sa[i].s[3]);
        sa[i].p[4] = own_array(m, sd->q0, sd->nq, first, count, &sa[i].s[4]);
        sa[i].p[5] = own_array(m, sd->fq, (int64_t)sd->nq * sd->nn, first, count, &sa[i].s[5]);
        sa[i].p[6] = own_array(m, sd->init_z, sd->nn, first, count, &sa[i].s[6]);
        sa[i].p[7] = own_array(m, sd->params, sd->nparams, first, count, &sa[i].s[7]);
    }
    for (int64_t b = 0; b < count; b++) {
        Instance *in = &m->inst[b];
        in->a = A + b * s_a; in->b = Bm + b * s_b; in->c = C + b * s_c; in->x0 = X0 + b * s_x0;
        in->dy = DY + b * s_dy; in->ey = EY + b * s_ey; in->fy = FY + b * s_fy; in->y0 = Y0 + b * s_y0;
        in->x = dalloc(nx); in->xnew = dalloc(nx); in->ucur = dalloc(nu); in->ycur = dalloc(ny);
        in->z = dalloc(nn_total);
        in->first_fail = -1;
        in->subs = (SubSolver *)calloc((size_t)d->nsub + 1, sizeof(SubSolver));
        for (int i = 0; i < d->nsub; i++) {
            const acmeb200_sub_desc *sd = &d->subs[i];
            SubSolver *s = &in->subs[i];
            s->v.nn = sd->nn; s->v.nq = sd->nq; s->v.np = sd->np; s->v.nelem = sd->nelem;
            s->v.nparams = sd->nparams; s->v.nn_total = nn_total; s->v.nx = nx; s->v.nu = nu;
            s->v.dq = sa[i].p[0] + b * sa[i].s[0]; s->v.eq = sa[i].p[1] + b * sa[i].s[1];
            s->v.fqprev = sa[i].p[2] + b * sa[i].s[2]; s->v.pexp = sa[i].p[3] + b * sa[i].s[3];
            s->v.q0 = sa[i].p[4] + b * sa[i].s[4]; s->v.fq = sa[i].p[5] + b * sa[i].s[5];
            s->v.params = sa[i].p[7] + b * sa[i].s[7];
            s->v.elems = m->elem_tables[i];
            const int nn = sd->nn, nq = sd->nq, np = sd->np;
            s->res = dalloc(nn); s->Jp = dalloc(nn * np); s->J = dalloc(nn * nn);
            s->pfull = dalloc(nq); s->Jq = dalloc(nn * nq); s->q = dalloc(nq);
            s->z = dalloc(nn); s->last_z = dalloc(nn); s->last_p = dalloc(np);
            s->last_Jp = dalloc(nn * np); s->tmp_nn = dalloc(nn); s->tmp_np = dalloc(np);
            ls_init(&s->linsolver, nn); ls_init(&s->last_linsolver, nn);
            s->tol = tol; s->maxiter = maxiter;
            s->start_p = dalloc(np); s->pa = dalloc(np); s->p_tmp = dalloc(np);
            s->delta_scratch = dalloc(np);
            alts_init_alloc(&s->alts, np);
            sub_init_state(s, sa[i].p[6] + b * sa[i].s[6], m->solver);
        }
    }
    /* keep init_z pointers for reset: stored as owned arrays; remember per-sub */
    m->sub_nn = (int *)calloc((size_t)d->nsub + 1, sizeof(int));
    m->sub_np = (int *)calloc((size_t)d->nsub + 1, sizeof(int));
    for (int i = 0; i < d->nsub; i++) { m->sub_nn[i] = d->subs[i].nn; m->sub_np[i] = d->subs[i].np; }
    free(sa);
    return m;
}

void oracle_destroy(oracle_model *m) {
    if (!m) return;
    for (int64_t b = 0; b < m->B; b++) {
        Instance *in = &m->inst[b];
        for (int i = 0; i < m->nsub; i++) {
            SubSolver *s = &in->subs[i];
            free(s->res); free(s->Jp); free(s->J); free(s->pfull); free(s->Jq); free(s->q);
            free(s->z); free(s->last_z); free(s->last_p); free(s->last_Jp); free(s->tmp_nn); free(s->tmp_np);
            ls_free(&s->linsolver); ls_free(&s->last_linsolver);
            free(s->start_p); free(s->pa); free(s->p_tmp); free(s->delta_scratch);
            alts_free(&s->alts);
            kdtree_free_index(&s->tree); free(s->tree.ps); free(s->zs);
        }
        free(in->subs); free(in->x); free(in->xnew); free(in->ucur); free(in->ycur); free(in->z);
    }
    for (int i = 0; i < m->n_owned; i++) free(m->owned[i]);
    for (int i = 0; i < m->nsub; i++) free(m->elem_tables[i]);
    free(m->owned); free(m->elem_tables); free(m->inst); free(m->sub_nn); free(m->sub_np);
    free(m);
}

/* step!  src/ACME.jl:666-715; returns 0 to continue, 1 if the reference would throw */
static int step(oracle_model *m, Instance *in, const double *u, double *y, int64_t n,
                acmeb200_stats *st) {
    const int nx = m->nx, nu = m->nu, ny = m->ny, nnt = m->nn_total;
    memcpy(in->ucur, u + n * nu, sizeof(double) * (size_t)nu);
    int zoff = 0;
    memset(in->z, 0, sizeof(double) * (size_t)nnt);
    for (int idx = 0; idx < m->nsub; idx++) {
        SubSolver *s = &in->subs[idx];
        double *p = s->p_tmp;
        if (nx == 0) memset(p, 0, sizeof(double) * (size_t)s->v.np);
        else gemv(s->v.np, nx, s->v.dq, in->x, 0.0, p);
        gemv(s->v.np, nu, s->v.eq, in->ucur, 1.0, p);
        if (idx > 0) gemv(s->v.np, nnt, s->v.fqprev, in->z, 1.0, p);
        double *zsub = homotopy_solve(s, p, m->solver);
        st->solves++;
        st->newton_iters += (uint64_t)s->h_iters;
        st->homotopy_solves += (uint64_t)s->used_homotopy;
        {
            int bin = s->h_iters < 1 ? 1 : s->h_iters;
            if (bin > ACMEB200_HIST_BINS) bin = ACMEB200_HIST_BINS;
            st->iter_hist[bin - 1]++;
        }
        if (!simple_converged(s)) {
            if (in->first_fail < 0) in->first_fail = n;
            if (all_finite(zsub, s->v.nn)) {
                in->status |= ACMEB200_STATUS_NOT_CONVERGED;
                st->not_converged++;
            } else {
                in->status |= ACMEB200_STATUS_NONFINITE;
                return 1;
            }
        }
        memcpy(in->z + zoff, zsub, sizeof(double) * (size_t)s->v.nn);
        zoff += s->v.nn;
    }
    if (ny > 0) {
        memcpy(in->ycur, in->y0, sizeof(double) * (size_t)ny);
        gemv(ny, nx, in->dy, in->x, 1.0, in->ycur);
        gemv(ny, nu, in->ey, in->ucur, 1.0, in->ycur);
        gemv(ny, nnt, in->fy, in->z, 1.0, in->ycur);
        memcpy(y + n * ny, in->ycur, sizeof(double) * (size_t)ny);
    }
    if (nx > 0) {
        memcpy(in->xnew, in->x0, sizeof(double) * (size_t)nx);
        gemv(nx, nx, in->a, in->x, 1.0, in->xnew);
        gemv(nx, nu, in->b, in->ucur, 1.0, in->xnew);
        gemv(nx, nnt, in->c, in->z, 1.0, in->xnew);
        memcpy(in->x, in->xnew, sizeof(double) * (size_t)nx);
    }
    st->samples++;
    return 0;
}

static void stats_add(acmeb200_stats *d, const acmeb200_stats *s) {
    d->samples += s->samples; d->solves += s->solves; d->newton_iters += s->newton_iters;
    d->homotopy_solves += s->homotopy_solves; d->not_converged += s->not_converged;
    for (int i = 0; i < ACMEB200_HIST_BINS; i++) d->iter_hist[i] += s->iter_hist[i];
}

/* run!(runner, y, u) for every instance; host threads over instances (each
 * instance has private solver state, like one deepcopy(model) per thread in Julia) */
typedef struct {
    oracle_model *m; const double *U; double *Y; int64_t u_stride, y_stride, N;
    int64_t *next; pthread_mutex_t *mu; acmeb200_stats local;
} RunJob;

static void *run_worker(void *arg) {
    RunJob *j = (RunJob *)arg;
    oracle_model *m = j->m;
    for (;;) {
        pthread_mutex_lock(j->mu);
        int64_t b = (*j->next)++;
        pthread_mutex_unlock(j->mu);
        if (b >= m->B) break;
        Instance *in = &m->inst[b];
        const double *u = j->U + b * j->u_stride;
        double *y = j->Y + b * j->y_stride;
        int64_t n = 0;
        if (!(in->status & ACMEB200_STATUS_NONFINITE))
            for (; n < j->N; n++)
                if (step(m, in, u, y, n, &j->local)) break;
        for (; n < j->N; n++) /* the reference throws here; mark the rest */
            for (int k = 0; k < m->ny; k++) y[n * m->ny + k] = NAN;
    }
    return NULL;
}

int oracle_num_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

int oracle_run(oracle_model *m, const double *U, int64_t u_stride, double *Y, int64_t y_stride,
               int64_t N, int nthreads) {
    if (y_stride == 0) y_stride = (int64_t)m->ny * N;
    if (nthreads <= 0) nthreads = oracle_num_threads();
    if (nthreads > m->B) nthreads = (int)(m->B > 0 ? m->B : 1);
    int64_t next = 0;
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    RunJob *jobs = (RunJob *)calloc((size_t)nthreads, sizeof(RunJob));
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    for (int t = 0; t < nthreads; t++) {
        jobs[t].m = m; jobs[t].U = U; jobs[t].Y = Y; jobs[t].u_stride = u_stride;
        jobs[t].y_stride = y_stride; jobs[t].N = N; jobs[t].next = &next; jobs[t].mu = &mu;
    }
    if (nthreads == 1) run_worker(&jobs[0]);
    else {
        for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, run_worker, &jobs[t]);
        for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    }
    for (int t = 0; t < nthreads; t++) stats_add(&m->stats, &jobs[t].local);
    free(jobs); free(th);
    m->samples_done += N;
    return 0;
}

int oracle_get_state(oracle_model *m, double *x) {
    for (int64_t b = 0; b < m->B; b++) memcpy(x + b * m->nx, m->inst[b].x, sizeof(double) * (size_t)m->nx);
    return 0;
}
int oracle_set_state(oracle_model *m, const double *x, int64_t stride) {
    for (int64_t b = 0; b < m->B; b++) memcpy(m->inst[b].x, x + b * stride, sizeof(double) * (size_t)m->nx);
    return 0;
}
int oracle_get_status(oracle_model *m, uint32_t *status, int64_t *first_fail) {
    for (int64_t b = 0; b < m->B; b++) {
        if (status) status[b] = m->inst[b].status;
        if (first_fail) first_fail[b] = m->inst[b].first_fail;
    }
    return 0;
}
int oracle_get_stats(oracle_model *m, acmeb200_stats *out) { *out = m->stats; return 0; }

/* direct access to one sub-problem solver of one instance (K3: test/runtests.jl:207-219;
 * also runtests.jl:721 `ACME.solve(model.solvers[1], [0.003, -0.0002])`) */
int oracle_solve_sub(oracle_model *m, int64_t inst, int sub, const double *p, double *z_out,
                     int *iters_out) {
    SubSolver *s = &m->inst[inst].subs[sub];
    double *z = homotopy_solve(s, p, m->solver);
    memcpy(z_out, z, sizeof(double) * (size_t)s->v.nn);
    if (iters_out) *iters_out = s->h_iters;
    return simple_converged(s);
}
/* cache introspection: number of stored solutions of (inst, sub) and a copy of them */
int oracle_cache_size(oracle_model *m, int64_t inst, int sub) { return m->inst[inst].subs[sub].num_ps; }
int oracle_cache_capacity(oracle_model *m, int64_t inst, int sub) { return m->inst[inst].subs[sub].tree.cap; }
int oracle_cache_export(oracle_model *m, int64_t inst, int sub, double *ps, double *zs) {
    SubSolver *s = &m->inst[inst].subs[sub];
    memcpy(ps, s->tree.ps, sizeof(double) * (size_t)s->v.np * s->tree.cap);
    memcpy(zs, s->zs, sizeof(double) * (size_t)s->v.nn * s->tree.cap);
    return s->num_ps;
}
