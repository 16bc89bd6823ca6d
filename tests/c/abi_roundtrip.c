/* Plain-C host of the C ABI (include/acmeb200.h): what a Julia `ccall` binding does, without Python in between.
 *
 *   1. the hand-derived diode-clipper fixture F1 of SURVEY.md section 7 (examples/diodeclipper.jl, z = (v_C, i_d2))
 *      reproduces the reference's doctest samples (docs/src/gettingstarted.md:106-113) to the printed digits;
 *   2. create -> run(chunk 1) -> get_solver_state -> destroy -> create -> set_solver_state -> run(chunk 2) equals one run
 *      bit for bit (x, extrapolation origins and the learnt solution stores travel in the blob);
 *   3. the same batch through acmeb200_multi_* (all visible GPUs from one process) equals the single-device run.
 *
 * Build: gcc -std=c99 -I include tests/c/abi_roundtrip.c -o abi_roundtrip -L acme.jl_b200 -lacmeb200 -lm
 * (tests/test_c_abi.py does it; exit code 0 = all checks passed). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "acmeb200.h"

#define NB 6      /* instances: swept saturation currents */
#define NS 4410   /* samples */
#define CHECK(x) do { int rc_ = (x); if (rc_ != 0) { fprintf(stderr, "%s failed: %d %s\n", #x, rc_, acmeb200_last_error()); return 1; } } while (0)

static const double T = 1.0 / 44100, R = 1e3, Cc = 47e-9;
static double a_[1] = {-1}, b_[1] = {0}, c_[2], x0_[1] = {0}, dy_[1] = {0}, ey_[1] = {0}, fy_[2] = {1, 0}, y0_[1] = {0};
static double dq_[1], eq_[1], fqprev_[2] = {0, 0}, pexp_[4] = {0, 1, 0, 0}, q0_[4] = {0, 0, 0, 0}, fq_[8], initz_[2] = {0, 0};
static double params_[4 * NB];
static acmeb200_elem elems_[2] = {{ACMEB200_ELEM_DIODE, 0, 0, 2}, {ACMEB200_ELEM_DIODE, 2, 2, 2}};
static acmeb200_sub_desc sub_;
static acmeb200_model_desc desc_;

static acmeb200_array shared(const double *p) { acmeb200_array a; a.ptr = p; a.stride = 0; return a; }

static void build_desc(void) {
    c_[0] = 2 * Cc; c_[1] = 0;
    dq_[0] = 2 / T; eq_[0] = 1 / R;
    /* fq (4 x 2, column-major): rows (1, 0), (-(1/R + 2C/T), 1), (-1, 0), (0, 1) */
    fq_[0] = 1; fq_[1] = -(1 / R + 2 * Cc / T); fq_[2] = -1; fq_[3] = 0;
    fq_[4] = 0; fq_[5] = 1; fq_[6] = 0; fq_[7] = 1;
    for (int b = 0; b < NB; b++) {  /* instance 0 is the example's (is, eta) */
        const double is = 1e-15 * pow(10.0, 0.5 * b);
        params_[4 * b + 0] = is; params_[4 * b + 1] = 1.0; params_[4 * b + 2] = 1.8 * is; params_[4 * b + 3] = 1.0;
    }
    memset(&sub_, 0, sizeof sub_);
    sub_.nn = 2; sub_.nq = 4; sub_.np = 1; sub_.nelem = 2;
    sub_.dq = shared(dq_); sub_.eq = shared(eq_); sub_.fqprev = shared(fqprev_); sub_.pexp = shared(pexp_); sub_.q0 = shared(q0_);
    sub_.fq = shared(fq_); sub_.init_z = shared(initz_); sub_.elems = elems_;
    sub_.params.ptr = params_; sub_.params.stride = 4; sub_.nparams = 4;
    memset(&desc_, 0, sizeof desc_);
    desc_.abi_version = ACMEB200_ABI_VERSION;
    desc_.nx = 1; desc_.nu = 1; desc_.ny = 1; desc_.nsub = 1;
    desc_.solver = ACMEB200_SOLVER_HOMOTOPY_CACHING;
    desc_.a = shared(a_); desc_.b = shared(b_); desc_.c = shared(c_); desc_.x0 = shared(x0_);
    desc_.dy = shared(dy_); desc_.ey = shared(ey_); desc_.fy = shared(fy_); desc_.y0 = shared(y0_);
    desc_.subs = &sub_;
}

static int printed_equal(double got, double want) {  /* 6 significant digits, as the doctest prints */
    char a[64], b[64];
    snprintf(a, sizeof a, "%.6g", got); snprintf(b, sizeof b, "%.6g", want);
    return strcmp(a, b) == 0 || fabs(got - want) <= 1e-6 * fabs(want);
}

int main(void) {
    int32_t ndev = 0;
    CHECK(acmeb200_device_count(&ndev));
    if (ndev == 0) { fprintf(stderr, "no CUDA device\n"); return 77; }
    CHECK(acmeb200_set_device(0));
    build_desc();
    double *u = malloc(sizeof(double) * NS * NB), *y = malloc(sizeof(double) * NS * NB), *y2 = malloc(sizeof(double) * NS * NB),
           *y3 = malloc(sizeof(double) * NS * NB);
    for (int b = 0; b < NB; b++)
        for (int n = 0; n < NS; n++) u[b * NS + n] = (1.0 + 3.0 * b) * sin(2 * M_PI * 1000 / 44100 * n);  /* harder drive for the later instances */

    /* 1: one run, golden samples of instance 0 */
    acmeb200_model *m = NULL;
    CHECK(acmeb200_model_create(&desc_, 0, NB, &m));
    CHECK(acmeb200_run(m, u, NS, y, NS, NS, 0, NULL));
    const double golden[4] = {0.0, 0.0275964, 0.0990996, 0.195777};
    for (int n = 0; n < 4; n++)
        if (!printed_equal(y[n], golden[n])) { fprintf(stderr, "golden sample %d: got %.9g want %.6g\n", n, y[n], golden[n]); return 1; }
    acmeb200_stats st1;
    CHECK(acmeb200_get_stats(m, &st1));
    acmeb200_model_destroy(m);

    /* 2: chunk 1 -> state -> destroy -> create -> state -> chunk 2 */
    const int n1 = 1777;
    CHECK(acmeb200_model_create(&desc_, 0, NB, &m));
    CHECK(acmeb200_run(m, u, NS, y2, NS, n1, 0, NULL));
    const int64_t bytes = acmeb200_solver_state_size(m);
    if (bytes <= 0) { fprintf(stderr, "state size %lld\n", (long long)bytes); return 1; }
    void *blob = malloc((size_t)bytes);
    CHECK(acmeb200_get_solver_state(m, blob, bytes));
    acmeb200_model_destroy(m);
    CHECK(acmeb200_model_create(&desc_, 0, NB, &m));
    CHECK(acmeb200_set_solver_state(m, blob, bytes));
    CHECK(acmeb200_run(m, u + n1, NS, y2 + n1, NS, NS - n1, 0, NULL));
    acmeb200_stats st2;
    CHECK(acmeb200_get_stats(m, &st2));
    acmeb200_model_destroy(m);
    if (memcmp(y, y2, sizeof(double) * NS * NB) != 0) { fprintf(stderr, "resumed run differs from the one-shot run\n"); return 1; }
    if (st1.samples != st2.samples || st1.newton_iters != st2.newton_iters) { fprintf(stderr, "statistics differ after the resume\n"); return 1; }
    if (acmeb200_set_solver_state(NULL, blob, bytes) == 0) { fprintf(stderr, "null model accepted\n"); return 1; }

    /* 3: every visible GPU from this one process */
    acmeb200_multi *mm = NULL;
    CHECK(acmeb200_multi_create(&desc_, NB, 0, &mm));
    CHECK(acmeb200_multi_run(mm, u, NS, y3, NS, NS, 0));
    const int shards = acmeb200_multi_shards(mm);
    int64_t covered = 0;
    for (int g = 0; g < shards; g++) {
        int64_t first = -1, count = 0;
        int32_t dev = -1;
        acmeb200_model *sm = acmeb200_multi_model(mm, g, &first, &count);
        if (!sm || first != covered) { fprintf(stderr, "shard %d does not start where the previous one ended\n", g); return 1; }
        CHECK(acmeb200_get_device(sm, &dev));
        if (dev != g) { fprintf(stderr, "shard %d on device %d\n", g, dev); return 1; }
        covered += count;
    }
    acmeb200_multi_destroy(mm);
    if (covered != NB || memcmp(y, y3, sizeof(double) * NS * NB) != 0) { fprintf(stderr, "multi-GPU run differs from the single-device run\n"); return 1; }

    printf("abi_roundtrip ok: %d instances x %d samples, %lld Newton iterations, state blob %lld bytes, %d device(s), %d shard(s)\n", NB, NS,
           (long long)st1.newton_iters, (long long)bytes, (int)ndev, shards);
    free(u); free(y); free(y2); free(y3); free(blob);
    return 0;
}
