// C-ABI implementation (include/acmeb200.h): model upload, kernel selection,
// run! for host or device streams, state/statistics access.
//
// There is no CPU fallback in this library: every entry point that computes
// needs a CUDA device and fails with ACMEB200_ENODEVICE / ACMEB200_ECUDA otherwise.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/acmeb200.h"
#include "devmodel.h"
#include "elements.cuh"
#define ACME_GENERIC_KERNEL_TU
#include "kernel_generic.cuh"
#include "hostmodel.h"

using namespace acme;


// ------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CUDA_TRY(x)                                                                           \
    do {                                                                                      \
        cudaError_t e_ = (x);                                                                 \
        if (e_ != cudaSuccess)                                                                \
            return fail(e_ == cudaErrorMemoryAllocation ? ACMEB200_ENOMEM : ACMEB200_ECUDA,   \
                        "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

extern "C" const char* acmeb200_last_error(void) { return g_err.c_str(); }
extern "C" int acmeb200_abi_version(void) { return ACMEB200_ABI_VERSION; }

// ------------------------------------------------------------------ creation
static void prep_consts(int kind, const double* P, double* C) {
    switch (kind) {
#define X(E) case E::KIND: E::prep(P, C); break;
        ACME_FOR_EACH_ELEM(X)
#undef X
    }
}

template <class E, int R>
static void probe_row(RowProg& rp, int q_off, int j_off) {
    double jv[8];
    for (int t = 0; t < 8; t++) jv[t] = 1000.0 + 7.5 * t;
    rp.n = 0;
    for (int k = 0; k < E::NQ; k++) {
        const double v = E::template row<R>(jv, [&](int kk) { return kk == k ? 1.0 : 0.0; });
        if (v == 0.0) continue;
        int which = -1;
        for (int t = 0; t < E::NJ; t++) if (v == jv[t]) which = t;
        rp.q[rp.n] = (unsigned char)(q_off + k);
        rp.jv[rp.n] = (signed char)(which >= 0 ? j_off + which : -1);
        rp.c[rp.n] = which >= 0 ? 0.0f : (float)v;
        rp.n++;
    }
}
template <class E>
static void probe_elem(RowProg* rows, int q_off, int j_off) {
    probe_row<E, 0>(rows[0], q_off, j_off);
    if constexpr (E::NN > 1) probe_row<E, 1>(rows[1], q_off, j_off);
}
static void probe_rows(int kind, RowProg* rows, int q_off, int j_off) {
    switch (kind) {
#define X(E) case E::KIND: probe_elem<E>(rows, q_off, j_off); break;
        ACME_FOR_EACH_ELEM(X)
#undef X
    }
}

static int select_kernel(acmeb200_model* m);
static int run_init(acmeb200_model* m);

extern "C" void acmeb200_model_destroy(acmeb200_model* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    cudaFree(m->d_blob); cudaFree(m->d_consts); cudaFree(m->d_initz); cudaFree(m->d_ws);
    cudaFree(m->d_status); cudaFree(m->d_first_fail); cudaFree(m->d_stats);
    for (void* p : m->d_cache) cudaFree(p);
    for (void* p : m->d_dyn) cudaFree(p);
    for (int i = 0; i < 2; i++) {
        if (m->h_stage[i]) cudaFreeHost(m->h_stage[i]);
        cudaFree(m->d_stage_u[i]); cudaFree(m->d_stage_y[i]);
        if (m->copy_streams[i]) cudaStreamDestroy(m->copy_streams[i]);
        if (m->ev[i]) cudaEventDestroy(m->ev[i]);
        if (m->ev_in[i]) cudaEventDestroy(m->ev_in[i]);
        if (m->ev_k[i]) cudaEventDestroy(m->ev_k[i]);
        if (m->ev_out[i]) cudaEventDestroy(m->ev_out[i]);
    }
    if (m->compute_stream) cudaStreamDestroy(m->compute_stream);
    delete m;
}

extern "C" int acmeb200_model_create(const acmeb200_model_desc* d, int64_t first, int64_t count,
                                     acmeb200_model** out) {
    if (!d || !out) return fail(ACMEB200_EINVAL, "null argument");
    *out = nullptr;
    if (d->abi_version != ACMEB200_ABI_VERSION)
        return fail(ACMEB200_EINVAL, "descriptor ABI version %d, library %d", d->abi_version, ACMEB200_ABI_VERSION);
    if (count <= 0 || first < 0) return fail(ACMEB200_EINVAL, "bad instance range [%lld, +%lld)", (long long)first, (long long)count);
    if (d->nx < 0 || d->nu < 0 || d->ny < 0 || d->nsub < 0) return fail(ACMEB200_EINVAL, "negative dimension");
    if (d->nsub > MAX_SUBS) return fail(ACMEB200_EUNSUPPORTED, "%d sub-problems, at most %d supported", d->nsub, MAX_SUBS);
    if (d->solver < 0 || d->solver > 2) return fail(ACMEB200_EINVAL, "unknown solver %d", d->solver);
    if (d->nsub > 0 && !d->subs) return fail(ACMEB200_EINVAL, "nsub = %d but subs is null", d->nsub);
    for (int i = 0; i < d->nsub; i++) {  // pointer checks up front, so a bad descriptor never reaches the device
        const acmeb200_sub_desc& sd = d->subs[i];
        if (sd.nn < 0 || sd.nq < 0 || sd.np < 0 || sd.nelem < 0 || sd.nparams < 0) return fail(ACMEB200_EINVAL, "sub %d: negative dimension", i);
        if (sd.nelem > 0 && !sd.elems) return fail(ACMEB200_EINVAL, "sub %d: nelem = %d but elems is null", i, sd.nelem);
        if (sd.nparams > 0 && !sd.params.ptr) return fail(ACMEB200_EINVAL, "sub %d: nparams = %d but params is null", i, sd.nparams);
        const acmeb200_cache& c = sd.cache;
        if (c.n_points > 0 && (!c.ps_idx || (sd.np > 0 && !c.ps) || (sd.nn > 0 && !c.zs) || (c.n_points > 1 && (!c.cut_dim || !c.cut_val)) || c.n_columns < c.n_points))
            return fail(ACMEB200_EINVAL, "sub %d: incomplete frozen cache (null array or n_columns < n_points)", i);
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(ACMEB200_ENODEVICE, "no CUDA device available (this library has no CPU fallback)");

    acmeb200_model* m = new acmeb200_model();
    cudaGetDevice(&m->device);
    m->B = count;
    DevModel& dm = m->dm;
    memset(&dm, 0, sizeof dm);
    dm.nx = d->nx; dm.nu = d->nu; dm.ny = d->ny; dm.nsub = d->nsub;
    dm.solver = d->solver;
    dm.maxiter = d->maxiter > 0 ? d->maxiter : 500;
    dm.tol = d->tol > 0 ? d->tol : 1e-10;
    m->cache_capacity = d->cache_capacity >= 2 && d->cache_capacity <= (1 << 20) ? d->cache_capacity : 0;
    int nnt = 0;
    for (int i = 0; i < d->nsub; i++) nnt += d->subs[i].nn;
    dm.nnt = nnt;

    // ---- blob layout: a b c x0 dy ey fy y0 | per sub: dq eq fqprev pexp q0 fq
    struct Piece { acmeb200_array arr; int len; int* off; };
    std::vector<Piece> pieces;
    int off = 0;
    auto add = [&](acmeb200_array arr, int len, int* o) { *o = off; pieces.push_back({arr, len, o}); off += len; };
    add(d->a, dm.nx * dm.nx, &dm.o_a); add(d->b, dm.nx * dm.nu, &dm.o_b); add(d->c, dm.nx * nnt, &dm.o_c);
    add(d->x0, dm.nx, &dm.o_x0); add(d->dy, dm.ny * dm.nx, &dm.o_dy); add(d->ey, dm.ny * dm.nu, &dm.o_ey);
    add(d->fy, dm.ny * nnt, &dm.o_fy); add(d->y0, dm.ny, &dm.o_y0);
    int elem_total = 0, const_total = 0, initz_total = 0, zoff = 0, wrow = 0;
    int max_nn = 0, max_nq = 0, max_np = 0, max_njv = 0;
    bool per_instance = false;
    std::vector<std::vector<double>> h_consts;  // [const row][instance]
    for (int i = 0; i < d->nsub; i++) {
        const acmeb200_sub_desc& sd = d->subs[i];
        DevSub& s = dm.subs[i];
        if (sd.nn < 0 || sd.nq < 0 || sd.np < 0 || sd.nelem < 0) { delete m; return fail(ACMEB200_EINVAL, "negative sub dimension"); }
        s.nn = sd.nn; s.nq = sd.nq; s.np = sd.np; s.nelem = sd.nelem; s.elem0 = elem_total; s.zoff = zoff;
        s.o_initz = initz_total;
        add(sd.dq, sd.np * dm.nx, &s.o_dq); add(sd.eq, sd.np * dm.nu, &s.o_eq);
        add(sd.fqprev, sd.np * nnt, &s.o_fqprev); add(sd.pexp, sd.nq * sd.np, &s.o_pexp);
        add(sd.q0, sd.nq, &s.o_q0); add(sd.fq, sd.nq * sd.nn, &s.o_fq);
        if (elem_total + sd.nelem > MAX_ELEMS) { delete m; return fail(ACMEB200_EUNSUPPORTED, "more than %d non-linear elements", MAX_ELEMS); }
        int row = 0, joff = 0, qcheck = 0;
        for (int e = 0; e < sd.nelem; e++) {
            const acmeb200_elem& el = sd.elems[e];
            if (elem_nn(el.kind) < 0) { delete m; return fail(ACMEB200_EINVAL, "unknown element kind %d", el.kind); }
            if (el.nparam != elem_npar(el.kind) || el.param_offset < 0 || el.param_offset + el.nparam > sd.nparams) {
                delete m; return fail(ACMEB200_EINVAL, "element %d of sub %d: bad parameter range", e, i);
            }
            if (el.q_offset < 0 || el.q_offset + elem_nq(el.kind) > sd.nq) { delete m; return fail(ACMEB200_EINVAL, "element %d of sub %d: q range outside the sub", e, i); }
            DevElem& de = dm.elems[elem_total + e];
            de.kind = el.kind; de.q_off = el.q_offset; de.c_off = const_total; de.row = row; de.j_off = joff;
            for (int r = 0; r < elem_nn(el.kind); r++)
                if (row + r < MAX_ROWS) dm.row_elem[i][row + r] = (unsigned char)e;
            if (zoff + row + elem_nn(el.kind) <= MAX_TOTAL_ROWS && joff + elem_nj(el.kind) <= 127 && el.q_offset + elem_nq(el.kind) <= 255)
                probe_rows(el.kind, &dm.rows[zoff + row], el.q_offset, joff);
            else
                m->rows_ok = false;
            // derived constants for every instance
            const int nc = elem_nc(el.kind);
            for (int k = 0; k < nc; k++) h_consts.emplace_back((size_t)count);
            for (int64_t b = 0; b < count; b++) {
                const double* P = sd.params.ptr + (sd.params.stride ? (first + b) * sd.params.stride : 0) + el.param_offset;
                double C[24];
                prep_consts(el.kind, P, C);
                for (int k = 0; k < nc; k++) h_consts[const_total + k][b] = C[k];
            }
            const_total += nc;
            row += elem_nn(el.kind);
            joff += elem_nj(el.kind);
            qcheck += elem_nq(el.kind);
        }
        if (row != sd.nn) { delete m; return fail(ACMEB200_EINVAL, "sub %d: elements provide %d equations, nn = %d", i, row, sd.nn); }
        (void)qcheck;
        s.njv = joff;
        elem_total += sd.nelem;
        initz_total += sd.nn;
        zoff += sd.nn;
        max_nn = std::max(max_nn, sd.nn); max_nq = std::max(max_nq, sd.nq);
        max_np = std::max(max_np, sd.np); max_njv = std::max(max_njv, joff);
        if (sd.init_z.stride || sd.params.stride) { /* per-instance state only; matrices may still be shared */ }
    }
    dm.nelem_total = elem_total;
    m->max_nn = max_nn;
    for (int i = 0; i < d->nsub; i++) m->max_nelem = std::max(m->max_nelem, d->subs[i].nelem);
    dm.nconst = const_total;
    dm.ninitz = initz_total;
    dm.blob_len = off;
    for (const Piece& p : pieces) if (p.arr.stride != 0 && p.len > 0) per_instance = true;
    for (const Piece& p : pieces)
        if (p.len > 0 && !p.arr.ptr) { delete m; return fail(ACMEB200_EINVAL, "null matrix pointer in descriptor"); }

    // ---- generic workspace layout (rows of the [row][instance] array)
    auto rows = [&](int n) { int r = wrow; wrow += n; return r; };
    dm.w_x = rows(dm.nx); dm.w_u = rows(dm.nu); dm.w_zall = rows(nnt); dm.w_xnew = rows(dm.nx);
    dm.w_p = rows(max_np); dm.w_pfull = rows(max_nq); dm.w_q = rows(max_nq); dm.w_res = rows(max_nn);
    dm.w_jv = rows(max_njv); dm.w_z = rows(max_nn); dm.w_tmp = rows(max_nn);
    dm.w_startp = rows(max_np); dm.w_pa = rows(max_np); dm.w_cp = rows(max_np);
    for (int i = 0; i < d->nsub; i++) {
        DevSub& s = dm.subs[i];
        s.w_lastp = rows(s.np); s.w_lastz = rows(s.nn); s.w_lastJp = rows(s.nn * s.np);
        s.w_LU[0] = rows(s.nn * s.nn); s.w_LU[1] = rows(s.nn * s.nn);
        s.w_ipiv[0] = rows(s.nn); s.w_ipiv[1] = rows(s.nn); s.w_sel = rows(1);
    }
    dm.w_rows = wrow;

    // ---- pack + upload
    auto destroy_fail = [&](int code) { acmeb200_model_destroy(m); return code; };
    const int64_t nblob = per_instance ? count : 1;
    std::vector<double> blob((size_t)std::max<int64_t>(2, (int64_t)off * nblob + 2));  // +pad: TMA copies an even count
    for (int64_t b = 0; b < nblob; b++)
        for (const Piece& p : pieces) {
            if (p.len == 0) continue;
            const double* src = p.arr.ptr + (p.arr.stride ? (first + b) * p.arr.stride : 0);
            memcpy(blob.data() + b * off + *p.off, src, sizeof(double) * (size_t)p.len);
        }
    m->blob_stride = per_instance ? off : 0;
    m->h_blob.assign(blob.begin(), blob.begin() + std::max(off, 1));
#define CT(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fail(e_ == cudaErrorMemoryAllocation ? ACMEB200_ENOMEM : ACMEB200_ECUDA, "%s failed: %s", #x, cudaGetErrorString(e_)); return destroy_fail(e_ == cudaErrorMemoryAllocation ? ACMEB200_ENOMEM : ACMEB200_ECUDA); } } while (0)
    CT(cudaMalloc(&m->d_blob, sizeof(double) * blob.size()));
    CT(cudaMemcpy(m->d_blob, blob.data(), sizeof(double) * blob.size(), cudaMemcpyHostToDevice));
    {
        std::vector<double> flat((size_t)std::max<int64_t>(1, (int64_t)const_total * count));
        for (int k = 0; k < const_total; k++) memcpy(flat.data() + (size_t)k * count, h_consts[k].data(), sizeof(double) * (size_t)count);
        CT(cudaMalloc(&m->d_consts, sizeof(double) * flat.size()));
        CT(cudaMemcpy(m->d_consts, flat.data(), sizeof(double) * flat.size(), cudaMemcpyHostToDevice));
    }
    {
        std::vector<double> flat((size_t)std::max<int64_t>(1, (int64_t)initz_total * count));
        for (int i = 0; i < d->nsub; i++) {
            const acmeb200_sub_desc& sd = d->subs[i];
            for (int k = 0; k < sd.nn; k++)
                for (int64_t b = 0; b < count; b++)
                    flat[(size_t)(dm.subs[i].o_initz + k) * count + b] =
                        sd.init_z.ptr ? sd.init_z.ptr[(sd.init_z.stride ? (first + b) * sd.init_z.stride : 0) + k] : 0.0;
        }
        CT(cudaMalloc(&m->d_initz, sizeof(double) * flat.size()));
        CT(cudaMemcpy(m->d_initz, flat.data(), sizeof(double) * flat.size(), cudaMemcpyHostToDevice));
    }
    for (int i = 0; i < d->nsub; i++) {
        // frozen, host-built cache (acmeb200_cache = the fields of KDTree + CachingSolver.zs): uploaded as ONE store image
        // (kdcache.cuh layout) shared by all instances and never written
        const acmeb200_cache& c = d->subs[i].cache;
        DevSub& s = dm.subs[i];
        s.kd_base = nullptr; s.kd_scr = nullptr; s.kd_stride = 0; s.kd_sstride = 0; s.kd_cap = 0; s.kd_frozen = 0;
        if (c.n_points <= 0) continue;
        if (c.n_points > (1 << 20) || c.n_columns > (1 << 20)) return destroy_fail(fail(ACMEB200_EUNSUPPORTED, "cache with %d points", c.n_points));
        if (s.np > KD_MAXNP) return destroy_fail(fail(ACMEB200_EUNSUPPORTED, "solution caches support np <= %d", KD_MAXNP));
        for (int k = 0; k < c.n_points; k++)
            if (c.ps_idx[k] < 1 || c.ps_idx[k] > c.n_columns) return destroy_fail(fail(ACMEB200_EINVAL, "cache index out of range"));
        for (int k = 0; k < c.n_points - 1; k++)
            if (c.cut_dim[k] < 1 || c.cut_dim[k] > s.np) return destroy_fail(fail(ACMEB200_EINVAL, "cache cut dimension out of range"));
        const int cap = c.n_columns;
        std::vector<double> img((size_t)kd_store_doubles(s.np, s.nn, cap), 0.0);
        KdStore st = KdStore::at(img.data(), nullptr, s.np, s.nn, cap);
        st.hdr[KD_H_NUM] = c.n_columns; st.hdr[KD_H_NEW] = 0; st.hdr[KD_H_LIMIT] = 2 * cap; st.hdr[KD_H_CAPREF] = cap;
        st.hdr[KD_H_TREEN] = c.n_points; st.hdr[KD_H_FLAGS] = KD_F_FROZEN;
        for (int k = 0; k < c.n_points - 1; k++) st.set_cut(k + 1, c.cut_dim[k], c.cut_val[k]);
        for (int k = 0; k < c.n_points; k++) st.set_psidx(k + 1, c.ps_idx[k]);
        for (int k = 0; k < cap; k++) {
            for (int j = 0; j < s.np; j++) st.col(k + 1)[j] = c.ps[(size_t)k * s.np + j];
            for (int j = 0; j < s.nn; j++) st.col(k + 1)[s.np + j] = c.zs[(size_t)k * s.nn + j];
        }
        void* p = nullptr;
        CT(cudaMalloc(&p, sizeof(double) * img.size()));
        m->d_cache.push_back(p);
        CT(cudaMemcpy(p, img.data(), sizeof(double) * img.size(), cudaMemcpyHostToDevice));
        s.kd_base = (double*)p; s.kd_cap = cap; s.kd_frozen = 1;
        s.kd_mir = nullptr; s.kd_mld = 1; s.kd_mshared = 1;
        if (s.np > 0) {  // shared leaf mirror (devmodel.h): the tree's points in leaf order, doubles [leaf][dimension]
            std::vector<double> mir((size_t)std::max(c.n_points, 1) * s.np, 0.0);
            for (int k = 0; k < c.n_points; k++)
                for (int j = 0; j < s.np; j++) mir[(size_t)k * s.np + j] = c.ps[(size_t)(c.ps_idx[k] - 1) * s.np + j];
            void* pm = nullptr;
            CT(cudaMalloc(&pm, sizeof(double) * mir.size()));
            m->d_cache.push_back(pm);
            CT(cudaMemcpy(pm, mir.data(), sizeof(double) * mir.size(), cudaMemcpyHostToDevice));
            s.kd_mir = (double*)pm;
        }
        m->has_cache = true;
    }
    CT(cudaMalloc(&m->d_status, sizeof(uint32_t) * (size_t)count));
    CT(cudaMalloc(&m->d_first_fail, sizeof(long long) * (size_t)count));
    CT(cudaMalloc(&m->d_stats, sizeof(DevStats)));
    CT(cudaMemset(m->d_stats, 0, sizeof(DevStats)));
#undef CT
    int rc = select_kernel(m);
    if (rc) return destroy_fail(rc);
    *out = m;
    return ACMEB200_OK;
}

static int select_kernel(acmeb200_model* m) {
    m->tpi = nullptr;
    m->coop_lanes = 0;
    m->rows = 0;
    if (m->kernel_mode == 0) m->tpi = find_tpi(m->dm);
    if (!m->tpi && (m->kernel_mode == 0 || m->kernel_mode == 3)) {
        // one warp per instance, LU rows in registers: compile-time shapes only
        if (m->rows_ok) m->rows = rows_shape(m->dm);  // shared or per-instance matrices, learning or frozen solution store
        if (m->kernel_mode == 3 && !m->rows)
            return fail(ACMEB200_EUNSUPPORTED, "the rows-in-registers kernel has no instantiation for this model shape");
    }
    if (!m->tpi && !m->rows && m->kernel_mode != 1) {
        const int lanes = coop_lanes_for(m);
        // automatic choice: large non-linear systems profit from lanes sharing one instance
        if (lanes && (m->kernel_mode == 2 || m->max_nn >= 4)) m->coop_lanes = lanes;
        m->coop_static = (m->coop_lanes >= 16 && coop_static_matches(m->dm)) ? 1 : 0;
        if (m->kernel_mode == 2 && !m->coop_lanes)
            return fail(ACMEB200_EUNSUPPORTED, "the cooperative kernel needs shared matrices and nn <= %d", MAX_ROWS);
    }
    m->ws_rows = m->tpi ? m->tpi->state_rows : m->dm.w_rows;
    if (m->tpi) m->kernel_name = m->tpi->name;
    else if (m->rows) m->kernel_name = std::string("rows<warp per instance, LU rows in registers, compile-time dims ") +
                                       (m->rows == 1 ? "[superover]" : "[superover, baked pots]") + (m->blob_stride ? ", per-instance matrices" : "") +
                                       (m->dm.solver == ACMEB200_SOLVER_HOMOTOPY_CACHING ? ", dynamic solution cache>" : ">");
    else if (m->coop_lanes) m->kernel_name = "coop<" + std::to_string(m->coop_lanes) + " lanes per instance, " +
                                            (m->coop_static ? "compile-time dims [superover]" : "runtime dims") + ", state in shared memory" +
                                            (m->dm.solver == ACMEB200_SOLVER_HOMOTOPY_CACHING ? ", dynamic solution cache>" : ">");
    else m->kernel_name = "generic<thread-per-instance, runtime dims>";
    // learning solution stores of the CachingSolver (kdcache.cuh), one per instance and sub-problem
    for (void* p : m->d_dyn) cudaFree(p);
    m->d_dyn.clear();
    m->dyn_bytes.clear();
    for (int i = 0; i < m->dm.nsub; i++) {
        DevSub& s = m->dm.subs[i];
        if (s.kd_frozen) continue;
        s.kd_base = nullptr; s.kd_scr = nullptr; s.kd_stride = 0; s.kd_sstride = 0; s.kd_cap = 0;
        s.kd_mir = nullptr; s.kd_mld = 0; s.kd_mshared = 0;
        if (m->dm.solver != ACMEB200_SOLVER_HOMOTOPY_CACHING || s.np > KD_MAXNP) continue;
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        // Physical capacity in stored solutions per instance.  The reference's arrays double for ever (solvers.jl:376-382;
        // superover stores ~900 solutions in its first second and ~150 per second after that), and a store that is full
        // stops learning, which costs iterations from then on (measured at capacity 1024 on config 4: 3.2 -> 8.6 iterations
        // per solve once the stores were full).  So: as large as a generous memory budget allows (a third of the free
        // memory, at most 32 GB -- 9 GB for the 8192 superover instances at 4096 solutions each), at most 4096; the
        // descriptor (cache_capacity) or ACMEB200_CACHE_CAP override it
        int cap = 4096;
        const size_t per_col = 8 * (size_t)(s.np + s.nn) + 16 + 24 + 8 * (size_t)s.np;
        const size_t budget = std::min<size_t>(free_b / 3, (size_t)32 << 30);
        while (cap > 64 && (size_t)m->B * cap * per_col > budget) cap /= 2;
        if (m->cache_capacity > 0) cap = m->cache_capacity;
        if (const char* e = getenv("ACMEB200_CACHE_CAP")) { const long v = atol(e); if (v >= 2 && v <= (1 << 20)) cap = (int)v; }  // tuning / test knob
        const int64_t stride = (kd_store_doubles(s.np, s.nn, cap) + 1) & ~int64_t(1), sstride = (kd_scratch_doubles(cap) + 1) & ~int64_t(1);
        void *st = nullptr, *sc = nullptr;
        CUDA_TRY(cudaMalloc(&st, sizeof(double) * (size_t)stride * m->B));
        m->d_dyn.push_back(st); m->dyn_bytes.push_back(sizeof(double) * (size_t)stride * m->B);
        CUDA_TRY(cudaMalloc(&sc, sizeof(double) * (size_t)sstride * m->B));
        m->d_dyn.push_back(sc); m->dyn_bytes.push_back(0);  // scratch needs no clearing
        s.kd_base = (double*)st; s.kd_stride = stride; s.kd_scr = (double*)sc; s.kd_sstride = sstride; s.kd_cap = cap;
        if (m->rows && s.np > 0) {  // leaf mirror of the warp-per-instance kernel (kernel_rows.cuh: centre, radius, floats [dimension][leaf])
            void* mir = nullptr;
            const int64_t mstride = rows_mirror_doubles(m->rows, cap);
            const size_t mbytes = sizeof(double) * (size_t)mstride * m->B;
            CUDA_TRY(cudaMalloc(&mir, mbytes));
            m->d_dyn.push_back(mir); m->dyn_bytes.push_back(mbytes);
            s.kd_mir = (double*)mir; s.kd_mld = mstride; s.kd_mshared = 0;
        }
        if (m->tpi && s.np > 0) {  // leaf mirror of the thread-per-instance kernels (devmodel.h), [leaf][dimension][instance]
            void* mir = nullptr;
            const size_t mbytes = sizeof(double) * (size_t)cap * s.np * m->B;
            CUDA_TRY(cudaMalloc(&mir, mbytes));
            m->d_dyn.push_back(mir); m->dyn_bytes.push_back(mbytes);
            s.kd_mir = (double*)mir; s.kd_mld = m->B; s.kd_mshared = 0;
        }
    }
    cudaFree(m->d_ws);
    m->d_ws = nullptr;
    CUDA_TRY(cudaMalloc(&m->d_ws, sizeof(double) * (size_t)std::max<int64_t>(1, m->ws_rows * m->B)));
    CUDA_TRY(cudaMemset(m->d_ws, 0, sizeof(double) * (size_t)std::max<int64_t>(1, m->ws_rows * m->B)));
    return run_init(m);
}

static RunArgs base_args(acmeb200_model* m) {
    RunArgs a;
    memset(&a, 0, sizeof a);
    a.blob = m->d_blob; a.blob_stride = m->blob_stride; a.consts = m->d_consts; a.initz = m->d_initz;
    a.ws = m->d_ws; a.ld = m->B; a.status = m->d_status; a.first_fail = m->d_first_fail; a.stats = m->d_stats;
    a.n_done = m->n_done;
    return a;
}

static cudaError_t launch(acmeb200_model* m, const RunArgs& a, cudaStream_t stream) {
    m->launches++;
    if (m->tpi) return m->tpi->launch(m, a, stream);
    if (m->rows && !a.init) return launch_rows_kernel(m, a, stream);  // the generic kernel initialises the (shared) state layout
    if (m->coop_lanes && !a.init) return launch_coop_kernel(m, a, stream);
    const int tpb = 128;
    ACME_LAUNCH(k_generic, (unsigned)((a.ninst + tpb - 1) / tpb), tpb, 0, stream, m->dm, a);
    return cudaGetLastError();
}

static int run_init(acmeb200_model* m) {
    RunArgs a = base_args(m);
    a.init = 1; a.inst0 = 0; a.ninst = m->B; a.N = 0;
    // the spare columns of a learning store are the zero-filled spare capacity of the reference's arrays
    for (size_t k = 0; k < m->d_dyn.size(); k++)
        if (m->dyn_bytes[k]) CUDA_TRY(cudaMemsetAsync(m->d_dyn[k], 0, m->dyn_bytes[k], nullptr));
    CUDA_TRY(launch(m, a, nullptr));
    CUDA_TRY(cudaMemsetAsync(m->d_stats, 0, sizeof(DevStats), nullptr));
    CUDA_TRY(cudaDeviceSynchronize());
    m->n_done = 0;
    return ACMEB200_OK;
}

extern "C" int acmeb200_reset(acmeb200_model* m) {
    if (!m) return fail(ACMEB200_EINVAL, "null model");
    CUDA_TRY(cudaSetDevice(m->device));
    return run_init(m);
}

extern "C" int acmeb200_set_kernel(acmeb200_model* m, int32_t mode) {
    if (!m) return fail(ACMEB200_EINVAL, "null model");
    if (mode < 0 || mode > 3) return fail(ACMEB200_EINVAL, "kernel mode must be 0 (auto), 1 (generic), 2 (cooperative) or 3 (rows in registers)");
    CUDA_TRY(cudaSetDevice(m->device));
    m->kernel_mode = mode;
    return select_kernel(m);
}
extern "C" const char* acmeb200_kernel_name(const acmeb200_model* m) { return m ? m->kernel_name.c_str() : ""; }
extern "C" int64_t acmeb200_launch_count(const acmeb200_model* m) { return m ? m->launches : 0; }

// ------------------------------------------------------------------ run
static int ensure_staging(acmeb200_model* m, size_t ubytes, size_t ybytes) {
    for (int i = 0; i < 2; i++) {
        if (!m->copy_streams[i]) CUDA_TRY(cudaStreamCreateWithFlags(&m->copy_streams[i], cudaStreamNonBlocking));
        if (!m->ev[i]) CUDA_TRY(cudaEventCreateWithFlags(&m->ev[i], cudaEventDisableTiming));
    }
    if (ubytes > m->stage_u_bytes) {
        for (int i = 0; i < 2; i++) { cudaFree(m->d_stage_u[i]); m->d_stage_u[i] = nullptr; CUDA_TRY(cudaMalloc(&m->d_stage_u[i], ubytes)); }
        m->stage_u_bytes = ubytes;
    }
    if (ybytes > m->stage_y_bytes) {
        for (int i = 0; i < 2; i++) { cudaFree(m->d_stage_y[i]); m->d_stage_y[i] = nullptr; CUDA_TRY(cudaMalloc(&m->d_stage_y[i], ybytes)); }
        m->stage_y_bytes = ybytes;
    }
    return ACMEB200_OK;
}

extern "C" int acmeb200_run(acmeb200_model* m, const double* U, int64_t u_stride, double* Y,
                            int64_t y_stride, int64_t N, uint32_t flags, void* stream_) {
    if (!m) return fail(ACMEB200_EINVAL, "null model");
    if (N < 0) return fail(ACMEB200_EINVAL, "negative sample count");
    if (N >= (1ll << 27)) return fail(ACMEB200_EUNSUPPORTED, "at most 2^27-1 samples per call; split the run");
    const DevModel& dm = m->dm;
    const bool smaj = (flags & ACMEB200_SAMPLE_MAJOR) != 0;  // strides are sample pitches, (nu, B, N) / (ny, B, N) streams
    if (smaj) {
        if (!m->tpi && (m->rows || m->coop_lanes))
            return fail(ACMEB200_EUNSUPPORTED, "sample-major streams are implemented by the thread-per-instance kernels only; this model runs on %s",
                        m->kernel_name.c_str());
        if (y_stride == 0) y_stride = (int64_t)dm.ny * m->B;
        if (u_stride != 0 && u_stride < (int64_t)dm.nu * m->B) return fail(ACMEB200_EINVAL, "sample-major u_stride smaller than nu*B");
        if (y_stride < (int64_t)dm.ny * m->B) return fail(ACMEB200_EINVAL, "sample-major y_stride smaller than ny*B");
    } else {
        if (y_stride == 0) y_stride = (int64_t)dm.ny * N;
        if (u_stride != 0 && u_stride < (int64_t)dm.nu * N) return fail(ACMEB200_EINVAL, "u_stride smaller than nu*N");
        if (y_stride < (int64_t)dm.ny * N) return fail(ACMEB200_EINVAL, "y_stride smaller than ny*N");
    }
    if (N == 0) return ACMEB200_OK;
    if ((dm.nu > 0 && !U) || (dm.ny > 0 && !Y)) return fail(ACMEB200_EINVAL, "null stream pointer");
    CUDA_TRY(cudaSetDevice(m->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    const bool udev = (flags & ACMEB200_U_DEVICE) || dm.nu == 0, ydev = (flags & ACMEB200_Y_DEVICE) || dm.ny == 0;

    if (udev && ydev) {
        RunArgs a = base_args(m);
        a.U = U; a.u_stride = u_stride; a.Y = Y; a.y_stride = y_stride; a.N = N; a.inst0 = 0; a.ninst = m->B;
        a.smaj = smaj;
        CUDA_TRY(launch(m, a, stream));
        m->n_done += N;
        return ACMEB200_OK;
    }

    // ---- host streams: chunks of TIME (all instances per chunk keep the whole GPU busy and the
    // per-instance state simply persists from chunk to chunk), three-stage pipeline
    // H2D(c+1) | kernel(c) | D2H(c-1) on separate streams with double-buffered staging
    const size_t row_u = (size_t)dm.nu * sizeof(double), row_y = (size_t)dm.ny * sizeof(double);
    const bool shared_u = (u_stride == 0) || dm.nu == 0;
    const size_t per_sample = std::max<size_t>(8, (size_t)m->B * std::max(shared_u ? 0 : row_u, row_y));
    size_t chunk_mb = 1024;  // staging chunk size; ACMEB200_CHUNK_MB overrides (tuning knob)
    if (const char* e = getenv("ACMEB200_CHUNK_MB")) { const long v = atol(e); if (v >= 1 && v <= 8192) chunk_mb = (size_t)v; }
    int64_t Tc = (int64_t)((chunk_mb << 20) / per_sample);
    Tc -= Tc % 16;
    Tc = std::max<int64_t>(16, std::min<int64_t>(Tc, N));
    const size_t ubytes = udev ? 8 : (shared_u ? std::max<size_t>(8, row_u * Tc) : row_u * Tc * m->B);
    const size_t ybytes = ydev ? 8 : std::max<size_t>(8, row_y * Tc * m->B);
    int rc = ensure_staging(m, ubytes, ybytes);
    if (rc) return rc;
    if (!m->compute_stream) CUDA_TRY(cudaStreamCreateWithFlags(&m->compute_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
        if (!m->ev_in[i]) CUDA_TRY(cudaEventCreateWithFlags(&m->ev_in[i], cudaEventDisableTiming));
        if (!m->ev_k[i]) CUDA_TRY(cudaEventCreateWithFlags(&m->ev_k[i], cudaEventDisableTiming));
        if (!m->ev_out[i]) CUDA_TRY(cudaEventCreateWithFlags(&m->ev_out[i], cudaEventDisableTiming));
    }
    CUDA_TRY(cudaStreamSynchronize(stream));
    cudaStream_t s_in = m->copy_streams[0], s_out = m->copy_streams[1], s_k = m->compute_stream;
    int64_t c = 0;
    for (int64_t n0 = 0; n0 < N; n0 += Tc, c++) {
        const int buf = (int)(c & 1);
        const int64_t nt = std::min<int64_t>(Tc, N - n0);
        RunArgs a = base_args(m);
        a.N = nt; a.inst0 = 0; a.ninst = m->B;
        a.smaj = smaj;
        a.n_done = m->n_done + n0;
        // sample-major streams: a chunk of time is `nt` whole rows of B instances, staged densely
        const int64_t u_step = smaj && !shared_u ? u_stride : dm.nu, y_step = smaj ? y_stride : dm.ny;  // doubles per sample
        // ---- H2D of chunk c into staging buffer `buf` (free once kernel c-2 has consumed it)
        if (!udev && dm.nu > 0) {
            if (c >= 2) CUDA_TRY(cudaStreamWaitEvent(s_in, m->ev_k[buf], 0));
            if (shared_u)
                CUDA_TRY(cudaMemcpyAsync(m->d_stage_u[buf], U + n0 * dm.nu, row_u * nt, cudaMemcpyHostToDevice, s_in));
            else if (smaj)
                CUDA_TRY(cudaMemcpy2DAsync(m->d_stage_u[buf], row_u * m->B, U + n0 * u_stride, (size_t)u_stride * sizeof(double),
                                           row_u * m->B, (size_t)nt, cudaMemcpyHostToDevice, s_in));
            else
                CUDA_TRY(cudaMemcpy2DAsync(m->d_stage_u[buf], row_u * nt, U + n0 * dm.nu, (size_t)u_stride * sizeof(double),
                                           row_u * nt, (size_t)m->B, cudaMemcpyHostToDevice, s_in));
            CUDA_TRY(cudaEventRecord(m->ev_in[buf], s_in));
            CUDA_TRY(cudaStreamWaitEvent(s_k, m->ev_in[buf], 0));
            a.U = m->d_stage_u[buf];
            a.u_stride = shared_u ? 0 : (smaj ? dm.nu * m->B : dm.nu * nt);
        } else {
            a.U = U ? U + n0 * u_step : nullptr;
            a.u_stride = u_stride;
        }
        // ---- kernel c (needs the output staging buffer `buf` drained by D2H c-2)
        if (!ydev && dm.ny > 0) {
            if (c >= 2) CUDA_TRY(cudaStreamWaitEvent(s_k, m->ev_out[buf], 0));
            a.Y = m->d_stage_y[buf];
            a.y_stride = smaj ? dm.ny * m->B : dm.ny * nt;
        } else {
            a.Y = Y ? Y + n0 * y_step : nullptr;
            a.y_stride = y_stride;
        }
        CUDA_TRY(launch(m, a, s_k));
        CUDA_TRY(cudaEventRecord(m->ev_k[buf], s_k));
        // ---- D2H of chunk c
        if (!ydev && dm.ny > 0) {
            CUDA_TRY(cudaStreamWaitEvent(s_out, m->ev_k[buf], 0));
            if (smaj)
                CUDA_TRY(cudaMemcpy2DAsync(Y + n0 * y_stride, (size_t)y_stride * sizeof(double), m->d_stage_y[buf], row_y * m->B,
                                           row_y * m->B, (size_t)nt, cudaMemcpyDeviceToHost, s_out));
            else
                CUDA_TRY(cudaMemcpy2DAsync(Y + n0 * dm.ny, (size_t)y_stride * sizeof(double), m->d_stage_y[buf], row_y * nt,
                                           row_y * nt, (size_t)m->B, cudaMemcpyDeviceToHost, s_out));
            CUDA_TRY(cudaEventRecord(m->ev_out[buf], s_out));
        }
    }
    CUDA_TRY(cudaStreamSynchronize(s_in));
    CUDA_TRY(cudaStreamSynchronize(s_k));
    CUDA_TRY(cudaStreamSynchronize(s_out));
    m->n_done += N;
    return ACMEB200_OK;
}

// ------------------------------------------------------------------ state / statistics
static int state_row_of_x(const acmeb200_model* m) { return m->tpi ? 0 : m->dm.w_x; }  // generic and coop share the layout

extern "C" int acmeb200_get_state(acmeb200_model* m, double* x_host) {
    if (!m || !x_host) return fail(ACMEB200_EINVAL, "null argument");
    CUDA_TRY(cudaSetDevice(m->device));
    CUDA_TRY(cudaDeviceSynchronize());
    const int nx = m->dm.nx;
    if (nx == 0) return ACMEB200_OK;
    std::vector<double> rows((size_t)nx * m->B);
    CUDA_TRY(cudaMemcpy(rows.data(), m->d_ws + (size_t)state_row_of_x(m) * m->B, sizeof(double) * rows.size(), cudaMemcpyDeviceToHost));
    for (int64_t b = 0; b < m->B; b++)
        for (int i = 0; i < nx; i++) x_host[b * nx + i] = rows[(size_t)i * m->B + b];
    return ACMEB200_OK;
}

extern "C" int acmeb200_set_state(acmeb200_model* m, const double* x_host, int64_t stride) {
    if (!m || !x_host) return fail(ACMEB200_EINVAL, "null argument");
    CUDA_TRY(cudaSetDevice(m->device));
    CUDA_TRY(cudaDeviceSynchronize());
    const int nx = m->dm.nx;
    if (nx == 0) return ACMEB200_OK;
    std::vector<double> rows((size_t)nx * m->B);
    for (int64_t b = 0; b < m->B; b++)
        for (int i = 0; i < nx; i++) rows[(size_t)i * m->B + b] = x_host[b * stride + i];
    CUDA_TRY(cudaMemcpy(m->d_ws + (size_t)state_row_of_x(m) * m->B, rows.data(), sizeof(double) * rows.size(), cudaMemcpyHostToDevice));
    return ACMEB200_OK;
}

extern "C" int acmeb200_get_status(acmeb200_model* m, uint32_t* status_host, int64_t* first_fail_host) {
    if (!m) return fail(ACMEB200_EINVAL, "null model");
    CUDA_TRY(cudaSetDevice(m->device));
    CUDA_TRY(cudaDeviceSynchronize());
    if (status_host) CUDA_TRY(cudaMemcpy(status_host, m->d_status, sizeof(uint32_t) * (size_t)m->B, cudaMemcpyDeviceToHost));
    if (first_fail_host) CUDA_TRY(cudaMemcpy(first_fail_host, m->d_first_fail, sizeof(long long) * (size_t)m->B, cudaMemcpyDeviceToHost));
    return ACMEB200_OK;
}

extern "C" int acmeb200_get_stats(acmeb200_model* m, acmeb200_stats* out) {
    if (!m || !out) return fail(ACMEB200_EINVAL, "null argument");
    CUDA_TRY(cudaSetDevice(m->device));
    CUDA_TRY(cudaDeviceSynchronize());
    DevStats s;
    CUDA_TRY(cudaMemcpy(&s, m->d_stats, sizeof s, cudaMemcpyDeviceToHost));
    out->samples = s.samples; out->solves = s.solves; out->newton_iters = s.newton_iters;
    out->homotopy_solves = s.homotopy_solves; out->not_converged = s.not_converged;
    for (int i = 0; i < ACMEB200_HIST_BINS; i++) out->iter_hist[i] = s.iter_hist[i];
    return ACMEB200_OK;
}

extern "C" int acmeb200_get_cache_sizes(acmeb200_model* m, int32_t sub, int32_t* sizes_host, int32_t* capacity_out) {
    if (!m || !sizes_host) return fail(ACMEB200_EINVAL, "null argument");
    if (sub < 0 || sub >= m->dm.nsub) return fail(ACMEB200_EINVAL, "sub-problem %d out of range", sub);
    CUDA_TRY(cudaSetDevice(m->device));
    CUDA_TRY(cudaDeviceSynchronize());
    const DevSub& s = m->dm.subs[sub];
    if (capacity_out) *capacity_out = s.kd_cap;
    if (s.kd_cap > 0 && !s.kd_frozen) {  // hdr[KD_H_NUM] of every instance's store
        CUDA_TRY(cudaMemcpy2DAsync(sizes_host, sizeof(int32_t), reinterpret_cast<const int*>(s.kd_base) + KD_H_NUM, sizeof(double) * (size_t)s.kd_stride,
                                   sizeof(int32_t), (size_t)m->B, cudaMemcpyDeviceToHost, nullptr));
        CUDA_TRY(cudaDeviceSynchronize());
    } else if (s.kd_cap > 0) {
        int32_t n = 0;
        CUDA_TRY(cudaMemcpy(&n, reinterpret_cast<const int*>(s.kd_base) + KD_H_NUM, sizeof n, cudaMemcpyDeviceToHost));
        for (int64_t b = 0; b < m->B; b++) sizes_host[b] = n;
    } else
        memset(sizes_host, 0, sizeof(int32_t) * (size_t)m->B);
    return ACMEB200_OK;
}

extern "C" int acmeb200_get_cache_info(acmeb200_model* m, int32_t sub, int32_t* info_host) {
    if (!m || !info_host) return fail(ACMEB200_EINVAL, "null argument");
    if (sub < 0 || sub >= m->dm.nsub) return fail(ACMEB200_EINVAL, "sub-problem %d out of range", sub);
    CUDA_TRY(cudaSetDevice(m->device));
    CUDA_TRY(cudaDeviceSynchronize());
    const DevSub& s = m->dm.subs[sub];
    memset(info_host, 0, sizeof(int32_t) * KD_HDR_INTS * (size_t)m->B);
    if (s.kd_cap <= 0) return ACMEB200_OK;
    if (!s.kd_frozen) {
        CUDA_TRY(cudaMemcpy2DAsync(info_host, sizeof(int32_t) * KD_HDR_INTS, s.kd_base, sizeof(double) * (size_t)s.kd_stride,
                                   sizeof(int32_t) * KD_HDR_INTS, (size_t)m->B, cudaMemcpyDeviceToHost, nullptr));
        CUDA_TRY(cudaDeviceSynchronize());
    } else {
        CUDA_TRY(cudaMemcpy(info_host, s.kd_base, sizeof(int32_t) * KD_HDR_INTS, cudaMemcpyDeviceToHost));
        for (int64_t b = 1; b < m->B; b++) memcpy(info_host + b * KD_HDR_INTS, info_host, sizeof(int32_t) * KD_HDR_INTS);
    }
    return ACMEB200_OK;
}

// ------------------------------------------------------------------ solver state (checkpoint / deepcopy)
namespace {
struct StateHeader {
    uint64_t magic;
    int32_t abi, layout;  // layout: 1 = thread-per-instance kernel state rows, 0 = generic workspace
    int64_t B, ws_rows, n_done, total_bytes;
    int32_t nx, nu, ny, nsub, nnt, solver;
    int32_t np[MAX_SUBS], nn[MAX_SUBS], kd_cap[MAX_SUBS];
    int64_t kd_stride[MAX_SUBS];
};
constexpr uint64_t STATE_MAGIC = 0x41434d4542323030ull;  // "ACMEB200"

StateHeader state_header(const acmeb200_model* m) {
    StateHeader h;
    memset(&h, 0, sizeof h);
    h.magic = STATE_MAGIC; h.abi = ACMEB200_ABI_VERSION; h.layout = m->tpi ? 1 : 0;
    h.B = m->B; h.ws_rows = m->ws_rows; h.n_done = m->n_done;
    h.nx = m->dm.nx; h.nu = m->dm.nu; h.ny = m->dm.ny; h.nsub = m->dm.nsub; h.nnt = m->dm.nnt; h.solver = m->dm.solver;
    int64_t bytes = sizeof(StateHeader) + sizeof(double) * m->ws_rows * m->B + sizeof(uint32_t) * m->B + sizeof(long long) * m->B + sizeof(DevStats);
    for (int i = 0; i < m->dm.nsub; i++) {
        const DevSub& s = m->dm.subs[i];
        h.np[i] = s.np; h.nn[i] = s.nn;
        if (s.kd_cap > 0 && !s.kd_frozen) {  // a frozen store belongs to the descriptor, not to the state
            h.kd_cap[i] = s.kd_cap; h.kd_stride[i] = s.kd_stride;
        }
    }
    for (size_t k = 0; k < m->d_dyn.size(); k++) bytes += (int64_t)m->dyn_bytes[k];  // learning stores and their leaf mirrors (scratch: 0)
    h.total_bytes = bytes;
    return h;
}
}  // namespace

extern "C" int64_t acmeb200_solver_state_size(acmeb200_model* m) {
    if (!m) return fail(ACMEB200_EINVAL, "null model");
    return state_header(m).total_bytes;
}

extern "C" int acmeb200_get_solver_state(acmeb200_model* m, void* buf, int64_t bytes) {
    if (!m || !buf) return fail(ACMEB200_EINVAL, "null argument");
    const StateHeader h = state_header(m);
    if (bytes < h.total_bytes) return fail(ACMEB200_EINVAL, "buffer of %lld bytes, the state needs %lld", (long long)bytes, (long long)h.total_bytes);
    CUDA_TRY(cudaSetDevice(m->device));
    CUDA_TRY(cudaDeviceSynchronize());
    char* out = static_cast<char*>(buf);
    memcpy(out, &h, sizeof h); out += sizeof h;
    auto take = [&](const void* dsrc, size_t n) -> cudaError_t { const cudaError_t e = n ? cudaMemcpy(out, dsrc, n, cudaMemcpyDeviceToHost) : cudaSuccess; out += n; return e; };
    CUDA_TRY(take(m->d_ws, sizeof(double) * (size_t)(m->ws_rows * m->B)));
    CUDA_TRY(take(m->d_status, sizeof(uint32_t) * (size_t)m->B));
    CUDA_TRY(take(m->d_first_fail, sizeof(long long) * (size_t)m->B));
    CUDA_TRY(take(m->d_stats, sizeof(DevStats)));
    for (size_t k = 0; k < m->d_dyn.size(); k++) CUDA_TRY(take(m->d_dyn[k], m->dyn_bytes[k]));
    return ACMEB200_OK;
}

extern "C" int acmeb200_set_solver_state(acmeb200_model* m, const void* buf, int64_t bytes) {
    if (!m || !buf) return fail(ACMEB200_EINVAL, "null argument");
    if (bytes < (int64_t)sizeof(StateHeader)) return fail(ACMEB200_EINVAL, "not a solver state");
    StateHeader h;
    memcpy(&h, buf, sizeof h);
    if (h.magic != STATE_MAGIC || h.abi != ACMEB200_ABI_VERSION) return fail(ACMEB200_EINVAL, "not a solver state of this ABI version");
    if (bytes < h.total_bytes) return fail(ACMEB200_EINVAL, "truncated solver state");
    CUDA_TRY(cudaSetDevice(m->device));
    // a learning store saved with another physical capacity (the automatic capacity depends on the free memory of the
    // device it was created on): re-create this model's stores with the saved one
    for (int i = 0; i < std::min(h.nsub, m->dm.nsub); i++) {
        const DevSub& s = m->dm.subs[i];
        if (h.kd_cap[i] > 0 && !s.kd_frozen && s.kd_cap != h.kd_cap[i]) {
            m->cache_capacity = h.kd_cap[i];
            const int rc = select_kernel(m);
            if (rc) return rc;
            break;
        }
    }
    const StateHeader mine = state_header(m);
    bool same = h.B == mine.B && h.ws_rows == mine.ws_rows && h.layout == mine.layout && h.nx == mine.nx && h.nu == mine.nu && h.ny == mine.ny &&
                h.nsub == mine.nsub && h.nnt == mine.nnt && h.solver == mine.solver && h.total_bytes == mine.total_bytes;
    for (int i = 0; same && i < h.nsub; i++)
        same = h.np[i] == mine.np[i] && h.nn[i] == mine.nn[i] && h.kd_cap[i] == mine.kd_cap[i] && h.kd_stride[i] == mine.kd_stride[i];
    if (!same)
        return fail(ACMEB200_EINVAL, "the solver state belongs to another model (dimensions, instance count, solver, cache capacity) or was "
                                     "saved with another kernel selected (this model runs on %s)", m->kernel_name.c_str());
    CUDA_TRY(cudaDeviceSynchronize());
    const char* in = static_cast<const char*>(buf) + sizeof h;
    auto put = [&](void* ddst, size_t n) -> cudaError_t { const cudaError_t e = n ? cudaMemcpy(ddst, in, n, cudaMemcpyHostToDevice) : cudaSuccess; in += n; return e; };
    CUDA_TRY(put(m->d_ws, sizeof(double) * (size_t)(m->ws_rows * m->B)));
    CUDA_TRY(put(m->d_status, sizeof(uint32_t) * (size_t)m->B));
    CUDA_TRY(put(m->d_first_fail, sizeof(long long) * (size_t)m->B));
    CUDA_TRY(put(m->d_stats, sizeof(DevStats)));
    for (size_t k = 0; k < m->d_dyn.size(); k++) CUDA_TRY(put(m->d_dyn[k], m->dyn_bytes[k]));
    m->n_done = h.n_done;
    return ACMEB200_OK;
}

extern "C" int acmeb200_get_extrapolation_origin(acmeb200_model* m, int32_t sub, double* p_host, double* z_host) {
    if (!m) return fail(ACMEB200_EINVAL, "null model");
    if (sub < 0 || sub >= m->dm.nsub) return fail(ACMEB200_EINVAL, "sub-problem %d out of range", sub);
    CUDA_TRY(cudaSetDevice(m->device));
    CUDA_TRY(cudaDeviceSynchronize());
    const DevSub& s = m->dm.subs[sub];
    // rows of the [row][instance] state: the thread-per-instance kernels keep x | last_p | last_z | Mx, the others the generic workspace
    const int prow = m->tpi ? m->dm.nx : s.w_lastp, zrow = m->tpi ? m->dm.nx + s.np : s.w_lastz;
    auto rows_to = [&](int row0, int n, double* dst) -> int {
        if (!dst || n == 0) return ACMEB200_OK;
        std::vector<double> r((size_t)n * m->B);
        CUDA_TRY(cudaMemcpy(r.data(), m->d_ws + (size_t)row0 * m->B, sizeof(double) * r.size(), cudaMemcpyDeviceToHost));
        for (int64_t b = 0; b < m->B; b++)
            for (int i = 0; i < n; i++) dst[b * n + i] = r[(size_t)i * m->B + b];
        return ACMEB200_OK;
    };
    int rc = rows_to(prow, s.np, p_host);
    if (rc) return rc;
    return rows_to(zrow, s.nn, z_host);
}

// ------------------------------------------------------------------ element Jacobians of the whole batch
// Jq = d(res)/dq of sub-problem `sub` at one q per instance (CircuitNLFunc's Jq, circuit.jl:10-17), with the
// instance's own element constants: what the batched linearize (ACME.jl:505-550, solvers.jl:407-414) needs at the
// steady state, for all instances in one launch instead of one host evaluation per instance.
__global__ void __launch_bounds__(128) k_eval_jq(const DevModel* mp, int sub, const double* consts, int64_t ld, int64_t first,
                                                 int64_t count, const double* q_in, double* jq_out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const DevModel& m = *mp;
    const DevSub& s = m.subs[sub];
    for (int k = 0; k < s.nn * s.nq; k++) jq_out[(int64_t)k * count + t] = 0.0;
    for (int e = 0; e < s.nelem; e++) {
        const DevElem& el = m.elems[s.elem0 + e];
        double C[20], q[5], res[2], jv[4];
        const int nc = elem_nc(el.kind), nq = elem_nq(el.kind), nn = elem_nn(el.kind);
        for (int k = 0; k < nc; k++) C[k] = consts[(int64_t)(el.c_off + k) * ld + first + t];
        for (int k = 0; k < nq; k++) q[k] = q_in[(int64_t)(el.q_off + k) * count + t];
        elem_eval(el.kind, C, q, res, jv);
        for (int r = 0; r < nn; r++)
            for (int c = 0; c < nq; c++)  // column c of the element's block: the row program applied to the unit vector e_c
                jq_out[((int64_t)(el.q_off + c) * s.nn + el.row + r) * count + t] =
                    elem_row(el.kind, r, jv, [&](int k) { return k == c ? 1.0 : 0.0; });
    }
}

extern "C" int acmeb200_eval_jq(acmeb200_model* m, int32_t sub, const double* q_host, double* jq_host) {
    if (!m) return fail(ACMEB200_EINVAL, "null model");
    if (sub < 0 || sub >= m->dm.nsub) return fail(ACMEB200_EINVAL, "sub-problem %d out of range", sub);
    if (!q_host || !jq_host) return fail(ACMEB200_EINVAL, "null argument");
    CUDA_TRY(cudaSetDevice(m->device));
    const DevSub& s = m->dm.subs[sub];
    const size_t nq = (size_t)s.nq * m->B, nj = (size_t)s.nn * s.nq * m->B;
    double *dq = nullptr, *dj = nullptr;
    DevModel* ddm = nullptr;
    CUDA_TRY(cudaMalloc(&dq, sizeof(double) * std::max<size_t>(nq, 1)));
    CUDA_TRY(cudaMalloc(&dj, sizeof(double) * std::max<size_t>(nj, 1)));
    CUDA_TRY(cudaMalloc(&ddm, sizeof(DevModel)));
    CUDA_TRY(cudaMemcpy(ddm, &m->dm, sizeof(DevModel), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(dq, q_host, sizeof(double) * nq, cudaMemcpyHostToDevice));
    ACME_LAUNCH(k_eval_jq, (unsigned)((m->B + 127) / 128), 128, 0, (cudaStream_t) nullptr, ddm, (int)sub, m->d_consts, m->B, (int64_t)0,
                m->B, dq, dj);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(jq_host, dj, sizeof(double) * nj, cudaMemcpyDeviceToHost));
    cudaFree(dq); cudaFree(dj); cudaFree(ddm);
    return ACMEB200_OK;
}

// ------------------------------------------------------------------ devices, the batch over all GPUs
extern "C" int acmeb200_device_count(int32_t* count_out) {
    if (!count_out) return fail(ACMEB200_EINVAL, "null argument");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) n = 0;
    *count_out = n;
    return ACMEB200_OK;
}
extern "C" int acmeb200_set_device(int32_t device) {
    CUDA_TRY(cudaSetDevice(device));
    return ACMEB200_OK;
}
extern "C" int acmeb200_get_device(acmeb200_model* m, int32_t* device_out) {
    if (!m || !device_out) return fail(ACMEB200_EINVAL, "null argument");
    *device_out = m->device;
    return ACMEB200_OK;
}

struct acmeb200_multi {
    std::vector<acmeb200_model*> shards;
    std::vector<int64_t> first, count;
    int nu = 0, ny = 0;
};

extern "C" void acmeb200_multi_destroy(acmeb200_multi* mm) {
    if (!mm) return;
    for (acmeb200_model* m : mm->shards) acmeb200_model_destroy(m);
    delete mm;
}

extern "C" int acmeb200_multi_create(const acmeb200_model_desc* d, int64_t B, int32_t n_gpus, acmeb200_multi** out) {
    if (!d || !out) return fail(ACMEB200_EINVAL, "null argument");
    *out = nullptr;
    if (B <= 0) return fail(ACMEB200_EINVAL, "no instances");
    int ndev = 0, prev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(ACMEB200_ENODEVICE, "no CUDA device available (this library has no CPU fallback)");
    if (n_gpus <= 0) n_gpus = ndev;
    if (n_gpus > ndev) return fail(ACMEB200_EINVAL, "%d GPUs asked for, %d visible", n_gpus, ndev);
    if ((int64_t)n_gpus > B) n_gpus = (int32_t)B;
    cudaGetDevice(&prev);
    acmeb200_multi* mm = new acmeb200_multi();
    mm->nu = d->nu; mm->ny = d->ny;
    const int64_t base = B / n_gpus, rem = B % n_gpus;
    for (int g = 0; g < n_gpus; g++) {
        const int64_t first = g * base + std::min<int64_t>(g, rem), count = base + (g < rem ? 1 : 0);
        acmeb200_model* m = nullptr;
        int rc = cudaSetDevice(g) == cudaSuccess ? acmeb200_model_create(d, first, count, &m) : fail(ACMEB200_ECUDA, "cudaSetDevice(%d) failed", g);
        if (rc) { acmeb200_multi_destroy(mm); cudaSetDevice(prev); return rc; }
        mm->shards.push_back(m); mm->first.push_back(first); mm->count.push_back(count);
    }
    cudaSetDevice(prev);
    *out = mm;
    return ACMEB200_OK;
}

extern "C" int acmeb200_multi_shards(const acmeb200_multi* mm) { return mm ? (int)mm->shards.size() : 0; }

extern "C" acmeb200_model* acmeb200_multi_model(acmeb200_multi* mm, int32_t shard, int64_t* first_out, int64_t* count_out) {
    if (!mm || shard < 0 || shard >= (int)mm->shards.size()) return nullptr;
    if (first_out) *first_out = mm->first[shard];
    if (count_out) *count_out = mm->count[shard];
    return mm->shards[shard];
}

extern "C" int acmeb200_multi_run(acmeb200_multi* mm, const double* U, int64_t u_stride, double* Y, int64_t y_stride,
                                  int64_t N, uint32_t flags) {
    if (!mm) return fail(ACMEB200_EINVAL, "null argument");
    if (flags & (ACMEB200_U_DEVICE | ACMEB200_Y_DEVICE | ACMEB200_SAMPLE_MAJOR))
        return fail(ACMEB200_EUNSUPPORTED, "acmeb200_multi_run takes instance-major host streams (one device cannot address all shards' device memory)");
    if (y_stride == 0) y_stride = (int64_t)mm->ny * N;
    const size_t G = mm->shards.size();
    std::vector<int> rcs(G, 0);
    std::vector<std::string> errs(G);
    auto work = [&](size_t g) {
        // every shard: its own device, its own pinned staging and streams, its contiguous block of the host arrays
        const double* u = U ? U + (u_stride ? mm->first[g] * u_stride : 0) : nullptr;
        double* y = Y ? Y + mm->first[g] * y_stride : nullptr;
        rcs[g] = acmeb200_run(mm->shards[g], u, u_stride, y, y_stride, N, 0, nullptr);
        if (rcs[g]) errs[g] = acmeb200_last_error();  // the message is thread local
    };
    std::vector<std::thread> threads;
    for (size_t g = 1; g < G; g++) threads.emplace_back(work, g);
    work(0);
    for (std::thread& t : threads) t.join();
    for (size_t g = 0; g < G; g++)
        if (rcs[g]) return fail(rcs[g], "shard %d: %s", (int)g, errs[g].c_str());
    return ACMEB200_OK;
}

// ------------------------------------------------------------------ k-d tree, host side (kdcache.cuh: the device's own code)
extern "C" int acmeb200_kdtree_build(int32_t np, int32_t n_columns, int32_t n_points, const double* ps, int32_t* cut_dim,
                                     double* cut_val, int32_t* ps_idx) {
    if (np < 0 || n_columns < 0 || n_points < 0 || n_points > n_columns) return fail(ACMEB200_EINVAL, "need 0 <= n_points <= n_columns");
    if (n_points == 0) return ACMEB200_OK;
    if ((np > 0 && !ps) || !ps_idx || (n_points > 1 && (!cut_dim || !cut_val))) return fail(ACMEB200_EINVAL, "null argument");
    if (np > 255) return fail(ACMEB200_EUNSUPPORTED, "np > 255");
    const int cap = n_columns;
    std::vector<double> img((size_t)kd_store_doubles(np, 0, cap), 0.0), scratch((size_t)kd_scratch_doubles(cap), 0.0);
    KdStore st = KdStore::at(img.data(), scratch.data(), np, 0, cap);
    if (np > 0) memcpy(st.cols, ps, sizeof(double) * (size_t)np * cap);  // nn = 0: the columns are the points
    kd_build(st, n_points, n_columns, n_columns);
    for (int k = 0; k < n_points - 1; k++) { cut_dim[k] = st.cutdim(k + 1); cut_val[k] = st.cutval(k + 1); }
    for (int k = 0; k < n_points; k++) ps_idx[k] = st.psidx(k + 1);
    return ACMEB200_OK;
}

extern "C" int acmeb200_kdtree_indnearest(int32_t np, int32_t n_columns, int32_t n_points, const int32_t* cut_dim, const double* cut_val,
                                          const int32_t* ps_idx, const double* ps, int32_t n_queries, const double* queries,
                                          double best_dist, int32_t best_pidx, int32_t* nearest_out) {
    if (np < 0 || n_columns < 0 || n_points < 0 || n_points > n_columns || n_queries < 0) return fail(ACMEB200_EINVAL, "bad sizes");
    if (np > KD_MAXNP) return fail(ACMEB200_EUNSUPPORTED, "the search supports np <= %d", KD_MAXNP);
    if (n_queries == 0) return ACMEB200_OK;
    if (!nearest_out || (np > 0 && (!queries || !ps)) || (n_points > 0 && !ps_idx) || (n_points > 1 && (!cut_dim || !cut_val)))
        return fail(ACMEB200_EINVAL, "null argument");
    for (int k = 0; k < n_points; k++)
        if (ps_idx[k] < 1 || ps_idx[k] > n_columns) return fail(ACMEB200_EINVAL, "ps_idx out of range");
    for (int k = 0; k < n_points - 1; k++)
        if (cut_dim[k] < 1 || cut_dim[k] > np) return fail(ACMEB200_EINVAL, "cut_dim out of range");
    const int cap = std::max(n_columns, 1);
    std::vector<double> img((size_t)kd_store_doubles(np, 0, cap), 0.0);
    KdStore st = KdStore::at(img.data(), nullptr, np, 0, cap);
    if (np > 0 && n_columns > 0) memcpy(st.cols, ps, sizeof(double) * (size_t)np * n_columns);
    for (int k = 0; k < n_points - 1; k++) st.set_cut(k + 1, cut_dim[k], cut_val[k]);
    for (int k = 0; k < n_points; k++) st.set_psidx(k + 1, ps_idx[k]);
    for (int q = 0; q < n_queries; q++) {
        const double* p = queries + (size_t)q * np;
        int ovf = 0;
        nearest_out[q] = kd_indnearest(st, n_points, [&](int i) { return p[i]; }, best_dist, best_pidx, &ovf);
    }
    return ACMEB200_OK;
}

// ------------------------------------------------------------------ device exp (diagnostic)
__global__ void __launch_bounds__(256) k_diag_exp(const double* x, double* out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
#ifdef ACME_HOST_EMU
    if (i < n) out[i] = exp(x[i]);  // the host emulation has no device exp: libm's
#else
    if (i < n) out[i] = acme_exp(x[i], ACME_EXPC);
#endif
}
extern "C" int acmeb200_diag_exp(const double* x_host, double* out_host, int64_t n) {
    if (!x_host || !out_host || n < 0) return fail(ACMEB200_EINVAL, "bad argument");
    if (n == 0) return ACMEB200_OK;
    double *dx = nullptr, *dy = nullptr;
    CUDA_TRY(cudaMalloc(&dx, sizeof(double) * (size_t)n));
    CUDA_TRY(cudaMalloc(&dy, sizeof(double) * (size_t)n));
    CUDA_TRY(cudaMemcpy(dx, x_host, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
    ACME_LAUNCH(k_diag_exp, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t) nullptr, dx, dy, n);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(out_host, dy, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
    cudaFree(dx); cudaFree(dy);
    return ACMEB200_OK;
}

// ------------------------------------------------------------------ FP64 pipe peak (diagnostic)
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

extern "C" int acmeb200_measure_fp64_peak(double* tflops_out) {
    if (!tflops_out) return fail(ACMEB200_EINVAL, "null argument");
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 8, tpb = 256, iters = 4096;
    double* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, sizeof(double) * (size_t)blocks * tpb));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CUDA_TRY(cudaEventRecord(e0));
        ACME_LAUNCH(k_dfma_peak, blocks, tpb, 0, (cudaStream_t) nullptr, d, iters, 0.999999, 1e-9);
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    const double flops = 2.0 * 64.0 * iters * (double)blocks * tpb;
    *tflops_out = flops / (best * 1e-3) / 1e12;
    return ACMEB200_OK;
}
