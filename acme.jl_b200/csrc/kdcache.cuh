// The CachingSolver's solution store on the device: the reference's k-d tree, moved on-device.
//
//   KDTree(p, Np)            /root/reference/src/kdtree.jl:11-73    -> kd_build
//   Alts / heap / indnearest /root/reference/src/kdtree.jl:75-234   -> kd_indnearest
//   solve(::CachingSolver)   /root/reference/src/solvers.jl:347-396 -> kd_lookup (start-point choice),
//                                                                      kd_after_solve (store, rebuild schedule)
//
// One store per (sub-problem, instance), instance-contiguous in global memory:
//   hdr[8] ints | nodes[cap] (16 bytes: cut value, cut dimension, leaf column) | columns[cap][np + nn] (p then z)
// plus a scratch block (keys + permutation, twice `cap` each) used only while a tree is rebuilt.
//
// What is reproduced on purpose, because it decides WHICH stored solution becomes the start point and
// therefore the iteration counts (tests compare them with the oracle's, which restates the same code):
//   * the tree is rebuilt on the reference's schedule (new_count / new_count_limit, solvers.jl:387-394), the
//     newest new_count entries are scanned linearly before the tree search (solvers.jl:354-363);
//   * KDTree(ps, num_ps) sorts `p[dim, :]` over ALL capacity columns (kdtree.jl:37), so zero-filled spare
//     columns of the doubling arrays (solvers.jl:376-382) can enter the tree and real points can drop out of
//     it.  The device arrays do not double: `cap_ref` tracks the capacity the reference's arrays would have,
//     spare columns up to the physical capacity are zero-filled memory, the ones beyond it are virtual zeros;
//   * the search is the reference's best-first search with its heap operations, so ties between equally
//     distant points resolve the same way; distances and bounds are accumulated unfused, in the same order.
// Bounded where the reference is unbounded: the physical capacity (a store that is full forgets its older half,
// kd_compact: flag KD_F_FULL) and the heap of alternatives (KD_HEAP entries; an overflow is counted, never seen
// in the tests: the deepest heap observed on the BASELINE circuits holds 32 entries).
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__) || defined(ACME_HOST_EMU)
#define KD_HD __host__ __device__ inline
#else
#define KD_HD inline
#endif

namespace acme {

constexpr int KD_HEAP = 64;    // alternatives kept by one search
constexpr int KD_MAXNP = 32;   // point dimension the search's delta vector holds
enum { KD_H_NUM = 0, KD_H_NEW = 1, KD_H_LIMIT = 2, KD_H_CAPREF = 3, KD_H_TREEN = 4, KD_H_FLAGS = 5, KD_HDR_INTS = 8 };
enum { KD_F_FROZEN = 1, KD_F_FULL = 2, KD_F_HEAP_OVERFLOW = 4 };

// unfused arithmetic on both sides (the reference's Julia code is not contracted)
KD_HD double kd_mul(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;
    return r;
#endif
}
KD_HD double kd_add(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;
    return r;
#endif
}
KD_HD double kd_sub(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dsub_rn(a, b);
#else
    volatile double r = a - b;
    return r;
#endif
}

// doubles of one instance's store / scratch
KD_HD int64_t kd_store_doubles(int np, int nn, int cap) { return KD_HDR_INTS / 2 + 2 * (int64_t)cap + (int64_t)cap * (np + nn); }
KD_HD int64_t kd_scratch_doubles(int cap) { return 2 * (int64_t)cap + cap; }  // 2*cap keys + 2*cap ints

// One tree node / leaf slot, 16 bytes: slot k-1 holds the cut of NODE k (nodes 1..Np-1) and the column of LEAF k
// (leaves 1..Np) -- two unrelated things that share an index range, so a small tree is one or two cache lines.
struct alignas(16) KdNode {
    double cut_val;
    int ti;  // bits 0..7 cut_dim (1-based), bits 8.. ps_idx (1-based column)
    int pad;
};

struct KdStore {
    int* hdr;
    KdNode* nodes;  // [cap]
    double* cols;   // column c (1-based) at [(c-1)*(np+nn)]: p (np values) then z (nn values)
    double* skey;   // scratch: [0,cap) keys of the range being sorted, [cap,2cap) merge buffer
    int* sidx;      // scratch: [0,cap) permutation p_idx (1-based columns), [cap,2cap) merge buffer
    int np, nn, cap;

    KD_HD static KdStore at(double* base, double* scratch, int np, int nn, int cap) {
        KdStore c;
        c.np = np; c.nn = nn; c.cap = cap;
        c.hdr = reinterpret_cast<int*>(base);
        c.nodes = reinterpret_cast<KdNode*>(base + KD_HDR_INTS / 2);
        c.cols = base + KD_HDR_INTS / 2 + 2 * (int64_t)cap;
        c.skey = scratch;
        c.sidx = scratch ? reinterpret_cast<int*>(scratch + 2 * (int64_t)cap) : nullptr;
        return c;
    }
    KD_HD double* col(int c1) const { return cols + (int64_t)(c1 - 1) * (np + nn); }
    // columns beyond the physical capacity are the virtual zero columns of the reference's doubled arrays
    KD_HD double P(int d, int c1) const { return c1 <= cap ? col(c1)[d] : 0.0; }
    KD_HD double Z(int i, int c1) const { return c1 <= cap ? col(c1)[np + i] : 0.0; }
    KD_HD int cutdim(int node) const { return nodes[node - 1].ti & 0xff; }
    KD_HD double cutval(int node) const { return nodes[node - 1].cut_val; }
    KD_HD int psidx(int leaf) const { return (int)((unsigned)nodes[leaf - 1].ti >> 8); }
    KD_HD void set_cut(int node, int dim, double v) {
        nodes[node - 1].ti = (nodes[node - 1].ti & ~0xff) | dim;
        nodes[node - 1].cut_val = v;
    }
    KD_HD void set_psidx(int leaf, int c1) { nodes[leaf - 1].ti = (nodes[leaf - 1].ti & 0xff) | (c1 << 8); }
};

// ---------------------------------------------------------------------------------------------- construction
// calc_cut_idx (kdtree.jl:12-20), 1-based
KD_HD int kd_calc_cut_idx(int min_idx, int max_idx) {
    const int N = max_idx - min_idx + 1;
    int e = 0;
    while ((1 << (e + 1)) <= N - 1) e++;
    const int N2 = 1 << e;
    if (3 * (N2 / 2) <= N) return min_idx + N2 - 1;
    return min_idx + N - N2 / 2 - 1;
}

// [lo, hi] of node n (heap numbering) of a tree over Np points: the min_idx/max_idx arrays of kdtree.jl:28-52,
// recomputed from the root instead of stored
KD_HD void kd_node_range(int n, int Np, int& lo, int& hi) {
    lo = 1; hi = Np;
    int depth = 0;
    while ((n >> (depth + 1)) != 0) depth++;
    for (int b = depth - 1; b >= 0; b--) {
        const int cut = kd_calc_cut_idx(lo, hi);
        if (((n >> b) & 1) == 0) hi = cut; else lo = cut + 1;
    }
}

// argmax(vec(var(p[:, cols], dims=2))) over the columns sidx[lo-1 .. hi-1]: first maximum, NaN counts as
// maximal (Julia's argmax); returns the 1-based dimension
KD_HD int kd_argmax_var(const KdStore& c, int lo, int hi) {
    const int n = hi - lo + 1;
    int best = 1;
    double bestv = 0.0;
    for (int d = 0; d < c.np; d++) {
        double mean = 0.0;
        for (int k = lo; k <= hi; k++) mean = kd_add(mean, c.P(d, c.sidx[k - 1]));
        mean /= n;
        double ss = 0.0;
        for (int k = lo; k <= hi; k++) {
            const double dv = kd_sub(c.P(d, c.sidx[k - 1]), mean);
            ss = kd_add(ss, kd_mul(dv, dv));
        }
        const double v = ss / (n - 1);
        if (d == 0) { best = 1; bestv = v; continue; }
        if (bestv != bestv) continue;
        if (v != v || v > bestv) { best = d + 1; bestv = v; }
    }
    return best;
}

// stable sort of the permutation entries [lo, hi] (1-based positions) by P(dim, column): what
// `p_idx[lo:hi] = p_idx[sortperm(p[dim, p_idx[lo:hi]]) .+ lo .- 1]` does (kdtree.jl:53-55).  Bottom-up merge sort
// with the keys carried along; the stable result is unique, so the algorithm is free.
KD_HD void kd_sort_range(KdStore& c, int lo, int hi, int dim /*1-based*/) {
    const int a = lo - 1, n = hi - lo + 1, cap = c.cap;
    for (int k = 0; k < n; k++) c.skey[a + k] = c.P(dim - 1, c.sidx[a + k]);
    int src = 0;  // 0: data in [0,cap), 1: in [cap,2cap)
    for (int w = 1; w < n; w *= 2) {
        const int so = src ? cap : 0, dof = src ? 0 : cap;
        for (int s = 0; s < n; s += 2 * w) {
            const int m = s + w < n ? s + w : n, e = s + 2 * w < n ? s + 2 * w : n;
            int i = s, j = m, o = s;
            while (i < m && j < e) {
                const double ki = c.skey[so + a + i], kj = c.skey[so + a + j];
                if (kj < ki) { c.skey[dof + a + o] = kj; c.sidx[dof + a + o] = c.sidx[so + a + j]; j++; }
                else { c.skey[dof + a + o] = ki; c.sidx[dof + a + o] = c.sidx[so + a + i]; i++; }
                o++;
            }
            for (; i < m; i++, o++) { c.skey[dof + a + o] = c.skey[so + a + i]; c.sidx[dof + a + o] = c.sidx[so + a + i]; }
            for (; j < e; j++, o++) { c.skey[dof + a + o] = c.skey[so + a + j]; c.sidx[dof + a + o] = c.sidx[so + a + j]; }
        }
        src ^= 1;
    }
    if (src)
        for (int k = 0; k < n; k++) { c.skey[a + k] = c.skey[cap + a + k]; c.sidx[a + k] = c.sidx[cap + a + k]; }
}

// KDTree(p, Np) (kdtree.jl:11-73) over the store's columns: the first `ncols` columns are memory (cap_ref of the
// dynamic store clipped to the physical capacity; a host-built frozen tree passes its n_columns), the columns
// ncols+1 .. cap_ref are the virtual zero columns.  Writes cut_dim / cut_val / ps_idx; needs the scratch block.
KD_HD void kd_build(KdStore& c, int Np, int ncols, int cap_ref) {
    c.hdr[KD_H_TREEN] = Np;
    if (Np <= 0) return;
    if (Np == 1) { c.set_psidx(1, 1); return; }
    for (int k = 0; k < Np; k++) c.sidx[k] = k + 1;
    int dim = kd_argmax_var(c, 1, Np);
    // p_idx = sortperm(vec(p[dim, :])) over every capacity column (kdtree.jl:37): only its first Np entries are used
    for (int k = 0; k < ncols; k++) c.sidx[k] = k + 1;
    kd_sort_range(c, 1, ncols, dim);
    const int nvirt = cap_ref > ncols ? cap_ref - ncols : 0;
    if (nvirt > 0) {
        // the virtual zero columns sort behind every real column whose key is <= 0 (stable: they have the larger indices)
        int a = 0;
        while (a < ncols && !(0.0 < c.skey[a])) a++;
        for (int k = Np - 1; k >= a; k--) c.sidx[k] = (k - a < nvirt) ? ncols + 1 + (k - a) : c.sidx[k - nvirt];
    }
    {
        const int cut = kd_calc_cut_idx(1, Np);
        c.set_cut(1, dim, (c.P(dim - 1, c.sidx[cut - 1]) + c.P(dim - 1, c.sidx[cut])) / 2);
    }
    for (int n = 2; n <= Np - 1; n++) {
        int lo, hi;
        kd_node_range(n, Np, lo, hi);
        dim = kd_argmax_var(c, lo, hi);
        kd_sort_range(c, lo, hi, dim);
        const int cut = kd_calc_cut_idx(lo, hi);
        c.set_cut(n, dim, (c.P(dim - 1, c.sidx[cut - 1]) + c.P(dim - 1, c.sidx[cut])) / 2);
    }
    for (int n = 1; n <= Np; n++) {
        int lo, hi;
        kd_node_range((n + Np - 1) / 2, Np, lo, hi);
        c.set_psidx(n, ((n + Np) % 2 == 1) ? c.sidx[lo - 1] : c.sidx[hi - 1]);
    }
}

// ---------------------------------------------------------------------------------------------- search
// The heap of alternatives (Alts, kdtree.jl:75-187).  An entry is (node, delta_norm); the reference also stores
// the entry's delta vector, which is a function of the node alone -- for every dimension the offset to the cut of
// the deepest ancestor at which the path root -> node left the query's side -- and is recomputed when the entry
// is dequeued.
struct KdHeap {
    int idx[KD_HEAP];
    double norm[KD_HEAP];
    int nvalid;
    double best_dist;
    int best_pidx;
    int overflow;
    // 1-based entry access like the reference
    KD_HD void swap(int i, int j) {
        const int ti_ = idx[i - 1]; idx[i - 1] = idx[j - 1]; idx[j - 1] = ti_;
        const double tn = norm[i - 1]; norm[i - 1] = norm[j - 1]; norm[j - 1] = tn;
    }
    KD_HD void siftup(int i) {  // kdtree.jl:102-113
        int parent = i / 2;
        while (i > 1 && norm[i - 1] < norm[parent - 1]) { swap(i, parent); i = parent; parent = i / 2; }
    }
    KD_HD void siftdown(int i) {  // kdtree.jl:115-135
        const int N = nvalid;
        for (;;) {
            int mn = i;
            if (2 * i <= N && norm[2 * i - 1] < norm[mn - 1]) mn = 2 * i;
            if (2 * i + 1 <= N && norm[2 * i] < norm[mn - 1]) mn = 2 * i + 1;
            if (mn == i) break;
            swap(i, mn);
            i = mn;
        }
    }
    KD_HD void deleteat(int i) {  // kdtree.jl:140-150
        swap(i, nvalid);
        nvalid -= 1;
        if (i <= nvalid) {
            if (i == 1 || norm[i - 1] > norm[i / 2 - 1]) siftdown(i); else siftup(i);
        }
    }
    KD_HD void enqueue(int new_idx, double new_norm) {  // kdtree.jl:158-175
        if (nvalid == KD_HEAP) { overflow = 1; return; }
        idx[nvalid] = new_idx;
        norm[nvalid] = new_norm;
        if (new_norm < best_dist) { nvalid += 1; siftup(nvalid); }
    }
    KD_HD void update_best(double dist, int p_idx) {  // kdtree.jl:177-187
        if (dist < best_dist) {
            best_dist = dist;
            best_pidx = p_idx;
            for (int i = nvalid; i >= 1; i--)
                if (norm[i - 1] >= best_dist) deleteat(i);
        }
    }
};

// indnearest(tree, p, alts) (kdtree.jl:192-234) with alts initialised by init!(alts, best_dist, best_pidx)
// (kdtree.jl:93-100): the 1-based column of the nearest tree point strictly nearer than best_dist, else best_pidx
template <class PF>
KD_HD int kd_indnearest(const KdStore& c, int tree_n, PF p, double best_dist, int best_pidx, int* overflow) {
    KdHeap h;
    h.nvalid = 1; h.idx[0] = 1; h.norm[0] = 0.0;
    h.best_dist = best_dist; h.best_pidx = best_pidx; h.overflow = 0;
    const int ncut = tree_n > 0 ? tree_n - 1 : 0;
    const int np = c.np < KD_MAXNP ? c.np : KD_MAXNP;
    double delta[KD_MAXNP];
    while (h.nvalid > 0) {
        int idx = h.idx[0];
        const double delta_norm = h.norm[0];
        h.deleteat(1);
        if (tree_n == 0) break;
        // the dequeued entry's delta vector
        for (int i = 0; i < np; i++) delta[i] = 0.0;
        {
            int depth = 0;
            while ((idx >> (depth + 1)) != 0) depth++;
            for (int b = depth - 1; b >= 0; b--) {
                const int a = idx >> (b + 1), child = idx >> b;
                const int dim = c.cutdim(a) - 1;
                const double cv = c.cutval(a), pd = p(dim);
                const int near = pd <= cv ? 2 * a : 2 * a + 1;
                if (child != near) delta[dim] = pd - cv;
            }
        }
        while (idx <= ncut) {
            const int dim = c.cutdim(idx) - 1;
            const double cv = c.cutval(idx), pd = p(dim);
            const double dcut = pd - cv;
            const double new_norm = kd_add(kd_sub(delta_norm, kd_mul(delta[dim], delta[dim])), kd_mul(dcut, dcut));
            if (new_norm < h.best_dist) h.enqueue(pd <= cv ? 2 * idx + 1 : 2 * idx, new_norm);
            idx = pd <= cv ? 2 * idx : 2 * idx + 1;
        }
        idx -= ncut;
        const int p_idx = c.psidx(idx);
        double dist = 0.0;
        for (int i = 0; i < c.np; i++) {
            const double d = p(i) - c.P(i, p_idx);
            dist = kd_add(dist, kd_mul(d, d));
        }
        h.update_best(dist, p_idx);
    }
    if (h.overflow && overflow) *overflow = 1;
    return h.best_pidx;
}

// ---------------------------------------------------------------------------------------------- CachingSolver
// The start-point choice of solve(::CachingSolver, p) (solvers.jl:348-366): nearest of {current origin (its squared
// distance is best_diff), the new_count newest stored solutions (linear scan), the tree}.  Returns the 1-based
// column to re-origin at, or 0 to keep the origin.
template <class PF>
KD_HD int kd_lookup(const KdStore& c, PF p, double best_diff, int* overflow) {
    const int num_ps = c.hdr[KD_H_NUM], new_count = c.hdr[KD_H_NEW];
    int idx = 0;
    for (int i = num_ps - new_count + 1; i <= num_ps; i++) {
        double diff = 0.0;
        for (int j = 0; j < c.np; j++) {
            const double d = c.P(j, i) - p(j);
            diff = kd_add(diff, kd_mul(d, d));
        }
        if (diff < best_diff) { best_diff = diff; idx = i; }
    }
    return kd_indnearest(c, c.hdr[KD_H_TREEN], p, best_diff, idx, overflow);
}

// A full store forgets its older half (the initial solution in column 1 stays): the newest cap/2 columns move down
// to columns 2.., the rest is zeroed (spare columns are the zero-filled spare capacity of the reference's arrays), and
// the tree is marked due for a rebuild at once.  This is where the device departs from the reference, whose arrays
// double for ever (solvers.jl:376-382): it keeps learning with bounded memory.  (A store that simply stops accepting
// solutions gets worse with time: measured on config 4 at capacity 1024, 3.2 -> 8.6 iterations per solve.)
KD_HD void kd_compact(KdStore& c) {
    int* const h = c.hdr;
    const int cap = c.cap, keep = cap / 2, w = c.np + c.nn;
    const int first = cap - keep + 1;  // oldest column that stays
    for (int k = 0; k < keep; k++) {
        const double* const src = c.col(first + k);
        double* const dst = c.col(2 + k);
        for (int j = 0; j < w; j++) dst[j] = src[j];
    }
    for (int col = keep + 2; col <= cap; col++) {
        double* const dst = c.col(col);
        for (int j = 0; j < w; j++) dst[j] = 0.0;
    }
    h[KD_H_NUM] = keep + 1;
    h[KD_H_CAPREF] = keep + 1;  // from here the reference's doubling rule again (solvers.jl:376-382)
    h[KD_H_NEW] = 1;     // with LIMIT = 0: the tree (whose leaves name the old columns) is rebuilt before the next search
    h[KD_H_LIMIT] = 0;
    h[KD_H_FLAGS] |= KD_F_FULL;
}

// What follows the base solve (solvers.jl:374-389): store (p, z) if the solve needed more than 5 iterations and
// converged, count down to the next rebuild.  Returns true when the tree is due for a rebuild (solvers.jl:390):
// the caller rebuilds (kd_build, or kd_build_warp with its whole warp) and calls kd_rebuilt.
template <class PF, class ZF>
KD_HD bool kd_store_step(KdStore& c, bool store, PF p, ZF z, int* stored_col = nullptr) {
    int* const h = c.hdr;
    if (stored_col) *stored_col = 0;
    if (h[KD_H_FLAGS] & KD_F_FROZEN) return false;
    bool compacted = false;
    if (store) {
        if (h[KD_H_NUM] >= c.cap && c.cap >= 8) { kd_compact(c); compacted = true; }
        if (h[KD_H_NUM] < c.cap) {
            const int n = ++h[KD_H_NUM];
            if (n > h[KD_H_CAPREF]) h[KD_H_CAPREF] = 2 * n;  // the reference's arrays double here (solvers.jl:376-382)
            double* const dst = c.col(n);
            for (int j = 0; j < c.np; j++) dst[j] = p(j);
            for (int j = 0; j < c.nn; j++) dst[c.np + j] = z(j);
            if (!compacted) h[KD_H_NEW] += 1;
            if (stored_col) *stored_col = compacted ? 0 : n;
        } else {
            h[KD_H_FLAGS] |= KD_F_FULL;  // capacity below 8: the store simply stops learning
        }
    }
    if (compacted) return true;  // rebuild at once
    if (h[KD_H_NEW] > 0) h[KD_H_LIMIT] -= 1;
    return h[KD_H_NEW] > h[KD_H_LIMIT];
}
KD_HD void kd_rebuilt(KdStore& c) {  // solvers.jl:391-393
    c.hdr[KD_H_NEW] = 0;
    c.hdr[KD_H_LIMIT] = 2 * c.hdr[KD_H_CAPREF];
}
// the whole of solvers.jl:374-394 by one thread.  Returns true when the tree was rebuilt.
template <class PF, class ZF>
KD_HD bool kd_after_solve(KdStore& c, bool store, PF p, ZF z, int* stored_col = nullptr) {
    if (!kd_store_step(c, store, p, z, stored_col)) return false;
    // spare columns, physical or not, are zeros: the virtual ones of kd_build.  A store that has been compacted has left
    // the reference's path anyway: its trees hold stored solutions only (no spare (p = 0, z = 0) columns as start points)
    const int cap_ref = (c.hdr[KD_H_FLAGS] & KD_F_FULL) ? c.hdr[KD_H_NUM] : c.hdr[KD_H_CAPREF];
    kd_build(c, c.hdr[KD_H_NUM], c.hdr[KD_H_NUM], cap_ref);
    kd_rebuilt(c);
    return true;
}

// CachingSolver(basesolver, initial_p = 0, initial_z, nn) (solvers.jl:327-333) on zero-filled memory
template <class ZF>
KD_HD void kd_init(KdStore& c, ZF init_z) {
    c.hdr[KD_H_NUM] = 1; c.hdr[KD_H_NEW] = 0; c.hdr[KD_H_LIMIT] = 2; c.hdr[KD_H_CAPREF] = 1; c.hdr[KD_H_TREEN] = 1;
    c.hdr[KD_H_FLAGS] = 0; c.hdr[6] = 0; c.hdr[7] = 0;
    for (int j = 0; j < c.np; j++) c.cols[j] = 0.0;
    for (int j = 0; j < c.nn; j++) c.cols[c.np + j] = init_z(j);
    c.nodes[0].ti = 1 << 8;  // KDTree(hcat(initial_p)): one leaf, column 1
}

}  // namespace acme
