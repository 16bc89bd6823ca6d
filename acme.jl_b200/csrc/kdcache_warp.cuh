// Warp-cooperative versions of the expensive parts of kdcache.cuh, bit-identical to the serial code.
//
// KDTree(ps, num_ps) (/root/reference/src/kdtree.jl:11-73) runs once per 2*capacity solves of a learning
// CachingSolver (solvers.jl:387-394); one thread doing it alone costs ~10^8 cycles at a thousand stored points
// (measured: config 4 six times slower).  Here the 32 lanes of a converged warp build one instance's tree:
//   * variances: lane d owns dimension d and adds the segment's points in permutation order -- the serial
//     summation order, so the arg-max (and with it the tree) cannot differ by rounding;
//   * sorts: a rank sort -- element i goes to position #{j : j sorts before i} under the stable order
//     (key, current position), which is the unique result of the reference's stable sortperm -- with the keys
//     passed around by shuffles;
//   * leaves in parallel.
// Every lane of the warp must call these functions together (full-mask collectives inside).
#pragma once
#include "kdcache.cuh"

namespace acme {

constexpr unsigned KDW_FULL = 0xffffffffu;

__device__ __forceinline__ double kdw_shfl(double v, int src) {
    const int hi = __shfl_sync(KDW_FULL, __double2hiint(v), src), lo = __shfl_sync(KDW_FULL, __double2loint(v), src);
    return __hiloint2double(hi, lo);
}

// kd_argmax_var over the permutation entries [lo, hi]
__device__ inline int kdw_argmax_var(const KdStore& c, int lo, int hi, int lane) {
    const int n = hi - lo + 1;
    const int d = lane < c.np ? lane : 0;
    double mean = 0.0;
    for (int b = 0; b < n; b += 32) {
        const int k = lo - 1 + b + lane;
        const int colv = k < hi ? c.sidx[k] : 1;
        const int m = n - b < 32 ? n - b : 32;
        for (int j = 0; j < m; j++) mean = kd_add(mean, c.P(d, __shfl_sync(KDW_FULL, colv, j)));
    }
    mean /= n;
    double ss = 0.0;
    for (int b = 0; b < n; b += 32) {
        const int k = lo - 1 + b + lane;
        const int colv = k < hi ? c.sidx[k] : 1;
        const int m = n - b < 32 ? n - b : 32;
        for (int j = 0; j < m; j++) {
            const double dv = kd_sub(c.P(d, __shfl_sync(KDW_FULL, colv, j)), mean);
            ss = kd_add(ss, kd_mul(dv, dv));
        }
    }
    const double v = ss / (n - 1);
    // first maximum, NaN counts as maximal (the serial loop of kd_argmax_var, one dimension per step)
    int best = 1;
    double bestv = kdw_shfl(v, 0);
    for (int dd = 1; dd < c.np; dd++) {
        const double vd = kdw_shfl(v, dd);
        if (bestv != bestv) continue;
        if (vd != vd || vd > bestv) { best = dd + 1; bestv = vd; }
    }
    return best;
}

// kd_sort_range: stable sort of the permutation entries [lo, hi] by P(dim, column); leaves the sorted keys in skey
__device__ inline void kdw_sort_range(const KdStore& c, int lo, int hi, int dim /*1-based*/, int lane) {
    const int a = lo - 1, n = hi - lo + 1, cap = c.cap;
    if (n <= 1) {
        if (n == 1 && lane == 0) c.skey[a] = c.P(dim - 1, c.sidx[a]);
        __syncwarp();
        return;
    }
    if (n <= 32) {  // one element per lane, everything in registers
        const bool on = lane < n;
        const int colv = on ? c.sidx[a + lane] : 1;
        const double key = on ? c.P(dim - 1, colv) : 0.0;
        int rank = 0;
        for (int j = 0; j < n; j++) {
            const double kj = kdw_shfl(key, j);
            rank += (kj < key || (!(key < kj) && j < lane)) ? 1 : 0;
        }
        __syncwarp();
        if (on) { c.sidx[a + rank] = colv; c.skey[a + rank] = key; }
        __syncwarp();
        return;
    }
    // keys first (one gather), then ranks chunk by chunk; results go to the merge buffer and are copied back
    for (int k = lane; k < n; k += 32) c.skey[a + k] = c.P(dim - 1, c.sidx[a + k]);
    __syncwarp();
    for (int bi = 0; bi < n; bi += 32) {
        const int i = bi + lane;
        const bool on = i < n;
        const double key = on ? c.skey[a + i] : 0.0;
        const int colv = on ? c.sidx[a + i] : 1;
        int rank = 0;
        for (int bj = 0; bj < n; bj += 32) {
            const double kl = bj + lane < n ? c.skey[a + bj + lane] : 0.0;
            const int m = n - bj < 32 ? n - bj : 32;
            for (int jj = 0; jj < m; jj++) {
                const double kj = kdw_shfl(kl, jj);
                rank += (kj < key || (!(key < kj) && bj + jj < i)) ? 1 : 0;
            }
        }
        if (on) { c.sidx[cap + a + rank] = colv; c.skey[cap + a + rank] = key; }
    }
    __syncwarp();
    for (int k = lane; k < n; k += 32) { c.sidx[a + k] = c.sidx[cap + a + k]; c.skey[a + k] = c.skey[cap + a + k]; }
    __syncwarp();
}

// kd_build by a whole warp; same arguments, same result
__device__ inline void kd_build_warp(KdStore& c, int Np, int ncols, int cap_ref, int lane) {
    if (lane == 0) c.hdr[KD_H_TREEN] = Np;
    if (Np <= 0) return;
    if (Np == 1) {
        if (lane == 0) c.set_psidx(1, 1);
        __syncwarp();
        return;
    }
    for (int k = lane; k < ncols; k += 32) c.sidx[k] = k + 1;
    __syncwarp();
    int dim = kdw_argmax_var(c, 1, Np, lane);
    kdw_sort_range(c, 1, ncols, dim, lane);
    const int nvirt = cap_ref > ncols ? cap_ref - ncols : 0;
    if (nvirt > 0) {
        // the virtual zero columns sort behind every real column whose key is <= 0: a = number of such keys
        int a = 0;
        for (int b = 0; b < ncols; b += 32) {
            const bool le = b + lane < ncols && !(0.0 < c.skey[b + lane]);
            a += __popc(__ballot_sync(KDW_FULL, le));
        }
        // shift the tail by nvirt, from the back, a round of 32 at a time: read, synchronise, write
        for (int top = Np - 1; top >= a; top -= 32) {
            const int k = top - lane;
            int v = 0;
            if (k >= a) v = (k - a < nvirt) ? ncols + 1 + (k - a) : c.sidx[k - nvirt];
            __syncwarp();
            if (k >= a) c.sidx[k] = v;
            __syncwarp();
        }
    }
    if (lane == 0) {
        const int cut = kd_calc_cut_idx(1, Np);
        c.set_cut(1, dim, (c.P(dim - 1, c.sidx[cut - 1]) + c.P(dim - 1, c.sidx[cut])) / 2);
    }
    __syncwarp();
    for (int n = 2; n <= Np - 1; n++) {
        int lo, hi;
        kd_node_range(n, Np, lo, hi);
        dim = kdw_argmax_var(c, lo, hi, lane);
        kdw_sort_range(c, lo, hi, dim, lane);
        if (lane == 0) {
            const int cut = kd_calc_cut_idx(lo, hi);
            c.set_cut(n, dim, (c.P(dim - 1, c.sidx[cut - 1]) + c.P(dim - 1, c.sidx[cut])) / 2);
        }
        __syncwarp();
    }
    for (int n = 1 + lane; n <= Np; n += 32) {
        int lo, hi;
        kd_node_range((n + Np - 1) / 2, Np, lo, hi);
        c.set_psidx(n, ((n + Np) % 2 == 1) ? c.sidx[lo - 1] : c.sidx[hi - 1]);
    }
    __syncwarp();
}

}  // namespace acme
