"""Host model of the device LU used by kernel_rows.cuh.

The warp-per-instance kernel never moves a matrix row: lane r keeps row r and only the POSITION the
reference's physically swapped matrix (setlhs!, /root/reference/src/solvers.jl:46-96) would hold it at
changes; the pivot search is a max over bit patterns with ties broken by position; the right-hand
side rides along as an extra column (= the forward substitution of solve!, solvers.jl:112-119) and
the back substitution (solvers.jl:120-129) broadcasts one component per step.  This numpy model
states that algorithm lane by lane exactly as the kernel does and pins it -- BIT FOR BIT -- to the
oracle's restatement of the reference's LinearSolver, on random matrices and on the adversarial cases
(exact ties, zero columns, singular matrices, rows needing every swap).  It needs no GPU; the GPU
tests then check the CUDA implementation against the cooperative kernel and the oracle."""
import numpy as np
import pytest

from oracle import oracle


def rows_lu(A, b):
    """returns (ok, LU rows by lane, pos, src per step, kp per step, b after elimination)"""
    n = A.shape[0]
    R = A.astype(np.float64).copy()        # R[lane] = that lane's row
    rhs = b.astype(np.float64).copy()
    pos = np.arange(n)
    src, kps = [], []
    for k in range(n):
        cand = pos >= k
        a = R[:, k]
        key = np.where(cand & ~np.isnan(a), np.abs(a), 0.0).view(np.uint64)   # |a| as a bit pattern, NaN never wins
        mx = key[cand].max()
        ties = np.flatnonzero(cand & (key == mx))
        s = ties[np.argmin(pos[ties])]        # first strict maximum in position order
        kp = pos[s]
        kps.append(kp)
        if mx == 0:                            # exactly zero pivot: the reference returns false here
            return False, R, pos, src, kps, rhs
        src.append(s)
        lk = np.flatnonzero(pos == k)[0]
        pos[lk], pos[s] = kp, k                # the row interchange, as a relabelling (no-op when kp == k)
        inv = 1.0 / R[s, k]
        prow, pb = R[s].copy(), rhs[s]
        below = cand.copy(); below[s] = False
        for i in np.flatnonzero(below):
            l = R[i, k] * inv
            R[i, k] = l
            for j in range(k + 1, n):
                R[i, j] = R[i, j] - l * prow[j]   # two roundings, like the reference's scalar loop
            rhs[i] = rhs[i] - l * pb
        R[s, k] = inv
    return True, R, pos, src, kps, rhs


def rows_backsolve(R, pos, src, c):
    """back substitution on rows-in-lanes; returns x indexed by COLUMN"""
    n = R.shape[0]
    b = c.copy()
    for j in range(n - 1, -1, -1):
        s = src[j]
        b[s] = R[s, j] * b[s]
        xj = b[s]
        for i in np.flatnonzero(pos < j):
            b[i] = b[i] - R[i, j] * xj
    x = np.empty(n)
    x[pos] = b
    return x


def adversarial(rng, n):
    A = rng.standard_normal((n, n))
    kind = rng.integers(0, 6)
    if kind == 0:      # exact ties in every column: +-1 entries
        A = rng.choice([-1.0, 1.0, 0.5, 2.0], size=(n, n))
    elif kind == 1:    # reversed diagonal dominance: every step swaps
        A = np.eye(n)[::-1] * 10 + 0.1 * rng.standard_normal((n, n))
    elif kind == 2:    # a zero first column except one entry, wide dynamic range
        A[:, 0] = 0.0; A[rng.integers(0, n), 0] = 1e-300
        A[:, -1] *= 1e200
    elif kind == 3:    # two equal rows: exactly singular, must fail like the reference
        A[n - 1] = A[0]
    elif kind == 4:    # a zero column
        A[:, rng.integers(0, n)] = 0.0
    return A


@pytest.mark.parametrize("n", [1, 2, 3, 7, 13])
def test_rows_lu_equals_reference_linearsolver_bitwise(n):
    rng = np.random.default_rng(100 + n)
    nok = 0
    for trial in range(120):
        A = adversarial(rng, n) if trial % 2 else rng.standard_normal((n, n)) * 10.0 ** rng.integers(-8, 8)
        b = rng.standard_normal(n)
        ok_ref, x_ref = oracle.linsolve(A, b)
        ok, R, pos, src, kps, c = rows_lu(A, b)
        assert ok == ok_ref, (n, trial)
        if not ok:
            continue
        nok += 1
        assert sorted(pos) == list(range(n))
        x = rows_backsolve(R, pos, src, c)
        assert np.array_equal(x.view(np.uint64), x_ref.view(np.uint64)), (n, trial, np.abs(x - x_ref).max())
    assert nok > 40


def test_rows_lu_positions_are_the_reference_permutation():
    """pos (lane -> position) reproduces ipiv: applying the reference's sequential swaps to the identity
    gives the same arrangement of original rows"""
    rng = np.random.default_rng(7)
    for _ in range(50):
        n = 13
        A = rng.standard_normal((n, n))
        ok, R, pos, src, kps, _ = rows_lu(A, np.zeros(n))
        assert ok
        perm = list(range(n))
        for k, kp in enumerate(kps):
            perm[k], perm[kp] = perm[kp], perm[k]
        assert all(pos[perm[i]] == i for i in range(n))
        assert all(pos[s] == k for k, s in enumerate(src))
