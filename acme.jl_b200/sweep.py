"""Model derivation at sweep scale.

A parameter sweep over values that are baked into the model matrices (resistors,
capacitors, fixed potentiometers: BASELINE config 3 and the alternative reading of
config 4) needs one host-side derivation per instance -- in the reference that is
``DiscreteModel(circ, t)`` (/root/reference/src/ACME.jl:150-262) called in a loop,
exact rationals and all.  The derivations are independent, so they are spread over
the host cores here; the results are stacked into the ``overrides`` / ``params`` /
``init_z`` arrays :class:`BatchRunner` takes (instance = last axis, column-major per
instance, i.e. exactly what the C ABI's per-instance strides describe).

Every instance must derive to the same *structure* as the first one (dimensions,
sub-problem decomposition, element kinds and q offsets): the device kernels are
chosen for one shape.  A sweep point where the exact-rational pipeline decides
differently (e.g. a resistor value of 0 changing the rank) raises.
"""
from __future__ import annotations

import multiprocessing as mp
import os
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

_LINEAR = ("a", "b", "c", "x0", "dy", "ey", "fy", "y0")
_SUB = ("dq", "eq", "fqprev", "pexp", "q0", "fq")

_builder: Optional[Callable] = None  # set in the parent before forking (closures need no pickling)


def _structure(m) -> tuple:
    return (m.nx, m.nu, m.ny, tuple((s.nn, s.nq, s.np_, tuple((e.kind, q) for e, q in s.elems)) for s in m.subs))


def _extract(m) -> dict:
    out = {k: np.asarray(getattr(m, k), dtype=np.float64) for k in _LINEAR}
    for i, s in enumerate(m.subs):
        for k in _SUB:
            out[f"{k}{i}"] = np.asarray(getattr(s, k), dtype=np.float64)
        out[f"init_z{i}"] = np.asarray(s.init_z, dtype=np.float64)
        flat: List[float] = []
        for e, _ in s.elems:
            flat.extend(e.params)
        out[f"params{i}"] = np.asarray(flat, dtype=np.float64)
    out["structure"] = _structure(m)
    return out


def _derive_chunk(chunk: Sequence) -> List[dict]:
    return [_extract(_builder(*p) if isinstance(p, tuple) else _builder(p)) for p in chunk]


def usable_cores() -> int:
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, -(-int(quota) // int(period))))
    except Exception:
        pass
    return n


def _cuda_initialised() -> bool:
    import sys
    torch = sys.modules.get("torch")
    try:
        return bool(torch is not None and torch.cuda.is_initialized())
    except Exception:
        return False


def derive_sweep(builder: Callable, points: Iterable, workers: Optional[int] = None, chunk: int = 16):
    """Derive ``builder(point)`` (a :class:`DiscreteModel`) for every sweep point.

    Returns ``(base_model, kwargs, B)`` with ``B = len(points)``: ``BatchRunner(base_model, B, **kwargs)``
    runs the sweep.  ``kwargs`` holds ``overrides`` (only the matrices that actually differ between
    instances; shape + (B,)), ``params`` (per sub, (nparams, B)) and ``init_z`` (per sub, (nn, B)).
    ``points`` are passed to ``builder`` as is (tuples are unpacked).
    """
    global _builder
    pts = list(points)
    if not pts:
        raise ValueError("empty sweep")
    if workers is None:
        # the pool forks: a process that already holds a CUDA context (driver threads, locks) should
        # not be forked behind the caller's back -- derive first, then create runners, or pass workers
        workers = 1 if _cuda_initialised() else usable_cores()
    chunks = [pts[i:i + chunk] for i in range(0, len(pts), chunk)]
    _builder = builder
    try:
        if workers > 1 and len(chunks) > 1:
            with mp.get_context("fork").Pool(min(workers, len(chunks))) as pool:
                parts = pool.map(_derive_chunk, chunks)
        else:
            parts = [_derive_chunk(c) for c in chunks]
    finally:
        _builder = None
    rows = [r for part in parts for r in part]
    base = builder(*pts[0]) if isinstance(pts[0], tuple) else builder(pts[0])
    st0 = _structure(base)
    for b, r in enumerate(rows):
        if r["structure"] != st0:
            raise ValueError(f"sweep point {b} ({pts[b]!r}) derives to a different model structure {r['structure']} "
                             f"than point 0 {st0}: one batch needs one shape")
    B = len(rows)
    nsub = len(base.subs)

    def stack(key):
        return np.stack([r[key] for r in rows], axis=-1)

    overrides: Dict[str, np.ndarray] = {}
    for key in list(_LINEAR) + [f"{k}{i}" for i in range(nsub) for k in _SUB]:
        s = stack(key)
        if s.size and not np.all(s == s[..., :1]):
            overrides[key] = np.asfortranarray(s)
    params = [np.asfortranarray(stack(f"params{i}")) for i in range(nsub)]
    init_z = [np.asfortranarray(stack(f"init_z{i}")) for i in range(nsub)]
    kwargs = {"overrides": overrides or None,
              "params": params if any(p.size and not np.all(p == p[:, :1]) for p in params) else None,
              "init_z": init_z if any(z.size and not np.all(z == z[:, :1]) for z in init_z) else None}
    assert all(v is None or len(v) for v in kwargs.values())
    return base, {k: v for k, v in kwargs.items() if v is not None}, B
