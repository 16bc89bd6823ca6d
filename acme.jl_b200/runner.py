"""``ModelRunner`` / ``run!`` over the CUDA library.

Mirrors /root/reference/src/ACME.jl:567-664: ``run_(model, u)`` is
``run!(model, u)``, ``ModelRunner`` pre-allocates (here: uploads the model and
keeps the per-instance state on the device), ``run_(runner, y, u)`` is the
in-place method.  :class:`BatchRunner` is the batched extension: B independent
instances of one circuit (swept element parameters / matrices, or many inputs),
``U`` of shape ``(nu, N, B)`` and ``Y`` of shape ``(ny, N, B)`` -- instance
slowest, so each instance's block is exactly the reference's ``nu x N`` matrix.

Nothing here computes samples on the CPU: the only implementation of the
per-sample path is the CUDA library; without it (or without a device) the
constructors raise.
"""
from __future__ import annotations

import ctypes as C
import warnings
from typing import Optional

import numpy as np

from . import _abi
from ._abi import make_desc, Stats
from ._lib import check, lib

WARN_NOT_CONVERGED = "Failed to converge while solving non-linear equation."       # ACME.jl:690
ERR_NONFINITE = "Failed to converge while solving non-linear equation, got non-finite result."  # ACME.jl:692


class DimensionMismatch(ValueError):
    pass


def _is_torch_cuda(t):
    return hasattr(t, "is_cuda") and t.is_cuda


class BatchRunner:
    """B instances of ``model`` resident on the current CUDA device.

    ``first``/``count`` select the slice of instances this runner owns (for
    sharding one descriptor over several GPUs); per-instance arrays are always
    given for the full batch.
    """

    def __init__(self, model, batch: int = 1, *, first: int = 0, count: Optional[int] = None,
                 kernel: str = "auto", **desc_kw):
        self.model = model
        self.batch_total = batch
        self.first = first
        self.batch = batch - first if count is None else count
        if first < 0 or self.batch <= 0 or first + self.batch > batch:
            raise ValueError(f"instance range [{first}, {first}+{self.batch}) outside the batch of {batch}")
        self._params = desc_kw.get("params")
        self._overrides = desc_kw.get("overrides")
        self._has_overrides = bool(desc_kw.get("overrides"))
        self._holder = make_desc(model, batch, **desc_kw)
        h = C.c_void_p()
        check(lib().acmeb200_model_create(C.byref(self._holder.desc), first, self.batch, C.byref(h)))
        self._h = h
        if kernel == "generic":
            check(lib().acmeb200_set_kernel(self._h, 1))
        elif kernel == "coop":
            check(lib().acmeb200_set_kernel(self._h, 2))
        elif kernel == "rows":
            check(lib().acmeb200_set_kernel(self._h, 3))
        elif kernel != "auto":
            raise ValueError("kernel must be 'auto', 'generic', 'coop' or 'rows'")

    def close(self):
        for r in getattr(self, "_steady_runners", {}).values():
            r.close()
        self._steady_runners = {}
        if getattr(self, "_h", None):
            lib().acmeb200_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ info
    @property
    def kernel_name(self) -> str:
        return lib().acmeb200_kernel_name(self._h).decode()

    @property
    def launch_count(self) -> int:
        return int(lib().acmeb200_launch_count(self._h))

    def stats(self) -> dict:
        s = Stats()
        check(lib().acmeb200_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def cache_sizes(self, sub: int = 0):
        """(sizes, capacity): solutions held by the learning CachingSolver per instance (solvers.jl:321-323)"""
        n = np.zeros(self.batch, dtype=np.int32)
        cap = C.c_int32(0)
        check(lib().acmeb200_get_cache_sizes(self._h, sub, n.ctypes.data_as(C.c_void_p), C.byref(cap)))
        return n, int(cap.value)

    def cache_info(self, sub: int = 0) -> dict:
        """the CachingSolver bookkeeping per instance (solvers.jl:321-325): num_ps, new_count, new_count_limit, the
        capacity the reference's doubling arrays would have, points in the current tree, flags"""
        a = np.zeros((self.batch, 8), dtype=np.int32)
        check(lib().acmeb200_get_cache_info(self._h, sub, a.ctypes.data_as(C.c_void_p)))
        return dict(num_ps=a[:, 0], new_count=a[:, 1], new_count_limit=a[:, 2], cap_ref=a[:, 3], tree_n=a[:, 4], flags=a[:, 5])

    def solver_state(self) -> bytes:
        """everything mutable -- x, extrapolation origins, learnt solution stores, status, statistics -- as one blob
        (``deepcopy(model)`` minus the matrices; acmeb200_get_solver_state)"""
        n = int(lib().acmeb200_solver_state_size(self._h))
        if n < 0:
            check(n)
        buf = C.create_string_buffer(n)
        check(lib().acmeb200_get_solver_state(self._h, buf, n))
        return buf.raw

    def set_solver_state(self, blob: bytes):
        """restore a blob of :meth:`solver_state` taken from a runner of the same model, batch and kernel"""
        check(lib().acmeb200_set_solver_state(self._h, blob, len(blob)))

    def extrapolation_origin(self, sub: int = 0):
        """``get_extrapolation_origin(solver)`` (solvers.jl:199) of every instance: (p (np, B), z (nn, B))"""
        s = self.model.subs[sub]
        p = np.zeros((s.np_, self.batch), order="F"); z = np.zeros((s.nn, self.batch), order="F")
        check(lib().acmeb200_get_extrapolation_origin(self._h, sub, p.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p)))
        return p, z

    def eval_jq(self, sub: int, q) -> np.ndarray:
        """Element Jacobians ``Jq`` (circuit.jl:10-17) of sub-problem `sub` at one ``q`` per instance, evaluated on the
        device with every instance's own element parameters: ``q`` (B, nq) -> (B, nn, nq)."""
        s = self.model.subs[sub]
        B = self.batch
        qh = np.ascontiguousarray(np.asarray(q, dtype=np.float64).reshape(B, s.nq).T)       # [nq][B]
        out = np.zeros((s.nq, s.nn, B))                                                       # [nq][nn][B]
        check(lib().acmeb200_eval_jq(self._h, sub, qh.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
        return np.ascontiguousarray(out.transpose(2, 1, 0))

    @property
    def device(self) -> int:
        d = C.c_int32(-1)
        check(lib().acmeb200_get_device(self._h, C.byref(d)))
        return int(d.value)

    def status(self):
        st = np.zeros(self.batch, dtype=np.uint32)
        ff = np.zeros(self.batch, dtype=np.int64)
        check(lib().acmeb200_get_status(self._h, st.ctypes.data_as(C.c_void_p), ff.ctypes.data_as(C.c_void_p)))
        return st, ff

    def reset(self):
        check(lib().acmeb200_reset(self._h))

    @property
    def x(self) -> np.ndarray:
        """model.x of every instance, shape (nx, B)"""
        out = np.zeros(max(self.model.nx * self.batch, 1))
        check(lib().acmeb200_get_state(self._h, out.ctypes.data_as(C.c_void_p)))
        return out[:self.model.nx * self.batch].reshape((self.model.nx, self.batch), order="F")

    @x.setter
    def x(self, val):
        if self.model.nx == 0:
            return
        v = np.asarray(val, dtype=np.float64).reshape(self.model.nx, -1)
        v = np.ascontiguousarray(np.broadcast_to(v, (self.model.nx, self.batch)).ravel(order="F"))
        if v.size:
            check(lib().acmeb200_set_state(self._h, v.ctypes.data_as(C.c_void_p), self.model.nx))

    # ------------------------------------------------------------------ run
    def _check_sizes(self, u_rows, u_cols, y_rows=None, y_cols=None):
        """checkiosizes (ACME.jl:625-635)"""
        m = self.model
        if u_rows != m.nu:
            raise DimensionMismatch(f"input matrix has {u_rows} rows, but model has {m.nu} inputs")
        if y_rows is not None and y_rows != m.ny:
            raise DimensionMismatch(f"output matrix has {y_rows} rows, but model has {m.ny} outputs")
        if y_cols is not None and u_cols != y_cols:
            raise DimensionMismatch(f"input matrix has {u_cols} columns, output matrix has {y_cols} columns")

    def run(self, u, y=None, *, stream=None, check_status: bool = True, layout: str = "instance"):
        """``run!`` for all instances.

        ``layout="sample"`` selects the sample-major streams of the C ABI
        (ACMEB200_SAMPLE_MAJOR; thread-per-instance kernels, specialised or generic): per-instance
        input (nu, B, N) and output (ny, B, N) in Julia layout -- torch tensors
        (N, B, nu) / (N, B, ny) -- so that one time step of all instances is
        contiguous.  A shared (nu, N) input is the same in both layouts.  The
        values are bit-identical to the default layout.

        u: numpy array (host) or torch CUDA tensor (device), shape (nu, N) -- one
        input shared by all instances -- or (nu, N, B) in Julia (column-major)
        layout, i.e. ``u[:, :, b]`` contiguous per instance.  Torch tensors must
        be float64 with that memory layout: shape (B, N, nu) C-contiguous is the
        same bytes and is accepted as such.
        Returns y with the matching convention.
        """
        m = self.model
        if layout not in ("instance", "sample"):
            raise ValueError("layout must be 'instance' or 'sample'")
        if layout == "sample":
            return self._run_sample_major(u, y, stream, check_status)
        if _is_torch_cuda(u):
            return self._run_torch(u, y, stream, check_status)
        u = np.asarray(u, dtype=np.float64)
        if u.ndim == 2:
            N = u.shape[1]
            self._check_sizes(u.shape[0], N)
            ubuf = np.ascontiguousarray(u.ravel(order="F"))
            ustride = 0
        elif u.ndim == 3:
            N = u.shape[1]
            self._check_sizes(u.shape[0], N)
            if u.shape[2] != self.batch:
                raise DimensionMismatch(f"input has {u.shape[2]} instances, runner has {self.batch}")
            ubuf = np.ascontiguousarray(u.ravel(order="F"))
            ustride = m.nu * N
        else:
            raise DimensionMismatch("u must be (nu, N) or (nu, N, B)")
        if y is None:
            ybuf = np.empty(max(m.ny * N * self.batch, 1))
        else:
            if y.ndim < 2 or y.shape[:2] != (m.ny, N):
                self._check_sizes(u.shape[0], N, y.shape[0] if y.ndim else -1, y.shape[1] if y.ndim > 1 else -1)
            # the instance axis too: acmeb200_run writes ny*N*batch doubles
            if not (y.shape == (m.ny, N, self.batch) or (y.ndim == 2 and self.batch == 1)):
                raise DimensionMismatch(f"output has shape {tuple(y.shape)}, runner needs {(m.ny, N, self.batch)}")
            if not (y.flags.f_contiguous and y.dtype == np.float64):
                raise ValueError("y must be a Fortran-contiguous float64 array")
            ybuf = y.reshape(-1, order="F") if y.size else np.empty(1)
        if ubuf.size == 0:
            ubuf = np.zeros(1)
        check(lib().acmeb200_run(self._h, ubuf.ctypes.data_as(C.c_void_p), ustride,
                                 ybuf.ctypes.data_as(C.c_void_p), m.ny * N, N, 0, None))
        if check_status:
            self.raise_for_status()
        if y is not None:
            return y
        return ybuf[:m.ny * N * self.batch].reshape((m.ny, N, self.batch), order="F")

    def _run_torch(self, u, y, stream, check_status):
        import torch
        m = self.model
        if u.dtype != torch.float64 or not u.is_contiguous():
            raise ValueError("device input must be a contiguous float64 tensor")
        # torch layout: (N, nu) shared or (B, N, nu), C-contiguous == Julia (nu, N[, B])
        if u.dim() == 2:
            N, nu = u.shape
            ustride = 0
        elif u.dim() == 3:
            B, N, nu = u.shape
            if B != self.batch:
                raise DimensionMismatch(f"input has {B} instances, runner has {self.batch}")
            ustride = N * nu
        else:
            raise DimensionMismatch("device u must be (N, nu) or (B, N, nu)")
        self._check_sizes(nu, N)
        if y is None:
            y = torch.empty((self.batch, N, m.ny), dtype=torch.float64, device=u.device)
        elif tuple(y.shape) != (self.batch, N, m.ny) or y.dtype != torch.float64 or not y.is_contiguous():
            raise DimensionMismatch(f"device y must be a contiguous float64 tensor of shape {(self.batch, N, m.ny)}")
        s = stream if stream is not None else torch.cuda.current_stream(u.device)
        check(lib().acmeb200_run(self._h, C.c_void_p(u.data_ptr()), ustride, C.c_void_p(y.data_ptr()),
                                 N * m.ny, N, _abi.U_DEVICE | _abi.Y_DEVICE, C.c_void_p(s.cuda_stream)))
        if check_status:
            self.raise_for_status()
        return y

    def _run_sample_major(self, u, y, stream, check_status):
        """sample-major streams: numpy (nu, B, N) / (ny, B, N) Fortran order, torch (N, B, nu) / (N, B, ny)"""
        m, B = self.model, self.batch
        if _is_torch_cuda(u):
            import torch
            if u.dtype != torch.float64 or not u.is_contiguous():
                raise ValueError("device input must be a contiguous float64 tensor")
            if u.dim() == 2:
                N, nu = u.shape
                ustride = 0
            elif u.dim() == 3:
                N, Bu, nu = u.shape
                if Bu != B:
                    raise DimensionMismatch(f"input has {Bu} instances, runner has {B}")
                ustride = B * nu
            else:
                raise DimensionMismatch("sample-major device u must be (N, nu) or (N, B, nu)")
            self._check_sizes(nu, N)
            if y is None:
                y = torch.empty((N, B, m.ny), dtype=torch.float64, device=u.device)
            elif tuple(y.shape) != (N, B, m.ny) or y.dtype != torch.float64 or not y.is_contiguous():
                raise DimensionMismatch(f"sample-major device y must be a contiguous float64 tensor of shape {(N, B, m.ny)}")
            s = stream if stream is not None else torch.cuda.current_stream(u.device)
            check(lib().acmeb200_run(self._h, C.c_void_p(u.data_ptr()), ustride, C.c_void_p(y.data_ptr()), B * m.ny, N,
                                     _abi.U_DEVICE | _abi.Y_DEVICE | _abi.SAMPLE_MAJOR, C.c_void_p(s.cuda_stream)))
            if check_status:
                self.raise_for_status()
            return y
        u = np.asarray(u, dtype=np.float64)
        if u.ndim == 2:
            N = u.shape[1]
            ustride = 0
        elif u.ndim == 3:
            N = u.shape[2]
            if u.shape[1] != B:
                raise DimensionMismatch(f"input has {u.shape[1]} instances, runner has {B}")
            ustride = m.nu * B
        else:
            raise DimensionMismatch("sample-major u must be (nu, N) or (nu, B, N)")
        self._check_sizes(u.shape[0], N)
        ubuf = np.ascontiguousarray(u.ravel(order="F"))
        if y is None:
            ybuf = np.empty(max(m.ny * N * B, 1))
        else:
            if y.shape != (m.ny, B, N):
                self._check_sizes(u.shape[0], N, y.shape[0], y.shape[-1])
                raise DimensionMismatch(f"sample-major y must have shape {(m.ny, B, N)}")
            if not (y.flags.f_contiguous and y.dtype == np.float64):
                raise ValueError("y must be a Fortran-contiguous float64 array")
            ybuf = y.reshape(-1, order="F") if y.size else np.empty(1)
        if ubuf.size == 0:
            ubuf = np.zeros(1)
        check(lib().acmeb200_run(self._h, ubuf.ctypes.data_as(C.c_void_p), ustride, ybuf.ctypes.data_as(C.c_void_p),
                                 m.ny * B, N, _abi.SAMPLE_MAJOR, None))
        if check_status:
            self.raise_for_status()
        if y is not None:
            return y
        return ybuf[:m.ny * N * B].reshape((m.ny, B, N), order="F")

    def run_host_pinned(self, u_ptr: int, ustride: int, y_ptr: int, N: int, layout: str = "instance"):
        """raw-pointer host run (used by bench.py for the end-to-end leg with pinned torch buffers);
        ``ustride`` is the instance pitch, or the sample pitch with ``layout="sample"``"""
        if layout == "sample":
            check(lib().acmeb200_run(self._h, C.c_void_p(u_ptr), ustride, C.c_void_p(y_ptr),
                                     self.model.ny * self.batch, N, _abi.SAMPLE_MAJOR, None))
            return
        check(lib().acmeb200_run(self._h, C.c_void_p(u_ptr), ustride, C.c_void_p(y_ptr),
                                 self.model.ny * N, N, 0, None))

    # ------------------------------------------------------------------ steady state (batched)
    def steadystate(self, u=None) -> np.ndarray:
        """Batched ``steadystate(model, u)`` (ACME.jl:474-497); see ``_steady``.  Returns x of shape (nx, B)."""
        return self._steady(u)[0]

    def _stack(self, name, default):
        """matrix `name` as a stack (Bk, rows, cols) with Bk = B for a per-instance override (this
        runner's slice of it), else 1"""
        ov = getattr(self, "_overrides", None) or {}
        B = self.batch
        arr = np.asarray(ov[name], dtype=np.float64) if name in ov else np.asarray(default, dtype=np.float64)
        base = np.asarray(default)
        if arr.ndim == base.ndim + 1:
            arr = arr[..., self.first:self.first + B] if arr.shape[-1] != B else arr
            arr = np.moveaxis(arr, -1, 0)
        else:
            arr = arr[None]
        return arr.reshape(arr.shape[0], *(base.shape if base.ndim == 2 else (base.shape[0], 1)))

    def _steady(self, u=None):
        """Batched ``steadystate(model, u)`` (ACME.jl:474-497): the state every instance settles to
        for a constant input ``u`` (``(nu,)`` shared or ``(nu, B)`` per instance).  The non-linear
        part -- one Newton/homotopy solve per sub-problem with the state recursion folded into
        ``fq``/``q0`` and tolerance 1e-15 -- runs on the device through the same C ABI: a derived
        zero-state model whose single "sample" is that solve.  Returns (x (nx, B), z (nn_total, B),
        u (nu, B))."""
        from .model import DiscreteModel, SubProblem
        m = self.model
        B = self.batch
        u = np.zeros((m.nu, B)) if u is None else np.broadcast_to(
            np.asarray(u, dtype=np.float64).reshape(m.nu, -1), (m.nu, B)).copy()
        stk = self._stack
        A_, B_, C_, X0 = stk("a", m.a), stk("b", m.b), stk("c", m.c), stk("x0", m.x0)
        IAinv = np.linalg.inv(np.eye(m.nx)[None] - A_) if m.nx else np.zeros((1, 0, 0))
        uB = u.T[:, :, None]                                        # (B, nu, 1)
        if not m.subs:  # linear: x = (I - a) \ (b u + x0), possibly per-instance matrices
            return ((IAinv @ (B_ @ uB + X0))[:, :, 0].T if m.nx else np.zeros((0, B))), np.zeros((0, B)), u
        nnt = m.nn_total
        steady_z = np.zeros((nnt, B))
        zoff = 0
        for i, s in enumerate(m.subs):
            dq, eq, fqprev = stk(f"dq{i}", s.dq), stk(f"eq{i}", s.eq), stk(f"fqprev{i}", s.fqprev)
            pexp, q0, fqm = stk(f"pexp{i}", s.pexp), stk(f"q0{i}", s.q0), stk(f"fq{i}", s.fq)
            dqIA = dq @ IAinv if m.nx else dq
            Eu = pexp @ (dqIA @ B_ + eq) if m.nx else pexp @ eq
            Ez = pexp @ ((dqIA @ C_ if m.nx else 0.0) + fqprev)
            const = q0 + (pexp @ dqIA @ X0 if m.nx else 0.0)
            fq = (pexp @ dqIA @ C_[:, :, zoff:zoff + s.nn] if m.nx else 0.0) + fqm
            nin = m.nu + nnt + 1
            bk = max(Eu.shape[0], Ez.shape[0], const.shape[0], fq.shape[0])   # 1: every matrix shared
            eq_d = np.concatenate([np.broadcast_to(Eu, (bk,) + Eu.shape[1:]), np.broadcast_to(Ez, (bk,) + Ez.shape[1:]),
                                   np.broadcast_to(const, (bk,) + const.shape[1:])], axis=2)   # (bk, nq, nin)
            fq_d = np.broadcast_to(fq, (bk,) + fq.shape[1:])
            sub = SubProblem(nn=s.nn, nq=s.nq, np_=s.nq, dq=np.zeros((s.nq, 0)),
                             eq=eq_d[0], fqprev=np.zeros((s.nq, s.nn)),
                             pexp=np.eye(s.nq), q0=np.zeros(s.nq), fq=fq_d[0], init_z=np.zeros(s.nn), elems=s.elems)
            sm = DiscreteModel.from_matrices(a=np.zeros((0, 0)), b=np.zeros((0, nin)), c=np.zeros((0, s.nn)),
                                             x0=np.zeros(0), dy=np.zeros((s.nn, 0)), ey=np.zeros((s.nn, nin)),
                                             fy=np.eye(s.nn), y0=np.zeros(s.nn), subs=[sub],
                                             solver="HomotopySolver{SimpleSolver}")
            params = None
            if getattr(self, "_params", None) is not None and self._params[i] is not None:
                params = [self._params[i][:, self.first:self.first + B]]
            sov = None
            if bk > 1:  # per-instance matrices: the derived model has per-instance eq / fq
                sov = {"eq0": np.asfortranarray(np.moveaxis(eq_d, 0, -1)), "fq0": np.asfortranarray(np.moveaxis(fq_d, 0, -1))}
            # the derived model does not depend on u: one device model per sub-problem for the life of this runner, reset
            # before every use (the reference builds a fresh solver per call, ACME.jl:481-485)
            cache = self.__dict__.setdefault("_steady_runners", {})
            r = cache.get(i)
            if r is None:
                r = cache[i] = BatchRunner(sm, B, params=params, overrides=sov, tol=1e-15)
            else:
                r.reset()
            uin = np.asfortranarray(np.vstack([u, steady_z, np.ones((1, B))]).reshape(nin, 1, B))
            z = r.run(uin, check_status=False)[:, 0, :]
            if (r.status()[0] != 0).any():
                raise RuntimeError("Failed to find steady state solution")  # ACME.jl:492
            steady_z[zoff:zoff + s.nn] = z
            zoff += s.nn
        if not m.nx:
            return np.zeros((0, B)), steady_z, u
        rhs = B_ @ uB + C_ @ steady_z.T[:, :, None] + X0
        return (IAinv @ rhs)[:, :, 0].T, steady_z, u

    def steadystate_(self, u=None) -> np.ndarray:
        """``steadystate!``: also makes it the current state of every instance (ACME.jl:499-503)."""
        x = self.steadystate(u)
        self.x = x
        return x

    # ------------------------------------------------------------------ linearize (batched)
    def linearize(self, usteady=None, **runner_kw) -> "BatchRunner":
        """Batched ``linearize(model, usteady)`` (ACME.jl:505-550): the small-signal linear model of
        every instance around its own steady state, as a new runner of B linear instances (which the
        linear path of the sample loop then runs at memory speed).  The steady states come from the
        device (``_steady``), and so do the element Jacobians at those points (``eval_jq``: one launch for the
        batch); the chain rule through the sub-problems (ACME.jl:520-546) is batched numpy.  Instances that share
        everything give a runner without per-instance matrices."""
        from .model import DiscreteModel
        m, B = self.model, self.batch
        xs, zs, u = self._steady(usteady)
        stk = self._stack
        bc = lambda arr: np.broadcast_to(arr, (B,) + arr.shape[1:])
        a, b, c, x0 = (bc(stk(n, d)).copy() for n, d in (("a", m.a), ("b", m.b), ("c", m.c), ("x0", m.x0)))
        dy, ey, fy, y0 = (bc(stk(n, d)).copy() for n, d in (("dy", m.dy), ("ey", m.ey), ("fy", m.fy), ("y0", m.y0)))
        xB, uB, zB = xs.T[:, :, None], u.T[:, :, None], zs.T[:, :, None]
        zranges, dzdps, dqlins, eqlins = [], [], [], []
        zoff = 0
        for i, s in enumerate(m.subs):
            dq, eq, fqprev = bc(stk(f"dq{i}", s.dq)), bc(stk(f"eq{i}", s.eq)), bc(stk(f"fqprev{i}", s.fqprev))
            pexp, q0, fq = bc(stk(f"pexp{i}", s.pexp)), bc(stk(f"q0{i}", s.q0)), bc(stk(f"fq{i}", s.fq))
            zr = slice(zoff, zoff + s.nn)
            psteady = dq @ xB + eq @ uB + fqprev @ zB                     # (B, np, 1)
            q = (q0 + pexp @ psteady + fq @ zB[:, zr])[:, :, 0]            # (B, nq)
            Jq = self.eval_jq(i, q)                                        # (B, nn, nq), one launch for the batch
            try:
                dzdp = -np.linalg.solve(Jq @ fq, Jq @ pexp)               # solvers.jl:407-414
            except np.linalg.LinAlgError:
                raise ValueError("Cannot linearize: singular Jacobian at the steady state")
            fqdzdps = [fqprev[:, :, zranges[n]] @ dzdps[n] for n in range(i)]
            dqlin = dq + sum((f @ d for f, d in zip(fqdzdps, dqlins)), np.zeros_like(dq))
            eqlin = eq + sum((f @ e for f, e in zip(fqdzdps, eqlins)), np.zeros_like(eq))
            zranges.append(zr); dzdps.append(dzdp); dqlins.append(dqlin); eqlins.append(eqlin)
            zlin = zB[:, zr] - dzdp @ psteady
            x0 = x0 + c[:, :, zr] @ zlin
            a = a + c[:, :, zr] @ dzdp @ dqlin
            b = b + c[:, :, zr] @ dzdp @ eqlin
            y0 = y0 + fy[:, :, zr] @ zlin
            dy = dy + fy[:, :, zr] @ dzdp @ dqlin
            ey = ey + fy[:, :, zr] @ dzdp @ eqlin
            zoff += s.nn
        lin0 = DiscreteModel.from_matrices(a=a[0], b=b[0], c=np.zeros((m.nx, 0)), x0=x0[0, :, 0], dy=dy[0],
                                           ey=ey[0], fy=np.zeros((m.ny, 0)), y0=y0[0, :, 0], subs=[],
                                           solver=m.solver)
        mats = {"a": a, "b": b, "dy": dy, "ey": ey, "x0": x0[:, :, 0], "y0": y0[:, :, 0]}
        ov = {k: np.asfortranarray(np.moveaxis(v, 0, -1)) for k, v in mats.items()
              if v.size and not np.all(v == v[:1])}
        return BatchRunner(lin0, B, overrides=ov or None, **runner_kw)

    def raise_for_status(self):
        """Re-raises the reference's failure semantics (ACME.jl:688-694)."""
        st, _ = self.status()
        if (st & _abi.STATUS_NONFINITE).any():
            raise RuntimeError(ERR_NONFINITE)
        if (st & _abi.STATUS_NOT_CONVERGED).any():
            warnings.warn(WARN_NOT_CONVERGED)


class _ShardView(BatchRunner):
    """one shard of a :class:`MultiGpuRunner`: the handle is owned by the multi object"""

    def __init__(self, model, handle, first, count, batch_total):
        self.model, self._h, self.first, self.batch, self.batch_total = model, handle, first, count, batch_total
        self._params = self._overrides = None
        self._has_overrides = False

    def close(self):
        self._h = None


class MultiGpuRunner:
    """The batch over all the GPUs of the box from ONE host process through the C ABI alone (acmeb200_multi_*): contiguous
    instance shards, one device model per GPU, no collective; ``run`` takes host arrays for the whole batch and drives
    every shard's host-buffer pipeline from its own host thread.  (``distributed.ShardedBatchRunner`` is the
    one-process-per-GPU variant for device-resident streams.)"""

    def __init__(self, model, batch: int, n_gpus: int = 0, **desc_kw):
        self.model, self.batch = model, batch
        self._holder = make_desc(model, batch, **desc_kw)
        h = C.c_void_p()
        check(lib().acmeb200_multi_create(C.byref(self._holder.desc), batch, n_gpus, C.byref(h)))
        self._h = h
        self.shards = []
        for g in range(lib().acmeb200_multi_shards(self._h)):
            first, count = C.c_int64(0), C.c_int64(0)
            mh = lib().acmeb200_multi_model(self._h, g, C.byref(first), C.byref(count))
            self.shards.append(_ShardView(model, C.c_void_p(mh), int(first.value), int(count.value), batch))

    def close(self):
        if getattr(self, "_h", None):
            for s in self.shards:
                s.close()
            lib().acmeb200_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, u, y=None, check_status: bool = True):
        """u: (nu, N) shared or (nu, N, B) Fortran-ordered host array; returns y (ny, N, B)"""
        m, B = self.model, self.batch
        u = np.asarray(u, dtype=np.float64)
        if u.ndim not in (2, 3) or u.shape[0] != m.nu:
            raise DimensionMismatch(f"input matrix has {u.shape[0]} rows, but model has {m.nu} inputs")
        N = u.shape[1]
        if u.ndim == 3 and u.shape[2] != B:
            raise DimensionMismatch(f"input has {u.shape[2]} instances, runner has {B}")
        ubuf = np.ascontiguousarray(u.ravel(order="F")) if u.size else np.zeros(1)
        if y is None:
            y = np.empty((m.ny, N, B), order="F")
        elif y.shape != (m.ny, N, B) or not y.flags.f_contiguous or y.dtype != np.float64:
            raise DimensionMismatch(f"output must be a Fortran-contiguous float64 array of shape {(m.ny, N, B)}")
        ybuf = y.reshape(-1, order="F") if y.size else np.empty(1)
        check(lib().acmeb200_multi_run(self._h, ubuf.ctypes.data_as(C.c_void_p), m.nu * N if u.ndim == 3 else 0,
                                       ybuf.ctypes.data_as(C.c_void_p), m.ny * N, N, 0))
        if check_status:
            for s in self.shards:
                s.raise_for_status()
        return y

    def stats(self) -> dict:
        tot = None
        for s in self.shards:
            d = s.stats()
            tot = d if tot is None else {k: ([a + b for a, b in zip(tot[k], d[k])] if isinstance(d[k], list) else tot[k] + d[k]) for k in d}
        return tot


class ModelRunner(BatchRunner):
    """``ModelRunner(model, showprogress)`` (ACME.jl:570-604) -- a batch of one."""

    def __init__(self, model, showprogress: bool = False, **kw):
        super().__init__(model, 1, **kw)
        self.showprogress = showprogress
        self.x = model.x


def run_(target, *args, showprogress=False, **kw):
    """``run!`` (ACME.jl:567, 619, 650/658).

    run_(model, u)        -> y, model.x is advanced
    run_(runner, u)       -> y
    run_(runner, y, u)    -> fills y in place
    """
    if isinstance(target, BatchRunner):
        runner = target
        if len(args) == 1:
            y = runner.run(args[0], **kw)
        elif len(args) == 2:
            y = runner.run(args[1], args[0], **kw)
        else:
            raise TypeError("run_(runner, [y,] u)")
        if isinstance(runner, ModelRunner) and hasattr(y, "ndim") and y.ndim == 3:
            y = y[:, :, 0]
        if isinstance(runner, ModelRunner):
            runner.model.x = runner.x[:, 0].copy()
        return y
    # run!(model, u): in the reference the solvers live in the DiscreteModel (ACME.jl:118-148), so the extrapolation
    # origin and what the CachingSolver has learnt persist from call to call.  Same here: the device runner is kept on
    # the model (one per set of runner options) and reused; only model.x is handed over each call, as ACME.jl:561-562 does.
    model = target
    (u,) = args
    key = repr(sorted(kw.items()))
    cache = model.__dict__.setdefault("_device_runners", {})
    runner = cache.get(key)
    if runner is None or getattr(runner, "_h", None) is None:
        runner = cache[key] = ModelRunner(model, showprogress, **kw)
    else:
        runner.x = model.x
    return run_(runner, u)
