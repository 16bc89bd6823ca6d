"""Runs in a subprocess with ACMEB200_LIB = the emulated library (tests/test_emu.py): the device
library's own kernel sources on the CPU against the oracle.  Prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np

import acme_jl_b200 as A
from acme_jl_b200 import BatchRunner, examples as ex
from oracle.oracle import OracleModel

assert os.environ.get("ACMEB200_LIB", "").endswith("libacmeb200_emu.so")
H, HC = "HomotopySolver{SimpleSolver}", "HomotopySolver{CachingSolver{SimpleSolver}}"
case = sys.argv[1]
sine = lambda n: np.sin(2 * np.pi * 1000 / 44100 * np.arange(n)).reshape(1, -1)
out = {"case": case}
if case == "superover_rows":
    m = ex.superover()
    B, N = 2, 70
    u = np.zeros((4, N, B), order="F"); u[0] = sine(N)[0][:, None]; u[1] = np.array([0.3, 0.8])[None, :]; u[2] = 0.5; u[3] = 1.0
    for solver in (H, HC):
        o = OracleModel(m, B, solver=solver); yref = o.run(u, threads=0)
        r = BatchRunner(m, B, solver=solver); y = r.run(u)
        out[solver] = dict(kernel=r.kernel_name, err=float(np.abs(y - yref).max() / np.abs(yref).max()),
                           hist=r.stats()["iter_hist"], hist_ref=o.stats()["iter_hist"],
                           hom=r.stats()["homotopy_solves"], hom_ref=o.stats()["homotopy_solves"])
        # chunked == one shot (state conversion rows-in-lanes <-> generic layout), and the generic kernel agrees
        r.reset()
        y2 = np.concatenate([r.run(np.asfortranarray(u[:, :31])), r.run(np.asfortranarray(u[:, 31:]))], axis=1)
        out[solver]["chunked_equal"] = bool(np.array_equal(y, y2))
        rg = BatchRunner(m, B, solver=solver, kernel="generic"); yg = rg.run(u)
        out[solver]["generic_diff"] = float(np.abs(yg - y).max() / np.abs(yref).max())
        r.close(); rg.close()
elif case == "baked_perinst":
    base, kw, B = A.derive_sweep(lambda d, t: ex.superover(d, t, 1.0), [(0.2, 0.5), (0.7, 0.4), (0.45, 0.9)], workers=1)
    u = sine(60)
    o = OracleModel(base, B, solver=H, **kw); yref = o.run(u, threads=0)
    r = BatchRunner(base, B, solver=H, **kw); y = r.run(u)
    out.update(kernel=r.kernel_name, err=float(np.abs(y - yref).max() / np.abs(yref).max()), hist=r.stats()["iter_hist"],
               hist_ref=o.stats()["iter_hist"])
    r.close()
elif case == "rows_4warp":
    # the 4-warp build of the rows kernel (normally chosen from 1185 instances per GPU), forced at a small batch:
    # a ragged last CTA, shared and per-instance matrices
    os.environ["ACMEB200_ROWS_SMALL_MAX"] = "0"
    base, kw, B = A.derive_sweep(lambda d, t: ex.superover(d, t, 1.0), [(0.15 + 0.12 * k, 0.3 + 0.1 * k) for k in range(6)], workers=1)
    u = sine(40)
    yref = OracleModel(base, B, solver=HC, **kw).run(u, threads=0)
    r = BatchRunner(base, B, solver=HC, **kw); y = r.run(u)
    out["perinst"] = dict(kernel=r.kernel_name, err=float(np.abs(y - yref).max() / np.abs(yref).max()), samples=int(r.stats()["samples"]))
    r.close()
    m = ex.superover()
    B, N = 5, 30
    us = np.zeros((4, N, B), order="F"); us[0] = sine(N)[0][:, None]; us[1] = ((np.arange(B) + 0.5) / B)[None, :]; us[2] = 0.5; us[3] = 1.0
    yref = OracleModel(m, B, solver=H).run(us, threads=0)
    r = BatchRunner(m, B, solver=H); y = r.run(us)
    out["shared"] = dict(kernel=r.kernel_name, err=float(np.abs(y - yref).max() / np.abs(yref).max()), samples=int(r.stats()["samples"]))
    r.close()
elif case == "steady":
    # batched steadystate (ACME.jl:474-497) through the ABI: shared matrices, per-instance matrices of a non-linear
    # model (derived zero-state model with per-instance eq/fq), per-instance matrices of a linear model
    m = ex.birdie(vol=0.8)
    r = BatchRunner(m, 3); xs = r.steadystate(np.array([[0.0, 0.1, 0.3]]))
    xs2 = r.steadystate(np.array([[0.3, 0.0, -0.2]])); r.close()   # second call: the derived device model is reused (reset)
    out["shared"] = max(float(np.abs(xs[:, b] - m.steadystate([uv])).max()) for b, uv in enumerate([0.0, 0.1, 0.3]))
    out["reuse"] = max(float(np.abs(xs2[:, b] - m.steadystate([uv])).max()) for b, uv in enumerate([0.3, 0.0, -0.2]))
    pts = [(0.2, 0.5), (0.7, 0.4), (0.45, 0.9)]
    base, kw, B = A.derive_sweep(lambda d, t: ex.superover(d, t, 1.0), pts, workers=1)
    r = BatchRunner(base, B, **kw); xs = r.steadystate_(np.array([[0.0, 0.05, -0.02]]))
    out["perinst"] = max(float(np.abs(xs[:, b] - ex.superover(*p, 1.0).steadystate([uv])).max())
                         for b, (p, uv) in enumerate(zip(pts, [0.0, 0.05, -0.02])))
    # started from its steady state with that constant input, every instance stays there (checksteady!, runtests.jl:664-682)
    y = r.run(np.asfortranarray(np.repeat(np.array([0.0, 0.05, -0.02])[None, None, :], 50, axis=1)))
    out["drift"] = float(np.abs(r.x - xs).max()); out["y_span"] = float(np.abs(y - y[:, :1, :]).max())
    r.close()
    Rs = [1e3, 5e3, 2e4]
    base, kw, B = A.derive_sweep(lambda R: ex.sallenkey(fs=96000, r1=R, r2=R), Rs, workers=1)
    r = BatchRunner(base, B, **kw); xs = r.steadystate(np.array([[0.3, 0.3, 0.3]])); r.close()
    out["linear"] = max(float(np.abs(xs[:, b] - ex.sallenkey(fs=96000, r1=R, r2=R).steadystate([0.3])).max()) for b, R in enumerate(Rs))
elif case == "linearize":
    # batched linearize (ACME.jl:505-550): per-instance steady inputs, element parameters and baked matrices;
    # each instance's small-signal model == the host linearize of the separately built circuit
    def lin_err(lr, refs):
        e = 0.0
        for b, ref in enumerate(refs):
            for nme in ("a", "b", "dy", "ey", "x0", "y0"):
                got = lr._stack(nme, getattr(lr.model, nme))
                got = got[b if got.shape[0] > 1 else 0]
                want = np.asarray(getattr(ref, nme), dtype=float).reshape(got.shape)
                e = max(e, float(np.abs(got - want).max() / max(1.0, np.abs(want).max())))
        return e
    m = ex.birdie(vol=0.8)
    us = [0.0, 0.1, -0.2]
    r = BatchRunner(m, 3); lr = r.linearize(np.array([us])); r.close()
    out["inputs"] = lin_err(lr, [m.linearize([uv]) for uv in us]); out["inputs_kernel"] = lr.kernel_name
    # the linearised batch run from its steady state tracks the non-linear one for a small signal (runtests.jl:673-682)
    N = 400
    u = np.asfortranarray(np.array(us)[None, None, :] + 1e-4 * sine(N)[:, :, None])
    lr.x = lr.steadystate(np.array([us])); yl = lr.run(u); lr.close()
    r = BatchRunner(m, 3); r.steadystate_(np.array([us])); yn = r.run(u); r.close()
    out["tracking"] = float(np.abs(yl - yn).max())
    base = ex.diodeclipper()
    isv, etav = [1e-15, 3e-14, 2e-13], [1.0, 1.5, 2.0]
    par = np.array([[i, e, 1.8 * i, e] for i, e in zip(isv, etav)]).T.copy()
    r = BatchRunner(base, 3, params=[par]); lr = r.linearize(np.array([[0.2, 0.2, 0.2]])); r.close()
    out["params"] = lin_err(lr, [ex.diodeclipper(is1=i, is2=1.8 * i, η1=e, η2=e).linearize([0.2]) for i, e in zip(isv, etav)])
    lr.close()
    pts = [(0.2, 0.5), (0.7, 0.4), (0.45, 0.9)]
    base, kw, B = A.derive_sweep(lambda d, t: ex.superover(d, t, 1.0), pts, workers=1)
    r = BatchRunner(base, B, **kw); lr = r.linearize(); r.close()
    out["baked"] = lin_err(lr, [ex.superover(*p, 1.0).linearize() for p in pts]); lr.close()
    # identical instances -> no per-instance matrices
    r = BatchRunner(m, 4); lr = r.linearize(); r.close()
    out["shared_has_overrides"] = bool(lr._has_overrides); lr.close()
elif case == "failure":
    m = ex.superover()
    B, N = 2, 25
    u = np.zeros((4, N, B), order="F"); u[0] = sine(N)[0][:, None] * np.array([3e2, 3e4])[None, :]; u[1] = 0.9; u[2] = 0.5; u[3] = 1.0
    for solver in (H, "SimpleSolver"):
        o = OracleModel(m, B, solver=solver); o.run(u, threads=0)
        r = BatchRunner(m, B, solver=solver); y = r.run(u, check_status=False)
        st, ff = r.status()
        out[solver] = dict(kernel=r.kernel_name, status=[int(v) for v in st], status_ref=[int(v) for v in o.status()[0]],
                           first=[int(v) for v in ff], first_ref=[int(v) for v in o.status()[1]],
                           hist=r.stats()["iter_hist"], hist_ref=o.stats()["iter_hist"], nan=int(np.isnan(y).sum()))
        r.close()
elif case == "coop":
    # the cooperative kernel: one warp per instance (compile-time superover shape) against the oracle and, bit for bit,
    # against the rows-in-registers kernel; sub-warp groups of 16 lanes with runtime dimensions on a model with three
    # non-linear sub-problems (the "simplified superover" of runtests.jl:751-756)
    m = ex.superover()
    B, N = 2, 40
    u = np.zeros((4, N, B), order="F"); u[0] = sine(N)[0][:, None]; u[1] = np.array([0.3, 0.8])[None, :]; u[2] = 0.5; u[3] = 1.0
    o = OracleModel(m, B, solver=HC); yref = o.run(u, threads=0)
    rr = BatchRunner(m, B, solver=HC, kernel="rows"); yr = rr.run(u)
    r = BatchRunner(m, B, solver=HC, kernel="coop"); y = r.run(u)
    out["static"] = dict(kernel=r.kernel_name, err=float(np.abs(y - yref).max() / np.abs(yref).max()), equals_rows=bool(np.array_equal(y, yr)),
                         hist_equals_rows=r.stats()["iter_hist"] == rr.stats()["iter_hist"])
    r.close(); rr.close()
    ms = A.DiscreteModel(ex.superover_circuit(1.0, 1.0, 1.0, vb_source=True), 1 / 44100)
    u1 = sine(60)
    o = OracleModel(ms, 1, solver=H); yref = o.run(u1, threads=0)
    r = BatchRunner(ms, 40, solver=H, kernel="coop"); y = r.run(u1)
    out["multisub"] = dict(kernel=r.kernel_name, nsub=len(ms.subs), err=float(np.abs(y[:, :, 0] - yref[:, :, 0]).max() / np.abs(yref).max()),
                           all_instances_equal=bool(np.all(y == y[:, :, :1])), hist=r.stats()["iter_hist"][:8],
                           hist_ref=[40 * v for v in o.stats()["iter_hist"][:8]])
    r.close()
elif case == "tpi":
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    # G1: the reference's doctest vector through the thread-per-instance kernel (TMA tiles, swizzle, mbarriers emulated)
    r = BatchRunner(ex.diodeclipper(), 1, solver=HC); y = r.run(cases.sine())[:, :, 0]
    out["g1"] = dict(kernel=r.kernel_name, n=int(y.shape[1]), first=[float(v) for v in y[0, :4]], last=[float(v) for v in y[0, -3:]],
                     hist=r.stats()["iter_hist"][:6])
    r.close()
    # a batch that fills neither its last warp nor its last tile, per-instance element parameters
    B, N = 70, 203
    Is = 10.0 ** (-16 + 4 * np.arange(B) / 69); eta = 1 + np.arange(B) / 69
    P = np.vstack([Is, eta, 1.8 * Is, eta])
    u = np.asfortranarray(np.repeat(sine(N)[:, :, None], B, axis=2))
    o = OracleModel(ex.diodeclipper(), B, params=[P], solver=H); yref = o.run(u, threads=0)
    r = BatchRunner(ex.diodeclipper(), B, params=[P], solver=H); y = r.run(u)
    out["clipper"] = dict(err=float(np.abs(y - yref).max() / np.abs(yref).max()), hist=r.stats()["iter_hist"][:8], hist_ref=o.stats()["iter_hist"][:8])
    r.close()
    # the host-buffer pipeline (time chunks, state carried from chunk to chunk): 1 MiB staging chunks = 3 chunks here
    N2 = 5003
    u2 = np.asfortranarray(np.repeat(sine(N2)[:, :, None], B, axis=2))
    r = BatchRunner(ex.diodeclipper(), B, params=[P], solver=HC); y_one = r.run(u2); r.close()
    os.environ["ACMEB200_CHUNK_MB"] = "1"
    r = BatchRunner(ex.diodeclipper(), B, params=[P], solver=HC); y_chunked = r.run(u2); launches = r.launch_count; r.close()
    del os.environ["ACMEB200_CHUNK_MB"]
    out["chunks"] = dict(equal=bool(np.array_equal(y_one, y_chunked)), launches=int(launches))
    # linear model, per-instance matrices: the whole-tile register path
    base, kw, B = A.derive_sweep(lambda R: ex.sallenkey(fs=96000, r1=R, r2=R), [1e3 * (1 + k) for k in range(37)], workers=1)
    u = np.asfortranarray(np.repeat(sine(101)[:, :, None], B, axis=2))
    yref = OracleModel(base, B, **kw).run(u, threads=0)
    r = BatchRunner(base, B, **kw); y = r.run(u)
    out["linear"] = dict(kernel=r.kernel_name, err=float(np.abs(y - yref).max() / np.abs(yref).max()))
    r.close()
    # birdie with white noise: homotopy, learning cache; tolerance tightened on both sides for the strict criterion
    rng = np.random.default_rng(1)
    u = np.asfortranarray(np.clip(0.2 * rng.standard_normal((1, 400, 5)), -1, 1))
    m = ex.birdie(vol=0.8)
    yref = OracleModel(m, 5, solver=H, tol=1e-13).run(u, threads=0)
    r = BatchRunner(m, 5, solver=HC, tol=1e-13); y = r.run(u)
    peak = np.abs(yref).max()
    out["birdie"] = dict(kernel=r.kernel_name, err=float(np.max(np.abs(y - yref) / np.maximum(np.abs(yref), 1e-3 * peak))),
                         stored=[int(v) for v in r.cache_sizes()[0]], bad=int((r.status()[0] != 0).sum()))
    r.close()
elif case == "sample_major":
    # ACMEB200_SAMPLE_MAJOR: (nu, B, N) / (ny, B, N) streams through the transposed tiles of k_tpi.  Every run is
    # compared BIT FOR BIT with the instance-major run of the same library (same arithmetic, other data movement)
    # and with the oracle.
    smaj = lambda a: np.asfortranarray(np.transpose(a, (0, 2, 1)))   # (c, N, B) -> (c, B, N)
    def both(mk, u, **kw):
        r = mk(); y = r.run(u); st = r.stats(); r.close()
        r = mk(); ys = r.run(smaj(u), layout="sample", **kw); sts = r.stats(); name = r.kernel_name; launches = r.launch_count; r.close()
        return y, ys, dict(kernel=name, equal=bool(np.array_equal(smaj(y), ys)), hist_equal=st["iter_hist"] == sts["iter_hist"],
                           samples=int(sts["samples"]), launches=int(launches))
    # non-linear, per-instance element parameters, a batch that fills neither its last warp nor its last tile
    B, N = 70, 203
    Is = 10.0 ** (-16 + 4 * np.arange(B) / 69); eta = 1 + np.arange(B) / 69
    P = np.vstack([Is, eta, 1.8 * Is, eta])
    amp = 0.5 + np.arange(B) / B                                      # distinct inputs, so a transposition error shows
    u = np.asfortranarray(sine(N)[:, :, None] * amp[None, None, :])
    mk = lambda: BatchRunner(ex.diodeclipper(), B, params=[P], solver=HC)
    y, ys, d = both(mk, u)
    yref = OracleModel(ex.diodeclipper(), B, params=[P], solver=HC).run(u, threads=0)
    d["err"] = float(np.abs(y - yref).max() / np.abs(yref).max())
    out["clipper"] = d
    # odd batch: the sample pitch nu*B is odd, the tensor maps cannot be built -> synchronous tile path
    B3 = 37
    u3 = np.asfortranarray(u[:, :, :B3])
    mk3 = lambda: BatchRunner(ex.diodeclipper(), B3, params=[P[:, :B3].copy()], solver=H)
    out["odd_batch"] = both(mk3, u3)[2]
    # the host-buffer pipeline in time chunks (1 MiB staging chunks) == one device-resident chunk
    N2 = 5003
    u2 = np.asfortranarray(sine(N2)[:, :, None] * amp[None, None, :])
    os.environ["ACMEB200_CHUNK_MB"] = "1"
    out["chunks"] = both(mk, u2)[2]
    del os.environ["ACMEB200_CHUNK_MB"]
    # one shared input, sample-major output; output written into a caller-provided array
    r = mk(); y1 = r.run(sine(N)); r.close()
    r = mk(); yo = np.zeros((1, B, N), order="F"); r.run(sine(N), yo, layout="sample"); r.close()
    out["shared_u"] = bool(np.array_equal(smaj(y1), yo))
    # linear model with per-instance matrices: the whole-tile register path
    base, kw, Bl = A.derive_sweep(lambda R: ex.sallenkey(fs=96000, r1=R, r2=R), [1e3 * (1 + k) for k in range(38)], workers=1)
    ul = np.asfortranarray(sine(101)[:, :, None] * (1 + np.arange(Bl))[None, None, :])
    yl, yls, d = both(lambda: BatchRunner(base, Bl, **kw), ul)
    d["err"] = float(np.abs(yl - OracleModel(base, Bl, **kw).run(ul, threads=0)).max() / np.abs(yl).max())
    out["linear"] = d
    # two input channels (birdie with the volume pot as an input): 512-byte tile rows; learning cache
    m = ex.birdie()
    Bb, Nb = 6, 150
    rng = np.random.default_rng(3)
    ub = np.zeros((2, Nb, Bb), order="F"); ub[0] = np.clip(0.2 * rng.standard_normal((Nb, Bb)), -1, 1); ub[1] = (0.3 + 0.1 * np.arange(Bb))[None, :]
    out["birdie_vol"] = both(lambda: BatchRunner(m, Bb, solver=HC), ub)[2]
    # state persists across calls in either layout: first half sample-major, second half instance-major
    r = mk(); ya = r.run(smaj(u[:, :100]), layout="sample"); yb = r.run(np.asfortranarray(u[:, 100:])); r.close()
    out["mixed_calls"] = bool(np.array_equal(np.concatenate([np.transpose(ya, (0, 2, 1)), yb], axis=1), y))
    # the generic thread-per-instance kernel takes the flag too (any model the ABI can describe) ...
    mkg = lambda: BatchRunner(ex.diodeclipper(), B3, params=[P[:, :B3].copy()], solver=HC, kernel="generic")
    out["generic"] = both(mkg, u3)[2]
    # ... the lane-parallel kernels refuse it loudly
    ms = ex.superover()
    us = np.zeros((4, 8, 2), order="F"); us[1:] = 0.5
    try:
        r = BatchRunner(ms, 2, solver=H); r.run(smaj(us), layout="sample")
        out["rows_refused"] = False
    except Exception as e:
        out["rows_refused"] = "thread-per-instance" in str(e)
elif case == "eval_jq":
    # acmeb200_eval_jq: Jq of the whole batch (circuit.jl:10-17) == the host element laws, per-instance parameters
    from dataclasses import replace
    from acme_jl_b200 import hostsolve
    rng = np.random.default_rng(3)
    B = 37
    m = ex.diodeclipper()
    P = np.vstack([10.0 ** rng.uniform(-16, -12, B), rng.uniform(1, 2, B), 10.0 ** rng.uniform(-16, -12, B), rng.uniform(1, 2, B)])
    r = BatchRunner(m, B, params=[P]); s = m.subs[0]
    q = rng.uniform(-0.6, 0.6, (B, s.nq)); J = r.eval_jq(0, q); r.close()
    err = 0.0
    for k in range(B):
        table, o = [], 0
        for e, off in s.elems:
            table.append((replace(e, params=tuple(P[o:o + len(e.params), k])), off)); o += len(e.params)
        want = np.asarray(hostsolve.eval_table(table, q[k], s.nn)[1], dtype=float)
        err = max(err, float(np.max(np.abs(J[k] - want) / np.maximum(np.abs(want), 1e-300))))
    out["clipper"] = err
    m = ex.superover(); r = BatchRunner(m, 3); s = m.subs[0]
    q = rng.uniform(-0.4, 0.4, (3, s.nq)); J = r.eval_jq(0, q); r.close()
    out["superover"] = max(float(np.max(np.abs(J[k] - np.asarray(hostsolve.eval_table(s.elems, q[k], s.nn)[1], dtype=float)))) for k in range(3))
    out["shape"] = list(J.shape) == [3, s.nn, s.nq]
print(json.dumps(out))
