"""Element library (host side; runs once per model, never on the hot path).

Mirror of the reference's ``Element`` struct (/root/reference/src/ACME.jl:21-112)
and of the element constructors in /root/reference/src/elements.jl.  The linear
stamps are kept as exact rationals (``fractions.Fraction``; the reference
converts every stamp to ``Rational{BigInt}`` at /root/reference/src/circuit.jl:42-46).

The reference stores non-linear element laws as Julia closures
(elements.jl:25-30, 107-129, 238-244, 323-401, 453-479, 540-546).  A closure
cannot cross a C ABI, so each element here carries an *element-table* entry
instead: ``NLElem(kind, params, nq, nn)``.  ``kind``/``params`` layouts are the
ones declared in ``include/acmeb200.h`` and are shared by the CUDA kernels, the
C-ABI and the CPU oracle.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from fractions import Fraction
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# element kinds -- keep in sync with include/acmeb200.h
KIND_DIODE = 1
KIND_BJT = 2
KIND_POT = 3
KIND_MOSFET = 4
KIND_OPAMP_TANH = 5
KIND_JA = 6

NPARAMS = {KIND_DIODE: 2, KIND_BJT: 14, KIND_POT: 1, KIND_MOSFET: 12,
           KIND_OPAMP_TANH: 2, KIND_JA: 5}
KIND_NQ = {KIND_DIODE: 2, KIND_BJT: 4, KIND_POT: 5, KIND_MOSFET: 3,
           KIND_OPAMP_TANH: 2, KIND_JA: 4}
KIND_NN = {KIND_DIODE: 1, KIND_BJT: 2, KIND_POT: 2, KIND_MOSFET: 1,
           KIND_OPAMP_TANH: 1, KIND_JA: 1}


def fr(x) -> Fraction:
    """Exact conversion, like ``Rational{BigInt}(x)`` in the reference."""
    if isinstance(x, Fraction):
        return x
    if isinstance(x, (int, np.integer)):
        return Fraction(int(x))
    return Fraction(float(x))


def rmat(x, ncols_hint: Optional[int] = None) -> np.ndarray:
    """``hcat(x)`` of ACME.jl:35: scalar -> 1x1, vector -> column, matrix as is."""
    a = np.array(x, dtype=object)
    if a.ndim == 0:
        a = a.reshape(1, 1)
    elif a.ndim == 1:
        a = a.reshape(-1, 1)
    out = np.empty(a.shape, dtype=object)
    for idx, v in np.ndenumerate(a):
        out[idx] = fr(v)
    return out


def rzeros(m: int, n: int) -> np.ndarray:
    out = np.empty((m, n), dtype=object)
    out[...] = Fraction(0)
    return out


def reye(n: int, sign: int = 1) -> np.ndarray:
    out = rzeros(n, n)
    for i in range(n):
        out[i, i] = Fraction(sign)
    return out


@dataclass
class NLElem:
    """One row of the element table (replaces a ``nonlinear_eq`` closure)."""
    kind: int
    params: Tuple[float, ...]
    nq: int
    nn: int


_MAT_DIMS = {
    "mv": ("nl", "nb"), "mi": ("nl", "nb"), "mx": ("nl", "nx"),
    "mxd": ("nl", "nx"), "mq": ("nl", "nq"), "mu": ("nl", "nu"),
    "u0": ("nl", "n0"),
    "pv": ("ny", "nb"), "pi": ("ny", "nb"), "px": ("ny", "nx"),
    "pxd": ("ny", "nx"), "pq": ("ny", "nq"),
}


class Element:
    """ACME.jl:58-98.  ``nl_elems`` replaces the ``nonlinear_eq`` closure."""

    def __init__(self, nl_elems: Sequence[NLElem] = (), ports=None, pins=None, **mats):
        matrices: Dict[str, np.ndarray] = {}
        sizes: Dict[str, int] = {"n0": 1}
        for name, val in mats.items():
            if name not in _MAT_DIMS:
                raise TypeError(f"unknown element matrix {name}")
            if val is None:
                continue
            m = rmat(val)
            matrices[name] = m
            for sym, s in zip(_MAT_DIMS[name], m.shape):
                if sizes.setdefault(sym, s) != s:
                    raise ValueError(f"Inconsistent sizes for {sym}")
        for name, (r, c) in _MAT_DIMS.items():
            if name not in matrices:
                matrices[name] = rzeros(sizes.setdefault(r, 0), sizes.setdefault(c, 0))
        self.m = matrices
        self.nl_elems: List[NLElem] = list(nl_elems)
        if ports is not None:
            pins = {}
            for b, (p_plus, p_minus) in enumerate(ports, start=1):
                pins.setdefault(str(p_plus), []).append((b, 1))
                pins.setdefault(str(p_minus), []).append((b, -1))
        if pins is None:
            pins = {str(i): [((i + 1) // 2, 2 * (i % 2) - 1)] for i in range(1, 2 * sizes["nb"] + 1)}
        self.pins: Dict[str, List[Tuple[int, int]]] = pins

    # sizes, ACME.jl:105-110
    @property
    def nb(self): return self.m["mv"].shape[1]
    @property
    def nx(self): return self.m["mx"].shape[1]
    @property
    def nq(self): return self.m["mq"].shape[1]
    @property
    def nu(self): return self.m["mu"].shape[1]
    @property
    def nl(self): return self.m["mv"].shape[0]
    @property
    def ny(self): return self.m["pv"].shape[0]
    @property
    def nn(self): return self.nb + self.nx + self.nq - self.nl

    def __eq__(self, other):
        if not isinstance(other, Element):
            return NotImplemented
        return (all(self.m[k].shape == other.m[k].shape and (self.m[k] == other.m[k]).all()
                    for k in _MAT_DIMS)
                and self.pins == other.pins
                and [(e.kind, e.params) for e in self.nl_elems]
                == [(e.kind, e.params) for e in other.nl_elems])


# ---------------------------------------------------------------- constructors
def resistor(r):
    """elements.jl:16"""
    return Element(mv=-1, mi=r)


def potentiometer(r, pos=None):
    """elements.jl:18-31.  With ``pos`` baked in it is linear; without, ``pos``
    becomes a model input and the element is non-linear (kind POT)."""
    if pos is not None:
        # products are formed in Float64 first, as the reference does (elements.jl:18)
        return Element(mv=[[-1, 0], [0, -1]],
                       mi=[[r * pos, 0], [0, r * (1 - pos)]],
                       ports=[(1, 2), (2, 3)])
    return Element(mv=[[1, 0], [0, 1], [0, 0], [0, 0], [0, 0]],
                   mi=[[0, 0], [0, 0], [1, 0], [0, 1], [0, 0]],
                   mq=[[-1 if i == j else 0 for j in range(5)] for i in range(5)],
                   mu=[0, 0, 0, 0, -1],
                   nl_elems=[NLElem(KIND_POT, (float(r),), 5, 2)],
                   ports=[(1, 2), (2, 3)])


def capacitor(c):
    """elements.jl:40"""
    return Element(mv=[c, 0], mi=[0, 1], mx=[-1, 0], mxd=[0, -1])


def inductor(l=None, *, ja: bool = False, n=230, **kw):
    """elements.jl:49 (linear) and :167-168 (Jiles-Atherton, ``ja=True``)."""
    if ja:
        return transformer(ja=True, ns=[n], **kw)
    return Element(mv=[1, 0], mi=[0, l], mx=[0, -1], mxd=[-1, 0])


def transformer(l1=None, l2=None, *, coupling_coefficient=1, mutual_coupling=None,
                ja: bool = False, **kw):
    """elements.jl:63-68 (linear) and :100-135 (Jiles-Atherton)."""
    if ja:
        return _transformer_ja(**kw)
    if mutual_coupling is None:
        mutual_coupling = coupling_coefficient * math.sqrt(l1 * l2)
    return Element(mv=[[1, 0], [0, 1], [0, 0], [0, 0]],
                   mi=[[0, 0], [0, 0], [l1, mutual_coupling], [mutual_coupling, l2]],
                   mx=[[0, 0], [0, 0], [-1, 0], [0, -1]],
                   mxd=[[-1, 0], [0, -1], [0, 0], [0, 0]],
                   ports=[("primary1", "primary2"), ("secondary1", "secondary2")])


def _transformer_ja(D=2.4e-2, A=4.54e-5, ns=(), a=14.1, α=5e-5, c=0.55, k=17.8, Ms=2.75e5):
    """elements.jl:104-135"""
    μ0 = 1.2566370614e-6
    ns = list(ns)
    w = len(ns)
    mv = [[1 if i == j else 0 for j in range(w)] for i in range(w + 5)]
    mi = [[0] * w for _ in range(w)] + [list(ns)] + [[0] * w for _ in range(4)]
    mx = [[0, 0] for _ in range(w)] + [[-math.pi * D, 0], [-1 / a, -α / a], [0, -1], [0, 0], [0, 0]]
    mxd = [[-μ0 * A * n_, -μ0 * n_ * A] for n_ in ns] + [[0, 0], [0, 0], [0, 0], [-1, 0], [0, -1]]
    mq = [[0] * 4 for _ in range(w + 1)] + [[1 if i == j else 0 for j in range(4)] for i in range(4)]
    return Element(mv=np.array(mv, dtype=object).reshape(w + 5, w),
                   mi=np.array(mi, dtype=object).reshape(w + 5, w),
                   mx=mx, mxd=mxd, mq=mq,
                   nl_elems=[NLElem(KIND_JA, (float(Ms), float(a), float(α), float(c), float(k)), 4, 1)])


def voltagesource(v=None, *, rs=0):
    """elements.jl:181-183"""
    if v is None:
        return Element(mv=1, mi=-rs, mu=1, ports=[("+", "-")])
    return Element(mv=1, mi=-rs, u0=v, ports=[("+", "-")])


def currentsource(i=None, *, gp=0):
    """elements.jl:197-199"""
    if i is None:
        return Element(mv=gp, mi=-1, mu=1, ports=[("+", "-")])
    return Element(mv=gp, mi=-1, u0=i, ports=[("+", "-")])


def voltageprobe(*, gp=0):
    """elements.jl:210-211"""
    return Element(mv=-gp, mi=1, pv=1, ports=[("+", "-")])


def currentprobe(*, rs=0):
    """elements.jl:223-224"""
    return Element(mv=1, mi=-rs, pi=1, ports=[("+", "-")])


def diode(*, is_=1e-12, η=1, **kw):
    """elements.jl:235-245.  ``is`` is a Python keyword: pass ``is_=`` (or ``**{'is': x}``)."""
    if "is" in kw:
        is_ = kw.pop("is")
    if kw:
        raise TypeError(f"unexpected arguments {list(kw)}")
    return Element(mv=[1, 0], mi=[0, 1], mq=[[-1, 0], [0, -1]], ports=[("+", "-")],
                   nl_elems=[NLElem(KIND_DIODE, (float(is_), float(η)), 2, 1)])


def bjt(typ, *, is_=1e-12, η=1, isc=None, ise=None, ηc=None, ηe=None, βf=1000, βr=10,
        ile=0, ilc=0, ηcl=None, ηel=None, vaf=math.inf, var=math.inf,
        ikf=math.inf, ikr=math.inf, re=0, rc=0, rb=0, **kw):
    """elements.jl:307-406"""
    if "is" in kw:
        is_ = kw.pop("is")
    if kw:
        raise TypeError(f"unexpected arguments {list(kw)}")
    isc = is_ if isc is None else isc
    ise = is_ if ise is None else ise
    ηc = η if ηc is None else ηc
    ηe = η if ηe is None else ηe
    ηcl = ηc if ηcl is None else ηcl
    ηel = ηe if ηel is None else ηel
    if typ == "npn":
        polarity = 1
    elif typ == "pnp":
        polarity = -1
    else:
        raise ValueError(f"Unknown bjt type {typ}, must be :npn or :pnp")
    params = tuple(float(v) for v in (ise, isc, ηe, ηc, βf, βr, ile, ilc, ηel, ηcl, vaf, var, ikf, ikr))
    return Element(mv=[[1, 0], [0, 1], [0, 0], [0, 0]],
                   mi=[[-(re + rb), -rb], [-rb, -(rc + rb)], [1, 0], [0, 1]],
                   mq=[[-polarity if i == j else 0 for j in range(4)] for i in range(4)],
                   nl_elems=[NLElem(KIND_BJT, params, 4, 2)],
                   ports=[("base", "emitter"), ("base", "collector")])


def mosfet(typ, *, vt=0.7, α=2e-5, λ=0):
    """elements.jl:433-481.  Polynomial ``vt``/``α`` limited to 4 coefficients."""
    if typ == "n":
        polarity = 1
    elif typ == "p":
        polarity = -1
    else:
        raise ValueError(f"Unknown mosfet type {typ}, must be :n or :p")
    vt = tuple(float(v) for v in (vt if isinstance(vt, (tuple, list)) else (vt,)))
    α = tuple(float(v) for v in (α if isinstance(α, (tuple, list)) else (α,)))
    if len(vt) > 4 or len(α) > 4:
        raise ValueError("at most 4 polynomial coefficients are supported for vt and α")
    params = (float(polarity), float(λ), float(len(vt)), float(len(α))) \
        + vt + (0.0,) * (4 - len(vt)) + α + (0.0,) * (4 - len(α))
    return Element(mv=[[-1, 0], [0, -1], [0, 0], [0, 0]],
                   mi=[[0, 0], [0, 0], [0, -1], [1, 0]],
                   mq=[[polarity * v for v in row] for row in ([1, 0, 0], [0, 1, 0], [0, 0, 1], [0, 0, 0])],
                   nl_elems=[NLElem(KIND_MOSFET, params, 3, 1)],
                   ports=[("gate", "source"), ("drain", "source")])


def opamp(*args, maxgain=math.inf, gain_bw_prod=math.inf):
    """elements.jl:508-517 (linear) and :536-551 (``opamp('macak', gain, vomin, vomax)``)."""
    if args:
        if args[0] != "macak" or len(args) != 4:
            raise TypeError("opamp('macak', gain, vomin, vomax)")
        _, gain, vomin, vomax = args
        offset = 0.5 * (vomin + vomax)
        scale = 0.5 * (vomax - vomin)
        return Element(mv=[[0, 0], [1, 0], [0, 1]], mi=[[1, 0], [0, 0], [0, 0]],
                       mq=[[0, 0], [-1, 0], [0, -1]], u0=[0, 0, offset],
                       nl_elems=[NLElem(KIND_OPAMP_TANH, (float(gain), float(scale)), 2, 1)],
                       ports=[("in+", "in-"), ("out+", "out-")])
    if gain_bw_prod == math.inf:
        inv_gain = 0.0 if maxgain == math.inf else 1 / maxgain
        return Element(mv=[[0, 0], [1, -inv_gain]], mi=[[1, 0], [0, 0]],
                       ports=[("in+", "in-"), ("out+", "out-")])
    # 1/sqrt(1-1/maxgain^2), 1/sqrt(maxgain^2-1) with maxgain=Inf -> 1, 0 as in IEEE
    g1 = -1.0 if maxgain == math.inf else -1 / math.sqrt(1 - 1 / maxgain ** 2)
    g2 = 0.0 if maxgain == math.inf else 1 / math.sqrt(maxgain ** 2 - 1)
    return Element(mv=[[0, 0], [g1, 0], [0, -1]], mi=[[1, 0], [0, 0], [0, 0]],
                   mx=[0, g2, 1], mxd=[0, 1 / (2 * math.pi * gain_bw_prod), 0],
                   ports=[("in+", "in-"), ("out+", "out-")])
