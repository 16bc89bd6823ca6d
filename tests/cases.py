"""Test circuits shared by the oracle tests (CPU) and the CUDA parity tests (GPU).

Each builder mirrors a circuit of the reference's test suite
(/root/reference/test/runtests.jl, line numbers in the docstrings).
"""
import math
from fractions import Fraction

import numpy as np

import acme_jl_b200 as A
from acme_jl_b200 import (bjt, capacitor, circuit, currentprobe, currentsource, diode, mosfet,
                          opamp, resistor, voltageprobe, voltagesource)


def sine(n=44100, f=1000.0, fs=44100.0):
    """``sin.(2π*1000/44100*(0:44099)')`` (runtests.jl:689, 700)"""
    return np.sin(2 * np.pi * f / fs * np.arange(n)).reshape(1, -1)


def rc_ladder():
    """docs/src/ug.md:40-56: cascade of 20 RC low-passes"""
    c = circuit([("src", voltagesource(), {"-": "gnd"}), ("output", voltageprobe(), {"-": "gnd"})])
    pin = ("src", "+")
    for _ in range(20):
        r = c.add(resistor(1000))
        k = c.add(capacitor(10e-9))
        c.connect((r, "1"), pin)
        c.connect((r, "2"), (k, "1"))
        c.connect((k, "2"), "gnd")
        pin = (r, "2")
    c.connect(pin, ("output", "+"))
    return c


def resistor_diode():
    """runtests.jl:68-86"""
    i, r, is_ = 1e-3, 10e3, 1e-12
    v_r = i * r
    v_d = 25e-3 * math.log(i / is_ + 1)
    c = circuit([
        ("vsrc", voltagesource(v_r + v_d), {"+": "supply voltage", "-": "gnd"}),
        ("r1", resistor(r), {}),
        ("d", diode(is_=is_), {"-": "gnd", "+": ("r1", "2")}),
        ("vprobe", voltageprobe(), {"-": "gnd", "+": ("r1", "2")}),
    ])
    c.connect(("r1", "1"), "supply voltage")
    return c, v_d


def no_solution():
    """runtests.jl:170-183: diode driven by a current source"""
    return circuit([
        ("d", diode(), {}),
        ("src", currentsource(), {"+": ("d", "+"), "-": ("d", "-")}),
        ("probe", voltageprobe(), {"+": ("d", "+"), "-": ("d", "-")}),
    ])


def three_diodes():
    """runtests.jl:267-279"""
    c = circuit([
        ("src1", voltagesource(), {}),
        ("probe1", currentprobe(), {}),
        ("d1", diode(), {"+": ("src1", "+")}),
        ("d2", diode(), {"+": ("d1", "-"), "-": ("probe1", "+")}),
    ])
    c.connect(("probe1", "-"), ("src1", "-"))
    c.add("src2", voltagesource())
    c.add("probe2", currentprobe())
    c.add("d3", diode())
    c.connect(("src2", "+"), ("d3", "+"))
    c.connect(("d3", "-"), ("probe2", "+"))
    c.connect(("probe2", "-"), ("src2", "-"))
    return c


BJT_BASE = dict(isc=1e-6, ise=2e-6, ηc=1.1, ηe=1.0, βf=100, βr=10)


def bjt_circuit(typ, **kw):
    """runtests.jl:490-498 / 518-527"""
    params = dict(BJT_BASE)
    params.update(kw)
    return circuit([
        ("t", bjt(typ, **params), {}),
        ("isrc", currentsource(), {"+": ("t", "base")}),
        ("vsrc", voltagesource(), {"-": ("isrc", "-")}),
        ("veprobe", voltageprobe(), {"+": ("t", "base"), "-": ("isrc", "-")}),
        ("vcprobe", voltageprobe(), {"+": ("t", "base"), "-": ("vsrc", "+")}),
        ("ieprobe", currentprobe(), {"+": ("t", "emitter"), "-": ("isrc", "-")}),
        ("icprobe", currentprobe(), {"+": ("t", "collector"), "-": ("vsrc", "+")}),
    ])


def bjt_input(typ, N=100):
    """runtests.jl:501"""
    ib = 1e-3 if typ == "npn" else -1e-3
    return np.vstack([np.linspace(0, ib, N),
                      np.concatenate([np.linspace(1, -1, N // 2), np.linspace(-1, 1, N // 2)])])


def bjt_expected(out, ile=0, ilc=0, ηcl=None, ηel=None, vaf=math.inf, var=math.inf,
                 ikf=math.inf, ikr=math.inf):
    """closed form of runtests.jl:535-543; returns (ie, ic) for measured (ve, vc)"""
    isc, ise, ηc, ηe, βf, βr = (BJT_BASE[k] for k in ("isc", "ise", "ηc", "ηe", "βf", "βr"))
    ηcl = ηc if ηcl is None else ηcl
    ηel = ηe if ηel is None else ηel
    ve, vc = out[0], out[1]
    i_f = βf / (1 + βf) * ise * (np.exp(ve / (ηe * 25e-3)) - 1)
    i_r = βr / (1 + βr) * isc * (np.exp(vc / (ηc * 25e-3)) - 1)
    icc = (2 * (1 - ve / var - vc / vaf)) / (1 + np.sqrt(1 + 4 * (i_f / ikf + i_r / ikr))) * (i_f - i_r)
    ibe = 1 / βf * i_f + ile * (np.exp(ve / (ηel * 25e-3)) - 1)
    ibc = 1 / βr * i_r + ilc * (np.exp(vc / (ηcl * 25e-3)) - 1)
    return icc + ibe, -icc + ibc


def mosfet_circuit(typ, **kw):
    """runtests.jl:592-597"""
    return circuit([
        ("vgs", voltagesource(), {"-": "gnd"}),
        ("vds", voltagesource(), {"-": "gnd"}),
        ("J", mosfet(typ, **kw), {"gate": ("vgs", "+"), "drain": ("vds", "+")}),
        ("out", currentprobe(), {"+": ("J", "source"), "-": "gnd"}),
    ])


def opamp_shelving(Amax, GBP):
    """runtests.jl:629-636"""
    return circuit([
        ("input", voltagesource(), {"-": "gnd"}),
        ("op", opamp(maxgain=Amax, gain_bw_prod=GBP), {"in+": ("input", "+"), "out-": "gnd"}),
        ("r1", resistor(109e3), {"1": ("op", "out+"), "2": ("op", "in-")}),
        ("r2", resistor(1e3), {"1": ("op", "in-")}),
        ("c", capacitor(22e-9), {"1": ("r2", "2"), "2": "gnd"}),
        ("output", voltageprobe(), {"+": ("op", "out+"), "-": "gnd"}),
    ])


def opamp_tanh():
    """runtests.jl:652-656"""
    return circuit([
        ("input", voltagesource(), {"-": "gnd"}),
        ("op", opamp("macak", 100, -3, 4), {"in+": ("input", "+"), "in-": [("op", "out-"), "gnd"]}),
        ("output", voltageprobe(), {"+": ("op", "out+"), "-": "gnd"}),
    ])


def source_probe_circuits():
    """runtests.jl:386-429: sources and probes with internal resistance / conductance; entries
    (circuit, input column, expected output)"""
    R = 100000
    g = Fraction(1, R)
    mk = lambda src, probe: circuit([("src", src, {}), ("probe", probe, {"+": ("src", "+"), "-": ("src", "-")})])
    return [(mk(currentsource(100e-3, gp=g), voltageprobe()), [], R * 100e-3),
            (mk(currentsource(gp=g), voltageprobe()), [100e-3], R * 100e-3),
            (mk(currentsource(100e-3), voltageprobe(gp=g)), [], R * 100e-3),
            (mk(voltagesource(10, rs=R), currentprobe()), [], 10 / R),
            (mk(voltagesource(rs=R), currentprobe()), [10.0], 10 / R),
            (mk(voltagesource(10), currentprobe(rs=R)), [], 10 / R)]


def bjt_internal_resistances(typ):
    """runtests.jl:547-587: a BJT with external base/collector/emitter resistors next to one with the same internal
    resistances, same bias; the four probes of each must agree"""
    ib, vce = (1e-3, 1) if typ == "npn" else (-1e-3, -1)
    rb, re, rc = 100, 10, 20
    bj = dict(BJT_BASE)
    c = circuit([
        ("t1", bjt(typ, **bj), {}), ("rbref", resistor(rb), {}), ("rcref", resistor(rc), {}), ("reref", resistor(re), {}),
        ("isrc1", currentsource(ib), {}), ("vscr1", voltagesource(vce), {}),
        ("veprobe1", voltageprobe(), {}), ("vcprobe1", voltageprobe(), {}),
        ("ieprobe1", currentprobe(), {}), ("icprobe1", currentprobe(), {})])
    c.connect(("t1", "base"), ("rbref", "1"))
    c.connect(("rbref", "2"), ("isrc1", "+"), ("veprobe1", "+"), ("vcprobe1", "+"))
    c.connect(("t1", "collector"), ("rcref", "1"))
    c.connect(("rcref", "2"), ("icprobe1", "+"))
    c.connect(("vcprobe1", "-"), ("icprobe1", "-"), ("vscr1", "+"))
    c.connect(("t1", "emitter"), ("reref", "1"))
    c.connect(("reref", "2"), ("ieprobe1", "+"))
    c.connect(("veprobe1", "-"), ("ieprobe1", "-"), ("vscr1", "-"), ("isrc1", "-"))
    for name, el in (("t2", bjt(typ, rb=rb, re=re, rc=rc, **bj)), ("isrc2", currentsource(ib)), ("vscr2", voltagesource(vce)),
                     ("veprobe2", voltageprobe()), ("vcprobe2", voltageprobe()),
                     ("ieprobe2", currentprobe()), ("icprobe2", currentprobe())):
        c.add(name, el)
    c.connect(("t2", "base"), ("isrc2", "+"), ("veprobe2", "+"), ("vcprobe2", "+"))
    c.connect(("t2", "collector"), ("icprobe2", "+"))
    c.connect(("vcprobe2", "-"), ("icprobe2", "-"), ("vscr2", "+"))
    c.connect(("t2", "emitter"), ("ieprobe2", "+"))
    c.connect(("veprobe2", "-"), ("ieprobe2", "-"), ("vscr2", "-"), ("isrc2", "-"))
    return c


def ja_inductor():
    """runtests.jl:433-440: Jiles-Atherton inductor next to its approximate linear equivalent (174 mH)"""
    from acme_jl_b200 import inductor
    return circuit([("Jin", voltagesource(), {}),
                    ("Jout1", currentprobe(), {"+": ("Jin", "+")}),
                    ("Jout2", currentprobe(), {"+": ("Jin", "+")}),
                    ("L_JA", inductor(ja=True), {"1": ("Jout1", "-"), "2": ("Jin", "-")}),
                    ("L_lin", inductor(174e-3), {"1": ("Jout2", "-"), "2": ("Jin", "-")})])


def ja_transformer():
    """runtests.jl:459-468: Jiles-Atherton transformer next to its approximate small-signal equivalent"""
    from acme_jl_b200 import transformer
    return circuit([("Jin", voltagesource(), {}),
                    ("R1", resistor(10), {"1": ("Jin", "+")}),
                    ("R2", resistor(10), {"1": ("Jin", "+")}),
                    ("T_JA", transformer(ja=True, ns=[10, 100]), {"1": ("R1", "2"), "2": ("Jin", "-")}),
                    ("T_lin", transformer(330e-6, 33e-3), {"primary1": ("R2", "2"), "primary2": ("Jin", "-")}),
                    ("Jout1", voltageprobe(gp=1e-3), {"+": ("T_JA", "3"), "-": ("T_JA", "4")}),
                    ("Jout2", voltageprobe(gp=1e-3), {"+": ("T_lin", "secondary1"), "-": ("T_lin", "secondary2")})])


def julia_isapprox(a, b, rtol):
    """isapprox for arrays: norm(a - b) <= rtol * max(norm(a), norm(b))"""
    a, b = np.asarray(a, float), np.asarray(b, float)
    return bool(np.linalg.norm(a - b) <= rtol * max(np.linalg.norm(a), np.linalg.norm(b)))


def test_quad_model(p0=0.0, z0=1.0):
    """The scalar equation z^2 - 1 + p = 0 of runtests.jl:207-219 as a one-sub
    model: q = [z; p], u = p, y = z."""
    from acme_jl_b200.elements import NLElem
    from acme_jl_b200.model import DiscreteModel, SubProblem
    sub = SubProblem(nn=1, nq=2, np_=1,
                     dq=np.zeros((1, 0)), eq=np.ones((1, 1)), fqprev=np.zeros((1, 1)),
                     pexp=np.array([[0.0], [1.0]]), q0=np.zeros(2), fq=np.array([[1.0], [0.0]]),
                     init_z=np.array([z0]), elems=[(NLElem(100, (), 2, 1), 0)])
    return DiscreteModel.from_matrices(a=np.zeros((0, 0)), b=np.zeros((0, 1)), c=np.zeros((0, 1)), x0=np.zeros(0),
                                       dy=np.zeros((1, 0)), ey=np.zeros((1, 1)), fy=np.ones((1, 1)), y0=np.zeros(1),
                                       subs=[sub], solver="HomotopySolver{SimpleSolver}")


def golden():
    """golden vectors the reference holds for the run! path (tests/golden/make_golden.py extracts them
    from the reference's doctests and test suite; the JSON is committed because the GPU box has no
    /root/reference)"""
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.json")))


def assert_printed_equal(got, want, digits=6):
    """compare like the doctest does: the value printed with `digits` significant digits"""
    assert f"{got:.{digits}g}" == f"{want:.{digits}g}", (got, want)
