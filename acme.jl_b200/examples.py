"""The four example circuits that BASELINE.json's configs name
(/root/reference/examples/{diodeclipper,sallenkey,birdie,superover}.jl),
written with :func:`circuit` instead of the ``@circuit`` macro.  Element order,
values and connections follow the example files line by line; keyword overrides
(``is1``, ``r``, ``c1`` ...) exist so the benchmark sweeps can rebuild the same
topology with swept values.
"""
from __future__ import annotations

from fractions import Fraction

from .circuit import Circuit, circuit
from .elements import (bjt, capacitor, diode, opamp, potentiometer, resistor,
                       voltageprobe, voltagesource)
from .model import DiscreteModel


def diodeclipper_circuit(is1=1e-15, is2=1.8e-15, η1=1, η2=1, r=1e3, c=47e-9) -> Circuit:
    """examples/diodeclipper.jl:6-15"""
    return circuit([
        ("j_in", voltagesource(), {"-": "gnd"}),
        ("r1", resistor(r), {"1": ("j_in", "+")}),
        ("c1", capacitor(c), {"1": ("r1", "2"), "2": "gnd"}),
        ("d1", diode(is_=is1, η=η1), {"-": "gnd", "+": ("r1", "2")}),
        ("d2", diode(is_=is2, η=η2), {"-": ("r1", "2"), "+": "gnd"}),
        ("j_out", voltageprobe(), {"-": "gnd", "+": ("r1", "2")}),
    ])


def diodeclipper(fs=44100, **kw) -> DiscreteModel:
    return DiscreteModel(diodeclipper_circuit(**kw), Fraction(1, fs))


def sallenkey_circuit(r1=10e3, r2=10e3, c1=10e-9, c2=10e-9) -> Circuit:
    """examples/sallenkey.jl:6-17"""
    return circuit([
        ("j_in", voltagesource(), {"-": "gnd"}),
        ("r1", resistor(r1), {"1": ("j_in", "+")}),
        ("r2", resistor(r2), {"1": ("r1", "2")}),
        ("c1", capacitor(c1), {"1": ("r1", "2")}),
        ("u1", opamp(), {"in+": ("r2", "2"), "in-": [("u1", "out+"), ("c1", "2")], "out-": "gnd"}),
        ("c2", capacitor(c2), {"1": ("u1", "in+"), "2": "gnd"}),
        ("j_out", voltageprobe(), {"-": "gnd", "+": ("u1", "out+")}),
    ])


def sallenkey(fs=44100, **kw) -> DiscreteModel:
    return DiscreteModel(sallenkey_circuit(**kw), Fraction(1, fs))


def birdie_circuit(vol=None) -> Circuit:
    """examples/birdie.jl:13-31"""
    return circuit([
        ("j3", voltagesource(9), {"-": "gnd", "+": "vcc"}),
        ("c5", capacitor(100e-6), {"1": "gnd", "2": "vcc"}),
        ("d1", diode(is_=350e-12, η=1.6), {"-": "vcc", "+": "gnd"}),
        ("j1", voltagesource(), {"-": "gnd"}),
        ("r1", resistor(1e6), {"1": ("j1", "+"), "2": "gnd"}),
        ("c1", capacitor(2.2e-9), {"1": ("j1", "+")}),
        ("r2", resistor(43e3), {"1": ("c1", "2"), "2": "gnd"}),
        ("r3", resistor(430e3), {"1": ("c1", "2"), "2": "vcc"}),
        ("t1", bjt("npn", isc=154.1e-15, ise=64.53e-15, ηc=1.10, ηe=1.06, βf=500, βr=12),
         {"base": ("c1", "2")}),
        ("r4", resistor(390), {"1": ("t1", "emitter"), "2": "gnd"}),
        ("r5", resistor(10e3), {"1": ("t1", "collector"), "2": "vcc"}),
        ("c3", capacitor(2.2e-9), {"1": ("t1", "collector")}),
        ("p1", potentiometer(100e3, vol), {"1": "gnd", "3": ("c3", "2")}),
        ("j2", voltageprobe(), {"-": "gnd", "+": ("p1", "2")}),
    ])


def birdie(vol=None, fs=44100) -> DiscreteModel:
    return DiscreteModel(birdie_circuit(vol), Fraction(1, fs))


def superover_circuit(drive=None, tone=None, level=None, sym=False, vb_source=False) -> Circuit:
    """examples/superover.jl:10-75"""
    spec = [
        # power supply
        ("j3", voltagesource(9), {"+": "vcc", "-": "gnd"}),
        ("d4", diode(is_=12e-9, η=2), {"-": "vcc", "+": "gnd"}),
        ("c11", capacitor(100e-6), {"1": "vcc", "2": "gnd"}),
        ("r17", resistor(33e3), {"1": "vcc", "2": "vb"}),
        ("r18", resistor(33e3), {"1": "vb", "2": "gnd"}),
        ("c12", capacitor(47e-6), {"1": "vb", "2": "gnd"}),
        # input stage
        ("j1", voltagesource(), {"-": "gnd"}),
        ("r1", resistor(2.2e6), {"1": ("j1", "+"), "2": "gnd"}),
        ("c1", capacitor(47e-9), {"1": ("j1", "+")}),
        ("r2", resistor(10e3), {"1": ("c1", "2")}),
        ("r3", resistor(470e3), {"1": ("r2", "2"), "2": "vb"}),
        ("q1", bjt("npn", is_=80e-15, βf=500, βr=10), {"base": ("r2", "2"), "collector": "vcc"}),
        ("r4", resistor(10e3), {"1": ("q1", "emitter"), "2": "gnd"}),
        ("c2", capacitor(18e-9), {"1": ("q1", "emitter")}),
        ("r5", resistor(100e3), {"1": ("c2", "2"), "2": "vb"}),
        # distortion stage
        ("ic1a", opamp(), {"in+": ("c2", "2"), "out-": "gnd"}),
        ("d1", diode(is_=4e-9, η=2), {"-": ("ic1a", "out+"), "+": ("ic1a", "in-")}),
        ("d2", diode(is_=3e-9, η=2), {"-": ("ic1a", "in-")}),
        ("d3", diode(is_=5e-9, η=2), {"+": ("ic1a", "out+"), "-": ("d2", "+")}),
        ("p1", potentiometer(1e6, drive), {"2": [("p1", "3"), ("ic1a", "out+")]}),
        ("r6", resistor(33e3), {"1": ("ic1a", "in-"), "2": ("p1", "1")}),
        ("c4", capacitor(47e-9), {"1": ("ic1a", "in-")}),
        ("r7", resistor(4.7e3), {"1": ("c4", "2"), "2": "vb"}),
        # tone control stage
        ("r8", resistor(10e3), {"1": ("ic1a", "out+")}),
        ("ic1b", opamp(), {"in+": ("r8", "2"), "out-": "gnd"}),
        ("c5", capacitor(18e-9), {"1": ("ic1b", "in+"), "2": "gnd"}),
        ("r10", resistor(10e3), {"1": ("ic1b", "out+"), "2": ("ic1b", "in-")}),
        ("c7", capacitor(10e-9), {"1": ("ic1b", "out+"), "2": ("ic1b", "in-")}),
        ("p2", potentiometer(20e3, tone), {"1": ("ic1b", "in+"), "3": ("ic1b", "in-")}),
        ("c6", capacitor(27e-9), {"1": ("p2", "2")}),
        ("r11", resistor(470), {"1": ("c6", "2"), "2": "gnd"}),
        # output stage
        ("c8", capacitor(1e-3), {"1": ("ic1b", "out+")}),
        ("r12", resistor(4.7e3), {"1": ("c8", "2")}),
        ("p3", potentiometer(10e3, level), {"1": "vb", "3": ("r12", "2")}),
        ("r20", resistor(22e3), {"1": ("p3", "2")}),
        ("c9", capacitor(47e-9), {"1": ("r20", "2")}),
        ("r13", resistor(1e6), {"1": ("c9", "2"), "2": "vb"}),
        ("q2", bjt("npn", is_=80e-15, βf=500, βr=10), {"base": ("c9", "2"), "collector": "vcc"}),
        ("r14", resistor(10e3), {"1": ("q2", "emitter"), "2": "gnd"}),
        ("r15", resistor(1e3), {"1": ("q2", "emitter")}),
        ("c10", capacitor(1e-6), {"1": ("r15", "2")}),
        ("r16", resistor(100e3), {"1": ("c10", "2"), "2": "gnd"}),
        ("j2", voltageprobe(), {"+": ("c10", "2"), "-": "gnd"}),
    ]
    circ = circuit(spec)
    if sym:
        circ.connect(("d3", "-"), ("d3", "+"))
    if vb_source:
        # the "simplified" variant of test/runtests.jl:751-756: ideal 4.5 V bias source
        circ.add("vbsrc", voltagesource(4.5))
        circ.connect(("vbsrc", "+"), "vb")
        circ.connect(("vbsrc", "-"), "gnd")
    return circ


def superover(drive=None, tone=None, level=None, sym=False, fs=44100, vb_source=False, **kw) -> DiscreteModel:
    return DiscreteModel(superover_circuit(drive, tone, level, sym, vb_source), Fraction(1, fs), **kw)
