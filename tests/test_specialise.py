"""Shapes the library does not carry (missing #5): `specialise(model)` generates the one-line
instantiation of the thread-per-instance kernel for the model's shape, builds it with nvcc into a plugin library and
registers it (acmeb200_register_tpi); the model then runs on a specialised kernel instead of the run-time-dimension one."""
import ctypes as C
import os
from fractions import Fraction

import numpy as np
import pytest

import acme_jl_b200 as A
from acme_jl_b200 import BatchRunner, _specialise as sp
from acme_jl_b200.circuit import circuit
from acme_jl_b200.elements import capacitor, diode, mosfet, resistor, voltageprobe, voltagesource
from oracle.oracle import OracleModel

import cases

HC = "HomotopySolver{CachingSolver{SimpleSolver}}"


def halfwave(is_=1e-15):
    """a series resistor into a capacitor with ONE diode across it: the clipper's topology, another element sequence"""
    return circuit([("j_in", voltagesource(), {"-": "gnd"}), ("r1", resistor(1e3), {"1": ("j_in", "+")}),
                    ("c1", capacitor(47e-9), {"1": ("r1", "2"), "2": "gnd"}), ("d1", diode(is_=is_), {"+": ("r1", "2"), "-": "gnd"}),
                    ("j_out", voltageprobe(), {"-": "gnd", "+": ("r1", "2")})])


def mosfet_stage():
    """common-source stage with a load capacitor (the element of test/runtests.jl:590-624 in a dynamic circuit)"""
    return circuit([("vdd", voltagesource(9), {"-": "gnd"}), ("j_in", voltagesource(), {"-": "gnd"}),
                    ("rd", resistor(4.7e3), {"1": ("vdd", "+")}),
                    ("m1", mosfet("n", vt=1.0, α=2e-3), {"gate": ("j_in", "+"), "drain": ("rd", "2"), "source": "gnd"}),
                    ("cl", capacitor(10e-9), {"1": ("rd", "2"), "2": "gnd"}),
                    ("j_out", voltageprobe(), {"+": ("rd", "2"), "-": "gnd"})])


def test_generator_builds_a_plugin_the_library_accepts():
    m = A.DiscreteModel(halfwave(), Fraction(1, 44100))
    shape = sp.shape_of(m)
    assert shape == (1, 1, 1, 1, (1,)) and sp.shape_key(shape) == "nx1_nu1_ny1_np1_e1"
    assert "TpiCfg<1, 1, 1, 1, Diode>" in sp.plugin_source(shape)
    so = sp.build_plugin(shape)
    plug = C.CDLL(so)
    plug.acmeb200_shape_entry.restype = C.c_void_p
    assert plug.acmeb200_shape_entry()
    sp.register_plugin(so)          # needs no device: the registry is host state
    sp.register_plugin(so)          # idempotent
    big = A.DiscreteModel(A.examples.superover_circuit(), Fraction(1, 44100))
    with pytest.raises(ValueError, match="too large for one thread"):
        sp.shape_of(big)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["halfwave", "mosfet"])
def test_specialised_kernel_for_a_shape_outside_the_registry(name):
    m = A.DiscreteModel(halfwave() if name == "halfwave" else mosfet_stage(), Fraction(1, 44100))
    B, N = 96, 3000
    u = np.asfortranarray(np.repeat(cases.sine(N)[:, :, None], B, axis=2) * np.linspace(0.5, 4.0, B)[None, None, :])
    if name == "mosfet":
        u = u + 1.5   # bias the gate around the threshold
    yref = OracleModel(m, B, solver=HC).run(u, threads=0)
    r = BatchRunner(m, B, solver=HC, kernel="generic")
    yg = r.run(u)
    r.close()
    A.specialise(m)
    r = BatchRunner(m, B, solver=HC)
    assert r.kernel_name.startswith("tpi<specialised"), r.kernel_name
    y = r.run(u)
    st = r.stats()
    r.close()
    scale = np.maximum(np.abs(yref), 1e-3 * np.abs(yref).max())
    assert (np.abs(y - yref) <= 1e-6 * scale).all() and (np.abs(yg - yref) <= 1e-6 * scale).all()
    assert st["samples"] == B * N and st["not_converged"] == 0
