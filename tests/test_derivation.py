"""Host-side derivation vs the reference's exact-rational unit tests
(/root/reference/test/runtests.jl:12-21, 221-265) and the np/nn pins of the
example circuits (runtests.jl:283-290, 699, 724, 734, 744, 757-759, 768, 777, 788-791)."""
from fractions import Fraction

import numpy as np
import pytest

import acme_jl_b200 as A
from acme_jl_b200 import examples as ex
from acme_jl_b200.model import _dot, gensolve, rank_factorize, reduce_pdims
from acme_jl_b200.elements import rmat

import cases


def R(rows):
    return rmat(rows)


def test_topomat():
    """runtests.jl:12-21"""
    tv, ti = A.topomat([[1, -1, 1], [-1, 1, -1]])
    prod = np.array(tv) @ np.array(ti).T
    assert prod.shape == (2, 1) and not prod.any()
    assert A.topomat([[0], [0]]) == ([[1]], [])
    tv, ti = A.topomat([[1], [-1]])
    assert tv == [] and ti == [[1]]


def test_gensolve_rank_factorize():
    """runtests.jl:221-228"""
    a = R([[1, 1, 1], [1, 1, 2], [1, 2, 1], [1, 2, 2], [2, 1, 1], [2, 1, 2]])
    b = R([[1, 2, 3, 4, 5, 6], [6, 5, 4, 3, 2, 1], [1, 0, 1, 0, 1, 0]])
    nullspace = gensolve(a.T.copy(), R(np.zeros((3, 0)).tolist()).reshape(3, 0))[1]
    assert nullspace.shape == (6, 3)
    assert all(v == 0 for v in _dot(nullspace.T.copy(), a).flat)
    ab = _dot(a, b)
    c, f = rank_factorize(ab)
    assert c.shape[1] == 3
    assert (_dot(c, f) == ab).all()


@pytest.mark.parametrize("zx_zero", [True, False])
@pytest.mark.parametrize("zu_zero", [True, False])
def test_reduce_pdims(zx_zero, zu_zero):
    """runtests.jl:230-265"""
    a = R([[-1, -1, -4, -3, 0, -1], [2, -1, -5, 3, -4, 0], [-2, 2, -5, -2, 5, 1],
           [-5, 4, -3, 0, 5, -5], [4, 3, 0, -1, 0, 2], [0, -3, -4, -4, -3, 4]])
    b = R([1, 2, 3, -2, -1, 0])
    c = R([[4, 2, -1], [-1, -3, 0], [-3, 5, 3], [0, 0, 0], [-4, -1, -1], [-1, -1, 5]])
    dy = R([[1, 2, 3, -2, -1, 0]])
    ey = R([[5]])
    fy = R([[-2, -1, 3]])
    p = R([[1, 1, 1], [1, 1, 2], [1, 2, 1], [1, 2, 2], [2, 1, 1], [2, 1, 2]])
    dq = R([[1, 2, 3, 4, 5, 6], [6, 5, 4, 3, 2, 1], [1, 0, 1, 0, 1, 0]])
    eq = R([1, 2, 3])
    fq = R([[1, 0, 0], [10, 0, 0], [0, 1, 0], [0, 10, 0], [0, 0, 1], [0, 0, 10]])
    zxin = R(np.zeros((3, 6), dtype=int).tolist()) if zx_zero else R([[1, 2, 0, 0, 2, 1], [0, 1, 2, 2, 0, 1], [0, 0, 1, 0, 1, 1]])
    zuin = R([0, 0, 0]) if zu_zero else R([1, 2, -1])
    dq_full = _dot(p, dq) + _dot(fq, zxin)
    eq_full = _dot(p, eq) + _dot(fq, zuin)
    mats = dict(a=a.copy(), b=b.copy(), c=c, dy=dy.copy(), ey=ey.copy(), fy=fy,
                dq_fulls=[dq_full.copy()], eq_fulls=[eq_full.copy()], fqprev_fulls=[eq_full.copy()], fqs=[fq])
    mats = reduce_pdims(mats)
    assert mats["pexps"][0].shape[1] == 3
    assert (_dot(mats["pexps"][0], mats["dqs"][0]) == mats["dq_fulls"][0]).all()
    assert (_dot(mats["pexps"][0], mats["eqs"][0]) == mats["eq_fulls"][0]).all()
    fqT = fq.T.copy()
    pinv = gensolve(_dot(fqT, fq), fqT)[0]
    zx = _dot(pinv, dq_full - mats["dq_fulls"][0])
    zu = _dot(pinv, eq_full - mats["eq_fulls"][0])
    assert (mats["a"] == a - _dot(c, zx)).all()
    assert (mats["b"] == b - _dot(c, zu)).all()
    assert (mats["dy"] == dy - _dot(fy, zx)).all()
    assert (mats["ey"] == ey - _dot(fy, zu)).all()


def test_element_equality():
    """runtests.jl:43-51"""
    assert A.resistor(1e3) == A.resistor(1e3)
    assert A.resistor(1e3) != A.resistor(2.2e3)
    assert A.resistor(1) != A.voltagesource(1)
    assert A.bjt("npn") == A.bjt("npn")
    assert A.bjt("npn") != A.bjt("pnp")


def np_list(m):
    return [s.np_ for s in m.subs]


def test_example_dims():
    """K10: the np/nn pins of the example circuits"""
    m = ex.diodeclipper()
    assert (m.nx, m.nu, m.ny) == (1, 1, 1) and np_list(m) == [1]          # runtests.jl:699
    m = ex.sallenkey()
    assert (m.nx, m.nu, m.ny, len(m.subs)) == (2, 1, 1, 0)
    assert np_list(ex.birdie(vol=0.8)) == [2]                              # :724
    assert np_list(ex.birdie()) == [3]                                     # :734
    assert np_list(ex.superover(1.0, 1.0, 1.0)) == [5]                     # :744
    m = ex.superover()
    assert np_list(m) == [11] and m.nu == 4                                # :777


def test_simplified_superover_dims():
    m = A.DiscreteModel(ex.superover_circuit(1.0, 1.0, 1.0, vb_source=True), 1 / 44100)
    assert np_list(m) == [2, 1, 2]                                         # :757-759
    m = A.DiscreteModel(ex.superover_circuit(1.0, 1.0, 1.0, vb_source=True), 1 / 44100,
                        decompose_nonlinearity=False)
    assert np_list(m) == [5]                                               # :768
    m = A.DiscreteModel(ex.superover_circuit(vb_source=True), 1 / 44100)
    assert np_list(m) == [2, 2, 2, 4]                                      # :788-791


def test_three_diode_decomposition_dims():
    """runtests.jl:280-291"""
    c = cases.three_diodes()
    m = A.DiscreteModel(c, 1, decompose_nonlinearity=False)
    assert [s.nn for s in m.subs] == [3]
    m = A.DiscreteModel(c, 1)
    assert [s.nn for s in m.subs] == [1, 2]


def test_indeterminate_warnings():
    """runtests.jl:153-168"""
    c = A.circuit([("r", A.resistor(0), {}),
                   ("probe", A.currentprobe(), {"+": ("r", "1"), "-": ("r", "2")})])
    with pytest.warns(UserWarning, match="Model output depends on indeterminate quantity"):
        A.DiscreteModel(c, 1)
    c = A.circuit([("u", A.opamp(), {"in+": ("u", "in-")}),
                   ("c", A.capacitor(1e-6), {"1": ("u", "out-"), "2": ("u", "out+")})])
    with pytest.warns(UserWarning, match="State update depends on indeterminate quantity"):
        A.DiscreteModel(c, 1)


def test_steadystate_ladder():
    """docs/src/ug.md:148-174"""
    m = A.DiscreteModel(cases.rc_ladder(), 1 / 44100)
    assert np.allclose(m.steadystate_(), np.zeros(20))


def test_K10_pins_from_golden_fixture():
    """every np/nn pin of test/runtests.jl, as extracted into tests/golden/reference_vectors.json
    (line number -> value), against the host-side derivation"""
    pins = {(p["line"], p["what"]): p["value"] for p in cases.golden()["K10_dimension_pins"]}
    simp = A.DiscreteModel(ex.superover_circuit(1.0, 1.0, 1.0, vb_source=True), 1 / 44100)
    simp_in = A.DiscreteModel(ex.superover_circuit(vb_source=True), 1 / 44100)
    three = A.DiscreteModel(cases.three_diodes(), 1)
    three_nd = A.DiscreteModel(cases.three_diodes(), 1, decompose_nonlinearity=False)
    ours = {
        (699, "np"): ex.diodeclipper().subs[0].np_, (724, "np"): ex.birdie(vol=0.8).subs[0].np_,
        (734, "np"): ex.birdie().subs[0].np_, (744, "np"): ex.superover(1.0, 1.0, 1.0).subs[0].np_,
        (757, "np"): simp.subs[0].np_, (758, "np"): simp.subs[1].np_, (759, "np"): simp.subs[2].np_,
        (768, "np"): A.DiscreteModel(ex.superover_circuit(1.0, 1.0, 1.0, vb_source=True), 1 / 44100,
                                     decompose_nonlinearity=False).subs[0].np_,
        (777, "np"): ex.superover().subs[0].np_,
        (788, "np"): simp_in.subs[0].np_, (789, "np"): simp_in.subs[1].np_, (790, "np"): simp_in.subs[2].np_,
        (791, "np"): simp_in.subs[3].np_,
        (283, "nn"): three_nd.subs[0].nn, (289, "nn"): three.subs[0].nn, (290, "nn"): three.subs[1].nn,
    }
    checked = 0
    for key, val in ours.items():
        assert key in pins, key
        assert pins[key] == val, (key, pins[key], val)
        checked += 1
    assert checked == len(ours) >= 16


def test_compiled_kernel_shapes_match_derived_models():
    """The compile-time shapes the device library instantiates (csrc/tpi.cu, csrc/rows.cu) are the
    shapes the derivation produces for the BASELINE circuits (SURVEY.md section 8 size table):
    (nx, nu, ny, nn, nq, np, element kinds in CircuitNLFunc order)."""
    D, B, P = 1, 2, 3   # ACMEB200_ELEM_DIODE / BJT / POT (include/acmeb200.h)
    nj = {D: 1, B: 4, P: 4}  # variable Jacobian entries per element (csrc/elements.cuh)

    def shape(m):
        s = m.subs[0]
        kinds = [e.kind for e, _ in s.elems]
        return (m.nx, m.nu, m.ny, s.nn, s.nq, s.np_, kinds, sum(nj[k] for k in kinds))

    assert shape(ex.diodeclipper()) == (1, 1, 1, 2, 4, 1, [D, D], 2)                      # tpi<diodeclipper>
    assert shape(ex.birdie(vol=0.8)) == (3, 1, 1, 2, 4, 2, [B], 4)                          # tpi<birdie ... [bjt]>
    assert shape(ex.birdie()) == (3, 2, 1, 4, 9, 3, [B, P], 8)                              # tpi<birdie ... [bjt,pot]>
    m = ex.sallenkey(fs=96000)
    assert (m.nx, m.nu, m.ny, len(m.subs)) == (2, 1, 1, 0)                                  # tpi<linear>
    so = shape(ex.superover())
    assert so[:6] == (11, 4, 1, 13, 29, 11) and len(so[6]) == 8 and so[7] == 23             # rows: RowsSuperover
    assert sorted(so[6]) == sorted([B, D, D, D, P, P, P, B])
    sb = shape(ex.superover(0.3, 0.7, 1.0))
    assert sb[:6] == (11, 1, 1, 7, 14, 5) and sb[6] == [B, D, D, D, B] and sb[7] == 11       # rows: RowsSuperoverBaked
