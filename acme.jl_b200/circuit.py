"""Circuit description (host side, one-off; mirrors /root/reference/src/circuit.jl).

Only what the model derivation needs: ``Circuit`` with ``add/connect/disconnect/
delete`` (circuit.jl:24-206), ``incidence`` (:51-66), ``topomat`` (:208-252) and
the element-table analogue of ``nonlinear_eq_func`` (:68-86).  The ``@circuit``
macro (:317-406) has no Python equivalent; :func:`circuit` takes the same
information as a dict ``{refdes: (element, {pin: target, ...})}`` where a target
is a net name (``"gnd"``) or a ``(refdes, pin)`` tuple.
"""
from __future__ import annotations

from collections import OrderedDict
from fractions import Fraction
from typing import Dict, List, Tuple

import numpy as np

from .elements import Element, NLElem, rzeros


class Circuit:
    def __init__(self):
        self.elements: "OrderedDict[str, Element]" = OrderedDict()
        self.nets: List[List[Tuple[str, str]]] = []
        self.net_names: Dict[str, List[Tuple[str, str]]] = {}
        self._gensym = 0

    # ------------------------------------------------------------ sizes
    def _sum(self, attr):
        return sum(getattr(e, attr) for e in self.elements.values())

    nb = property(lambda s: s._sum("nb"))
    nx = property(lambda s: s._sum("nx"))
    nq = property(lambda s: s._sum("nq"))
    nu = property(lambda s: s._sum("nu"))
    nl = property(lambda s: s._sum("nl"))
    ny = property(lambda s: s._sum("ny"))
    nn = property(lambda s: s._sum("nn"))

    def blockdiag(self, name: str) -> np.ndarray:
        """circuit.jl:37-47"""
        mats = [e.m[name] for e in self.elements.values()]
        out = rzeros(sum(m.shape[0] for m in mats), sum(m.shape[1] for m in mats))
        r = c = 0
        for m in mats:
            out[r:r + m.shape[0], c:c + m.shape[1]] = m
            r += m.shape[0]
            c += m.shape[1]
        return out

    def u0(self) -> np.ndarray:
        """circuit.jl:49"""
        mats = [e.m["u0"] for e in self.elements.values()]
        if not mats:
            return rzeros(0, 1)
        return np.vstack(mats)

    # ------------------------------------------------------------ editing
    def add(self, designator, elem: Element = None):
        """circuit.jl:94-117"""
        if elem is None:
            designator, elem = None, designator
        if designator is None:
            self._gensym += 1
            designator = f"##{self._gensym}"
        designator = str(designator)
        if designator in self.elements:
            self.delete(designator)
        for pin in elem.pins:
            self.nets.append([(designator, pin)])
        self.elements[designator] = elem
        return designator

    def delete(self, designator):
        """circuit.jl:125-130"""
        designator = str(designator)
        for net in self.nets:
            net[:] = [ep for ep in net if ep[0] != designator]
        del self.elements[designator]

    def _netfor(self, p):
        """circuit.jl:141-152"""
        if isinstance(p, tuple):
            key = (str(p[0]), str(p[1]))
            for net in self.nets:
                if key in net:
                    return net
            raise ValueError(f"Unknown pin {p}")
        name = str(p)
        if name not in self.net_names:
            net = []
            self.net_names[name] = net
            self.nets.append(net)
        return self.net_names[name]

    def connect(self, *pins):
        """circuit.jl:175-188"""
        nets = []
        for pin in pins:
            n = self._netfor(pin)
            if not any(n is m for m in nets):
                nets.append(n)
        for net in nets[1:]:
            nets[0].extend(net)
            idx = next(i for i, m in enumerate(self.nets) if m is net)
            del self.nets[idx]
            for name, named in list(self.net_names.items()):
                if named is net:
                    self.net_names[name] = nets[0]

    def disconnect(self, pin):
        """circuit.jl:190-206"""
        key = (str(pin[0]), str(pin[1]))
        net = self._netfor(key)
        net[:] = [p for p in net if p != key]
        self.nets.append([key])

    # ------------------------------------------------------------ topology
    def branch_offset(self, designator) -> int:
        off = 0
        for des, el in self.elements.items():
            if des == designator:
                return off
            off += el.nb
        raise ValueError("Element not found in circuit")

    def incidence(self) -> List[List[int]]:
        """circuit.jl:51-66 (dense int matrix; duplicates add up like ``sparse``)."""
        inc = [[0] * self.nb for _ in self.nets]
        for row, pins in enumerate(self.nets):
            for elemname, pinname in pins:
                off = self.branch_offset(elemname)
                for branch, polarity in self.elements[elemname].pins[pinname]:
                    inc[row][off + branch - 1] += polarity
        return inc

    def topomat(self):
        return topomat(self.incidence())

    def nl_table(self, elem_idxs=None):
        """Element-table analogue of ``nonlinear_eq_func(c, elem_idxs)``
        (circuit.jl:68-86): entries ``(NLElem, q_offset)`` in element order,
        q-offsets cumulative in ``nq(elem)``, elements without nn and nq skipped."""
        elems = list(self.elements.values())
        if elem_idxs is not None:
            elems = [elems[i] for i in elem_idxs]
        table = []
        col = 0
        for el in elems:
            if el.nn == 0 and el.nq == 0:
                continue
            sub = 0
            for nle in el.nl_elems:
                table.append((nle, col + sub))
                sub += nle.nq
            col += el.nq
        return table


def topomat(incidence: List[List[int]]):
    """circuit.jl:208-249.  Returns ``(tv, ti)`` as lists of int rows."""
    inc = [list(r) for r in incidence]
    nrows = len(inc)
    ncols = len(inc[0]) if nrows else 0
    if nrows == 0:
        # no nets: every branch is a self-loop; handled by the general code below
        pass
    assert all(abs(v) == 1 for r in inc for v in r if v != 0)
    for c in range(ncols):
        assert sum(inc[r][c] for r in range(nrows)) == 0
    t = [False] * ncols
    row = 0
    for col in range(ncols):
        rows = [r for r in range(row, nrows) if inc[r][col] != 0]
        assert len(rows) <= 2
        if not rows:
            continue
        t[col] = True
        if rows[0] != row:
            inc[rows[0]], inc[row] = inc[row], inc[rows[0]]
        if len(rows) == 2:
            assert inc[row][col] + inc[rows[1]][col] == 0
            inc[rows[1]] = [a + b for a, b in zip(inc[rows[1]], inc[row])]
        if inc[row][col] < 0:
            inc[row] = [-v for v in inc[row]]
        for r in range(row):
            if inc[r][col] == 1:
                inc[r] = [a - b for a, b in zip(inc[r], inc[row])]
            elif inc[r][col] == -1:
                inc[r] = [a + b for a, b in zip(inc[r], inc[row])]
        row += 1
    ti = [inc[r] for r in range(row)]
    tcols = [c for c in range(ncols) if t[c]]
    lcols = [c for c in range(ncols) if not t[c]]
    tv = [[0] * ncols for _ in lcols]
    for k, lc in enumerate(lcols):
        for r, tc in enumerate(tcols):
            tv[k][tc] = -ti[r][lc]
        tv[k][lc] = 1
    return tv, ti


def circuit(spec) -> Circuit:
    """Python stand-in for ``@circuit`` (circuit.jl:317-406).

    ``spec`` is an ordered mapping ``refdes -> (element, connections)`` or a list
    of ``(refdes, element, connections)``; ``connections`` maps a pin of this
    element to a net name, a ``(refdes, pin)`` tuple, or a list of those.
    """
    c = Circuit()
    items = spec.items() if isinstance(spec, dict) else [(s[0], (s[1], s[2] if len(s) > 2 else {})) for s in spec]
    for refdes, val in items:
        elem, conns = val if isinstance(val, tuple) else (val, {})
        c.add(refdes, elem)
        for pin, targets in conns.items():
            if not isinstance(targets, list):
                targets = [targets]
            c.connect((refdes, pin), *targets)
    return c
