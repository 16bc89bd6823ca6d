"""acme.jl_b200 -- B200-native batched DK-method circuit simulation.

Host-side mirror of the reference's ``DiscreteModel`` / ``ModelRunner`` / ``run!``
interface (/root/reference/src/ACME.jl:118-148, 567-664) over a C-ABI
(``include/acmeb200.h``) into hand-written sm_100a CUDA kernels (``csrc/``).

The directory is named after the reference repo (``acme.jl_b200``); because of
the dot it is imported through the root-level shim module ``acme_jl_b200``.
"""
from .elements import (Element, NLElem, bjt, capacitor, currentprobe, currentsource,
                       diode, inductor, mosfet, opamp, potentiometer, resistor,
                       transformer, voltageprobe, voltagesource)
from .circuit import Circuit, circuit, topomat
from .model import DiscreteModel, SubProblem, gensolve, rank_factorize
from . import examples
from .runner import BatchRunner, ModelRunner, MultiGpuRunner, run_, DimensionMismatch
from .sweep import derive_sweep
from .kdtree import KDTree, frozen_cache
from ._specialise import specialise

__all__ = [
    "Element", "NLElem", "Circuit", "circuit", "topomat", "DiscreteModel", "SubProblem",
    "gensolve", "rank_factorize", "examples",
    "BatchRunner", "ModelRunner", "MultiGpuRunner", "run_", "DimensionMismatch", "derive_sweep", "KDTree", "frozen_cache", "specialise",
    "resistor", "potentiometer", "capacitor", "inductor", "transformer",
    "voltagesource", "currentsource", "voltageprobe", "currentprobe",
    "diode", "bjt", "mosfet", "opamp",
]
