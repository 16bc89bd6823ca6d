#!/usr/bin/env python
"""Extracts the golden vectors the reference itself holds for the run! path into
tests/golden/reference_vectors.json.

The reference (HSU-ANT/ACME.jl, mounted read-only at /root/reference in the build container)
is pure Julia and cannot be executed here or on the GPU box, so the golden data are the OUTPUTS
THE REFERENCE PRINTS IN ITS OWN DOCTESTS (run by its CI through Documenter) and the known
answers of its test suite:

  G1  docs/src/gettingstarted.md:106-113   diode clipper, 1 s of a 1 kHz sine at 44.1 kHz
  G2  docs/src/ug.md:107-114               20-stage RC ladder, first 100 samples of the impulse response
  K1  test/runtests.jl:23-41               LinearSolver known answer (3x3)
  K4  test/runtests.jl:70-86               resistor-diode DC operating point
  K10 test/runtests.jl (np / nn pins)      dimensions of the four example circuits

Run in the build container:  python tests/golden/make_golden.py
(the GPU box has no /root/reference: tests only read the committed JSON).
"""
import json
import os
import re

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")


def doctest_matrix(path, call_pattern):
    """the printed matrix row that follows `call_pattern` + '# output' in a jldoctest block"""
    lines = open(path, encoding="utf-8").read().splitlines()
    for i, line in enumerate(lines):
        if re.search(call_pattern, line):
            for j in range(i, i + 12):
                m = re.match(r"^\s*1×(\d+) Matrix\{Float64\}:", lines[j])
                if m:
                    row = lines[j + 1]
                    head, tail = row.split("…")
                    return {"file": os.path.relpath(path, REF), "line": j + 2, "n": int(m.group(1)),
                            "first": [float(x) for x in head.split()], "last": [float(x) for x in tail.split()],
                            "printed_digits": 6}
    raise RuntimeError(f"doctest output not found in {path}")


def main():
    g1 = doctest_matrix(os.path.join(REF, "docs/src/gettingstarted.md"), r"y = run!\(model, sin\.")
    g1["input"] = "sin(2*pi*1000/44100*n), n = 0..44099; diodeclipper example circuit, fs = 44100"
    g2 = doctest_matrix(os.path.join(REF, "docs/src/ug.md"), r"run!\(model, \[1 zeros\(1,99\)\]\)")
    g2["input"] = "unit impulse, 100 samples; 20-stage RC ladder of docs/src/ug.md:40-56"
    tests = open(os.path.join(REF, "test/runtests.jl"), encoding="utf-8").read()
    pins = []
    for m in re.finditer(r"@test ACME\.(np|nn)\(model(?:, (\d+))?\) == (\d+)", tests):
        line = tests.count("\n", 0, m.start()) + 1
        pins.append({"what": m.group(1), "sub": int(m.group(2) or 0), "value": int(m.group(3)), "line": line})
    k4 = re.search(r"25e-3 \* log\(1e-3 / 1e-12 ?\+ ?1\)", tests)
    out = {
        "source": "HSU-ANT/ACME.jl doctests and test/runtests.jl (see make_golden.py)",
        "G1_diodeclipper_doctest": g1,
        "G2_rc_ladder_doctest": g2,
        "K4_resistor_diode": {"file": "test/runtests.jl", "line": tests.count("\n", 0, k4.start()) + 1 if k4 else None,
                              "expression": "25e-3*log(1e-3/1e-12+1)", "is": 1e-12, "r": 1e3, "v": 1.0},
        "K10_dimension_pins": pins,
    }
    json.dump(out, open(OUT, "w"), indent=1)
    print("wrote", OUT)
    print(json.dumps(out)[:600])


if __name__ == "__main__":
    main()
