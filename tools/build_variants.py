"""Kernel tuning: build variants of the thread-per-instance translation unit (tpi.cu) with different
CTA sizes / tile lengths / pipeline depths and link each into its own library under tools/libs/
(select one with ACMEB200_LIB=tools/libs/lib_<name>.so; tools/kbench_cfg3.py times configs 2 and 3).
usage: python tools/build_variants.py name:TPB:MINB:T:STAGES:OSTAGES[:DEFINE=VALUE,...] [...]
e.g.  base:64:8:8:2:2 t16:64:8:16:2:1 creg:64:8:8:2:2:ACME_TPI_CACHE_REG=1"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import acme_jl_b200._build as b

b.build()  # objects of the other translation units
libs = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libs")
os.makedirs(libs, exist_ok=True)
procs = []
for spec in sys.argv[1:]:
    name, tpb, mb, T, st, ost, *extra = spec.split(":")
    defines = [f"-D{d}" for d in extra[0].split(",")] if extra else []
    obj = f"{libs}/tpi_{name}.o"
    flags = [f"-DACME_TPI_TPB={tpb}", f"-DACME_TPI_MINB={mb}", f"-DACME_TPI_T={T}", f"-DACME_TPI_STAGES={st}", f"-DACME_TPI_OSTAGES={ost}"] + defines
    cmd = [b.nvcc()] + b.NVCC_FLAGS + flags + ["-Xptxas", "-v", "-c", "-o", obj, os.path.join(b.CSRC, "tpi.cu")]
    wobj = f"{libs}/tpi_wide_{name}.o"  # the 255-register build of the same variant (tpi_wide.cu), compiled alongside
    wp = subprocess.Popen([b.nvcc()] + b.NVCC_FLAGS + flags + ["-c", "-o", wobj, os.path.join(b.CSRC, "tpi_wide.cu")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    procs.append((name, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True), wobj, wp))
for name, obj, p, wobj, wp in procs:
    out = p.communicate()[0]
    wout = wp.communicate()[0]
    if wp.returncode:
        print(name, 'FAILED (wide)\n', wout[-1500:])
        continue
    if p.returncode:
        print(name, "FAILED\n", out[-1500:])
        continue
    lines = out.splitlines()
    for i, l in enumerate(lines):  # registers / spills of the clipper and the linear kernel
        if "Compiling entry function" in l and ("Li1ELi1ELi1ELi1EJNS_5DiodeES2_EEELb0" in l or "Li2ELi1ELi1ELi0EJEEELb1" in l):
            print(name, l.split("'")[1][20:60], "|", lines[i + 2].strip()[:70], "|", lines[i + 3].strip()[:40])
    lib = f"{libs}/lib_{name}.so"
    objs = [obj if s == "tpi.cu" else (wobj if s == "tpi_wide.cu" else b._obj(s)) for s in b.SOURCES]
    r = subprocess.run([b.nvcc(), "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
                        "-Xcompiler", "-fPIC", "-o", lib] + objs, capture_output=True, text=True)
    print(name, "->", lib if r.returncode == 0 else r.stderr[-300:])
    os.remove(obj); os.remove(wobj)
