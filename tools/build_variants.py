"""kernel tuning: build variants of the thread-per-instance translation unit (tpi.cu) with different
CTA sizes / tile lengths and link each into its own library under tools/libs/.
usage: python tools/build_variants.py name:TPB:MINB:T [...]   e.g.  t8:64:8:8 t16:64:8:16"""
import glob, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import acme_jl_b200._build as b

b.build()  # objects of the other translation units
os.makedirs(os.path.join(b.HERE, "..", "tools", "libs"), exist_ok=True)
libs = os.path.abspath(os.path.join(b.HERE, "..", "tools", "libs"))
procs = []
for spec in sys.argv[1:]:
    name, tpb, mb, T = spec.split(":")
    obj = os.path.join(libs, f"tpi_{name}.o")
    cmd = [b.nvcc()] + b.NVCC_FLAGS + [f"-DACME_TPI_TPB={tpb}", f"-DACME_TPI_MINB={mb}", f"-DACME_TPI_T={T}", "-Xptxas", "-v",
                                      "-c", "-o", obj, os.path.join(b.CSRC, "tpi.cu")]
    procs.append((name, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for name, obj, p in procs:
    out = p.communicate()[0]
    if p.returncode:
        print(name, "FAILED\n", out[-2000:]); continue
    lines = out.splitlines()
    for i, l in enumerate(lines):
        if "Compiling entry function" in l and ("Li1ELi1ELi1ELi1EJNS_5DiodeES2_EEELb0" in l or "Li2ELi1ELi1ELi0EJEEELb1" in l):
            print(name, l.split("'")[1][20:60], "|", lines[i + 2].strip()[:70], "|", lines[i + 3].strip()[:40])
    lib = os.path.join(libs, f"lib_{name}.so")
    objs = [obj if s == "tpi.cu" else b._obj(s) for s in b.SOURCES]
    r = subprocess.run([b.nvcc(), "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
                        "-Xcompiler", "-fPIC", "-o", lib] + objs, capture_output=True, text=True)
    print(name, "->", lib if r.returncode == 0 else r.stderr[-500:])
    os.remove(obj)
