import os, subprocess, sys, glob
sys.path.insert(0, '/root/repo')
import acme_jl_b200._build as b
variants = {}
for spec in sys.argv[1:]:
    name, tpb, mb, T = spec.split(':')
    variants[name] = (int(tpb), int(mb), int(T))
os.makedirs('/root/repo/tools/libs', exist_ok=True)
for f in glob.glob('/root/repo/tools/libs/*.so'): os.remove(f)
procs = []
for name, (tpb, mb, T) in variants.items():
    out = f'/root/repo/tools/libs/lib_{name}.so'
    cmd = [b.nvcc()] + b.NVCC_FLAGS + [f'-DACME_TPI_TPB={tpb}', f'-DACME_TPI_MINB={mb}', f'-DACME_TPI_T={T}', '-Xptxas', '-v', '-o', out] + [os.path.join(b.CSRC, s) for s in b.SOURCES]
    procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for name, p in procs:
    o = p.communicate()[0]
    lines = o.splitlines()
    for i, l in enumerate(lines):
        if 'Compiling entry function' in l and 'Li1ELi1ELi1ELi1EJNS_5DiodeES2_EEELb0' in l:
            print(name, lines[i+2].strip()[:75], '|', lines[i+3].strip()[:40])
