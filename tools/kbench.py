import glob, os, subprocess, sys
libs = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libs', 'lib_*.so')))
sel = sys.argv[1:]
for lib in libs:
    if sel and not any(s in lib for s in sel): continue
    env = dict(os.environ, ACMEB200_LIB=lib)
    out = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'kbench_one.py')], env=env, capture_output=True, text=True)
    print(out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:])
