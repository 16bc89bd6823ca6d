"""static SASS statistics of one kernel of the library: registers/spills from the loop body's point of view"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'acme.jl_b200', 'libacmeb200.so')
pat = sys.argv[2] if len(sys.argv) > 2 else 'Li1ELi1ELi1ELi1EJNS_5DiodeES2_EEELb0'
out = subprocess.check_output(['cuobjdump', '-sass', lib], text=True)
blocks = out.split('Function : ')
for b in blocks[1:]:
    name = b.split('\n', 1)[0]
    if pat not in name: continue
    ins = []
    for line in b.splitlines():
        m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);', line)
        if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
    print(name[:90], 'static instructions', len(ins))
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    # find backward branches (loops)
    loops = []
    for i, (a, s) in enumerate(ins):
        m = re.search(r'BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)', s)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_index:
                loops.append((addr_index[tgt], i))
    def opname(s):
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', s)
        return m.group(2).split('.')[0]
    for (lo, hi) in sorted(loops, key=lambda x: x[1]-x[0]):
        body = ins[lo:hi+1]
        ops = collections.Counter(opname(s) for _, s in body)
        if ops.get('MUFU', 0) >= 1 or len(body) > 100:
            print(f'  loop [{lo},{hi}] len {len(body)}:', dict(ops.most_common(14)), 'LDL', ops.get('LDL',0), 'STL', ops.get('STL',0))
    allops = collections.Counter(opname(s) for _, s in ins)
    print('  total LDL', allops.get('LDL', 0), 'STL', allops.get('STL', 0), 'UBLKCP', allops.get('UBLKCP', 0))
