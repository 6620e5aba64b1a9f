"""Per-kernel SASS opcode histogram of the in-tree library: which kernels are tcgen05 / TMEM / bulk-copy (Blackwell-native) and
which still use the warp-level mma.sync path.  Usage: python scripts/sass_histogram.py > profiles/r02_sass_histogram.txt"""
import collections, re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'strive_b200', 'libstrive_b200.so')
out = subprocess.run(['cuobjdump', '-sass', so], stdout=subprocess.PIPE, text=True).stdout
KEYS = ['UTCHMMA', 'UTCIMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTCBAR', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'SYNCS', 'HMMA', 'IMMA', 'FFMA', 'LDG', 'STG', 'LDS', 'STS', 'ATOM', 'RED', 'BAR']
cur, hist, total = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        total[cur] = 0
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if cur and m:
        op = m.group(1)
        total[cur] += 1
        for k in KEYS:
            if op.split('.')[0] == k or op.startswith(k + '.'):
                hist[cur][k] += 1
def demangle(n):
    r = subprocess.run(['c++filt', n], stdout=subprocess.PIPE, text=True).stdout.strip()
    return re.sub(r'\(.*', '', r)[:70]
print('SASS opcode histogram per kernel of strive_b200/libstrive_b200.so (cuobjdump -sass, sm_100a)')
print('UTCHMMA/UTCIMMA = tcgen05.mma (f16 / i8 kinds), LDTM/STTM = tcgen05.ld/st (TMEM), UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk,')
print('UTMALDG/UTMASTG = tensor-map TMA (not used: operands are staged by producer warps, see DESIGN.md 5), HMMA = warp-level mma.sync')
print()
print('%-70s %6s  %s' % ('kernel', 'instr', ' '.join('%7s' % k for k in KEYS)))
for k, h in hist.items():
    if total[k] < 50:
        continue
    print('%-70s %6d  %s' % (demangle(k), total[k], ' '.join('%7d' % h[x] for x in KEYS)))
tot = collections.Counter()
for h in hist.values():
    tot.update(h)
print()
print('%-70s %6d  %s' % ('TOTAL', sum(total.values()), ' '.join('%7d' % tot[x] for x in KEYS)))
