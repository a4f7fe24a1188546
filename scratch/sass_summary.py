"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): UTC*MMA (tcgen05.mma),
LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA), UTCBAR (tcgen05.commit), plus legacy HMMA for contrast.

    python scratch/sass_summary.py > profiles/r2_sass_summary.md        (needs cuobjdump; no GPU)
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'diffspectra_b200', 'libdiffspectra_b200.so')
sys.path.insert(0, ROOT)
from bench import csrc_sha
out = subprocess.run(['cuobjdump', '-sass', SO], capture_output=True, text=True).stdout
KEYS = ['UTCHMMA', 'UTCHMMA.2CTA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTCBAR', 'UBLKCP', 'HMMA', 'LDGSTS', 'MUFU.TANH', 'FFMA2']
counts, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r'^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if not m:
        continue
    op = m.group(1)
    for k in KEYS:
        if op == k or op.startswith(k + '.'):
            counts[cur][k] += 1
def short(name):
    d = subprocess.run(['cu++filt', name], capture_output=True, text=True).stdout.strip() or name
    d = d.replace('void ', '').replace('(anonymous namespace)::', '').replace('<unnamed>::', '')
    if '>(' in d:
        d = d.split('>(')[0] + '>'
    else:
        d = d.split('(')[0]
    d = d.replace('(GemmMode)0', 'STORE').replace('(GemmMode)1', 'LNMOD').replace('(GemmMode)2', 'RESGATE').replace('(GemmMode)3', 'COORD').replace('(GemmMode)4', 'EHEAD')
    d = d.replace('(bool)1', 'true').replace('(bool)0', 'false')
    return d[:90]
print('# SASS evidence, round 2 (`cuobjdump -sass diffspectra_b200/libdiffspectra_b200.so`, csrc_sha %s)\n' % csrc_sha())
print('Counts of instructions per kernel; only kernels with at least one tensor-core / TMEM / TMA instruction are listed.\n')
print('| kernel | ' + ' | '.join(KEYS) + ' |')
print('|---|' + '---|' * len(KEYS))
tot = collections.Counter()
for name, c in counts.items():
    tot.update(c)
    if any(c[k] for k in KEYS[:8]):
        print('| `%s` | ' % short(name) + ' | '.join(str(c[k]) for k in KEYS) + ' |')
print('| **all %d kernels** | ' % len(counts) + ' | '.join(str(tot[k]) for k in KEYS) + ' |')
