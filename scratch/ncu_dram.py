"""profiles/ncu_dram.json from an `ncu --set full` capture of `bench.py --workload configs1` (one denoiser step is enough):

    ncu -i gpurun_out/<capture>.ncu-rep --page raw --csv > /tmp/raw.csv
    python scratch/ncu_dram.py /tmp/raw.csv "<what was captured>" > profiles/ncu_dram.json

For every kernel name (bench.py's short names) the mean over its captured launches of dram__bytes_read.sum +
dram__bytes_write.sum, plus the csrc_sha of the build the capture was taken from: bench.py reports `roofline.traffic` only
when that hash equals the hash of the sources it runs."""
import collections, csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import _demangle, csrc_sha

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
iN, iR, iW, iT = hdr.index('Kernel Name'), hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum'), hdr.index('gpu__time_duration.sum')
mul = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
tmul = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}
acc = collections.defaultdict(list)
for r in rows[2:]:
    name = r[iN]
    # ncu prints demangled names: map them to bench.py's short names
    short = name.split('(')[0].replace('void ', '').replace('<unnamed>::', '')
    if short.startswith('gemm_tc_kernel<'):
        a = short[len('gemm_tc_kernel<'):].rstrip('>').replace('(int)', '').replace('(GemmMode)', '').replace(' ', '').split(',')
        mode = ('STORE', 'LNMOD', 'RESGATE', 'COORD', 'EHEAD')[int(a[1])] if a[1].isdigit() else a[1]
        short = 'gemm_tc_kernel<%s,%s,%s>' % (a[0], mode, {'true': '1', 'false': '0'}.get(a[2], a[2]))
    else:
        short = short.split('<')[0]
    by = float(r[iR].replace(',', '')) * mul[units[iR]] + float(r[iW].replace(',', '')) * mul[units[iW]]
    acc[short].append((by, float(r[iT].replace(',', '')) * tmul[units[iT]]))
out = {'source': sys.argv[2] if len(sys.argv) > 2 else sys.argv[1], 'csrc_sha': csrc_sha(),
       'kernels': {k: sum(b for b, _ in v) / len(v) for k, v in acc.items()},
       'launches': {k: len(v) for k, v in acc.items()},
       'us_under_ncu': {k: sum(t for _, t in v) / len(v) for k, v in acc.items()}}
print(json.dumps(out, indent=1, sort_keys=True))
