"""Summarise an `ncu --set full` report exported with `ncu -i X.ncu-rep --page raw --csv` into a markdown table."""
import csv, sys, re
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
def col(name):
    return hdr.index(name) if name in hdr else None
C = {k: col(k) for k in ['Kernel Name', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
     'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
     'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
     'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
     'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct']}
def f(r, k, fmt='%.1f'):
    i = C.get(k)
    if i is None or r[i] == '': return 'n/a'
    try: return fmt % float(r[i].replace(',', ''))
    except ValueError: return r[i]
def tobytes(r, k):
    i = C[k]; v = float(r[i].replace(',', '')); u = units[i]
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
def tous(r):
    i = C['gpu__time_duration.sum']; v = float(r[i].replace(',', '')); u = units[i]
    return v * {'ns': 1e-3, 'us': 1, 'ms': 1e3, 'usecond': 1, 'nsecond': 1e-3, 'msecond': 1e3}[u]
tp = [k for k in hdr if 'pipe_tensor' in k and 'pct' in k]
print('| # | kernel | grid | us | DRAM rd MB | DRAM wr MB | DRAM GB/s | dram % | L2 % | L1 % | SM % | issue % | warps act % | regs | L2 hit % | tensor pipe % | Minst |')
print('|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|')
for n, r in enumerate(rows[2:]):
    name = re.sub(r'\(.*', '', r[C['Kernel Name']]).replace('void ', '').replace('<unnamed>::', '')
    us = tous(r); rd = tobytes(r, 'dram__bytes_read.sum'); wr = tobytes(r, 'dram__bytes_write.sum')
    def _num(v):
        try: return float(v.replace(',', ''))
        except ValueError: return 0.0
    tens = max([_num(r[hdr.index(k)]) for k in tp] or [0])
    print('| %d | `%s` | %s | %.1f | %.1f | %.1f | %.0f | %s | %s | %s | %s | %s | %s | %s | %s | %.1f | %s |' % (
        n, name[:48], r[C['Grid Size']].replace(', 1, 1', '').strip('()'), us, rd / 1e6, wr / 1e6, (rd + wr) / us / 1e3,
        f(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'), f(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'),
        f(r, 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed'), f(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'),
        f(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'), f(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'),
        f(r, 'launch__registers_per_thread', '%d'), f(r, 'lts__t_sector_hit_rate.pct'), tens,
        '%.1f' % (float(r[C['smsp__inst_executed.sum']].replace(',', '')) / 1e6) if C['smsp__inst_executed.sum'] is not None else 'n/a'))
