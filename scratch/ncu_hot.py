"""Top sampled SASS instructions of one kernel: ncu -i rep --page source --csv (sass view) piped through this."""
import csv, sys, subprocess
rep, kid, top = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--launch-skip', kid, '--launch-count', '1'], capture_output=True, text=True).stdout
lines = out.splitlines()
print(lines[0][:150])
rows = list(csv.reader(lines[1:]))
hdr = rows[0]
iS, iSrc, iEx = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
data = [(int(r[iS] or 0), n, r[iSrc].strip(), r[iEx]) for n, r in enumerate(rows[1:]) if len(r) > iS and r[iS].isdigit()]
tot = sum(d[0] for d in data)
print('total samples', tot, 'instructions', len(data))
for s, n, src, ex in sorted(data, reverse=True)[:top]:
    print('%5.1f%%  #%4d  x%-8s %s' % (100.0 * s / tot, n, ex, src[:100]))
