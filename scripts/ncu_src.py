"""Summarise an `ncu --page source --csv` dump: instructions executed per SASS line (hot ones) and stall samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
hdr = rows[1]
ie, src, smp = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('# Samples')
agg = []
for r in rows[2:]:
    try:
        agg.append((int(r[ie]), int(r[smp] or 0), r[src]))
    except (ValueError, IndexError):
        continue
tot = sum(a[0] for a in agg); ts = sum(a[1] for a in agg)
print('total warp-instructions', tot, 'SASS lines', len(agg), 'samples', ts)
for i, (n, s, t) in enumerate(agg):
    if n > frac * tot or s > frac * 2 * ts:
        print('%5d %11d %5.2f%% smp %5.2f%%  %s' % (i, n, 100.0 * n / tot, 100.0 * s / max(ts, 1), t[:100]))
