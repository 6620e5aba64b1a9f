"""Stall-sample summary of one kernel's `ncu --page source --csv` section: totals per stall reason and the hottest SASS lines."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.012
hdr = rows[1]
ie, src, smp = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('# Samples')
stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
agg = []
for r in rows[2:]:
    try:
        agg.append((int(r[ie]), int(r[smp] or 0), r[src], r))
    except (ValueError, IndexError):
        pass
tot = sum(a[0] for a in agg); ts = sum(a[1] for a in agg)
print('warp-instructions', tot, 'samples', ts)
st = {hdr[i]: 0 for i in stalls}
for a in agg:
    for i in stalls:
        try:
            st[hdr[i]] += int(a[3][i] or 0)
        except ValueError:
            pass
print(sorted(st.items(), key=lambda kv: -kv[1])[:10])
for i, (n, s, t, r) in enumerate(agg):
    if s > frac * ts:
        top = sorted([(int(r[j] or 0), hdr[j]) for j in stalls], reverse=True)[:2]
        print('%5d n=%9d smp %5.2f%% %-72s %s' % (i, n, 100.0 * s / ts, t[:72], top))
