#!/bin/bash
# One GPU visit: parity tests, bench line, ncu launch list.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/diag_gpu.txt
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu_info.txt 2>&1
echo "nproc=$(nproc)" >> gpurun_out/gpu_info.txt
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps ${BENCH_STEPS:-5} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "${RUN_NCU:-1}" = "1" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_SKIP:-1200} -c ${NCU_COUNT:-420} --csv --log-file gpurun_out/launches.csv \
     python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
  python - <<'PY'
import csv, collections
rows = []
with open('gpurun_out/launches.csv') as f:
    lines = [l for l in f if not l.startswith('==')]
r = csv.DictReader(lines)
agg = collections.OrderedDict()
for row in r:
    try:
        v = float(row['Metric Value'].replace(',', ''))
    except Exception:
        continue
    k = row['Kernel Name'].split('(')[0]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print('ncu launch list: %d kernels, total %.3f ms' % (sum(a[0] for a in agg.values()), tot / 1e6))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print('  %-60s n=%4d  %.3f ms  %.1f%%' % (k[:60], a[0], a[1] / 1e6, 100 * a[1] / tot))
PY
fi
echo "---- diag"; tail -40 gpurun_out/diag_gpu.txt 2>/dev/null
