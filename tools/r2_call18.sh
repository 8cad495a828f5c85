# round 2, call 18: ncu launch list of the bench command with the final build (kernel share of a step)
OUT=gpurun_out
python bench.py --steps 2 --warmup 1 --no-extra > $OUT/r2c_bench_short.json 2> $OUT/r2c_bench_short.err; echo rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r2c_launches.csv python bench.py --steps 2 --warmup 1 --no-extra > $OUT/r2c_ncu_launches.log 2>&1; echo ncu rc=$?
python - <<'P'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r2c_launches.csv')) if len(r) > 5 and r[0].isdigit()]
t = collections.Counter(); n = collections.Counter()
for r in rows:
    name = r[4].split('(')[0][:60]; v = float(r[-1].replace(',', '')); u = r[-2]
    v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1, 'msecond': 1, 'usecond': 1e-3, 'nsecond': 1e-6, 's': 1e3}.get(u, 1)
    t[name] += v; n[name] += 1
for k, v in t.most_common(12): print(f'{v:10.3f} ms {n[k]:4d}  {k}')
P
