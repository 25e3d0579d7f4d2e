mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/prefill_launches.csv python tools/prof_prefill.py 8 2048 1 > gpurun_out/prof_prefill.log 2>&1
tail -2 gpurun_out/prof_prefill.log
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/prefill_launches.csv') if l.startswith('"')))
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(',', ''));
    v = v / 1e6 if r[ui] == 'ns' else v / 1e3 if r[ui] in ('us', 'usecond') else v
    name = r[ki].split('(')[0][:60]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{ms:9.2f} ms {n:5d} x {k}  ({100 * ms / tot:.1f}%)")
print(f"{tot:9.2f} ms total under ncu (serialised, cold caches)")
PY
