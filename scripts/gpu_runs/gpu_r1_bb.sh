timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 800 --csv --log-file gpurun_out/llm_small_m_launches.csv python scripts/llm_probe.py 5 > gpurun_out/llm_probe.log 2>&1
tail -2 gpurun_out/llm_probe.log
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/llm_small_m_launches.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size')
agg=collections.OrderedDict()
for r in rows[1:]:
    key=(r[ki][:50], r[gi])
    a=agg.setdefault(key,[0,0.0]); a[0]+=1; a[1]+=float(r[vi].replace(',',''))
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print(f"{k[0]:52s} {k[1]:14s} n={n:4d} total_us={t/1000 if t>1e5 else t:10.1f} avg={t/n:9.1f}")
PY
