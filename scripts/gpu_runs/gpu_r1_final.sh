# end-of-round check: what the driver runs
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference 2>/dev/null | tail -1 | cut -c1-200
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
r=d['relation_tokens_per_sec']
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'sep_calls',round(d['ms_per_step_separate_calls'],3),'e2e',round(d['e2e']['value']),d['e2e']['ms_per_step_both_passes'],
      'gemm TF/s',round(d['roofline']['achieved']),'frac',round(d['roofline']['frac'],3),'xattn TF/s',round(d['roofline_xattn']['achieved']),
      'llm',round(r['value']),round(r['ms_per_image'],1),'roofline',r['roofline']['frac'],'cpu',round(d['cpu_baseline']['value']),'clocks',d['clocks'],'launches',d['gpu_launches'])
PY
