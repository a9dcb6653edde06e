python bench.py > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; tail -c 200 gpurun_out/bench_b.err; python tools/bench_models.py --breakdown > gpurun_out/models_b.jsonl 2> gpurun_out/models_breakdown_b.log; bash tools/gpu_round2b.sh; python -c "
import json
l=json.loads(open('gpurun_out/bench_b.json').read().strip().splitlines()[-1])
print(l['value'], l['ms_per_step'], l['roofline']['frac'], l['e2e']['value'], l['clocks'])
for k,v in l['other_configs'].items(): print(k, v['ms_per_step'], v.get('frac_of_sustained_bf16_peak'))
"
