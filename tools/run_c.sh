# closing verification visit: whole GPU suite with the parity report, smoke, default bench line
rm -f gpurun_out/parity_report_f.jsonl
BSL_PARITY_REPORT=gpurun_out/parity_report_f.jsonl python -m pytest tests -q -m gpu 2>&1 | tail -n 6 > gpurun_out/pytest_full_f.log; cat gpurun_out/pytest_full_f.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1 | tee gpurun_out/smoke_f.log
python bench.py > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err
python -c "
import json
l=json.loads(open('gpurun_out/bench_f.json').read().strip().splitlines()[-1])
print(l['value'], l['ms_per_step'], l['roofline']['frac'], l['e2e']['value'], l['clocks'])
for k,v in l['other_configs'].items(): print(k, v['ms_per_step'], v.get('frac_of_sustained_bf16_peak'))
"
