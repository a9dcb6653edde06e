# closing verification visit: whole GPU suite with the parity report, smoke, default bench line
rm -f gpurun_out/parity_report_e.jsonl
BSL_PARITY_REPORT=gpurun_out/parity_report_e.jsonl python -m pytest tests -q -m gpu 2>&1 | tail -n 6 > gpurun_out/pytest_full_e.log; cat gpurun_out/pytest_full_e.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1 | tee gpurun_out/smoke_e.log
python bench.py > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err
python -c "
import json
l=json.loads(open('gpurun_out/bench_e.json').read().strip().splitlines()[-1])
print(l['value'], l['ms_per_step'], l['roofline']['frac'], l['e2e']['value'], l['clocks'])
for k,v in l['other_configs'].items(): print(k, v['ms_per_step'], v.get('frac_of_sustained_bf16_peak'))
"
