# Generic A/B of environment switches on the headline bench, one gpurun call:  bash tools/ab.sh "label VAR=val ..." ...
set -u
mkdir -p gpurun_out
for spec in "$@"; do
  set -- $spec; label=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$label.log 2> gpurun_out/bench_$label.err; rc=$?
  python - <<PY
import json
l=[x for x in open("gpurun_out/bench_$label.log") if x.startswith("{")]
d=json.loads(l[-1]) if l else {}
print("$label rc=$rc", d.get("ms_per_step"), d.get("value"), (d.get("clocks") or {}).get("sm_mhz"))
PY
  [ $rc -ne 0 ] && tail -5 gpurun_out/bench_$label.err
done
