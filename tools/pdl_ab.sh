# A/B of programmatic dependent launch modes (engine.py BSL_PDL) on the headline bench; one gpurun call.
set -u
mkdir -p gpurun_out
run() {  # label, env...
  label=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$label.log 2> gpurun_out/bench_$label.err; rc=$?
  python - <<PY
import json
l=[x for x in open("gpurun_out/bench_$label.log") if x.startswith("{")]
d=json.loads(l[-1]) if l else {}
print("$label rc=$rc", d.get("ms_per_step"), d.get("value"), (d.get("clocks") or {}).get("sm_mhz"))
PY
}
run pdl0 BSL_PDL=0
run pdl1 BSL_PDL=1
run pdl0b BSL_PDL=0
run pdl1b BSL_PDL=1
