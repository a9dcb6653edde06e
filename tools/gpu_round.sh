#!/bin/bash
# One GPU-box visit: parity suite, headline bench, ncu launch list of the same bench command, and
# ncu --set full captures of the dominant kernels (each only after its command exited 0 without ncu).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -2 gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit=$?"
cat gpurun_out/bench.log
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2>&1; echo "ref exit=$?"
cat gpurun_out/bench_ref.log
python tools/step_breakdown.py --json gpurun_out/breakdown.json > gpurun_out/breakdown.log 2>&1; echo "breakdown exit=$?"
python tools/timeline.py --out gpurun_out/timeline.txt > /dev/null 2>&1; echo "timeline exit=$?"
head -24 gpurun_out/breakdown.log
if [ "${SKIP_NCU:-0}" != "1" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit=$?"
  [ "${SKIP_FULL:-0}" = "1" ] && exit 0
  for spec in "dec1_1 wgrad wgrad_halo_kernel" "dec1_1 fprop conv_halo_kernel" "dec3_1 dgrad conv_halo_kernel" "enc1_2 dgrad conv_halo_kernel"; do
    set -- $spec
    extra=""; [ "$2" = "fprop" ] && extra="--stats"
    python tools/gpu_conv_bench.py --layers $1 --ops $2 --reps 1 $extra > /dev/null 2>&1 || { echo "plain $spec failed"; continue; }
    ncu --set full --clock-control none --import-source on -k regex:$3 -c 1 -f -o gpurun_out/full_$1_$2 \
        python tools/gpu_conv_bench.py --layers $1 --ops $2 --reps 1 $extra > gpurun_out/ncu_full_$1_$2.log 2>&1; echo "ncu full $spec exit=$?"
  done
fi
