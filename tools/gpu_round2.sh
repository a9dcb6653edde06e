#!/bin/bash
# Round-2 profiling visit (one GPU): ncu launch lists of the bench command for cfg2 and cfg3, and ncu --set full captures of
# the kernels VERDICT r1 found without one: wgrad_halo2 (all 14 launches of one step), the transposed-conv kernels
# (scatter-epilogue conv_halo forward, igemm backward-data / backward-filter). Each only after its command ran clean.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs"
$B > gpurun_out/ncu_plain_cfg2.log 2>&1 || { echo "plain cfg2 bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_cfg2.csv $B > /dev/null 2>&1; echo "launch list cfg2 exit=$?"
$B --config cfg3 > gpurun_out/ncu_plain_cfg3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02_launches_cfg3.csv $B --config cfg3 > /dev/null 2>&1; echo "launch list cfg3 exit=$?"
S="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs"
ncu --set full --clock-control none --import-source on -k regex:wgrad_halo2_kernel --launch-skip 42 -c 14 -f -o gpurun_out/r02_full_wgrad_halo2 $S > /dev/null 2>&1; echo "full wgrad_halo2 exit=$?"
ncu --set full --clock-control none --import-source on -k regex:igemm_kernel --launch-skip 15 -c 5 -f -o gpurun_out/r02_full_convT_bwd $S > /dev/null 2>&1; echo "full igemm (convT bwd) exit=$?"
ncu --set full --clock-control none --import-source on -k "regex:conv_halo_kernel<(128|256), 2, 0, 0, 1" --launch-skip 12 -c 4 -f -o gpurun_out/r02_full_convT_fwd $S > /dev/null 2>&1; echo "full convT fwd exit=$?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null
