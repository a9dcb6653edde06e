#!/bin/bash
# Round-2 closing profiling visit of the final build (one GPU): ncu launch list of the cfg2 bench command, and ncu --set full
# captures of every tensor-core launch of one cfg2 training step (CTA-pair conv_halo kernels, wgrad_halo3, the 64-wide
# resident-filter kernels), each only after the plain command ran clean. Reports are summarised ON THE BOX
# (tools/ncu_summary.py) and deleted: gpurun_out/ may not exceed 64 MiB.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs"
$B > gpurun_out/ncu_plain_cfg2_c.log 2>&1 || { echo "plain cfg2 bench failed"; tail -5 gpurun_out/ncu_plain_cfg2_c.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02c_launches_cfg2.csv $B > /dev/null 2>&1; echo "launch list cfg2 exit=$?"
python tools/ncu_summary.py --launches gpurun_out/r02c_launches_cfg2.csv --out gpurun_out/r02c_ncu_launches_cfg2.txt > /dev/null
rm -f gpurun_out/r02c_launches_cfg2.csv
P="python tools/step_profile.py --model unet --steps 1 --warmup 2"
$P > gpurun_out/ncu_plain_step_c.log 2>&1 || { echo "plain cfg2 step failed"; tail -5 gpurun_out/ncu_plain_step_c.log; exit 1; }
ncu --profile-from-start off --set full --clock-control none -k "regex:conv_halo_kernel|wgrad_halo" -c 80 -f -o /tmp/r02c_full_cfg2_conv $P > /dev/null 2>&1; echo "full cfg2 conv kernels exit=$?"
python tools/ncu_summary.py /tmp/r02c_full_cfg2_conv.ncu-rep --out gpurun_out/r02c_ncu_full_cfg2_conv.txt > /dev/null; echo "summary exit=$?"
python tools/ncu_table.py gpurun_out/r02c_ncu_full_cfg2_conv.txt --out gpurun_out/r02c_ncu_full_cfg2_conv_table.txt; echo "table exit=$?"
rm -f /tmp/r02c_full_cfg2_conv.ncu-rep
du -sh gpurun_out
