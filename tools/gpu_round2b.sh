#!/bin/bash
# Round-2 closing profiling visit (one GPU): ncu launch list of one cfg4 (UNet3D) training step and ncu --set full
# captures of its tensor-core launches (pixel-pair packed level, (3,3,3) halo kernels, strided layers), each only after the
# plain command ran clean. The reports are summarised ON THE BOX (tools/ncu_summary.py) and deleted: gpurun_out/ may not
# exceed 64 MiB.
set -u
mkdir -p gpurun_out
P="python tools/step_profile.py --model unet3d --steps 1 --warmup 2"
$P > gpurun_out/ncu_plain_cfg4.log 2>&1 || { echo "plain cfg4 step failed"; cat gpurun_out/ncu_plain_cfg4.log; exit 1; }
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_cfg4.csv $P > /dev/null 2>&1; echo "launch list cfg4 exit=$?"
python tools/ncu_summary.py --launches gpurun_out/r02_launches_cfg4.csv --out gpurun_out/r02_ncu_launches_cfg4.txt > /dev/null
ncu --profile-from-start off --set full --clock-control none -k "regex:conv_halo_kernel|wgrad_halo|igemm_kernel" -c 100 -f -o /tmp/r02_full_cfg4_conv $P > /dev/null 2>&1; echo "full cfg4 conv kernels exit=$?"
python tools/ncu_summary.py /tmp/r02_full_cfg4_conv.ncu-rep --out gpurun_out/r02_ncu_full_cfg4_conv.txt > /dev/null; echo "summary exit=$?"
rm -f /tmp/r02_full_cfg4_conv.ncu-rep
du -sh gpurun_out
