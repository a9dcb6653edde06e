#!/bin/bash
# One GPU-box visit late in the round: the full round script with the committed defaults, then the pending A/B
# experiments (head-backward fusion, late PDL trigger).
set -u
bash tools/gpu_round.sh
echo "=== experiments"
BSL_FUSE_HEAD_BWD=1 timeout 600 python -m pytest tests/test_gpu_unet.py tests/test_golden.py tests/test_gpu_gunet.py tests/test_gpu_host_api.py -m gpu -x -q 2>&1 | tail -3
BSL_FUSE_HEAD_BWD=1 python tools/step_breakdown.py > gpurun_out/bd_hb.log 2>&1; head -24 gpurun_out/bd_hb.log | tail -20
late=""; [ -f boxsegliver_b200/libbsl_b200_late.so ] && late="late BSL_LIB=boxsegliver_b200/libbsl_b200_late.so"
bash tools/ab.sh "base X=0" "hb BSL_FUSE_HEAD_BWD=1" "$late" "base2 X=0" "hb2 BSL_FUSE_HEAD_BWD=1" "${late/late /late2 }"
