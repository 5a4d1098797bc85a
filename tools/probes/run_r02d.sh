#!/bin/bash
# Round 2, last GPU call (~70 s of box time left): staged ZRLT stores, warp fold, rotate-shuffle deep RANK step.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 60 python tools/probes/final_check.py 256 > gpurun_out/r02d_final_check.log 2>&1
echo "final_check rc=$?"
cat gpurun_out/r02d_final_check.log
FINAL_CHECK_SKIP_BIG=1 KNZ_ZRLT_BYTEWALK=1 timeout 30 python tools/probes/final_check.py 256 > gpurun_out/r02d_final_check_bytewalk.log 2>&1
echo "bytewalk rc=$?"
tail -4 gpurun_out/r02d_final_check_bytewalk.log
