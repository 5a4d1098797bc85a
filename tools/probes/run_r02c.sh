#!/bin/bash
# Round 2, last part: GPU suite (minus the three slowest FPAQ / golden cases), ZRLT A/B, bench, ZRLT launch list.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x --durations=8 \
  -k "not (large_blocks and FPAQ) and not (matches_golden and 21)" > gpurun_out/r02c_gpu_suite.log 2>&1
echo "suite rc=$?"
tail -3 gpurun_out/r02c_gpu_suite.log
timeout 120 python tools/probes/profile_all.py headline 256 > gpurun_out/r02c_zrlt_lean.log 2>&1
KNZ_ZRLT_BYTEWALK=1 timeout 120 python tools/probes/profile_all.py headline 256 > gpurun_out/r02c_zrlt_bytewalk.log 2>&1
tail -2 gpurun_out/r02c_zrlt_lean.log gpurun_out/r02c_zrlt_bytewalk.log
timeout 200 python bench.py --steps 3 --warmup 3 > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err
echo "bench rc=$?"
timeout 120 ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  -k regex:zrlt --csv --log-file gpurun_out/r02c_zrlt_launches.csv python tools/probes/profile_all.py headline 64 \
  > gpurun_out/r02c_zrlt_ncu.log 2>&1
echo "ncu rc=$?"
