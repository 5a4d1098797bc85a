#!/bin/bash
# Round 2, closing run: bench N=1 on the final tree, then the ZRLT launch list (time + DRAM bytes).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 75 python bench.py --steps 3 --warmup 3 > gpurun_out/r02e_bench_n1.json 2> gpurun_out/r02e_bench_n1.err
echo "bench rc=$?"
tail -c 600 gpurun_out/r02e_bench_n1.json
timeout 40 ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  -k regex:zrlt -c 150 --csv --log-file gpurun_out/r02e_zrlt_launches.csv python tools/probes/final_check.py 64 \
  > gpurun_out/r02e_zrlt_ncu.log 2>&1
echo "ncu rc=$?"
