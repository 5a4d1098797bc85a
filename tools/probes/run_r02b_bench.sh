python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; tail -c 3000 gpurun_out/r02b_bench_n1.json; tail -5 gpurun_out/r02b_bench_n1.err
