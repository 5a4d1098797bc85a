python -m pytest tests -x -q -m gpu --durations=15 > gpurun_out/r02b_gpu_suite.log 2>&1; tail -30 gpurun_out/r02b_gpu_suite.log
