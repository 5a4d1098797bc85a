python tools/probes/levels_probe.py 256 2>&1 | tee gpurun_out/r02b_levels_probe.log
ncu --clock-control none --metrics gpu__time_duration.sum -c 3000 --csv --log-file gpurun_out/r02b_launches_levels.csv python tools/probes/levels_probe.py 64 > /dev/null 2>&1
