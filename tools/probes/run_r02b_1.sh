set -x
python -m pytest tests/test_stream_features.py -x -q -m gpu > gpurun_out/r02b_stream_features.log 2>&1; tail -3 gpurun_out/r02b_stream_features.log
for v in 1 2; do KNZ_SBRT_INV=$v python tools/probes/rank_inv_variants.py 64 2>&1 | tail -1 | tee -a gpurun_out/r02b_rank_inv_variants.log; done
