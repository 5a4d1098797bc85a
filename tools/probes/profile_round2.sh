set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
NCU="ncu --clock-control none"
timeout 400 $NCU --metrics gpu__time_duration.sum -c 4000 --csv --log-file $O/r2_launches_256blocks.csv python tools/probes/profile_all.py headline 256 > $O/r2_prof_headline.log 2>&1
timeout 300 $NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum -c 4000 --csv --log-file $O/r2_traffic_64blocks.csv python tools/probes/profile_all.py headline 64 > /dev/null 2>&1
timeout 300 $NCU --set full --import-source on -k regex:"sbrt_inverse_fast|sbrt_rank_quad" -c 4 -o $O/r2_full_rank python tools/probes/profile_all.py headline 256 > /dev/null 2>&1
timeout 300 $NCU --set full --import-source on -k regex:"ans0_encode_kernel|ans0_decode_kernel|huf_encode_kernel|huf_decode_kernel|ans1_code_kernel|ans1_decode_kernel|ans1_hist_kernel|ans1_map_kernel" -c 12 -o $O/r2_full_entropy python tools/probes/profile_all.py entropy 64 > $O/r2_prof_entropy.log 2>&1
timeout 300 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -c 4000 --csv --log-file $O/r2_launches_entropy.csv python tools/probes/profile_all.py entropy 64 > /dev/null 2>&1
timeout 400 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -c 4000 --csv --log-file $O/r2_launches_config5.csv python tools/probes/profile_all.py config5 16 > $O/r2_prof_config5.log 2>&1
ls -la $O/r2_*
tail -5 $O/r2_prof_headline.log $O/r2_prof_entropy.log $O/r2_prof_config5.log
