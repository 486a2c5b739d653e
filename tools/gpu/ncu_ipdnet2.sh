set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:sn_freq_kernel -s 0 -c 1 -o gpurun_out/prof_ipdnet2_freq python tools/bench_ipdnet2.py default > gpurun_out/ncu_ipdnet2_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sn_time_kernel -s 0 -c 1 -o gpurun_out/prof_ipdnet2_time python tools/bench_ipdnet2.py default > gpurun_out/ncu_ipdnet2_b.log 2>&1
tail -3 gpurun_out/ncu_ipdnet2_a.log
