set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/gputests_stft.log 2>&1
tail -5 gpurun_out/gputests_stft.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 -k regex:"stft|assemble|reflect|norm_scan|head" --csv --log-file gpurun_out/launches_stft.csv python tools/bench_ipdnet2.py cfg5 > gpurun_out/ncu_stft.log 2>&1
grep -E "stft|assemble|reflect|head" gpurun_out/launches_stft.csv | awk -F'","' '{print $5, $(NF)}' | cut -c1-120 | head -12
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 12 -k regex:"stft|assemble|norm_scan|head" --csv --log-file gpurun_out/launches_stft2.csv python bench.py --no-cpu-baseline --steps 2 > gpurun_out/ncu_stft2.log 2>&1
grep -E "stft|assemble|head|norm" gpurun_out/launches_stft2.csv | awk -F'","' '{print $5, $(NF)}' | cut -c1-120 | head -12
