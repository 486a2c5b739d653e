set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/gputests_fe.log 2>&1
tail -5 gpurun_out/gputests_fe.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 12 -k regex:"stft|assemble|norm_scan|head" --csv --log-file gpurun_out/launches_fe.csv python bench.py --no-cpu-baseline --steps 2 > gpurun_out/ncu_fe.log 2>&1
grep -E "stft|assemble|head|norm" gpurun_out/launches_fe.csv | awk -F'","' '{print $5, $(NF)}' | cut -c1-100 | head -8
python bench.py --no-cpu-baseline | cut -c1-200
