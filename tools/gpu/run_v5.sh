set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/gputests_v5.log 2>&1
python bench.py > gpurun_out/bench_v5_offline.json 2> gpurun_out/bench_v5_offline.err
python bench.py --variant online --no-cpu-baseline > gpurun_out/bench_v5_online.json 2> gpurun_out/bench_v5_online.err
python tools/bench_extra.py ipdnet fnssl_b64 > gpurun_out/extra_v5.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v5.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_v5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lstm_tc4 -s 10 -c 3 -o gpurun_out/prof_tc4_v5 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tc4_v5.log 2>&1
ls -la gpurun_out
