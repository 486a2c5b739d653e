set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_ipdnet2.py -m gpu -x -q ) > gpurun_out/gputests_ipdnet2.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python tools/bench_ipdnet2.py tiny > gpurun_out/sanitizer_ipdnet2.log 2>&1
timeout 600 python tools/bench_ipdnet2.py cfg5 default --cpu > gpurun_out/bench_ipdnet2.jsonl 2> gpurun_out/bench_ipdnet2.err
tail -5 gpurun_out/gputests_ipdnet2.log; tail -5 gpurun_out/sanitizer_ipdnet2.log; cat gpurun_out/bench_ipdnet2.jsonl
