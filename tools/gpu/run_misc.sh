set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_ipdnet2.py tests/test_gpu_parity.py -m gpu -x -q -k "ipdnet or lstm or stream or network" 2>&1 | tail -3 ) > gpurun_out/gputests_misc.log 2>&1
cat gpurun_out/gputests_misc.log
python tools/bench_ipdnet2.py cfg5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['frames_per_s'], {k:v['avg_ms'] for k,v in d['kernels'].items()})"
python tools/bench_extra.py ipdnet 2>&1 | tail -1
