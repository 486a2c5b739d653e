set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_ipdnet2.py -m gpu -x -q 2>&1 | tail -3 ) > gpurun_out/gputests_misc.log 2>&1
cat gpurun_out/gputests_misc.log
python tools/bench_ipdnet2.py cfg5 default 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['ms_per_step'], d['frames_per_s'], {k:v['avg_ms'] for k,v in d['kernels'].items()})"
