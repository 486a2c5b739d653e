set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lstm or stream" 2>&1 | tail -3 ) > gpurun_out/gputests_small1.log 2>&1
python bench.py --no-cpu-baseline > gpurun_out/bench_s1_offline.json 2>/dev/null
python bench.py --variant online --no-cpu-baseline > gpurun_out/bench_s1_online.json 2>/dev/null
python tools/bench_extra.py ipdnet > gpurun_out/extra_s1.jsonl 2>&1
cat gpurun_out/gputests_small1.log
python - <<'PY'
import json
for f in ("offline","online"):
    d=json.load(open(f"gpurun_out/bench_s1_{f}.json"))
    print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], [(k["kernel"], k["avg_ms"]) for k in d["kernels"]])
PY
cat gpurun_out/extra_s1.jsonl
