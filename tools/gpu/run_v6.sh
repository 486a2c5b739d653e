set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/gputests_v6.log 2>&1
python bench.py --no-cpu-baseline > gpurun_out/bench_v6_offline.json 2> gpurun_out/bench_v6_offline.err
python bench.py --variant online --no-cpu-baseline > gpurun_out/bench_v6_online.json 2> gpurun_out/bench_v6_online.err
tail -4 gpurun_out/gputests_v6.log; cut -c1-400 gpurun_out/bench_v6_offline.json; python - <<'PY'
import json
for f in ("offline","online"):
    d=json.load(open(f"gpurun_out/bench_v6_{f}.json"))
    print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], [(k["kernel"], k["avg_ms"]) for k in d["kernels"]])
PY
