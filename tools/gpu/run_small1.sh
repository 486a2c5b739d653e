set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
FNSSL_TC_SMALL1=1 python bench.py --no-cpu-baseline > gpurun_out/bench_s1f_offline.json 2>/dev/null
python bench.py --no-cpu-baseline > gpurun_out/bench_s1d_offline.json 2>/dev/null
FNSSL_TC_SMALL1=1 python tools/bench_extra.py ipdnet > gpurun_out/extra_s1f.jsonl 2>&1
python - <<'PY'
import json
for f in ("s1f","s1d"):
    d=json.load(open(f"gpurun_out/bench_{f}_offline.json"))
    print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], [(k["kernel"], k["avg_ms"]) for k in d["kernels"]])
PY
cat gpurun_out/extra_s1f.jsonl
