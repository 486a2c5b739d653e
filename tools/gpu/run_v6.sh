set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/gputests_v7.log 2>&1
python __graft_entry__.py smoke > gpurun_out/smoke_v7.log 2>&1
python bench.py > gpurun_out/bench_v7_offline.json 2> gpurun_out/bench_v7_offline.err
python bench.py --variant online --no-cpu-baseline > gpurun_out/bench_v7_online.json 2> gpurun_out/bench_v7_online.err
python tools/bench_ipdnet2.py cfg5 default --cpu > gpurun_out/bench_ipdnet2_v7.jsonl 2> gpurun_out/bench_ipdnet2_v7.err
python tools/bench_extra.py ipdnet fnssl_b64 fnssl_b64_online > gpurun_out/extra_v7.jsonl 2>&1
tail -4 gpurun_out/gputests_v7.log; tail -3 gpurun_out/smoke_v7.log; python - <<'PY'
import json
for f in ("offline","online"):
    d=json.load(open(f"gpurun_out/bench_v7_{f}.json"))
    print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], [(k["kernel"], k["avg_ms"]) for k in d["kernels"]])
print(open("gpurun_out/extra_v7.jsonl").read())
for l in open("gpurun_out/bench_ipdnet2_v7.jsonl"):
    d=json.loads(l); print(d["workload"][:40], d["ms_per_step"], d["frames_per_s"], {k:v["avg_ms"] for k,v in d["kernels"].items()})
PY
