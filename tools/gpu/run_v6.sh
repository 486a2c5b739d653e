set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/gputests_v6.log 2>&1
python __graft_entry__.py smoke > gpurun_out/smoke_v6.log 2>&1
python bench.py > gpurun_out/bench_v6_offline.json 2> gpurun_out/bench_v6_offline.err
python bench.py --variant online --no-cpu-baseline > gpurun_out/bench_v6_online.json 2> gpurun_out/bench_v6_online.err
python tools/bench_ipdnet2.py cfg5 default --cpu > gpurun_out/bench_ipdnet2_v6.jsonl 2> gpurun_out/bench_ipdnet2_v6.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 70 --csv --log-file gpurun_out/launches_ipdnet2_v6.csv python tools/bench_ipdnet2.py cfg5 > gpurun_out/ncu_ipdnet2_v6.log 2>&1
tail -4 gpurun_out/gputests_v6.log; tail -3 gpurun_out/smoke_v6.log; python - <<'PY'
import json
for f in ("offline","online"):
    d=json.load(open(f"gpurun_out/bench_v6_{f}.json"))
    print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], [(k["kernel"], k["avg_ms"]) for k in d["kernels"]])
for l in open("gpurun_out/bench_ipdnet2_v6.jsonl"):
    d=json.loads(l); print(d["workload"][:40], d["ms_per_step"], d["frames_per_s"], {k:v["avg_ms"] for k,v in d["kernels"].items()})
PY
