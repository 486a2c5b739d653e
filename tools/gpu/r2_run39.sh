# round 2, call 39: final validation + bench of the shipped dispatch (tc5 / tc6<256> / tc6<128> two-source / tc4)
cd $GRAFT_REPO_ROOT; O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -x -q -rs ) > $O/r2_gputests_39.log 2>&1; tail -8 $O/r2_gputests_39.log
python __graft_entry__.py smoke > $O/r2_smoke_39.log 2>&1; tail -3 $O/r2_smoke_39.log
python bench.py > $O/r2_bench_39.json 2> $O/r2_bench_39.err; tail -c 300 $O/r2_bench_39.err
python tools/bench_extra.py ipdnet > $O/r2_extra_39.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2_launches_v39_cfg4_b256.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $O/r2_ncu_39b.log 2>&1
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_39.json"))
print("cfg4", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], [(k["kernel"], k["avg_ms"], k["tflops"]) for k in d["kernels"]])
for k,v in (d.get("extra") or {}).items(): print(k, v["value"], v["ms_per_step"], v["e2e"], [(x["kernel"], x["avg_ms"]) for x in v["kernels"]])
print(open("gpurun_out/r2_extra_39.jsonl").read()[:600])
PY
