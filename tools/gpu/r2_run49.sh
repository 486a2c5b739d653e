# round 2, call 49: the shipped library -- full GPU suite, smoke, bench
cd $GRAFT_REPO_ROOT; O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -x -q -rs ) > $O/r2_gputests_49.log 2>&1; tail -5 $O/r2_gputests_49.log
python __graft_entry__.py smoke > $O/r2_smoke_49.log 2>&1; tail -3 $O/r2_smoke_49.log
python bench.py > $O/r2_bench_49.json 2> $O/r2_bench_49.err; tail -c 300 $O/r2_bench_49.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_49.json"))
print("cfg4", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], d["gpu_launches"], [(k["kernel"], k["avg_ms"], k["tflops"]) for k in d["kernels"]], d["clocks"])
for k,v in (d.get("extra") or {}).items(): print(k, v["value"], v["ms_per_step"], v["e2e"], [(x["kernel"], x["avg_ms"]) for x in v["kernels"]])
PY
