cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
export FNSSL_TC_WAIT_TIMEOUT=1
timeout 300 python tools/lstm_time.py > $O/r2_time_v8.log 2>&1; cat $O/r2_time_v8.log
( time python -m pytest tests -m gpu -x -q -rs ) > $O/r2_gputests_8.log 2>&1; tail -14 $O/r2_gputests_8.log
unset FNSSL_TC_WAIT_TIMEOUT
python __graft_entry__.py smoke > $O/r2_smoke_8.log 2>&1; tail -3 $O/r2_smoke_8.log
python bench.py > $O/r2_bench_8.json 2> $O/r2_bench_8.err; tail -c 600 $O/r2_bench_8.err
python tools/bench_ipdnet2.py cfg5 > $O/r2_bench_ipdnet2_8.jsonl 2> $O/r2_bench_ipdnet2_8.err
python tools/bench_extra.py ipdnet > $O/r2_extra_8.jsonl 2>&1
ncu --set full --clock-control none --import-source on -k regex:lstm_tc4 -s 6 -c 6 -o $O/r2_prof_v8_cfg4_b256 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > $O/r2_ncu_8a.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2_launches_v8_cfg4_b256.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $O/r2_ncu_8b.log 2>&1
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_8.json"))
print("cfg4", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], [(k["kernel"], k["avg_ms"], k["tflops"]) for k in d["kernels"]])
print("torch", json.dumps(d["gpu_torch_baseline"])[:400])
for k,v in (d.get("extra") or {}).items(): print(k, v["value"], v["ms_per_step"], v["e2e"], [(x["kernel"], x["avg_ms"]) for x in v["kernels"]])
for l in open("gpurun_out/r2_bench_ipdnet2_8.jsonl"):
    dd=json.loads(l); print(dd["workload"][:40], dd["ms_per_step"], dd["frames_per_s"], {k:v["avg_ms"] for k,v in dd["kernels"].items()})
print(open("gpurun_out/r2_extra_8.jsonl").read()[:600])
PY
