# round 2, final single-GPU validation + measurement of the shipped library
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
( time python -m pytest tests -m gpu -x -q -rs ) > $O/r2_gputests_31.log 2>&1; tail -14 $O/r2_gputests_31.log
python __graft_entry__.py smoke > $O/r2_smoke_31.log 2>&1; tail -3 $O/r2_smoke_31.log
python bench.py > $O/r2_bench_31.json 2> $O/r2_bench_31.err; tail -c 400 $O/r2_bench_31.err
python bench.py --variant online --no-extra --no-cpu-baseline > $O/r2_bench_31_online.json 2>> $O/r2_bench_31.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_31_reference.json 2>> $O/r2_bench_31.err
python tools/bench_ipdnet2.py cfg5 > $O/r2_bench_ipdnet2_31.jsonl 2> $O/r2_bench_ipdnet2_31.err
python tools/bench_extra.py ipdnet > $O/r2_extra_31.jsonl 2>&1
FNSSL_TC_WAIT_TIMEOUT=1 timeout 300 python tools/lstm_time.py > $O/r2_time_v31.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lstm_tc5 -s 6 -c 4 -o $O/r2_prof_v31_cfg4_b256 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > $O/r2_ncu_31a.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2_launches_v31_cfg4_b256.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $O/r2_ncu_31b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lstm_tc6 -c 2 -o $O/r2_prof_v31_tc6 python tools/lstm_time.py H256x1_add_b15 > $O/r2_ncu_31c.log 2>&1
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_31.json"))
print("cfg4", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], [(k["kernel"], k["avg_ms"], k["tflops"]) for k in d["kernels"]])
print("torch", json.dumps(d.get("gpu_torch_baseline"))[:400])
for k,v in (d.get("extra") or {}).items(): print(k, v["value"], v["ms_per_step"], v["e2e"], [(x["kernel"], x["avg_ms"]) for x in v["kernels"]])
o=json.load(open("gpurun_out/r2_bench_31_online.json")); print("online", o["value"], o["ms_per_step"], [(k["kernel"], k["avg_ms"], k["tflops"]) for k in o["kernels"]])
print(open("gpurun_out/r2_bench_31_reference.json").read()[:500])
for l in open("gpurun_out/r2_bench_ipdnet2_31.jsonl"):
    dd=json.loads(l); print(dd["workload"][:40], dd["ms_per_step"], dd["frames_per_s"])
print(open("gpurun_out/r2_extra_31.jsonl").read()[:600])
PY
