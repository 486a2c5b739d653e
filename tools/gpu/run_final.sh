set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
( time python -m pytest tests -m gpu -x -q ) > $O/gputests_v8.log 2>&1
python __graft_entry__.py smoke > $O/smoke_v8.log 2>&1
python bench.py > $O/bench_v8_offline.json 2> $O/bench_v8_offline.err
python bench.py --variant online --no-cpu-baseline > $O/bench_v8_online.json 2> $O/bench_v8_online.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_v8_reference_arm.json 2> $O/bench_v8_reference_arm.err
python tools/bench_ipdnet2.py cfg5 default --cpu > $O/bench_ipdnet2_v8.jsonl 2> $O/bench_ipdnet2_v8.err
python tools/bench_extra.py ipdnet fnssl_b64 fnssl_b64_online fnssl_b4 stream > $O/extra_v8.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_v8.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_v8_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lstm_tc4 -s 10 -c 3 -o $O/prof_tc4_v8 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_v8_b.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 60 --csv --log-file $O/launches_ipdnet2_v8.csv python tools/bench_ipdnet2.py cfg5 > $O/ncu_v8_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sn_freq_kernel -s 0 -c 1 -o $O/prof_ipdnet2_freq_v8 python tools/bench_ipdnet2.py default > $O/ncu_v8_d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sn_time_kernel -s 0 -c 1 -o $O/prof_ipdnet2_time_v8 python tools/bench_ipdnet2.py default > $O/ncu_v8_e.log 2>&1
tail -4 $O/gputests_v8.log; tail -3 $O/smoke_v8.log
python - <<'PY'
import json
for f in ("offline","online"):
    d=json.load(open(f"gpurun_out/bench_v8_{f}.json"))
    print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"] if d.get("roofline") else None, [(k["kernel"], k["avg_ms"]) for k in d["kernels"]])
print(open("gpurun_out/bench_v8_reference_arm.json").read()[:300])
print(open("gpurun_out/extra_v8.jsonl").read())
for l in open("gpurun_out/bench_ipdnet2_v8.jsonl"):
    d=json.loads(l); print(d["workload"][:40], d["ms_per_step"], d["frames_per_s"], {k:v["avg_ms"] for k,v in d["kernels"].items()}, d.get("cpu_baseline",{}).get("value"))
PY
