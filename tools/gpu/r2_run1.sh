# round 2, call 1: validation of the refactor + new bench + sanitizer logs + H=256 ncu capture
set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/r2_gpu.txt
( time python -m pytest tests -m gpu -x -q ) > $O/r2_gputests_1.log 2>&1
tail -5 $O/r2_gputests_1.log
python __graft_entry__.py smoke > $O/r2_smoke_1.log 2>&1; tail -3 $O/r2_smoke_1.log
python bench.py > $O/r2_bench_1.json 2> $O/r2_bench_1.err; tail -c 1500 $O/r2_bench_1.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_1_ref.json 2> $O/r2_bench_1_ref.err
python bench.py --workload cfg4 --variant online --no-extra --no-cpu-baseline --steps 5 > $O/r2_bench_1_cfg4_online.json 2> $O/r2_bench_1_cfg4_online.err
for c in lstm128 lstm256n conv; do
  for tool in racecheck synccheck; do
    timeout 200 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py $c > $O/r2_sanitizer_${tool}_${c}.log 2>&1
    echo "$tool $c rc=$?"; tail -4 $O/r2_sanitizer_${tool}_${c}.log
  done
done
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_cases.py lstm128 lstm256n conv > $O/r2_sanitizer_memcheck.log 2>&1; tail -4 $O/r2_sanitizer_memcheck.log
python tools/tc4_trace.py > $O/r2_tc4_trace_1.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2_launches_cfg4_b32.csv python bench.py --workload cfg4 --batch 32 --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $O/r2_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lstm_tc4 -s 6 -c 6 -o $O/r2_prof_tc4_online python bench.py --workload cfg2 --variant online --steps 1 --warmup 1 --no-cpu-baseline --no-extra > $O/r2_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lstm_tc4 -s 6 -c 6 -o $O/r2_prof_tc4_cfg4_b32 python bench.py --workload cfg4 --batch 32 --steps 1 --warmup 1 --no-cpu-baseline --no-extra > $O/r2_ncu_c.log 2>&1
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_1.json"))
print("cfg4", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], [(k["kernel"], k["avg_ms"], k["tflops"]) for k in d["kernels"]])
print("clocks", d["clocks"]); print("cpu", d["cpu_baseline"]); print("torch", json.dumps(d["gpu_torch_baseline"]))
for k,v in (d.get("extra") or {}).items(): print(k, v["value"], v["ms_per_step"], v["e2e"], [(x["kernel"], x["avg_ms"]) for x in v["kernels"]])
d=json.load(open("gpurun_out/r2_bench_1_cfg4_online.json")); print("cfg4 online", d["value"], d["ms_per_step"], [(k["kernel"], k["avg_ms"], k["tflops"]) for k in d["kernels"]])
print(open("gpurun_out/r2_bench_1_ref.json").read()[:400])
PY
cat $O/r2_tc4_trace_1.txt | head -80
