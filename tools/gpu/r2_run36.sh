# round 2, call 36: validation + measurements after lstm_tc6 learned the narrow second source
cd $GRAFT_REPO_ROOT; O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -x -q -rs ) > $O/r2_gputests_36.log 2>&1; tail -8 $O/r2_gputests_36.log
python __graft_entry__.py smoke > $O/r2_smoke_36.log 2>&1; tail -3 $O/r2_smoke_36.log
python bench.py > $O/r2_bench_36.json 2> $O/r2_bench_36.err; tail -c 300 $O/r2_bench_36.err
python bench.py --variant online --no-extra --no-cpu-baseline > $O/r2_bench_36_online.json 2>> $O/r2_bench_36.err
python tools/bench_extra.py ipdnet > $O/r2_extra_36.jsonl 2>&1
for tool in racecheck synccheck memcheck; do
  timeout 300 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py pair256n > $O/r2_sanitizer36_${tool}_pair256n.log 2>&1
  echo "$tool pair256n rc=$?"; tail -3 $O/r2_sanitizer36_${tool}_pair256n.log
done
ncu --set full --clock-control none --import-source on -k regex:lstm_tc6 -c 2 -o $O/r2_prof_v36_tc6_in272 python tools/lstm_time.py in256+16_H256x1_add_b15 > $O/r2_ncu_36c.log 2>&1
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_36.json"))
print("cfg4", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], [(k["kernel"], k["avg_ms"], k["tflops"]) for k in d["kernels"]])
for k,v in (d.get("extra") or {}).items(): print(k, v["value"], v["ms_per_step"], v["e2e"], [(x["kernel"], x["avg_ms"]) for x in v["kernels"]])
o=json.load(open("gpurun_out/r2_bench_36_online.json")); print("online", o["value"], o["ms_per_step"], o["e2e"]["value"], [(k["kernel"], k["avg_ms"], k["tflops"]) for k in o["kernels"]])
print(open("gpurun_out/r2_extra_36.jsonl").read()[:600])
PY
