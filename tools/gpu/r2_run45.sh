cd $GRAFT_REPO_ROOT; O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -x -q -rs ) > $O/r2_gputests_45.log 2>&1; tail -5 $O/r2_gputests_45.log
FNSSL_TC_PAIR=0 FNSSL_TC_PAIR256=0 timeout 300 python tools/lstm_time.py in16_H128x2 H128x2_add H256x1_add 2>&1 | grep -v "_b15\|_b60\|_b256\|_B" | tee $O/r2_tc4_store_time_45.log
python bench.py --no-cpu-baseline > $O/r2_bench_45.json 2> $O/r2_bench_45.err; tail -c 300 $O/r2_bench_45.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_45.json"))
print("cfg4", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], d["gpu_launches"], [(k["kernel"], k["avg_ms"], k["tflops"]) for k in d["kernels"]], d["clocks"])
for k,v in (d.get("extra") or {}).items(): print(k, v["value"], v["ms_per_step"], v["e2e"], [(x["kernel"], x["avg_ms"]) for x in v["kernels"]])
PY
