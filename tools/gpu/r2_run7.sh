cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
export FNSSL_TC_WAIT_TIMEOUT=1
export FNSSL_B200_LIB=$GRAFT_REPO_ROOT/fn_ssl_b200/variants/libfnssl_b200_exp.so
for d in 0 2 4 16 18 22; do echo "== FNSSL_TC_DEBUG=$d"; FNSSL_TC_DEBUG=$d timeout 300 python tools/lstm_time.py full_in256_H128x2_add narrow_in256_H128x2 full_in16_H128x2 narrow_in256_H256 2>&1 | grep -v "_B"; done
unset FNSSL_B200_LIB
( time python -m pytest tests -m gpu -x -q ) > $O/r2_gputests_7.log 2>&1; tail -4 $O/r2_gputests_7.log
unset FNSSL_TC_WAIT_TIMEOUT
python bench.py --no-cpu-baseline > $O/r2_bench_7.json 2> $O/r2_bench_7.err; tail -c 600 $O/r2_bench_7.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_7.json"))
print("cfg4", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], [(k["kernel"], k["avg_ms"], k["tflops"]) for k in d["kernels"]])
for k,v in (d.get("extra") or {}).items(): print(k, v["value"], v["ms_per_step"], v["e2e"], [(x["kernel"], x["avg_ms"]) for x in v["kernels"]])
PY
