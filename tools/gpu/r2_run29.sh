cd $GRAFT_REPO_ROOT; O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/r2_gputests_v29.log
export FNSSL_TC_WAIT_TIMEOUT=1
V=$PWD/fn_ssl_b200/variants/libfnssl_b200_clus.so
for lib in "" $V; do
  echo "lib=[$lib]" | tee -a $O/r2_scope_time_29.log
  FNSSL_B200_LIB=$lib timeout 300 python tools/lstm_time.py in16_H128x2 H128x2_add H256x1_add 2>&1 | grep -v "_b15\|_b60\|_b256\|_B" | tee -a $O/r2_scope_time_29.log
done
unset FNSSL_TC_WAIT_TIMEOUT
timeout 300 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline 2>&1 | tail -1 | tee $O/r2_bench_29_cfg4.json
for tool in racecheck synccheck; do
  for c in lstm128 lstm256n pair128 pair256; do
    timeout 300 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py $c > $O/r2_sanitizer29_${tool}_${c}.log 2>&1
    echo "$tool $c rc=$?"; tail -3 $O/r2_sanitizer29_${tool}_${c}.log
  done
done
