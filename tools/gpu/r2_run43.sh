cd $GRAFT_REPO_ROOT; O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -x -q -rs ) > $O/r2_gputests_43.log 2>&1; tail -5 $O/r2_gputests_43.log
V=$PWD/fn_ssl_b200/variants/libfnssl_b200_q4.so
for lib in "" $V; do
  echo "lib=[$lib]" | tee -a $O/r2_store_time_43.log
  FNSSL_B200_LIB=$lib timeout 300 python tools/lstm_time.py H128x2_b256 H128x2_add_b256 2>&1 | tee -a $O/r2_store_time_43.log
  FNSSL_B200_LIB=$lib timeout 300 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cfg4', d['value'], d['ms_per_step'], [(k['kernel'], k['avg_ms']) for k in d['kernels']])" | tee -a $O/r2_store_time_43.log
done
