cd $GRAFT_REPO_ROOT; O=gpurun_out; mkdir -p $O
FNSSL_TC_WAIT_TIMEOUT=1 timeout 120 python -m pytest tests/test_training_backward.py -q -m gpu -x --tb=short > $O/r2_train_bwd_54.log 2>&1; rc=$?; echo "tests rc=$rc"; tail -5 $O/r2_train_bwd_54.log
if [ $rc -eq 0 ]; then
  timeout 150 python tools/bench_train.py 2 8 16 > $O/r2_train_bench_54.jsonl 2> $O/r2_train_bench_54.err; echo "bench rc=$?"; cat $O/r2_train_bench_54.jsonl; tail -3 $O/r2_train_bench_54.err
  timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2_ncu_launches_v54_train_b8.csv python tools/bench_train.py 8 --once > /dev/null 2>&1; echo "ncu rc=$?"
fi
