# round 2, last call: full GPU suite on the shipped library, the candidate weight-gradient kernel (FNSSL_TRAIN_DW=2), training bench
cd $GRAFT_REPO_ROOT; O=gpurun_out; mkdir -p $O
( time FNSSL_TC_WAIT_TIMEOUT=1 timeout 85 python -m pytest tests -m gpu -q -rs --tb=short ) > $O/r2_gputests_55.log 2>&1; echo "suite rc=$?"; grep -E "passed|failed|FAILED|ERROR" $O/r2_gputests_55.log | tail -12
FNSSL_TRAIN_DW=2 timeout 30 python -m pytest tests/test_training_backward.py -m gpu -q --tb=short -k "lstm_layer or training_step or fnblock_train" > $O/r2_train_bwd_55_dw2.log 2>&1; echo "dw2 rc=$?"; tail -3 $O/r2_train_bwd_55_dw2.log
timeout 40 python tools/bench_train.py 8 16 --ours-only --dw-compare --ipdnet > $O/r2_train_bench_55.jsonl 2> $O/r2_train_bench_55.err; echo "bench rc=$?"; cat $O/r2_train_bench_55.jsonl; tail -2 $O/r2_train_bench_55.err
