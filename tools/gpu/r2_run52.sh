cd $GRAFT_REPO_ROOT; O=gpurun_out; mkdir -p $O
FNSSL_TC_WAIT_TIMEOUT=1 timeout 200 python -m pytest tests/test_training_backward.py -q -m gpu -x --tb=short > $O/r2_train_bwd_52.log 2>&1; echo "rc=$?"; tail -40 $O/r2_train_bwd_52.log
