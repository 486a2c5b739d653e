cd $GRAFT_REPO_ROOT; O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -x -q -rs ) > $O/r2_gputests_37.log 2>&1; tail -8 $O/r2_gputests_37.log
