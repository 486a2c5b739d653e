cd $GRAFT_REPO_ROOT; O=gpurun_out; mkdir -p $O
python __graft_entry__.py smoke > $O/r2_smoke_50.log 2>&1; tail -6 $O/r2_smoke_50.log
