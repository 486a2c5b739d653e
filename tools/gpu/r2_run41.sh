cd $GRAFT_REPO_ROOT; O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -x -q -rs ) > $O/r2_gputests_41.log 2>&1; tail -6 $O/r2_gputests_41.log
python __graft_entry__.py smoke > $O/r2_smoke_41.log 2>&1; tail -3 $O/r2_smoke_41.log
for tool in racecheck synccheck; do
  timeout 300 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py pair128n6 > $O/r2_sanitizer41_${tool}_pair128n6.log 2>&1
  echo "$tool pair128n6 rc=$?"; tail -3 $O/r2_sanitizer41_${tool}_pair128n6.log
done
