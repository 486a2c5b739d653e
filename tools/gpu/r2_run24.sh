# round 2, call 24: full GPU suite + smoke on the final library (pair kernels on by default), sanitizer on the pair kernels
cd $GRAFT_REPO_ROOT; O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $O/r2_gputests_v24.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $O/r2_smoke_v24.log
for tool in racecheck synccheck; do
  for c in pair128 pair128f pair256; do
    timeout 300 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py $c > $O/r2_sanitizer_${tool}_${c}.log 2>&1
    echo "$tool $c rc=$?"; tail -4 $O/r2_sanitizer_${tool}_${c}.log
  done
done
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_cases.py pair128 pair256 > $O/r2_sanitizer_memcheck_pair.log 2>&1; tail -4 $O/r2_sanitizer_memcheck_pair.log
