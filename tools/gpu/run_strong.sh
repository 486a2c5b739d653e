set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
: > gpurun_out/strong_scaling.jsonl
for b in 256 128 64 32; do
  python bench.py --batch $b --steps 5 --warmup 3 --no-cpu-baseline >> gpurun_out/strong_scaling.jsonl 2>> gpurun_out/strong_scaling.err
done
python - <<'PY'
import json
for l in open("gpurun_out/strong_scaling.jsonl"):
    d=json.loads(l); print(d["config"]["global_batch"], d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"])
PY
