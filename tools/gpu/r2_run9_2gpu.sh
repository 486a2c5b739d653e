cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_parity.py -q -k "non_current_device" -rs 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2_bench_9_2gpu.json 2> $O/r2_bench_9_2gpu.err; tail -c 400 $O/r2_bench_9_2gpu.err
head -c 300 $O/r2_bench_9_2gpu.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > $O/r2_bench_9_2gpu_ref.json 2>/dev/null; head -c 200 $O/r2_bench_9_2gpu_ref.json; echo
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $O/r2_bench_9_1gpu.json 2>/dev/null
python - <<'PY'
import json
for f in ("r2_bench_9_1gpu","r2_bench_9_2gpu"):
    d=json.load(open(f"gpurun_out/{f}.json"))
    print(f, d["n_gpus"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["scaling"], d["config"]["global_batch"], d["run"], [(k["kernel"], k["avg_ms"]) for k in d["kernels"][:4]])
PY
