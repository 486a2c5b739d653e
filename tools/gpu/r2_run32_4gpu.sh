# round 2: strong scaling of cfg4 (global batch 256) at N = 4, 2, 1 on ONE box (4 GPUs visible)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L
for N in 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 10 --warmup 3 > $O/r2_bench_32_${N}gpu.json 2> $O/r2_bench_32_${N}gpu.err; tail -c 300 $O/r2_bench_32_${N}gpu.err
done
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $O/r2_bench_32_1gpu.json 2>/dev/null
python -m pytest tests/test_gpu_parity.py -q -k "non_current_device" -rs 2>&1 | tail -2
python - <<'PY'
import json
for n in (1, 2, 4):
    d=json.load(open(f"gpurun_out/r2_bench_32_{n}gpu.json"))
    print(n, d["n_gpus"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["scaling"], d["config"]["global_batch"], d.get("run"), [(k["kernel"], k["avg_ms"]) for k in d["kernels"][:4]], d["clocks"])
PY
