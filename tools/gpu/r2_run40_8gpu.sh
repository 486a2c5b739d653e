# round 2: strong scaling of cfg4 (global batch 256) at N = 8 on one box
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --steps 10 --warmup 3 > $O/r2_bench_40_8gpu.json 2> $O/r2_bench_40_8gpu.err; tail -c 300 $O/r2_bench_40_8gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 bench.py --impl reference --gpus 8 --steps 1 --warmup 0 > $O/r2_bench_40_8gpu_ref.json 2>/dev/null; head -c 200 $O/r2_bench_40_8gpu_ref.json; echo
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_40_8gpu.json"))
print(d["n_gpus"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["scaling"], d["config"]["global_batch"], d.get("run"), [(k["kernel"], k["avg_ms"]) for k in d["kernels"][:4]], d["clocks"])
PY
