#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): 2-GPU parity tests, then the bench at 2/4/8 ranks.
TAG=${1:-rXX}
NMAX=${2:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/${TAG}_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/${TAG}_pytest_multi.log
for N in 2 4 8; do
  if [ $N -le $NMAX ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/${TAG}_bench_c3_f32_n$N.json 2> gpurun_out/${TAG}_bench_c3_f32_n$N.err; echo "bench N=$N rc=$?"
    python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/${TAG}_bench_c3_f32_n$N.json") if l.startswith("{")][-1])
    print("N=$N value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"] and d["e2e"]["value"], "kernel_ms", d["roofline"]["kernel_ms"])
except Exception as e:
    print("N=$N parse failed", e)
PY
  fi
done
