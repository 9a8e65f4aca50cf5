#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): sharded parity tests for the worlds that fit, then bench at the given rank counts.
# usage: scripts/gpu_multi.sh TAG "<pytest -k expr>" "<list of N>"
TAG=${1:-rXX}
KEXPR=${2:-""}
NS=${3:-"2"}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs ${KEXPR:+-k "$KEXPR"} > gpurun_out/${TAG}_pytest_multi.log 2>&1; echo "pytest multi rc=$?" | tee -a gpurun_out/${TAG}_pytest_multi.log; tail -12 gpurun_out/${TAG}_pytest_multi.log
for N in $NS; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_c3_f32_n$N.json 2> gpurun_out/${TAG}_bench_c3_f32_n$N.err; echo "bench N=$N rc=$?"; tail -3 gpurun_out/${TAG}_bench_c3_f32_n$N.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/${TAG}_bench_c3_f32_n$N.json") if l.startswith("{")][-1])
    print("N=$N it/s %.1f ms %.4f"%(d["value"], d["ms_per_step"]), "e2e", d["e2e"] and round(d["e2e"]["value"],1), {k:round(v,4) for k,v in d["roofline"]["kernel_ms"].items()})
    print("   check", d["check"])
    print("   compact", d.get("compact") and {k:d["compact"].get(k) for k in ("storage","value","ms_per_step","W_max_rel_diff_vs_dense")})
except Exception as e:
    print("N=$N parse failed", e)
PY
done
