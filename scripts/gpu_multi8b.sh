#!/bin/bash
# 8-GPU visit (short): sharded parity at 8 ranks through peer memory, device timeline at 8 ranks, C3 at 8 and 4 ranks, C4 at 8.
TAG=${1:-rXX}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs -k "memory-8" > gpurun_out/${TAG}_pytest_multi.log 2>&1; echo "pytest multi rc=$?" | tee -a gpurun_out/${TAG}_pytest_multi.log; tail -5 gpurun_out/${TAG}_pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 scripts/timeline.py --workload C3 --out gpurun_out/${TAG}_timeline_c3_n8.json 2>&1 | tail -9
run() {  # N workload extra...
  N=$1; WL=$2; shift; shift
  OUT=gpurun_out/${TAG}_bench_${WL}_n${N}$(echo "$@" | tr -d ' -' | cut -c1-20)
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 20 --warmup 5 --workload $WL "$@" > $OUT.json 2> $OUT.err; echo "bench $WL N=$N $@ rc=$?"; tail -2 $OUT.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT.json") if l.startswith("{")][-1])
    print("   it/s %.1f ms %.4f"%(d["value"], d["ms_per_step"]), "e2e", d["e2e"] and round(d["e2e"]["value"],1), {k:round(v,4) for k,v in d["roofline"].get("kernel_ms",{}).items()}, "host", d["roofline"].get("host_enqueue_ms_per_step"))
    print("   check", {k:v for k,v in d["check"].items() if k!="what"})
    c=d.get("compact"); print("   compact", c and {k:c.get(k) for k in ("storage","value","ms_per_step")})
except Exception as e:
    print("   parse failed", e)
PY
}
run 8 C3
run 4 C3
run 2 C3 --no-compact
run 8 C4 --no-e2e --no-compact
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 scripts/timeline.py --workload C3 --out gpurun_out/${TAG}_timeline_c3_n4.json 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 scripts/e2e_stages.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -22 | tee gpurun_out/${TAG}_e2e_stages_n8.log
