#!/bin/bash
TAG=${1:-r02m}
mkdir -p gpurun_out
python scripts/timeline.py --workload C3 --out gpurun_out/${TAG}_timeline_c3_n1.json 2>&1 | tail -4
python scripts/timeline.py --workload C3r8 --steps 50 --out gpurun_out/${TAG}_timeline_c3r8_n1.json 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/timeline.py --workload C3 --out gpurun_out/${TAG}_timeline_c3_n2.json 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_speculation.py -m gpu -q -x 2>&1 | tail -3
