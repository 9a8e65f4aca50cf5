#!/bin/bash
TAG=${1:-r02aa}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "2" > gpurun_out/${TAG}_pytest_multi.log 2>&1; echo "pytest multi rc=$?" | tee -a gpurun_out/${TAG}_pytest_multi.log; tail -4 gpurun_out/${TAG}_pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/e2e_stages.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -20 | tee gpurun_out/${TAG}_e2e_stages_n2.log
