#!/bin/bash
# L2-pin sweep: timeline at shard size and at C3 for several sizes of the L2-resident part of X
TAG=${1:-r02t}
mkdir -p gpurun_out
for MB in 0 24 40 56 72 88 104; do
  echo "== L2PIN $MB MB"
  ESPM_B200_L2PIN_MB=$MB python scripts/timeline.py --workload C3r8 --steps 50 2>&1 | tail -1
  ESPM_B200_L2PIN_MB=$MB python scripts/timeline.py --workload C3 --steps 20 2>&1 | tail -1
done 2>&1 | tee gpurun_out/${TAG}_l2pin_sweep.log
