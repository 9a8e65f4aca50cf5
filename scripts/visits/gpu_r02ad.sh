#!/bin/bash
# ring-depth sweep of the two X passes at the 1/8-shard size (the W-pass cap of 3 was tuned at full size)
TAG=${1:-r02ad}
mkdir -p gpurun_out
for D in 2 3 4 5; do echo "== WDEPTH $D"; ESPM_B200_WDEPTH=$D python scripts/timeline.py --workload C3r8 --steps 50 2>&1 | tail -1; done 2>&1 | tee gpurun_out/${TAG}_depth_sweep.log
for D in 2 3 4; do echo "== HDEPTH $D"; ESPM_B200_HDEPTH=$D python scripts/timeline.py --workload C3r8 --steps 50 2>&1 | tail -1; done 2>&1 | tee -a gpurun_out/${TAG}_depth_sweep.log
