#!/bin/bash
# ncu evidence of round 2: launch list of the bench command, full capture of the two X passes (C3) and of the small kernels in
# steady state at the 1/8-shard size
TAG=${1:-r02w}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_c3_f32.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-compact > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"h_finish|h_apply|w_finish" -s 75 -c 3 -f -o gpurun_out/${TAG}_small_c3r8 python bench.py --workload C3r8 --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --no-compact > gpurun_out/${TAG}_ncu_small.log 2>&1; echo "ncu small rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"h_pass|w_pass" -s 12 -c 2 -f -o gpurun_out/${TAG}_full_c3_f32 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-compact > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/${TAG}_*
