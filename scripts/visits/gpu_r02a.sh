#!/bin/bash
# Round 2, visit A (1 GPU): parity tests incl. the new ones, a C3 bench line, and a source-level ncu capture of the
# k x p / m x k kernels at the 8-rank shard size (workload C3r8 = one rank's 64 image rows of C3).
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_c3_f32.json 2> gpurun_out/${TAG}_bench_c3_f32.err; echo "bench f32 rc=$?"
timeout 600 python bench.py --workload C3r8 --steps 50 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_bench_c3r8_f32.json 2> gpurun_out/${TAG}_bench_c3r8_f32.err; echo "bench c3r8 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"h_finish|h_apply|w_finish" -s 18 -c 3 -f -o gpurun_out/${TAG}_small_c3r8 python bench.py --workload C3r8 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_ncu_small.log 2>&1; echo "ncu small rc=$?"
head -c 2500 gpurun_out/${TAG}_bench_c3_f32.json; echo
head -c 2500 gpurun_out/${TAG}_bench_c3r8_f32.json; echo
