#!/bin/bash
# One GPU-box visit: parity tests, bench lines, ncu launch list and a full capture of the two X passes.
# usage: scripts/gpu_round.sh TAG   (outputs under gpurun_out/TAG_*)
TAG=${1:-rXX}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/${TAG}_bench_c3_f32.json 2> gpurun_out/${TAG}_bench_c3_f32.err; echo "bench f32 rc=$?"
timeout 600 python bench.py --steps 30 --warmup 5 --dtype f64 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_bench_c3_f64.json 2> gpurun_out/${TAG}_bench_c3_f64.err; echo "bench f64 rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "bench ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_c3_f32.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"h_pass|w_pass" -s 6 -c 2 -f -o gpurun_out/${TAG}_full_c3_f32 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
cat gpurun_out/${TAG}_bench_c3_f32.json | head -c 3000
# fp64 mode: full capture of the H pass (DP-pipe side of the roofline) and the other workloads' bench lines
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"h_pass|w_pass" -s 6 -c 2 -f -o gpurun_out/${TAG}_full_c3_f64 python bench.py --dtype f64 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_ncu_full_f64.log 2>&1; echo "ncu full f64 rc=$?"
for WL in C2 C5; do
  timeout 600 python bench.py --workload $WL --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_${WL}_f32.json 2> gpurun_out/${TAG}_bench_${WL}_f32.err; echo "bench $WL f32 rc=$?"
done
timeout 600 python bench.py --workload C2 --dtype f64 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_C2_f64.json 2> gpurun_out/${TAG}_bench_C2_f64.err; echo "bench C2 f64 rc=$?"
