#!/bin/bash
TAG=${1:-r02c}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q -k "full_size or fp32" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -8 gpurun_out/${TAG}_pytest.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"h_finish|h_apply|w_finish" -s 18 -c 3 -f -o gpurun_out/${TAG}_small_c3r8 python bench.py --workload C3r8 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_ncu_small.log 2>&1; echo "ncu small rc=$?"
