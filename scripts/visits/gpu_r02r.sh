#!/bin/bash
TAG=${1:-r02r}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log; tail -3 gpurun_out/${TAG}_pytest.log
python scripts/timeline.py --workload C3r8 --steps 50 --out gpurun_out/${TAG}_timeline_c3r8_n1.json 2>&1 | tail -2
python scripts/timeline.py --workload C3 --out gpurun_out/${TAG}_timeline_c3_n1.json 2>&1 | tail -1
python scripts/e2e_stages.py 2>&1 | tail -20 | tee gpurun_out/${TAG}_e2e_stages.log
