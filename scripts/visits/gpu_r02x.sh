#!/bin/bash
TAG=${1:-r02x}
mkdir -p gpurun_out
for V in 0 1; do
  echo "== GWRES $V"
  ESPM_B200_GWRES=$V python scripts/timeline.py --workload C3 --steps 20 2>&1 | tail -1
  ESPM_B200_GWRES=$V python scripts/timeline.py --workload C3r8 --steps 50 2>&1 | tail -1
done 2>&1 | tee gpurun_out/${TAG}_gwres.log
ESPM_B200_GWRES=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_compact.py -m gpu -q -x 2>&1 | tail -3
ESPM_B200_GWRES=1 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); r=d['roofline']; print('bench gwres=1', d['value'], r['h_pass_ms'], r['w_pass_ms'], d['compact'] and d['compact']['value'])"
