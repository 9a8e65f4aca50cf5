#!/bin/bash
# fp64 H pass compiled for 1 CTA / SM (library built with -DESPM_H_OCC_F64=1): f64 bench line + the fp64 parity tests
TAG=${1:-r02ab}
mkdir -p gpurun_out
timeout 600 python bench.py --dtype f64 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_c3_f64_occ1.json 2>/dev/null
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/${TAG}_bench_c3_f64_occ1.json") if l.startswith("{")][-1]); r=d["roofline"]
print("f64 occ1: it/s %.1f ms %.4f h %.4f w %.4f"%(d["value"], d["ms_per_step"], r["h_pass_ms"], r["w_pass_ms"]), r["kernel_ms"])
PY
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_log2_table.py -m gpu -q -x 2>&1 | tail -3
