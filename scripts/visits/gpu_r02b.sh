#!/bin/bash
TAG=${1:-r02b}
mkdir -p gpurun_out
timeout 300 python scripts/diag_fp32_step.py > gpurun_out/${TAG}_diag_fp32.log 2>&1; echo "diag rc=$?"; cat gpurun_out/${TAG}_diag_fp32.log | tail -12
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -30 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_c3_f32.json 2> gpurun_out/${TAG}_bench_c3_f32.err; echo "bench f32 rc=$?"
timeout 600 python bench.py --workload C3r8 --steps 50 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_bench_c3r8_f32.json 2> gpurun_out/${TAG}_bench_c3r8_f32.err; echo "bench c3r8 rc=$?"
python - <<PY
import json
for f in ("c3","c3r8"):
    try:
        d=json.loads([l for l in open("gpurun_out/${TAG}_bench_%s_f32.json"%f) if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"] and d["e2e"]["value"], d["check"]["kl_raw"], d["check"]["bisect_its_H"])
    except Exception as e: print(f, "parse failed", e)
PY
