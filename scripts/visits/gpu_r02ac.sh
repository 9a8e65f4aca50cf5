#!/bin/bash
TAG=${1:-r02ac}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_c3_f32.json 2> gpurun_out/${TAG}_bench_c3_f32.err; echo "bench rc=$?"; tail -2 gpurun_out/${TAG}_bench_c3_f32.err
timeout 600 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-compact > gpurun_out/${TAG}_bench_c3_f32_k4.json 2> gpurun_out/${TAG}_bench_c3_f32_k4.err; echo "bench k4 rc=$?"; tail -2 gpurun_out/${TAG}_bench_c3_f32_k4.err
timeout 900 python bench.py --workload C5 --steps 5 --warmup 3 --no-e2e > gpurun_out/${TAG}_bench_C5_n1.json 2> gpurun_out/${TAG}_bench_C5_n1.err; echo "bench C5 rc=$?"; tail -2 gpurun_out/${TAG}_bench_C5_n1.err
python - <<PY
import json
for f in ("c3_f32","c3_f32_k4","C5_n1"):
    try:
        d=json.loads([l for l in open("gpurun_out/${TAG}_bench_%s.json"%f) if l.startswith("{")][-1]); r=d["roofline"]
        print(f, "it/s %.1f ms %.4f"%(d["value"], d["ms_per_step"]), "h %.4f w %.4f frac %.3f"%(r["h_pass_ms"], r["w_pass_ms"], r["frac"]), r.get("launches_timed"), "e2e", d["e2e"] and round(d["e2e"]["value"],1), d["clocks"])
    except Exception as e: print(f, "parse failed", e)
PY
