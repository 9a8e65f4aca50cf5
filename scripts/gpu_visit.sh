#!/bin/bash
# One 1-GPU visit: parity tests, the C3 bench line (dense + compact + e2e) and the shard-size line.
# usage: scripts/gpu_visit.sh TAG [pytest-args]
TAG=${1:-rXX}
shift
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q "$@" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -25 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_c3_f32.json 2> gpurun_out/${TAG}_bench_c3_f32.err; echo "bench f32 rc=$?"; tail -3 gpurun_out/${TAG}_bench_c3_f32.err
timeout 600 python bench.py --workload C3r8 --steps 50 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_bench_c3r8_f32.json 2> gpurun_out/${TAG}_bench_c3r8_f32.err; echo "bench c3r8 rc=$?"
python - <<PY
import json
for f in ("c3","c3r8"):
    try:
        d=json.loads([l for l in open("gpurun_out/${TAG}_bench_%s_f32.json"%f) if l.startswith("{")][-1])
        print(f, "it/s %.1f ms %.4f"%(d["value"], d["ms_per_step"]), {k:round(v,4) for k,v in d["roofline"]["kernel_ms"].items()}, "e2e", d["e2e"] and round(d["e2e"]["value"],1), "kl", d["check"]["kl_raw"], d["check"]["bisect_its_H"])
        print("   compact", d.get("compact"))
    except Exception as e: print(f, "parse failed", e)
PY
