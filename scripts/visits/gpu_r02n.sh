#!/bin/bash
# 2-GPU visit: whole 1-GPU suite, sharded parity at world 2, timelines (N=1 shard size, N=2), bench N=2, bench N=1
TAG=${1:-r02n}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log; tail -4 gpurun_out/${TAG}_pytest.log
python scripts/timeline.py --workload C3 --out gpurun_out/${TAG}_timeline_c3_n1.json 2>&1 | tail -2
python scripts/timeline.py --workload C3r8 --steps 50 --out gpurun_out/${TAG}_timeline_c3r8_n1.json 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/timeline.py --workload C3 --out gpurun_out/${TAG}_timeline_c3_n2.json 2>&1 | tail -3
bash scripts/gpu_multi.sh ${TAG} "" "2"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_c3_f32.json 2> gpurun_out/${TAG}_bench_c3_f32.err
timeout 600 python bench.py --workload C3r8 --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-compact > gpurun_out/${TAG}_bench_c3r8_f32.json 2> gpurun_out/${TAG}_bench_c3r8_f32.err
python - <<PY
import json
for f in ("c3_f32","c3r8_f32"):
    try:
        d=json.loads([l for l in open("gpurun_out/${TAG}_bench_%s.json"%f) if l.startswith("{")][-1])
        print(f, "it/s %.1f ms %.4f"%(d["value"], d["ms_per_step"]), {k:round(v,4) for k,v in d["roofline"]["kernel_ms"].items()}, "frac", round(d["roofline"]["frac"],3), "e2e", d["e2e"] and round(d["e2e"]["value"],1), "kl", d["check"]["kl_raw"], d["check"]["bisect_its_H"])
    except Exception as e: print(f, "parse failed", e)
PY
