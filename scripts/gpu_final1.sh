#!/bin/bash
# Final 1-GPU visit of a round: the whole GPU suite, smoke(), the driver's bench line + reference arm, the ncu launch list of
# the same command, DRAM traffic of the X passes at the shard sizes of 2 / 4 / 8-GPU fits (ncu cannot attach to a
# multi-rank run: the slab of one rank is run alone, same grid, same bytes), timelines, e2e stages.
TAG=${1:-rXX}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log; tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_c3_f32.json 2> gpurun_out/${TAG}_bench_c3_f32.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "bench ref rc=$?"
timeout 600 python bench.py --dtype f64 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_c3_f64.json 2> gpurun_out/${TAG}_bench_c3_f64.err; echo "bench f64 rc=$?"
for WL in C3r2 C3r4 C3r8; do
  timeout 600 python bench.py --workload $WL --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-compact > gpurun_out/${TAG}_bench_${WL}_f32.json 2>/dev/null
  timeout 600 ncu --set full --clock-control none -k regex:"h_pass|w_pass" -s 12 -c 2 -f -o gpurun_out/${TAG}_full_${WL} python bench.py --workload $WL --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-compact > /dev/null 2>&1; echo "ncu $WL rc=$?"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_c3_f32.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-compact > /dev/null 2>&1; echo "ncu launches rc=$?"
python scripts/timeline.py --workload C3 --out gpurun_out/${TAG}_timeline_c3_n1.json 2>&1 | tail -2
python scripts/timeline.py --workload C3r8 --steps 50 --out gpurun_out/${TAG}_timeline_c3r8_n1.json 2>&1 | tail -1
python scripts/e2e_stages.py 2>&1 | tail -18 > gpurun_out/${TAG}_e2e_stages.log
python - <<PY
import json
for f in ("c3_f32","c3_f64","C3r2_f32","C3r4_f32","C3r8_f32","ref"):
    try:
        d=json.loads([l for l in open("gpurun_out/${TAG}_bench_%s.json"%f) if l.startswith("{")][-1])
        r=d.get("roofline") or {}
        print(f, "it/s %.3f ms %.4f"%(d["value"], d["ms_per_step"]), "frac", r.get("frac"), "e2e", d["e2e"] and round(d["e2e"]["value"],1), (d.get("cpu_baseline") or {}).get("sample","")[:160])
    except Exception as e: print(f, "parse failed", e)
PY
