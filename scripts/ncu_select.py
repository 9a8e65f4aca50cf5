"""Reduce an ncu --set full capture to the rows quoted in profiles/*_summary.md.

    python scripts/ncu_select.py gpurun_out/TAG_full.ncu-rep profiles/TAG_full_selected.csv

Reads the report with `ncu -i ... --page raw --csv` (one row per launch, one column per metric) and writes a
transposed table (one row per metric, one column per launch) restricted to the metric families below.
"""
import csv
import io
import subprocess
import sys

KEEP = ("dram__bytes", "gpu__time_duration", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared", "launch__", "lts__t_sector_hit_rate",
        "sm__inst_executed_pipe_", "sm__pipe_", "sm__warps_active", "sm__throughput", "smsp__inst_executed.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__warp_issue_stalled", "smsp__average_warp",
        "gpu__compute_memory_throughput", "lts__t_bytes.sum", "l1tex__throughput", "smsp__issue_active")


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    body = [r for r in rows[2:] if len(r) == len(hdr)]          # rows[1] holds the units
    units = rows[1]
    with open(out, "w", newline="") as fh:
        w = csv.writer(fh)
        for ci, name in enumerate(hdr):
            if name in ("Kernel Name", "Block Size", "Grid Size") or name.startswith(KEEP):
                w.writerow([name, units[ci]] + [r[ci] for r in body])
    print("wrote %s: %d launches" % (out, len(body)))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
