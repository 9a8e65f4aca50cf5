"""The reference arm of bench.py (`--impl reference`) runs on the host only: check its JSON contract here, and that the
GPU arm refuses to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          timeout=timeout, env=e, cwd=ROOT)


def test_reference_arm_json_line():
    out = _run(["--impl", "reference", "--workload", "C1", "--steps", "2", "--warmup", "1", "--cpu-rows", "8"])
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "smoothnmf_iterations_per_s" and d["unit"] == "it/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("C1:")
    cb = d["cpu_baseline"]
    # the unmodified reference where its tree is present (this container), else the oracle port (the GPU box)
    sys.path.insert(0, ROOT)
    from oracle import ref_import
    assert cb["kind"] == ("reference" if ref_import.reference_available() else "port")
    assert cb["cores"] == os.cpu_count() and cb["value"] == d["value"] and cb["sample"] and cb["blas_threads"] >= 1
    assert cb["single_thread"]["blas_threads"] == 1 and cb["single_thread"]["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    out = _run(["--impl", "reference", "--workload", "C1", "--gpus", "2", "--steps", "1"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    out = _run(["--workload", "C1", "--steps", "1", "--no-e2e", "--no-cpu-baseline"], timeout=900)
    assert out.returncode != 0
    assert "{\"metric\"" not in out.stdout


def test_event_plan_samples_every_fourth_iteration():
    """bench.event_plan: which timed iterations carry CUDA events (an event between two kernels costs their PDL overlap,
    so the launches are sampled); every plan must time both passes at least once."""
    sys.path.insert(0, ROOT)
    import bench
    mk = object
    for count in (1, 2, 3, 4, 5, 7, 8, 20, 50):
        plan = bench.event_plan(count, mk)
        assert len(plan) == count
        h = [t for t in plan if t is not None and t[2] is not None]
        w = [t for t in plan if t is not None and t[0] is not None]
        assert h and w, count
        for t in plan:
            if t is not None:
                assert (t[0] is None) == (t[1] is None) and (t[2] is None) == (t[3] is None)
    p20 = bench.event_plan(20, mk)
    assert [i for i, t in enumerate(p20) if t is not None] == [3, 7, 11, 15, 19]
    assert sum(t is not None and t[2] is not None for t in p20) == 3 and sum(t is not None and t[0] is not None for t in p20) == 2
    assert all(t is None or (t[0] is None) != (t[2] is None) for t in p20)      # one pass per sampled iteration
