#!/usr/bin/env python
"""Where one iteration's time goes on every rank (ESPM_FLAG_TIMING): the small kernels leave %globaltimer stamps in
the scalar records; this script runs the bench workload for a few iterations and prints, per rank, the mean length of
every segment between two stamps.  Works on one GPU and under torch.distributed.run.

    python scripts/timeline.py [--workload C3] [--steps 30] [--warmup 5] [--out gpurun_out/timeline.json]

Segments (record slots ESPM_S_T0 + i):
    mask_wait    h_apply start        -> lock-step count known (all ranks' trace masks seen)
    apply+wpass  count known          -> w_finish start  (rest of h_apply, the whole W pass)
    wf_rows      w_finish start       -> this CTA's row of G^T S pushed
    wf_wait      rows pushed          -> every rank's flag seen (or the grid barrier on one GPU)
    wf_rest      flags seen           -> w_finish end (W', G W', column sums)
    hpass        w_finish end         -> h_finish start  (the whole H pass)
    hfinish      h_finish start       -> h_finish end
    gap          h_finish end         -> next h_apply start
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C3")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--seed", type=int, default=93)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import bench
    import espm_b200
    from espm_b200 import _lib as L, synth
    from espm_b200.engine import FitEngine
    wl = dict(bench.WORKLOADS[args.workload])
    nx, ny, n, k = wl["nx"], wl["ny"], wl["n"], wl["k"]
    p = nx * ny
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    espm_b200.config.x_storage = "dense"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    prob = synth.make_problem(nx, ny, n, k, wl["n_elements"], seed=args.seed)
    shard = None
    j0, j1 = 0, p
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        from espm_b200.dist import make_shard, shard_bounds
        shard = make_shard()
        j0, j1, _ = shard_bounds(p, nx, ny, rank, world)
    X = synth.poisson_X_torch(prob, j0, j1, args.seed, dev, torch.float32)
    G = prob["G_full"].astype(np.float32)
    W0, H0 = synth.init_factors(G.shape[1], k, p, args.seed, dtype=np.float32)
    Wm, K = args.warmup, args.steps
    eng = FitEngine(X, G, W0, H0, shape_2d=(nx, ny), max_records=Wm + K + 16, shard=shard, x_local=True, tol=0.0,
                    **wl["kw"])
    eng.st.flags |= L.FLAG_TIMING
    eng.evaluate(0)
    eng.run_iterations(1, Wm)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.run_iterations(Wm + 1, K)
    e1.record()
    torch.cuda.synchronize()
    ms_iter = e0.elapsed_time(e1) / K
    t = eng.rec_np[Wm + 1: Wm + K + 1, L.S_T0: L.S_T0 + 8].copy()         # (K, 8) ns
    names = ["mask_wait", "apply+wpass", "wf_rows", "wf_wait", "wf_rest", "hpass", "hfinish", "gap"]
    seg = np.zeros((K - 1, 8))
    for i in range(7):
        seg[:, i] = (t[:-1, i + 1] - t[:-1, i]) * 1e-3                   # us
    seg[:, 7] = (t[1:, 0] - t[:-1, 7]) * 1e-3
    mean = seg.mean(0)
    med = np.median(seg, 0)
    row = torch.tensor(np.concatenate([mean, med, [ms_iter * 1e3]]), dtype=torch.float64, device=dev)
    rows = [row]
    if world > 1:
        rows = [torch.zeros_like(row) for _ in range(world)]
        dist.all_gather(rows, row)
    if rank == 0:
        out = {"workload": args.workload, "world": world, "steps": K, "segments_us": names, "ranks": []}
        print("%-5s " % "rank" + " ".join("%12s" % s for s in names) + "   sum_us  iter_us(events)")
        for r, v in enumerate(rows):
            v = v.cpu().numpy()
            print("%-5d " % r + " ".join("%12.1f" % x for x in v[:8]) + "   %6.1f  %6.1f" % (v[:8].sum(), v[16]))
            out["ranks"].append({"rank": r, "mean_us": v[:8].tolist(), "median_us": v[8:16].tolist(),
                                 "iter_us_events": float(v[16])})
        if args.out:
            with open(args.out, "w") as fh:
                json.dump(out, fh, indent=1)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
