"""World-size-2 (and 3) gloo tests of the pixel-sharding host logic (espm_b200/dist.py) on CPU tensors.

The kernels need a GPU, the collectives around them do not: the same ``Shard`` class drives NCCL on the
B200 box and gloo here.  Each test spawns one process per rank, rendezvous on 127.0.0.1.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, fn_name, out_dir):
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        globals()[fn_name](rank, world)
        open(os.path.join(out_dir, "ok%d" % rank), "w").close()
    finally:
        dist.destroy_process_group()


def _run(fn_name, world, tmp_path):
    mp.spawn(_worker, args=(world, _free_port(), fn_name, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / ("ok%d" % r)).exists(), "rank %d did not finish" % r


# ------------------------------------------------------------------------------------------ bodies
def _body_halo(rank, world):
    from espm_b200.dist import Shard, shard_bounds
    nx, ny, k = 7, 5, 3
    p = nx * ny
    Hfull = torch.arange(k * p, dtype=torch.float64).reshape(k, p) + 1.0
    j0, j1, row0 = shard_bounds(p, nx, ny, rank, world)
    halo, p_loc = 8, j1 - j0
    buf = torch.full((k, halo + p_loc + halo), -7.0, dtype=torch.float64)
    buf[:, halo:halo + p_loc] = Hfull[:, j0:j1]
    Shard().exchange_halo(buf, halo, p_loc, ny)
    if rank > 0:        # the image row above my first row
        assert torch.equal(buf[:, halo - ny:halo], Hfull[:, j0 - ny:j0])
    else:               # Neumann edge: untouched
        assert torch.all(buf[:, :halo] == -7.0)
    if rank < world - 1:
        assert torch.equal(buf[:, halo + p_loc:halo + p_loc + ny], Hfull[:, j1:j1 + ny])
    else:
        assert torch.all(buf[:, halo + p_loc:] == -7.0)
    assert torch.equal(buf[:, halo:halo + p_loc], Hfull[:, j0:j1])


def _body_masks_and_stats(rank, world):
    from espm_b200.dist import Shard
    sh = Shard()
    mask = torch.zeros(4 * world, dtype=torch.int32)
    mask[0] = 1 << rank
    mask[1] = 0x10 if rank == world - 1 else 0
    mask[3] = -2147483648 if rank == 0 else 0          # top bit survives the signed OR
    sh.gather_masks(mask)
    assert int(mask[0]) == (1 << world) - 1
    assert int(mask[1]) == 0x10 and int(mask[2]) == 0 and int(mask[3]) == -2147483648
    kp = 4
    hstats = torch.zeros(3 * kp, dtype=torch.float64)
    hstats[:kp] = rank + 1.0
    hstats[kp:2 * kp] = 10.0 * (rank + 1)
    hstats[2 * kp:] = torch.tensor([rank, -rank, 5.0, rank * 2.0], dtype=torch.float64)
    sh.allreduce_hstats(hstats, kp)
    tot = world * (world + 1) / 2
    assert torch.all(hstats[:kp] == tot) and torch.all(hstats[kp:2 * kp] == 10 * tot)
    assert hstats[2 * kp:].tolist() == [world - 1.0, 0.0, 5.0, 2.0 * (world - 1)]
    s = torch.full((6, 4), float(rank + 1), dtype=torch.float64)
    sh.allreduce_sum(s)
    assert torch.all(s == tot)


def _body_records_and_gather(rank, world):
    from espm_b200 import _lib as L
    from espm_b200.dist import Shard, shard_bounds
    sh = Shard()
    rec = np.zeros((3, L.NSCALARS))
    rec[:, L.S_XLOGY] = rank + 1.0
    rec[:, L.S_LOGREG] = 0.5 * (rank + 1)
    rec[:, L.S_LAPL] = 2.0
    rec[:, L.S_REL_H] = [0.1 * rank, 0.3, 0.2 * (world - rank)]
    rec[:, L.S_SUMY] = 42.0                      # replicated: taken from rank 0, not summed
    rec[:, L.S_DEV_FLAGS] = float(1 << rank)
    out = sh.combine_records(rec)
    tot = world * (world + 1) / 2
    assert np.all(out[:, L.S_XLOGY] == tot) and np.all(out[:, L.S_LOGREG] == 0.5 * tot)
    assert np.all(out[:, L.S_LAPL] == 2.0 * world) and np.all(out[:, L.S_SUMY] == 42.0)
    assert np.allclose(out[:, L.S_REL_H], [0.1 * (world - 1), 0.3, 0.2 * world])
    assert np.all(out[:, L.S_DEV_FLAGS] == float((1 << world) - 1))
    # ragged H shards (image rows do not divide evenly)
    nx, ny, k = 5, 3, 2
    p = nx * ny
    Hfull = np.arange(k * p, dtype=np.float64).reshape(k, p)
    j0, j1, _ = shard_bounds(p, nx, ny, rank, world)
    got = sh.gather_H(Hfull[:, j0:j1], p)
    assert np.array_equal(got, Hfull)


# ------------------------------------------------------------------------------------------ tests
@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange(world, tmp_path):
    _run("_body_halo", world, tmp_path)


def test_masks_stats_and_sums_world2(tmp_path):
    _run("_body_masks_and_stats", 2, tmp_path)


@pytest.mark.parametrize("world", [2, 3])
def test_records_and_gather(world, tmp_path):
    _run("_body_records_and_gather", world, tmp_path)


def test_shard_bounds_cover_image_rows():
    from espm_b200.dist import shard_bounds
    for nx, ny, world in ((512, 512, 8), (80, 80, 3), (7, 5, 7), (9, 4, 2)):
        p = nx * ny
        prev = 0
        for r in range(world):
            j0, j1, row0 = shard_bounds(p, nx, ny, r, world)
            assert j0 == prev and j0 == row0 * ny and (j1 - j0) % ny == 0 and j1 > j0
            prev = j1
        assert prev == p
    # no image shape: plain pixel ranges
    assert [shard_bounds(10, 0, 0, r, 3)[:2] for r in range(3)] == [(0, 4), (4, 7), (7, 10)]
    with pytest.raises(ValueError):
        shard_bounds(12, 3, 4, 0, 4)


def _body_sharded_init(rank, world):
    """NNDSVD initialisation against a pixel-SHARDED X (init_device.py): local products, all-gathered tall panels,
    all-reduced n x r results.  CPU tensors + gloo stand in for the device + NCCL; the reference is scikit-learn's
    _initialize_nmf on the whole matrix (updates.py:179)."""
    import types
    from sklearn.decomposition._nmf import _initialize_nmf
    from espm_b200 import _lib as L
    from espm_b200.dist import Shard, shard_bounds
    from espm_b200.init_device import initialize_nmf_device
    rng = np.random.default_rng(5)
    n, nx, ny, k, m = 96, 21, 24, 3, 7            # 21 image rows: ragged shards
    p = nx * ny
    x = np.linspace(0, 1, n)
    G = np.stack([np.exp(-0.5 * ((x - c) / 0.04) ** 2) for c in np.linspace(0.1, 0.9, m - 2)]
                 + [np.exp(-3 * x) + 0.05, (1 - x) * 0.5 + 0.05], axis=1)
    Ht = rng.uniform(size=(k, p)) ** 3
    lam = G @ (rng.uniform(size=(m, k)) ** 2) @ (Ht / Ht.sum(0, keepdims=True))
    X = rng.poisson(lam / lam.sum(0, keepdims=True) * 60.0).astype(np.float64) + 1e-14
    j0, j1, _ = shard_bounds(p, nx, ny, rank, world)
    p_loc = j1 - j0
    n_pad, nt = (n + 31) // 32 * 32, (p_loc + 127) // 128
    Xt = np.zeros((nt, n_pad, 128))
    for t in range(nt):
        w = min(128, p_loc - t * 128)
        Xt[t, :n, :w] = X[:, j0 + t * 128:j0 + t * 128 + w]
    eng = types.SimpleNamespace(n=n, p=p, p_loc=p_loc, j0=j0, shard=Shard(), Xt=torch.from_numpy(Xt.reshape(-1)),
                                x_code=L.F64, c_code=L.F64, st=types.SimpleNamespace(n_pad=n_pad, n_tiles=nt))
    for init in (None, "nndsvd", "random"):
        Wd, Hd = initialize_nmf_device(eng, k, init, random_state=3)
        Ws, Hs = _initialize_nmf(X, k, init=init, random_state=3)
        assert Wd.shape == Ws.shape and Hd.shape == Hs.shape
        assert np.max(np.abs(Wd - Ws)) <= 1e-10 * np.abs(Ws).max(), init
        assert np.max(np.abs(Hd - Hs)) <= 1e-10 * np.abs(Hs).max(), init


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_nndsvd_init(tmp_path, world):
    _run("_body_sharded_init", world, tmp_path)


def _body_gather_device(rank, world):
    """Shard.gather_H_device (one padded all-gather + one copy; sizes from shard_bounds, no collective) against the
    original gather_H on ragged shards."""
    from espm_b200.dist import Shard, shard_bounds
    nx, ny, k = 11, 6, 3
    p = nx * ny
    Hfull = torch.arange(k * p, dtype=torch.float32).reshape(k, p) * 0.5 + 1.0
    sizes = []
    for r in range(world):
        a, b, _ = shard_bounds(p, nx, ny, r, world)
        sizes.append(b - a)
    j0, j1, _ = shard_bounds(p, nx, ny, rank, world)
    sh = Shard()
    out = sh.gather_H_device(Hfull[:, j0:j1].contiguous(), sizes)
    ref = sh.gather_H(Hfull[:, j0:j1].numpy(), p)
    assert out.shape == (k, p) and np.array_equal(out, Hfull.numpy()) and np.array_equal(out, ref)


@pytest.mark.parametrize("world", [2, 3])
def test_gather_h_device(tmp_path, world):
    _run("_body_gather_device", world, tmp_path)
