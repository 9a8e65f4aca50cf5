"""Pixel sharding over the GPUs of one box (one process per GPU, torch.distributed).

The path shards over pixels: rank r owns a contiguous block of image rows, keeps
its X slab and H columns local, and exchanges per iteration only
  * one image row of H with each neighbour (Laplacian halo),
  * the 128-bit lock-step trace mask of the H bisection (OR),
  * the n x k ratio sums S of the W update and the k row statistics of H (sum / max),
  * the handful of loss scalars (sum / max), off the critical path.
All messages are < 128 KiB, i.e. latency bound; NCCL is used for them through torch.distributed.
The same class runs on CPU tensors with the gloo backend (used by the CPU tests of the host logic).
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _lib as L


def shard_bounds(p, nx, ny, rank, world):
    """Pixel range [j0, j1) and first image row of ``rank``: contiguous image rows (pixel = i*ny + j)."""
    if nx > 0 and ny > 0:
        if world > nx:
            raise ValueError("more ranks (%d) than image rows (%d)" % (world, nx))
        base, rem = divmod(nx, world)
        r0 = rank * base + min(rank, rem)
        r1 = r0 + base + (1 if rank < rem else 0)
        return r0 * ny, r1 * ny, r0
    if world > p:
        raise ValueError("more ranks (%d) than pixels (%d)" % (world, p))
    base, rem = divmod(p, world)
    j0 = rank * base + min(rank, rem)
    j1 = j0 + base + (1 if rank < rem else 0)
    return j0, j1, 0


class Shard:
    """Collectives of the sharded fit.  ``group`` is a torch.distributed process group (None = world)."""

    def __init__(self, rank=None, world=None, group=None):
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self._mask_all = None

    # ---- W update inputs -------------------------------------------------------------------------
    def allreduce_sum(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def allreduce_max_scalar(self, t):
        """Maximum over the ranks of a 0-d / 1-element float64 tensor (any device the backend accepts)."""
        dev = t.device
        v = t.reshape(1).clone()
        if dist.get_backend(self.group) == "nccl" and v.device.type != "cuda":
            v = v.cuda()
        dist.all_reduce(v, op=dist.ReduceOp.MAX, group=self.group)
        return v.to(dev)[0]

    def allreduce_max_int(self, t):
        """In-place elementwise maximum over the ranks of an integer tensor (marks: OR)."""
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)

    def allreduce_hstats(self, hstats, kp):
        """hstats = {rowsum[kp], rowsum(max(.,ls))[kp], rowmax[kp]} (float64)."""
        dist.all_reduce(hstats[:2 * kp], op=dist.ReduceOp.SUM, group=self.group)
        dist.all_reduce(hstats[2 * kp:3 * kp], op=dist.ReduceOp.MAX, group=self.group)

    # ---- X ingest (once per fit) -----------------------------------------------------------------
    def combine_ingest(self, info, row_nz):
        """info = [ESPM_X_* flags, sum of X, number of all-zero pixels] of this rank's slab (float64),
        row_nz = per-channel "has a non-zero" marks (int32).  Returns the global versions: flags OR-ed,
        sums added, marks OR-ed (a channel is an all-zero row only if it is zero on every rank)."""
        gathered = [torch.empty_like(info) for _ in range(self.world)]
        dist.all_gather(gathered, info.contiguous(), group=self.group)
        allr = torch.stack(gathered)
        flags = 0
        for v in allr[:, 0].cpu().tolist():
            flags |= int(v)
        out = torch.stack([torch.tensor(float(flags), dtype=info.dtype, device=info.device),
                           allr[:, 1].sum(), allr[:, 2].sum()])
        row = row_nz.clone()
        dist.all_reduce(row, op=dist.ReduceOp.MAX, group=self.group)
        return out, row

    # ---- lock-step bisection ---------------------------------------------------------------------
    def gather_masks(self, mask):
        """mask[:4] <- OR over ranks of the 128-bit trace masks (NCCL has no bitwise OR: gather + fold)."""
        if self._mask_all is None or self._mask_all.device != mask.device:
            self._mask_all = torch.zeros(self.world * 4, dtype=mask.dtype, device=mask.device)
        dist.all_gather_into_tensor(self._mask_all, mask[:4].contiguous(), group=self.group)
        folded = self._mask_all.view(self.world, 4)
        acc = folded[0].clone()
        for r in range(1, self.world):
            acc = torch.bitwise_or(acc, folded[r])
        mask[:4].copy_(acc)

    # ---- Laplacian halo --------------------------------------------------------------------------
    def exchange_halo(self, Hbuf, halo, p_loc, ny):
        """Hbuf: (k, ldh).  Sends the first / last owned image row to the previous / next rank and
        receives theirs into the ny elements just before / after the owned pixels."""
        ops = []
        first = Hbuf[:, halo:halo + ny].contiguous()
        last = Hbuf[:, halo + p_loc - ny:halo + p_loc].contiguous()
        recv_lo = torch.empty_like(first)
        recv_hi = torch.empty_like(last)
        if self.rank > 0:
            ops.append(dist.P2POp(dist.isend, first, self._peer(self.rank - 1), group=self.group))
            ops.append(dist.P2POp(dist.irecv, recv_lo, self._peer(self.rank - 1), group=self.group))
        if self.rank < self.world - 1:
            ops.append(dist.P2POp(dist.isend, last, self._peer(self.rank + 1), group=self.group))
            ops.append(dist.P2POp(dist.irecv, recv_hi, self._peer(self.rank + 1), group=self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        if self.rank > 0:
            Hbuf[:, halo - ny:halo] = recv_lo
        if self.rank < self.world - 1:
            Hbuf[:, halo + p_loc:halo + p_loc + ny] = recv_hi

    def _peer(self, r):
        return r if self.group is None else dist.get_global_rank(self.group, r)

    # ---- scalars / results -----------------------------------------------------------------------
    def combine_records(self, rec):
        """rec: (m, NSCALARS) local records (NumPy) -> global records, identical on every rank."""
        t = torch.from_numpy(np.ascontiguousarray(rec))
        backend = dist.get_backend(self.group)
        if backend == "nccl":
            t = t.cuda()
        gathered = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(gathered, t, group=self.group)
        allr = np.stack([g.cpu().numpy() for g in gathered])          # (world, m, NSCALARS)
        return combine_record_arrays(allr)

    def gather_H_device(self, H_dev, sizes):
        """Concatenate the H shards along pixels (rank order) from DEVICE slabs: one padded all-gather and one D2H copy.
        ``sizes``: the pixel count of every rank (known to every rank from shard_bounds, no collective needed)."""
        k, mx = H_dev.shape[0], max(sizes)
        pad = torch.zeros(k, mx, dtype=H_dev.dtype, device=H_dev.device)
        pad[:, :H_dev.shape[1]] = H_dev
        if dist.get_backend(self.group) == "nccl":
            out = torch.empty(self.world, k, mx, dtype=H_dev.dtype, device=H_dev.device)
            dist.all_gather_into_tensor(out, pad, group=self.group)
        else:                                   # gloo (CPU tests): no all_gather_into_tensor
            parts = [torch.empty_like(pad) for _ in range(self.world)]
            dist.all_gather(parts, pad, group=self.group)
            out = torch.stack(parts)
        host = out.cpu().numpy()
        return np.concatenate([host[r][:, :s] for r, s in enumerate(sizes)], axis=1)

    def gather_H(self, H_local, p):
        """Concatenate the H shards along pixels (rank order)."""
        k = H_local.shape[0]
        t = torch.from_numpy(np.ascontiguousarray(H_local))
        backend = dist.get_backend(self.group)
        sizes = [None] * self.world
        dist.all_gather_object(sizes, H_local.shape[1], group=self.group)
        mx = max(sizes)
        pad = torch.zeros(k, mx, dtype=t.dtype)
        pad[:, :t.shape[1]] = t
        if backend == "nccl":
            pad = pad.cuda()
        out = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(out, pad, group=self.group)
        return np.concatenate([o.cpu().numpy()[:, :s] for o, s in zip(out, sizes)], axis=1)


class _DevMem:
    """Read/write view of raw device memory for torch (``__cuda_array_interface__``, no ownership)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2}


def _up(a, b):
    return (a + b - 1) // b * b


# One peer region per process is kept across fits (module level): allocating it, exporting / opening the CUDA-IPC
# handles of `world` ranks and the barriers around that cost tens of milliseconds, more than a 20-iteration fit of the
# benchmark image on 8 GPUs.  Key = everything that determines the layout; a different problem shape frees the old one.
_REGION = None


class _PeerRegion:
    def __init__(self, key, region, total, bases, opened, metas, lib):
        self.key, self.region, self.total = key, region, total
        self.bases, self.opened, self.metas, self.lib = bases, opened, metas, lib
        self.seq_s = self.seq_m = 0        # exchange sequence numbers continue across fits (flags are never reset)
        self.group = None

    def free(self, group):
        """Collective: unmap the peers' regions, then free this rank's."""
        import ctypes
        torch.cuda.synchronize()
        dist.barrier(group=group)
        for p in self.opened:
            self.lib.espm_peer_close(ctypes.c_void_p(p))
        self.opened = []
        dist.barrier(group=group)
        self.lib.espm_peer_free(ctypes.c_void_p(self.region))
        self.region = None


def release_peer_memory(group=None):
    """Collective: free the cached peer region (call on every rank, e.g. before destroying the process group)."""
    global _REGION
    if _REGION is not None and _REGION.region is not None:
        _REGION.free(_REGION.group)
    _REGION = None
    release_record_inbox()


# Record inboxes of the sharded fit (see espm_state.peer_rec): one POSIX shared-memory segment per rank, opened by every
# process of the box and page-locked / device-mapped in each (espm_host_register), so that every GPU can store its share
# of a scalar record straight into every host's memory.  Cached across fits like the peer region.
_INBOX = None
_INBOX_SEQ = 0


class _RecordInbox:
    WORDS = 8          # doubles per share: sum X log Y, log-reg, Laplacian, rel_H, device flags, -, -, stamp

    def __init__(self, shard, cap):
        import ctypes
        import mmap
        import os
        global _INBOX_SEQ
        self.world, self.rank, self.cap, self.group = shard.world, shard.rank, int(cap), shard.group
        self.nbytes = self.world * self.cap * self.WORDS * 8
        self.lib = L.load()
        self.stamp = 0                     # record stamps continue across fits: a stale share never matches
        _INBOX_SEQ += 1
        path = "/dev/shm/espm_b200_inbox_%d_%d_%d" % (os.getpid(), self.rank, _INBOX_SEQ)
        self.maps, self.dev_ptrs, self.registered = [], [], []
        ok = True
        try:
            fd = os.open(path, os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
            os.ftruncate(fd, self.nbytes)            # zero-filled
            os.close(fd)
        except OSError:
            ok, path = False, None
        paths = [None] * self.world
        dist.all_gather_object(paths, path, group=self.group)
        ok = ok and all(q is not None for q in paths)
        if ok:
            try:
                for q in paths:
                    fd = os.open(q, os.O_RDWR)
                    m = mmap.mmap(fd, self.nbytes)
                    os.close(fd)
                    self.maps.append(m)
                    arr = np.frombuffer(m, dtype=np.float64)
                    d = ctypes.c_void_p()
                    L.check(self.lib.espm_host_register(ctypes.c_void_p(arr.ctypes.data), self.nbytes, ctypes.byref(d)))
                    self.registered.append(arr.ctypes.data)
                    self.dev_ptrs.append(d.value)
            except (OSError, L.EspmError, ValueError):
                ok = False
        oks = [None] * self.world
        dist.all_gather_object(oks, ok, group=self.group)      # also: everybody has opened every segment
        if path is not None:
            try:
                os.unlink(path)                      # the mappings keep the memory alive; nothing is left behind
            except OSError:
                pass
        self.ok = all(oks)
        if self.ok:
            # views[r][r, slot, :] is rank r's share of record `slot`: every GPU writes only into its OWN host's inbox
            # (slot row [rank]), every host reads the shares out of the inboxes it has mapped
            self.views = [np.frombuffer(m, dtype=np.float64).reshape(self.world, self.cap, self.WORDS) for m in self.maps]
            self.mine = self.views[self.rank]
        else:
            self.free()

    def free(self):
        import ctypes
        for ptr in self.registered:
            self.lib.espm_host_unregister(ctypes.c_void_p(ptr))
        self.registered, self.dev_ptrs = [], []
        self.mine = None
        self.views = []
        self.maps = []                     # (the mmap objects are closed by the garbage collector once unreferenced)


def release_record_inbox():
    """Collective: unmap the cached record inboxes."""
    global _INBOX
    if _INBOX is not None and _INBOX.ok:
        torch.cuda.synchronize()
        dist.barrier(group=_INBOX.group)
        _INBOX.free()
    _INBOX = None


class PeerShard(Shard):
    """Sharded fit whose per-iteration exchanges run INSIDE the kernels through CUDA-IPC peer memory over
    NVLink (ESPM_FLAG_PEER, include/espm_b200.h): no host-launched collective on the critical path.
    torch.distributed is still used for the one-off set-up (IPC handles, initial halo / statistics) and for
    gathering results."""

    use_peer = True

    def __init__(self, rank=None, world=None, group=None):
        super().__init__(rank, world, group)
        if self.world > L.MAX_RANKS:
            raise ValueError("at most %d pixel shards are supported" % L.MAX_RANKS)
        self.reg = None

    # kept for callers that address the region directly
    @property
    def bases(self):
        return self.reg.bases

    @property
    def meta(self):
        return self.reg.metas

    def _layout(self, st):
        sz = 8 if st.c_dtype == L.F64 else 4
        hs_off = _up(st.n_pad * st.kp * sz, 128)
        slot = _up(hs_off + 3 * st.kp * 8, 256)          # one source rank: S [n_pad][kp] + 3 kp statistics
        stride = _up(self.world * slot, 256)             # one parity: a slot per source rank
        x_off = L.PF_WORDS * 4
        h_bytes = _up(st.k * st.ldh * sz, 256)
        h_off = [_up(x_off + 2 * stride, 256) + i * h_bytes for i in range(3)]
        total = h_off[2] + h_bytes
        return sz, hs_off, slot, stride, x_off, h_bytes, h_off, total

    def _all_ok(self, ok):
        """Collective AND of a per-rank success flag."""
        oks = [None] * self.world
        dist.all_gather_object(oks, bool(ok), group=self.group)
        return all(oks)

    def setup_peer(self, eng):
        """Allocate (or reuse) this rank's region (flags | receive buffers | 3 H buffers), map every peer's region and
        fill the peer fields of ``eng.st``.  Returns the three H tensors (views of the region), or None -- on EVERY
        rank -- when any rank could not allocate, export or map (the caller then falls back to the NCCL exchange)."""
        import ctypes
        global _REGION
        lib = L.load()
        st = eng.st
        sz, hs_off, slot, stride, x_off, h_bytes, h_off, total = self._layout(st)
        key = (self.world, self.rank, int(total), int(st.ldh), int(st.p_loc), int(st.halo), tuple(h_off), int(x_off),
               int(slot), eng.device.index, id(self.group))
        reg = _REGION
        if reg is not None and (reg.key != key or reg.region is None):
            release_peer_memory()                    # another problem shape: collective free (every rank misses)
            reg = None
        if reg is None:
            ptr = ctypes.c_void_p()
            handle = ctypes.create_string_buffer(64)
            mine = None
            try:
                L.check(lib.espm_peer_alloc(total, ctypes.byref(ptr)))
                L.check(lib.espm_peer_export(ptr, handle))
                mine = dict(handle=handle.raw, ldh=int(st.ldh), p_loc=int(st.p_loc), h_off=h_off, x_off=x_off,
                            halo=int(st.halo))
            except L.EspmError:
                mine = None
            metas = [None] * self.world
            dist.all_gather_object(metas, mine, group=self.group)
            opened, bases, ok = [], [], all(m is not None for m in metas)
            if ok:
                try:
                    for r, m in enumerate(metas):
                        if r == self.rank:
                            bases.append(ptr.value)
                            continue
                        q = ctypes.c_void_p()
                        L.check(lib.espm_peer_open(m["handle"], ctypes.byref(q)))
                        opened.append(q.value)
                        bases.append(q.value)
                except L.EspmError:
                    ok = False
            if not self._all_ok(ok):
                # some rank cannot reach a peer (other host, no P2P): nobody uses peer memory
                for q in opened:
                    lib.espm_peer_close(ctypes.c_void_p(q))
                dist.barrier(group=self.group)
                if ptr.value:
                    lib.espm_peer_free(ptr)
                return None
            reg = _PeerRegion(key, ptr.value, total, bases, opened, metas, lib)
            reg.group = self.group
            _REGION = reg
        self.reg = reg
        st.rank, st.world = self.rank, self.world
        st.xchg_stride, st.xchg_slot, st.xchg_hs_off = stride, slot, hs_off
        for r in range(self.world):
            st.peer_flags[r] = reg.bases[r]
            st.peer_xchg[r] = reg.bases[r] + reg.metas[r]["x_off"]
        st.flags |= L.FLAG_PEER | L.FLAG_FUSED_WREDUCE
        tdt = torch.float64 if st.c_dtype == L.F64 else torch.float32
        raw = torch.as_tensor(_DevMem(reg.region, total), device=eng.device)
        H = []
        for i in range(3):
            t = raw[h_off[i]:h_off[i] + st.k * st.ldh * sz].view(tdt).view(st.k, st.ldh)
            t.fill_(1.0)
            H.append(t)
        # every region is mapped and initialised before any kernel of any rank writes into it
        torch.cuda.current_stream(eng.device).synchronize()
        dist.barrier(group=self.group)
        return H

    @property
    def seq(self):
        return self.reg.seq_s, self.reg.seq_m

    def setup_inbox(self, eng, n_slots):
        """Record inboxes for ``n_slots`` record slots (cached; grown collectively when a fit needs more).  Returns the
        inbox, or None when shared memory / host registration is not available on some rank (the records are then
        gathered with a collective at read-back, as with the NCCL exchange)."""
        global _INBOX
        box = _INBOX
        if box is not None and (not box.ok or box.cap < n_slots or box.world != self.world or box.rank != self.rank
                                or box.group is not self.group):
            release_record_inbox()
            box = None
        if box is None:
            box = _RecordInbox(self, max(int(n_slots), 1024))
            _INBOX = box
        if not box.ok:
            return None
        st = eng.st
        for r in range(self.world):
            st.peer_rec[r] = box.dev_ptrs[r]
        st.rec_cap = box.cap
        return box

    def halo_targets(self, eng, ibuf):
        """(prev pointer, prev ldh, next pointer, next ldh) for pushing the boundary rows of H buffer ``ibuf``."""
        if eng.ny <= 0:
            return 0, 0, 0, 0
        sz = 8 if eng.st.c_dtype == L.F64 else 4
        pp = pl = npn = nl = 0
        if self.rank > 0:
            m = self.meta[self.rank - 1]
            pp = self.bases[self.rank - 1] + m["h_off"][ibuf] + (m["halo"] + m["p_loc"]) * sz
            pl = m["ldh"]
        if self.rank < self.world - 1:
            m = self.meta[self.rank + 1]
            npn = self.bases[self.rank + 1] + m["h_off"][ibuf] + (m["halo"] - eng.ny) * sz
            nl = m["ldh"]
        return pp, pl, npn, nl

    def close(self, seq_s=None, seq_m=None):
        """End of a fit: every rank is done with the region (collective), which stays mapped for the next fit of the
        same shape; ``release_peer_memory()`` frees it."""
        if self.reg is None:
            return
        if seq_s is not None:
            self.reg.seq_s, self.reg.seq_m = int(seq_s), int(seq_m)
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        self.reg = None


def _peer_capable(group):
    """Collective: True on every rank iff all ranks run on one host and no visible pair of their devices lacks peer
    access.  (A pair that cannot be checked from here -- the peer's device is not visible to this process -- is
    settled by ``PeerShard.setup_peer``, whose failure is collective as well.)"""
    import socket
    world = dist.get_world_size(group)
    dev = torch.cuda.current_device()
    try:
        uuid = str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:
        uuid = None
    infos = [None] * world
    dist.all_gather_object(infos, (socket.gethostname(), uuid), group=group)
    ok = all(h == infos[0][0] for h, _ in infos)
    if ok and uuid is not None:
        local = {}
        for i in range(torch.cuda.device_count()):
            try:
                local[str(torch.cuda.get_device_properties(i).uuid)] = i
            except Exception:
                pass
        for _, u in infos:
            j = local.get(u)
            if j is not None and j != dev and not torch.cuda.can_device_access_peer(dev, j):
                ok = False
    oks = [None] * world
    dist.all_gather_object(oks, ok, group=group)
    return all(oks)


_PEER_CAPABLE = {}


def make_shard(group=None):
    """PeerShard when the process group runs NCCL, every rank sits on the same host with peer access between the
    devices and ESPM_B200_PEER != 0; else Shard (NCCL collectives).  The decision is collective; the capability check
    (two object collectives, ~2 ms at 8 ranks) is made once per process group and device."""
    import os
    if dist.get_backend(group) == "nccl" and os.environ.get("ESPM_B200_PEER", "1") != "0":
        key = (id(group) if group is not None else None, dist.get_world_size(group), torch.cuda.current_device())
        if key not in _PEER_CAPABLE:
            _PEER_CAPABLE[key] = _peer_capable(group)
        if _PEER_CAPABLE[key]:
            return PeerShard(group=group)
    return Shard(group=group)


def combine_record_arrays(allr):
    """(world, m, NSCALARS) -> (m, NSCALARS): sums for the additive loss parts, max for rel_H, OR for
    flags; everything else is replicated and taken from rank 0."""
    out = allr[0].copy()
    for s in (L.S_XLOGY, L.S_LOGREG, L.S_LAPL):
        out[:, s] = allr[:, :, s].sum(axis=0)
    out[:, L.S_REL_H] = allr[:, :, L.S_REL_H].max(axis=0)
    flags = allr[:, :, L.S_DEV_FLAGS].astype(np.int64)
    out[:, L.S_DEV_FLAGS] = np.bitwise_or.reduce(flags, axis=0).astype(np.float64)
    return out
