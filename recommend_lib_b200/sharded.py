"""Row-sharded BPR-MF for catalogues larger than one GPU (BASELINE.json configs[4]).

No counterpart in the reference (single process, single device).  One process per GPU; tables are
block-sharded by row: rank r owns users [r*u_per, (r+1)*u_per) and items [r*i_per, (r+1)*i_per).  Triples are routed
to the owner of their user (the sampler draws each rank's users from its own block), so the user gather / update is
local and only item rows travel.  One step:

    plan      unique item ids of the batch (sorted => already grouped by owner), remap triples to cache indices
    exchange  ids -> owners                                   (all-to-all, int32)
    serve     owners gather the requested rows                (local gather)
    exchange  rows -> requesters  = the batch's row cache     (all-to-all, 4*D B per row)
    compute   daisy_bpr_shard_step: fused step on (P_local, cache); users updated in place, one descent sum per
              cache row written out
    exchange  descent sums -> owners                          (all-to-all, 4*D B per row)
    apply     daisy_owner_apply: per row, contributions summed in sender-rank order, Q_local updated once

Every accumulation order is fixed (stable sorts, rank order), so the sharded step is deterministic and equals the
single-GPU step up to fp32 summation order.  The collectives are NCCL all-to-all(v) over NVLink; on the gloo backend
(CPU tests of the routing logic) the same exchanges run as batched isend/irecv.
"""
from __future__ import annotations

import ctypes
import math
import os
import time

import numpy as np
import torch
import torch.distributed as dist


class ShardLayout:
    """Block sharding of both tables over `world` ranks."""

    def __init__(self, user_num, item_num, world):
        self.user_num, self.item_num, self.world = int(user_num), int(item_num), int(world)
        self.u_per = math.ceil(self.user_num / self.world)
        self.i_per = math.ceil(self.item_num / self.world)

    def user_range(self, rank):
        return min(rank * self.u_per, self.user_num), min((rank + 1) * self.u_per, self.user_num)

    def item_range(self, rank):
        return min(rank * self.i_per, self.item_num), min((rank + 1) * self.i_per, self.item_num)

    def item_bounds(self, device):
        b = [min(r * self.i_per, self.item_num) for r in range(self.world + 1)]
        return torch.tensor(b, dtype=torch.int64, device=device)


class DistComm:
    """Variable-size all-to-all over torch.distributed (NCCL: all_to_all_single; gloo: batched P2P)."""

    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.nccl = dist.get_backend(group) == "nccl"

    def counts(self, send_counts, device):
        """send_counts[r] = how many rows this rank sends to r  ->  recv_counts[r] = how many it receives from r."""
        t = torch.tensor(send_counts, dtype=torch.int64, device=device)
        if self.nccl:
            out = torch.empty_like(t)
            dist.all_to_all_single(out, t, group=self.group)
            return out.tolist()
        allc = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(allc, t, group=self.group)
        return [int(allc[r][self.rank]) for r in range(self.world)]

    def exchange(self, x, send_counts, recv_counts):
        """x: rows grouped by destination rank (send_counts each); returns rows grouped by source rank."""
        out = x.new_empty((int(sum(recv_counts)),) + tuple(x.shape[1:]))
        if self.nccl:
            dist.all_to_all_single(out, x.contiguous(), list(recv_counts), list(send_counts), group=self.group)
            return out
        ops, so, ro = [], 0, 0
        keep = []
        for r in range(self.world):
            s, c = send_counts[r], recv_counts[r]
            if r == self.rank:
                out[ro:ro + c] = x[so:so + s]
            else:
                if s:
                    chunk = x[so:so + s].contiguous()
                    keep.append(chunk)
                    ops.append(dist.P2POp(dist.isend, chunk, r, group=self.group))
                if c:
                    ops.append(dist.P2POp(dist.irecv, out[ro:ro + c], r, group=self.group))
            so += s
            ro += c
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return out


class CudaBackend:
    """The product backend: libdaisy_b200 kernels on this rank's GPU.  No CPU fallback."""

    def __init__(self, layout, rank, dim, max_batch, device):
        from . import _lib
        _lib.require_cuda()
        self._lib = _lib
        self.device = torch.device(device)
        u0, u1 = layout.user_range(rank)
        i0, i1 = layout.item_range(rank)
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.h = _lib.Handle(idx, max(u1 - u0, 1), max(i1 - i0, 1), dim, max_batch)
        self.dim = dim
        self.loss = torch.zeros(1, dtype=torch.float64, device=self.device)

    def _s(self):
        return self._lib.stream_ptr(torch, self.device)

    def gather_rows(self, Q_local, rows_local):
        return Q_local.index_select(0, rows_local.long())

    def shard_step(self, P_local, cache, tri_local, lr, wd):
        vp = self._lib.c_vp
        grads = torch.empty_like(cache)
        self._lib.check(self.h.L.daisy_bpr_shard_step(self.h.ptr, vp(P_local.data_ptr()), vp(cache.data_ptr()),
                                                      cache.shape[0], vp(tri_local.data_ptr()), tri_local.shape[0],
                                                      lr, wd, vp(grads.data_ptr()), vp(self.loss.data_ptr()), self._s()))
        return grads

    def owner_apply(self, Q_local, rows_local, grads, lr, wd):
        vp = self._lib.c_vp
        self._lib.check(self.h.L.daisy_owner_apply(self.h.ptr, vp(Q_local.data_ptr()), vp(rows_local.data_ptr()),
                                                   vp(grads.data_ptr()), rows_local.shape[0], lr, wd, self._s()))

    def materialize(self, P_local, Q_local):
        vp = self._lib.c_vp
        self._lib.check(self.h.L.daisy_materialize(self.h.ptr, vp(P_local.data_ptr()), vp(Q_local.data_ptr()), self._s()))

    def check(self):
        self._lib.check(self.h.L.daisy_check(self.h.ptr, self._s()))

    def loss_sum(self, reset=True):
        v = float(self.loss.item())
        if reset:
            self.loss.zero_()
        return v


class ShardedBPR:
    """One rank's share of a row-sharded BPR-MF model.

    ``step(triples)``: ``triples`` int32 [B, 3] on this rank's device with columns (LOCAL user index, GLOBAL positive
    item, GLOBAL negative item).  ``P_full`` / ``Q_full`` (optional, host tensors) initialise the shards from given
    tables (parity tests); otherwise shards are drawn N(0, 0.01^2) from a (seed, rank) generator
    (BPRMFRecommender.py:39-40).
    """

    def __init__(self, user_num, item_num, factor_num, lr=0.01, wd=0.001, max_batch=4096, rank=None, world=None,
                 device=None, comm=None, backend=None, P_full=None, Q_full=None, seed=2019):
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.layout = ShardLayout(user_num, item_num, self.world)
        self.dim, self.lr, self.wd = int(factor_num), float(lr), float(wd)
        self.device = torch.device(device if device is not None else "cpu")
        u0, u1 = self.layout.user_range(self.rank)
        i0, i1 = self.layout.item_range(self.rank)
        self.u0, self.i0 = u0, i0
        if P_full is not None:
            self.P = torch.as_tensor(P_full)[u0:u1].to(self.device, torch.float32).contiguous()
            self.Q = torch.as_tensor(Q_full)[i0:i1].to(self.device, torch.float32).contiguous()
        else:
            g = torch.Generator(device=self.device).manual_seed(seed * 1000 + self.rank)
            self.P = torch.empty((u1 - u0, self.dim), device=self.device).normal_(0, 0.01, generator=g)
            self.Q = torch.empty((i1 - i0, self.dim), device=self.device).normal_(0, 0.01, generator=g)
        self.comm = comm
        self.backend = backend if backend is not None else CudaBackend(self.layout, self.rank, self.dim, max_batch,
                                                                       self.device)
        self._bounds = self.layout.item_bounds(self.device)
        self.profile = None
        self.wire_rows = 0         # rows received + sent over the interconnect (both exchanges), for reporting

    # ---- phases (also driven one by one by the in-process multi-rank emulation in the tests) --------------------
    def plan(self, triples):
        B = triples.shape[0]
        items = triples[:, 1:].t().reshape(-1)                       # [i_0..i_B-1, j_0..j_B-1]
        uniq, inv = torch.unique(items, sorted=True, return_inverse=True)
        pos = torch.searchsorted(uniq.long(), self._bounds)          # sorted ids are already grouped by owner block
        send_counts = (pos[1:] - pos[:-1]).tolist()                  # host sync: split sizes of the exchanges
        tri_local = torch.stack([triples[:, 0], inv[:B].to(torch.int32), inv[B:].to(torch.int32)], dim=1).contiguous()
        return uniq.to(torch.int32), send_counts, tri_local

    def serve(self, recv_ids):
        return self.backend.gather_rows(self.Q, recv_ids - self.i0)

    def compute(self, tri_local, cache):
        return self.backend.shard_step(self.P, cache, tri_local, self.lr, self.wd)

    def apply(self, recv_ids, grads_in):
        self.backend.owner_apply(self.Q, (recv_ids - self.i0).to(torch.int32).contiguous(), grads_in, self.lr, self.wd)

    # ---- one full step over torch.distributed -----------------------------------------------------------------------
    PHASES = ("plan", "counts", "ids", "serve", "rows", "compute", "grads", "apply")

    def _mark(self, name):
        if self.profile is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.profile.append((name, e, time.perf_counter()))

    def step(self, triples):
        c = self.comm
        self._mark("start")
        ids, send_counts, tri_local = self.plan(triples)
        self._mark("plan")
        recv_counts = c.counts(send_counts, self.device)
        self._mark("counts")
        recv_ids = c.exchange(ids, send_counts, recv_counts)
        self._mark("ids")
        rows = self.serve(recv_ids)
        self._mark("serve")
        cache = c.exchange(rows, recv_counts, send_counts)
        self._mark("rows")
        grads = self.compute(tri_local, cache)
        self._mark("compute")
        grads_in = c.exchange(grads, send_counts, recv_counts)
        self._mark("grads")
        self.apply(recv_ids, grads_in)
        self._mark("apply")
        remote = sum(send_counts) - send_counts[self.rank]
        self.wire_rows += 2 * remote

    def profile_summary(self):
        """Mean device ms and host ms per phase over the profiled steps (enable with ``self.profile = []``)."""
        torch.cuda.synchronize(self.device)
        dev, host, n = {}, {}, 0
        prev = None
        for name, ev, t in self.profile:
            if name == "start":
                n += 1
            else:
                dev[name] = dev.get(name, 0.0) + prev[1].elapsed_time(ev)
                host[name] = host.get(name, 0.0) + (t - prev[2]) * 1e3
            prev = (name, ev, t)
        return {k: (round(dev[k] / n, 3), round(host[k] / n, 3)) for k in dev}

    def materialize(self):
        self.backend.materialize(self.P, self.Q)

    def full_tables(self):
        """All-gather the (materialised) shards: (P [U, D], Q [I, D]) on every rank.  Test / small-model helper."""
        self.materialize()
        out = []
        for t in (self.P, self.Q):
            n = torch.tensor([t.shape[0]], device=self.device)
            ns = [torch.zeros_like(n) for _ in range(self.world)]
            dist.all_gather(ns, n)
            mx = int(max(int(x) for x in ns))
            pad = torch.zeros((mx, self.dim), device=self.device)
            pad[:t.shape[0]] = t
            parts = [torch.empty_like(pad) for _ in range(self.world)]
            dist.all_gather(parts, pad)
            out.append(torch.cat([p[:int(k)] for p, k in zip(parts, ns)]))
        return out[0], out[1]


class _DeviceView:
    """A [rows, dim] fp32 window on device memory owned by the library (the item shard inside the rank's arena),
    exposed to torch through __cuda_array_interface__."""

    def __init__(self, ptr, rows, dim):
        self.__cuda_array_interface__ = {"shape": (int(rows), int(dim)), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def exchange_ipc_handles(blob, group=None, device=None):
    """All-gather one 64-byte CUDA IPC handle per rank over torch.distributed (any backend) -> list of bytes."""
    world = dist.get_world_size(group)
    dev = device if (device is not None and dist.get_backend(group) == "nccl") else torch.device("cpu")
    mine = torch.tensor(list(blob), dtype=torch.uint8, device=dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    return [bytes(p.cpu().tolist()) for p in parts]


class PeerShardedBPR:
    """One rank's share of a row-sharded BPR-MF model whose exchange runs over PEER MEMORY (NVLink) inside the step
    kernels -- the product path of BASELINE.json configs[4] (include/daisy_b200.h, daisy_shard_*).  No CPU fallback.

    ``P`` [local users, D] is an ordinary CUDA tensor; ``Q`` [local items, D] is a view of the item shard inside the
    rank's arena, which the peer GPUs map.  ``mapping``:
      "symm"  (default for world > 1) the arena is a torch symmetric-memory buffer (cuMem VMM, rendezvous over the
              process group) -- peer reads of random rows run at NVLink speed through this mapping;
      "ipc"   the library cudaMallocs the arena and the ranks exchange legacy CUDA IPC handles (measured 2x slower
              for random row reads; kept for platforms without symmetric memory);
      "local" all ranks live in this process on one device (tests): wire with ``connect_in_process``.
    ``step(triples)``: int32 [B, 3] with columns (LOCAL user index, GLOBAL positive item, GLOBAL negative item), on the
    device or in pinned host memory.  Every rank must call ``step`` the same number of times.
    """

    def __init__(self, user_num, item_num, factor_num, lr=0.01, wd=0.001, max_batch=4096, rank=None, world=None,
                 device=None, P_full=None, Q_full=None, seed=2019, mapping=None, group=None):
        from . import _lib
        _lib.require_cuda()
        self._lib = _lib
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.group = group
        self.mapping = mapping or ("symm" if self.world > 1 else "local")
        assert self.mapping in ("symm", "ipc", "local")
        self.layout = ShardLayout(user_num, item_num, self.world)
        self.dim, self.lr, self.wd = int(factor_num), float(lr), float(wd)
        self.device = torch.device(device if device is not None else "cuda")
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        u0, u1 = self.layout.user_range(self.rank)
        i0, i1 = self.layout.item_range(self.rank)
        self.u0, self.i0 = u0, i0
        self.h = _lib.Handle(idx, max(u1 - u0, 1), max(i1 - i0, 1), self.dim, max_batch)
        L, vp = self.h.L, _lib.c_vp
        self._symm_buf = self._symm_hdl = None
        arena_arg = None
        if self.mapping == "symm":
            import torch.distributed._symmetric_memory as symm
            need = _lib.c_i64()
            _lib.check(L.daisy_shard_arena_size(self.dim, int(max_batch), self.world, int(item_num), ctypes.byref(need)))
            with torch.cuda.device(self.device):
                self._symm_buf = symm.empty((need.value + 3) // 4, dtype=torch.float32, device=self.device)
            arena_arg = vp(self._symm_buf.data_ptr())
        _lib.check(L.daisy_shard_init(self.h.ptr, self.rank, self.world, int(item_num), arena_arg))
        arena, q, nbytes = vp(), vp(), _lib.c_i64()
        _lib.check(L.daisy_shard_arena(self.h.ptr, ctypes.byref(arena), ctypes.byref(q), ctypes.byref(nbytes)))
        self.arena_ptr, self.arena_bytes = arena.value, nbytes.value
        with torch.cuda.device(self.device):
            self.Q = torch.as_tensor(_DeviceView(q.value, max(i1 - i0, 1), self.dim), device=self.device)[:i1 - i0]
            if P_full is not None:
                self.P = torch.as_tensor(P_full)[u0:u1].to(self.device, torch.float32).contiguous()
                self.Q.copy_(torch.as_tensor(Q_full)[i0:i1].to(self.device, torch.float32))
            else:
                g = torch.Generator(device=self.device).manual_seed(seed * 1000 + self.rank)
                self.P = torch.empty((u1 - u0, self.dim), device=self.device).normal_(0, 0.01, generator=g)
                self.Q.normal_(0, 0.01, generator=g)
            self.loss = torch.zeros(1, dtype=torch.float64, device=self.device)
            self._P_keep = self.P if self.P.shape[0] else torch.zeros((1, self.dim), device=self.device)
        self._P_ptr = self._P_keep.data_ptr()
        self.attached = self.world == 1
        torch.cuda.synchronize(self.device)

    # ---- wiring ---------------------------------------------------------------------------------------------------
    def ipc_handle(self):
        buf = (ctypes.c_ubyte * 64)()
        self._lib.check(self.h.L.daisy_shard_ipc_handle(self.h.ptr, ctypes.cast(buf, ctypes.c_void_p)))
        return bytes(buf)

    def connect(self):
        """Map the peers' arenas (one process per GPU): symmetric-memory rendezvous, or legacy IPC handle exchange."""
        if self.world > 1:
            L = self.h.L
            if self.mapping == "symm":
                import torch.distributed._symmetric_memory as symm
                self._symm_hdl = symm.rendezvous(self._symm_buf, self.group if self.group is not None else dist.group.WORLD)
                ptrs = [int(p) for p in self._symm_hdl.buffer_ptrs]
                assert ptrs[self.rank] == self.arena_ptr, "symmetric-memory rendezvous returned a different local address"
                arr = (ctypes.c_void_p * self.world)(*ptrs)
                self._lib.check(L.daisy_shard_attach(self.h.ptr, None, arr, 0))
            elif self.mapping == "ipc":
                blobs = exchange_ipc_handles(self.ipc_handle(), self.group, self.device)
                raw = b"".join(blobs)
                buf = (ctypes.c_ubyte * len(raw)).from_buffer_copy(raw)
                self._lib.check(L.daisy_shard_attach(self.h.ptr, ctypes.cast(buf, ctypes.c_void_p), None, 0))
            else:
                raise ValueError('mapping "local" is wired with PeerShardedBPR.connect_in_process(shards)')
            self.attached = True
            torch.cuda.synchronize(self.device)
            dist.barrier(self.group)
        return self

    @staticmethod
    def connect_in_process(shards):
        """All ranks live in this process on one device (tests): hand every rank the others' arena pointers."""
        G = len(shards)
        arr = (ctypes.c_void_p * G)(*[s.arena_ptr for s in shards])
        for s in shards:
            if G > 1:
                s._lib.check(s.h.L.daisy_shard_attach(s.h.ptr, None, arr, 1))
            s.attached = True

    def _s(self):
        return self._lib.stream_ptr(torch, self.device)

    # ---- the step ---------------------------------------------------------------------------------------------------
    def step(self, triples):
        vp = self._lib.c_vp
        B = int(triples.shape[0])
        fn = self.h.L.daisy_shard_step if triples.is_cuda else self.h.L.daisy_shard_step_host
        self._lib.check(fn(self.h.ptr, vp(self._P_ptr), vp(triples.data_ptr() if B else 0), B, self.lr, self.wd,
                           vp(self.loss.data_ptr()), self._s()))

    def compute(self, triples):
        vp = self._lib.c_vp
        B = int(triples.shape[0])
        self._lib.check(self.h.L.daisy_shard_compute(self.h.ptr, vp(self._P_ptr), vp(triples.data_ptr() if B else 0), B,
                                                     self.lr, self.wd, vp(self.loss.data_ptr()), self._s()))

    def apply(self):
        self._lib.check(self.h.L.daisy_shard_apply(self.h.ptr, self.lr, self.wd, self._s()))

    def materialize(self):
        self._lib.check(self.h.L.daisy_shard_materialize(self.h.ptr, self._lib.c_vp(self._P_ptr), self._s()))

    def check(self):
        self._lib.check(self.h.L.daisy_check(self.h.ptr, self._s()))

    def last_counts(self):
        """Distinct item rows the most recent step exchanged with every owner (list of `world` ints).  Synchronises."""
        buf = (ctypes.c_uint32 * (self.world + 1))()
        self._lib.check(self.h.L.daisy_shard_last_counts(self.h.ptr, ctypes.cast(buf, ctypes.c_void_p), self._s()))
        return [int(buf[o + 1]) - int(buf[o]) for o in range(self.world)]

    SHARD_PHASES = ("bookkeeping", "fetch", "compute_push", "barrier1", "apply", "barrier2")

    def phase_ms(self):
        arr = (ctypes.c_double * 6)()
        n = ctypes.c_int64()
        self._lib.check(self.h.L.daisy_shard_phase_ms(self.h.ptr, arr, ctypes.byref(n)))
        return dict(zip(self.SHARD_PHASES, [round(x, 4) for x in arr])), n.value

    def loss_sum(self, reset=True, group=None, reduce=False):
        t = self.loss.clone()
        if reduce and self.world > 1:
            dist.all_reduce(t, group=group)
        if reset:
            self.loss.zero_()
        return float(t.item())

    def full_tables(self):
        """All-gather the (materialised) shards: (P [U, D], Q [I, D]) on every rank.  Test / small-model helper."""
        self.materialize()
        out = []
        for t, per in ((self.P, self.layout.u_per), (self.Q, self.layout.i_per)):
            pad = torch.zeros((per, self.dim), device=self.device)
            pad[:t.shape[0]] = t
            parts = [torch.empty_like(pad) for _ in range(self.world)]
            dist.all_gather(parts, pad)
            out.append(torch.cat(parts))
        return out[0][:self.layout.user_num], out[1][:self.layout.item_num]

    def close(self):
        torch.cuda.synchronize(self.device)
        self.Q = None
        self.h.close()
        self._symm_hdl = self._symm_buf = None


# ----------------------------------------------------------------------------------------------------------------
# bench.py --gpus N  (launched by torchrun, one rank per GPU)
# ----------------------------------------------------------------------------------------------------------------
def bench_sharded(args, cfg, metric, unit):
    import json
    from .sampler import _rng, zipf_items
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    U, I, D, B = cfg["user_num"], cfg["item_num"], cfg["dim"], cfg["batch"]
    if args.scale != 1.0:
        U, I = int(U * args.scale), int(I * args.scale)
    if args.batch:
        B = args.batch
    K, W = args.steps, max(args.warmup, 3)
    peer = args.exchange == "peer"
    if peer:
        model = PeerShardedBPR(U, I, D, lr=cfg["lr"], wd=cfg["wd"], max_batch=B, rank=rank, world=world, device=dev,
                               seed=2019, mapping=args.mapping).connect()
        handle, loss_dev, check = model.h, model.loss, model.check
    else:
        model = ShardedBPR(U, I, D, lr=cfg["lr"], wd=cfg["wd"], max_batch=B, rank=rank, world=world, device=dev,
                           comm=DistComm(), seed=2019)
        handle, loss_dev, check = model.backend.h, model.backend.loss, model.backend.check
    u0, u1 = model.layout.user_range(rank)
    nb = K + W
    g = _rng(2019, 40, rank)
    host = np.empty((nb * B, 3), dtype=np.int32)
    host[:, 0] = g.integers(0, u1 - u0, size=nb * B)                       # local user index: routed by owner
    host[:, 1] = zipf_items(g, nb * B, I, cfg["zipf"], perm_seed=2019)     # global ids, Zipf over a permuted catalogue
    host[:, 2] = g.integers(0, I, size=nb * B)
    host = torch.from_numpy(host.reshape(nb, B, 3)).pin_memory()
    devtri = host.to(dev)
    if peer:
        handle.set_inputs_ready(True)      # device triples are uploaded and synchronised before they are used

    def run(first, count, src):
        for s in range(first, first + count):
            if peer:
                model.step(src[s])
            else:
                model.step(src[s] if src is devtri else src[s].to(dev, non_blocking=True))

    run(0, W, devtri)
    model.materialize()                             # warm the lazy-decay pass too (first launch loads its code)
    check()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if not peer:
        model.wire_rows = 0
    launches0 = handle.launches
    clocks = None
    try:                                            # every rank samples its own GPU (rank 0's goes into `clocks`)
        import bench as _bench
        clocks = _bench.ClockSampler(local)         # NVML initialisation takes tens of ms: before the barrier
    except Exception:
        clocks = None
    torch.cuda.synchronize()
    dist.barrier()                                  # all ranks enter the timed region together
    if clocks is not None:
        clocks.start()
    if os.environ.get("DAISY_TRACE_TIMED") and peer:
        handle.trace_start()
    ev0.record()
    run(W, K, devtri)
    model.materialize()
    ev1.record()
    torch.cuda.synchronize()
    if clocks is not None:
        clocks.stop()
    if os.environ.get("DAISY_TRACE_TIMED") and peer:
        rows = [[round(x, 2) for x in row] for row in handle.trace_dump()]
        print(f"rank {rank} timed-region trace:", rows, flush=True)
    dist.barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    launches = handle.launches - launches0
    if peer:
        cnt = model.last_counts()
        remote = sum(cnt) - cnt[rank]
        wire_rows = 2 * remote * K         # fetched rows in + pushed sums out, per rank (last step's count, every step alike)
    else:
        wire_rows = model.wire_rows
    # e2e: host triples, per-step loss read-back
    loss_host = torch.zeros(nb, dtype=torch.float64).pin_memory()
    run(0, W, host)
    torch.cuda.synchronize()
    dist.barrier()
    ev0.record()
    for s in range(W, W + K):
        if peer:
            model.step(host[s])
        else:
            model.step(host[s].to(dev, non_blocking=True))
        loss_host[s:s + 1].copy_(loss_dev, non_blocking=True)
    model.materialize()
    ev1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms2 = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    check()
    phases = None
    # the fused peer path always adds its per-rank phase profile (10 serialised steps AFTER both timed regions): it is
    # what says which rank the others wait for at the barrier; --no-phases drops it
    want_phases = args.phases or (peer and not getattr(args, "no_phases", False))
    if args.phases and not peer:
        model.profile = []
        run(0, min(nb, 10), devtri)
        phases = model.profile_summary()
        model.profile = None
    elif want_phases:
        handle.set_timing(2)
        run(0, min(nb, 10), devtri)
        phases, _ = model.phase_ms()
        inner, _ = handle.phase_ms()
        phases["compute_push_detail"] = {k: round(v, 4) for k, v in inner.items()}
        handle.set_timing(0)
    if args.trace and peer:
        torch.cuda.synchronize()
        dist.barrier()
        handle.trace_start()
        run(0, min(nb, 12), devtri)
        rows = [[round(x, 3) for x in row] for row in handle.trace_dump()]
        for r in range(world):
            if r == rank:
                print(f"rank {rank} trace_ms(book_begin, book_end, kernels_begin, compute_end):", rows, flush=True)
            dist.barrier()
    by_rank = None
    if want_phases and peer:                        # which rank waits for which: the phases of every rank, side by side
        mine = {"rank": rank, "sm_mhz": (clocks.summary()["sm_mhz"] if clocks is not None else None)}
        mine.update({k: round(v, 4) for k, v in phases.items() if k != "compute_push_detail"})
        mine["main"] = phases["compute_push_detail"].get("main")
        by_rank = [None] * world
        dist.all_gather_object(by_rank, mine)
    if rank == 0:
        ms_total, ms_e2e = float(ms), float(ms2)
        value = B * world * K / (ms_total * 1e-3)
        # NVLink bytes per step and direction at one GPU: it RECEIVES the rows it fetches and SERVES the rows its peers
        # fetch from it (egress), and it PUSHES its row sums (egress) and receives its peers' (ingress); by symmetry
        # every direction carries (remote rows) x 4D bytes twice.  wire_rows = 2 x remote rows of this rank.
        each_way = wire_rows / K * 4 * D
        peak_nvl = 770.0
        step_s = ms_total / K * 1e-3
        line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": cfg["workload"], "user_num": U, "item_num": I, "dim": D,
                           "batch_per_gpu": B, "global_batch": B * world, "lr": cfg["lr"], "wd": cfg["wd"],
                           "sharding": ("block rows, triples routed to the user's owner; item rows read from / row "
                                        "sums stored to the owners' memory by the step kernels over NVLink (peer "
                                        "pointers), flag barriers, deterministic owner-side merge") if peer else
                                       ("block rows, triples routed to the user's owner, item rows + row gradients "
                                        "exchanged by NCCL all-to-all"),
                           "exchange": args.exchange, "peer_mapping": args.mapping if peer else None,
                           "main_schedule": (f"chunks dealt round-robin over "
                                             f"{os.environ.get('DAISY_SHARD_INTERLEAVE') or world} owner ranges "
                                             f"(DAISY_SHARD_INTERLEAVE; 0 = sorted order)") if peer else None,
                           "l2": "inputs larger than L2", "lazy_decay_materialized_in_timed_region": True},
                "clocks": clocks.summary() if clocks is not None else None,
                "e2e": {"value": B * world * K / (ms_e2e * 1e-3), "unit": unit, "ms_per_step": ms_e2e / K,
                        "h2d_bytes_per_step": B * 12 * world, "d2h_bytes_per_step": 8 * world},
                "gpu_launches": int(launches),
                "roofline": {"bound": "nvlink", "achieved": each_way / step_s / 1e9, "peak": peak_nvl,
                             "unit": "GB/s per direction per GPU", "frac": each_way / step_s / 1e9 / peak_nvl,
                             "traffic": None, "peak_source": "measured peer copy (B200_PROFILING.md)",
                             "remote_rows_per_step_per_gpu": wire_rows / K / 2,
                             "nvlink_bytes_per_step_per_gpu_each_way": each_way,
                             "note": "whole step (compute + barriers + owner merge), not the fused kernel alone",
                             "hbm_whole_step_frac": (B * (24 * D + 12) / step_s / 1e9) / 6461.8}}
        if phases:
            line["phase_ms(device,host)"] = phases
        if by_rank:
            line["phase_ms_by_rank"] = by_rank
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()
