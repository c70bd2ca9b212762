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
    K, W = args.steps, max(args.warmup, 3)
    model = ShardedBPR(U, I, D, lr=cfg["lr"], wd=cfg["wd"], max_batch=B, rank=rank, world=world, device=dev,
                       comm=DistComm(), seed=2019)
    u0, u1 = model.layout.user_range(rank)
    nb = K + W
    g = _rng(2019, 40, rank)
    host = np.empty((nb * B, 3), dtype=np.int32)
    host[:, 0] = g.integers(0, u1 - u0, size=nb * B)                       # local user index: routed by owner
    host[:, 1] = zipf_items(g, nb * B, I, cfg["zipf"], perm_seed=2019)     # global ids, Zipf over a permuted catalogue
    host[:, 2] = g.integers(0, I, size=nb * B)
    host = torch.from_numpy(host.reshape(nb, B, 3)).pin_memory()
    devtri = host.to(dev)

    def run(first, count, src):
        for s in range(first, first + count):
            model.step(src[s] if src is devtri else src[s].to(dev, non_blocking=True))

    run(0, W, devtri)
    model.backend.check()
    torch.cuda.synchronize()
    dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    model.wire_rows = 0
    launches0 = model.backend.h.launches
    torch.cuda.synchronize()
    ev0.record()
    run(W, K, devtri)
    model.materialize()
    ev1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    launches = model.backend.h.launches - launches0
    wire_rows = model.wire_rows
    # e2e: host triples, per-step loss read-back
    loss_host = torch.zeros(nb, dtype=torch.float64).pin_memory()
    run(0, W, host)
    torch.cuda.synchronize()
    dist.barrier()
    ev0.record()
    for s in range(W, W + K):
        model.step(host[s].to(dev, non_blocking=True))
        loss_host[s:s + 1].copy_(model.backend.loss, non_blocking=True)
    model.materialize()
    ev1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms2 = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    model.backend.check()
    phases = None
    if args.phases:
        model.profile = []
        run(0, min(nb, 10), devtri)
        phases = model.profile_summary()
        model.profile = None
    if rank == 0:
        ms_total, ms_e2e = float(ms), float(ms2)
        value = B * world * K / (ms_total * 1e-3)
        wire_bytes = wire_rows / K * 4 * D                      # per step, this rank, rows in + rows out
        line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": cfg["workload"], "user_num": U, "item_num": I, "dim": D,
                           "batch_per_gpu": B, "global_batch": B * world, "lr": cfg["lr"], "wd": cfg["wd"],
                           "sharding": "block rows, triples routed to the user's owner, item rows + row gradients "
                                       "exchanged by NCCL all-to-all", "l2": "inputs larger than L2"},
                "e2e": {"value": B * world * K / (ms_e2e * 1e-3), "unit": unit, "ms_per_step": ms_e2e / K,
                        "h2d_bytes_per_step": B * 12 * world, "d2h_bytes_per_step": 8 * world},
                "gpu_launches": int(launches),
                "nvlink": {"rows_exchanged_per_step_per_gpu": wire_rows / K, "bytes_per_step_per_gpu_each_way": wire_bytes / 2,
                           "achieved_GBs_each_way": wire_bytes / 2 / (ms_total / K * 1e-3) / 1e9,
                           "measured_peer_copy_GBs": 770.0}}
        if phases:
            line["phase_ms(device,host)"] = phases
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()
