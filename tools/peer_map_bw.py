"""Diagnostic (2 GPUs, torchrun): random 512-byte row gather from the peer GPU through (a) the legacy CUDA-IPC mapping
of the library's arena and (b) a torch symmetric-memory (cuMem VMM) mapping.  Decides how the arenas are mapped.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/peer_map_bw.py
"""
import ctypes
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recommend_lib_b200 import _lib  # noqa: E402
from recommend_lib_b200.sharded import PeerShardedBPR  # noqa: E402

rank = int(os.environ["RANK"])
world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
D, ROWS, N = 128, 10_000_000, 700_000
m = PeerShardedBPR(1000, ROWS * world, D, max_batch=1 << 20, rank=rank, world=world, device=dev, mapping="ipc").connect()
L, vp = m.h.L, _lib.c_vp
idx = torch.randint(0, ROWS, (N,), device=dev, dtype=torch.int32).sort().values
dst = torch.empty((N, D), device=dev)


def timeit(name, src_ptr):
    s = _lib.stream_ptr(torch, dev)
    for _ in range(2):
        _lib.check(L.daisy_gather_rows(m.h.ptr, vp(src_ptr), vp(idx.data_ptr()), N, vp(dst.data_ptr()), s))
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        _lib.check(L.daisy_gather_rows(m.h.ptr, vp(src_ptr), vp(idx.data_ptr()), N, vp(dst.data_ptr()), s))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"rank {rank} {name:40s} {ms:7.3f} ms {N * D * 4 / ms * 1e-6:8.1f} GB/s", flush=True)
    dist.barrier()


# peer arena pointers as the library mapped them: read them back through a tiny ctypes peek at the handle is not part
# of the ABI, so re-open is avoided -- instead use a second exported pointer: the peer q base equals the peer arena base
# (q sits at offset 0), which daisy_shard_attach stored; expose it via daisy_shard_peer_q below if present.
get = getattr(L, "daisy_shard_peer_q", None)
timeit("local arena q (cudaMalloc)", m.Q.data_ptr())
if get is not None:
    get.argtypes = [vp, ctypes.c_int, ctypes.POINTER(vp)]
    get.restype = ctypes.c_int
    p = vp()
    _lib.check(get(m.h.ptr, (rank + 1) % world, ctypes.byref(p)))
    timeit("peer q via legacy cudaIpc mapping", p.value)
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty((ROWS, D), dtype=torch.float32, device=dev)
    t.fill_(1.0)
    hdl = symm.rendezvous(t, dist.group.WORLD)
    ptrs = list(hdl.buffer_ptrs)
    timeit("local symmetric-memory buffer", ptrs[rank])
    timeit("peer symmetric-memory buffer (VMM mapping)", ptrs[(rank + 1) % world])
except Exception as e:  # noqa: BLE001
    print(f"rank {rank}: symmetric memory unavailable: {type(e).__name__}: {e}", flush=True)
dist.barrier()
dist.destroy_process_group()
