"""World-size-2 (and 3) gloo tests of the row-sharded BPR path on CPU: routing by user owner, id / row / gradient
exchanges (DistComm over gloo P2P), owner-side accumulation.  The per-rank compute is the oracle backend
(tests/sharded_testing.py) -- the CUDA kernels of the same path are checked on the GPU in test_sharded_gpu.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem(U, I, D, B, steps, seed):
    rng = np.random.default_rng(seed)
    P0 = (rng.standard_normal((U, D)) * 0.3).astype(np.float32)
    Q0 = (rng.standard_normal((I, D)) * 0.3).astype(np.float32)
    batches = []
    for _ in range(steps):
        b = np.stack([rng.integers(0, U, B), rng.integers(0, I, B), rng.integers(0, I, B)], 1).astype(np.int32)
        b[: B // 4, 1] = 3                       # a hot positive item owned by rank 0, requested by every rank
        b[B // 2: B // 2 + 5, 0] = U - 1          # a repeated user on the last rank
        batches.append(b)
    return P0, Q0, batches


def _worker(rank, world, port, U, I, D, B, steps, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from recommend_lib_b200.sharded import ShardedBPR, DistComm
        from sharded_testing import OracleBackend, route
        P0, Q0, batches = _problem(U, I, D, B, steps, seed=7)
        m = ShardedBPR(U, I, D, lr=0.05, wd=0.01, max_batch=B, rank=rank, world=world, device="cpu", comm=DistComm(),
                       backend=OracleBackend(), P_full=P0, Q_full=Q0)
        for b in batches:
            m.step(torch.from_numpy(route(b, m.layout, rank)))
        P, Q = m.full_tables()
        loss = torch.tensor([m.backend._loss], dtype=torch.float64)
        dist.all_reduce(loss)
        if rank == 0:
            np.savez(out, P=P.numpy(), Q=Q.numpy(), loss=loss.numpy(), wire=m.wire_rows)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,U,I", [(2, 41, 30), (3, 50, 31)])
def test_sharded_equals_single_process_oracle(tmp_path, world, U, I):
    from oracle import bpr_oracle
    D, B, steps = 8, 300, 3
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(world, _free_port(), U, I, D, B, steps, out), nprocs=world, join=True)
    r = np.load(out)
    P0, Q0, batches = _problem(U, I, D, B, steps, seed=7)
    Pr, Qr, losses = bpr_oracle.bpr_run_closed_form(P0, Q0, batches, 0.05, 0.01, np.float64)
    assert np.abs(r["P"] - Pr).max() / np.abs(Pr).max() < 1e-5
    assert np.abs(r["Q"] - Qr).max() / np.abs(Qr).max() < 1e-5
    assert abs(float(r["loss"][0]) - sum(losses)) / sum(losses) < 1e-6
    assert int(r["wire"]) > 0


def test_layout_and_plan_grouping():
    from recommend_lib_b200.sharded import ShardLayout, ShardedBPR
    from sharded_testing import OracleBackend
    lay = ShardLayout(10, 7, 3)
    assert [lay.user_range(r) for r in range(3)] == [(0, 4), (4, 8), (8, 10)]
    assert [lay.item_range(r) for r in range(3)] == [(0, 3), (3, 6), (6, 7)]
    m = ShardedBPR(10, 7, 4, rank=1, world=3, device="cpu", backend=OracleBackend(),
                   P_full=np.zeros((10, 4), np.float32), Q_full=np.zeros((7, 4), np.float32))
    tri = torch.tensor([[0, 6, 1], [1, 1, 5], [2, 6, 6]], dtype=torch.int32)
    ids, counts, local = m.plan(tri)
    assert ids.tolist() == [1, 5, 6] and counts == [1, 1, 1]          # sorted unique ids, grouped by owner block
    assert local.tolist() == [[0, 2, 0], [1, 0, 1], [2, 2, 2]]        # item columns are cache indices
    assert m.P.shape == (4, 4) and m.Q.shape == (3, 4)
