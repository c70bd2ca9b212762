"""GPU tests of the row-sharded kernels (daisy_bpr_shard_step, daisy_owner_apply) on ONE device: G shards live in this
process and step in lockstep (exchanges by slicing, tests/sharded_testing.py), as the profiling guide prescribes when
there are fewer GPUs than ranks.  Checked against the CPU oracle and against the unsharded CUDA step."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import rel_err  # noqa: E402

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _problem(U, I, D, B, steps, seed):
    rng = np.random.default_rng(seed)
    P0 = (rng.standard_normal((U, D)) * 0.3).astype(np.float32)
    Q0 = (rng.standard_normal((I, D)) * 0.3).astype(np.float32)
    batches = []
    for _ in range(steps):
        b = np.stack([rng.integers(0, U, B), rng.integers(0, I, B), rng.integers(0, I, B)], 1).astype(np.int32)
        b[: B // 4, 1] = 3
        b[B // 2: B // 2 + B // 10, 0] = U - 1
        b[-B // 8:, 2] = 3
        batches.append(b)
    return P0, Q0, batches


@pytest.mark.parametrize("G,U,I,D,B", [(2, 400, 300, 64, 6000), (3, 1000, 701, 128, 20000), (4, 64, 50, 32, 37)])
def test_sharded_lockstep_matches_oracle_and_unsharded(G, U, I, D, B):
    assert torch.cuda.is_available()
    from oracle import bpr_oracle
    from recommend_lib_b200.bpr import BPR, BPRSGD
    from recommend_lib_b200.sharded import ShardedBPR
    from sharded_testing import lockstep_step, route
    dev = torch.device("cuda:0")
    lr, wd, steps = 0.05, 0.01, 3
    P0, Q0, batches = _problem(U, I, D, B, steps, seed=G)
    shards = [ShardedBPR(U, I, D, lr=lr, wd=wd, max_batch=B, rank=r, world=G, device=dev, P_full=P0, Q_full=Q0)
              for r in range(G)]
    for b in batches:
        lockstep_step(shards, [torch.from_numpy(route(b, s.layout, s.rank)).to(dev) for s in shards])
    for s in shards:
        s.backend.check()
        s.materialize()
    P = torch.cat([s.P for s in shards]).cpu().numpy()
    Q = torch.cat([s.Q for s in shards]).cpu().numpy()
    loss = sum(s.backend.loss_sum() for s in shards)
    Pr, Qr, losses = bpr_oracle.bpr_run_closed_form(P0, Q0, batches, lr, wd, np.float64)
    assert rel_err(P, Pr) <= 1e-5 and rel_err(Q, Qr) <= 1e-5
    assert abs(loss - sum(losses)) / sum(losses) < 1e-5
    # the unsharded CUDA step on the same global batches
    m = BPR(U, I, D, max_batch=B)
    with torch.no_grad():
        m.embed_user.weight.copy_(torch.from_numpy(P0))
        m.embed_item.weight.copy_(torch.from_numpy(Q0))
    m = m.to(dev)
    opt = BPRSGD(m, lr=lr, weight_decay=wd)
    for b in batches:
        opt.step(torch.from_numpy(b).to(dev))
    m.materialize()
    # both are fp32 with different (fixed) summation orders for the hot rows: same tolerance as against the oracle
    assert rel_err(P, m.embed_user.weight.detach().cpu().numpy()) <= 1e-5
    assert rel_err(Q, m.embed_item.weight.detach().cpu().numpy()) <= 1e-5


def test_sharded_step_is_deterministic():
    from recommend_lib_b200.sharded import ShardedBPR
    from sharded_testing import lockstep_step, route
    dev = torch.device("cuda:0")
    G, U, I, D, B = 4, 3000, 2000, 128, 50000
    P0, Q0, batches = _problem(U, I, D, B, 2, seed=9)
    outs = []
    for _ in range(2):
        shards = [ShardedBPR(U, I, D, max_batch=B, rank=r, world=G, device=dev, P_full=P0, Q_full=Q0) for r in range(G)]
        for b in batches:
            lockstep_step(shards, [torch.from_numpy(route(b, s.layout, s.rank)).to(dev) for s in shards])
        for s in shards:
            s.materialize()
        outs.append((torch.cat([s.P for s in shards]).clone(), torch.cat([s.Q for s in shards]).clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
