"""GPU parity tests of the Item2Vec / SGNS step: daisy_sgns_step through the drop-in
Item2Vec + SGNS + SGNSAdam classes against the golden run of the unmodified reference (recorded negatives, both sampling
branches) and the closed-form oracle (oracle/sgns_oracle.py).

First run on a B200 in round 2 (profiles/r02a_*).
Tolerance 1e-5 relative (max-abs-diff / max-abs) on both tables and on the loss."""
import os

import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device (there is no CPU fallback to test)"
    return torch.device("cuda:0")


def make(iv0, ov0, n_negs, dev):
    from recommend_lib_b200.item2vec import Item2Vec, SGNS, SGNSAdam
    V, D = iv0.shape
    model = Item2Vec(vocab_size=V, embedding_size=D)
    with torch.no_grad():
        model.ivectors.weight.copy_(torch.from_numpy(np.asarray(iv0, np.float32)))
        model.ovectors.weight.copy_(torch.from_numpy(np.asarray(ov0, np.float32)))
    sgns = SGNS(embedding=model, vocab_size=V, n_negs=n_negs).to(dev)
    return model, sgns, SGNSAdam(sgns)


def tables(model):
    return model.ivectors.weight.detach().cpu().numpy(), model.ovectors.weight.detach().cpu().numpy()


@pytest.mark.parametrize("branch", ["w", "u"])
def test_sgns_golden_five_steps(golden, dev, branch):
    g = golden("sgns_small.npz")
    k = lambda name: g[f"{branch}_{name}"]
    C = k("owords").shape[2]
    n_negs = k("nwords").shape[2] // C
    model, sgns, opt = make(k("iv0"), k("ov0"), n_negs, dev)
    for s in range(len(k("losses"))):
        opt.step(torch.from_numpy(k("iword")[s]), torch.from_numpy(k("owords")[s]), torch.from_numpy(k("nwords")[s]))
        loss = opt.loss_sum()
        iv, ov = tables(model)
        assert abs(loss - k("losses")[s]) <= 1e-5 * k("losses")[s], s
        assert rel_err(iv, k("iv")[s]) <= 1e-5 and rel_err(ov, k("ov")[s]) <= 1e-5, (s, rel_err(iv, k("iv")[s]))
        assert np.abs(iv[0]).max() == 0 and np.abs(ov[0]).max() == 0          # the padding row never moves
    opt.check()


@pytest.mark.parametrize("V,D,B,C,N", [(1682, 300, 512, 10, 20), (50, 16, 33, 3, 2), (5000, 100, 4096, 4, 5), (64, 36, 7, 1, 0)])
def test_sgns_against_oracle(dev, V, D, B, C, N):
    from oracle import sgns_oracle
    rng = np.random.default_rng(V + B)
    iv0 = (rng.standard_normal((V, D)) * 0.3).astype(np.float32)
    ov0 = (rng.standard_normal((V, D)) * 0.3).astype(np.float32)
    iv0[0] = 0
    ov0[0] = 0
    model, sgns, opt = make(iv0, ov0, N, dev)
    ora = sgns_oracle.SGNSAdam(iv0, ov0)
    pop = rng.zipf(1.3, size=4 * B * C * max(N, 1)) % V                         # skewed negatives: hot output rows
    for s in range(3):
        iw = rng.integers(1, V, B)
        ow = rng.integers(0, V, (B, C))
        nw = pop[rng.integers(0, len(pop), (B, C * N))] if N else np.zeros((B, 0), np.int64)
        iw[: max(1, B // 4)] = 3
        opt.step(torch.from_numpy(iw), torch.from_numpy(ow), torch.from_numpy(nw))
        loss = opt.loss_sum()
        lo = ora.step(iw, ow, nw)
        iv, ov = tables(model)
        assert abs(loss - lo) <= 1e-5 * abs(lo), s
        assert rel_err(iv, ora.iv) <= 1e-5 and rel_err(ov, ora.ov) <= 1e-5, (s, rel_err(iv, ora.iv), rel_err(ov, ora.ov))
    opt.check()


def test_sgns_is_bit_reproducible_and_draws_its_own_negatives(dev):
    rng = np.random.default_rng(3)
    V, D, B, C, N = 400, 64, 256, 4, 5
    iv0 = (rng.standard_normal((V, D)) * 0.3).astype(np.float32)
    ov0 = (rng.standard_normal((V, D)) * 0.3).astype(np.float32)
    iw, ow, nw = rng.integers(1, V, B), rng.integers(0, V, (B, C)), rng.integers(0, V, (B, C * N))
    outs = []
    for rep in range(2):
        model, sgns, opt = make(iv0, ov0, N, dev)
        for s in range(3):
            opt.step(torch.from_numpy(iw), torch.from_numpy(ow), torch.from_numpy(nw))
        outs.append(tables(model) + (opt.loss_sum(),))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1]) and outs[0][2] == outs[1][2]
    losses = []
    for rep in range(2):                                        # negatives drawn by the step: seeded runs agree
        torch.manual_seed(11)
        model, sgns, opt = make(iv0, ov0, N, dev)
        opt.step(torch.from_numpy(iw), torch.from_numpy(ow))
        losses.append(opt.loss_sum())
    assert losses[0] == losses[1] and np.isfinite(losses[0])


def test_sgns_reports_bad_ids(dev):
    from recommend_lib_b200 import _lib
    rng = np.random.default_rng(5)
    V, D, B, C, N = 40, 8, 16, 2, 3
    model, sgns, opt = make((rng.standard_normal((V, D)) * 0.1).astype(np.float32),
                            (rng.standard_normal((V, D)) * 0.1).astype(np.float32), N, dev)
    iw, ow, nw = rng.integers(1, V, B), rng.integers(0, V, (B, C)), rng.integers(0, V, (B, C * N))
    nw[5, 2] = V                                                # out of range
    opt.step(torch.from_numpy(iw), torch.from_numpy(ow), torch.from_numpy(nw))
    with pytest.raises(IndexError):                             # DAISY_EINDEX, like nn.Embedding's own error
        opt.check()
    with pytest.raises(RuntimeError):
        sgns(torch.from_numpy(iw), torch.from_numpy(ow))
