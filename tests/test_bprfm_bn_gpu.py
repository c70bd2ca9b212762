"""GPU parity tests of BPR-FM at the script's DEFAULTS (batch norm + dropout):
daisy_fmbn_step / daisy_fmbn_forward through BPRFMBN + FMBNAdagrad against the golden run of the unmodified reference
(recorded dropout masks) and the closed-form oracle (oracle/bprfm_oracle.py: BPRFMFull).

First run on a B200 in round 2 (profiles/r02a_*).  Tolerances
are the ones of tests/test_bprfm_gpu.py: 1e-5 on well-conditioned quantities; with the script's Adagrad accumulator of
1e-8 an element whose gradient nearly cancels is ill-conditioned (embeddings 1e-4, biases 2e-4, see
test_oracle_golden.py::test_bprfm_full_oracle_matches_reference_with_batch_norm_and_dropout)."""
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


def make(E0, b0, bias_, U, p, dev):
    from recommend_lib_b200.bprfm_bn import BPRFMBN
    m = BPRFMBN(E0.shape[0], E0.shape[1], True, [p, 0.2], user_num=U)
    with torch.no_grad():
        m.embeddings.weight.copy_(torch.from_numpy(np.asarray(E0, np.float32)))
        m.biases.weight.copy_(torch.from_numpy(np.asarray(b0, np.float32).reshape(-1, 1)))
        m.bias_.fill_(float(bias_))
    return m.to(dev)


def state(m):
    bn = m.FM_layers[0]
    c = lambda t: t.detach().cpu().numpy()
    return dict(E=c(m.embeddings.weight), b=c(m.biases.weight).reshape(-1), gamma=c(bn.weight), beta=c(bn.bias),
                rm=c(bn.running_mean), rv=c(bn.running_var))


@pytest.mark.parametrize("branch,tol_E,tol_b", [("cond", 1e-5, 1e-5), ("script", 1e-4, 2e-4)])
def test_fmbn_golden_four_steps(golden, dev, branch, tol_E, tol_b):
    from recommend_lib_b200.bprfm_bn import FMBNAdagrad
    g = golden("bprfm_bn_small.npz")
    k = lambda name: g[f"{branch}_{name}"]
    U = int(k("user_num"))
    m = make(k("E0"), k("b0"), float(k("g0")), U, float(k("p")), dev)
    m.train()
    opt = FMBNAdagrad(m, lr=float(k("lr")), initial_accumulator_value=float(k("acc0")))
    ones = torch.ones(k("fi")[0].shape)
    for s in range(len(k("loss"))):
        opt.step(torch.from_numpy(k("fi")[s]), ones, torch.from_numpy(k("fj")[s]), ones,
                 mask_i=torch.from_numpy(k("mi")[s]), mask_j=torch.from_numpy(k("mj")[s]))
        loss = opt.loss_sum()
        st = state(m)
        assert abs(loss - k("loss")[s]) <= 1e-5 * k("loss")[s], s
        assert rel_err(st["E"], k("E")[s]) <= tol_E, (s, rel_err(st["E"], k("E")[s]))
        assert rel_err(st["b"], k("b")[s]) <= tol_b, (s, rel_err(st["b"], k("b")[s]))
        assert rel_err(st["gamma"], k("gamma")[s]) <= 1e-5 and rel_err(st["beta"], k("beta")[s]) <= 2e-5, s
        assert rel_err(st["rm"], k("rm")[s]) <= 1e-5 and rel_err(st["rv"], k("rv")[s]) <= 1e-5, s
    m.check()
    assert int(m.FM_layers[0].num_batches_tracked) == 2 * len(k("loss"))
    m.eval()
    pi, pj = m(torch.from_numpy(k("fi")[0]), ones, torch.from_numpy(k("fj")[0]), ones)
    assert np.allclose(pi.cpu().numpy(), k("fwd_i"), rtol=1e-4, atol=2e-5)
    assert np.allclose(pj.cpu().numpy(), k("fwd_j"), rtol=1e-4, atol=2e-5)


def _problem(U, I, F, B, p, seed):
    rng = np.random.default_rng(seed)
    E0 = (rng.standard_normal((U + I, F)) * 0.4).astype(np.float32)
    b0 = (rng.standard_normal(U + I) * 0.05).astype(np.float32)
    u = rng.integers(0, U, B)
    i = rng.integers(0, I, B)
    j = rng.integers(0, I, B)
    u[: B // 5] = 3                                             # a hot user, a hot item on both sides
    i[B // 2: B // 2 + B // 7] = 5
    j[-B // 9:] = 5
    keep = lambda: ((rng.random((B, F)) >= p) / (1.0 - p)).astype(np.float32) if p > 0 else None
    return E0, b0, np.stack([u, i, j], 1).astype(np.int32), keep(), keep()


@pytest.mark.parametrize("U,I,F,B,p", [(300, 200, 64, 4096, 0.5), (50, 40, 8, 96, 0.5), (2000, 900, 32, 20000, 0.0),
                                        (64, 64, 100, 333, 0.2)])
def test_fmbn_against_oracle(dev, U, I, F, B, p):
    from oracle import bprfm_oracle
    from recommend_lib_b200.bprfm_bn import FMBNAdagrad
    E0, b0, tri, mi, mj = _problem(U, I, F, B, p, seed=U + B)
    m = make(E0, b0, 0.0, U, p, dev)
    m.train()
    opt = FMBNAdagrad(m, lr=0.05, initial_accumulator_value=0.1)
    ora = bprfm_oracle.BPRFMFull(E0, b0, 0.0, True, lr=0.05, initial_accumulator_value=0.1)
    fi = np.stack([tri[:, 0], U + tri[:, 1]], 1)
    fj = np.stack([tri[:, 0], U + tri[:, 2]], 1)
    ones = np.ones((B, 2))
    t = torch.from_numpy(tri).to(dev)
    tm = lambda a: torch.from_numpy(a) if a is not None else None
    for s in range(3):
        opt.step(t, mask_i=tm(mi), mask_j=tm(mj))
        loss = opt.loss_sum()
        lo = ora.step(fi, ones, fj, ones, mi, mj)
        st = state(m)
        assert abs(loss - lo) <= 1e-5 * abs(lo), s
        assert rel_err(st["E"], ora.E) <= 1e-5 and rel_err(st["b"], ora.bias) <= 1e-5, (s, rel_err(st["E"], ora.E))
        assert rel_err(st["gamma"], ora.gamma) <= 1e-5 and rel_err(st["beta"], ora.beta) <= 2e-5, s
        assert rel_err(st["rm"], ora.running_mean) <= 1e-5 and rel_err(st["rv"], ora.running_var) <= 1e-5, s
    m.check()
    m.eval()
    ones_t = torch.ones(B, 2)
    pi, pj = m(torch.from_numpy(fi), ones_t, torch.from_numpy(fj), ones_t)
    oi, oj = ora.forward(fi, ones, fj, ones)
    assert np.allclose(pi.cpu().numpy(), oi, rtol=1e-4, atol=1e-5) and np.allclose(pj.cpu().numpy(), oj, rtol=1e-4, atol=1e-5)


def test_fmbn_is_bit_reproducible_and_draws_its_own_masks(dev):
    from recommend_lib_b200.bprfm_bn import FMBNAdagrad
    E0, b0, tri, mi, mj = _problem(300, 200, 64, 4096, 0.5, seed=7)
    outs = []
    for rep in range(2):
        m = make(E0, b0, 0.0, 300, 0.5, dev)
        m.train()
        opt = FMBNAdagrad(m, lr=0.05)
        for s in range(3):
            opt.step(torch.from_numpy(tri).to(dev), mask_i=torch.from_numpy(mi), mask_j=torch.from_numpy(mj))
        outs.append((state(m), opt.loss_sum()))
    for key in outs[0][0]:
        assert np.array_equal(outs[0][0][key], outs[1][0][key]), key
    assert outs[0][1] == outs[1][1]
    # without given masks the step draws them (torch CUDA generator): seeded runs agree, the loss stays finite
    losses = []
    for rep in range(2):
        torch.manual_seed(5)
        m = make(E0, b0, 0.0, 300, 0.5, dev)
        m.train()
        opt = FMBNAdagrad(m, lr=0.05)
        opt.step(torch.from_numpy(tri).to(dev))
        losses.append(opt.loss_sum())
    assert losses[0] == losses[1] and np.isfinite(losses[0])


def test_fmbn_reports_bad_ids_and_refuses_training_mode_forward(dev):
    from recommend_lib_b200 import _lib
    from recommend_lib_b200.bprfm_bn import FMBNAdagrad
    E0, b0, tri, _, _ = _problem(50, 40, 8, 96, 0.0, seed=1)
    m = make(E0, b0, 0.0, 50, 0.0, dev)
    m.train()
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 2, dtype=torch.long), None, torch.zeros(2, 2, dtype=torch.long), None)
    bad = tri.copy()
    bad[17, 2] = 40                                             # item id == item_num
    opt = FMBNAdagrad(m, lr=0.05)
    opt.step(torch.from_numpy(bad).to(dev))
    with pytest.raises(IndexError):                             # DAISY_EINDEX, like nn.Embedding's own error
        m.check()
    with pytest.raises(ValueError):                             # batch norm needs two samples (torch raises ValueError too)
        opt.step(torch.from_numpy(tri[:1]).to(dev))
