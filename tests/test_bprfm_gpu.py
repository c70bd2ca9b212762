"""GPU parity tests of the BPR-FM path: daisy_bprfm_adagrad_step through the drop-in BPRFM
class against the golden run of the unmodified reference and the closed-form oracle.  Tolerance: biases 1e-5, embeddings
1e-4 relative (max-abs-diff / max-abs) -- Adagrad with state_sum starting at 1e-8 is ill-conditioned on elements whose
gradient nearly cancels (tests/test_oracle_golden.py::test_bprfm_oracle_matches_reference_with_adagrad)."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device (there is no CPU fallback to test)"
    return torch.device("cuda:0")


def make(E0, b0, bias_, U, dev, max_batch=4096):
    from recommend_lib_b200.bprfm import BPRFM
    m = BPRFM(E0.shape[0], E0.shape[1], False, [0.0, 0.0], user_num=U, max_batch=max_batch)
    with torch.no_grad():
        m.embeddings.weight.copy_(torch.from_numpy(np.asarray(E0, np.float32)))
        m.biases.weight.copy_(torch.from_numpy(np.asarray(b0, np.float32).reshape(-1, 1)))
        m.bias_.fill_(float(bias_))
    return m.to(dev)


def state(m):
    m.sync()
    return m.embeddings.weight.detach().cpu().numpy(), m.biases.weight.detach().cpu().numpy().reshape(-1)


def test_bprfm_golden_four_steps(golden, dev):
    from recommend_lib_b200.bprfm import FMAdagrad
    g = golden("bprfm_small.npz")
    U = int(g["user_num"])
    m = make(g["E0"], g["b0"], float(g["bias_"]), U, dev)
    opt = FMAdagrad(m, lr=float(g["lr"]))
    ones = torch.ones(g["feats_i"][0].shape)
    for k in range(len(g["losses"])):
        opt.step(torch.from_numpy(g["feats_i"][k]), ones, torch.from_numpy(g["feats_j"][k]), ones)
        loss = opt.loss_sum()
        E, b = state(m)
        assert abs(loss - g["losses"][k]) <= 1e-5 * g["losses"][k], k
        assert rel_err(E, g["E"][k]) <= 1e-4 and rel_err(b, g["b"][k]) <= 1e-5, (k, rel_err(E, g["E"][k]), rel_err(b, g["b"][k]))
    m.check()
    pi, pj = m(torch.from_numpy(g["feats_i"][0]), ones, torch.from_numpy(g["feats_j"][0]), ones)
    assert np.allclose(pi.cpu().numpy(), g["fwd_i"], rtol=1e-4, atol=1e-5)
    assert np.allclose(pj.cpu().numpy(), g["fwd_j"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("U,I,F,B,steps", [(300, 200, 64, 4096, 3), (50, 40, 32, 700, 3), (5000, 3000, 8, 20000, 2),
                                            (2000, 500, 16, 100000, 1)])
def test_bprfm_against_oracle(dev, U, I, F, B, steps):
    """Small / mid / general bookkeeping paths (the step is the BPR step with another optimiser functor)."""
    from oracle import bprfm_oracle
    from recommend_lib_b200.bprfm import FMAdagrad
    rng = np.random.default_rng(U + B)
    E0 = (rng.standard_normal((U + I, F)) * 0.3).astype(np.float32)
    b0 = (rng.standard_normal(U + I) * 0.05).astype(np.float32)
    m = make(E0, b0, 0.25, U, dev, max_batch=B)
    opt = FMAdagrad(m, lr=0.05, initial_accumulator_value=1e-2)     # a well-conditioned accumulator: 1e-5 holds
    E, b = E0.astype(np.float64), b0.astype(np.float64)
    aE, ab = np.full_like(E, 1e-2), np.full_like(b, 1e-2)
    for k in range(steps):
        u, i, j = rng.integers(0, U, B), rng.integers(0, I, B), rng.integers(0, I, B)
        u[: B // 4] = 2
        i[B // 2: B // 2 + B // 6] = 1
        j[-B // 8:] = 1
        fi, fj = np.stack([u, U + i], 1), np.stack([u, U + j], 1)
        opt.step(torch.from_numpy(fi), None, torch.from_numpy(fj), None)
        loss = bprfm_oracle.bprfm_adagrad_step(E, b, 0.25, aE, ab, fi, fj, lr=0.05)
        assert abs(opt.loss_sum() - loss) <= 1e-5 * loss, k
    m.check()
    Eg, bg = state(m)
    # biases are small (|b| <= 0.2) next to one Adagrad step (lr = 0.05): 5e-5 of max|b| is 2e-4 of a step -- an item with
    # hundreds of +-s contributions that nearly cancel carries the fp32 rounding of that sum
    assert rel_err(Eg, E) <= 1e-5 and rel_err(bg, b) <= 5e-5, (rel_err(Eg, E), rel_err(bg, b))
    assert np.array_equal(bg[:U], b0[:U])                            # user biases never move


def test_bprfm_rejects_what_is_not_accelerated(dev):
    from recommend_lib_b200.bprfm import BPRFM, FMAdagrad
    from recommend_lib_b200.bprfm_bn import BPRFMBN, FMBNAdagrad
    # the script's defaults (batch norm + dropout, BPRFMRecommender.py:116-125) build the batch-norm model ...
    bn = BPRFM(10, 8, True, [0.5, 0.2], user_num=4)
    assert isinstance(bn, BPRFMBN) and bn.drop_prob == [0.5, 0.2] and "FM_layers.0.weight" in bn.state_dict()
    assert isinstance(FMAdagrad(bn, lr=0.05), FMBNAdagrad)
    # ... dropout WITHOUT batch norm (not selectable from the script's command line) is on no accelerated path
    with pytest.raises(NotImplementedError):
        BPRFM(10, 8, False, [0.5, 0.0], user_num=4)
    with pytest.raises(ValueError):
        BPRFM(10, 8, False, [0.0, 0.0])
    m = BPRFM(10, 8, False, [0.0, 0.0], user_num=4).to(dev)
    with pytest.raises(NotImplementedError):
        m.triples(torch.tensor([[0, 5]]), torch.tensor([[1.0, 0.5]]), torch.tensor([[0, 6]]), None)
