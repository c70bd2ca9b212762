"""GPU parity tests of NCF with an MLP tower (model 'MLP' and the script's default 'NeuMF-end'; SURVEY.md section 8f, row
N3): daisy_neumf_step / daisy_neumf_forward through the drop-in NeuMF + NeuMFAdam classes against the golden run of the
unmodified reference and the closed-form oracle (oracle/neumf_oracle.py).

First run on a B200 in round 2 (profiles/r02a_*).
Tolerance 1e-5 relative (max-abs-diff / max-abs) on every parameter and on the loss (tower weights: plus 1 % of one Adam step, see the test)."""
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


def make(name, init, U, I, F, L, dev, lr=1e-3):
    from recommend_lib_b200.ncf_mlp import NeuMF, NeuMFAdam
    m = NeuMF(U, I, F, L, 0.0, name)
    with torch.no_grad():
        m.embed_user_GMF.weight.copy_(torch.from_numpy(np.asarray(init["Pg"], np.float32)))
        m.embed_item_GMF.weight.copy_(torch.from_numpy(np.asarray(init["Qg"], np.float32)))
        m.embed_user_MLP.weight.copy_(torch.from_numpy(np.asarray(init["Pm"], np.float32)))
        m.embed_item_MLP.weight.copy_(torch.from_numpy(np.asarray(init["Qm"], np.float32)))
        for l, lin in enumerate(m.linears()):
            lin.weight.copy_(torch.from_numpy(np.asarray(init[f"W{l}"], np.float32)))
            lin.bias.copy_(torch.from_numpy(np.asarray(init[f"b{l}"], np.float32)))
        m.predict_layer.weight.copy_(torch.from_numpy(np.asarray(init["wp"], np.float32).reshape(1, -1)))
        m.predict_layer.bias.copy_(torch.from_numpy(np.asarray(init["bp"], np.float32).reshape(-1)))
    m = m.to(dev)
    return m, NeuMFAdam(m, lr=lr)


def state(m):
    c = lambda t: t.detach().cpu().numpy()
    d = dict(Pg=c(m.embed_user_GMF.weight), Qg=c(m.embed_item_GMF.weight), Pm=c(m.embed_user_MLP.weight),
             Qm=c(m.embed_item_MLP.weight), wp=c(m.predict_layer.weight).reshape(-1), bp=c(m.predict_layer.bias))
    for l, lin in enumerate(m.linears()):
        d[f"W{l}"], d[f"b{l}"] = c(lin.weight), c(lin.bias)
    return d


@pytest.mark.parametrize("tag,name", [("mlp", "MLP"), ("neumf", "NeuMF-end")])
def test_neumf_golden_four_steps(golden, dev, tag, name):
    g = golden("neumf_small.npz")
    k = lambda n: g[f"{tag}_{n}"]
    L, F = int(k("num_layers")), int(k("factor_num"))
    keys = ["Pg", "Qg", "Pm", "Qm", "wp", "bp"] + [f"W{l}" for l in range(L)] + [f"b{l}" for l in range(L)]
    m, opt = make(name, {key: k(key + "_0") for key in keys}, k("Pg_0").shape[0], k("Qg_0").shape[0], F, L, dev, float(k("lr")))
    for s in range(len(k("loss"))):
        opt.step(torch.from_numpy(k("users")[s]), torch.from_numpy(k("items")[s]), torch.from_numpy(k("labels")[s]))
        loss = opt.loss_sum()
        st = state(m)
        assert abs(loss - k("loss")[s]) <= 1e-5 * k("loss")[s], s
        for key in keys:
            tol = 2e-5 if key.startswith("b") else 1e-5
            assert rel_err(st[key], k(key)[s]) <= tol, (s, key, rel_err(st[key], k(key)[s]))
    m.check()
    fwd = m(torch.from_numpy(k("users")[0]), torch.from_numpy(k("items")[0])).cpu().numpy()
    assert np.allclose(fwd, k("fwd_last"), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name,U,I,F,L,B", [("NeuMF-end", 943, 1682, 32, 3, 256), ("MLP", 50, 70, 8, 1, 33),
                                             ("NeuMF-end", 300, 200, 16, 4, 1000), ("NeuMF-end", 64, 64, 128, 3, 64)])
def test_neumf_against_oracle(dev, name, U, I, F, L, B):
    from oracle import neumf_oracle
    rng = np.random.default_rng(U + B)
    Dm = F << (L - 1)
    rnd = lambda *sh: (rng.standard_normal(sh) * 0.2).astype(np.float32)
    init = dict(Pg=rnd(U, F), Qg=rnd(I, F), Pm=rnd(U, Dm), Qm=rnd(I, Dm), wp=rnd(F if name == "MLP" else 2 * F), bp=rnd(1))
    n_in = 2 * Dm
    for l in range(L):
        init[f"W{l}"], init[f"b{l}"] = rnd(n_in // 2, n_in) * (2.0 / np.sqrt(n_in)), rnd(n_in // 2) * 0.1
        n_in //= 2
    m, opt = make(name, init, U, I, F, L, dev)
    ora = neumf_oracle.NeuMFAdam(name, init["Pg"], init["Qg"], init["Pm"], init["Qm"], [init[f"W{l}"] for l in range(L)],
                                 [init[f"b{l}"] for l in range(L)], init["wp"], init["bp"])
    for s in range(3):
        u, i, y = rng.integers(0, U, B), rng.integers(0, I, B), (rng.random(B) < 0.25).astype(np.float32)
        u[: B // 4] = 3
        i[B // 2: B // 2 + B // 8] = 5
        opt.step(torch.from_numpy(u), torch.from_numpy(i), torch.from_numpy(y))
        loss, lo = opt.loss_sum(), ora.step(u, i, y)
        st = state(m)
        assert abs(loss - lo) <= 1e-5 * lo, s
        ref = dict(Pg=ora.Pg, Qg=ora.Qg, Pm=ora.Pm, Qm=ora.Qm, wp=ora.wp, bp=ora.bp)
        ref.update({f"W{l}": ora.Ws[l] for l in range(L)})
        ref.update({f"b{l}": ora.bs[l] for l in range(L)})
        for key, r in ref.items():
            # Adam's first steps move an element by lr * g / (|g| + eps): an element of the tower weights whose gradient
            # is a near-cancelling sum over the batch (|g| within a few orders of eps = 1e-8) turns fp32 rounding of g
            # into a fraction of the lr-sized step, whatever the implementation (measured on the B200: 0.24 % of one
            # step on W0 of the widest tower).  Tower weights: 1e-5 relative plus 1 % of a step; everything else 1e-5.
            if key.startswith("W"):
                assert np.abs(st[key] - r).max() <= 1e-5 * np.abs(r).max() + 0.01 * opt.lr, (s, key, rel_err(st[key], r))
            else:
                assert rel_err(st[key], r) <= 1e-5, (s, key, rel_err(st[key], r))
    m.check()
    assert np.allclose(m(torch.from_numpy(u), torch.from_numpy(i)).cpu().numpy(), ora.forward(u, i), rtol=1e-4, atol=1e-5)


def test_neumf_is_bit_reproducible_and_reports_bad_ids(dev):
    from recommend_lib_b200 import _lib
    rng = np.random.default_rng(2)
    U, I, F, L, B = 200, 150, 32, 3, 256
    Dm = F << (L - 1)
    rnd = lambda *sh: (rng.standard_normal(sh) * 0.2).astype(np.float32)
    init = dict(Pg=rnd(U, F), Qg=rnd(I, F), Pm=rnd(U, Dm), Qm=rnd(I, Dm), wp=rnd(2 * F), bp=rnd(1))
    n_in = 2 * Dm
    for l in range(L):
        init[f"W{l}"], init[f"b{l}"] = rnd(n_in // 2, n_in) * 0.1, rnd(n_in // 2) * 0.1
        n_in //= 2
    u, i, y = rng.integers(0, U, B), rng.integers(0, I, B), (rng.random(B) < 0.25).astype(np.float32)
    outs = []
    for rep in range(2):
        m, opt = make("NeuMF-end", init, U, I, F, L, dev)
        for s in range(3):
            opt.step(torch.from_numpy(u), torch.from_numpy(i), torch.from_numpy(y))
        outs.append((state(m), opt.loss_sum()))
    for key in outs[0][0]:
        assert np.array_equal(outs[0][0][key], outs[1][0][key]), key
    assert outs[0][1] == outs[1][1]
    bad = u.copy()
    bad[9] = U
    opt.step(torch.from_numpy(bad), torch.from_numpy(i), torch.from_numpy(y))
    with pytest.raises(IndexError):                             # DAISY_EINDEX, like nn.Embedding's own error
        m.check()
