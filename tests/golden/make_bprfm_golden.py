"""Golden fixture for BPR-FM: the unmodified reference BPRFM class with batch_norm=False and
drop_prob=[0, 0] (the configuration whose output does not depend on torch's global RNG), the script's loss and its
default optimiser Adagrad(lr=0.05, initial_accumulator_value=1e-8) (BPRFMRecommender.py:191-193,214-219), on a tiny
seeded problem with heavy in-batch duplicates.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_bprfm_golden.py      # build container only (needs /root/reference)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
import torch  # noqa: E402
from BPRFMRecommender import BPRFM  # noqa: E402  (reference, unmodified)


def main():
    torch.manual_seed(2019)
    U, I, F, B, steps = 50, 40, 8, 128, 4
    model = BPRFM(U + I, F, False, [0.0, 0.0])
    with torch.no_grad():
        model.embeddings.weight.mul_(20.0)
        model.biases.weight.copy_(torch.randn(U + I, 1) * 0.05)
    optimizer = torch.optim.Adagrad(model.parameters(), lr=0.05, initial_accumulator_value=1e-8)   # :191-193
    rng = np.random.default_rng(11)
    get = lambda: (model.embeddings.weight.detach().numpy().copy(), model.biases.weight.detach().numpy().copy().reshape(-1),
                   float(model.bias_.detach()))
    E0, b0, g0 = get()
    fi, fj, losses, Es, bs = [], [], [], [], []
    for k in range(steps):
        u = rng.integers(0, U, B)
        i = rng.integers(0, I, B)
        j = rng.integers(0, I, B)
        u[: B // 4] = 3
        i[B // 2: B // 2 + B // 6] = 5
        j[-B // 8:] = 5                                         # the hot item also shows up as a negative
        j[:2] = i[:2]                                           # i == j
        feat_i = np.stack([u, U + i], 1)
        feat_j = np.stack([u, U + j], 1)
        ti, tj = torch.from_numpy(feat_i).long(), torch.from_numpy(feat_j).long()
        ones = torch.ones(B, 2)
        model.zero_grad()                                       # :214-219 verbatim
        pred_i, pred_j = model(ti, ones, tj, ones)
        loss = -(pred_i - pred_j).sigmoid().log().sum()
        loss.backward()
        optimizer.step()
        E, b, g = get()
        assert g == g0                                          # bias_ has zero gradient: it never moves
        fi.append(feat_i); fj.append(feat_j); losses.append(float(loss)); Es.append(E); bs.append(b)
    with torch.no_grad():
        ones = torch.ones(B, 2)
        pi, pj = model(torch.from_numpy(fi[0]).long(), ones, torch.from_numpy(fj[0]).long(), ones)
    np.savez_compressed(os.path.join(HERE, "bprfm_small.npz"), E0=E0, b0=b0, bias_=g0, user_num=U, feats_i=np.stack(fi),
                        feats_j=np.stack(fj), losses=np.array(losses), E=np.stack(Es), b=np.stack(bs), fwd_i=pi.numpy(),
                        fwd_j=pj.numpy(), lr=0.05)
    print("wrote bprfm_small.npz: losses", losses)


if __name__ == "__main__":
    main()
