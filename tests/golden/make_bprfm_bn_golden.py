"""Golden fixture for BPR-FM in the script's DEFAULT configuration: the unmodified reference BPRFM class
with batch_norm=True and drop_prob=[0.5, 0.2] (BPRFMRecommender.py:116-125), the script's loss and optimiser
(:191-193, 214-219).  nn.Dropout draws its mask inside forward from torch's global generator; the draw is recorded by
seeding the generator, making the two draws forward makes (positive call, then negative call), and seeding it again
before the reference's forward.  Two optimiser settings: the script's initial_accumulator_value 1e-8 (ill-conditioned
where a gradient nearly cancels) and 0.1 (well conditioned, tight comparison).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_bprfm_bn_golden.py      # build container only
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
import torch  # noqa: E402
import torch.nn.functional as Fn  # noqa: E402
from BPRFMRecommender import BPRFM  # noqa: E402  (reference, unmodified)


def run(acc0, batch_norm=True, p=0.5, U=50, I=40, F=8, B=96, steps=4):
    torch.manual_seed(2019)
    model = BPRFM(U + I, F, batch_norm, [p, 0.2])
    with torch.no_grad():
        model.embeddings.weight.mul_(40.0)
        model.biases.weight.copy_(torch.randn(U + I, 1) * 0.05)
    optimizer = torch.optim.Adagrad(model.parameters(), lr=0.05, initial_accumulator_value=acc0)   # :191-193
    rng = np.random.default_rng(13)
    bn = model.FM_layers[0] if batch_norm else None

    def get():
        d = dict(E=model.embeddings.weight.detach().numpy().copy(), b=model.biases.weight.detach().numpy().copy().reshape(-1),
                 g=np.float32(model.bias_.detach().item()))
        if bn is not None:
            d.update(gamma=bn.weight.detach().numpy().copy(), beta=bn.bias.detach().numpy().copy(),
                     rm=bn.running_mean.numpy().copy(), rv=bn.running_var.numpy().copy())
        return d

    s0 = get()
    rec = {k: [] for k in ("fi", "fj", "mi", "mj", "loss", "E", "b", "g", "gamma", "beta", "rm", "rv")}
    model.train()
    for k in range(steps):
        u = rng.integers(0, U, B)
        i = rng.integers(0, I, B)
        j = rng.integers(0, I, B)
        u[: B // 4] = 3
        i[B // 2: B // 2 + B // 6] = 5
        j[-B // 8:] = 5
        feat_i, feat_j = np.stack([u, U + i], 1), np.stack([u, U + j], 1)
        ti, tj = torch.from_numpy(feat_i).long(), torch.from_numpy(feat_j).long()
        ones = torch.ones(B, 2)
        seed = 500 + k
        torch.manual_seed(seed)                                  # the two dropout draws of forward, recorded
        mi = Fn.dropout(torch.ones(B, F), p, True).numpy().copy()
        mj = Fn.dropout(torch.ones(B, F), p, True).numpy().copy()
        torch.manual_seed(seed)
        model.zero_grad()                                        # :214-219 verbatim
        pred_i, pred_j = model(ti, ones, tj, ones)
        loss = -(pred_i - pred_j).sigmoid().log().sum()
        loss.backward()
        optimizer.step()
        s = get()
        rec["fi"].append(feat_i); rec["fj"].append(feat_j); rec["mi"].append(mi); rec["mj"].append(mj)
        rec["loss"].append(float(loss))
        for key in ("E", "b", "g") + (("gamma", "beta", "rm", "rv") if bn is not None else ()):
            rec[key].append(s[key])
    model.eval()
    with torch.no_grad():
        ones = torch.ones(B, 2)
        pi, pj = model(torch.from_numpy(rec["fi"][0]).long(), ones, torch.from_numpy(rec["fj"][0]).long(), ones)
    out = dict(E0=s0["E"], b0=s0["b"], g0=s0["g"], fwd_i=pi.numpy(), fwd_j=pj.numpy(), acc0=acc0, p=p, lr=0.05, user_num=U)
    out.update({k: np.stack(v) for k, v in rec.items() if v})
    return out


def main():
    out = {}
    for name, acc0 in (("script", 1e-8), ("cond", 0.1)):
        r = run(acc0)
        out.update({f"{name}_{k}": v for k, v in r.items()})
        print(name, "losses", r["loss"])
    np.savez_compressed(os.path.join(HERE, "bprfm_bn_small.npz"), **out)
    print("wrote bprfm_bn_small.npz")


if __name__ == "__main__":
    main()
