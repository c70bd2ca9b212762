"""Golden fixture for NCF with an MLP tower: the unmodified reference NCF class with model='MLP' and with
the script's default model='NeuMF-end' (NCFRecommender.py:175-178), dropout 0 (the script's default), trained for a few
steps with the script's own loss and optimiser (:255-260, 283-287) on a tiny seeded problem with in-batch duplicates and
rows that stay untouched for several steps (dense Adam keeps moving them).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_neumf_golden.py      # build container only (needs /root/reference)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
import torch  # noqa: E402
from NCFRecommender import NCF  # noqa: E402  (reference, unmodified)


def run(model_name, U=40, I=30, F=8, L=3, B=64, steps=4):
    torch.manual_seed(2019)
    model = NCF(U, I, F, L, 0.0, model_name, None, None)
    with torch.no_grad():                                   # larger than N(0, .01): gradients that matter
        for e in (model.embed_user_GMF, model.embed_item_GMF, model.embed_user_MLP, model.embed_item_MLP):
            e.weight.mul_(30.0)
    loss_function = torch.nn.BCEWithLogitsLoss()            # :255
    optimizer = torch.optim.Adam(model.parameters(), lr=0.001)   # :260
    rng = np.random.default_rng(17)
    lin = [m for m in model.MLP_layers if isinstance(m, torch.nn.Linear)]
    n = lambda t: t.detach().numpy().copy()

    def get():
        d = dict(Pg=n(model.embed_user_GMF.weight), Qg=n(model.embed_item_GMF.weight), Pm=n(model.embed_user_MLP.weight),
                 Qm=n(model.embed_item_MLP.weight), wp=n(model.predict_layer.weight).reshape(-1), bp=n(model.predict_layer.bias))
        for l, m in enumerate(lin):
            d[f"W{l}"], d[f"b{l}"] = n(m.weight), n(m.bias)
        return d

    out = {f"{k}_0": v for k, v in get().items()}
    rec = {k: [] for k in list(get().keys()) + ["users", "items", "labels", "loss"]}
    for k in range(steps):
        u = rng.integers(0, U // 2 if k % 2 else U, B)      # odd steps leave half of the users untouched
        i = rng.integers(0, I, B)
        u[: B // 4] = 3
        i[B // 2: B // 2 + B // 8] = 5
        y = (rng.random(B) < 0.2).astype(np.float32)
        user, item, label = torch.from_numpy(u).long(), torch.from_numpy(i).long(), torch.from_numpy(y).float()
        model.zero_grad()                                   # :283-287 verbatim
        prediction = model(user, item)
        loss = loss_function(prediction, label)
        loss.backward()
        optimizer.step()
        for key, v in get().items():
            rec[key].append(v)
        rec["users"].append(u); rec["items"].append(i); rec["labels"].append(y); rec["loss"].append(float(loss))
    with torch.no_grad():
        fwd = model(torch.from_numpy(rec["users"][0]).long(), torch.from_numpy(rec["items"][0]).long()).numpy()
    out.update({k: np.stack(v) for k, v in rec.items()})
    out.update(fwd_last=fwd, lr=0.001, num_layers=L, factor_num=F)
    return out


def main():
    res = {}
    for tag, name in (("mlp", "MLP"), ("neumf", "NeuMF-end")):
        r = run(name)
        res.update({f"{tag}_{k}": v for k, v in r.items()})
        print(name, "losses", r["loss"])
    np.savez_compressed(os.path.join(HERE, "neumf_small.npz"), **res)
    print("wrote neumf_small.npz")


if __name__ == "__main__":
    main()
