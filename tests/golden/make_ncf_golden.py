"""Golden fixture for the NEXT row (NCF-GMF): the unmodified reference NCF class (model='GMF') trained for a
few steps with the script's own loss and optimiser (NCFRecommender.py:255-260,283-287) on a tiny seeded problem with
heavy in-batch duplicates and rows that stay untouched for several steps (dense Adam keeps moving them).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_ncf_golden.py      # build container only (needs /root/reference)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
import torch  # noqa: E402
from NCFRecommender import NCF  # noqa: E402  (reference, unmodified)


def main():
    torch.manual_seed(2019)
    U, I, F, B, steps = 60, 45, 8, 96, 5
    model = NCF(U, I, F, 3, 0.0, "GMF", None, None)
    with torch.no_grad():                                   # larger weights than N(0, .01): gradients that matter
        model.embed_user_GMF.weight.mul_(30.0)
        model.embed_item_GMF.weight.mul_(30.0)
    loss_function = torch.nn.BCEWithLogitsLoss()            # :255
    optimizer = torch.optim.Adam(model.parameters(), lr=0.001)   # :260
    rng = np.random.default_rng(7)
    get = lambda: dict(P=model.embed_user_GMF.weight.detach().numpy().copy(), Q=model.embed_item_GMF.weight.detach().numpy().copy(),
                       w=model.predict_layer.weight.detach().numpy().copy().reshape(-1), b=model.predict_layer.bias.detach().numpy().copy())
    s0 = get()
    users, items, labels, losses, Ps, Qs, ws, bs = [], [], [], [], [], [], [], []
    for k in range(steps):
        u = rng.integers(0, U // 2 if k % 2 else U, B)      # odd steps leave half of the users untouched
        i = rng.integers(0, I, B)
        u[: B // 4] = 3                                     # a hot user, a hot item
        i[B // 2: B // 2 + B // 8] = 5
        y = (rng.random(B) < 0.2).astype(np.float32)
        user, item, label = torch.from_numpy(u).long(), torch.from_numpy(i).long(), torch.from_numpy(y).float()
        model.zero_grad()                                   # :283-287 verbatim
        prediction = model(user, item)
        loss = loss_function(prediction, label)
        loss.backward()
        optimizer.step()
        s = get()
        users.append(u); items.append(i); labels.append(y); losses.append(float(loss))
        Ps.append(s["P"]); Qs.append(s["Q"]); ws.append(s["w"]); bs.append(s["b"])
    with torch.no_grad():
        fwd = model(torch.from_numpy(users[0]).long(), torch.from_numpy(items[0]).long()).numpy()
    np.savez_compressed(os.path.join(HERE, "gmf_small.npz"), P0=s0["P"], Q0=s0["Q"], w0=s0["w"], b0=s0["b"],
                        users=np.stack(users), items=np.stack(items), labels=np.stack(labels), losses=np.array(losses),
                        P=np.stack(Ps), Q=np.stack(Qs), w=np.stack(ws), b=np.stack(bs), fwd_last=fwd, lr=0.001)
    print("wrote gmf_small.npz: losses", losses)


def trajectory(epochs=20):
    """The script's loop (NCFRecommender.py:262-289) for model 'GMF' at its defaults (factor_num 32, batch 256, lr 0.001,
    num_ng 4, 20 epochs) on the ml-100k split of config 1, fed by the deterministic SampleSampler; evaluation with the
    reference's own metric_eval(algo='ncf') on 1 + 999 candidates per user."""
    import json
    import time
    ROOT = os.path.dirname(os.path.dirname(HERE))
    sys.path.insert(0, ROOT)
    from torch.utils.data import DataLoader
    from util.data_loader import NCFData            # reference, unmodified
    from util.metrics import metric_eval            # reference, unmodified
    from recommend_lib_b200 import data as hostdata
    from recommend_lib_b200.sampler import SampleSampler
    s = np.load(os.path.join(HERE, "ml100k_split.npz"))
    tr, te = s["train_pairs"].astype(np.int64), s["test_pairs"].astype(np.int64)
    U, I = int(s["user_num"]), int(s["item_num"])
    allp = np.concatenate([tr, te])
    eu, ec = hostdata.eval_candidates(allp[:, 0], allp[:, 1], te[:, 0], te[:, 1], I, 999, 2019)
    eval_list = [[int(u), int(c)] for u, row in zip(eu, ec) for c in row]
    test_loader = DataLoader(NCFData(eval_list, I, None, 0, False), batch_size=1000, shuffle=False, num_workers=0)
    torch.manual_seed(2019)
    model = NCF(U, I, 32, 3, 0.0, "GMF", None, None)
    init = dict(P0=model.embed_user_GMF.weight.detach().numpy().copy(), Q0=model.embed_item_GMF.weight.detach().numpy().copy(),
                w0=model.predict_layer.weight.detach().numpy().copy().reshape(-1), b0=model.predict_layer.bias.detach().numpy().copy())
    loss_function = torch.nn.BCEWithLogitsLoss()
    optimizer = torch.optim.Adam(model.parameters(), lr=0.001)
    sampler = SampleSampler(tr, I, num_ng=4, seed=2019)
    out = {"epochs": [], "eval_users": int(len(eu)), "batch": 256, "lr": 0.001, "factor_num": 32}
    t0 = time.time()
    for ep in range(epochs):
        smp = sampler.sample_epoch(ep)
        model.train()
        total = 0.0
        for b in range(0, len(smp), 256):
            user = torch.from_numpy(smp[b:b + 256, 0]).long()
            item = torch.from_numpy(smp[b:b + 256, 1]).long()
            label = torch.from_numpy(smp[b:b + 256, 2]).float()
            model.zero_grad()
            prediction = model(user, item)
            loss = loss_function(prediction, label)
            loss.backward()
            optimizer.step()
            total += float(loss)
        rec = {"epoch": ep + 1, "loss_sum": total}
        if ep + 1 in (1, epochs):
            model.eval()
            with torch.no_grad():
                hr, ndcg = metric_eval(model, test_loader, 10, algo="ncf")
            rec["hr"], rec["ndcg"] = float(hr), float(ndcg)
        out["epochs"].append(rec)
        print(rec, f"{time.time() - t0:.0f}s", flush=True)
    np.savez_compressed(os.path.join(HERE, "gmf_ml100k_init.npz"), **init)
    json.dump(out, open(os.path.join(HERE, "gmf_ml100k_traj.json"), "w"), indent=1)


if __name__ == "__main__":
    if "--trajectory" in sys.argv:
        trajectory()
    else:
        main()
