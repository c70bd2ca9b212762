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


if __name__ == "__main__":
    main()
