"""Golden fixture for a NEXT row (Item2Vec / SGNS): the unmodified reference Item2Vec + SGNS classes
trained for a few steps with the script's own optimiser (Item2VecRecommender.py:259-277) on a tiny seeded problem.
SGNS.forward draws its negatives from torch's global generator (:86-91); the draw is recorded by seeding the generator,
making the same call the reference makes, and seeding it again before the reference's forward.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_sgns_golden.py      # build container only (needs /root/reference)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
import torch  # noqa: E402
from Item2VecRecommender import Item2Vec, SGNS  # noqa: E402  (reference, unmodified)


def run(weighted, V=40, D=12, B=24, C=4, N=5, steps=5, seed=2019):
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed + 1)
    model = Item2Vec(vocab_size=V, embedding_size=D)                     # :259
    with torch.no_grad():                                                # larger than U(-.5/D, .5/D): gradients that matter
        model.ivectors.weight.mul_(20.0)
        model.ovectors.weight.mul_(20.0)
    counts = rng.integers(1, 50, V).astype(np.float64)
    weights = counts / counts.sum() if weighted else None
    sgns = SGNS(embedding=model, vocab_size=V, n_negs=N, weights=weights)   # :261
    optimizer = torch.optim.Adam(sgns.parameters())                      # :266
    iv0, ov0 = model.ivectors.weight.detach().numpy().copy(), model.ovectors.weight.detach().numpy().copy()
    iws, ows, nws, losses, ivs, ovs = [], [], [], [], [], []
    for k in range(steps):
        iw = rng.integers(1, V // 2 if k % 2 else V, B)                  # odd steps leave half of the rows untouched
        ow = rng.integers(0, V, (B, C))                                  # 0 = padding (short windows), as the corpus has it
        iw[: B // 4] = 3                                                 # a hot centre item
        iword, owords = torch.from_numpy(iw).long(), torch.from_numpy(ow).long()
        s = 1000 + k
        torch.manual_seed(s)                                             # the draw of :86-91, recorded
        if weighted:
            nw = torch.multinomial(sgns.weights, B * C * N, replacement=True).view(B, -1)
        else:
            nw = torch.FloatTensor(B, C * N).uniform_(0, V - 1).long()
        torch.manual_seed(s)
        loss = sgns(iword, owords)                                       # :274-277 verbatim
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        iws.append(iw); ows.append(ow); nws.append(nw.numpy().copy()); losses.append(float(loss))
        ivs.append(model.ivectors.weight.detach().numpy().copy()); ovs.append(model.ovectors.weight.detach().numpy().copy())
    return dict(iv0=iv0, ov0=ov0, iword=np.stack(iws), owords=np.stack(ows), nwords=np.stack(nws),
                losses=np.array(losses), iv=np.stack(ivs), ov=np.stack(ovs))


def main():
    out = {}
    for name, weighted in (("w", True), ("u", False)):
        r = run(weighted)
        out.update({f"{name}_{k}": v for k, v in r.items()})
        print(name, "losses", r["losses"])
    np.savez_compressed(os.path.join(HERE, "sgns_small.npz"), **out)
    print("wrote sgns_small.npz")


if __name__ == "__main__":
    main()
