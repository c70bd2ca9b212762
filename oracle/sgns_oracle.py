"""CPU oracle (TEST INFRASTRUCTURE, never imported by the product) for the next row of SURVEY.md section 8(f): N4,
Item2Vec / skip-gram with negative sampling -- 1 centre row, C context rows and C * n_negs negative rows per example.
No kernel is built on it yet; it is here, pinned to the unmodified reference, so that the kernel
has a checker.

Restates, in closed form (numpy, float64 by default):
  * Item2Vec.forward_i / forward_o     (Item2VecRecommender.py:60-68): row gathers from ivectors / ovectors
  * SGNS.forward                       (:82-97) with the negatives `nwords` GIVEN (the reference draws them inside
    forward from torch's global generator, :86-91; the golden script records the draw):
        loss = -mean_b [ mean_c log sigma(o_bc . i_b)  +  mean_c sum_n log sigma(-n_bcn . i_b) ]
  * loss.backward()                    (:276): dense gradients, repeated rows accumulate; row 0 of both tables is the
    padding row (nn.Embedding(padding_idx=0), :40-41) and receives no gradient
  * optim.Adam(sgns.parameters())      (:266, :277): torch defaults (lr 1e-3, betas .9/.999, eps 1e-8), dense --
    every row of both tables is stepped every batch (oracle/gmf_oracle.py: adam_update)
Pinned by tests/test_oracle_golden.py against tests/golden/sgns_small.npz (tests/golden/make_sgns_golden.py runs the
reference's own Item2Vec + SGNS classes and torch.optim.Adam).
"""
import numpy as np

from .gmf_oracle import adam_update


def _log_sigmoid(x):
    return -(np.maximum(-x, 0.0) + np.log1p(np.exp(-np.abs(x))))       # overflow-safe log(1 / (1 + exp(-x)))


def _sigmoid(x):
    return 0.5 * (1.0 + np.tanh(0.5 * x))


def sgns_loss_grads(iv, ov, iword, owords, nwords, padding_idx=0):
    """(loss, g_ivectors, g_ovectors) of SGNS.forward (:82-97) for given negatives.

    iword [B], owords [B, C], nwords [B, C * n_negs] (the layout of `.view(batch_size, -1)`, :86-87)."""
    iword, owords, nwords = (np.asarray(a, dtype=np.int64) for a in (iword, owords, nwords))
    B, C = owords.shape
    i = iv[iword]                                    # [B, D]
    o = ov[owords]                                   # [B, C, D]
    n = ov[nwords]                                   # [B, C*N, D]
    x = np.einsum("bcd,bd->bc", o, i)
    y = np.einsum("bkd,bd->bk", n, i)
    oloss = _log_sigmoid(x).mean(1)
    nloss = _log_sigmoid(-y).reshape(B, C, -1).sum(2).mean(1)
    loss = float(-(oloss + nloss).mean())
    dx = -_sigmoid(-x) / (B * C)                     # d loss / d x_bc
    dy = _sigmoid(y) / (B * C)                       # d loss / d y_bk
    gi, go = np.zeros_like(iv), np.zeros_like(ov)
    np.add.at(gi, iword, np.einsum("bc,bcd->bd", dx, o) + np.einsum("bk,bkd->bd", dy, n))
    np.add.at(go, owords.reshape(-1), (dx[:, :, None] * i[:, None, :]).reshape(-1, iv.shape[1]))
    np.add.at(go, nwords.reshape(-1), (dy[:, :, None] * i[:, None, :]).reshape(-1, iv.shape[1]))
    if padding_idx is not None:                      # embedding backward skips the padding row
        gi[padding_idx] = 0.0
        go[padding_idx] = 0.0
    return loss, gi, go


class SGNSAdam:
    """Tables + Adam moments of the reference loop; step() == Item2VecRecommender.py:274-277 with given negatives."""

    def __init__(self, ivectors, ovectors, lr=1e-3, dtype=np.float64):
        self.iv, self.ov = np.array(ivectors, dtype=dtype), np.array(ovectors, dtype=dtype)
        self.lr, self.t = lr, 0
        self.m = [np.zeros_like(self.iv), np.zeros_like(self.ov)]
        self.v = [np.zeros_like(self.iv), np.zeros_like(self.ov)]

    def step(self, iword, owords, nwords):
        loss, gi, go = sgns_loss_grads(self.iv, self.ov, iword, owords, nwords)
        self.t += 1
        adam_update(self.iv, gi, self.m[0], self.v[0], self.t, self.lr)
        adam_update(self.ov, go, self.m[1], self.v[1], self.t, self.lr)
        return loss
