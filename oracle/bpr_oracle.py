"""CPU oracle for the BPR-MF training step and its HR/NDCG evaluation.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Not imported by the product.

Two independent restatements of the same arithmetic are kept on purpose:

* ``bpr_step_closed_form``: numpy, any float dtype, the maths of SURVEY.md
  section 3.2 written out by hand (gather, score, sigmoid coefficient,
  duplicate-accumulating scatter, dense SGD + L2 over every row).
* ``TorchPort``: the third-party arithmetic the reference actually executes
  (PyTorch ``embedding`` / autograd ``embedding_dense_backward`` /
  ``optim.SGD`` -- ``requirements.txt:3`` pins ``pytorch>=1.0.1``; the build image
  has torch 2.11.0) driven through the reference's call sequence
  ``zero_grad -> forward -> -(pi-pj).sigmoid().log().sum() -> backward -> step``
  (``BPRMFRecommender.py:172-176``).  This is also the CPU baseline timed by
  ``bench.py`` (kind "port": the reference is Python and does not travel to the
  GPU box).

Both are pinned against the unmodified reference classes by
``tests/golden/make_golden.py`` -> ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------
# forward / loss                                   BPRMFRecommender.py:42-50,174
# --------------------------------------------------------------------------
def bpr_scores(P, Q, u, i, j):
    """pred_i, pred_j = <P[u],Q[i]>, <P[u],Q[j]>   (BPRMFRecommender.py:43-50)."""
    pu = P[u]
    return (pu * Q[i]).sum(-1), (pu * Q[j]).sum(-1)


def bpr_loss(pred_i, pred_j):
    """sum_t -log sigmoid(pred_i - pred_j)   (BPRMFRecommender.py:174, SUM reduction).

    Written in the overflow-safe softplus form; for the |x| the reference can
    reach (std 0.01 init) it is the same number.
    """
    x = (pred_i - pred_j).astype(np.float64)
    return float(np.logaddexp(0.0, -x).sum())


# --------------------------------------------------------------------------
# one optimisation step                            BPRMFRecommender.py:154,172-176
# --------------------------------------------------------------------------
def bpr_step_closed_form(P, Q, triples, lr, wd, dtype=np.float64):
    """One reference step on a batch of (u, i, j) triples, closed form.

    grads are taken at the PRE-step tables, repeated rows accumulate
    (``embedding_dense_backward``), and SGD with ``weight_decay`` touches every
    row:  W <- W - lr * (dW + wd * W)      (torch.optim.SGD, BPRMFRecommender.py:154).

    Returns (P_new, Q_new, loss) with the tables in ``dtype``.
    """
    P = np.asarray(P, dtype=dtype)
    Q = np.asarray(Q, dtype=dtype)
    t = np.asarray(triples).reshape(-1, 3)
    u, i, j = t[:, 0], t[:, 1], t[:, 2]
    pu, qi, qj = P[u], Q[i], Q[j]
    x = (pu * qi).sum(-1) - (pu * qj).sum(-1)
    # d/dx -log sigmoid(x) = -(1 - sigmoid(x)) = -sigmoid(-x)
    s = (1.0 / (1.0 + np.exp(x.astype(np.float64)))).astype(dtype)
    gP = np.zeros_like(P)
    gQ = np.zeros_like(Q)
    np.add.at(gP, u, -s[:, None] * (qi - qj))
    np.add.at(gQ, i, -s[:, None] * pu)
    np.add.at(gQ, j, s[:, None] * pu)
    lr = dtype(lr)
    wd = dtype(wd)
    P_new = P - lr * (gP + wd * P)
    Q_new = Q - lr * (gQ + wd * Q)
    loss = float(np.logaddexp(0.0, -x.astype(np.float64)).sum())
    return P_new, Q_new, loss


def bpr_run_closed_form(P, Q, batches, lr, wd, dtype=np.float64):
    """Apply ``bpr_step_closed_form`` over a list of batches; returns tables and per-step losses."""
    losses = []
    for b in batches:
        P, Q, l = bpr_step_closed_form(P, Q, b, lr, wd, dtype)
        losses.append(l)
    return P, Q, losses


class TorchPort:
    """The reference's CPU training path, restated on the installed PyTorch.

    Two embedding tables of shape [user_num, D] / [item_num, D] (the role of
    ``BPR.embed_user`` / ``BPR.embed_item``, BPRMFRecommender.py:36-37), dense
    gradients, ``optim.SGD(lr, weight_decay)`` (``:154``).  ``step`` performs exactly
    the five calls of ``:172-176``.
    """

    def __init__(self, P0, Q0, lr, wd, threads=None):
        import torch
        self.torch = torch
        if threads:
            torch.set_num_threads(int(threads))
        as_t = lambda a: (a.detach().clone() if isinstance(a, torch.Tensor)
                          else torch.from_numpy(np.array(a, dtype=np.float32, copy=True)))
        self.P = as_t(P0).float().requires_grad_(True)
        self.Q = as_t(Q0).float().requires_grad_(True)
        self.opt = torch.optim.SGD([self.P, self.Q], lr=lr, weight_decay=wd)

    def forward(self, u, i, j):
        F = self.torch.nn.functional
        pu = F.embedding(u, self.P)
        return (pu * F.embedding(i, self.Q)).sum(dim=-1), (pu * F.embedding(j, self.Q)).sum(dim=-1)

    def step(self, triples):
        torch = self.torch
        t = torch.as_tensor(np.asarray(triples)).long().reshape(-1, 3)
        u, i, j = t[:, 0].contiguous(), t[:, 1].contiguous(), t[:, 2].contiguous()
        self.opt.zero_grad()
        pi, pj = self.forward(u, i, j)
        loss = -(pi - pj).sigmoid().log().sum()
        loss.backward()
        self.opt.step()
        return float(loss.detach())

    def tables(self):
        return self.P.detach().numpy().copy(), self.Q.detach().numpy().copy()


# --------------------------------------------------------------------------
# evaluation                                        util/metrics.py:35-66,88-94
# --------------------------------------------------------------------------
def topk_order(scores, k):
    """Indices of the k largest scores ordered (score desc, position asc).

    ``torch.topk``'s tie order is implementation defined; the
    product defines it as position-ascending and so does this oracle.
    """
    s = np.asarray(scores)
    order = np.lexsort((np.arange(s.shape[-1]), -s.astype(np.float64)))
    return order[:k]


def candidate_scores(P, Q, user, cands):
    """score[c] = <P[user], Q[cands[c]]> in the dtype of the tables (util/metrics.py:58)."""
    return (P[user][None, :] * Q[cands]).sum(-1)


def bpr_topk_eval(P, Q, users, cands, top_k, scores=None):
    """HR@k / NDCG@k over groups of ``1 positive + C-1 negatives`` (util/metrics.py:46-66).

    ``users`` [N]; ``cands`` [N, C] with the held-out positive in column 0
    (``gt_item = item_i[0]``, util/metrics.py:61).  ``scores`` (optional, [N, C])
    replaces the recomputed scores so that top-K can be checked bit-exactly on
    the scores another implementation produced.

    Returns (HR, NDCG, topk_items [N, k]).
    """
    users = np.asarray(users)
    cands = np.asarray(cands)
    n = users.shape[0]
    hr = np.zeros(n)
    ndcg = np.zeros(n)
    top = np.zeros((n, top_k), dtype=np.int64)
    for g in range(n):
        sc = candidate_scores(P, Q, users[g], cands[g]) if scores is None else scores[g]
        idx = topk_order(sc, top_k)
        rec = cands[g][idx]
        top[g] = rec
        gt = cands[g][0]
        hit = np.nonzero(rec == gt)[0]
        if hit.size:                                   # _hit / _ndcg, util/metrics.py:35-44
            hr[g] = 1.0
            ndcg[g] = 1.0 / np.log2(hit[0] + 2.0)
    return float(hr.mean()), float(ndcg.mean()), top


def full_topk(P, Q, users, k, exclude=None, dtype=np.float32):
    """Full-catalogue top-k: score every item for each user, order (score desc, item asc).

    ``exclude``: optional list (per user) of item ids to mask out (the user's
    training positives).  Returns (items [N,k] int64, scores [N,k]).
    """
    P = np.asarray(P, dtype=dtype)
    Q = np.asarray(Q, dtype=dtype)
    items = np.zeros((len(users), k), dtype=np.int64)
    vals = np.zeros((len(users), k), dtype=dtype)
    for n, u in enumerate(users):
        sc = Q @ P[u]
        if exclude is not None and len(exclude[n]):
            sc = sc.copy()
            sc[np.asarray(exclude[n], dtype=np.int64)] = -np.inf
        idx = topk_order(sc, k)
        items[n] = idx
        vals[n] = sc[idx]
    return items, vals


# --------------------------------------------------------------------------
# lazy sparse Adam (no Daisy counterpart: BPR-MF uses SGD only,
# BPRMFRecommender.py:154; semantics = torch.optim.SparseAdam on
# nn.Embedding(sparse=True), i.e. only rows present in the batch are touched)
# --------------------------------------------------------------------------
def bpr_adam_step_closed_form(P, Q, mP, vP, mQ, vQ, triples, step, lr,
                              beta1=0.9, beta2=0.999, eps=1e-8, dtype=np.float64):
    """One lazy (sparse) Adam step on the BPR loss; ``step`` is 1-based.

    Rows that do not occur in the batch keep weights AND moments unchanged
    (torch.optim.SparseAdam).  Returns new (P, Q, mP, vP, mQ, vQ, loss).
    """
    P = np.asarray(P, dtype=dtype); Q = np.asarray(Q, dtype=dtype)
    mP = np.array(mP, dtype=dtype); vP = np.array(vP, dtype=dtype)
    mQ = np.array(mQ, dtype=dtype); vQ = np.array(vQ, dtype=dtype)
    t = np.asarray(triples).reshape(-1, 3)
    u, i, j = t[:, 0], t[:, 1], t[:, 2]
    pu, qi, qj = P[u], Q[i], Q[j]
    x = (pu * (qi - qj)).sum(-1)
    s = (1.0 / (1.0 + np.exp(x.astype(np.float64)))).astype(dtype)
    gP = np.zeros_like(P); gQ = np.zeros_like(Q)
    np.add.at(gP, u, -s[:, None] * (qi - qj))
    np.add.at(gQ, i, -s[:, None] * pu)
    np.add.at(gQ, j, s[:, None] * pu)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    step_size = lr * np.sqrt(bc2) / bc1          # torch/optim/_functional.py: sparse_adam

    def upd(W, m, v, g, rows):
        rows = np.unique(rows)
        W = W.copy()
        m[rows] = m[rows] + (1 - beta1) * (g[rows] - m[rows])
        v[rows] = v[rows] + (1 - beta2) * (g[rows] * g[rows] - v[rows])
        denom = np.sqrt(v[rows]) + eps           # eps is not bias-corrected in SparseAdam (dense Adam divides it in)
        W[rows] = W[rows] - step_size * m[rows] / denom
        return W, m, v

    Pn, mP, vP = upd(P, mP, vP, gP, u)
    Qn, mQ, vQ = upd(Q, mQ, vQ, gQ, np.concatenate([i, j]))
    loss = float(np.logaddexp(0.0, -x.astype(np.float64)).sum())
    return Pn, Qn, mP, vP, mQ, vQ, loss
