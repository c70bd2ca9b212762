"""CPU oracle (TEST INFRASTRUCTURE, never imported by the product) for the next row of SURVEY.md section 8(f): N3,
NCF in its GMF variant -- the first member of the gather / score / scatter family after BPR-MF.  No kernel is built on
it yet; it is here, pinned to the unmodified reference, so that the kernel has a checker.

Restates, in closed form (numpy, float64 by default):
  * NCF.forward with model == 'GMF'   (NCFRecommender.py:105-125): logit = w . (P[u] * Q[i]) + b
  * nn.BCEWithLogitsLoss()            (:255, mean reduction), overflow-safe form
  * loss.backward()                   (:286): dense gradients; repeated rows accumulate
  * optim.Adam(lr) .step()            (:260, :287): torch defaults (betas .9/.999, eps 1e-8, no amsgrad, no weight
    decay), applied to EVERY row of both tables every step -- a row without a gradient still moves while its first
    moment decays, which is what a sparse / lazy kernel has to reproduce
Pinned by tests/test_oracle_golden.py against tests/golden/gmf_small.npz (tests/golden/make_ncf_golden.py runs the
reference's own NCF class + torch.optim.Adam).
"""
import numpy as np


def gmf_forward(P, Q, w, b, users, items):
    """logit[t] = sum_f w[f] * P[u_t, f] * Q[i_t, f] + b   (NCFRecommender.py:106-109,122-125)."""
    return (P[users] * Q[items]) @ w + b


def bce_with_logits_mean(x, y):
    """mean_t [ max(x,0) - x*y + log(1 + exp(-|x|)) ]  ==  nn.BCEWithLogitsLoss()(x, y)."""
    return float(np.mean(np.maximum(x, 0.0) - x * y + np.log1p(np.exp(-np.abs(x)))))


def gmf_grads(P, Q, w, b, users, items, labels):
    """Dense gradients of the mean BCE loss; returns (loss, gP, gQ, gw, gb)."""
    pu, qi = P[users], Q[items]
    x = (pu * qi) @ w + b
    y = labels.astype(P.dtype)
    dx = (1.0 / (1.0 + np.exp(-x)) - y) / len(users)          # d loss / d logit
    gP, gQ = np.zeros_like(P), np.zeros_like(Q)
    np.add.at(gP, users, dx[:, None] * (qi * w))
    np.add.at(gQ, items, dx[:, None] * (pu * w))
    gw = (pu * qi).T @ dx
    gb = dx.sum()
    return bce_with_logits_mean(x, y), gP, gQ, gw, gb


def adam_update(theta, g, m, v, t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam single-tensor step, in place (t is 1-based): every element, zero gradient or not."""
    m *= b1
    m += (1.0 - b1) * g
    v *= b2
    v += (1.0 - b2) * g * g
    step_size = lr / (1.0 - b1 ** t)
    denom = np.sqrt(v) / np.sqrt(1.0 - b2 ** t) + eps
    theta -= step_size * m / denom


class GMFAdam:
    """The reference's training loop state: tables, predict layer, Adam moments.  step() == NCFRecommender.py:283-287."""

    def __init__(self, P, Q, w, b, lr=1e-3, dtype=np.float64):
        self.P, self.Q = np.array(P, dtype=dtype), np.array(Q, dtype=dtype)
        self.w, self.b = np.array(w, dtype=dtype).reshape(-1), np.array([float(np.asarray(b).reshape(-1)[0])], dtype=dtype)
        self.lr, self.t = lr, 0
        self.m = [np.zeros_like(a) for a in (self.P, self.Q, self.w, self.b)]
        self.v = [np.zeros_like(a) for a in (self.P, self.Q, self.w, self.b)]

    def step(self, users, items, labels):
        loss, gP, gQ, gw, gb = gmf_grads(self.P, self.Q, self.w, self.b[0], users, items, labels)
        self.t += 1
        for theta, g, m, v in zip((self.P, self.Q, self.w, self.b), (gP, gQ, gw, np.array([gb])), self.m, self.v):
            adam_update(theta, g, m, v, self.t, self.lr)
        return loss
