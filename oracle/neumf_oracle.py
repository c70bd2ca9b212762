"""CPU oracle (TEST INFRASTRUCTURE, never imported by the product) for NCF with an MLP tower -- ``model='MLP'`` and the
script's default ``model='NeuMF-end'`` (NCFRecommender.py:28-125, 175-178), dropout 0 (the script's default, :136-139).
SURVEY.md section 8f, row N3.  The GMF variant is oracle/gmf_oracle.py.

Restates, in closed form (numpy, float64 by default):
  * NCF.forward (:105-125): GMF branch eu_g * ei_g; MLP branch x0 = [eu_m | ei_m] -> (Linear + ReLU) x num_layers
    (:51-57); concat -> predict_layer -> logit
  * nn.BCEWithLogitsLoss() (:255, mean), loss.backward() (:286): dense gradients, repeated embedding rows accumulate
  * optim.Adam(model.parameters(), lr) (:260, :287): torch defaults, applied to EVERY element of every parameter
    (for model 'MLP' the GMF tables and for every model unused parameters have no gradient: torch skips them)
Pinned by tests/test_oracle_golden.py against tests/golden/neumf_small.npz (tests/golden/make_neumf_golden.py runs the
reference's own NCF class + torch.optim.Adam).
"""
import numpy as np

from .gmf_oracle import adam_update, bce_with_logits_mean


class NeuMFAdam:
    """State of the reference's training loop for model in {'MLP', 'NeuMF-end'}; step() == NCFRecommender.py:283-287."""

    def __init__(self, model, Pg, Qg, Pm, Qm, Ws, bs, wp, bp, lr=1e-3, dtype=np.float64):
        assert model in ("MLP", "NeuMF-end")
        self.model = model
        c = lambda a: np.array(a, dtype=dtype)
        self.Pg, self.Qg, self.Pm, self.Qm = c(Pg), c(Qg), c(Pm), c(Qm)
        self.Ws, self.bs = [c(W) for W in Ws], [c(b).reshape(-1) for b in bs]
        self.wp, self.bp = c(wp).reshape(-1), c(bp).reshape(-1)
        self.lr, self.t = lr, 0
        self.params = ([self.Pg, self.Qg] if model != "MLP" else []) + [self.Pm, self.Qm] + \
            [x for W, b in zip(self.Ws, self.bs) for x in (W, b)] + [self.wp, self.bp]
        self.m = [np.zeros_like(p) for p in self.params]
        self.v = [np.zeros_like(p) for p in self.params]

    def _forward(self, users, items):
        acts = [np.concatenate([self.Pm[users], self.Qm[items]], 1)]           # :113-116
        for W, b in zip(self.Ws, self.bs):
            acts.append(np.maximum(acts[-1] @ W.T + b, 0.0))                    # Dropout(0) + Linear + ReLU
        if self.model == "MLP":
            concat = acts[-1]
        else:
            concat = np.concatenate([self.Pg[users] * self.Qg[items], acts[-1]], 1)   # :107-109, :122-123
        return acts, concat, concat @ self.wp + self.bp[0]

    def forward(self, users, items):
        return self._forward(np.asarray(users, np.int64), np.asarray(items, np.int64))[2]

    def step(self, users, items, labels):
        users, items = np.asarray(users, np.int64), np.asarray(items, np.int64)
        y = np.asarray(labels, dtype=self.Pm.dtype)
        acts, concat, x = self._forward(users, items)
        loss = bce_with_logits_mean(x, y)
        dx = (1.0 / (1.0 + np.exp(-x)) - y) / len(users)
        g_wp, g_bp = concat.T @ dx, np.array([dx.sum()])
        dconcat = dx[:, None] * self.wp[None, :]
        grads = []
        if self.model != "MLP":
            F = self.Pg.shape[1]
            dg, dh = dconcat[:, :F], dconcat[:, F:]
            gPg, gQg = np.zeros_like(self.Pg), np.zeros_like(self.Qg)
            np.add.at(gPg, users, dg * self.Qg[items])
            np.add.at(gQg, items, dg * self.Pg[users])
            grads += [gPg, gQg]
        else:
            dh = dconcat
        layer_grads = []
        for l in range(len(self.Ws) - 1, -1, -1):
            dz = dh * (acts[l + 1] > 0)
            layer_grads.append((dz.T @ acts[l], dz.sum(0)))
            dh = dz @ self.Ws[l]
        layer_grads.reverse()
        Dm = self.Pm.shape[1]
        gPm, gQm = np.zeros_like(self.Pm), np.zeros_like(self.Qm)
        np.add.at(gPm, users, dh[:, :Dm])
        np.add.at(gQm, items, dh[:, Dm:])
        grads += [gPm, gQm] + [g for pair in layer_grads for g in pair] + [g_wp, g_bp]
        self.t += 1
        for theta, g, m, v in zip(self.params, grads, self.m, self.v):
            adam_update(theta, g, m, v, self.t, self.lr)
        return loss
