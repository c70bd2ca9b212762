"""CPU oracle (TEST INFRASTRUCTURE, never imported by the product) for BPR-FM with two one-hot features per example
 in the configuration that has a reproducible reference output: ``batch_norm=False``,
``drop_prob=[0, 0]`` (dropout draws from torch's global generator inside forward; batch-norm is orthogonal to the
gather/score/scatter path), optimiser Adagrad -- the script's default (BPRFMRecommender.py:122-125,191-193).

Restates, in closed form (numpy, float64 by default):
  * BPRFM._out (BPRFMRecommender.py:61-80) for features [user_feature, item_feature] with values [1, 1]:
        pred = 0.5 * sum_f ((e_u + e_i)^2 - e_u^2 - e_i^2) + b_u + b_i + bias_  =  <e_u, e_i> + b_u + b_i + bias_
  * loss = -(pred_i - pred_j).sigmoid().log().sum()   (:217)  -- the user bias and bias_ cancel in pred_i - pred_j
  * loss.backward() + optim.Adagrad(lr, initial_accumulator_value=1e-8).step()  (:191-193,218-219): torch defaults
    lr_decay 0, eps 1e-10; an element with zero gradient does not move (Adagrad has no momentum), so the dense
    optimiser equals a sparse one on the touched rows
Pinned by tests/test_oracle_golden.py against tests/golden/bprfm_small.npz (tests/golden/make_bprfm_golden.py).
"""
import numpy as np


def pred(E, bias, bias_, feats):
    """feats int [B,2] (user feature, item feature) -> FM prediction (BPRFMRecommender.py:61-80, values all 1)."""
    eu, ei = E[feats[:, 0]], E[feats[:, 1]]
    return (eu * ei).sum(1) + bias[feats[:, 0]] + bias[feats[:, 1]] + bias_


def bprfm_adagrad_step(E, bias, bias_, accE, accb, feats_i, feats_j, lr=0.05, eps=1e-10):
    """One step, in place on E [N,F], bias [N], accE, accb (Adagrad state_sum); returns the batch-sum loss."""
    u, i, j = feats_i[:, 0], feats_i[:, 1], feats_j[:, 1]
    eu, ei, ej = E[u], E[i], E[j]
    x = (eu * (ei - ej)).sum(1) + bias[i] - bias[j]
    s = 1.0 / (1.0 + np.exp(x))                                # -dloss/dx
    loss = float(np.sum(np.maximum(-x, 0.0) + np.log1p(np.exp(-np.abs(x)))))
    gE, gb = np.zeros_like(E), np.zeros_like(bias)
    np.add.at(gE, u, -s[:, None] * (ei - ej))
    np.add.at(gE, i, -s[:, None] * eu)
    np.add.at(gE, j, s[:, None] * eu)
    np.add.at(gb, i, -s)
    np.add.at(gb, j, s)                                        # the user bias gets +g - g = 0 exactly
    for theta, g, acc in ((E, gE, accE), (bias, gb, accb)):
        acc += g * g
        theta -= lr * g / (np.sqrt(acc) + eps)
    return loss


# ---------------------------------------------------------------------------------------------------------------------
# The script's DEFAULT configuration: batch_norm=True, drop_prob=[0.5, 0.2] (BPRFMRecommender.py:116-125), any number of
# features per example with values.  Dropout masks are GIVEN (the reference draws them inside forward from torch's
# global generator; tests/golden/make_bprfm_bn_golden.py records the draw).  No kernel is built on this yet.
# ---------------------------------------------------------------------------------------------------------------------
class BPRFMFull:
    """Closed form of BPRFM._out + the script's step (BPRFMRecommender.py:57-80, 214-219) with
    nn.BatchNorm1d(num_factors) in training mode (:49-50: batch statistics, biased variance, eps 1e-5; running
    statistics updated with momentum 0.1 and the unbiased variance, once per _out call: positive first, then negative),
    nn.Dropout(drop_prob[0]) as a given mask (kept elements scaled by 1 / (1 - p)), and optim.Adagrad over ALL
    parameters (embeddings, biases, bias_, BN weight / bias)."""

    def __init__(self, E, bias, bias_=0.0, batch_norm=True, lr=0.05, initial_accumulator_value=1e-8, eps=1e-10,
                 dtype=np.float64):
        self.E, self.bias = np.array(E, dtype=dtype), np.array(bias, dtype=dtype).reshape(-1)
        self.bias_ = np.array([float(bias_)], dtype=dtype)
        F = self.E.shape[1]
        self.batch_norm = bool(batch_norm)
        self.gamma, self.beta = np.ones(F, dtype=dtype), np.zeros(F, dtype=dtype)
        self.running_mean, self.running_var = np.zeros(F, dtype=dtype), np.ones(F, dtype=dtype)
        self.lr, self.eps, self.bn_eps, self.momentum = lr, eps, 1e-5, 0.1
        self.params = [self.E, self.bias, self.bias_] + ([self.gamma, self.beta] if self.batch_norm else [])
        self.acc = [np.full_like(p, initial_accumulator_value) for p in self.params]

    # -- forward of one _out call; returns pred and what backward needs --------------------------------------------------
    def _out(self, feats, vals, mask, train):
        ne = self.E[feats] * vals[:, :, None]                      # :62-64  [B, K, F]
        S = ne.sum(1)
        x = 0.5 * (S * S - (ne * ne).sum(1))                       # :67-71  bi-interaction
        ctx = dict(feats=feats, vals=vals, ne=ne, S=S, mask=mask)
        z = x
        if self.batch_norm:
            if train:
                mu, var = x.mean(0), x.var(0)                      # biased variance normalises
                n = x.shape[0]
                self.running_mean += self.momentum * (mu - self.running_mean)
                self.running_var += self.momentum * (var * n / (n - 1) - self.running_var)
            else:
                mu, var = self.running_mean, self.running_var
            inv = 1.0 / np.sqrt(var + self.bn_eps)
            xhat = (x - mu) * inv
            z = self.gamma * xhat + self.beta
            ctx.update(inv=inv, xhat=xhat)
        if train and mask is not None:
            z = z * mask                                           # mask already carries 1 / (1 - p)
        pred = z.sum(1) + (self.bias[feats] * vals).sum(1) + self.bias_[0]     # :72-78
        return pred, ctx

    def forward(self, feats_i, vals_i, feats_j, vals_j):
        """Evaluation-mode (pred_i, pred_j): running statistics, no dropout."""
        return self._out(feats_i, vals_i, None, False)[0], self._out(feats_j, vals_j, None, False)[0]

    def _backward(self, ctx, g, grads):
        """g [B] = d loss / d pred of this call; accumulates into grads (same order as self.params)."""
        feats, vals, ne, S, mask = ctx["feats"], ctx["vals"], ctx["ne"], ctx["S"], ctx["mask"]
        F = self.E.shape[1]
        dz = np.repeat(g[:, None], F, 1)
        if mask is not None:
            dz = dz * mask
        if self.batch_norm:
            xhat, inv = ctx["xhat"], ctx["inv"]
            grads[3] += (dz * xhat).sum(0)
            grads[4] += dz.sum(0)
            dxh = dz * self.gamma
            dx = inv * (dxh - dxh.mean(0) - xhat * (dxh * xhat).mean(0))
        else:
            dx = dz
        dne = dx[:, None, :] * (S[:, None, :] - ne)               # d FM / d ne_k = S - ne_k
        np.add.at(grads[0], feats.reshape(-1), (dne * vals[:, :, None]).reshape(-1, F))
        np.add.at(grads[1], feats.reshape(-1), (g[:, None] * vals).reshape(-1))
        grads[2] += g.sum()

    def step(self, feats_i, vals_i, feats_j, vals_j, mask_i=None, mask_j=None):
        """One training step (:214-219); returns the batch-sum loss."""
        dt = self.E.dtype
        vals_i, vals_j = np.asarray(vals_i, dtype=dt), np.asarray(vals_j, dtype=dt)
        pi, ci = self._out(feats_i, vals_i, mask_i, True)
        pj, cj = self._out(feats_j, vals_j, mask_j, True)
        x = pi - pj
        s = 1.0 / (1.0 + np.exp(x))
        loss = float(np.sum(np.maximum(-x, 0.0) + np.log1p(np.exp(-np.abs(x)))))
        grads = [np.zeros_like(p) for p in self.params]
        self._backward(ci, -s, grads)
        self._backward(cj, s, grads)
        for theta, g, acc in zip(self.params, grads, self.acc):
            acc += g * g
            theta -= self.lr * g / (np.sqrt(acc) + self.eps)
        return loss
