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
