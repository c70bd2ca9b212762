"""Host-side data preparation for the BPR-MF path (numpy; no pandas needed at run time).

Restates the parts of ``load_mat`` that the hot path depends on, with the
reference's defects routed around:

* ``encode_ids``      -- ``pd.Categorical(...).codes``            util/data_loader.py:447-448
* ``split_loo_by_time`` -- ``_split_loo(by_time=1)``               util/data_loader.py:410-414
* ``eval_candidates``  -- ``_negative_sampling`` + test_data layout util/data_loader.py:433-441,461-469
  (positive first, then 999 negatives drawn without replacement from the items the
  user never interacted with; users with fewer than 999 such items are dropped --
  the reference's ``random.sample`` raises for them, D2).
"""
from __future__ import annotations

import numpy as np

from .sampler import _rng


def load_ml100k(path):
    """``u.data``: tab-separated ``user item rating timestamp`` (util/data_loader.py:28-30),
    rows sorted by (user, item, timestamp) like ``load_rate`` does (``:117``)."""
    raw = np.loadtxt(path, dtype=np.int64)
    order = np.lexsort((raw[:, 3], raw[:, 1], raw[:, 0]))
    return raw[order]


def encode_ids(col):
    """Dense 0-based codes in sorted order of the raw ids (pd.Categorical codes)."""
    uniq, codes = np.unique(col, return_inverse=True)
    return codes.astype(np.int64), int(uniq.shape[0])


def split_loo_by_time(users, items, timestamps):
    """Leave-one-out by latest timestamp; ties go to the first row in input order
    (``rank(method='first', ascending=False) == 1``).  Returns (train_idx, test_idx) row indices,
    both in input order."""
    n = users.shape[0]
    # stable sort by (user, -timestamp): first row of each user block is the held-out one
    order = np.lexsort((np.arange(n), -timestamps, users))
    first = np.ones(n, dtype=bool)
    first[1:] = users[order][1:] != users[order][:-1]
    test_mask = np.zeros(n, dtype=bool)
    test_mask[order[first]] = True
    return np.nonzero(~test_mask)[0], np.nonzero(test_mask)[0]


def eval_candidates(all_users, all_items, test_users, test_items, item_num, num_neg=999, seed=2019):
    """Candidate lists for ``metric_eval``: ``[N, 1 + num_neg]`` with the positive in column 0.

    Negatives are uniform without replacement from items the user has no interaction with in
    the FULL frame (train and test).  Users with fewer than ``num_neg`` such items are dropped.
    Returns (users [N], cands [N, 1+num_neg]) as int32.
    """
    g = _rng(seed, 20)
    order = np.argsort(all_users, kind="stable")
    su, si = all_users[order], all_items[order]
    starts = np.searchsorted(su, test_users, side="left")
    ends = np.searchsorted(su, test_users, side="right")
    keep_u, rows = [], []
    for u, it, a, b in zip(test_users, test_items, starts, ends):
        seen = np.zeros(item_num, dtype=bool)
        seen[si[a:b]] = True
        pool = np.nonzero(~seen)[0]
        if pool.size < num_neg:
            continue
        neg = g.choice(pool, size=num_neg, replace=False)
        keep_u.append(u)
        rows.append(np.concatenate([[it], neg]))
    return np.asarray(keep_u, dtype=np.int32), np.asarray(rows, dtype=np.int32)
