"""Deterministic host-side triple sampler (replaces ``BPRData.ng_sample`` + DataLoader shuffle).

The reference draws, for every training positive ``(u, i)``, ``num_ng`` negatives
``j ~ U[0, item_num)`` re-drawn while ``(u, j)`` is a training positive
(util/data_loader.py:680-690), then lets ``DataLoader(shuffle=True)`` permute the
``num_ng * |train|`` triples into batches (BPRMFRecommender.py:141-142).  Both use
global, unseeded RNG state.  This sampler produces triples with the same
distribution from a counter-based Philox stream keyed by ``(seed, epoch)``, fully
vectorised, and hands out packed ``int32 [N, 3]`` batches that are fed
*identically* to the CUDA path and to the CPU oracle.
"""
from __future__ import annotations

import numpy as np


def _rng(seed, *stream):
    """Philox generator keyed by (seed, stream...) -- independent streams per epoch / purpose."""
    key = np.random.SeedSequence([int(seed), *[int(s) for s in stream]]).generate_state(2, dtype=np.uint64)
    return np.random.Generator(np.random.Philox(key=key))


class TripleSampler:
    """``(u, i, j)`` triples for BPR training.

    Parameters mirror ``BPRData(features, num_item, train_mat, num_ng, is_training=True)``
    (util/data_loader.py:668): ``train_pairs`` are the ``features`` (``[N, 2]`` ints), the
    rejection set is the pair set itself (the role of ``train_mat``).
    """

    def __init__(self, train_pairs, item_num, num_ng=4, seed=2019, reject=True):
        p = np.ascontiguousarray(np.asarray(train_pairs)[:, :2], dtype=np.int64)
        self.users = p[:, 0].copy()
        self.items = p[:, 1].copy()
        self.item_num = int(item_num)
        self.num_ng = int(num_ng)
        self.seed = int(seed)
        self.reject = bool(reject)
        # sorted 64-bit keys u * item_num + i: membership test == "(u, j) in train_mat"
        self._keys = np.unique(self.users * self.item_num + self.items)

    def __len__(self):
        return self.num_ng * len(self.users)

    def _is_positive(self, u, j):
        k = u * self.item_num + j
        pos = np.searchsorted(self._keys, k)
        pos[pos == len(self._keys)] = 0
        return self._keys[pos] == k

    def sample_epoch(self, epoch, shuffle=True):
        """All triples of one epoch, ``int32 [num_ng * |train|, 3]``.

        Un-shuffled order is the reference's ``features_fill`` order (positive-major,
        ``num_ng`` consecutive negatives each, util/data_loader.py:684-690); ``shuffle``
        applies one epoch-keyed permutation (DataLoader ``shuffle=True``).
        """
        g = _rng(self.seed, 1, epoch)
        u = np.repeat(self.users, self.num_ng)
        i = np.repeat(self.items, self.num_ng)
        j = g.integers(0, self.item_num, size=u.shape[0], dtype=np.int64)
        if self.reject:
            bad = np.nonzero(self._is_positive(u, j))[0]
            while bad.size:                       # re-draw only the rejected ones, in index order
                j[bad] = g.integers(0, self.item_num, size=bad.size, dtype=np.int64)
                bad = bad[self._is_positive(u[bad], j[bad])]
        out = np.empty((u.shape[0], 3), dtype=np.int32)
        out[:, 0], out[:, 1], out[:, 2] = u, i, j
        if shuffle:
            out = out[_rng(self.seed, 2, epoch).permutation(out.shape[0])]
        return np.ascontiguousarray(out)

    def batches(self, epoch, batch_size, shuffle=True):
        """Yield consecutive ``int32 [<=batch_size, 3]`` batches; the last one may be short (drop_last=False)."""
        t = self.sample_epoch(epoch, shuffle)
        for s in range(0, t.shape[0], batch_size):
            yield t[s:s + batch_size]


class DeviceTripleSampler:
    """The same job on the GPU (``daisy_sample_triples``, csrc/sampler.cu): one epoch of
    ``(u, i, j)`` triples is produced directly in device memory -- no host sampling, no H2D copy.

    Same constructor as ``TripleSampler``.  Deterministic: every draw is Philox4x32-10 keyed by
    ``(seed, epoch, slot, attempt)``; the rule (a different stream than the host sampler's numpy generator, same
    distribution and order conventions) is restated in ``oracle/sampler_oracle.py`` and checked bit for bit.
    No CPU fallback.
    """

    def __init__(self, train_pairs, item_num, user_num=None, num_ng=4, seed=2019, reject=True, device="cuda"):
        import torch
        from . import _lib
        _lib.require_cuda()
        self._lib, self._torch = _lib, torch
        self.device = torch.device(device)
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        p = np.ascontiguousarray(np.asarray(train_pairs)[:, :2], dtype=np.int64)
        self.item_num, self.num_ng, self.seed = int(item_num), int(num_ng), int(seed)
        self.user_num = int(user_num) if user_num is not None else int(p[:, 0].max()) + 1 if len(p) else 1
        self.n_pairs = len(p)
        self.pairs = torch.from_numpy(p.astype(np.int32)).to(self.device)
        keys = np.unique(p[:, 0] * self.item_num + p[:, 1]) if reject else np.zeros(0, np.int64)
        self.pos_keys = torch.from_numpy(keys).to(self.device)
        self.h = _lib.Handle(idx, self.user_num, self.item_num, 4, 0)

    def __len__(self):
        return self.num_ng * self.n_pairs

    def sample_epoch(self, epoch, shuffle=True, out=None):
        """int32 [num_ng * |train|, 3] on the device (asynchronous on the current stream)."""
        torch, _lib = self._torch, self._lib
        n = len(self)
        if out is None:
            out = torch.empty((n, 3), dtype=torch.int32, device=self.device)
        vp = _lib.c_vp
        _lib.check(self.h.L.daisy_sample_triples(self.h.ptr, vp(self.pairs.data_ptr()), self.n_pairs, self.num_ng,
                                                 vp(self.pos_keys.data_ptr()) if self.pos_keys.numel() else None,
                                                 int(self.pos_keys.numel()), self.seed, int(epoch), int(bool(shuffle)),
                                                 vp(out.data_ptr()), _lib.stream_ptr(torch, self.device)))
        return out

    def check(self):
        self._lib.check(self.h.L.daisy_check(self.h.ptr, self._lib.stream_ptr(self._torch, self.device)))


class SampleSampler(TripleSampler):
    """``(user, item, label)`` samples for pointwise training -- the deterministic stand-in for ``NCFData.ng_sample`` +
    ``DataLoader(shuffle=True)`` (util/data_loader.py:945-960, NCFRecommender.py:240-241): every training pair once
    with label 1, ``num_ng`` negatives per pair with label 0 (uniform items, re-drawn while ``(u, j)`` is a training
    pair), one epoch-keyed permutation.  Fed identically to the reference loop and to the CUDA path."""

    def __len__(self):
        return (self.num_ng + 1) * len(self.users)

    def sample_epoch(self, epoch, shuffle=True):
        tri = super().sample_epoch(epoch, shuffle=False)                       # (u, i, j), positive-major
        n_pos = len(self.users)
        out = np.empty(((self.num_ng + 1) * n_pos, 3), dtype=np.int32)
        out[:n_pos, 0], out[:n_pos, 1], out[:n_pos, 2] = self.users, self.items, 1      # features_ps, labels_ps
        out[n_pos:, 0], out[n_pos:, 1], out[n_pos:, 2] = tri[:, 0], tri[:, 2], 0        # features_ng, labels_ng
        if shuffle:
            out = out[_rng(self.seed, 4, epoch).permutation(out.shape[0])]
        return np.ascontiguousarray(out)


# --------------------------------------------------------------------------
# synthetic workloads of BASELINE.json configs 2-5
# --------------------------------------------------------------------------
def zipf_items(g, n, item_num, exponent=1.0, perm_seed=None):
    """``n`` item ids with P(rank r) ~ 1/r^exponent over a seeded permutation of the catalogue.

    Inverse-CDF on the exact cumulative weights for catalogues up to 2^25 items.
    """
    cdf = _zipf_cdf(int(item_num), float(exponent))
    ranks = np.searchsorted(cdf, g.random(n), side="right").astype(np.int64)
    np.minimum(ranks, item_num - 1, out=ranks)
    if perm_seed is None:
        return ranks
    return _item_perm(int(perm_seed), int(item_num))[ranks]


_CACHE = {}


def _zipf_cdf(item_num, exponent):
    key = ("cdf", item_num, exponent)
    if key not in _CACHE:
        w = 1.0 / np.power(np.arange(1, item_num + 1, dtype=np.float64), exponent)
        cdf = np.cumsum(w)
        cdf /= cdf[-1]
        _CACHE[key] = cdf
    return _CACHE[key]


def _item_perm(seed, item_num):
    key = ("perm", seed, item_num)
    if key not in _CACHE:
        _CACHE[key] = _rng(seed, 3).permutation(item_num)
    return _CACHE[key]


def synthetic_triples(n, user_num, item_num, seed=2019, stream=0, zipf=1.0, permute_items=True):
    """Config 3/4/5 triples: users uniform, positives Zipf(zipf) over a permuted catalogue, negatives uniform
    without rejection (collision probability ~1e-6 at these sizes; both implementations get the same triples)."""
    g = _rng(seed, 10, stream)
    out = np.empty((n, 3), dtype=np.int32)
    out[:, 0] = g.integers(0, user_num, size=n, dtype=np.int64)
    out[:, 1] = zipf_items(g, n, item_num, zipf, perm_seed=seed if permute_items else None)
    out[:, 2] = g.integers(0, item_num, size=n, dtype=np.int64)
    return out


def synthetic_ratings(n, user_num, item_num, seed=2019, zipf=1.0):
    """Config 2 ratings: users uniform, items Zipf over a permutation, ratings 1..5 with the ml-100k histogram."""
    g = _rng(seed, 11)
    users = g.integers(0, user_num, size=n, dtype=np.int64).astype(np.int32)
    items = zipf_items(g, n, item_num, zipf, perm_seed=seed).astype(np.int32)
    ratings = g.choice(np.arange(1, 6, dtype=np.float64), size=n, p=[0.061, 0.114, 0.271, 0.342, 0.212])
    return users, items, ratings
