"""SVD++ behind Daisy's Cython class, computed by libdaisy_b200.so on a B200.

GPU-verified in round 2 (tests/test_svdpp_gpu.py, part of the default ``-m gpu`` suite); csrc/svdpp.cu also reproduces
the reference's golden run under the host emulation of tests/emu.

Drop-in for ``util.matrix_factorization.SVDpp`` (util/matrix_factorization.pyx:169-288; call site
SVDppRecommender.py:142): same constructor keywords, ``fit(train_set)`` on a DataFrame with ``user, item, rating``
columns (returns ``None`` and sets ``pu/qi/yj/bu/bi/global_mean/ur``), ``predict(u, i)`` raising
``ValueError('Invalid user code' / 'Invalid item code')``.

``fit`` keeps the reference's strictly sequential per-rating semantics in float64 (csrc/svdpp.cu).  Initial factors are
drawn exactly like the reference does -- ``np.random.normal`` from the GLOBAL numpy RNG in the order pu, qi, yj
(:221-224; ``random_state`` is stored and ignored) -- so seeding numpy reproduces the reference's start.
"""
from __future__ import annotations

import ctypes
from collections import defaultdict

import numpy as np
import torch

from . import _lib
from ._lib import c_vp, SVDppParams
from .mf import _columns, _predict_many


def user_histories(users, items, user_num):
    """``ur`` of SVDpp.fit (:231-234) in CSR form.  Returns (ptr int64 [U+1], idx int32, mult int32 or None): per user
    the items of its ratings in frame order; ``mult[k]`` = occurrences of ``idx[k]`` in its user's list when ``k`` is the
    first one, else 0 (``None`` when no list holds an item twice) -- see daisy_svdpp_fit in include/daisy_b200.h."""
    users = np.asarray(users, dtype=np.int64)
    items = np.asarray(items, dtype=np.int64)
    if len(users) and (users.min() < 0 or users.max() >= user_num):
        raise ValueError('Invalid user code')
    order = np.argsort(users, kind="stable")
    ptr = np.zeros(user_num + 1, dtype=np.int64)
    np.add.at(ptr, users + 1, 1)
    ptr = np.cumsum(ptr)
    idx = np.ascontiguousarray(items[order], dtype=np.int32)
    su = users[order]
    span = int(items.max()) + 1 if len(items) else 1
    key = su * span + idx                                   # (user, item) as one integer
    by_key = np.argsort(key, kind="stable")                 # equal keys stay in list order
    ks = key[by_key]
    first = np.ones(len(ks), dtype=bool)
    first[1:] = ks[1:] != ks[:-1]
    if first.all():
        return ptr, idx, None
    starts = np.flatnonzero(first)
    counts = np.diff(np.append(starts, len(ks)))
    mult = np.zeros(len(ks), dtype=np.int32)
    mult[by_key[starts]] = counts
    return ptr, idx, mult


class SVDpp(object):
    """util/matrix_factorization.pyx:169-288."""

    def __init__(self, user_num, item_num, n_factors=20, n_epochs=20, init_mean=0, init_std_dev=.1,
                 lr_all=.007, reg_all=.02, lr_bu=None, lr_bi=None, lr_pu=None, lr_qi=None, lr_yj=None,
                 reg_bu=None, reg_bi=None, reg_pu=None, reg_qi=None, reg_yj=None, random_state=None, verbose=True,
                 device="cuda"):
        self.user_num = user_num
        self.item_num = item_num
        self.n_factors = n_factors
        self.n_epochs = n_epochs
        self.init_mean = init_mean
        self.init_std_dev = init_std_dev
        self.lr_bu = lr_bu if lr_bu is not None else lr_all
        self.lr_bi = lr_bi if lr_bi is not None else lr_all
        self.lr_pu = lr_pu if lr_pu is not None else lr_all
        self.lr_qi = lr_qi if lr_qi is not None else lr_all
        self.lr_yj = lr_yj if lr_yj is not None else lr_all
        self.reg_bu = reg_bu if reg_bu is not None else reg_all
        self.reg_bi = reg_bi if reg_bi is not None else reg_all
        self.reg_pu = reg_pu if reg_pu is not None else reg_all
        self.reg_qi = reg_qi if reg_qi is not None else reg_all
        self.reg_yj = reg_yj if reg_yj is not None else reg_all
        self.random_state = random_state
        self.verbose = verbose
        self.device = device

    def _params(self, global_mean):
        return SVDppParams(lr_bu=self.lr_bu, lr_bi=self.lr_bi, lr_pu=self.lr_pu, lr_qi=self.lr_qi, lr_yj=self.lr_yj,
                           reg_bu=self.reg_bu, reg_bi=self.reg_bi, reg_pu=self.reg_pu, reg_qi=self.reg_qi,
                           reg_yj=self.reg_yj, global_mean=float(global_mean))

    def fit(self, train_set):
        users, items, ratings = _columns(train_set)
        global_mean = float(ratings.mean())
        bu = np.zeros(self.user_num, np.double)
        bi = np.zeros(self.item_num, np.double)
        pu = np.random.normal(self.init_mean, self.init_std_dev, size=(self.user_num, self.n_factors))
        qi = np.random.normal(self.init_mean, self.init_std_dev, size=(self.item_num, self.n_factors))
        yj = np.random.normal(self.init_mean, self.init_std_dev, size=(self.item_num, self.n_factors))
        self.global_mean = global_mean
        self._fit_arrays(users, items, ratings, pu, qi, yj, bu, bi)
        if self.verbose:
            for e in range(self.n_epochs):
                print(f'processing epoch {e + 1}')

    def _fit_arrays(self, users, items, ratings, pu, qi, yj, bu, bi):
        """The device part of ``fit`` on given start tables (float64 host arrays, updated in place)."""
        _lib.require_cuda()
        if len(items) and (items.min() < 0 or items.max() >= self.item_num):
            raise ValueError('Invalid item code')
        ptr, idx, mult = user_histories(users, items, self.user_num)
        self._ur_ptr, self._ur_idx, self._ur_rating = ptr, idx, ratings[np.argsort(users, kind="stable")]
        self._ur = None
        dev = torch.device(self.device)
        di = dev.index if dev.index is not None else torch.cuda.current_device()
        h = _lib.Handle(di, self.user_num, self.item_num, self.n_factors, 0)
        try:
            t = lambda a: torch.from_numpy(a if a.flags.writeable else a.copy()).to(dev)
            dpu, dqi, dyj, dbu, dbi = t(pu), t(qi), t(yj), t(bu), t(bi)
            du, dit, dr = t(users), t(items), t(ratings)
            dptr, didx = t(ptr), t(idx)
            dmult = t(mult) if mult is not None else None
            sse = torch.zeros(max(self.n_epochs, 1), dtype=torch.float64, device=dev)
            prm = self._params(self.global_mean)
            s = _lib.stream_ptr(torch, dev)
            _lib.check(h.L.daisy_svdpp_fit(h.ptr, c_vp(dpu.data_ptr()), c_vp(dqi.data_ptr()), c_vp(dyj.data_ptr()),
                                           c_vp(dbu.data_ptr()), c_vp(dbi.data_ptr()), c_vp(du.data_ptr()),
                                           c_vp(dit.data_ptr()), c_vp(dr.data_ptr()), len(ratings), int(self.n_epochs),
                                           c_vp(dptr.data_ptr()), c_vp(didx.data_ptr()),
                                           c_vp(dmult.data_ptr()) if dmult is not None else None,
                                           ctypes.byref(prm), c_vp(sse.data_ptr()), s))
            _lib.check(h.L.daisy_check(h.ptr, s))
            pu[...] = dpu.cpu().numpy(); qi[...] = dqi.cpu().numpy(); yj[...] = dyj.cpu().numpy()
            bu[...] = dbu.cpu().numpy(); bi[...] = dbi.cpu().numpy()
            self.sse_ = sse.cpu().numpy()
        finally:
            h.close()
        self.bu, self.bi, self.pu, self.qi, self.yj = bu, bi, pu, qi, yj

    @property
    def ur(self):
        """The reference's ``self.ur`` (:231-234, :265): ``{user: [(item, rating), ...]}``, built on first use."""
        if self._ur is None:
            ur = defaultdict(list)
            p = self._ur_ptr
            for u in range(self.user_num):
                if p[u + 1] > p[u]:
                    ur[u] = list(zip(self._ur_idx[p[u]:p[u + 1]].tolist(), self._ur_rating[p[u]:p[u + 1]].tolist()))
            self._ur = ur
        return self._ur

    def predict(self, u, i):
        est = self.global_mean
        if u >= self.user_num:
            raise ValueError('Invalid user code')
        if i >= self.item_num:
            raise ValueError('Invalid item code')
        est += self.bu[u] + self.bi[i]
        Iu = self._ur_idx[self._ur_ptr[u]:self._ur_ptr[u + 1]]
        if len(Iu) == 0:
            u_impl_feedback = 0
        else:
            u_impl_feedback = self.yj[Iu].sum(axis=0) / np.sqrt(len(Iu))
        est += np.dot(self.qi[i], self.pu[u] + u_impl_feedback)
        return est

    def user_factors(self):
        """``pu[u] + sum_{j in Iu} yj[j] / sqrt|Iu|`` for every user, on the device (daisy_svdpp_user_factors)."""
        _lib.require_cuda()
        dev = torch.device(self.device)
        di = dev.index if dev.index is not None else torch.cuda.current_device()
        h = _lib.Handle(di, self.user_num, self.item_num, self.n_factors, 0)
        try:
            t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
            dpu, dyj, dptr, didx = t(self.pu), t(self.yj), t(self._ur_ptr), t(self._ur_idx)
            z = torch.empty_like(dpu)
            s = _lib.stream_ptr(torch, dev)
            _lib.check(h.L.daisy_svdpp_user_factors(h.ptr, c_vp(dpu.data_ptr()), c_vp(dyj.data_ptr()),
                                                    c_vp(dptr.data_ptr()), c_vp(didx.data_ptr()), c_vp(z.data_ptr()), s))
            _lib.check(h.L.daisy_check(h.ptr, s))
            return z.cpu().numpy()
        finally:
            h.close()

    def predict_many(self, users, items):
        """Batched ``predict`` on the device (replaces the Python ``predict`` call per candidate of
        SVDppRecommender.py's ranking loop): the histories are summed once per user, then it is funk-SVD's predict."""
        z = self.user_factors()
        return _predict_many(self.user_num, self.item_num, self.n_factors, z, self.qi, self.bu, self.bi, users, items,
                             True, float(self.global_mean), self.device)
