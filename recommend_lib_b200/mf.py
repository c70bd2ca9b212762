"""funk-SVD / RSVD behind Daisy's Cython classes, computed by libdaisy_b200.so on a B200.

Drop-ins for ``util.matrix_factorization.SVD`` / ``RSVD`` (util/matrix_factorization.pyx:81-167, 5-78): same
constructor keywords, ``fit(train_set)`` on a DataFrame with ``user, item, rating`` columns (returns ``None`` and
sets ``pu/qi/bu/bi/global_mean`` resp. ``ui/vj/ci/dj`` as float64 ndarrays), ``predict(u, i)`` raising
``ValueError('Invalid user code' / 'Invalid item code')``.

``fit`` keeps the reference's strictly sequential per-rating semantics (update t+1 sees update t) in float64: the
device kernel is a dataflow over per-row version counters (csrc/mf.cu).  Initial factors are drawn exactly like
the reference does -- ``np.random.normal`` from the GLOBAL numpy RNG, user table first (:38-39, :124-125; the
``random_state`` argument is stored and ignored) -- so seeding numpy reproduces the reference's start.

Deviation (documented): the reference's ``RSVD.fit`` only trains when ``verbose`` is true because the
rating loop is nested under ``if self.verbose`` (:42-44); here ``verbose`` only controls printing.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import c_vp, MFParams


def _columns(train_set):
    """user / item / rating arrays from a DataFrame (or any mapping / structured object with those fields)."""
    get = (lambda k: train_set[k].values) if hasattr(train_set, "columns") else (lambda k: np.asarray(train_set[k]))
    return (np.ascontiguousarray(get("user"), dtype=np.int32), np.ascontiguousarray(get("item"), dtype=np.int32),
            np.ascontiguousarray(get("rating"), dtype=np.float64))


def _device_fit(user_num, item_num, n_factors, n_epochs, prm, users, items, ratings, A, Bm, ba, bb, device, want_sse):
    """Run daisy_mf_fit; A, Bm, ba, bb are float64 host arrays updated with the trained values."""
    _lib.require_cuda()
    dev = torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    h = _lib.Handle(idx, user_num, item_num, n_factors, 0)
    try:
        t = lambda a: torch.from_numpy(a if a.flags.writeable else a.copy()).to(dev)   # read-only views (np.load) warn
        dA, dB, dba, dbb = t(A), t(Bm), t(ba), t(bb)
        du, di, dr = t(users), t(items), t(ratings)
        sse = torch.zeros(max(n_epochs, 1), dtype=torch.float64, device=dev) if want_sse else None
        _lib.check(h.L.daisy_mf_fit(h.ptr, c_vp(dA.data_ptr()), c_vp(dB.data_ptr()), c_vp(dba.data_ptr()),
                                    c_vp(dbb.data_ptr()), c_vp(du.data_ptr()), c_vp(di.data_ptr()),
                                    c_vp(dr.data_ptr()), len(ratings), int(n_epochs), ctypes.byref(prm),
                                    c_vp(sse.data_ptr()) if sse is not None else None,
                                    _lib.stream_ptr(torch, dev)))
        _lib.check(h.L.daisy_check(h.ptr, _lib.stream_ptr(torch, dev)))
        A[...] = dA.cpu().numpy(); Bm[...] = dB.cpu().numpy(); ba[...] = dba.cpu().numpy(); bb[...] = dbb.cpu().numpy()
        return sse.cpu().numpy() if sse is not None else None
    finally:
        h.close()


class SVD(object):
    """util/matrix_factorization.pyx:81-167."""

    def __init__(self, user_num, item_num, n_factors=100, n_epochs=20, biased=True, init_mean=0, init_std_dev=.1,
                 lr_all=.005, reg_all=.02, lr_bu=None, lr_bi=None, lr_pu=None, lr_qi=None, reg_bu=None, reg_bi=None,
                 reg_pu=None, reg_qi=None, random_state=None, verbose=True, device="cuda"):
        self.user_num = user_num
        self.item_num = item_num
        self.n_factors = n_factors
        self.n_epochs = n_epochs
        self.biased = biased
        self.init_mean = init_mean
        self.init_std_dev = init_std_dev
        self.lr_bu = lr_bu if lr_bu is not None else lr_all
        self.lr_bi = lr_bi if lr_bi is not None else lr_all
        self.lr_pu = lr_pu if lr_pu is not None else lr_all
        self.lr_qi = lr_qi if lr_qi is not None else lr_all
        self.reg_bu = reg_bu if reg_bu is not None else reg_all
        self.reg_bi = reg_bi if reg_bi is not None else reg_all
        self.reg_pu = reg_pu if reg_pu is not None else reg_all
        self.reg_qi = reg_qi if reg_qi is not None else reg_all
        self.random_state = random_state
        self.verbose = verbose
        self.device = device

    def fit(self, train_set):
        users, items, ratings = _columns(train_set)
        bu = np.zeros(self.user_num)
        bi = np.zeros(self.item_num)
        pu = np.random.normal(self.init_mean, self.init_std_dev, size=(self.user_num, self.n_factors))
        qi = np.random.normal(self.init_mean, self.init_std_dev, size=(self.item_num, self.n_factors))
        global_mean = float(ratings.mean()) if self.biased else 0
        self.global_mean = global_mean
        prm = MFParams(variant=0, biased=int(bool(self.biased)), lr_bu=self.lr_bu, lr_bi=self.lr_bi, lr_pu=self.lr_pu,
                       lr_qi=self.lr_qi, reg_bu=self.reg_bu, reg_bi=self.reg_bi, reg_pu=self.reg_pu,
                       reg_qi=self.reg_qi, reg2=0.0, global_mean=float(global_mean))
        self.sse_ = _device_fit(self.user_num, self.item_num, self.n_factors, self.n_epochs, prm, users, items, ratings,
                                pu, qi, bu, bi, self.device, True)
        if self.verbose:
            for e in range(self.n_epochs):
                print(f'Processing epoch {e + 1}')
        self.bu, self.bi, self.pu, self.qi = bu, bi, pu, qi

    def predict(self, u, i):
        if u >= self.user_num:
            raise ValueError('Invalid user code')
        if i >= self.item_num:
            raise ValueError('Invalid item code')
        if self.biased:
            return self.global_mean + self.bu[u] + self.bi[i] + np.dot(self.qi[i], self.pu[u])
        return np.dot(self.qi[i], self.pu[u])

    def predict_many(self, users, items):
        """Batched ``predict`` on the device (replaces the 1000 Python ``predict`` calls per user of
        MFRecommender.py:195-217)."""
        return _predict_many(self.user_num, self.item_num, self.n_factors, self.pu, self.qi, self.bu, self.bi, users,
                             items, bool(self.biased), float(self.global_mean), self.device)


class RSVD(object):
    """util/matrix_factorization.pyx:5-78."""

    def __init__(self, user_num, item_num, n_factors=96, n_epochs=20, version=2, init_mean=0, init_std_dev=.1,
                 lr=.001, reg=.02, reg2=.05, random_state=None, verbose=True, device="cuda"):
        self.user_num = user_num
        self.item_num = item_num
        self.n_factors = n_factors
        self.n_epochs = n_epochs
        self.version = version
        self.lr = lr
        self.reg = reg
        self.reg2 = reg2
        self.init_mean = init_mean
        self.init_std_dev = init_std_dev
        self.random_state = random_state
        self.verbose = verbose
        self.device = device

    def fit(self, train_set):
        users, items, ratings = _columns(train_set)
        global_mean = float(ratings.mean())
        ci = np.zeros(self.user_num, np.double)
        dj = np.zeros(self.item_num, np.double)
        ui = np.random.normal(self.init_mean, self.init_std_dev, size=(self.user_num, self.n_factors))
        vj = np.random.normal(self.init_mean, self.init_std_dev, size=(self.item_num, self.n_factors))
        prm = MFParams(variant=2 if self.version == 2 else 1, biased=0, lr_bu=self.lr, lr_bi=self.lr, lr_pu=self.lr,
                       lr_qi=self.lr, reg_bu=0.0, reg_bi=0.0, reg_pu=self.reg, reg_qi=self.reg, reg2=self.reg2,
                       global_mean=global_mean)
        self.sse_ = _device_fit(self.user_num, self.item_num, self.n_factors, self.n_epochs, prm, users, items, ratings,
                                ui, vj, ci, dj, self.device, True)
        if self.verbose:
            for e in range(self.n_epochs):
                print(f'Processing epoch {e + 1}')
        self.ci, self.dj, self.ui, self.vj = ci, dj, ui, vj

    def predict(self, i, j):
        if i >= self.user_num:
            raise ValueError('Invalid user code')
        if j >= self.item_num:
            raise ValueError('Invalid item code')
        if self.version == 2:
            return self.ci[i] + self.dj[j] + np.dot(self.ui[i], self.vj[j])
        elif self.version == 1:
            return np.dot(self.ui[i], self.vj[j])

    def predict_many(self, users, items):
        return _predict_many(self.user_num, self.item_num, self.n_factors, self.ui, self.vj, self.ci, self.dj, users,
                             items, self.version == 2, 0.0, self.device)


def _predict_many(user_num, item_num, n_factors, A, Bm, ba, bb, users, items, with_bias, mu, device):
    _lib.require_cuda()
    dev = torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    h = _lib.Handle(idx, user_num, item_num, n_factors, 0)
    try:
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
        dA, dB, dba, dbb = t(A, np.float64), t(Bm, np.float64), t(ba, np.float64), t(bb, np.float64)
        du, di = t(users, np.int32), t(items, np.int32)
        est = torch.empty(du.shape[0], dtype=torch.float64, device=dev)
        s = _lib.stream_ptr(torch, dev)
        _lib.check(h.L.daisy_mf_predict(h.ptr, c_vp(dA.data_ptr()), c_vp(dB.data_ptr()), c_vp(dba.data_ptr()),
                                        c_vp(dbb.data_ptr()), c_vp(du.data_ptr()), c_vp(di.data_ptr()), du.shape[0],
                                        int(bool(with_bias)), float(mu), c_vp(est.data_ptr()), s))
        rc = h.L.daisy_check(h.ptr, s)
        if rc == _lib.EINDEX:
            u = np.asarray(users)
            raise ValueError('Invalid user code' if (u >= user_num).any() or (u < 0).any() else 'Invalid item code')
        _lib.check(rc)
        return est.cpu().numpy()
    finally:
        h.close()


class MFRecommender:
    """``fit()/predict()`` convenience wrapper around the call sites of MFRecommender.py:146-148,197,214 and
    RSVDRecommender.py:154-157,206,223: ``algo in {'svd', 'rsvd'}``, remaining keywords go to the class."""

    def __init__(self, user_num, item_num, algo="svd", **kw):
        self.algo = SVD(user_num, item_num, **kw) if algo == "svd" else RSVD(user_num, item_num, **kw)

    def fit(self, train_set):
        self.algo.fit(train_set)
        return self

    def predict(self, u, i):
        if np.ndim(u) == 0:
            return self.algo.predict(u, i)
        return self.algo.predict_many(u, i)
