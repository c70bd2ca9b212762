"""ctypes front-end of the C oracle ``oracle/mf_oracle.c`` (funk-SVD / RSVD).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Not imported by the product.

``svd_fit`` / ``rsvd_fit`` take the initial tables explicitly (the reference
draws them from the global numpy RNG inside ``fit``,
util/matrix_factorization.pyx:38-39,124-125; ``draw_init`` reproduces that draw
order so that seeding numpy the same way gives the same start).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "mf_oracle.c")
_SO = os.path.join(_HERE, "libmf_oracle.so")
_lib = None


def build(force=False):
    """gcc -O2 -ffp-contract=off: no FMA contraction => same rounding as the Cython-generated C."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", _SO, _SRC])
    return _SO


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int32)
        d, i32, i64 = ctypes.c_double, ctypes.c_int, ctypes.c_int64
        L.mf_oracle_svd_fit.restype = d
        L.mf_oracle_svd_fit.argtypes = [i64, ip, ip, dp, i32, i32, i32, d, d, d, d, d, d, d, d, d, dp, dp, dp, dp]
        L.mf_oracle_rsvd_fit.restype = d
        L.mf_oracle_rsvd_fit.argtypes = [i64, ip, ip, dp, i32, i32, i32, d, d, d, d, dp, dp, dp, dp]
        L.mf_oracle_svdpp_fit.restype = d
        L.mf_oracle_svdpp_fit.argtypes = [i64, ip, ip, dp, i32, i32, d, d, d, d, d, d, d, d, d, d, d,
                                          ctypes.POINTER(ctypes.c_int64), ip, dp, dp, dp, dp, dp, dp]
        L.mf_oracle_predict.restype = i32
        L.mf_oracle_predict.argtypes = [i64, i64, i64, i64, i32, i32, d, dp, dp, dp, dp, dp]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def draw_init(user_num, item_num, n_factors, init_mean=0.0, init_std_dev=0.1):
    """The two ``np.random.normal`` draws of ``fit`` in reference order (user table first)."""
    a = np.random.normal(init_mean, init_std_dev, size=(user_num, n_factors))
    b = np.random.normal(init_mean, init_std_dev, size=(item_num, n_factors))
    return a, b


def svd_fit(users, items, ratings, pu, qi, n_epochs=20, biased=True, lr_all=.005, reg_all=.02,
            lr_bu=None, lr_bi=None, lr_pu=None, lr_qi=None,
            reg_bu=None, reg_bi=None, reg_pu=None, reg_qi=None):
    """SVD.fit (util/matrix_factorization.pyx:104-155).  Returns dict(pu, qi, bu, bi, global_mean, sse)."""
    users = np.ascontiguousarray(users, dtype=np.int32)
    items = np.ascontiguousarray(items, dtype=np.int32)
    ratings = np.ascontiguousarray(ratings, dtype=np.float64)
    pu = np.array(pu, dtype=np.float64, order="C")
    qi = np.array(qi, dtype=np.float64, order="C")
    bu = np.zeros(pu.shape[0])
    bi = np.zeros(qi.shape[0])
    pick = lambda v, d: d if v is None else v
    mu = float(ratings.mean()) if biased else 0.0
    sse = lib().mf_oracle_svd_fit(len(ratings), _ip(users), _ip(items), _dp(ratings), pu.shape[1], n_epochs,
                                  int(bool(biased)),
                                  pick(lr_bu, lr_all), pick(lr_bi, lr_all), pick(lr_pu, lr_all), pick(lr_qi, lr_all),
                                  pick(reg_bu, reg_all), pick(reg_bi, reg_all), pick(reg_pu, reg_all),
                                  pick(reg_qi, reg_all), mu, _dp(pu), _dp(qi), _dp(bu), _dp(bi))
    return dict(pu=pu, qi=qi, bu=bu, bi=bi, global_mean=mu, sse=sse)


def rsvd_fit(users, items, ratings, ui, vj, n_epochs=20, version=2, lr=.001, reg=.02, reg2=.05):
    """RSVD.fit (util/matrix_factorization.pyx:22-66).  Returns dict(ui, vj, ci, dj, global_mean, sse)."""
    users = np.ascontiguousarray(users, dtype=np.int32)
    items = np.ascontiguousarray(items, dtype=np.int32)
    ratings = np.ascontiguousarray(ratings, dtype=np.float64)
    ui = np.array(ui, dtype=np.float64, order="C")
    vj = np.array(vj, dtype=np.float64, order="C")
    ci = np.zeros(ui.shape[0])
    dj = np.zeros(vj.shape[0])
    mu = float(ratings.mean())
    sse = lib().mf_oracle_rsvd_fit(len(ratings), _ip(users), _ip(items), _dp(ratings), ui.shape[1], n_epochs,
                                   int(version), lr, reg, reg2, mu, _dp(ui), _dp(vj), _dp(ci), _dp(dj))
    return dict(ui=ui, vj=vj, ci=ci, dj=dj, global_mean=mu, sse=sse)


def predict(u, i, pu, qi, bu, bi, with_bias, mu=0.0):
    """SVD.predict / RSVD.predict (util/matrix_factorization.pyx:157-167, 68-78)."""
    est = ctypes.c_double()
    rc = lib().mf_oracle_predict(int(u), int(i), pu.shape[0], qi.shape[0], pu.shape[1], int(bool(with_bias)),
                                 float(mu), _dp(pu), _dp(qi), _dp(bu), _dp(bi), ctypes.byref(est))
    if rc == -1:
        raise ValueError('Invalid user code')
    if rc == -2:
        raise ValueError('Invalid item code')
    return est.value



def user_item_lists(users, items, user_num):
    """``ur`` of SVDpp.fit (util/matrix_factorization.pyx:231-234) in CSR form: per user the items of its ratings in
    frame order (a repeated (user, item) rating appears twice, as in the reference's list)."""
    users = np.asarray(users, dtype=np.int64)
    order = np.argsort(users, kind="stable")
    ptr = np.zeros(user_num + 1, dtype=np.int64)
    np.add.at(ptr, users + 1, 1)
    return np.cumsum(ptr), np.ascontiguousarray(np.asarray(items)[order], dtype=np.int32)


def svdpp_fit(users, items, ratings, pu, qi, yj, n_epochs=20, lr_all=.007, reg_all=.02, lists=None, global_mean=None):
    """SVDpp.fit (util/matrix_factorization.pyx:193-271).  Returns dict(pu, qi, yj, bu, bi, global_mean, sse).
    ``lists`` = (ptr, idx) and ``global_mean`` of a larger frame let a bounded PREFIX of that frame's ratings be walked with
    the full frame's histories (bench.py's cpu_baseline sample); by default both come from the given ratings."""
    users = np.ascontiguousarray(users, dtype=np.int32)
    items = np.ascontiguousarray(items, dtype=np.int32)
    ratings = np.ascontiguousarray(ratings, dtype=np.float64)
    pu, qi, yj = (np.array(a, dtype=np.float64, order="C") for a in (pu, qi, yj))
    bu, bi = np.zeros(pu.shape[0]), np.zeros(qi.shape[0])
    ptr, idx = user_item_lists(users, items, pu.shape[0]) if lists is None else lists
    mu = float(ratings.mean()) if global_mean is None else float(global_mean)
    scratch = np.zeros(pu.shape[1])
    sse = lib().mf_oracle_svdpp_fit(len(ratings), _ip(users), _ip(items), _dp(ratings), pu.shape[1], n_epochs,
                                    lr_all, lr_all, lr_all, lr_all, lr_all, reg_all, reg_all, reg_all, reg_all, reg_all, mu,
                                    ptr.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _ip(idx), _dp(pu), _dp(qi), _dp(yj),
                                    _dp(bu), _dp(bi), _dp(scratch))
    return dict(pu=pu, qi=qi, yj=yj, bu=bu, bi=bi, global_mean=mu, sse=sse, ur_ptr=ptr, ur_idx=idx)


def svdpp_predict(u, i, fit):
    """SVDpp.predict (util/matrix_factorization.pyx:273-288)."""
    pu, qi, yj = fit["pu"], fit["qi"], fit["yj"]
    if u >= pu.shape[0]:
        raise ValueError('Invalid user code')
    if i >= qi.shape[0]:
        raise ValueError('Invalid item code')
    Iu = fit["ur_idx"][fit["ur_ptr"][u]:fit["ur_ptr"][u + 1]]
    impl = 0 if len(Iu) == 0 else sum(yj[j] for j in Iu) / np.sqrt(len(Iu))
    return fit["global_mean"] + fit["bu"][u] + fit["bi"][i] + np.dot(qi[i], pu[u] + impl)
