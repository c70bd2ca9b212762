"""Diagnostic: per-link cost of daisy_mf_fit's schedules on (a) one pure item chain, (b) two interleaved hot chains,
(c) the config-2 shape.  Run on a GPU box:  DAISY_MF_STATS=1 python tools/mf_chain_probe.py"""
import ctypes
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recommend_lib_b200 import _lib  # noqa: E402
from recommend_lib_b200._lib import MFParams, c_vp  # noqa: E402
from recommend_lib_b200.sampler import synthetic_ratings  # noqa: E402

dev = torch.device("cuda:0")


def run(name, users, items, ratings, U, I, D=128, epochs=(1, 5)):
    rng = np.random.default_rng(0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    dpu, dqi = t(rng.normal(0, 0.1, (U, D))), t(rng.normal(0, 0.1, (I, D)))
    dbu, dbi = t(np.zeros(U)), t(np.zeros(I))
    du, di, dr = t(users.astype(np.int32)), t(items.astype(np.int32)), t(ratings.astype(np.float64))
    prm = MFParams(0, 1, .005, .005, .005, .005, .02, .02, .02, .02, 0.0, float(ratings.mean()))
    h = _lib.Handle(0, U, I, D, 0)
    s = _lib.stream_ptr(torch, dev)
    times = []
    for ep in (epochs[0],) + tuple(epochs):
        torch.cuda.synchronize()
        t0 = time.time()
        _lib.check(h.L.daisy_mf_fit(h.ptr, c_vp(dpu.data_ptr()), c_vp(dqi.data_ptr()), c_vp(dbu.data_ptr()),
                                    c_vp(dbi.data_ptr()), c_vp(du.data_ptr()), c_vp(di.data_ptr()), c_vp(dr.data_ptr()),
                                    len(ratings), ep, ctypes.byref(prm), None, s))
        torch.cuda.synchronize()
        times.append(time.time() - t0)
    per_epoch = (times[2] - times[1]) / (epochs[1] - epochs[0])
    fixed = times[1] - per_epoch * epochs[0]
    hot = np.bincount(items).max()
    print(f"{name:36s} n={len(ratings):8d} D={D:3d} hottest chain {hot:7d}: {per_epoch * 1e3:8.2f} ms/epoch (+{fixed * 1e3:6.1f} ms fixed) "
          f"= {per_epoch / hot * 1e6:6.3f} us per hottest-chain link, {len(ratings) / per_epoch / 1e6:6.2f} M ratings/s", flush=True)
    h.close()


rng = np.random.default_rng(1)
N = 200_000
r = rng.integers(1, 6, N).astype(np.float64)
for U in (6040, 50_000, 400_000):
    run(f"one item, {U} users", rng.integers(0, U, N), np.zeros(N, np.int64), r, U, 4)
run("one item, 6040 users, D=32", rng.integers(0, 6040, N), np.zeros(N, np.int64), r, 6040, 4, D=32)
run("one item, users round-robin 6040", np.arange(N) % 6040, np.zeros(N, np.int64), r, 6040, 4)
run("two items alternating, 6040 users", rng.integers(0, 6040, N), np.arange(N) % 2, r, 6040, 4)
u, i, rr = synthetic_ratings(1_000_209, 6040, 3706, seed=2019)
run("config 2 (zipf 3706 items)", u, i, rr, 6040, 3706)
