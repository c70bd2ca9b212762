"""Scratch diagnostic (not a test): device funk-SVD vs golden / C oracle in a fresh process."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pandas as pd
from oracle import mf_oracle
from recommend_lib_b200.mf import SVD, RSVD

e = lambda x, y: float(np.abs(x - y).max() / np.abs(y).max())
g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mf_small.npz"))
df = pd.DataFrame({"user": g["users"].astype(np.int64), "item": g["items"].astype(np.int64), "rating": g["ratings"]})
for rep in range(2):
    np.random.seed(2019)
    a = SVD(int(g["U"]), int(g["I"]), n_factors=int(g["D"]), n_epochs=int(g["E"]), verbose=False, biased=True)
    a.fit(df)
    print("golden svd_b: pu", e(a.pu, g["svd_b_pu"]), "qi", e(a.qi, g["svd_b_qi"]), "bu", e(a.bu, g["svd_b_bu"]), "mu",
          a.global_mean - float(g["svd_b_mu"]), flush=True)
    o = mf_oracle.svd_fit(g["users"], g["items"], g["ratings"], g["svd_b_pu0"], g["svd_b_qi0"], n_epochs=int(g["E"]))
    print("   oracle vs golden", e(o["pu"], g["svd_b_pu"]), " device vs oracle", e(a.pu, o["pu"]), flush=True)
    np.random.seed(2019)
    p0 = np.random.normal(0, .1, size=(int(g["U"]), int(g["D"])))
    print("   init equal:", np.array_equal(p0, g["svd_b_pu0"]))

rng = np.random.default_rng(0)
U, I, D, N = 500, 7, 16, 20000
users = rng.integers(0, U, N).astype(np.int32)
items = np.full(N, 3, dtype=np.int32)
ratings = rng.integers(1, 6, N).astype(np.float64)
df = pd.DataFrame({"user": users, "item": items, "rating": ratings})
for biased in (True, False):
    np.random.seed(9)
    a = SVD(U, I, n_factors=D, n_epochs=2, verbose=False, biased=biased)
    a.fit(df)
    np.random.seed(9)
    p0, q0 = mf_oracle.draw_init(U, I, D)
    o = mf_oracle.svd_fit(users, items, ratings, p0, q0, n_epochs=2, biased=biased)
    print(f"chain biased={biased}: pu {e(a.pu, o['pu']):.2e} qi {e(a.qi, o['qi']):.2e}", flush=True)
