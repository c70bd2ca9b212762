"""Golden fixture for SVD++: the reference's own compiled Cython class SVDpp
(util/matrix_factorization.pyx:169-288, built into oracle/_ref by oracle/build_ref.py) fitted on a tiny seeded frame.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_svdpp_golden.py      # build container only (needs /root/reference)
"""
import contextlib
import io
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.build_ref import load_ref  # noqa: E402


def main():
    m = load_ref()
    assert m is not None, "oracle/_ref could not be built"
    rng = np.random.default_rng(23)
    U, I, D, N, E = 25, 18, 6, 220, 3
    users, items = rng.integers(0, U, N), rng.integers(0, I, N)          # repeated (user, item) ratings included
    ratings = rng.integers(1, 6, N).astype(np.float64)
    df = pd.DataFrame({"user": users.astype(np.int64), "item": items.astype(np.int64), "rating": ratings})
    np.random.seed(2019)                                                   # fit draws pu, qi, yj from the global RNG (:216-219)
    a = m.SVDpp(U, I, n_factors=D, n_epochs=E, verbose=False)
    with contextlib.redirect_stdout(io.StringIO()):
        a.fit(df)
    np.random.seed(2019)
    pu0 = np.random.normal(0, .1, size=(U, D))
    qi0 = np.random.normal(0, .1, size=(I, D))
    yj0 = np.random.normal(0, .1, size=(I, D))
    pred = np.array([a.predict(int(u), int(i)) for u, i in zip(users[:15], items[:15])])
    np.savez_compressed(os.path.join(HERE, "svdpp_small.npz"), U=U, I=I, D=D, E=E, users=users, items=items, ratings=ratings,
                        pu0=pu0, qi0=qi0, yj0=yj0, pu=np.asarray(a.pu), qi=np.asarray(a.qi), yj=np.asarray(a.yj),
                        bu=np.asarray(a.bu), bi=np.asarray(a.bi), mu=a.global_mean, pred=pred)
    print("wrote svdpp_small.npz; |pu| max", np.abs(a.pu).max())


if __name__ == "__main__":
    main()
