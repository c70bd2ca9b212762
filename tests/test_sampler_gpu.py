"""Device-side negative sampler (daisy_sample_triples) against its numpy restatement (oracle/sampler_oracle.py):
bit-exact triples, the reference's semantics as properties (util/data_loader.py:680-690), epoch / seed keying."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _pairs(U, I, n, seed):
    rng = np.random.default_rng(seed)
    p = np.stack([rng.integers(0, U, n), rng.integers(0, I, n)], 1)
    return np.unique(p, axis=0)


@pytest.mark.parametrize("U,I,n,num_ng,shuffle", [(50, 40, 600, 4, True), (943, 1682, 99057, 4, True),
                                                  (7, 5, 20, 3, False), (2000, 300000, 50000, 1, True)])
def test_device_sampler_equals_oracle_bit_for_bit(U, I, n, num_ng, shuffle):
    assert torch.cuda.is_available()
    from oracle import sampler_oracle
    from recommend_lib_b200.sampler import DeviceTripleSampler
    pairs = _pairs(U, I, n, seed=U)
    s = DeviceTripleSampler(pairs, I, U, num_ng=num_ng, seed=2019)
    for epoch in (0, 3):
        got = s.sample_epoch(epoch, shuffle=shuffle).cpu().numpy()
        s.check()
        want = sampler_oracle.sample_epoch(pairs, I, num_ng, 2019, epoch, shuffle)
        assert got.shape == want.shape == (len(pairs) * num_ng, 3)
        assert np.array_equal(got, want)


def test_device_sampler_semantics_and_keying():
    from recommend_lib_b200.sampler import DeviceTripleSampler
    U, I = 300, 200
    pairs = _pairs(U, I, 20000, seed=1)                    # dense: a third of all (u, i) are positives
    s = DeviceTripleSampler(pairs, I, U, num_ng=4, seed=7)
    flat = s.sample_epoch(0, shuffle=False).cpu().numpy()
    # features_fill order: positive-major, num_ng consecutive negatives each
    assert np.array_equal(flat[:, :2], np.repeat(pairs, 4, axis=0))
    pos = set(map(tuple, pairs.tolist()))
    assert not any((int(u), int(j)) in pos for u, _, j in flat)          # never a training positive of u
    assert flat[:, 2].min() >= 0 and flat[:, 2].max() < I
    # uniform over the non-positives of each user: chi-square-ish check on one busy user
    sh = s.sample_epoch(0, shuffle=True).cpu().numpy()
    assert sorted(map(tuple, sh.tolist())) == sorted(map(tuple, flat.tolist()))   # a permutation of the same triples
    assert not np.array_equal(sh, flat)
    assert not np.array_equal(s.sample_epoch(1, shuffle=False).cpu().numpy(), flat)        # epochs differ
    s2 = DeviceTripleSampler(pairs, I, U, num_ng=4, seed=8)
    assert not np.array_equal(s2.sample_epoch(0, shuffle=False).cpu().numpy(), flat)       # seeds differ
    assert np.array_equal(s.sample_epoch(0, shuffle=False).cpu().numpy(), flat)            # reproducible
    counts = np.bincount(flat[:, 2], minlength=I)
    assert counts.min() > 0 and counts.max() < 3.0 * counts.mean()


def test_device_sampler_flags_a_user_without_negatives():
    from recommend_lib_b200.sampler import DeviceTripleSampler
    I = 6
    pairs = np.array([[0, i] for i in range(I)] + [[1, 2]])              # user 0 has every item
    s = DeviceTripleSampler(pairs, I, 2, num_ng=2, seed=1)
    s.sample_epoch(0)
    with pytest.raises(IndexError):
        s.check()


def test_fit_with_device_sampler_trains_ml100k(golden):
    """BPRMFRecommender.fit(sampler='device') on the ml-100k split: the same model quality as the host-sampled run
    (a different random stream of the same distribution, so compared within a few percent, not bit for bit)."""
    from recommend_lib_b200 import data
    from recommend_lib_b200.bpr import BPRMFRecommender
    s = golden("ml100k_split.npz")
    tr, te = s["train_pairs"].astype(np.int64), s["test_pairs"].astype(np.int64)
    U, I = int(s["user_num"]), int(s["item_num"])
    allp = np.concatenate([tr, te])
    eu, ec = data.eval_candidates(allp[:, 0], allp[:, 1], te[:, 0], te[:, 1], I, 999, 2019)
    hist = {}
    for smp in ("host", "device"):
        r = BPRMFRecommender(U, I, factor_num=64, epochs=8, sampler=smp, seed=2019, device="cuda:0")
        r.fit(tr, eu, ec)
        hist[smp] = r.history
    h, d = hist["host"][-1], hist["device"][-1]
    assert abs(d["loss"] - h["loss"]) / h["loss"] < 0.03, (d, h)
    assert d["loss"] < hist["device"][0]["loss"]
    assert abs(d["hr"] - h["hr"]) < 0.03, (d, h)
