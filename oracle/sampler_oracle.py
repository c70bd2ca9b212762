"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the device-side negative sampler's rule (csrc/sampler.cu).

The reference's sampler (BPRData.ng_sample, util/data_loader.py:680-690, plus DataLoader(shuffle=True),
BPRMFRecommender.py:141-142) draws from unseeded global RNG state, so there is no reference output to match; what is
pinned here is (i) the generator itself against the published Random123 known-answer vectors of Philox4x32-10
(tests/test_oracle_golden.py), (ii) the reference's semantics -- features_fill order, j uniform over [0, item_num),
never a training positive of u -- as properties, and (iii) bit equality between this restatement and the device.
"""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(key, ctr):
    """key: 2 uint32 (scalars or arrays), ctr: 4 uint32 arrays -> 4 uint32 arrays (Salmon et al., SC'11)."""
    k0, k1 = (np.asarray(k, dtype=np.uint64) & MASK for k in key)
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK for c in ctr)
    for r in range(10):
        if r > 0:
            k0 = (k0 + np.uint64(W0)) & MASK
            k1 = (k1 + np.uint64(W1)) & MASK
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
    return tuple(np.asarray(c, dtype=np.uint64).astype(np.uint32) for c in (c0, c1, c2, c3))


def sample_epoch(pairs, item_num, num_ng, seed, epoch, shuffle=True):
    """int32 [len(pairs) * num_ng, 3]: the rule in the header of csrc/sampler.cu."""
    pairs = np.asarray(pairs, dtype=np.int64)[:, :2]
    n = len(pairs) * num_ng
    slots = np.arange(n, dtype=np.uint64)
    s_lo, s_hi = slots & MASK, slots >> np.uint64(32)
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    u = np.repeat(pairs[:, 0], num_ng)
    i = np.repeat(pairs[:, 1], num_ng)
    keys = np.unique(pairs[:, 0] * item_num + pairs[:, 1])

    def is_pos(uu, jj):
        k = uu * item_num + jj
        pos = np.searchsorted(keys, k)
        pos[pos == len(keys)] = 0
        return keys[pos] == k

    def draw(idx, attempt):
        r = philox4x32_10((k0, k1), (s_lo[idx], s_hi[idx], np.full(len(idx), epoch, np.uint64),
                                     np.full(len(idx), attempt, np.uint64)))[0]
        return ((r.astype(np.uint64) * np.uint64(item_num)) >> np.uint64(32)).astype(np.int64)

    todo = np.arange(n)
    j = np.zeros(n, dtype=np.int64)
    attempt = 0
    while todo.size and attempt < 4096:
        j[todo] = draw(todo, attempt)
        todo = todo[is_pos(u[todo], j[todo])] if len(keys) else todo[:0]
        attempt += 1
    out = np.stack([u, i, j], 1).astype(np.int32)
    if shuffle:
        sk = philox4x32_10((k0, k1 ^ np.uint64(0x9E3779B9)), (s_lo, s_hi, np.full(n, epoch, np.uint64),
                                                               np.full(n, 0xFFFFFFFF, np.uint64)))[0]
        out = out[np.argsort(sk, kind="stable")]
    return np.ascontiguousarray(out)


def ng_sample_loop(pairs, item_num, num_ng, positives=None, rng=None):
    """The reference's own per-positive Python loop (BPRData.ng_sample, util/data_loader.py:680-690), restated with a
    set standing in for the dok_matrix ``train_mat`` and a local RandomState for ``np.random``: the CPU baseline the
    sampler bench times (it is what the reference runs once per epoch, single-threaded)."""
    rng = rng if rng is not None else np.random.RandomState(2019)
    train = positives if positives is not None else set(map(tuple, np.asarray(pairs)[:, :2].tolist()))
    fill = []
    for x in pairs:
        u, i = int(x[0]), int(x[1])
        for _ in range(num_ng):
            j = rng.randint(item_num)
            while (u, j) in train:
                j = rng.randint(item_num)
            fill.append([u, i, j])
    return fill
