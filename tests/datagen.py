"""Seeded synthetic descriptor generators shared by tests and bench."""
import numpy as np

QUERY_SEED, DB_SEED = 1234, 5678


def uniform(n, seed):
    return np.random.default_rng(seed).integers(0, 256, (n, 32), np.uint8)


def low_entropy(n, seed):
    """bytes in {0,1}: ~44 % of queries have tied best matches -> stresses the lowest-index tie-break."""
    return np.random.default_rng(seed).integers(0, 2, (n, 32), np.uint8)


def flip_bits(rows, p, rng):
    """Bernoulli(p) bit flips of every row."""
    bits = np.unpackbits(rows, axis=1)
    bits ^= (rng.random(bits.shape) < p).astype(np.uint8)
    return np.packbits(bits, axis=1)


def noisy_copy(query, seed, p=0.05, replace=0.3, n=None):
    """A keyframe that 'sees the same place': per-row bit flips, a fraction of rows replaced by random,
    rows shuffled; n rows (default len(query))."""
    rng = np.random.default_rng(seed)
    n = len(query) if n is None else n
    src = rng.integers(0, len(query), n) if n != len(query) else rng.permutation(len(query))
    kf = flip_bits(query[src], p, rng)
    rep = rng.random(n) < replace
    kf[rep] = rng.integers(0, 256, (int(rep.sum()), 32), np.uint8)
    return kf


def keyframe(k, n=2000, seed=DB_SEED):
    """Keyframe k of the synthetic DB: bytes depend only on (seed, k), never on the shard layout."""
    return np.random.default_rng([seed, k]).integers(0, 256, (n, 32), np.uint8)


def structured_db(query, K, n=2000, seed=DB_SEED, every=10, ragged=False):
    """K keyframes; every `every`-th is a noisy copy of the query so filter/similarity are non-trivial."""
    rng = np.random.default_rng([seed, 99])
    out = []
    for k in range(K):
        nk = int(rng.integers(max(1, n - 200), n + 100)) if ragged else n
        if every and k % every == every - 1:
            out.append(noisy_copy(query, seed * 1000 + k, p=(0.02, 0.1)[(k // every) % 2], n=nk))
        else:
            out.append(keyframe(k, nk, seed))
    return out


def planted_pair(nq, nt, seed, kmax=40):
    """train = random rows + for some queries a copy with k in [0,kmax] flipped bits (so 2*min filters)."""
    rng = np.random.default_rng(seed)
    q = rng.integers(0, 256, (nq, 32), np.uint8)
    t = rng.integers(0, 256, (nt, 32), np.uint8)
    m = min(nq, nt) // 2
    if m:
        qi = rng.choice(nq, m, replace=False)
        ti = rng.choice(nt, m, replace=False)
        for a, b in zip(qi, ti):
            bits = np.unpackbits(q[a])
            flip = rng.choice(256, int(rng.integers(0, kmax + 1)), replace=False)
            bits[flip] ^= 1
            t[b] = np.packbits(bits)
    return q, t
