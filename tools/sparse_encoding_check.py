"""Exhaustive check of the 6-bits-per-chunk encoding for a pair-granular 4:8 sparse FP4 MMA.

A chunk is 8 operand positions = 4 pairs.  Train side (dense operand), signs s in {+-1}^6:
    b = (s1+s2, s5 | s1-s2, s5 | s3+s4, s6 | s3-s4, s6)                 values 0, +-1, +-2 (all E2M1)
Query side (sparse operand), signs sigma: pair 1 if sigma1 == sigma2 else pair 2 holds (sigma1, sigma5); pair 3 if
sigma3 == sigma4 else pair 4 holds (sigma3, sigma6); the other two pairs are zero.
Claim: a . b == sigma . s == 6 - 2 * hamming for all 64 x 64 combinations, with exactly two non-zero pairs in every a.
Also the element-granular variant: 4 bits in 4 positions, a = (s1+s2, s1-s2, s3+s4, s3-s4) / 2, b likewise without the / 2.
"""
import itertools
import numpy as np

E2M1 = {0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0}


def dense6(s):
    s1, s2, s3, s4, s5, s6 = s
    return np.array([s1 + s2, s5, s1 - s2, s5, s3 + s4, s6, s3 - s4, s6], float)


def sparse6(g):
    g1, g2, g3, g4, g5, g6 = g
    a = np.zeros(8)
    p = 0 if g1 == g2 else 1
    a[2 * p:2 * p + 2] = (g1, g5)
    p = 2 if g3 == g4 else 3
    a[2 * p:2 * p + 2] = (g3, g6)
    return a


def main():
    signs6 = list(itertools.product((-1, 1), repeat=6))
    for g in signs6:
        a = sparse6(g)
        active = [bool(np.any(a[2 * p:2 * p + 2])) for p in range(4)]
        assert sum(active) == 2, (g, a)
        assert all(abs(v) in E2M1 for v in a)
        for s in signs6:
            b = dense6(s)
            assert all(abs(v) in E2M1 for v in b)
            ham = sum(x != y for x, y in zip(g, s))
            assert a @ b == np.dot(g, s) == 6 - 2 * ham, (g, s)
    signs4 = list(itertools.product((-1, 1), repeat=4))
    for g in signs4:
        a = np.array([g[0] + g[1], g[0] - g[1], g[2] + g[3], g[2] - g[3]]) / 2
        assert np.count_nonzero(a) == 2
        for s in signs4:
            b = np.array([s[0] + s[1], s[0] - s[1], s[2] + s[3], s[2] - s[3]], float)
            assert a @ b == np.dot(g, s)
    print("ok: 64 x 64 six-bit chunks (pair-granular 4:8) and 16 x 16 four-bit chunks (2:4) reproduce sigma . s exactly")


if __name__ == "__main__":
    main()
