"""CPU tier: the plain-C restatement of the shipped pipeline's SIFT / L2 matcher (oracle/l2_oracle.c; reference
src/main.cpp:509-534, 1375-1392) pinned against the committed cv2 fixtures (tests/golden/l2_golden.npz, made by
make_l2_golden.py from cv2.SIFT + cv2.BFMatcher(NORM_L2).knnMatch) and live against cv2."""
import os

import numpy as np
import pytest

from oracle import l2_oracle

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "l2_golden.npz"))
PAIRS = sorted({k.split("/")[1] for k in GOLD.files if k.startswith("pair/")})


def loop_case():
    n = GOLD["loop/n"]
    offs = np.concatenate([[0], np.cumsum(n)])
    descs = [GOLD["loop/rows"][offs[i]:offs[i + 1]] for i in range(len(n))]
    gap, min_rows, ratio = GOLD["loop/params"]
    return descs, int(gap), int(min_rows), float(ratio)


@pytest.mark.parametrize("name", PAIRS)
def test_l2_oracle_matches_golden_pairs(name):
    q, t = GOLD[f"pair/{name}/q"], GOLD[f"pair/{name}/t"]
    for ratio in (0.75, 0.7):
        r = l2_oracle.match_features(q, t, ratio, nthreads=2)
        np.testing.assert_array_equal(r["train_idx"], GOLD[f"pair/{name}/train_idx"])
        has1, has2 = r["dist2"] >= 0, r["second_dist2"] >= 0
        np.testing.assert_array_equal(r["distance"][has1], GOLD[f"pair/{name}/distance"][has1])  # float32, bit for bit
        np.testing.assert_array_equal(np.sqrt(r["second_dist2"][has2].astype(np.float32)), GOLD[f"pair/{name}/second"][has2])
        np.testing.assert_array_equal(r["good"], GOLD[f"pair/{name}/good_{ratio}"])


def test_l2_oracle_matches_golden_sift_loop_sequence():
    descs, gap, min_rows, ratio = loop_case()
    want = GOLD["loop/n_good"]
    for curr in range(len(descs)):
        res = l2_oracle.loop_check(descs[curr], descs[:curr + 1], list(range(curr + 1)), curr, gap, ratio, min_rows, nthreads=4)
        np.testing.assert_array_equal(res["n_good"], want[curr, :curr + 1])
    r = l2_oracle.match_features(descs[-1], descs[0], ratio)
    np.testing.assert_array_equal(r["train_idx"], GOLD["loop/last_vs_first/train_idx"])
    np.testing.assert_array_equal(r["distance"], GOLD["loop/last_vs_first/distance"])
    np.testing.assert_array_equal(r["good"], GOLD["loop/last_vs_first/good"])


@pytest.mark.skipif(not l2_oracle.cv2_available(), reason="cv2 not importable")
def test_l2_oracle_vs_cv2_live():
    rng = np.random.default_rng(3)
    for nq, nt in ((257, 300), (64, 1), (33, 500)):
        q = rng.integers(0, 256, (nq, 128)).astype(np.uint8)
        t = rng.integers(0, 256, (nt, 128)).astype(np.uint8)
        a, b = l2_oracle.match_features(q, t, 0.8), l2_oracle.cv2_match_features(q, t, 0.8)
        np.testing.assert_array_equal(a["train_idx"], b["train_idx"])
        np.testing.assert_array_equal(a["distance"], b["distance"])
        np.testing.assert_array_equal(a["good"], b["good"])


def test_to_u8_rejects_non_sift_values():
    with pytest.raises(ValueError):
        l2_oracle.to_u8(np.full((2, 128), 0.5, np.float32))
    with pytest.raises(ValueError):
        l2_oracle.to_u8(np.full((2, 128), 300.0, np.float32))
