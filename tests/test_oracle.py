"""CPU tier: the plain-C restatement (oracle/lc_oracle.c) pinned against the real OpenCV matcher
(cv2.BFMatcher(NORM_HAMMING), the call behind the reference's matcher_, include/loop_closing.hpp:73)
and against the committed golden fixtures (tests/golden/hamming_golden.npz, made by make_golden.py)."""
import os

import numpy as np
import pytest

import datagen
from oracle import c_oracle, cv2_ref
from oracle.params import Params

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "hamming_golden.npz"))
PAIRS = sorted({k.split("/")[1] for k in GOLD.files if k.startswith("pair/")})
KNOBS = {"default": Params(), "floor30": Params(filter_floor=30), "strict": Params(filter_strict=True),
         "init100": Params(min_init=100)}
needs_cv2 = pytest.mark.skipif(not cv2_ref.available(), reason="cv2 not importable")


@pytest.mark.parametrize("name", PAIRS)
def test_c_oracle_matches_golden_pairs(name):
    q, t = GOLD[f"pair/{name}/q"], GOLD[f"pair/{name}/t"]
    idx, dist = c_oracle.match(q, t, nthreads=2)
    np.testing.assert_array_equal(idx, GOLD[f"pair/{name}/train_idx"])
    np.testing.assert_array_equal(dist, GOLD[f"pair/{name}/dist"])
    for kn, p in KNOBS.items():
        good, mn, ng = c_oracle.filter_matches(dist, p)
        np.testing.assert_array_equal(good, GOLD[f"pair/{name}/{kn}/good"])
        assert [mn, ng] == GOLD[f"pair/{name}/{kn}/summary"].tolist()


def _orb_case():
    n = GOLD["loop/orb/n"]
    rows = GOLD["loop/orb/rows"]
    offs = np.concatenate([[0], np.cumsum(n)])
    descs = [rows[offs[i]:offs[i + 1]] for i in range(len(n))]
    thr, gap, mg, floor = GOLD["loop/orb/params"]
    return descs, Params(loop_threshold=float(thr), min_loop_gap=int(gap), min_good=int(mg), filter_floor=int(floor))


def test_c_oracle_matches_golden_orb_loop_sequence():
    """real ORB descriptors of a synthetic closed-loop sequence, processFrame order (append i, detectLoops(i))."""
    descs, p = _orb_case()
    all_good, all_min, cands = [], [], []
    for i, d in enumerate(descs):
        res, best = c_oracle.loop_check(d, np.concatenate(descs[:i + 1]), [len(x) for x in descs[:i + 1]],
                                        list(range(i + 1)), i, p, nthreads=4)
        all_good.append(res["n_good"])
        all_min.append(res["min_dist"])
        cands += c_oracle.loop_candidates(res, len(d), i, p)
    np.testing.assert_array_equal(res, GOLD["loop/orb/last_results"])
    np.testing.assert_array_equal(best, GOLD["loop/orb/last_best"])
    np.testing.assert_array_equal(np.concatenate(all_good), GOLD["loop/orb/all_n_good"])
    np.testing.assert_array_equal(np.concatenate(all_min), GOLD["loop/orb/all_min_dist"])
    want = GOLD["loop/orb/candidates"]
    assert len(cands) == len(want) > 0
    assert np.array_equal(np.array(cands, np.float64).reshape(-1, 4), want)  # similarity doubles bit-identical


def test_c_oracle_matches_golden_multi_image_tiebreak():
    """Appendix A step 6: packed (dist<<32 | global_row) min == OpenCV's add([...]).match tie-break."""
    q, n, rows = GOLD["multi/q"], GOLD["multi/n"], GOLD["multi/rows"]
    _, best = c_oracle.loop_check(q, rows, n, np.zeros(len(n), np.int32), 100, Params(), nthreads=3)
    np.testing.assert_array_equal(best, GOLD["multi/best"])


@needs_cv2
@pytest.mark.parametrize("gen,nq,nt,seed", [
    ("uniform", 2000, 2000, 1), ("low_entropy", 300, 700, 2), ("low_entropy", 1, 1, 3), ("low_entropy", 33, 31, 4),
    ("uniform", 2081, 1999, 5), ("low_entropy", 700, 2081, 6)])
def test_c_oracle_vs_cv2_live(gen, nq, nt, seed):
    g = getattr(datagen, gen)
    q, t = g(nq, seed), g(nt, seed + 100)
    a, b = c_oracle.match(q, t, nthreads=4), cv2_ref.match(q, t)
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])
    for p in KNOBS.values():
        ga, gb = c_oracle.filter_matches(a[1], p), cv2_ref.filter_matches(b[1], p)
        np.testing.assert_array_equal(ga[0], gb[0])
        assert ga[1:] == gb[1:]


@needs_cv2
def test_c_oracle_vs_cv2_live_planted_and_loop():
    q, t = datagen.planted_pair(600, 800, 9)
    a, b = c_oracle.match_features(q, t), cv2_ref.match_features(q, t)
    for k in ("train_idx", "dist", "good"):
        np.testing.assert_array_equal(a[k], b[k])
    assert (a["min_dist"], a["n_good"]) == (b["min_dist"], b["n_good"]) and 0 < a["n_good"] < 600
    kfs = datagen.structured_db(q, 25, n=300, every=4, ragged=True) + [np.zeros((0, 32), np.uint8)]
    fids = list(range(0, 52, 2))
    for p in (Params(min_loop_gap=10, loop_threshold=0.15), Params(min_loop_gap=10, gap_strict=True, filter_floor=40)):
        r1, b1 = c_oracle.loop_check(q, np.concatenate(kfs), [len(k) for k in kfs], fids, 45, p, nthreads=4)
        r2, b2 = cv2_ref.loop_check(q, kfs, fids, 45, p)
        np.testing.assert_array_equal(r1, r2)
        np.testing.assert_array_equal(b1, b2)


def test_empty_sides_produce_no_matches():
    q, e = datagen.uniform(4, 1), np.zeros((0, 32), np.uint8)
    for a, b in ((q, e), (e, q), (e, e)):
        idx, dist = c_oracle.match(a, b)
        assert len(idx) == 0 and len(dist) == 0
        assert c_oracle.filter_matches(dist)[1:] == (-1, 0)


def test_oracle_thread_count_does_not_change_results():
    q = datagen.low_entropy(500, 1)
    kfs = [datagen.low_entropy(60 + i, 10 + i) for i in range(30)]
    args = (q, np.concatenate(kfs), [len(k) for k in kfs], list(range(30)), 100, Params())
    r1, b1 = c_oracle.loop_check(*args, nthreads=1)
    r8, b8 = c_oracle.loop_check(*args, nthreads=8)
    np.testing.assert_array_equal(r1, r8)
    np.testing.assert_array_equal(b1, b8)
