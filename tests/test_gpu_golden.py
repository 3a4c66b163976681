"""GPU parity against the COMMITTED golden fixtures (tests/golden/hamming_golden.npz, written by make_golden.py from
cv2.BFMatcher(NORM_HAMMING) - the call behind the reference's matcher_, include/loop_closing.hpp:73): no oracle code
runs here, the C-ABI results are compared with the stored OpenCV answers bit for bit."""
import os

import numpy as np
import pytest

from oracle.params import Params

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "hamming_golden.npz"))
PAIRS = sorted({k.split("/")[1] for k in GOLD.files if k.startswith("pair/")})
KNOBS = {"default": Params(), "floor30": Params(filter_floor=30), "strict": Params(filter_strict=True),
         "init100": Params(min_init=100)}


@pytest.fixture(scope="module")
def lc():
    import slam_loop_closing_b200 as m
    return m


def _ctx(lc, p, **kw):
    return lc.Context(loop_threshold=p.loop_threshold, min_loop_gap=p.min_loop_gap, min_good=p.min_good,
                      filter_floor=p.filter_floor, filter_strict=p.filter_strict, min_init=p.min_init,
                      gap_strict=p.gap_strict, thr_inclusive=p.thr_inclusive, **kw)


@pytest.mark.parametrize("kernel", [1, 2])  # every pair match runs on the integer kernels (POPC, CSA)
@pytest.mark.parametrize("name", PAIRS)
def test_match_pair_equals_golden(lc, name, kernel):
    """matchFeatures (include/loop_closing.hpp:40): train index, distance, 2 x min filter under every knob set."""
    q, t = GOLD[f"pair/{name}/q"], GOLD[f"pair/{name}/t"]
    for kn, p in KNOBS.items():
        with _ctx(lc, p, kernel=kernel) as c:
            r = c.match_pair(q, t)
        if len(q) == 0 or len(t) == 0:
            assert len(r["train_idx"]) == 0 and [r["min_dist"], r["n_good"]] == GOLD[f"pair/{name}/{kn}/summary"].tolist() == [-1, 0]
            continue
        np.testing.assert_array_equal(r["train_idx"], GOLD[f"pair/{name}/train_idx"])
        np.testing.assert_array_equal(r["dist"], GOLD[f"pair/{name}/dist"])
        np.testing.assert_array_equal(r["good"].astype(bool), GOLD[f"pair/{name}/{kn}/good"].astype(bool))
        assert [r["min_dist"], r["n_good"]] == GOLD[f"pair/{name}/{kn}/summary"].tolist()


def _orb_case():
    n = GOLD["loop/orb/n"]
    rows = GOLD["loop/orb/rows"]
    offs = np.concatenate([[0], np.cumsum(n)])
    descs = [rows[offs[i]:offs[i + 1]] for i in range(len(n))]
    thr, gap, mg, floor = GOLD["loop/orb/params"]
    return descs, Params(loop_threshold=float(thr), min_loop_gap=int(gap), min_good=int(mg), filter_floor=int(floor))


@pytest.mark.parametrize("kernel", [0, 2, 3, 4])  # AUTO, CSA, int8 tensor, FP4 tensor (small stores fall back to CSA)
def test_orb_loop_sequence_equals_golden(lc, kernel):
    """BASELINE config 1 stand-in: real ORB descriptors of a synthetic closed-loop sequence in processFrame order
    (append frame i, then detectLoops(i)); per-keyframe records, candidates and similarity doubles as OpenCV gives."""
    descs, p = _orb_case()
    all_good, all_min, cands = [], [], []
    with _ctx(lc, p, kernel=kernel) as c:
        for i, d in enumerate(descs):
            c.db_append(i, d)
            res, best = c.loop_check(d, i)
            all_good.append(res["n_good"])
            all_min.append(res["min_dist"])
            cands += [(x["current_frame_id"], x["matched_frame_id"], x["num_matches"], x["similarity_score"])
                      for x in c.select_candidates(res, len(d), i)]
    np.testing.assert_array_equal(res, GOLD["loop/orb/last_results"])
    np.testing.assert_array_equal(best, GOLD["loop/orb/last_best"])
    np.testing.assert_array_equal(np.concatenate(all_good), GOLD["loop/orb/all_n_good"])
    np.testing.assert_array_equal(np.concatenate(all_min), GOLD["loop/orb/all_min_dist"])
    want = GOLD["loop/orb/candidates"]
    assert len(cands) == len(want) > 0
    assert np.array_equal(np.array([tuple(float(v) for v in x) for x in cands], np.float64).reshape(-1, 4), want)


@pytest.mark.parametrize("fused", [True, False])
def test_loop_closing_system_twin_equals_golden(lc, fused):
    """the same sequence through the LoopClosingSystem mirror (include/loop_closing.hpp:29-66): one GPU submission per
    frame (lc_process_frame_begin / _matches / _end) or the three separate calls."""
    descs, p = _orb_case()
    s = lc.LoopClosingSystem(p.loop_threshold, p.min_loop_gap, min_good=p.min_good, filter_floor=p.filter_floor, fused=fused)
    for i, d in enumerate(descs):
        loops = s.processDescriptors(d, i)
        if loops:  # the candidates' DMatch lists in one GPU call: sizes equal the candidates' match counts
            m = s.loopMatches(i, loops)
            assert [len(m[c.matched_frame_id]) for c in loops] == [c.num_matches for c in loops]
    got = np.array([(c.current_frame_id, c.matched_frame_id, c.num_matches, c.similarity_score) for c in s.getLoopClosures()],
                   np.float64).reshape(-1, 4)
    assert np.array_equal(got, GOLD["loop/orb/candidates"])


@pytest.mark.parametrize("kernel", [2, 3, 4])
def test_multi_image_tiebreak_equals_golden(lc, kernel):
    """Appendix A step 6: packed (dist << 32 | global_row) min == OpenCV's add([...]).match tie-break.  The store is
    replicated so that the tensor paths have a machine's worth of work; copies never win a tie (higher global rows)."""
    q, n, rows = GOLD["multi/q"], GOLD["multi/n"], GOLD["multi/rows"]
    offs = np.concatenate([[0], np.cumsum(n)])
    kfs = [rows[offs[i]:offs[i + 1]] for i in range(len(n))]
    reps = 1 if kernel == 2 else 80
    with _ctx(lc, Params(), kernel=kernel) as c:
        c.db_append_many([0] * (len(kfs) * reps), kfs * reps)
        _, best = c.loop_check(q, 100)
        if kernel != 2:
            assert c.last_timing().path == kernel
    np.testing.assert_array_equal(best, GOLD["multi/best"])
