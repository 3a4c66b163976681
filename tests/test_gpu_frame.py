"""GPU tier: lc_process_frame (and its _begin / _matches / _end form) - processFrame's matcher side
(include/loop_closing.hpp:34, README.md:94-102) in one GPU submission - against the three separate calls it replaces
and against the oracles, bit-exact."""
import numpy as np
import pytest

import datagen
from oracle import c_oracle, cv2_ref
from oracle.params import Params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lc():
    import slam_loop_closing_b200 as m
    return m


def _ctx(lc, p, **kw):
    return lc.Context(loop_threshold=p.loop_threshold, min_loop_gap=p.min_loop_gap, min_good=p.min_good,
                      filter_floor=p.filter_floor, filter_strict=p.filter_strict, min_init=p.min_init,
                      gap_strict=p.gap_strict, thr_inclusive=p.thr_inclusive, **kw)


def _sequence(n_frames, seed, lo=900, hi=2100, revisit_from=None, empty_at=()):
    """descriptor frames of a walk that returns to its first places (noisy copies), ragged sizes, some empty frames"""
    rng = np.random.default_rng(seed)
    frames = []
    for i in range(n_frames):
        n = 0 if i in empty_at else int(rng.integers(lo, hi))
        if revisit_from is not None and i >= revisit_from and len(frames[i - revisit_from]) and n:
            frames.append(datagen.noisy_copy(frames[i - revisit_from], seed * 100 + i, p=0.04, replace=0.3, n=n))
        elif i and len(frames[-1]) and n and i % 3:
            frames.append(datagen.noisy_copy(frames[-1], seed * 100 + i, p=0.08, replace=0.5, n=n))  # consecutive frames overlap
        else:
            frames.append(datagen.uniform(n, seed * 100 + i))
    return frames


def _cands(a):
    return [(int(x["current_frame_id"]), int(x["matched_frame_id"]), int(x["num_matches"]), float(x["similarity_score"])) for x in a]


def _assert_pair_equal(got, want):
    for k in ("train_idx", "dist", "good"):
        np.testing.assert_array_equal(got[k], want[k])
    assert (got["min_dist"], got["n_good"]) == (want["min_dist"], want["n_good"])


@pytest.mark.parametrize("kernel", [0, 2, 4])
@pytest.mark.parametrize("p", [Params(loop_threshold=0.15, min_loop_gap=5), Params(loop_threshold=0.1, min_loop_gap=0, filter_floor=30),
                               Params(loop_threshold=0.15, min_loop_gap=4, gap_strict=True, filter_strict=True, thr_inclusive=True)])
def test_process_frame_equals_the_three_separate_calls_and_the_oracle(lc, kernel, p):
    """every frame of a 36-frame walk: fused call == match_pair(prev, cur) + db_append + loop_check + select_candidates on a
    second context, and == the C oracle (all frames) / OpenCV (sampled frames).  The store grows across the switch from
    the integer kernels to the tensor path, frames are ragged and two are empty."""
    frames = _sequence(36, 7, revisit_from=20, empty_at=(4, 27))
    with _ctx(lc, p, kernel=kernel) as a, _ctx(lc, p, kernel=kernel) as b:
        n_loops = 0
        for i, d in enumerate(frames):
            fid = 10 + 2 * i  # frame ids are not store indices
            pair, cands, k = a.process_frame(fid, d)
            assert k == i
            # the three separate calls
            if i:
                _assert_pair_equal(pair, b.match_pair(frames[i - 1], d))
            else:
                assert len(pair["train_idx"]) == 0 and (pair["min_dist"], pair["n_good"]) == (-1, 0)
            b.db_append(fid, d)
            res, _ = b.loop_check(d, fid, want_best=False)
            want = _cands(b.select_candidates(res, len(d), fid))
            assert _cands(cands) == want
            n_loops += len(want)
            # the oracles
            if i:
                _assert_pair_equal(pair, c_oracle.match_features(frames[i - 1], d, p, nthreads=8))
            rows = np.concatenate(frames[:i + 1]) if i else frames[0]
            w_res, _ = c_oracle.loop_check(d, rows, [len(f) for f in frames[:i + 1]], [10 + 2 * j for j in range(i + 1)], fid, p, nthreads=8)
            assert want == [(c, m, n, s) for (c, m, n, s) in c_oracle.loop_candidates(w_res, len(d), fid, p)]
            if i in (1, 21, 35) and cv2_ref.available():
                _assert_pair_equal(pair, cv2_ref.match_features(frames[i - 1], d, p))
        assert n_loops > 0
        assert a.db_size() == b.db_size()


def test_process_frame_in_three_steps_overlaps_host_work(lc):
    """_begin returns with everything queued; _matches hands out the pair match while the loop check is still running;
    _end the candidates and (on request) the per-keyframe records - all equal to the one-call form."""
    p = Params(loop_threshold=0.15, min_loop_gap=3)
    frames = _sequence(30, 11, lo=1800, hi=2050, revisit_from=15)
    with _ctx(lc, p) as a, _ctx(lc, p) as b:
        for i, d in enumerate(frames):
            assert a.process_frame_begin(i, d) == i
            d_copy = d.copy()
            d[:] = 0  # the library has copied the rows
            busy = sum(int(x) for x in np.arange(3000))  # "estimatePose"
            pair = a.process_frame_matches()
            pair_again = a.process_frame_matches()  # may be asked twice
            cands, rec = a.process_frame_end(want_records=True)
            d[:] = d_copy
            w_pair, w_cands, _ = b.process_frame(i, d)
            _assert_pair_equal(pair, w_pair)
            _assert_pair_equal(pair_again, w_pair)
            assert _cands(cands) == _cands(w_cands) and busy > 0
            res, _ = b.loop_check(d, i, want_best=False)
            np.testing.assert_array_equal(rec, res)
        assert a.last_timing().path == 4  # 30 keyframes of ~1900 rows: the FP4 tensor path
        # skipping _matches is allowed
        a.process_frame_begin(99, frames[0])
        assert len(a.process_frame_end()) >= 0


def test_process_frame_errors_and_state(lc):
    p = Params(loop_threshold=0.15, min_loop_gap=1)
    f0, f1 = datagen.uniform(300, 1), datagen.uniform(280, 2)
    with _ctx(lc, p) as c:
        with pytest.raises(lc.LcError):
            c.process_frame_end()  # nothing pending
        with pytest.raises(lc.LcError):
            c.process_frame_matches()
        c.process_frame(0, f0)
        # cap smaller than the previous frame's rows: refused before anything is appended
        import ctypes as C
        k, nm, mn, ng, n_out = C.c_int64(-5), C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
        idx = np.zeros(10, np.int32)
        out = np.zeros(4, lc.LOOP_CANDIDATE)
        rc = c._lib.lc_process_frame(c._h, 1, f1.ctypes.data, len(f1), 32, idx.ctypes.data, None, None, 10, C.byref(nm), C.byref(mn), C.byref(ng),
                                     out.ctypes.data, 4, C.byref(n_out), C.byref(k))
        assert rc == 4 and k.value == -1 and c.db_size()["keyframes"] == 1  # LC_ERR_CAPACITY, store untouched
        # a frame in flight: no second frame, no loop check; other calls are fine
        c.process_frame_begin(1, f1)
        with pytest.raises(lc.LcError):
            c.process_frame_begin(2, f0)
        with pytest.raises(lc.LcError):
            c.loop_check(f0, 5)
        assert c._lib.lc_loop_check_end(c._h, None, 0, None, None) == 1  # LC_ERR_INVALID: the frame owns the pending check
        assert c.db_size()["keyframes"] == 2
        got = c.match_pair(f0, f1)  # allowed between _begin and _end
        pair = c.process_frame_matches()
        _assert_pair_equal(pair, got)
        c.process_frame_end()
        # pending loop check blocks a frame too
        c.loop_check_begin(f0, 9)
        with pytest.raises(lc.LcError):
            c.process_frame(2, f0)
        c.loop_check_end()
        assert c.db_size()["keyframes"] == 2
    with _ctx(lc, p, shard_rank=0, shard_world=2) as c:
        with pytest.raises(lc.LcError):
            c.process_frame(0, f0)  # unsharded contexts only
        assert c.db_size()["keyframes"] == 0


def test_process_frame_strided_rows_and_table_growth(lc):
    """a strided cv::Mat-like view is accepted, and the keyframe table grows past its first allocation (1024 entries)
    while frames keep coming through the kernel-argument road."""
    p = Params(loop_threshold=0.15, min_loop_gap=2)
    wide = np.random.default_rng(3).integers(0, 256, (400, 48), np.uint8)
    with _ctx(lc, p) as a, _ctx(lc, p) as b:
        for i in range(1100):
            d = wide[(i * 7) % 350:(i * 7) % 350 + 40, 8:40]  # 40 rows, row stride 48
            pair, cands, k = a.process_frame(i, d)
            if i % 97 == 0 or i > 1090:
                if i:
                    b.db_append_many(list(range(b.db_size()["keyframes"], i)),
                                     [wide[(j * 7) % 350:(j * 7) % 350 + 40, 8:40] for j in range(b.db_size()["keyframes"], i)])
                    prev = wide[((i - 1) * 7) % 350:((i - 1) * 7) % 350 + 40, 8:40]
                    _assert_pair_equal(pair, b.match_pair(prev, d))
                b.db_append(i, d)
                res, _ = b.loop_check(d, i, want_best=False)
                assert _cands(cands) == _cands(b.select_candidates(res, len(d), i))
        assert a.db_size()["keyframes"] == 1100
