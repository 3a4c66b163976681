"""GPU parity: liblc_b200 (through the C ABI) vs the CPU oracle, bit-exact (integer work).

Checkers: oracle/c_oracle (plain-C restatement) everywhere, and the real OpenCV matcher
(oracle/cv2_ref: cv2.BFMatcher(NORM_HAMMING)) at sizes it finishes in seconds.
"""
import numpy as np
import pytest

import datagen
from oracle import c_oracle, cv2_ref
from oracle.params import Params

pytestmark = pytest.mark.gpu

KERNELS = [1, 2]  # LC_KERNEL_POPC, LC_KERNEL_CSA


@pytest.fixture(scope="module")
def lc():
    import slam_loop_closing_b200 as m
    return m


def _ctx(lc, p: Params = Params(), **kw):
    return lc.Context(loop_threshold=p.loop_threshold, min_loop_gap=p.min_loop_gap, min_good=p.min_good,
                      filter_floor=p.filter_floor, filter_strict=p.filter_strict, min_init=p.min_init,
                      gap_strict=p.gap_strict, thr_inclusive=p.thr_inclusive, **kw)


def _check_pair(got, q, t, p=Params(), use_cv2=True):
    want = c_oracle.match_features(q, t, p, nthreads=8)
    assert len(got["train_idx"]) == len(want["train_idx"])
    np.testing.assert_array_equal(got["train_idx"], want["train_idx"])
    np.testing.assert_array_equal(got["dist"], want["dist"])
    np.testing.assert_array_equal(got["good"], want["good"])
    assert got["min_dist"] == want["min_dist"] and got["n_good"] == want["n_good"]
    if use_cv2 and cv2_ref.available():
        ref = cv2_ref.match_features(q, t, p)
        np.testing.assert_array_equal(got["train_idx"], ref["train_idx"])
        np.testing.assert_array_equal(got["dist"], ref["dist"])
        np.testing.assert_array_equal(got["good"], ref["good"])
        assert got["min_dist"] == ref["min_dist"] and got["n_good"] == ref["n_good"]


@pytest.mark.parametrize("kernel", KERNELS)
def test_pair_c2_uniform_2000x2000(lc, kernel):
    """BASELINE config 2: 2000 x 2000 consecutive-frame match + 2 x min filter."""
    q, t = datagen.uniform(2000, datagen.QUERY_SEED), datagen.uniform(2000, datagen.DB_SEED)
    with _ctx(lc, kernel=kernel) as c:
        _check_pair(c.match_pair(q, t), q, t)
        assert c.last_timing().path == 5 and c.launch_count() == 1  # one launch per pair match


def test_pair_timing_modes(lc):
    """lc_config.timing: pair matches are event-timed only on request (LC_TIMING_ON); path and launches always report."""
    q, t = datagen.uniform(700, 1), datagen.uniform(900, 2)
    with lc.Context() as c:
        _check_pair(c.match_pair(q, t), q, t)
        tm = c.last_timing()
        assert (tm.path, tm.kernel_launches) == (5, 1) and tm.kernel_ms == 0.0 and tm.total_ms == 0.0
    with lc.Context(timing=1) as c:
        _check_pair(c.match_pair(q, t), q, t)
        tm = c.last_timing()
        assert (tm.path, tm.kernel_launches) == (5, 1) and tm.kernel_ms > 0.0 and tm.total_ms >= tm.kernel_ms
        assert tm.match_ms == tm.kernel_ms and tm.d2h_ms == 0.0  # results are stored straight into pinned host memory
    with pytest.raises(lc.LcError):
        lc.Context(timing=7)


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("nq,nt", [(1, 1), (1, 700), (700, 1), (31, 33), (32, 32), (33, 31), (255, 257), (300, 700),
                                   (1999, 2000), (2000, 1999), (2081, 2081), (5000, 300), (300, 5000), (4097, 513)])
def test_pair_shapes_low_entropy_ties(lc, kernel, nq, nt):
    """low-entropy rows: many tied minima -> lowest train index must win, exactly as cv::BFMatcher."""
    q, t = datagen.low_entropy(nq, 10 + nq), datagen.low_entropy(nt, 20 + nt)
    with _ctx(lc, kernel=kernel) as c:
        _check_pair(c.match_pair(q, t), q, t)


@pytest.mark.parametrize("nq,nt", [(40000, 900),     # results past the direct-to-pinned limit: device buffer + one copy
                                   (300, 530000),    # rows past the pinned upload buffer: the two-copy staging road
                                   (70000, 70),      # many query chunks, one short segment each
                                   (9, 300000)])     # one warp of queries against many full waves of segments
def test_pair_large_shapes_take_the_other_roads(lc, nq, nt):
    """the single-launch pair kernel beyond the sizes its fast paths are built for; same context reused afterwards."""
    q, t = datagen.low_entropy(nq, 3), datagen.uniform(nt, 4)
    t[nt // 3] = q[0]  # an exact copy: min_dist 0, so the 2 x min filter keeps only distance-0 matches
    want = c_oracle.match_features(q, t, nthreads=16)
    with _ctx(lc) as c:
        for _ in range(2):
            got = c.match_pair(q, t)
            assert c.last_timing().path == 5
            np.testing.assert_array_equal(got["train_idx"], want["train_idx"])
            np.testing.assert_array_equal(got["dist"], want["dist"])
            np.testing.assert_array_equal(got["good"], want["good"])
            assert (got["min_dist"], got["n_good"]) == (want["min_dist"], want["n_good"]) and got["min_dist"] == 0
        small_q, small_t = datagen.uniform(50, 8), datagen.uniform(60, 9)
        _check_pair(c.match_pair(small_q, small_t), small_q, small_t)  # the self-cleaning buffers are back at idle


@pytest.mark.parametrize("kernel", KERNELS)
def test_pair_planted_duplicates_filter_bites(lc, kernel):
    q, t = datagen.planted_pair(1500, 1800, 7)
    with _ctx(lc, kernel=kernel) as c:
        got = c.match_pair(q, t)
        _check_pair(got, q, t)
        assert 0 < got["n_good"] < len(q)  # the 2 x min filter really removes matches here


def test_pair_empty_sides(lc):
    q = datagen.uniform(10, 1)
    e = np.zeros((0, 32), np.uint8)
    with _ctx(lc) as c:
        for a, b in ((q, e), (e, q), (e, e)):
            got = c.match_pair(a, b)
            assert len(got["train_idx"]) == 0 and got["min_dist"] == -1 and got["n_good"] == 0


def test_pair_identical_rows_and_zero_distance(lc):
    row = datagen.uniform(1, 3)
    q = np.repeat(row, 100, axis=0)
    t = np.repeat(row, 257, axis=0)
    with _ctx(lc) as c:
        got = c.match_pair(q, t)
        assert (got["train_idx"] == 0).all() and (got["dist"] == 0).all() and got["min_dist"] == 0
        assert got["n_good"] == 100  # d <= 2*0
        c.set_params(filter_strict=1)  # B1: "<" keeps nothing when min_dist == 0
        assert c.match_pair(q, t)["n_good"] == 0
    _check_pair(got, q, t)


def test_pair_max_distance_256(lc):
    q = np.zeros((5, 32), np.uint8)
    t = np.full((3, 32), 255, np.uint8)
    with _ctx(lc) as c:
        got = c.match_pair(q, t)
    assert (got["dist"] == 256).all() and (got["train_idx"] == 0).all()
    _check_pair(got, q, t)


def test_pair_strided_rows(lc):
    """non-continuous cv::Mat: row stride 48 bytes."""
    rng = np.random.default_rng(5)
    big_q = rng.integers(0, 256, (400, 48), np.uint8)
    big_t = rng.integers(0, 256, (333, 64), np.uint8)
    q, t = big_q[:, :32], big_t[:, 16:48]
    with _ctx(lc) as c:
        _check_pair(c.match_pair(q, t), np.ascontiguousarray(q), np.ascontiguousarray(t))


@pytest.mark.parametrize("p", [Params(filter_floor=30), Params(filter_strict=True), Params(min_init=100),
                               Params(filter_floor=64, filter_strict=True, min_init=20)])
def test_pair_filter_knobs(lc, p):
    q, t = datagen.planted_pair(900, 1100, 11)
    with _ctx(lc, p) as c:
        _check_pair(c.match_pair(q, t), q, t, p)


# ------------------------------------------------------------------ loop check
def _check_loop(c, q, kfs, fids, cur, p, use_cv2=True, nthreads=16):
    res, best = c.loop_check(q, cur)
    db = np.concatenate(kfs) if len(kfs) else np.zeros((0, 32), np.uint8)
    want_res, want_best = c_oracle.loop_check(q, db, [len(k) for k in kfs], fids, cur, p, nthreads=nthreads)
    np.testing.assert_array_equal(res, want_res)
    np.testing.assert_array_equal(best, want_best)
    if use_cv2 and cv2_ref.available():
        ref_res, ref_best = cv2_ref.loop_check(q, kfs, fids, cur, p)
        np.testing.assert_array_equal(res, ref_res)
        np.testing.assert_array_equal(best, ref_best)
    cands = c.select_candidates(res, len(q), cur)
    want_c = c_oracle.loop_candidates(want_res, len(q), cur, p)
    assert [(int(x["current_frame_id"]), int(x["matched_frame_id"]), int(x["num_matches"]), float(x["similarity_score"]))
            for x in cands] == want_c
    det = c.detect_loops(q, cur)
    assert det.tobytes() == cands.tobytes()
    return res, best, cands


@pytest.mark.parametrize("kernel", KERNELS)
def test_loop_small_db_split_path(lc, kernel):
    """config-1-like: a few dozen keyframes -> split kernel + finalize; gap excludes the recent ones."""
    q = datagen.uniform(2000, datagen.QUERY_SEED)
    kfs = datagen.structured_db(q, 40, every=5, ragged=True)
    fids = list(range(0, 120, 3))
    p = Params(loop_threshold=0.15)
    with _ctx(lc, p, kernel=kernel) as c:
        c.db_append_many(fids, kfs)
        res, best, cands = _check_loop(c, q, kfs, fids, 119, p)
        assert c.last_timing().path == 2
        assert (res["n_good"][-9:] == -1).all() and (res["n_good"][:30] >= 0).all()  # 119 - id >= 30
        assert len(cands) > 0


@pytest.mark.parametrize("kernel", KERNELS)
def test_loop_scan_path_ragged_ties(lc, kernel):
    """enough keyframes for the whole-keyframe scan kernel; short low-entropy ragged keyframes incl. empty ones."""
    rng = np.random.default_rng(21)
    q = datagen.low_entropy(700, 5)
    ns = rng.integers(0, 400, 6000)
    ns[::97] = 0
    kfs = [datagen.low_entropy(int(n), 1000 + i) for i, n in enumerate(ns)]
    fids = list(range(len(kfs)))
    p = Params(loop_threshold=0.15, min_loop_gap=30)
    with _ctx(lc, p, kernel=kernel) as c:
        c.db_append_many(fids, kfs)
        res, best, _ = _check_loop(c, q, kfs, fids, 5990, p, use_cv2=False)
        assert c.last_timing().path == 1
        assert (res["n_good"][ns == 0] == -1).all()


@pytest.mark.parametrize("kernel", KERNELS)
def test_loop_c3_1k_keyframes(lc, kernel):
    """BASELINE config 3: 2000-descriptor query vs 1k keyframes x 2000 rows (structured variant), vs the C oracle."""
    q = datagen.uniform(2000, datagen.QUERY_SEED)
    kfs = datagen.structured_db(q, 1000, every=10)
    fids = list(range(1000))
    p = Params(loop_threshold=0.15)
    with _ctx(lc, p, kernel=kernel) as c:
        c.db_append_many(fids, kfs)
        res, best, cands = _check_loop(c, q, kfs, fids, 1029, p, use_cv2=False)
        assert (res["n_good"] >= 0).all()
        # README-literal rule: unrelated random keyframes pass "<= 2*min" wholesale (similarity 1.0) while the
        # planted noisy copies (small min_dist) keep few matches - both kinds must appear
        assert 0 < len(cands) < 1000 and (res["n_good"][9::10] < 2000).all()
        # cv2 on a sample of keyframes (the real OpenCV matcher)
        if cv2_ref.available():
            for k in (0, 9, 19, 500, 999):
                ref = cv2_ref.match_features(q, kfs[k], p)
                assert (res[k]["min_dist"], res[k]["n_good"]) == (ref["min_dist"], ref["n_good"])


def test_loop_scan_path_big_query_frames(lc):
    """nq just under / over the register-tile capacities (2048, 4096) on the scan path."""
    for nq in (2049, 4096):
        q = datagen.uniform(nq, 77)
        kfs = [datagen.uniform(24, 5000 + i) for i in range(5000)]
        fids = list(range(5000))
        p = Params(min_loop_gap=0)
        for kernel, path in ((2, 1), (0, 4)):  # carry-save scan kernel; AUTO -> tensor cores
            with _ctx(lc, p, kernel=kernel) as c:
                c.db_append_many(fids, kfs)
                _check_loop(c, q, kfs, fids, 6000, p, use_cv2=False)
                assert c.last_timing().path == path


def test_loop_query_larger_than_scan_capacity(lc):
    q = datagen.low_entropy(4500, 3)
    kfs = [datagen.low_entropy(50, 100 + i) for i in range(300)]
    fids = list(range(300))
    p = Params(min_loop_gap=1)
    for kernel, path in ((2, 2), (0, 4)):  # split + finalize on the integer kernels; AUTO -> tensor cores
        with _ctx(lc, p, kernel=kernel) as c:
            c.db_append_many(fids, kfs)
            _check_loop(c, q, kfs, fids, 400, p, use_cv2=False)
            assert c.last_timing().path == path


def test_loop_empty_query_empty_db_and_gap_knobs(lc):
    q = datagen.uniform(100, 1)
    with _ctx(lc, Params(min_loop_gap=5)) as c:
        res, best = c.loop_check(q, 10)
        assert len(res) == 0 and (best == np.iinfo(np.uint64).max).all()
        kfs = [datagen.uniform(60, 10 + i) for i in range(8)]
        c.db_append_many(list(range(8)), kfs)
        res, best = c.loop_check(np.zeros((0, 32), np.uint8), 100)
        assert len(res) == 8 and (res["n_good"] == -1).all() and len(best) == 0
        res, _ = c.loop_check(q, 10)  # ids 0..5 eligible with >=, 6,7 not
        assert (res["n_good"][:6] >= 0).all() and (res["n_good"][6:] == -1).all()
        c.set_params(gap_strict=1)
        res, _ = c.loop_check(q, 10)
        assert (res["n_good"][:5] >= 0).all() and (res["n_good"][5:] == -1).all()
        # nothing eligible -> best stays UINT64_MAX
        res, best = c.loop_check(q, 0)
        assert (res["n_good"] == -1).all() and (best == np.iinfo(np.uint64).max).all()


def test_loop_incremental_appends_and_clear(lc):
    """processFrame's pattern: append one keyframe, check, repeat (device table grows between calls)."""
    p = Params(min_loop_gap=3, loop_threshold=0.15)
    base = datagen.uniform(500, 9)
    kfs, fids = [], []
    with _ctx(lc, p, slab_rows=4096) as c:  # tiny slabs: keyframes spill into new slabs
        for i in range(40):
            kf = datagen.noisy_copy(base, 100 + i, p=0.03) if i % 4 == 0 else datagen.uniform(500 + i, 200 + i)
            assert c.db_append(i, kf) == i
            kfs.append(kf)
            fids.append(i)
            _check_loop(c, kf, kfs, fids, i, p, use_cv2=(i % 13 == 0))
        assert c.db_size() == dict(keyframes=40, rows=sum(map(len, kfs)), local_keyframes=40, local_rows=sum(map(len, kfs)))
        fid, rows = c.db_read(17)
        assert fid == 17 and (rows == kfs[17]).all()
        c.db_clear()
        assert c.db_size()["keyframes"] == 0
        assert len(c.loop_check(base, 100)[0]) == 0


def test_loop_batch_matches_single(lc):
    """throughput mode (BASELINE config 5 shape, scaled down): batch == the same frames one by one."""
    p = Params(loop_threshold=0.15, min_loop_gap=30)
    qs = [datagen.uniform(n, 300 + i) for i, n in enumerate((2000, 1500, 2000, 0, 700))]
    curs = [500, 400, 40, 500, 35]
    kfs = datagen.structured_db(qs[0], 120, n=600, every=7, ragged=True)
    fids = list(range(120))
    with _ctx(lc, p) as c:
        c.db_append_many(fids, kfs)
        res_b, best_b = c.loop_check_batch(qs, curs)
        for f, (q, cur) in enumerate(zip(qs, curs)):
            res, best = c.loop_check(q, cur)
            np.testing.assert_array_equal(res_b[f], res)
            np.testing.assert_array_equal(best_b[f], best)
            db = np.concatenate(kfs)
            w_res, w_best = c_oracle.loop_check(q, db, [len(k) for k in kfs], fids, cur, p, nthreads=16)
            np.testing.assert_array_equal(res, w_res)
            np.testing.assert_array_equal(best, w_best)


def test_match_vs_keyframe_resident(lc):
    q = datagen.uniform(800, 1)
    kfs = [datagen.uniform(900, 2), datagen.noisy_copy(q, 3, p=0.05), np.zeros((0, 32), np.uint8)]
    with _ctx(lc) as c:
        c.db_append_many([0, 1, 2], kfs)
        for k in (0, 1):
            _check_pair(c.match_vs_keyframe(q, k), q, kfs[k])
        assert len(c.match_vs_keyframe(q, 2)["dist"]) == 0


def test_match_keyframes_compacted_candidate_matches(lc):
    """lc_match_keyframes: the DMatch lists the loop candidates hand to estimatePose (include/loop_closing.hpp:43-53)."""
    p = Params(loop_threshold=0.15, min_loop_gap=30)
    q = datagen.uniform(700, 5)
    kfs = datagen.structured_db(q, 40, n=600, every=4, ragged=True)
    kfs[9] = np.zeros((0, 32), np.uint8)
    with _ctx(lc, p) as c:
        c.db_append_many(list(range(40)), kfs)
        sel = [0, 4, 9, 17, 39, 4]
        got = c.match_keyframes(q, sel)
        assert len(got) == len(sel)
        for k, (m, mind) in zip(sel, got):
            want = c_oracle.match_features(q, kfs[k], p) if len(kfs[k]) else None
            if want is None:
                assert len(m) == 0 and mind == -1
                continue
            keep = np.flatnonzero(want["good"])
            np.testing.assert_array_equal(m["query_idx"], keep)
            np.testing.assert_array_equal(m["train_idx"], want["train_idx"][keep])
            np.testing.assert_array_equal(m["distance"], want["dist"][keep])
            assert mind == want["min_dist"] and len(m) == want["n_good"]
            if cv2_ref.available():
                ref = cv2_ref.match_features(q, kfs[k], p)
                np.testing.assert_array_equal(m["train_idx"], ref["train_idx"][np.flatnonzero(ref["good"])])
        assert c.match_keyframes(q, []) == []
        assert all(len(m) == 0 and d == -1 for m, d in c.match_keyframes(np.zeros((0, 32), np.uint8), [1, 2]))


def test_loop_check_begin_end_overlaps_host_work(lc):
    """lc_loop_check_begin / _end: the same results as lc_loop_check, with the host free in between."""
    p = Params(loop_threshold=0.15, min_loop_gap=30)
    q = datagen.uniform(2000, 21)
    kfs = datagen.structured_db(q, 600, n=500, every=9, ragged=True)
    with _ctx(lc, p) as c:
        c.db_append_many(list(range(600)), kfs)
        want_res, want_best = c.loop_check(q, 700)
        c.loop_check_begin(q, 700)
        q[:] = 0  # the library copied the rows: the caller's buffer is free again
        busy = sum(int(x) for x in np.arange(20000))  # "ORB of the next frame"
        with pytest.raises(lc.LcError):  # no second check while one is pending
            c.loop_check(datagen.uniform(10, 1), 700)
        res, best = c.loop_check_end()
        assert busy > 0
        np.testing.assert_array_equal(res, want_res)
        np.testing.assert_array_equal(best, want_best)
        with pytest.raises(lc.LcError):
            c.loop_check_end()  # nothing pending any more
        c.loop_check_begin(datagen.uniform(300, 2), 700, want_best=False)
        res2, best2 = c.loop_check_end()
        assert best2 is None and len(res2) == 600


def test_other_calls_between_loop_check_begin_and_end(lc):
    """the streamed front-end in use: while a loop check is pending the caller matches the NEXT pair of frames, appends a
    keyframe and fetches match lists - none of it may disturb the pending check's results (their pinned buffer is its own)."""
    p = Params(loop_threshold=0.15, min_loop_gap=30)
    q = datagen.uniform(2000, 31)
    kfs = datagen.structured_db(q, 300, n=700, every=7, ragged=True)
    nxt, nxt2 = datagen.uniform(1800, 32), datagen.uniform(1900, 33)
    with _ctx(lc, p) as c:
        c.db_append_many(list(range(300)), kfs)
        want_res, want_best = c.loop_check(q, 400)
        c.loop_check_begin(q, 400)
        _check_pair(c.match_pair(nxt, nxt2), nxt, nxt2)          # results land in the synchronous calls' buffer
        _check_pair(c.match_vs_keyframe(nxt, 5), nxt, kfs[5])
        ml = c.match_keyframes(nxt, [0, 7])
        assert len(ml) == 2
        assert c.db_append(300, nxt) == 300                       # not part of the pending check
        res, best = c.loop_check_end()
        np.testing.assert_array_equal(res, want_res)
        np.testing.assert_array_equal(best, want_best)
        res3, _ = c.loop_check(q, 400)                            # the appended keyframe shows up in the next one
        assert len(res3) == 301 and np.array_equal(res3[:300], want_res)


def test_plain_c_client_finds_its_loop(lc, tmp_path):
    """examples/abi_client.c: the C ABI driven from plain C99 - no Python, torch or OpenCV between caller and library."""
    import os
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("gcc not found")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib, exe = lc.library_path(), str(tmp_path / "abi_client")
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "abi_client.c"),
                           lib, "-Wl,-rpath," + os.path.dirname(lib), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "keyframe 7 recognised with all matches" in r.stdout, r.stdout + r.stderr


def test_sharded_contexts_merge_equals_single(lc):
    """two shard contexts on one GPU: merged per-keyframe records and min-merged best == unsharded."""
    p = Params(loop_threshold=0.15, min_loop_gap=10)
    q = datagen.low_entropy(900, 4)  # ties straddle shards
    kfs = [datagen.low_entropy(int(n), 50 + i) for i, n in enumerate(np.random.default_rng(1).integers(0, 300, 101))]
    fids = list(range(101))
    with _ctx(lc, p) as c0:
        c0.db_append_many(fids, kfs)
        res0, best0 = c0.loop_check(q, 150)
    for world in (2, 3, 8):
        merged = np.zeros(len(kfs), res0.dtype)
        best = np.full(len(q), np.iinfo(np.uint64).max, np.uint64)
        for r in range(world):
            with _ctx(lc, p, shard_rank=r, shard_world=world) as c:
                c.db_append_many(fids, [kf if k % world == r else len(kf) for k, kf in enumerate(kfs)])
                res, b = c.loop_check(q, 150)
                merged[r::world] = res
                best = np.minimum(best, b)
        np.testing.assert_array_equal(merged, res0)
        np.testing.assert_array_equal(best, best0)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_merge_shards_kernel_equals_unsharded(lc, world):
    """the device side of the multi-GPU exchange on ONE GPU: every rank's NativeShard writes its [records | bests]
    contribution, the buffers are concatenated as the all_gather would leave them, and lc_merge_shards_device must give
    the unsharded check's answer (and what the torch restatement of the merge gives)."""
    import torch
    from slam_loop_closing_b200.sharded import NativeShard, merge_torch
    p = Params(loop_threshold=0.15, min_loop_gap=10)
    qs = [datagen.low_entropy(700, 4), datagen.low_entropy(33, 5), np.zeros((0, 32), np.uint8)]  # ties straddle shards
    curs = [150, 12, 99]  # the second frame sees keyframes 0..2 only: some shards compare nothing (UINT64_MAX bests)
    kfs = [datagen.low_entropy(int(n), 50 + i) for i, n in enumerate(np.random.default_rng(1).integers(0, 300, 61))]
    fids = list(range(61))
    with _ctx(lc, p) as c0:
        c0.db_append_many(fids, kfs)
        res0, best0 = c0.loop_check_batch(qs, curs)
    nq = np.array([len(q) for q in qs], np.int32)
    F, rows, slots = len(qs), int(nq.sum()), (len(kfs) + world - 1) // world
    rec_bytes = F * slots * 16
    contrib_bytes = (rec_bytes + rows * 8 + 15) & ~15
    shards = [NativeShard(r, world, 0, loop_threshold=p.loop_threshold, min_loop_gap=p.min_loop_gap) for r in range(world)]
    gathered = torch.zeros(world * contrib_bytes, dtype=torch.uint8, device="cuda")
    d_q = shards[0].to_device(np.concatenate(qs))
    for r, sh in enumerate(shards):
        for k, (fid, kf) in enumerate(zip(fids, kfs)):
            sh.append(fid, kf if k % world == r else None, len(kf))
        contrib = gathered[r * contrib_bytes:(r + 1) * contrib_bytes]
        sh.loop_check(d_q, nq, np.array(curs, np.int32), slots,
                      out=(contrib[:rec_bytes].view(torch.int32).view(F, slots, 4), contrib[rec_bytes:rec_bytes + rows * 8].view(torch.int64)))
    rec, best = shards[0].merge(gathered, contrib_bytes, world, F, slots, rows)
    rec_t, best_t = merge_torch(gathered, contrib_bytes, world, F, slots, rows)
    torch.cuda.synchronize()
    assert torch.equal(rec, rec_t) and torch.equal(best, best_t)
    got = rec.cpu().numpy()[:, :len(kfs)].copy().view(res0.dtype).reshape(F, len(kfs))
    np.testing.assert_array_equal(got, res0)
    np.testing.assert_array_equal(best.cpu().numpy().view(np.uint64), np.concatenate(best0))
    for sh in shards:
        sh.ctx.close()


def test_store_save_load_roundtrip(lc, tmp_path):
    p = Params(min_loop_gap=0)
    q = datagen.uniform(300, 8)
    kfs = [datagen.uniform(int(n), 60 + i) for i, n in enumerate((100, 0, 257, 31))]
    path = str(tmp_path / "store.lcdb")
    with _ctx(lc, p) as c:
        c.db_append_many([5, 6, 7, 8], kfs)
        res, best = c.loop_check(q, 50)
        c.db_save(path)
    with _ctx(lc, p) as c:
        c.db_load(path)
        res2, best2 = c.loop_check(q, 50)
        np.testing.assert_array_equal(res, res2)
        np.testing.assert_array_equal(best, best2)
        assert c.db_append(9, kfs[0]) == 4  # appends continue after a load


def test_errors_are_status_codes_not_crashes(lc):
    with _ctx(lc) as c:
        with pytest.raises(lc.LcError):
            c.db_read(0)
        with pytest.raises((ValueError, TypeError)):
            c.match_pair(np.zeros((3, 31), np.uint8), np.zeros((3, 32), np.uint8))
        with pytest.raises(lc.LcError):
            lc.Context(device=99)
        with pytest.raises(lc.LcError):
            lc.Context(shard_rank=2, shard_world=2)
