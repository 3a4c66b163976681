"""GPU parity of the tensor-core paths (LC_KERNEL_TENSOR: tcgen05.mma kind::i8, LC_KERNEL_MXF4: kind::mxf4, both on
+1/-1 expanded bits) vs the CPU oracle - bit-exact, same records / packed best / tie-breaks as the integer kernels."""
import numpy as np
import pytest

import datagen
from oracle import c_oracle, cv2_ref
from oracle.params import Params

pytestmark = pytest.mark.gpu
KERNELS = [3, 4]  # LC_KERNEL_TENSOR (int8), LC_KERNEL_MXF4; timing.path reports the same numbers


@pytest.fixture(scope="module")
def lc():
    import slam_loop_closing_b200 as m
    return m


def _ctx(lc, p, kernel, **kw):
    return lc.Context(loop_threshold=p.loop_threshold, min_loop_gap=p.min_loop_gap, min_good=p.min_good,
                      filter_floor=p.filter_floor, filter_strict=p.filter_strict, min_init=p.min_init,
                      gap_strict=p.gap_strict, thr_inclusive=p.thr_inclusive, kernel=kernel, **kw)


def _check(c, q, kfs, fids, cur, p, expect_path):
    res, best = c.loop_check(q, cur)
    assert c.last_timing().path == expect_path
    db = np.concatenate(kfs) if len(kfs) else np.zeros((0, 32), np.uint8)
    w_res, w_best = c_oracle.loop_check(q, db, [len(k) for k in kfs], fids, cur, p, nthreads=16)
    np.testing.assert_array_equal(res, w_res)
    np.testing.assert_array_equal(best, w_best)
    return res, best


@pytest.mark.parametrize("nq", [1, 127, 128, 129, 512, 513, 700, 1023, 1024, 1025, 2000, 2081])
@pytest.mark.parametrize("kernel", KERNELS)
def test_tensor_query_sizes_low_entropy_ties(lc, kernel, nq):
    """query counts around the 128-row MMA tile and the 512-query resident block; tied minima everywhere."""
    p = Params(min_loop_gap=0, loop_threshold=0.15)
    q = datagen.low_entropy(nq, 1000 + nq)
    ns = np.random.default_rng(nq).integers(1, 420, 300)
    kfs = [datagen.low_entropy(int(n), 2000 + i) for i, n in enumerate(ns)]
    with _ctx(lc, p, kernel) as c:
        c.db_append_many(list(range(300)), kfs)
        _check(c, q, kfs, list(range(300)), 1000, p, kernel)


@pytest.mark.parametrize("rows", [1, 15, 16, 17, 31, 33, 127, 128, 129, 191, 192, 193, 223, 224, 225, 255, 256, 257, 383, 384, 385, 2000])
@pytest.mark.parametrize("kernel", KERNELS)
def test_tensor_keyframe_row_counts(lc, kernel, rows):
    """train rows around the 16-row MMA N granule and the 192-row B tile."""
    p = Params(min_loop_gap=0)
    q = datagen.uniform(600, 7)
    kfs = [datagen.uniform(rows, 3000 + i) for i in range(200)]
    kfs[5] = datagen.noisy_copy(q, 11, p=0.02, n=rows)
    with _ctx(lc, p, kernel) as c:
        c.db_append_many(list(range(200)), kfs)
        _check(c, q, kfs, list(range(200)), 500, p, kernel)


@pytest.mark.parametrize("kernel", KERNELS)
def test_tensor_extreme_distances(lc, kernel):
    """distance 0 (dot = +256) and 256 (dot = -256), identical rows -> lowest index, complement rows."""
    p = Params(min_loop_gap=0)
    q = datagen.uniform(300, 1)
    kfs = []
    for i in range(200):
        kf = datagen.uniform(100, 50 + i)
        if i % 3 == 0:
            kf[10:20] = q[:10]          # exact copies (dist 0), duplicated below
            kf[60:70] = q[:10]
        if i % 3 == 1:
            kf[:] = ~q[3]               # every row is the complement of query 3 (dist 256 for it)
        kfs.append(kf)
    with _ctx(lc, p, kernel) as c:
        c.db_append_many(list(range(200)), kfs)
        res, best = _check(c, q, kfs, list(range(200)), 500, p, kernel)
        assert res["min_dist"][0] == 0
    allq = np.repeat(q[:1], 64, axis=0)
    kfs = [np.repeat(~q[:1], 40, axis=0) for _ in range(200)]
    with _ctx(lc, p, kernel) as c:
        c.db_append_many(list(range(200)), kfs)
        res, best = _check(c, allq, kfs, list(range(200)), 500, p, kernel)
        assert (res["min_dist"] == 256).all() and (best == (np.uint64(256) << np.uint64(32))).all()


@pytest.mark.parametrize("kernel", KERNELS)
def test_tensor_ragged_db_with_gap_and_empty_keyframes(lc, kernel):
    p = Params(loop_threshold=0.15, min_loop_gap=30, filter_floor=30)
    q = datagen.uniform(2000, datagen.QUERY_SEED)
    kfs = datagen.structured_db(q, 400, n=500, every=6, ragged=True)
    for i in range(0, 400, 41):
        kfs[i] = np.zeros((0, 32), np.uint8)
    fids = list(range(400))
    with _ctx(lc, p, kernel) as c:
        c.db_append_many(fids, kfs)
        res, _ = _check(c, q, kfs, fids, 420, p, kernel)
        assert (res["n_good"][391:] == -1).all() and (res["n_good"][::41] == -1).all()
        cands = c.select_candidates(res, len(q), 420)
        want = c_oracle.loop_candidates(res, len(q), 420, p)
        assert [(int(x["matched_frame_id"]), int(x["num_matches"]), float(x["similarity_score"])) for x in cands] == \
               [(m, n, s) for (_, m, n, s) in want]


@pytest.mark.parametrize("kernel", KERNELS)
def test_tensor_c3_1k_keyframes_vs_oracle_and_cv2(lc, kernel):
    """BASELINE config 3 on the tensor path."""
    p = Params(loop_threshold=0.15)
    q = datagen.uniform(2000, datagen.QUERY_SEED)
    kfs = datagen.structured_db(q, 1000, every=10)
    fids = list(range(1000))
    with _ctx(lc, p, kernel) as c:
        c.db_append_many(fids, kfs)
        res, _ = _check(c, q, kfs, fids, 1029, p, kernel)
        if cv2_ref.available():
            for k in (0, 9, 19, 500, 999):
                ref = cv2_ref.match_features(q, kfs[k], p)
                assert (res[k]["min_dist"], res[k]["n_good"]) == (ref["min_dist"], ref["n_good"])


@pytest.mark.parametrize("kernel", KERNELS)
def test_tensor_batch_of_frames(lc, kernel):
    p = Params(loop_threshold=0.15, min_loop_gap=30)
    qs = [datagen.uniform(n, 300 + i) for i, n in enumerate((2000, 1500, 600, 0, 700))]
    curs = [500, 400, 40, 500, 35]
    kfs = datagen.structured_db(qs[0], 150, n=400, every=7, ragged=True)
    fids = list(range(150))
    with _ctx(lc, p, kernel) as c:
        c.db_append_many(fids, kfs)
        res_b, best_b = c.loop_check_batch(qs, curs)
        assert c.last_timing().path == kernel
        db = np.concatenate(kfs)
        for f, (q, cur) in enumerate(zip(qs, curs)):
            w_res, w_best = c_oracle.loop_check(q, db, [len(k) for k in kfs], fids, cur, p, nthreads=16)
            np.testing.assert_array_equal(res_b[f], w_res)
            np.testing.assert_array_equal(best_b[f], w_best)


@pytest.mark.parametrize("kernel", KERNELS)
def test_tensor_small_problem_falls_back_to_integer_kernels(lc, kernel):
    """too little work to fill the GPU (and every pair match): the CSA kernels answer, results identical."""
    p = Params(min_loop_gap=0)
    q = datagen.uniform(500, 1)
    kfs = [datagen.uniform(300, 10 + i) for i in range(20)]
    with _ctx(lc, p, kernel) as c:
        c.db_append_many(list(range(20)), kfs)
        _check(c, q, kfs, list(range(20)), 100, p, 2)
        got = c.match_pair(q, kfs[0])
        want = c_oracle.match_features(q, kfs[0], p)
        np.testing.assert_array_equal(got["train_idx"], want["train_idx"])
        np.testing.assert_array_equal(got["dist"], want["dist"])
