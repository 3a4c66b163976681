"""GPU parity of the shipped pipeline's SIFT / L2 matcher (lc_l2_*: knnMatch(k = 2) + Lowe ratio, reference
src/main.cpp:509-534 and its loop search :1375-1392) against the committed cv2 fixtures (tests/golden/l2_golden.npz)
and the CPU oracle (oracle/l2_oracle.c) - bit-exact: indices, float32 distances, ratio decisions, counts."""
import os

import numpy as np
import pytest

from oracle import l2_oracle

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "l2_golden.npz"))
PAIRS = sorted({k.split("/")[1] for k in GOLD.files if k.startswith("pair/")})


@pytest.fixture(scope="module")
def lc():
    import slam_loop_closing_b200 as m
    return m


def _check_pair(r, want):
    np.testing.assert_array_equal(r["train_idx"], want["train_idx"])
    has1 = want["train_idx"] >= 0
    np.testing.assert_array_equal(r["distance"][has1], want["distance"][has1])
    np.testing.assert_array_equal(r["good"], want["good"])
    assert r["n_good"] == int(np.sum(want["good"]))


@pytest.mark.parametrize("name", PAIRS)
def test_l2_match_pair_equals_golden(lc, name):
    q, t = GOLD[f"pair/{name}/q"], GOLD[f"pair/{name}/t"]
    with lc.Context() as c:
        for ratio in (0.75, 0.7):
            r = c.l2_match_pair(q, t, ratio)
            want = dict(train_idx=GOLD[f"pair/{name}/train_idx"], distance=GOLD[f"pair/{name}/distance"],
                        good=GOLD[f"pair/{name}/good_{ratio}"])
            _check_pair(r, want)
            has2 = GOLD[f"pair/{name}/second"] >= 0
            np.testing.assert_array_equal(r["second"][has2], GOLD[f"pair/{name}/second"][has2])
            assert (r["second"][~has2] == -1).all()
        rf = c.l2_match_pair(q.astype(np.float32), t.astype(np.float32), 0.7)  # CV_32F rows, as cv::SIFT hands them out
        np.testing.assert_array_equal(rf["train_idx"], r["train_idx"])
        np.testing.assert_array_equal(rf["distance"], r["distance"])


def _loop_case():
    n = GOLD["loop/n"]
    offs = np.concatenate([[0], np.cumsum(n)])
    descs = [GOLD["loop/rows"][offs[i]:offs[i + 1]] for i in range(len(n))]
    gap, min_rows, ratio = GOLD["loop/params"]
    return descs, int(gap), int(min_rows), float(ratio)


def test_l2_sift_loop_sequence_equals_golden(lc):
    """real cv2.SIFT descriptors of a synthetic closed-loop sequence: every (curr, past <= curr - gap) pair the loop
    search of src/main.cpp:1378-1392 visits, match counts as OpenCV gives them."""
    descs, gap, min_rows, ratio = _loop_case()
    want = GOLD["loop/n_good"]
    with lc.Context() as c:
        for curr, d in enumerate(descs):
            assert c.l2_db_append(curr, d.astype(np.float32)) == curr
            res = c.l2_loop_check(d, curr, gap, ratio, min_rows)
            np.testing.assert_array_equal(res["n_good"], want[curr, :curr + 1])
            np.testing.assert_array_equal(res["frame_id"], np.arange(curr + 1))
        r = c.l2_match_pair(descs[-1], descs[0], ratio)
        _check_pair(r, dict(train_idx=GOLD["loop/last_vs_first/train_idx"], distance=GOLD["loop/last_vs_first/distance"],
                            good=GOLD["loop/last_vs_first/good"]))
        c.l2_db_clear()
        assert c.l2_db_size()["keyframes"] == 0 and len(c.l2_loop_check(descs[0], 5, 1)) == 0


@pytest.mark.parametrize("nq,nt", [(1, 2), (127, 129), (128, 128), (1025, 300), (4000, 4000), (2500, 1)])
def test_l2_pair_shapes_vs_oracle(lc, nq, nt):
    rng = np.random.default_rng(nq * 7 + nt)
    q = rng.integers(0, 256, (nq, 128)).astype(np.uint8)
    t = rng.integers(0, 256, (nt, 128)).astype(np.uint8)
    if nt > 40 and nq > 20:  # exact copies and duplicates: zero distances, ties on the train index
        t[7], t[33], t[34] = q[3], q[3], q[10]
    with lc.Context() as c:
        r = c.l2_match_pair(q, t, 0.8)
    w = l2_oracle.match_features(q, t, 0.8, nthreads=16)
    np.testing.assert_array_equal(r["train_idx"], w["train_idx"])
    np.testing.assert_array_equal(r["distance"], w["distance"])
    np.testing.assert_array_equal(r["good"], w["good"])
    assert r["n_good"] == w["n_good"]


def test_l2_loop_check_many_views_vs_oracle(lc):
    """a machine's worth of views (ragged, some below min_rows, some too recent) against the oracle"""
    rng = np.random.default_rng(5)
    q = rng.integers(0, 200, (1500, 128)).astype(np.uint8)
    views = []
    for k in range(220):
        n = int(rng.integers(90, 400))
        v = rng.integers(0, 200, (n, 128)).astype(np.uint8)
        if k % 9 == 0:  # a view that really overlaps the query
            m = min(n, 300)
            v[:m] = np.clip(q[:m].astype(np.int32) + rng.integers(-2, 3, (m, 128)), 0, 255).astype(np.uint8)
        views.append(v)
    with lc.Context() as c:
        for k, v in enumerate(views):
            c.l2_db_append(k, v)
        res = c.l2_loop_check(q, 230, 20, 0.7, 100)
    want = l2_oracle.loop_check(q, views, list(range(220)), 230, 20, 0.7, 100, nthreads=16)
    np.testing.assert_array_equal(res, want)
    big = np.array([len(v) >= 100 for v in views])
    assert (res["n_good"][:211][big[:211]] >= 0).all() and (res["n_good"][~big] == -1).all() and (res["n_good"][211:] == -1).all()
    assert (res["n_good"][:211:9][big[:211:9]] > 100).all()  # the overlapping views stand out


def test_l2_second_neighbour_inside_and_outside_the_best_rows_group(lc):
    """the kernel ranks aligned groups of 8 train rows first and looks inside the best group only when the ratio test can
    still pass: views whose two nearest rows sit in one group, straddle a group boundary, or ARE the whole view
    (n <= 8), through the loop check (counts only) and the pair API (exact second distance)."""
    rng = np.random.default_rng(11)
    q = rng.integers(0, 200, (600, 128)).astype(np.uint8)

    def near(rows, amp):
        return np.clip(rows.astype(np.int32) + rng.integers(-amp, amp + 1, rows.shape), 0, 255).astype(np.uint8)

    views = []
    for n in (3, 8, 9, 130, 700, 700, 700, 1500):
        views.append(rng.integers(0, 200, (n, 128)).astype(np.uint8))
    views[0][:2] = near(q[:2], 1)                      # the whole view is one (partial) group
    views[1][:8] = near(q[:8], 1)                      # exactly one full group
    views[2][7], views[2][8] = near(q[5:6], 1)[0], near(q[5:6], 2)[0]   # nearest and second straddle rows 7 | 8
    for v, amp2 in ((views[4], 1), (views[5], 3), (views[6], 40)):
        for i in range(0, 300, 3):                     # nearest at row 2i + 100, second right beside it (same group or next)
            v[2 * i + 100] = near(q[i:i + 1], 1)[0]
            v[2 * i + 101] = near(q[i:i + 1], amp2)[0]
    views[7][::2][:600] = near(q, 2)                   # a well-separated nearest row for every query: most pass
    with lc.Context() as c:
        for k, v in enumerate(views):
            c.l2_db_append(k, v)
        for ratio in (0.7, 0.95):
            res = c.l2_loop_check(q, 100, 1, ratio, 1)
            want = l2_oracle.loop_check(q, views, list(range(len(views))), 100, 1, ratio, 1, nthreads=16)
            np.testing.assert_array_equal(res, want)
        assert res["n_good"][7] > 400 and res["n_good"][4] < res["n_good"][6]
        for v in views:
            r = c.l2_match_pair(q, v, 0.8)
            w = l2_oracle.match_features(q, v, 0.8, nthreads=16)
            np.testing.assert_array_equal(r["train_idx"], w["train_idx"])
            np.testing.assert_array_equal(r["distance"], w["distance"])
            np.testing.assert_array_equal(r["good"], w["good"])
            has2 = w["second_dist2"] >= 0  # the exact second distance, whichever side of the group boundary it came from
            np.testing.assert_array_equal(r["second"][has2], np.sqrt(w["second_dist2"][has2].astype(np.float32)))
            assert (r["second"][~has2] == -1).all()


@pytest.mark.parametrize("world", [2, 3])
def test_l2_sharded_stores_merge_equals_unsharded(lc, world):
    """view k lives on the context with k % world == rank (lc_config.shard_rank / shard_world, as the Hamming store);
    the per-view records of the shards, interleaved, are the unsharded check's (ShardedL2DB does this over NCCL)."""
    from slam_loop_closing_b200.sharded import NativeL2Shard
    rng = np.random.default_rng(13)
    q = rng.integers(0, 200, (900, 128)).astype(np.uint8)
    views = [rng.integers(0, 200, (int(rng.integers(60, 500)), 128)).astype(np.uint8) for _ in range(31)]
    for k in range(0, 31, 4):
        m = min(len(views[k]), 250)
        views[k][:m] = np.clip(q[:m].astype(np.int32) + rng.integers(-2, 3, (m, 128)), 0, 255).astype(np.uint8)
    with lc.Context() as c0:
        for k, v in enumerate(views):
            c0.l2_db_append(k, v)
        want = c0.l2_loop_check(q, 40, 12, 0.7, 100)
    merged = np.zeros(len(views), want.dtype)
    for r in range(world):
        sh = NativeL2Shard(r, world, 0)
        for k, v in enumerate(views):
            sh.append(k, v if k % world == r else None, len(v))
        assert sh.ctx.l2_db_size() == dict(keyframes=len(views), local_keyframes=len(range(r, len(views), world)))
        merged[r::world] = sh.loop_check(q, 40, 12, 0.7, 100)
        sh.ctx.close()
    np.testing.assert_array_equal(merged, want)
    assert (want["n_good"][:29:4] > 100).all() and (want["n_good"][29:] == -1).all()


def test_l2_rejects_non_sift_values(lc):
    with lc.Context() as c:
        with pytest.raises(lc.LcError):
            c.l2_db_append(0, np.full((3, 128), 0.5, np.float32))
        with pytest.raises(lc.LcError):
            c.l2_match_pair(np.full((3, 128), 300.0, np.float32), np.zeros((3, 128), np.float32))


def test_l2_loop_search_over_the_resident_store(lc):
    """lc_l2_loop_search: the shipped loop search's double loop (src/main.cpp:1375-1389: every curr >= loopGap against every
    past <= curr - loopGap) in one call, the queries being the stored views themselves - against the oracle pair by pair and
    against the per-view call; more query views than one launch takes (chunks of 64), ragged views, one below min_rows, one empty."""
    rng = np.random.default_rng(11)
    V, gap = 75, 6
    views = []
    for k in range(V):
        n = 0 if k == 40 else 60 if k == 13 else int(rng.integers(150, 420))
        v = rng.integers(0, 200, (n, 128)).astype(np.uint8)
        if k >= 30 and k % 4 == 0 and len(views[k - 30]) and n:  # the walk returns: view k overlaps view k - 30
            m = min(n, len(views[k - 30]), 250)
            v[:m] = np.clip(views[k - 30][:m].astype(np.int32) + rng.integers(-2, 3, (m, 128)), 0, 255).astype(np.uint8)
        views.append(v)
    with lc.Context() as c:
        for k, v in enumerate(views):
            c.l2_db_append(k, v)
        res = c.l2_loop_search(gap, V - gap, gap, 0.7, 100)
        assert res.shape == (V - gap, V)
        for f, curr in enumerate(range(gap, V)):
            want = l2_oracle.loop_check(views[curr], views, list(range(V)), curr, gap, 0.7, 100, nthreads=16)
            np.testing.assert_array_equal(res[f], want, err_msg=f"curr {curr}")
            if curr % 17 == 0:
                np.testing.assert_array_equal(res[f], c.l2_loop_check(views[curr], curr, gap, 0.7, 100))
        assert (res["n_good"] > 60).sum() >= 5  # the revisits stand out
        assert (res[40 - gap]["n_good"] == -1).all() and (res[13 - gap]["n_good"] == -1).all()  # empty / too small query views
        with pytest.raises(lc.LcError):
            c.l2_loop_search(70, 10, gap)  # past the end of the store
        assert c.l2_loop_search(0, 0, gap).shape == (0, V)
    with lc.Context(shard_rank=0, shard_world=2) as c:
        c.l2_db_append(0, views[0])
        with pytest.raises(lc.LcError):
            c.l2_loop_search(0, 1, 0)
