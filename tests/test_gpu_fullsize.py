"""GPU tier, BASELINE config 4 at FULL size (100k keyframes x 2000 rows, 6.4 GB): size-independent properties of the
loop check - a planted row-permuted copy of the query is found exactly (record and every per-query best), repeated
checks are identical, the gap rule cuts exactly the most recent keyframes, and sampled keyframes agree with the CPU oracle."""
import os
import sys

import numpy as np
import pytest

from oracle import c_oracle
from oracle.params import Params

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_c4_full_size_properties():
    import torch
    import bench
    import slam_loop_closing_b200 as lc
    from slam_loop_closing_b200.sharded import ShardedLoopDB

    n_kf, nq = 100_000, bench.NQ
    free, _ = torch.cuda.mem_get_info()
    if free < 12 << 30:
        pytest.skip("needs 12 GB of free device memory")
    q = bench.query_rows()
    db = ShardedLoopDB(device=0, loop_threshold=0.15)
    ctx = db.shard.ctx
    q_dev = db.shard.to_device(q)
    plant_at = n_kf // 3
    perm = bench.build_db(db, torch, n_kf, 0, 1, torch.device("cuda", 0), plant_at, q_dev)
    assert ctx.db_size()["rows"] == n_kf * bench.ROWS_PER_KF
    cur = n_kf + 29
    res, best = ctx.loop_check(q, cur)
    assert ctx.last_timing().path == 4
    # records: every keyframe compared, ids in insertion order
    assert len(res) == n_kf and (res["frame_id"] == np.arange(n_kf)).all() and (res["n_kf"] == bench.ROWS_PER_KF).all()
    assert (res["min_dist"] >= 0).all() and (res["n_good"] >= 1).all() and (res["n_good"] <= nq).all()
    # the planted permuted copy: distance 0 for every query, all of them good, best row = the copy's row
    assert res[plant_at]["min_dist"] == 0 and res[plant_at]["n_good"] == nq
    want_best = np.uint64(plant_at) * np.uint64(bench.ROWS_PER_KF) + np.argsort(perm).astype(np.uint64)
    np.testing.assert_array_equal(best, want_best)
    # uniform random rows: the 2 x min filter passes everything elsewhere, min distances ~ 80-100
    others = np.delete(res, plant_at)
    assert (others["n_good"] == nq).all() and 60 <= others["min_dist"].min() and others["min_dist"].max() <= 110
    # idempotence
    res2, best2 = ctx.loop_check(q, cur)
    np.testing.assert_array_equal(res, res2)
    np.testing.assert_array_equal(best, best2)
    # the gap rule at full size: only keyframes at least min_loop_gap back are compared
    res3, _ = ctx.loop_check(q, n_kf - 1 + 10)
    assert (res3["n_good"][:n_kf - 20] >= 0).all() and (res3["n_good"][n_kf - 20:] == -1).all()
    np.testing.assert_array_equal(res3[:n_kf - 20], res[:n_kf - 20])
    # sampled keyframes against the CPU oracle
    for k in (0, 1, plant_at - 1, plant_at, n_kf // 2, n_kf - 1):
        _, rows = ctx.db_read(k)
        w = c_oracle.match_features(q, rows, Params(), nthreads=16)
        assert (int(res[k]["min_dist"]), int(res[k]["n_good"])) == (w["min_dist"], w["n_good"]), k
    ctx.close()
