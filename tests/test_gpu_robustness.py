"""GPU tier, robustness of the store and the boundary (round-2 hardening): the 2^32-row limit of the packed best,
stream ordering of the device-pointer entry points, corrupt store files, timing of a pending check."""
import struct

import numpy as np
import pytest

import datagen
from oracle import c_oracle
from oracle.params import Params

pytestmark = pytest.mark.gpu

MAX_ROWS_PER_KF = (1 << 20) - 1  # LC_MAX_ROWS_PER_KF
MAX_STORE_ROWS = 1 << 32         # LC_MAX_STORE_ROWS


@pytest.fixture(scope="module")
def lc():
    import slam_loop_closing_b200 as m
    return m


@pytest.mark.parametrize("kernel", [0, 2, 4])  # AUTO, CSA (integer kernels), MXF4 (tensor path + resolve)
def test_global_rows_up_to_the_32_bit_limit(lc, kernel):
    """best_packed = dist << 32 | global_row.  A shard that owns only the LAST keyframe of a store of almost 2^32 rows
    (the other shard's keyframes are recorded by size only, so no memory is needed) must report global rows above 2^31
    exactly, accept appends up to LC_MAX_STORE_ROWS and refuse the one that crosses it without changing the store."""
    q = datagen.uniform(500, 77)
    kf = np.concatenate([datagen.uniform(2500, 78), q[::-1]])  # rows 2500.. are the query reversed: distance 0
    with lc.Context(shard_rank=1, shard_world=2, min_loop_gap=0, kernel=kernel) as c:
        n_pairs = 4096
        # even keyframes live on shard 0 (size only), odd ones here and are empty
        c.db_append_many(list(range(2 * n_pairs)), [MAX_ROWS_PER_KF if k % 2 == 0 else None for k in range(2 * n_pairs)])
        base = n_pairs * MAX_ROWS_PER_KF
        assert base == MAX_STORE_ROWS - 4096 and c.db_size()["rows"] == base
        c.db_append_meta(2 * n_pairs, 0)
        assert c.db_append(2 * n_pairs + 1, kf) == 2 * n_pairs + 1
        res, best = c.loop_check(q, 10 ** 6)
        assert len(res) == n_pairs + 1 and (res["n_good"][:-1] == -1).all()  # empty keyframes are not compared
        assert (int(res[-1]["min_dist"]), int(res[-1]["n_good"]), int(res[-1]["n_kf"])) == (0, len(q), len(kf))
        want = np.uint64(base + 2500) + np.arange(len(q) - 1, -1, -1, dtype=np.uint64)  # distance 0: high word is 0
        np.testing.assert_array_equal(best, want)
        assert int(best.min()) > (1 << 31)
        # checked against the oracle with the same global offset
        w_res, w_best = c_oracle.loop_check(q, kf, [len(kf)], [2 * n_pairs + 1], 10 ** 6, Params(min_loop_gap=0), nthreads=4)
        np.testing.assert_array_equal(best, w_best + np.uint64(base))
        # fill to the limit exactly, then one row too many
        left = MAX_STORE_ROWS - c.db_size()["rows"]
        assert 0 < left < MAX_ROWS_PER_KF
        c.db_append_meta(2 * n_pairs + 2, left)
        before = c.db_size()
        assert before["rows"] == MAX_STORE_ROWS
        with pytest.raises(lc.LcError) as e:
            c.db_append(2 * n_pairs + 3, q[:1])
        assert e.value.code == 4 and "LC_MAX_STORE_ROWS" in str(e.value)  # LC_ERR_CAPACITY
        with pytest.raises(lc.LcError):
            c.db_append_meta(2 * n_pairs + 3, 1)
        assert c.db_size() == before
        c.db_append_meta(2 * n_pairs + 3, 0)  # an empty keyframe still fits
        res2, best2 = c.loop_check(q, 10 ** 6)
        np.testing.assert_array_equal(best2, want)


def test_append_many_device_waits_for_the_producer_stream(lc):
    """lc_db_append_many_device(…, cuda_stream): the device-to-device copies are ordered behind the stream that is still
    WRITING the rows when the call is made (the copies used to race with torch's generator kernels)."""
    import torch
    src = torch.from_numpy(datagen.uniform(3 * 800, 5)).cuda()
    rows = torch.zeros_like(src)
    side = torch.cuda.Stream()
    torch.cuda.synchronize()
    with lc.Context(min_loop_gap=0) as c:
        with torch.cuda.stream(side):
            torch.cuda._sleep(400_000_000)  # ~0.2 s: the rows are certainly not written yet when the library is called
            rows.copy_(src)
            c.db_append_many_device(np.arange(3, dtype=np.int32), np.full(3, 800, np.int32), rows.data_ptr(), side.cuda_stream)
            rows.zero_()  # stream-ordered reuse right after the call must be safe too
        for k in range(3):
            _, got = c.db_read(k)
            np.testing.assert_array_equal(got, src[k * 800:(k + 1) * 800].cpu().numpy())
        side.synchronize()
        assert int(rows.max()) == 0


def test_calls_on_different_streams_share_the_scratch_safely(lc):
    """One context, its device-pointer entry points on two caller streams and its host API in between, no synchronisation
    by the caller except before reading: the library chains the calls by events (they share the pair keys, the distance
    scratch and the query images), so every result equals the serial one."""
    import torch
    p = Params(loop_threshold=0.15, min_loop_gap=0)
    q1, q2 = datagen.uniform(1500, 1), datagen.low_entropy(900, 2)
    kfs = datagen.structured_db(q1, 200, n=400, every=5, ragged=True)
    fids = list(range(200))
    rows, ns = np.concatenate(kfs), [len(k) for k in kfs]
    want = {id(q): c_oracle.loop_check(q, rows, ns, fids, 500, p, nthreads=8) for q in (q1, q2)}
    want_pair = c_oracle.match_features(q1, kfs[4], p, nthreads=8)
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    with lc.Context(loop_threshold=p.loop_threshold, min_loop_gap=0) as c:
        c.db_append_many(fids, kfs)
        d_q1, d_q2, d_t = (torch.from_numpy(x).cuda() for x in (q1, q2, kfs[4]))
        outs = []
        torch.cuda.synchronize()
        for it in range(6):
            st, q, d_q = ((sa, q1, d_q1), (sb, q2, d_q2))[it % 2]
            rec = torch.empty((200, 4), dtype=torch.int32, device="cuda")
            best = torch.empty(len(q), dtype=torch.int64, device="cuda")
            c.loop_check_device(d_q.data_ptr(), len(q), 500, rec.data_ptr(), best.data_ptr(), st.cuda_stream)
            idx = torch.empty(len(q1), dtype=torch.int32, device="cuda")
            dist = torch.empty(len(q1), dtype=torch.int32, device="cuda")
            good = torch.empty(len(q1), dtype=torch.uint8, device="cuda")
            summ = torch.empty(2, dtype=torch.int32, device="cuda")
            other = sb if st is sa else sa
            c.match_pair_device(d_q1.data_ptr(), len(q1), d_t.data_ptr(), len(kfs[4]), idx.data_ptr(), dist.data_ptr(), good.data_ptr(),
                                summ.data_ptr(), other.cuda_stream)
            outs.append((q, rec, best, idx, dist, good, summ))
            if it % 3 == 2:  # the host API joins in: it runs on the context's own stream
                got = c.match_pair(q1, kfs[4])
                np.testing.assert_array_equal(got["train_idx"], want_pair["train_idx"])
                r, b = c.loop_check(q2, 500)
                np.testing.assert_array_equal(r, want[id(q2)][0])
                np.testing.assert_array_equal(b, want[id(q2)][1])
        torch.cuda.synchronize()
        for q, rec, best, idx, dist, good, summ in outs:
            w_res, w_best = want[id(q)]
            np.testing.assert_array_equal(rec.cpu().numpy().view(w_res.dtype).reshape(-1), w_res)
            np.testing.assert_array_equal(best.cpu().numpy().view(np.uint64), w_best)
            np.testing.assert_array_equal(idx.cpu().numpy(), want_pair["train_idx"])
            np.testing.assert_array_equal(dist.cpu().numpy(), want_pair["dist"])
            np.testing.assert_array_equal(good.cpu().numpy().astype(bool), want_pair["good"])
            assert tuple(summ.cpu().numpy()) == (want_pair["min_dist"], want_pair["n_good"])


def _corrupt(tmp_path, name, blob, edit):
    b = bytearray(blob)
    b = edit(b) or b
    path = str(tmp_path / name)
    open(path, "wb").write(bytes(b))
    return path


def test_db_load_rejects_corrupt_files_and_leaves_an_empty_usable_store(lc, tmp_path):
    """lc_db_load does not trust the file: counts, global indices, row offsets and the length are checked; every
    rejected file leaves the store empty, and a good file still loads afterwards."""
    q = datagen.uniform(300, 8)
    kfs = [datagen.uniform(int(n), 60 + i) for i, n in enumerate((100, 0, 257, 31))]
    good = str(tmp_path / "good.lcdb")
    with lc.Context(min_loop_gap=0) as c:
        c.db_append_many([5, 6, 7, 8], kfs)
        want = c.loop_check(q, 50)
        c.db_save(good)
    blob = open(good, "rb").read()
    HDR = 8 + 12 + 24  # magic, {version, rank, world}, {g_kf, g_rows, n_local}
    REC = 24           # {i64 global index, i64 row offset, i32 frame id, i32 n}
    assert struct.unpack_from("<qqq", blob, 20) == (4, 388, 4)

    def put(fmt, off, *v):
        return lambda b: struct.pack_into(fmt, b, off, *v)

    bad = {
        "magic": put("<8s", 0, b"LCB200DX"),
        "version": put("<I", 8, 7),
        "other_shard": put("<II", 12, 1, 2),
        "n_local_negative": put("<q", 36, -1),
        "n_local_huge": put("<q", 36, 1 << 40),
        "n_local_vs_g_kf": put("<q", 20, 9),
        "g_rows_negative": put("<q", 28, -5),
        "g_rows_small": put("<q", 28, 200),
        "g_rows_over_limit": put("<q", 28, (1 << 32) + 1),
        "global_index": put("<q", HDR + 2 * REC, 5),
        "row_offset_backwards": put("<q", HDR + 2 * REC + 8, 50),
        "row_offset_past_end": put("<q", HDR + 3 * REC + 8, 380),
        "n_negative": put("<i", HDR + REC + 20, -3),
        "n_too_large": put("<i", HDR + 2 * REC + 20, 1 << 20),
        "truncated_rows": lambda b: b[:-100],
        "truncated_table": lambda b: b[:HDR + REC + 7],
        "truncated_header": lambda b: b[:30],
        "trailing_bytes": lambda b: b + b"\0" * 32,
    }
    with lc.Context(min_loop_gap=0) as c:
        for name, edit in bad.items():
            path = _corrupt(tmp_path, name + ".lcdb", blob, edit)
            with pytest.raises(lc.LcError) as e:
                c.db_load(path)
            assert e.value.code in (1, 5), name  # LC_ERR_INVALID (another shard's file) or LC_ERR_IO
            assert c.db_size() == dict(keyframes=0, rows=0, local_keyframes=0, local_rows=0), name
            with pytest.raises(lc.LcError):
                c.db_read(0)
        c.db_load(good)
        got = c.loop_check(q, 50)
        np.testing.assert_array_equal(got[0], want[0])
        np.testing.assert_array_equal(got[1], want[1])
        _, rows = c.db_read(2)
        np.testing.assert_array_equal(rows, kfs[2])
        with pytest.raises(lc.LcError):  # a loaded store is not empty any more
            c.db_load(good)


def test_sharded_store_file_roundtrip_and_validation(lc, tmp_path):
    """shard 1 of 3: the table must hold exactly the keyframes k % 3 == 1 of a store of g_kf keyframes."""
    kfs = [datagen.uniform(40 + k, 90 + k) for k in range(8)]
    path = str(tmp_path / "shard1.lcdb")
    q = datagen.uniform(64, 3)
    with lc.Context(min_loop_gap=0, shard_rank=1, shard_world=3) as c:
        c.db_append_many(list(range(8)), [kf if k % 3 == 1 else len(kf) for k, kf in enumerate(kfs)])
        want = c.loop_check(q, 100)
        c.db_save(path)
    with lc.Context(min_loop_gap=0, shard_rank=1, shard_world=3) as c:
        c.db_load(path)
        assert c.db_size() == dict(keyframes=8, rows=sum(len(k) for k in kfs), local_keyframes=3, local_rows=41 + 44 + 47)
        got = c.loop_check(q, 100)
        np.testing.assert_array_equal(got[0], want[0])
        np.testing.assert_array_equal(got[1], want[1])
        _, rows = c.db_read(7)
        np.testing.assert_array_equal(rows, kfs[7])
    blob = bytearray(open(path, "rb").read())
    struct.pack_into("<q", blob, 20, 11)  # 11 keyframes would put 4 on this shard, the table has 3
    bad = str(tmp_path / "bad.lcdb")
    open(bad, "wb").write(bytes(blob))
    with lc.Context(min_loop_gap=0, shard_rank=1, shard_world=3) as c:
        with pytest.raises(lc.LcError):
            c.db_load(bad)
        assert c.db_size()["keyframes"] == 0
    with lc.Context(min_loop_gap=0, shard_rank=0, shard_world=3) as c:
        with pytest.raises(lc.LcError):
            c.db_load(path)


def test_timing_of_a_pending_check_survives_calls_made_meanwhile(lc):
    """lc_loop_check_begin ... lc_match_pair ... lc_loop_check_end: the pair match runs untimed behind the check and
    lc_get_last_timing describes the check again after _end (it used to be reset by the pair match)."""
    q = datagen.uniform(2000, 21)
    kfs = datagen.structured_db(q, 300, n=500, every=9)
    a, b = datagen.uniform(700, 1), datagen.uniform(800, 2)
    with lc.Context(min_loop_gap=0, timing=1) as c:
        c.db_append_many(list(range(300)), kfs)
        want = c.loop_check(q, 400)
        t0 = c.last_timing()
        assert t0.path == 4 and t0.match_ms > 0 and t0.total_ms >= t0.kernel_ms > 0
        c.loop_check_begin(q, 400)
        got_pair = c.match_pair(a, b)
        tp = c.last_timing()
        assert tp.path == 5 and tp.kernel_launches == 1 and tp.total_ms == 0  # reported, not timed
        res, best = c.loop_check_end()
        t1 = c.last_timing()
        assert t1.path == 4 and t1.match_launches == t0.match_launches and t1.match_ms > 0 and t1.total_ms >= t1.kernel_ms > 0
        np.testing.assert_array_equal(res, want[0])
        np.testing.assert_array_equal(best, want[1])
        w = c_oracle.match_features(a, b, Params(), nthreads=4)
        np.testing.assert_array_equal(got_pair["train_idx"], w["train_idx"])
        # and an ordinary call afterwards is timed again
        c.match_pair(a, b)
        assert c.last_timing().total_ms > 0


def test_mxf4_peak_probe_reports_a_plausible_pipe_rate(lc):
    """lc_probe_mxf4_peak: the roofline denominator bench.py measures live - 16384 MAC/clk/SM x 148 SMs at 1.5-1.97 GHz."""
    with lc.Context() as c:
        macs, mhz = c.probe_mxf4_peak(30)
        assert 1400 < mhz < 2000, mhz
        per_clk_per_sm = macs / (mhz * 1e6) / 148
        assert 15000 < per_clk_per_sm < 16500, per_clk_per_sm
        with pytest.raises(lc.LcError):
            c.probe_mxf4_peak(0)


def test_sharded_db_results_as_views_survive_one_more_call(lc):
    """ShardedLoopDB.loop_check(copy=False): records and bests are views of two pinned download buffers used in turn."""
    from slam_loop_closing_b200.sharded import ShardedLoopDB
    p = Params(loop_threshold=0.15, min_loop_gap=0)
    q1, q2 = datagen.uniform(700, 1), datagen.low_entropy(650, 2)
    kfs = datagen.structured_db(q1, 90, n=300, every=4, ragged=True)
    fids = list(range(90))
    rows, ns = np.concatenate(kfs), [len(k) for k in kfs]
    db = ShardedLoopDB(device=0, loop_threshold=0.15, min_loop_gap=0)
    for fid, kf in zip(fids, kfs):
        db.append(fid, kf)
    rec_a, best_a = db.loop_check([q1], [200], copy=False)
    rec_b, best_b = db.loop_check([q2], [200], copy=False)
    for (rec, best), q in (((rec_a, best_a), q1), ((rec_b, best_b), q2)):
        w_res, w_best = c_oracle.loop_check(q, rows, ns, fids, 200, p, nthreads=8)
        np.testing.assert_array_equal(rec[0], w_res)
        np.testing.assert_array_equal(best[0], w_best)
    db.shard.ctx.close()
