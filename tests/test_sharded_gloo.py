"""CPU tier, world_size 2 and 3 over gloo: the host-side shard/merge logic of slam_loop_closing_b200.sharded
(round-robin ownership, sentinel padding, all_reduce(MIN) of the packed best, all_gather + re-interleave).
The per-shard compute is swapped for the CPU oracle here (the product shard needs a GPU); the merge code
under test is exactly the code that runs over NCCL on the GPUs."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import datagen
from oracle import c_oracle
from oracle.params import Params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleShard:
    """Test double with NativeShard's interface, computing on the CPU oracle."""

    def __init__(self, rank, world, p):
        self.rank, self.world, self.p = rank, world, p
        self.kfs, self.meta = [], []  # local rows; (frame_id, n, global_row_off) for local keyframes
        self.g_rows = 0
        self.g_kf = 0

    def append(self, frame_id, desc, n):
        if self.g_kf % self.world == self.rank:
            self.kfs.append(np.asarray(desc, np.uint8).reshape(-1, 32))
            self.meta.append((frame_id, n, self.g_rows))
        self.g_kf += 1
        self.g_rows += n

    def local_keyframes(self):
        return len(self.kfs)

    def to_device(self, rows):
        return torch.from_numpy(np.ascontiguousarray(rows))

    def loop_check(self, d_q, nq, cur_ids, slots, out=None):
        F = len(nq)
        rec = torch.tensor([-1, 0, -1, -1], dtype=torch.int32).repeat(F, slots, 1)
        best = torch.full((int(np.sum(nq)),), -1, dtype=torch.int64)
        if out is not None:  # fill the caller's contribution buffer halves
            out[0].copy_(rec)
            out[1].copy_(best)
            rec, best = out
        rows = d_q.numpy()
        off = 0
        for f in range(F):
            q = rows[off:off + nq[f]]
            b = np.full(nq[f], np.iinfo(np.uint64).max, np.uint64)
            for s, (kf, (fid, n, goff)) in enumerate(zip(self.kfs, self.meta)):
                res, bk = c_oracle.loop_check(q, kf, [n], [fid], int(cur_ids[f]), self.p)
                rec[f, s] = torch.tensor([res[0]["frame_id"], res[0]["n_kf"], res[0]["min_dist"], res[0]["n_good"]])
                hit = bk != np.iinfo(np.uint64).max
                bk = np.where(hit, bk + np.uint64(goff), bk)
                b = np.minimum(b, bk)
            best[off:off + nq[f]] = torch.from_numpy(b.view(np.int64))
            off += nq[f]
        return rec, best


def _case():
    q0 = datagen.low_entropy(150, 1)  # ties that straddle shards
    # the last query sees keyframes 0 and 1 only: with 3 ranks one shard compares nothing and reports UINT64_MAX
    qs = [q0, datagen.low_entropy(90, 2), np.zeros((0, 32), np.uint8), datagen.low_entropy(70, 4)]
    curs = [40, 25, 40, 6]
    ns = np.random.default_rng(3).integers(0, 60, 23)
    kfs = [datagen.low_entropy(int(n), 100 + i) for i, n in enumerate(ns)]
    return qs, curs, kfs, list(range(len(kfs))), Params(min_loop_gap=5, loop_threshold=0.15)


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from slam_loop_closing_b200.sharded import ShardedLoopDB
        qs, curs, kfs, fids, p = _case()
        db = ShardedLoopDB(shard=OracleShard(rank, world, p))
        for k, (fid, kf) in enumerate(zip(fids, kfs)):
            db.append(fid, kf if db.owns(k) else None, n=len(kf))
        rec, best = db.loop_check(qs, curs)
        if rank == 0:
            ret["rec"], ret["best"] = rec, [b.copy() for b in best]
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_merge_equals_unsharded_over_gloo(world):
    qs, curs, kfs, fids, p = _case()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    db = np.concatenate(kfs)
    for f, (q, cur) in enumerate(zip(qs, curs)):
        want_res, want_best = c_oracle.loop_check(q, db, [len(k) for k in kfs], fids, cur, p)
        np.testing.assert_array_equal(ret["rec"][f], want_res)
        np.testing.assert_array_equal(ret["best"][f], want_best)


def test_world_one_needs_no_process_group():
    from slam_loop_closing_b200.sharded import ShardedLoopDB
    qs, curs, kfs, fids, p = _case()
    db = ShardedLoopDB(shard=OracleShard(0, 1, p))
    for fid, kf in zip(fids, kfs):
        db.append(fid, kf)
    rec, best = db.loop_check(qs[:1], curs[:1])
    want_res, want_best = c_oracle.loop_check(qs[0], np.concatenate(kfs), [len(k) for k in kfs], fids, curs[0], p)
    np.testing.assert_array_equal(rec[0], want_res)
    np.testing.assert_array_equal(best[0], want_best)
    # copy=False hands out views of the download buffers: what one call returned survives the next call
    rec_a, best_a = db.loop_check(qs[:1], curs[:1], copy=False)
    rec_b, best_b = db.loop_check(qs[1:2], curs[1:2], copy=False)
    np.testing.assert_array_equal(rec_a[0], want_res)
    np.testing.assert_array_equal(best_a[0], want_best)
    w2 = c_oracle.loop_check(qs[1], np.concatenate(kfs), [len(k) for k in kfs], fids, curs[1], p)
    np.testing.assert_array_equal(rec_b[0], w2[0])
    np.testing.assert_array_equal(best_b[0], w2[1])


@pytest.mark.parametrize("world,F,slots,rows", [(1, 1, 5, 7), (2, 1, 3, 0), (3, 2, 4, 11), (8, 3, 1, 5), (4, 1, 0, 3)])
def test_merge_torch_equals_numpy_restatement(world, F, slots, rows):
    """the torch form of the shard merge (what lc_merge_shards_device is checked against on the GPU, and what the gloo run
    above uses): records re-interleaved to global keyframe order k = slot * world + rank, bests merged by UNSIGNED minimum."""
    from slam_loop_closing_b200.sharded import merge_torch
    rng = np.random.default_rng(world * 100 + F * 10 + slots + rows)
    rec_bytes = F * slots * 16
    contrib_bytes = (rec_bytes + rows * 8 + 15) & ~15
    recs = rng.integers(-5, 1000, (world, F, slots, 4)).astype(np.int32)
    bests = rng.integers(0, 1 << 63, (world, rows), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, (world, rows)).astype(np.uint64)
    if rows:
        bests[0, 0] = np.iinfo(np.uint64).max          # "nothing compared on this shard" must lose against any real value
        bests[-1, -1] = np.uint64(1) << np.uint64(63)  # the sign bit alone: larger than every value below 2^63
    buf = np.zeros((world, contrib_bytes), np.uint8)
    for r in range(world):
        buf[r, :rec_bytes] = recs[r].reshape(-1).view(np.uint8)
        buf[r, rec_bytes:rec_bytes + rows * 8] = bests[r].view(np.uint8)
    rec, best = merge_torch(torch.from_numpy(buf.reshape(-1)), contrib_bytes, world, F, slots, rows)
    want = recs.transpose(1, 2, 0, 3).reshape(F, slots * world, 4)
    np.testing.assert_array_equal(rec.numpy(), want)
    np.testing.assert_array_equal(best.numpy().view(np.uint64), bests.min(axis=0) if rows else np.zeros(0, np.uint64))


# ---- the SIFT / L2 loop search over the same partitioning (ShardedL2DB): per-view records, all_gather only
class L2OracleShard:
    """Test double with NativeL2Shard's interface, computing on the CPU L2 oracle."""

    def __init__(self, rank, world):
        self.rank, self.world, self.views, self.g = rank, world, [], 0

    def append(self, frame_id, desc, n):
        if self.g % self.world == self.rank:
            self.views.append((frame_id, np.asarray(desc)))
        self.g += 1

    def loop_check(self, q, current_frame_id, loop_gap, ratio, min_rows):
        from oracle import l2_oracle
        return l2_oracle.loop_check(q, [v for _, v in self.views], [f for f, _ in self.views], current_frame_id, loop_gap, ratio, min_rows)


def _l2_case():
    rng = np.random.default_rng(9)
    q = rng.integers(0, 200, (150, 128)).astype(np.uint8)
    views = []
    for k in range(11):
        v = rng.integers(0, 200, (int(rng.integers(5, 200)), 128)).astype(np.uint8)
        if k % 3 == 0:
            m = min(len(v), 100)
            v[:m] = np.clip(q[:m].astype(np.int32) + rng.integers(-2, 3, (m, 128)), 0, 255).astype(np.uint8)
        views.append(v)
    return q, views, list(range(0, 22, 2))


def _l2_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from slam_loop_closing_b200.sharded import ShardedL2DB
        q, views, fids = _l2_case()
        db = ShardedL2DB(shard=L2OracleShard(rank, world))
        for k, (fid, v) in enumerate(zip(fids, views)):
            db.append(fid, v if db.owns(k) else None, n=len(v))
        res = db.loop_check(q, 24, 6, 0.8, 20)
        if rank == world - 1:
            ret["res"] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_l2_loop_check_equals_unsharded_over_gloo(world):
    from oracle import l2_oracle
    q, views, fids = _l2_case()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_l2_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    want = l2_oracle.loop_check(q, views, fids, 24, 6, 0.8, 20)
    np.testing.assert_array_equal(ret["res"], want)
    assert (want["n_good"] > 0).any() and (want["n_good"] == -1).any()
