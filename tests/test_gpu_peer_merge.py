"""GPU tier: the peer-memory merge of a keyframe-sharded check (lc_exchange_* / lc_merge_shards_peer_device) on ONE GPU.
Several ranks are emulated by several contexts of this process whose exchange buffers are connected by device address;
kernels of different ranks must not wait for each other on one GPU, so every rank first signals (phase 1) and then merges
(phase 2).  The real thing - one process per GPU, CUDA IPC, phase 3 - is exercised by bench.py's full-size gate at N > 1."""
import numpy as np
import pytest

import datagen
from oracle.params import Params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lc():
    import slam_loop_closing_b200 as m
    return m


@pytest.mark.parametrize("world", [2, 3, 8])
def test_peer_merge_equals_unsharded_over_three_sequences(lc, world):
    import torch
    from slam_loop_closing_b200.sharded import NativeShard, _wrap_device_memory
    p = Params(loop_threshold=0.15, min_loop_gap=10)
    kfs = [datagen.low_entropy(int(n), 50 + i) for i, n in enumerate(np.random.default_rng(1).integers(0, 300, 61))]
    fids = list(range(61))
    shards = [NativeShard(r, world, 0, loop_threshold=p.loop_threshold, min_loop_gap=p.min_loop_gap) for r in range(world)]
    for r, sh in enumerate(shards):
        for k, (fid, kf) in enumerate(zip(fids, kfs)):
            sh.append(fid, kf if k % world == r else None, len(kf))
    slot_bytes = 1 << 16
    bases = [sh.ctx.exchange_create(slot_bytes)[1] for sh in shards]
    for r, sh in enumerate(shards):
        sh.ctx.exchange_connect(world, r, bases=bases)
    with lc.Context(loop_threshold=p.loop_threshold, min_loop_gap=p.min_loop_gap) as c0:
        c0.db_append_many(fids, kfs)
        # three checks in a row: both slots get reused, the flags keep counting
        for seq, (qs, curs) in enumerate([([datagen.low_entropy(700, 4), datagen.low_entropy(33, 5), np.zeros((0, 32), np.uint8)], [150, 12, 99]),
                                          ([datagen.low_entropy(500, 6)], [80]),
                                          ([datagen.uniform(64, 7), datagen.low_entropy(300, 8)], [61, 200])], start=1):
            res0, best0 = c0.loop_check_batch(qs, curs)
            nq = np.array([len(q) for q in qs], np.int32)
            F, rows, slots = len(qs), int(nq.sum()), (len(kfs) + world - 1) // world
            rec_bytes = F * slots * 16
            d_q = shards[0].to_device(np.concatenate(qs))
            for sh in shards:  # every rank's contribution into its own slot of this sequence number
                ptr, n = sh.ctx.exchange_slot(seq)
                assert n == slot_bytes and rec_bytes + rows * 8 <= n
                slot = _wrap_device_memory(ptr, n, sh.device)
                sh.loop_check(d_q, nq, np.array(curs, np.int32), slots,
                              out=(slot[:rec_bytes].view(torch.int32).view(F, slots, 4), slot[rec_bytes:rec_bytes + rows * 8].view(torch.int64)))
            outs = []
            for sh in shards:  # phase 1: signal only
                sh.ctx.merge_shards_peer_device(seq, F, slots, rows, None, None, phase=1, stream=1)
            for sh in shards:  # phase 2: the flags are all there, no kernel waits for another
                rec = torch.empty((F, slots * world, 4), dtype=torch.int32, device="cuda")
                best = torch.empty(rows, dtype=torch.int64, device="cuda")
                sh.ctx.merge_shards_peer_device(seq, F, slots, rows, rec.data_ptr(), best.data_ptr(), phase=2, stream=1)
                outs.append((rec, best))
            torch.cuda.synchronize()
            for rec, best in outs:  # every rank ends up with the unsharded answer
                got = rec.cpu().numpy()[:, :len(kfs)].copy().view(res0.dtype).reshape(F, len(kfs))
                np.testing.assert_array_equal(got, res0)
                np.testing.assert_array_equal(best.cpu().numpy().view(np.uint64), np.concatenate(best0))
    for sh in shards:
        sh.ctx.exchange_close()
    for sh in shards:
        sh.ctx.close()


def test_exchange_argument_checks(lc):
    with lc.Context() as c:
        with pytest.raises(lc.LcError):
            c.merge_shards_peer_device(1, 1, 4, 10, None, None)   # no exchange yet
        with pytest.raises(lc.LcError):
            c.exchange_create(24)                                   # not a multiple of 16
        handle, base = c.exchange_create(4096)
        assert len(handle) == 64 and base
        with pytest.raises(lc.LcError):
            c.exchange_connect(17, 0, bases=[base] * 17)            # more than 16 ranks
        c.exchange_connect(1, 0, bases=[base])
        with pytest.raises(lc.LcError):
            c.merge_shards_peer_device(1, 1, 1000, 0, None, None)  # 16 kB of records do not fit a 4 kB slot
        with pytest.raises(lc.LcError):
            c.exchange_slot(0)                                      # sequence numbers start at 1
        c.exchange_close()
