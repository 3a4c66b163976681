"""The keyframe-sharded ORB loop check on real GPUs (one process each) against the unsharded check on rank 0, while the
store grows: ragged and empty keyframes, keyframe counts that the world size does not divide, several query frames per check,
contributions that outgrow the exchange slots (the peer-memory exchange is re-created), both transports.
  torchrun --nproc-per-node N tools/sharded_check.py [rounds]"""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import datagen
import slam_loop_closing_b200 as lc
from slam_loop_closing_b200.sharded import ShardedLoopDB

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rng = np.random.default_rng(99)  # the same stream of cases on every rank
dbs = {x: ShardedLoopDB(device=local, loop_threshold=0.15, min_loop_gap=3, exchange=x) for x in ("p2p", "nccl")}
ref = lc.Context(device=local, loop_threshold=0.15, min_loop_gap=3) if rank == 0 else None
n_kf, checks, live = 0, 0, set()
for rnd in range(rounds):
    # grow the store: a burst of keyframes, some empty, some low-entropy (ties across shards)
    for _ in range(int(rng.integers(1, 2 + 40 * (rnd % 5 == 4)))):
        n = int(rng.choice([0, 37, 500, 1999, 2000, 2100]))
        kf = (datagen.low_entropy if rng.integers(0, 2) else datagen.uniform)(n, int(rng.integers(1 << 30)))
        for db in dbs.values():
            db.append(2 * n_kf, kf if db.owns(n_kf) else None, n=len(kf))
        if ref is not None:
            ref.db_append(2 * n_kf, kf)
        n_kf += 1
    F = int(rng.choice([1, 1, 2, 7, 40 if rnd % 6 == 5 else 1]))
    qs = [(datagen.low_entropy if rng.integers(0, 2) else datagen.uniform)(int(rng.integers(0, 2300)), int(rng.integers(1 << 30))) for _ in range(F)]
    curs = [int(rng.integers(0, 2 * n_kf + 10)) for _ in range(F)]
    want = ref.loop_check_batch(qs, curs) if ref is not None else None
    for name, db in dbs.items():
        rec, best = db.loop_check(qs, curs, copy=bool(rnd % 2))
        if db._xchg_bytes:
            live.add(name)
        if rank == 0:
            assert np.array_equal(rec, want[0]), (rnd, name, "records")
            for f in range(F):
                assert np.array_equal(best[f], want[1][f]), (rnd, name, "best", f)
        checks += 1
dist.barrier()
if rank == 0:
    print(f"world {world}: {checks} sharded loop checks over a store grown to {n_kf} keyframes == unsharded, bit-exact; "
          f"peer-memory exchange live on: {sorted(live)} (slots {dbs['p2p']._xchg_bytes} B)")
    ref.close()
for db in dbs.values():
    db.shard.ctx.close()
dist.destroy_process_group()
