"""ShardedL2DB over NCCL against the unsharded check on rank 0: torchrun --nproc-per-node N tools/l2_sharded_probe.py"""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
import slam_loop_closing_b200 as lc
from slam_loop_closing_b200.sharded import ShardedL2DB

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rng = np.random.default_rng(21)
q = rng.integers(0, 200, (3000, 128)).astype(np.uint8)
views = [rng.integers(0, 200, (int(rng.integers(150, 3000)), 128)).astype(np.uint8) for _ in range(61)]
for k in range(0, 61, 5):
    m = min(len(views[k]), 800)
    views[k][:m] = np.clip(q[:m].astype(np.int32) + rng.integers(-2, 3, (m, 128)), 0, 255).astype(np.uint8)
db = ShardedL2DB(device=local)
for k, v in enumerate(views):
    db.append(k, v if db.owns(k) else None, n=len(v))
res = db.loop_check(q, 70, 20, 0.7, 100)
if rank == 0:
    with lc.Context(device=local) as c:
        for k, v in enumerate(views):
            c.l2_db_append(k, v)
        want = c.l2_loop_check(q, 70, 20, 0.7, 100)
    ok = np.array_equal(res, want)
    print(f"world {world}: sharded L2 loop check over NCCL {'==' if ok else '!='} unsharded; n_good of the overlapping views {res['n_good'][:51:5].tolist()}")
    assert ok
dist.destroy_process_group()
