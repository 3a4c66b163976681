"""Where does a sharded end-to-end loop check spend its host time?  torchrun --nproc-per-node N tools/e2e_probe.py K"""
import os
import sys
import time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
import bench
from slam_loop_closing_b200.sharded import ShardedLoopDB

K = int(sys.argv[1]) if len(sys.argv) > 1 else 25000
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
db = ShardedLoopDB(device=local, loop_threshold=0.15)
q = bench.query_rows()
q_dev = db.shard.to_device(q)
bench.build_db(db, torch, K, rank, world, dev, K // 3, q_dev)
cur = K + 29
for _ in range(5):
    db.loop_check([q], [cur])
dist.barrier(); torch.cuda.synchronize()
T = {}
def lap(name, t0):
    torch.cuda.synchronize()
    T[name] = T.get(name, 0.0) + time.perf_counter() - t0
    return time.perf_counter()
n = 30
t_all = time.perf_counter()
for _ in range(n):
    db.loop_check([q], [cur])
torch.cuda.synchronize()
t_all = (time.perf_counter() - t_all) / n
for _ in range(n):
    t = time.perf_counter()
    nq = np.array([len(q)], np.int32)
    d_q = db.shard.to_device(q); t = lap("to_device", t)
    slots = (db.n_keyframes + world - 1) // world
    rec_bytes = slots * 16
    contrib_bytes = (rec_bytes + int(nq.sum()) * 8 + 15) & ~15
    contrib = db.shard.new_buffer(contrib_bytes)
    rec = contrib[:rec_bytes].view(torch.int32).view(1, slots, 4)
    best = contrib[rec_bytes:rec_bytes + int(nq.sum()) * 8].view(torch.int64)
    db.shard.loop_check(d_q, nq, np.array([cur], np.int32), slots, out=(rec, best)); t = lap("shard.loop_check (kernels)", t)
    g = db.shard.new_buffer(world * contrib_bytes)
    dist.all_gather_into_tensor(g, contrib); t = lap("all_gather [records | bests]", t)
    merged, best = db.shard.merge(g, contrib_bytes, world, 1, slots, int(nq.sum())); t = lap("merge kernel", t)
    r_h, b_h = db._to_host(merged.reshape(-1), best); t = lap("D2H + sync", t)
    out = r_h.view(np.int32).reshape(1, -1, 4)[:, :K].copy(); t = lap("numpy copy", t)
if rank == 0:
    print(f"world {world} K {K}: loop_check e2e {t_all * 1e3:.3f} ms per call")
    for k, v in T.items():
        print(f"  {k:32s} {v / n * 1e3:8.3f} ms (with a device sync after the phase)")
dist.destroy_process_group()
