"""Why is the e2e arm's matching kernel slower than the device arm's?  Same store, the two arms in alternation."""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import bench
import slam_loop_closing_b200 as lc
from slam_loop_closing_b200.sharded import ShardedLoopDB

K = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
dev = torch.device("cuda", 0)
db = ShardedLoopDB(device=0, loop_threshold=0.15, timing=1)
ctx = db.shard.ctx
q = bench.query_rows()
q_dev = db.shard.to_device(q)
bench.build_db(db, torch, K, 0, 1, dev, K // 3, q_dev)
cur = K + 29
nq, cu = np.array([2000], np.int32), np.array([cur], np.int32)
pairs = 2000.0 * 2000 * K


def device_arm(n=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        db.merged_device(q_dev, nq, cu)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def e2e_arm(n=20):
    t0 = time.perf_counter()
    mm = []
    for _ in range(n):
        ctx.loop_check(q, cur, want_best=True)
        mm.append(ctx.last_timing().match_ms)
    return (time.perf_counter() - t0) * 1e3 / n, float(np.mean(mm))


for _ in range(5):
    db.merged_device(q_dev, nq, cu)
for rnd in range(3):
    d = device_arm()
    e, m = e2e_arm()
    print(f"round {rnd}: device arm {d:.3f} ms/step ({pairs / d / 1e6:.0f} Gpairs/s) | e2e arm {e:.3f} ms/step, match kernels {m:.3f} ms of it", flush=True)
for pause in (0.5, 2.0, 5.0):
    time.sleep(pause)
    d = device_arm()
    time.sleep(pause)
    e, m = e2e_arm()
    print(f"after a {pause} s pause before each arm: device arm {d:.3f} ms/step | e2e arm {e:.3f} ms/step (match {m:.3f})", flush=True)
ctx.set_params(timing=0)
for rnd in range(2):
    e, m = e2e_arm()
    d = device_arm()
    print(f"timing auto, e2e first: e2e arm {e:.3f} ms/step (match {m:.3f}) | device arm {d:.3f} ms/step", flush=True)
