"""Latency of lc_loop_check (host buffers) against small keyframe stores - the regime of the reference's own run
(BASELINE config 1: <= ~70 eligible keyframes) up to C3 (1000).  python tools/loop_latency.py [kernel]"""
import sys
import time
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import datagen
import slam_loop_closing_b200 as lc

kernel = int(sys.argv[1]) if len(sys.argv) > 1 else 0
timing = int(sys.argv[2]) if len(sys.argv) > 2 else 0  # 1: event-time small checks too (costs ~20 us per call)
q = datagen.uniform(2000, 1234)
for K in (1, 4, 10, 30, 67, 100, 200, 400, 1000):
    with lc.Context(kernel=kernel, loop_threshold=0.15, min_loop_gap=0, timing=timing) as c:
        kfs = [datagen.uniform(2000, 5678 + k) for k in range(K)]
        c.db_append_many(list(range(K)), kfs)
        for _ in range(10):
            c.loop_check(q, K + 30)
        ts, ks, ms = [], [], []
        for _ in range(60):
            t0 = time.perf_counter()
            c.loop_check(q, K + 30)
            ts.append(time.perf_counter() - t0)
            tm = c.last_timing()
            ks.append(tm.kernel_ms); ms.append(tm.match_ms)
        med = float(np.median(ts))
        print(f"K {K:5d}: e2e {med * 1e6:8.1f} us  kernels {np.median(ks) * 1e3:8.1f} us  match {np.median(ms) * 1e3:8.1f} us  path {tm.path} "
              f"launches {tm.kernel_launches}  {K * 4e6 / med / 1e9:9.1f} Gpairs/s e2e")
