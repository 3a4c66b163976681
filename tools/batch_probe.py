"""Experiment driver: batched loop check (F query frames x 2000 rows) against K keyframes; kernel-only time."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import slam_loop_closing_b200 as lc

kernel, K, F = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
c = lc.Context(kernel=kernel, loop_threshold=0.15)
g = torch.Generator(device="cuda")
g.manual_seed(1)
rows = torch.randint(0, 256, (K * 2000, 32), dtype=torch.uint8, device="cuda", generator=g)
c.db_append_many_device(np.arange(K, dtype=np.int32), np.full(K, 2000, np.int32), rows.data_ptr(), torch.cuda.current_stream().cuda_stream or 1)
c.synchronize()
rng = np.random.default_rng(0)
qs = [rng.integers(0, 256, (2000, 32), np.uint8) for _ in range(F)]
cur = [K + 100] * F
import time
for i in range(2):
    t0 = time.perf_counter()
    res, best = c.loop_check_batch(qs, cur)
    tm = c.last_timing()
    print(f"warm-up call {i}: wall {(time.perf_counter() - t0) * 1e3:.1f} ms; h2d {tm.h2d_ms:.2f} kernels {tm.kernel_ms:.1f} (match {tm.match_ms:.1f} in "
          f"{tm.match_launches}) d2h {tm.d2h_ms:.2f} total {tm.total_ms:.1f}", flush=True)
ks = []
for _ in range(3):
    t0 = time.perf_counter()
    c.loop_check_batch(qs, cur)
    tm = c.last_timing()
    ks.append(tm.kernel_ms)
    print(f"call: wall {(time.perf_counter() - t0) * 1e3:.1f} ms; h2d {tm.h2d_ms:.2f} kernels {tm.kernel_ms:.1f} (match {tm.match_ms:.1f} in "
          f"{tm.match_launches}) d2h {tm.d2h_ms:.2f} total {tm.total_ms:.1f}", flush=True)
r1, b1 = c.loop_check(qs[F // 2], K + 100)
assert np.array_equal(res[F // 2], r1) and np.array_equal(best[F // 2], b1), "batch != single"
print("kernel", kernel, "K", K, "F", F, "kernel_ms", float(np.median(ks)), "Gpairs/s", 2000 * 2000 * K * F / float(np.median(ks)) / 1e6,
      "launches", c.last_timing().kernel_launches, flush=True)
c.close()
