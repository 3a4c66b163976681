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
for _ in range(2):
    res, best = c.loop_check_batch(qs, cur)
ks = []
for _ in range(3):
    c.loop_check_batch(qs, cur)
    ks.append(c.last_timing().kernel_ms)
r1, b1 = c.loop_check(qs[F // 2], K + 100)
assert np.array_equal(res[F // 2], r1) and np.array_equal(best[F // 2], b1), "batch != single"
print("kernel", kernel, "K", K, "F", F, "kernel_ms", float(np.median(ks)), "Gpairs/s", 2000 * 2000 * K * F / float(np.median(ks)) / 1e6,
      "launches", c.last_timing().kernel_launches, flush=True)
c.close()
