"""Experiment driver: kernel-only time of one loop check (K keyframes x 2000 rows) for a given kernel id."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import slam_loop_closing_b200 as lc

kernel, K = int(sys.argv[1]), int(sys.argv[2])
c = lc.Context(kernel=kernel, loop_threshold=0.15)
g = torch.Generator(device="cuda")
g.manual_seed(1)
rows = torch.randint(0, 256, (K * 2000, 32), dtype=torch.uint8, device="cuda", generator=g)
c.db_append_many_device(np.arange(K, dtype=np.int32), np.full(K, 2000, np.int32), rows.data_ptr(), torch.cuda.current_stream().cuda_stream or 1)
c.synchronize()
q = np.random.default_rng(0).integers(0, 256, (2000, 32), np.uint8)
for _ in range(3):
    c.loop_check(q, K + 100)
ks = []
for _ in range(5):
    c.loop_check(q, K + 100)
    ks.append(c.last_timing().kernel_ms)
print("kernel", kernel, "K", K, "kernel_ms", float(np.median(ks)), "Gpairs/s", 2000 * 2000 * K / float(np.median(ks)) / 1e6, flush=True)
c.close()
