"""Experiment driver: SIFT / L2 loop check (4000-descriptor query vs K views x 4000 rows); kernel-only time."""
import sys
import numpy as np
sys.path.insert(0, ".")
import slam_loop_closing_b200 as lc

K = int(sys.argv[1]) if len(sys.argv) > 1 else 400
N = 4000
rng = np.random.default_rng(0)
c = lc.Context()
base = rng.integers(0, 200, (N, 128)).astype(np.uint8)
for k in range(K):
    c.l2_db_append(k, np.roll(base, k, axis=0))
q = rng.integers(0, 200, (N, 128)).astype(np.uint8)
for _ in range(2):
    c.l2_loop_check(q, K + 10, 1)
ks, ms = [], []
for _ in range(5):
    c.l2_loop_check(q, K + 10, 1)
    ks.append(c.last_timing().kernel_ms)
    ms.append(c.last_timing().match_ms)
print("l2 K", K, "kernel_ms", float(np.median(ks)), "match_ms", float(np.median(ms)), "Gpairs/s (128-d u8)", N * N * K / float(np.median(ms)) / 1e6, flush=True)
c.close()
