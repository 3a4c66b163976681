"""Experiment driver: how much of the FP4 kernel's clock is set by the operand DATA (power)?  One loop check of K
keyframes x 2000 rows with train rows that are random / all zero bits / all one bits, and a query that is random / zero.
usage: power_probe.py K [reps]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import slam_loop_closing_b200 as lc

K = int(sys.argv[1])
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
rng = np.random.default_rng(0)
for db_fill in ("random", "zeros", "ones", "quarter"):
    c = lc.Context(kernel=4, loop_threshold=0.15)
    if db_fill == "random":
        g = torch.Generator(device="cuda")
        g.manual_seed(1)
        rows = torch.randint(0, 256, (K * 2000, 32), dtype=torch.uint8, device="cuda", generator=g)
    elif db_fill == "quarter":  # a quarter of the bits set (AND of two random bytes)
        g = torch.Generator(device="cuda")
        g.manual_seed(1)
        rows = torch.randint(0, 256, (K * 2000, 32), dtype=torch.uint8, device="cuda", generator=g) & \
            torch.randint(0, 256, (K * 2000, 32), dtype=torch.uint8, device="cuda", generator=g)
    else:
        rows = torch.full((K * 2000, 32), 0 if db_fill == "zeros" else 255, dtype=torch.uint8, device="cuda")
    c.db_append_many_device(np.arange(K, dtype=np.int32), np.full(K, 2000, np.int32), rows.data_ptr(), torch.cuda.current_stream().cuda_stream or 1)
    c.synchronize()
    for q_fill in ("random", "zeros"):
        q = rng.integers(0, 256, (2000, 32), np.uint8) if q_fill == "random" else np.zeros((2000, 32), np.uint8)
        for _ in range(3):
            c.loop_check(q, K + 100, want_best=False)
        ks = []
        for _ in range(reps):
            c.loop_check(q, K + 100, want_best=False)
            ks.append(c.last_timing().match_ms)
        print(f"train {db_fill:7s} query {q_fill:7s} K {K}: match kernel {np.median(ks):8.3f} ms  {2000 * 2000 * K / np.median(ks) / 1e6:9.1f} Gpairs/s", flush=True)
    c.close()
    del rows
