"""Latency of the consecutive-frame match (BASELINE config 2: 2000 x 2000) through the C ABI with host buffers.
python tools/pair_probe.py [nq] [nt] [kernel]"""
import sys
import time
import numpy as np
sys.path.insert(0, ".")
import slam_loop_closing_b200 as lc

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
kernel = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rng = np.random.default_rng(1)
q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
timing = int(sys.argv[4]) if len(sys.argv) > 4 else 0  # 1: event-time the pair call (costs ~10 us of the call)
with lc.Context(kernel=kernel, timing=timing) as c:
    for _ in range(20):
        c.match_pair(q, t)
    e2e, h2d, ker, d2h, mk = [], [], [], [], []
    for _ in range(200):
        t0 = time.perf_counter()
        c.match_pair(q, t)
        e2e.append(time.perf_counter() - t0)
        tm = c.last_timing()
        h2d.append(tm.h2d_ms); ker.append(tm.kernel_ms); d2h.append(tm.d2h_ms); mk.append(tm.match_ms)
    med = lambda v: float(np.median(v))
    print(f"pair {nq}x{nt} kernel {kernel} path {tm.path} launches {tm.kernel_launches}: e2e {med(e2e) * 1e6:.1f} us "
          f"(min {min(e2e) * 1e6:.1f}); device h2d {med(h2d) * 1e3:.1f} us, kernels {med(ker) * 1e3:.1f} us "
          f"(match kernel {med(mk) * 1e3:.1f} us), d2h {med(d2h) * 1e3:.1f} us; {nq * nt / med(e2e) / 1e9:.1f} Gpairs/s e2e")
