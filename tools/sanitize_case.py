"""Small end-to-end case for compute-sanitizer: every kernel family once (FP4 / int8 tensor, integer scan + split, L2)."""
import sys
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import datagen
import slam_loop_closing_b200 as lc

q = datagen.uniform(1100, 1)
kfs = datagen.structured_db(q, 170, n=300, every=7, ragged=True)
for kernel in (4, 3, 2):
    with lc.Context(kernel=kernel, loop_threshold=0.15) as c:
        c.db_append_many(list(range(170)), kfs)
        res, best = c.loop_check(q, 300)
        print("kernel", kernel, "path", c.last_timing().path, int(res["n_good"].sum()), int(best.sum() % 1000003))
        r = c.match_pair(q, kfs[3])
        m = c.match_keyframes(q, [0, 7, 14])
rng = np.random.default_rng(0)
with lc.Context() as c:
    a, b = rng.integers(0, 200, (300, 128)).astype(np.uint8), rng.integers(0, 200, (500, 128)).astype(np.uint8)
    print("l2", c.l2_match_pair(a, b)["n_good"])
    for k in range(5):
        c.l2_db_append(k, b)
    print(c.l2_loop_check(a, 10, 1)["n_good"])
# the fused per-frame call (pair match vs the resident previous frame + append + loop check in one submission), both stores sizes
with lc.Context(loop_threshold=0.15, min_loop_gap=3) as c:
    tot = 0
    for fid in range(14):
        pair, cand, k = c.process_frame(fid, kfs[fid % 9])
        tot += len(cand)
    print("process_frame", tot)
print("done")
