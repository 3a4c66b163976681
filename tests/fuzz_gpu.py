"""Randomised parity soak: random shapes / kernels / knobs through the C ABI against the CPU oracle, for a time budget.
   python tests/fuzz_gpu.py [seconds] [seed]      (tests/test_gpu_fuzz.py runs a short one in the GPU suite)"""
import os
import sys
import time
import numpy as np
_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
sys.path.insert(0, _HERE)
import datagen
import slam_loop_closing_b200 as lc
from oracle import c_oracle, l2_oracle
from oracle.params import Params



def soak(budget, seed):
    rng = np.random.default_rng(seed)
    t0, cases = time.time(), 0
    while time.time() - t0 < budget:
        cases += _one_case(rng, cases)
    return cases, time.time() - t0


def _one_case(rng, cases):
    if True:
        kernel = int(rng.choice([0, 2, 3, 4]))
        n_kf = int(rng.integers(1, 500))
        max_rows = int(rng.choice([40, 250, 700, 2300]))
        p = Params(loop_threshold=0.15, min_loop_gap=int(rng.integers(0, 40)), filter_floor=int(rng.choice([0, 30])),
                   filter_strict=bool(rng.integers(0, 2)), gap_strict=bool(rng.integers(0, 2)))
        gen = datagen.low_entropy if rng.integers(0, 2) else datagen.uniform
        kfs = [gen(int(rng.integers(0, max_rows)), int(rng.integers(1 << 30))) for _ in range(n_kf)]
        fids = np.sort(rng.choice(np.arange(n_kf * 2), n_kf, replace=False)).tolist()
        F = int(rng.choice([1, 1, 2, 5]))
        qs = [gen(int(rng.integers(0, 2600)), int(rng.integers(1 << 30))) for _ in range(F)]
        if len(qs[0]) and n_kf > 3 and len(kfs[2]):  # plant a noisy copy so that n_good is not trivial
            kfs[2] = datagen.noisy_copy(qs[0], 5, p=0.05, n=max(len(kfs[2]), 1))
        curs = [int(rng.integers(0, n_kf * 2 + 50)) for _ in range(F)]
        with lc.Context(kernel=kernel, loop_threshold=p.loop_threshold, min_loop_gap=p.min_loop_gap, filter_floor=p.filter_floor,
                        filter_strict=p.filter_strict, gap_strict=p.gap_strict) as c:
            c.db_append_many(fids, kfs)
            res, best = c.loop_check_batch(qs, curs)
            path = c.last_timing().path
        db = np.concatenate(kfs) if n_kf else np.zeros((0, 32), np.uint8)
        for f in range(F):
            w_res, w_best = c_oracle.loop_check(qs[f], db, [len(k) for k in kfs], fids, curs[f], p, nthreads=16)
            assert np.array_equal(res[f], w_res), (cases, kernel, path, "records")
            assert np.array_equal(best[f], w_best), (cases, kernel, path, "best")
        # matchFeatures (the single-launch pair kernel): several differently sized pairs through ONE context, so that the
        # kernel's self-cleaning key buffer and ticket counter are exercised across calls; strided rows now and then
        pp = Params(filter_floor=int(rng.choice([0, 30])), filter_strict=bool(rng.integers(0, 2)), min_init=int(rng.choice([-1, 100])))
        with lc.Context(kernel=int(rng.choice([0, 1, 2])), filter_floor=pp.filter_floor, filter_strict=pp.filter_strict, min_init=pp.min_init) as c:
            for _ in range(4):
                hi = int(rng.choice([40, 600, 2600, 7000]))
                a = gen(int(rng.integers(1, hi)), int(rng.integers(1 << 30)))
                b = gen(int(rng.integers(1, hi)), int(rng.integers(1 << 30)))
                if rng.integers(0, 3) == 0:  # a planted near-copy: small min_dist, selective filter
                    b = datagen.noisy_copy(a, int(rng.integers(1 << 30)), p=0.05, n=len(b))
                if rng.integers(0, 4) == 0:  # non-contiguous rows (a cv::Mat ROI): row stride 48
                    wide = np.zeros((len(a), 48), np.uint8)
                    wide[:, :32] = a
                    a = wide[:, :32]
                g = c.match_pair(a, b)
                w = c_oracle.match_features(a, b, pp, nthreads=16)
                assert c.last_timing().path == 5 and c.last_timing().kernel_launches == 1
                assert np.array_equal(g["train_idx"], w["train_idx"]) and np.array_equal(g["dist"], w["dist"]) and \
                    np.array_equal(g["good"].astype(bool), w["good"]) and (g["min_dist"], g["n_good"]) == (w["min_dist"], w["n_good"]), \
                    (cases, "pair", a.shape, b.shape)
        if cases % 3 == 0:  # lc_process_frame: a short random frame sequence, fused call == the oracle frame by frame
            pf = Params(loop_threshold=float(rng.choice([0.05, 0.15, 0.5])), min_loop_gap=int(rng.integers(0, 6)), filter_floor=int(rng.choice([0, 30])),
                        gap_strict=bool(rng.integers(0, 2)), thr_inclusive=bool(rng.integers(0, 2)))
            n_fr, hi = int(rng.integers(2, 14)), int(rng.choice([60, 900, 2300]))
            seq = []
            for i in range(n_fr):
                n = int(rng.integers(0, hi))
                full = [x for x in seq if len(x)]
                if full and n and rng.integers(0, 2):  # a revisit: a noisy copy of an earlier frame
                    seq.append(datagen.noisy_copy(full[int(rng.integers(0, len(full)))], int(rng.integers(1 << 30)), p=0.05, n=n))
                else:
                    seq.append(gen(n, int(rng.integers(1 << 30))))
            with lc.Context(kernel=int(rng.choice([0, 2, 4])), loop_threshold=pf.loop_threshold, min_loop_gap=pf.min_loop_gap, filter_floor=pf.filter_floor,
                            gap_strict=pf.gap_strict, thr_inclusive=pf.thr_inclusive) as c:
                for i, d in enumerate(seq):
                    if rng.integers(0, 2):
                        pair, cands, k = c.process_frame(3 * i, d)
                    else:
                        k = c.process_frame_begin(3 * i, d)
                        pair, cands = c.process_frame_matches(), c.process_frame_end()
                    assert k == i
                    if i:
                        w = c_oracle.match_features(seq[i - 1], d, pf, nthreads=16)
                        assert np.array_equal(pair["train_idx"], w["train_idx"]) and np.array_equal(pair["dist"], w["dist"]) and \
                            np.array_equal(pair["good"].astype(bool), w["good"]) and (pair["min_dist"], pair["n_good"]) == (w["min_dist"], w["n_good"]), \
                            (cases, "frame pair", i)
                    w_res, _ = c_oracle.loop_check(d, np.concatenate(seq[:i + 1]), [len(x) for x in seq[:i + 1]], [3 * j for j in range(i + 1)], 3 * i, pf,
                                                   nthreads=16)
                    want = c_oracle.loop_candidates(w_res, len(d), 3 * i, pf)
                    got = [(int(x["current_frame_id"]), int(x["matched_frame_id"]), int(x["num_matches"]), float(x["similarity_score"])) for x in cands]
                    assert got == [tuple(x) for x in want], (cases, "frame loops", i)
        if cases % 7 == 0:  # the SIFT / L2 path
            a = rng.integers(0, 256, (int(rng.integers(1, 1500)), 128)).astype(np.uint8)
            b = rng.integers(0, 256, (int(rng.integers(1, 1500)), 128)).astype(np.uint8)
            if rng.integers(0, 2):
                a, b = (a // 64) * 40, (b // 64) * 40  # few distinct values: many equal distances
            with lc.Context() as c:
                g = c.l2_match_pair(a, b, 0.8)
            w = l2_oracle.match_features(a, b, 0.8, nthreads=16)
            assert np.array_equal(g["train_idx"], w["train_idx"]) and np.array_equal(g["distance"], w["distance"]) and \
                np.array_equal(g["good"], w["good"]), (cases, "l2")
    return 1


if __name__ == "__main__":
    n, dt = soak(float(sys.argv[1]) if len(sys.argv) > 1 else 120.0, int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    print(f"fuzz ok: {n} random cases in {dt:.0f} s, all bit-exact")
