"""Generates tests/golden/*.npz with the REAL OpenCV matcher (cv2.BFMatcher(NORM_HAMMING)).

The reference ships no tests or golden vectors for this path and its own matchFeatures/detectLoops
source is absent, so the fixtures are produced by the OpenCV call the reference's
`matcher_` makes (include/loop_closing.hpp:73) plus README.md:117-126's rules (oracle/cv2_ref.py).
Run from the repo root:  python tests/golden/make_golden.py
Inputs are stored next to the outputs so the fixtures do not depend on numpy's RNG streams.
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import datagen  # noqa: E402
from oracle import cv2_ref  # noqa: E402
from oracle.params import Params  # noqa: E402


def textured_canvas(h, w, seed):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w), np.uint8)
    img = cv2.GaussianBlur(img, (0, 0), 2.0)
    return cv2.normalize(img, None, 0, 255, cv2.NORM_MINMAX)


def orb_loop_sequence(n_frames=24, nfeatures=300, seed=3):
    """Camera window sliding around a closed path on a textured canvas (frames 0 and n-1 overlap)."""
    canvas = textured_canvas(900, 1200, seed)
    orb = cv2.ORB_create(nfeatures)
    descs = []
    for i in range(n_frames):
        a = 2 * np.pi * i / (n_frames - 1)
        x, y = int(430 + 300 * np.cos(a)), int(330 + 200 * np.sin(a))
        _, d = orb.detectAndCompute(np.ascontiguousarray(canvas[y:y + 240, x:x + 320]), None)
        descs.append(d if d is not None else np.zeros((0, 32), np.uint8))
    return descs


def main():
    out = {}
    # ---- pair cases (match + filter), several knob settings
    pairs = {
        "uniform_300x700": (datagen.uniform(300, 1), datagen.uniform(700, 2)),
        "ties_300x700": (datagen.low_entropy(300, 3), datagen.low_entropy(700, 4)),
        "planted_200x260": datagen.planted_pair(200, 260, 5),
        "edge_1x1": (datagen.uniform(1, 6), datagen.uniform(1, 7)),
        "edge_31x33": (datagen.low_entropy(31, 8), datagen.low_entropy(33, 9)),
        "edge_33x31": (datagen.low_entropy(33, 10), datagen.low_entropy(31, 11)),
        "identical_rows": (np.repeat(datagen.uniform(1, 12), 40, 0), np.repeat(datagen.uniform(1, 12), 70, 0)),
        "empty_train": (datagen.uniform(5, 13), np.zeros((0, 32), np.uint8)),
    }
    knobs = {"default": Params(), "floor30": Params(filter_floor=30), "strict": Params(filter_strict=True),
             "init100": Params(min_init=100)}
    for name, (q, t) in pairs.items():
        out[f"pair/{name}/q"], out[f"pair/{name}/t"] = q, t
        idx, dist = cv2_ref.match(q, t)
        out[f"pair/{name}/train_idx"], out[f"pair/{name}/dist"] = idx, dist
        for kn, p in knobs.items():
            good, mn, ng = cv2_ref.filter_matches(dist, p)
            out[f"pair/{name}/{kn}/good"] = np.asarray(good, bool)
            out[f"pair/{name}/{kn}/summary"] = np.array([mn, ng], np.int32)
    # ---- loop check on real ORB descriptors of a synthetic closed-loop sequence
    descs = orb_loop_sequence()
    p = Params(loop_threshold=0.15, min_loop_gap=8, min_good=20, filter_floor=30)
    out["loop/orb/n"] = np.array([len(d) for d in descs], np.int32)
    out["loop/orb/rows"] = np.concatenate(descs)
    out["loop/orb/params"] = np.array([p.loop_threshold, p.min_loop_gap, p.min_good, p.filter_floor])
    recs, bests, cands = [], [], []
    for i, d in enumerate(descs):  # processFrame order: append frame i, then detectLoops(i)
        res, best = cv2_ref.loop_check(d, descs[:i + 1], list(range(i + 1)), i, p)
        recs.append(res)
        bests.append(best)
        for r in res:
            if r["n_good"] >= 0 and min(len(d), r["n_kf"]) > 0:
                sim = float(int(r["n_good"])) / float(min(len(d), int(r["n_kf"])))
                if sim > p.loop_threshold and r["n_good"] >= p.min_good:
                    cands.append((i, int(r["frame_id"]), int(r["n_good"]), sim))
    out["loop/orb/last_results"] = recs[-1]
    out["loop/orb/last_best"] = bests[-1]
    out["loop/orb/all_n_good"] = np.concatenate([r["n_good"] for r in recs])
    out["loop/orb/all_min_dist"] = np.concatenate([r["min_dist"] for r in recs])
    out["loop/orb/candidates"] = np.array(cands, np.float64).reshape(-1, 4)
    # ---- multi-image tie-break (OpenCV's own add([...]).match)
    q = datagen.low_entropy(120, 20)
    kfs = [datagen.low_entropy(n, 30 + i) for i, n in enumerate((40, 0, 75, 1, 64))]
    out["multi/q"] = q
    out["multi/n"] = np.array([len(k) for k in kfs], np.int32)
    out["multi/rows"] = np.concatenate(kfs)
    out["multi/best"] = cv2_ref.global_best_multi_image(q, kfs)
    np.savez_compressed(os.path.join(HERE, "hamming_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "hamming_golden.npz"), "cv2", cv2.__version__, "loop candidates:", len(cands))


if __name__ == "__main__":
    main()
