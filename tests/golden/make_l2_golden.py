"""Generates tests/golden/l2_golden.npz with the REAL OpenCV call of the shipped pipeline's matcher:
cv2.SIFT descriptors + cv2.BFMatcher(NORM_L2).knnMatch(k=2) + Lowe's ratio test
(/root/reference/src/main.cpp:497-534, loop search :1375-1392).  Run from the repo root:
    python tests/golden/make_l2_golden.py
Inputs (u8 descriptors) are stored next to the outputs.
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import l2_oracle  # noqa: E402


def textured_canvas(h, w, seed):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w), np.uint8)
    img = cv2.GaussianBlur(img, (0, 0), 2.0)
    return cv2.normalize(img, None, 0, 255, cv2.NORM_MINMAX)


def sift_loop_sequence(n_views=10, nfeatures=400, seed=5):
    """Camera window sliding around a closed path on a textured canvas (views 0 and n-1 overlap)."""
    canvas = textured_canvas(900, 1200, seed)
    sift = cv2.SIFT_create(nfeatures)
    descs = []
    for i in range(n_views):
        a = 2 * np.pi * i / (n_views - 1)
        x, y = int(430 + 300 * np.cos(a)), int(330 + 200 * np.sin(a))
        _, d = sift.detectAndCompute(np.ascontiguousarray(canvas[y:y + 300, x:x + 400]), None)
        descs.append(l2_oracle.to_u8(d if d is not None else np.zeros((0, 128), np.float32)))
    return descs


def main():
    out = {}
    rng = np.random.default_rng(11)
    q = rng.integers(0, 160, (300, 128)).astype(np.uint8)
    t = rng.integers(0, 160, (450, 128)).astype(np.uint8)
    t[40], t[41], t[200] = q[7], q[7], q[9]          # exact copies: distance 0, ties resolved by train index
    t[300:310] = np.clip(q[20:30].astype(np.int32) + rng.integers(-3, 4, (10, 128)), 0, 255).astype(np.uint8)
    lowent = rng.integers(0, 2, (260, 128)).astype(np.uint8) * 50  # many equal distances
    pairs = {"random_300x450": (q, t), "ties_120x140": (lowent[:120], lowent[120:]),
             "one_train_row": (q[:5], t[:1]), "two_train_rows": (q[:5], t[:2]), "empty_train": (q[:5], t[:0])}
    for name, (a, b) in pairs.items():
        for ratio in (0.75, 0.7):
            r = l2_oracle.cv2_match_features(a, b, ratio)
            out[f"pair/{name}/q"], out[f"pair/{name}/t"] = a, b
            out[f"pair/{name}/train_idx"], out[f"pair/{name}/distance"], out[f"pair/{name}/second"] = \
                r["train_idx"], r["distance"], r["second"]
            out[f"pair/{name}/good_{ratio}"] = r["good"]
    # ---- real SIFT loop sequence: every (curr, past <= curr - gap) pair as the loop search visits them (ratio 0.7)
    descs = sift_loop_sequence()
    gap, min_rows = 3, 100
    out["loop/n"] = np.array([len(d) for d in descs], np.int32)
    out["loop/rows"] = np.concatenate(descs)
    out["loop/params"] = np.array([gap, min_rows, 0.7])
    ng = np.full((len(descs), len(descs)), -1, np.int32)
    for curr in range(gap, len(descs)):
        for past in range(0, curr - gap + 1):
            if len(descs[curr]) < min_rows or len(descs[past]) < min_rows:
                continue
            ng[curr, past] = l2_oracle.cv2_match_features(descs[curr], descs[past], 0.7)["n_good"]
    out["loop/n_good"] = ng
    r = l2_oracle.cv2_match_features(descs[-1], descs[0], 0.7)
    out["loop/last_vs_first/train_idx"], out["loop/last_vs_first/distance"], out["loop/last_vs_first/good"] = \
        r["train_idx"], r["distance"], r["good"]
    path = os.path.join(HERE, "l2_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", "rows per view", out["loop/n"].tolist(),
          "max n_good", int(ng.max()), "last vs first", int(ng[-1, 0]))


if __name__ == "__main__":
    main()
