"""The reference's matcher as it really runs: OpenCV's BFMatcher(NORM_HAMMING) through cv2.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference's `matcher_` is a
cv::BFMatcher(NORM_HAMMING) (include/loop_closing.hpp:73, README.md:116); its matchFeatures /
detectLoops bodies (src/loop_closing.cpp) are absent, so README.md:117-126 is restated here in numpy
around the real OpenCV call.  Used to pin oracle/lc_oracle.c, to generate tests/golden/, as a second
checker in the GPU parity tests and as the reported CPU baseline in bench.py.
"""
import numpy as np

from .params import Params

try:
    import cv2
except Exception as e:  # pragma: no cover - cv2 is part of the image
    cv2 = None
    _cv2_err = e

KF_RESULT = np.dtype([("frame_id", "<i4"), ("n_kf", "<i4"), ("min_dist", "<i4"), ("n_good", "<i4")])


def available():
    return cv2 is not None


def _bf():
    if cv2 is None:
        raise RuntimeError(f"cv2 unavailable: {_cv2_err}")
    return cv2.BFMatcher(cv2.NORM_HAMMING)  # crossCheck=False (default), Appendix B8


def match(q, t):
    """cv::BFMatcher::match(q, t) -> (train_idx[nq], dist[nq]) with OpenCV's own tie-break."""
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
    t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
    if len(q) == 0 or len(t) == 0:
        return np.zeros(0, np.int32), np.zeros(0, np.int32)
    ms = _bf().match(q, t)
    assert len(ms) == len(q) and all(m.queryIdx == i for i, m in enumerate(ms))
    idx = np.fromiter((m.trainIdx for m in ms), np.int32, len(ms))
    dist = np.fromiter((int(m.distance) for m in ms), np.int32, len(ms))
    assert all(float(d) == m.distance for d, m in zip(dist, ms))
    return idx, dist


def filter_matches(dist, p: Params = Params()):
    dist = np.asarray(dist, np.int64)
    if len(dist) == 0:
        return np.zeros(0, bool), -1, 0
    mn = int(dist.min())
    if p.min_init >= 0:
        mn = min(mn, p.min_init)
    thr = max(2 * mn, p.filter_floor)
    good = dist < thr if p.filter_strict else dist <= thr
    return good, mn, int(good.sum())


def match_features(q, t, p: Params = Params()):
    idx, dist = match(q, t)
    good, mn, ng = filter_matches(dist, p)
    return dict(train_idx=idx, dist=dist, good=good, min_dist=mn, n_good=ng)


def loop_check(q, keyframes, frame_ids, current_frame_id, p: Params = Params(), want_best=True):
    """keyframes: list of (n_k,32) uint8 arrays in insertion order.  One match() per eligible keyframe,
    exactly the reference's access pattern."""
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
    K = len(keyframes)
    res = np.zeros(K, KF_RESULT)
    best = np.full(len(q), np.iinfo(np.uint64).max, np.uint64)
    off = 0
    for k, (kf, fid) in enumerate(zip(keyframes, frame_ids)):
        kf = np.ascontiguousarray(kf, np.uint8).reshape(-1, 32)
        res[k] = (fid, len(kf), -1, -1)
        gap = current_frame_id - fid
        eligible = gap > p.min_loop_gap if p.gap_strict else gap >= p.min_loop_gap
        if eligible and len(q) and len(kf):
            idx, dist = match(q, kf)
            _, mn, ng = filter_matches(dist, p)
            res[k]["min_dist"], res[k]["n_good"] = mn, ng
            if want_best:
                pk = (dist.astype(np.uint64) << np.uint64(32)) | (idx.astype(np.uint64) + np.uint64(off))
                best = np.minimum(best, pk)
        off += len(kf)
    return res, (best if want_best else None)


def global_best_multi_image(q, keyframes):
    """OpenCV's own multi-image tie-break (bf.add([...]); bf.match(q)): lowest imgIdx then trainIdx.
    Returns packed (dist<<32 | global_row) per query; pins Appendix A step 6."""
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
    kfs = [np.ascontiguousarray(k, np.uint8).reshape(-1, 32) for k in keyframes]
    offs = np.concatenate([[0], np.cumsum([len(k) for k in kfs])]).astype(np.int64)
    bf = _bf()
    bf.add([k for k in kfs if len(k)])
    nz = [i for i, k in enumerate(kfs) if len(k)]
    ms = bf.match(q)
    out = np.empty(len(q), np.uint64)
    for i, m in enumerate(ms):
        out[i] = (int(m.distance) << 32) | int(offs[nz[m.imgIdx]] + m.trainIdx)
    return out
