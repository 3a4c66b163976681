"""ctypes binding of oracle/l2_oracle.c + the cv2 reference for the shipped pipeline's SIFT / L2 matcher
(TEST INFRASTRUCTURE ONLY - see oracle/__init__.py).  Reference: /root/reference/src/main.cpp:509-534, 1375-1392."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

L2_RESULT = np.dtype([("frame_id", "<i4"), ("n_kf", "<i4"), ("n_good", "<i4"), ("reserved", "<i4")])


def lib():
    global _LIB
    if _LIB is None:
        so, src = os.path.join(_HERE, "libl2_oracle.so"), os.path.join(_HERE, "l2_oracle.c")
        if not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
            subprocess.check_call(["make", "-C", _HERE, "-s", "libl2_oracle.so"])
        L = C.CDLL(so)
        L.l2o_knn2.restype = None
        L.l2o_knn2.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_int]
        L.l2o_ratio.restype = C.c_int
        L.l2o_ratio.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_void_p]
        _LIB = L
    return _LIB


def to_u8(desc):
    """OpenCV SIFT descriptors (float32, integer-valued 0..255) -> the u8 rows every implementation here works on."""
    d = np.asarray(desc)
    if d.dtype != np.uint8:
        r = np.round(d)
        if d.size and (np.any(r != d) or d.min() < 0 or d.max() > 255):
            raise ValueError("not OpenCV SIFT descriptors: values must be integers in [0, 255]")
        d = r.astype(np.uint8)
    return np.ascontiguousarray(d).reshape(-1, 128)


def knn2(q, t, nthreads=8):
    q, t = to_u8(q), to_u8(t)
    out = [np.zeros(len(q), np.int32) for _ in range(4)]
    lib().l2o_knn2(q.ctypes.data, len(q), t.ctypes.data, len(t), *[o.ctypes.data for o in out], nthreads)
    return dict(idx1=out[0], d1=out[1], idx2=out[2], d2=out[3])


def match_features(q, t, ratio=0.75, nthreads=8):
    """matchFeatures of src/main.cpp:509-534: knn + Lowe ratio; distances as float32 like cv::DMatch::distance."""
    k = knn2(q, t, nthreads)
    good = np.zeros(len(k["d1"]), np.uint8)
    n = lib().l2o_ratio(k["d1"].ctypes.data, k["d2"].ctypes.data, len(good), float(ratio), good.ctypes.data)
    return dict(train_idx=k["idx1"], dist2=k["d1"], second_dist2=k["d2"], good=good.astype(bool), n_good=int(n),
                distance=np.sqrt(np.maximum(k["d1"], 0).astype(np.float32)))


def loop_check(q, kfs, frame_ids, current, loop_gap, ratio=0.7, min_rows=100, nthreads=8):
    """the loop search's matching step (src/main.cpp:1378-1392): n_good per past view, -1 where the pair is skipped."""
    res = np.zeros(len(kfs), L2_RESULT)
    for k, (kf, fid) in enumerate(zip(kfs, frame_ids)):
        res[k] = (fid, len(kf), -1, 0)
        if current - fid < loop_gap or len(q) < min_rows or len(kf) < min_rows:
            continue
        res[k]["n_good"] = match_features(q, kf, ratio, nthreads)["n_good"]
    return res


# ---- the real OpenCV call (what src/main.cpp:517-519 runs)
def cv2_available():
    try:
        import cv2  # noqa: F401
        return True
    except Exception:
        return False


def cv2_match_features(q, t, ratio=0.75):
    import cv2
    q32, t32 = np.asarray(q, np.float32).reshape(-1, 128), np.asarray(t, np.float32).reshape(-1, 128)
    n = len(q32)
    out = dict(train_idx=np.full(n, -1, np.int32), distance=np.zeros(n, np.float32), second=np.full(n, -1, np.float32),
               good=np.zeros(n, bool))
    if n == 0 or len(t32) == 0:
        out["n_good"] = 0
        return out
    for i, m in enumerate(cv2.BFMatcher(cv2.NORM_L2, False).knnMatch(q32, t32, 2)):
        if len(m) >= 1:
            out["train_idx"][i], out["distance"][i] = m[0].trainIdx, m[0].distance
        if len(m) >= 2:
            out["second"][i] = m[1].distance
            out["good"][i] = m[0].distance < ratio * m[1].distance
    out["n_good"] = int(out["good"].sum())
    return out
