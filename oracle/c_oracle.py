"""ctypes binding of oracle/lc_oracle.c (TEST INFRASTRUCTURE ONLY - see oracle/__init__.py)."""
import ctypes as C
import os
import subprocess

import numpy as np

from .params import Params

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _CParams(C.Structure):
    _fields_ = [("loop_threshold", C.c_double), ("min_loop_gap", C.c_int32), ("min_good", C.c_int32),
                ("filter_floor", C.c_int32), ("filter_strict", C.c_int32), ("min_init", C.c_int32),
                ("gap_strict", C.c_int32), ("thr_inclusive", C.c_int32)]


KF_RESULT = np.dtype([("frame_id", "<i4"), ("n_kf", "<i4"), ("min_dist", "<i4"), ("n_good", "<i4")])


def build():
    so = os.path.join(_HERE, "liblc_oracle.so")
    src = os.path.join(_HERE, "lc_oracle.c")
    if not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        u8p, i32p, i64p, u64p = (C.POINTER(t) for t in (C.c_uint8, C.c_int32, C.c_int64, C.c_uint64))
        L.lco_match.restype = C.c_int
        L.lco_match.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_size_t,
                                C.c_void_p, C.c_void_p, C.c_int]
        L.lco_filter.restype = None
        L.lco_filter.argtypes = [C.c_void_p, C.c_int, C.POINTER(_CParams), C.c_void_p,
                                 C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.lco_loop_check.restype = None
        L.lco_loop_check.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_int, C.c_int, C.POINTER(_CParams), C.c_void_p,
                                     C.c_void_p, C.c_int]
        _LIB = L
    return _LIB


def _cp(p: Params) -> _CParams:
    return _CParams(p.loop_threshold, p.min_loop_gap, p.min_good, p.filter_floor, int(p.filter_strict),
                    p.min_init, int(p.gap_strict), int(p.thr_inclusive))


def _rows(a):
    a = np.ascontiguousarray(a, dtype=np.uint8).reshape(-1, 32)
    return a


def match(q, t, nthreads=1):
    """BFMatcher(NORM_HAMMING).match: returns (train_idx[nq], dist[nq]); empty arrays if a side is empty."""
    q, t = _rows(q), _rows(t)
    nq, nt = len(q), len(t)
    if nq == 0 or nt == 0:
        return np.zeros(0, np.int32), np.zeros(0, np.int32)
    idx = np.empty(nq, np.int32)
    dist = np.empty(nq, np.int32)
    n = lib().lco_match(q.ctypes.data, nq, 32, t.ctypes.data, nt, 32, idx.ctypes.data, dist.ctypes.data, nthreads)
    assert n == nq
    return idx, dist


def filter_matches(dist, p: Params = Params()):
    """2x-min-distance filter: returns (good[n] bool, min_dist, n_good)."""
    dist = np.ascontiguousarray(dist, np.int32)
    good = np.zeros(len(dist), np.uint8)
    mn, ng = C.c_int32(), C.c_int32()
    cp = _cp(p)
    lib().lco_filter(dist.ctypes.data, len(dist), C.byref(cp), good.ctypes.data, C.byref(mn), C.byref(ng))
    return good.astype(bool), mn.value, ng.value


def match_features(q, t, p: Params = Params(), nthreads=1):
    """matchFeatures(frame1=q, frame2=t): dict(train_idx, dist, good, min_dist, n_good)."""
    idx, dist = match(q, t, nthreads)
    good, mn, ng = filter_matches(dist, p)
    return dict(train_idx=idx, dist=dist, good=good, min_dist=mn, n_good=ng)


def loop_check(q, db_rows, n_kf, frame_ids, current_frame_id, p: Params = Params(), nthreads=1, want_best=True):
    """Per-keyframe (frame_id, n_kf, min_dist, n_good) records + packed per-query global best."""
    q = _rows(q)
    db = _rows(db_rows)
    n_kf = np.ascontiguousarray(n_kf, np.int32)
    frame_ids = np.ascontiguousarray(frame_ids, np.int32)
    K = len(n_kf)
    row_off = np.zeros(K, np.int64)
    if K:
        row_off[1:] = np.cumsum(n_kf[:-1], dtype=np.int64)
    assert int(n_kf.sum()) == len(db)
    res = np.zeros(K, KF_RESULT)
    best = np.full(len(q), np.iinfo(np.uint64).max, np.uint64) if want_best else None
    cp = _cp(p)
    lib().lco_loop_check(q.ctypes.data, len(q), 32, db.ctypes.data, row_off.ctypes.data, n_kf.ctypes.data,
                         frame_ids.ctypes.data, K, int(current_frame_id), C.byref(cp), res.ctypes.data,
                         best.ctypes.data if want_best else None, nthreads)
    return res, best


def loop_candidates(res, nq, current_frame_id, p: Params = Params()):
    """Apply README.md:121-126 to per-keyframe records: list of (cur, matched, n_good, similarity)."""
    out = []
    for r in res:
        if r["n_good"] < 0:
            continue
        denom = min(int(nq), int(r["n_kf"]))
        if denom == 0:
            continue
        sim = float(int(r["n_good"])) / float(denom)
        ok_thr = sim >= p.loop_threshold if p.thr_inclusive else sim > p.loop_threshold
        if ok_thr and int(r["n_good"]) >= p.min_good:
            out.append((int(current_frame_id), int(r["frame_id"]), int(r["n_good"]), sim))
    return out
