"""ctypes binding of liblc_b200.so - the C ABI of include/lc_abi.h.

There is no Python or CPU implementation behind this module: if the shared library is missing or
cannot be loaded, importing the product fails loudly.
"""
import ctypes as C
import os

import numpy as np

from . import _build

DESC_BYTES = 32
UINT64_MAX = np.iinfo(np.uint64).max

KF_RESULT = np.dtype([("frame_id", "<i4"), ("n_kf", "<i4"), ("min_dist", "<i4"), ("n_good", "<i4")])
LOOP_CANDIDATE = np.dtype([("current_frame_id", "<i4"), ("matched_frame_id", "<i4"), ("num_matches", "<i4"),
                           ("similarity_score", "<f8")], align=True)

KERNEL_AUTO, KERNEL_POPC, KERNEL_CSA, KERNEL_TENSOR, KERNEL_MXF4 = 0, 1, 2, 3, 4


class Config(C.Structure):
    """lc_config (include/lc_abi.h)."""
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("loop_threshold", C.c_double),
                ("min_loop_gap", C.c_int32), ("min_good", C.c_int32), ("filter_floor", C.c_int32),
                ("filter_strict", C.c_int32), ("min_init", C.c_int32), ("gap_strict", C.c_int32),
                ("thr_inclusive", C.c_int32), ("shard_rank", C.c_int32), ("shard_world", C.c_int32),
                ("kernel", C.c_int32), ("timing", C.c_int32), ("slab_rows", C.c_int64)]


L2_RESULT = np.dtype([("frame_id", "<i4"), ("n_kf", "<i4"), ("n_good", "<i4"), ("reserved", "<i4")])  # lc_l2_result
MATCH = np.dtype([("query_idx", np.int32), ("train_idx", np.int32), ("distance", np.int32)])  # lc_match


class Timing(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("kernel_ms", C.c_float), ("d2h_ms", C.c_float), ("total_ms", C.c_float),
                ("kernel_launches", C.c_int32), ("path", C.c_int32), ("match_ms", C.c_float), ("match_launches", C.c_int32)]


class LcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"liblc_b200 error {code}: {msg}")
        self.code = code


# every symbol include/lc_abi.h declares: (restype, argtypes)
_vp, _i32, _i64, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
_pi32, _pi64 = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
SYMBOLS = {
    "lc_abi_version": (_i32, []),
    "lc_default_config": (None, [C.POINTER(Config)]),
    "lc_create": (_i32, [C.POINTER(Config), C.POINTER(_vp)]),
    "lc_destroy": (None, [_vp]),
    "lc_last_error": (C.c_char_p, [_vp]),
    "lc_get_config": (_i32, [_vp, C.POINTER(Config)]),
    "lc_set_params": (_i32, [_vp, C.POINTER(Config)]),
    "lc_db_append": (_i32, [_vp, _i32, _vp, _i32, _sz, _pi64]),
    "lc_db_append_many": (_i32, [_vp, _i32, _vp, _vp, _vp]),
    "lc_db_append_many_device": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp]),
    "lc_db_size": (_i32, [_vp, _pi64, _pi64, _pi64, _pi64]),
    "lc_db_clear": (_i32, [_vp]),
    "lc_db_read": (_i32, [_vp, _i64, _vp, _i32, _pi32, _pi32]),
    "lc_db_save": (_i32, [_vp, C.c_char_p]),
    "lc_db_load": (_i32, [_vp, C.c_char_p]),
    "lc_match_pair": (_i32, [_vp, _vp, _i32, _sz, _vp, _i32, _sz, _vp, _vp, _vp, _pi32, _pi32, _pi32]),
    "lc_match_vs_keyframe": (_i32, [_vp, _vp, _i32, _sz, _i64, _vp, _vp, _vp, _pi32, _pi32, _pi32]),
    "lc_match_keyframes": (_i32, [_vp, _vp, _i32, _sz, _i32, _vp, _vp, _i64, _vp, _vp]),
    "lc_l2_db_append": (_i32, [_vp, _i32, _vp, _i32, _sz, _i32, _vp]),
    "lc_l2_db_size": (_i32, [_vp, _vp, _vp]),
    "lc_l2_db_clear": (_i32, [_vp]),
    "lc_l2_match_pair": (_i32, [_vp, _vp, _i32, _sz, _vp, _i32, _sz, _i32, C.c_double, _vp, _vp, _vp, _vp, _pi32]),
    "lc_l2_loop_check": (_i32, [_vp, _vp, _i32, _sz, _i32, _i32, _i32, C.c_double, _i32, _vp, _i64, _vp]),
    "lc_l2_loop_search": (_i32, [_vp, _i64, _i32, _i32, C.c_double, _i32, _vp, _i64, _vp]),
    "lc_loop_check_begin": (_i32, [_vp, _vp, _i32, _sz, _i32, _i32]),
    "lc_loop_check_end": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "lc_loop_check": (_i32, [_vp, _vp, _i32, _sz, _i32, _vp, _i64, _pi64, _vp]),
    "lc_loop_check_batch": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _i64, _vp]),
    "lc_select_candidates": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _i64, _pi64]),
    "lc_detect_loops": (_i32, [_vp, _vp, _i32, _sz, _i32, _vp, _i64, _pi64]),
    "lc_process_frame": (_i32, [_vp, _i32, _vp, _i32, _sz, _vp, _vp, _vp, _i32, _pi32, _pi32, _pi32, _vp, _i64, _pi64, _pi64]),
    "lc_process_frame_begin": (_i32, [_vp, _i32, _vp, _i32, _sz, _pi64]),
    "lc_process_frame_matches": (_i32, [_vp, _vp, _vp, _vp, _i32, _pi32, _pi32, _pi32]),
    "lc_process_frame_end": (_i32, [_vp, _vp, _i64, _pi64, _vp, _i64, _pi64]),
    "lc_loop_check_device": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "lc_loop_check_batch_device": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lc_merge_shards_device": (_i32, [_vp, _vp, _i64, _i32, _i32, _i64, _i64, _vp, _vp, _vp]),
    "lc_exchange_create": (_i32, [_vp, _i64, _vp, C.POINTER(_vp)]),
    "lc_exchange_connect": (_i32, [_vp, _i32, _i32, _vp, _vp]),
    "lc_exchange_slot": (_i32, [_vp, _i64, C.POINTER(_vp), _pi64]),
    "lc_exchange_close": (_i32, [_vp]),
    "lc_merge_shards_peer_device": (_i32, [_vp, _i64, _i32, _i64, _i64, _vp, _vp, _i32, _vp]),
    "lc_match_pair_device": (_i32, [_vp, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "lc_probe_mxf4_peak": (_i32, [_vp, _i32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "lc_get_last_timing": (_i32, [_vp, C.POINTER(Timing)]),
    "lc_launch_count": (_i64, [_vp]),
    "lc_synchronize": (_i32, [_vp]),
}

_lib = None


def library_path():
    return _build.LIB


def load():
    """Load liblc_b200.so (building it first if the sources are newer and nvcc is present)."""
    global _lib
    if _lib is None:
        path = os.environ.get("LC_B200_LIB")  # developer knob: A/B another build of the same ABI
        if not path:
            path = _build.LIB
            if _build.stale():  # missing, or older than a source / header: never run kernels that do not match the tree
                if _build.have_nvcc():
                    path = _build.build()
                elif not os.path.exists(path):
                    raise RuntimeError("liblc_b200.so is not built and nvcc is not available (there is no CPU fallback)")
                else:
                    import warnings
                    warnings.warn("liblc_b200.so is older than its sources and nvcc is not available: using the stale library")
        lib = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError here = the .so does not export the declared ABI
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def default_config(**kw) -> Config:
    cfg = Config()
    load().lc_default_config(C.byref(cfg))
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise TypeError(f"lc_config has no field {k!r}")
        setattr(cfg, k, int(v) if isinstance(v, (bool, np.integer)) else v)
    return cfg


def _rows(a, name="descriptors"):
    """(n,32) uint8 view of `a` plus its row stride in bytes (no copy for row-strided inputs)."""
    a = np.asarray(a)
    if a.dtype != np.uint8:
        raise TypeError(f"{name} must be uint8, got {a.dtype}")
    if a.size == 0:
        return np.zeros((0, DESC_BYTES), np.uint8), DESC_BYTES
    if a.ndim != 2 or a.shape[1] != DESC_BYTES:
        raise ValueError(f"{name} must have shape (n, 32), got {a.shape}")
    if a.strides[1] != 1 or a.strides[0] < DESC_BYTES:
        a = np.ascontiguousarray(a)
    return a, a.strides[0]


def _ptr(a):
    return a.ctypes.data if a is not None and a.size else None


class Context:
    """Owning wrapper of an lc_ctx."""

    def __init__(self, cfg: Config = None, **kw):
        self._lib = load()
        self._h = _vp()
        cfg = cfg if cfg is not None else default_config(**kw)
        rc = self._lib.lc_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            raise LcError(rc, self._lib.lc_last_error(None).decode())

    def close(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self._lib.lc_destroy(h)
            except TypeError:  # interpreter shutdown: ctypes has already torn the function object down
                pass
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, rc):
        if rc != 0:
            raise LcError(rc, self._lib.lc_last_error(self._h).decode())

    # ---- config
    @property
    def config(self) -> Config:
        cfg = Config()
        self._ck(self._lib.lc_get_config(self._h, C.byref(cfg)))
        return cfg

    def set_params(self, **kw):
        cfg = self.config
        for k, v in kw.items():
            setattr(cfg, k, int(v) if isinstance(v, (bool, np.integer)) else v)
        self._ck(self._lib.lc_set_params(self._h, C.byref(cfg)))

    # ---- store
    def db_append(self, frame_id, desc):
        idx = C.c_int64(-1)
        if desc is None:
            self._ck(self._lib.lc_db_append(self._h, int(frame_id), None, 0, DESC_BYTES, C.byref(idx)))
            return idx.value
        a, stride = _rows(desc)
        self._ck(self._lib.lc_db_append(self._h, int(frame_id), _ptr(a), len(a), stride, C.byref(idx)))
        return idx.value

    def db_append_meta(self, frame_id, n):
        """Record a keyframe owned by another shard (no descriptor bytes)."""
        idx = C.c_int64(-1)
        self._ck(self._lib.lc_db_append(self._h, int(frame_id), None, int(n), DESC_BYTES, C.byref(idx)))
        return idx.value

    def db_append_many(self, frame_ids, keyframes):
        """keyframes: list of (n,32) uint8 arrays, or None / an int row count for unowned keyframes."""
        cnt = len(frame_ids)
        fids = np.ascontiguousarray(frame_ids, np.int32)
        keep, ns, ptrs = [], np.zeros(cnt, np.int32), (C.c_void_p * cnt)()
        for i, kf in enumerate(keyframes):
            if kf is None or isinstance(kf, (int, np.integer)):
                ns[i] = 0 if kf is None else int(kf)
                ptrs[i] = None
            else:
                a = np.ascontiguousarray(_rows(kf)[0])
                keep.append(a)
                ns[i] = len(a)
                ptrs[i] = _ptr(a)
        self._ck(self._lib.lc_db_append_many(self._h, cnt, fids.ctypes.data, ns.ctypes.data, ptrs))

    def db_append_many_device(self, frame_ids, n_rows, d_rows_ptr, stream=None):
        """stream: raw handle of the CUDA stream that produced the rows (torch: `torch.cuda.current_stream().cuda_stream
        or 1` - handle 0 is torch's default stream, which the C ABI names cudaStreamLegacy = 0x1); None = the rows are
        already complete and stay untouched until synchronize()."""
        fids = np.ascontiguousarray(frame_ids, np.int32)
        ns = np.ascontiguousarray(n_rows, np.int32)
        self._ck(self._lib.lc_db_append_many_device(self._h, len(fids), fids.ctypes.data, ns.ctypes.data, d_rows_ptr, stream))

    def db_size(self):
        v = [C.c_int64() for _ in range(4)]
        self._ck(self._lib.lc_db_size(self._h, *[C.byref(x) for x in v]))
        return dict(zip(("keyframes", "rows", "local_keyframes", "local_rows"), (x.value for x in v)))

    def db_clear(self):
        self._ck(self._lib.lc_db_clear(self._h))

    def db_read(self, kf_index):
        n, fid = C.c_int32(), C.c_int32()
        self._ck(self._lib.lc_db_read(self._h, int(kf_index), None, 0, C.byref(n), C.byref(fid)))
        out = np.empty((n.value, DESC_BYTES), np.uint8)
        self._ck(self._lib.lc_db_read(self._h, int(kf_index), _ptr(out), n.value, C.byref(n), C.byref(fid)))
        return fid.value, out

    def db_save(self, path):
        self._ck(self._lib.lc_db_save(self._h, os.fsencode(path)))

    def db_load(self, path):
        self._ck(self._lib.lc_db_load(self._h, os.fsencode(path)))

    # ---- matchFeatures
    def _pair_out(self, nq):
        return (np.empty(nq, np.int32), np.empty(nq, np.int32), np.empty(nq, np.uint8),
                C.c_int32(), C.c_int32(), C.c_int32())

    @staticmethod
    def _pair_ret(idx, dist, good, nm, mn, ng):
        m = nm.value
        return dict(train_idx=idx[:m], dist=dist[:m], good=good[:m].astype(bool), min_dist=mn.value, n_good=ng.value)

    def match_pair(self, q, t):
        qa, qs = _rows(q, "query")
        ta, ts = _rows(t, "train")
        idx, dist, good, nm, mn, ng = self._pair_out(len(qa))
        self._ck(self._lib.lc_match_pair(self._h, _ptr(qa), len(qa), qs, _ptr(ta), len(ta), ts, _ptr(idx), _ptr(dist),
                                         _ptr(good), C.byref(nm), C.byref(mn), C.byref(ng)))
        return self._pair_ret(idx, dist, good, nm, mn, ng)

    def match_vs_keyframe(self, q, kf_index):
        qa, qs = _rows(q, "query")
        idx, dist, good, nm, mn, ng = self._pair_out(len(qa))
        self._ck(self._lib.lc_match_vs_keyframe(self._h, _ptr(qa), len(qa), qs, int(kf_index), _ptr(idx), _ptr(dist),
                                                _ptr(good), C.byref(nm), C.byref(mn), C.byref(ng)))
        return self._pair_ret(idx, dist, good, nm, mn, ng)

    def match_keyframes(self, q, kf_indices):
        """matchFeatures(q, stored keyframe k) for every k in kf_indices, good matches only: a list of (MATCH array,
        min_dist) - MATCH has fields query_idx / train_idx / distance, the cv::DMatch a pose estimator consumes."""
        qa, qs = _rows(q, "query")
        sel = np.ascontiguousarray(kf_indices, np.int64)
        offs = np.zeros(len(sel) + 1, np.int64)
        mind = np.zeros(max(len(sel), 1), np.int32)
        buf = np.zeros(max(len(qa) * len(sel), 1), MATCH)  # an upper bound: every query survives
        self._ck(self._lib.lc_match_keyframes(self._h, _ptr(qa), len(qa), qs, len(sel), _ptr(sel), _ptr(buf), len(buf),
                                              _ptr(offs), _ptr(mind)))
        return [(buf[offs[i]:offs[i + 1]].copy(), int(mind[i])) for i in range(len(sel))]

    # ---- the shipped pipeline's SIFT / L2 matcher (src/main.cpp:509-534, 1375-1392)
    @staticmethod
    def _l2_rows(a, what):
        a = np.asarray(a)
        if a.dtype != np.uint8:
            a = a.astype(np.float32, copy=False)
        if a.ndim != 2 or a.shape[1] != 128:
            raise ValueError(f"{what}: expected (n, 128) SIFT descriptors, got {a.shape}")
        if a.strides[1] != a.itemsize:
            a = np.ascontiguousarray(a)
        return a, int(a.strides[0]) if len(a) else 128 * a.itemsize, 1 if a.dtype == np.uint8 else 0

    def l2_db_append(self, frame_id, desc):
        a, st, ty = self._l2_rows(desc, "descriptors")
        k = C.c_int64()
        self._ck(self._lib.lc_l2_db_append(self._h, int(frame_id), _ptr(a), len(a), st, ty, C.byref(k)))
        return k.value

    def l2_db_append_meta(self, frame_id, n):
        """Record a view this shard does not own (only its place in the global order matters)."""
        k = C.c_int64()
        self._ck(self._lib.lc_l2_db_append(self._h, int(frame_id), None, int(n), 128, 1, C.byref(k)))
        return k.value

    def l2_db_size(self):
        n, nl = C.c_int64(), C.c_int64()
        self._ck(self._lib.lc_l2_db_size(self._h, C.byref(n), C.byref(nl)))
        return dict(keyframes=n.value, local_keyframes=nl.value)

    def l2_db_clear(self):
        self._ck(self._lib.lc_l2_db_clear(self._h))

    def l2_match_pair(self, q, t, ratio=0.75):
        """matchFeatures(desc1 = q, desc2 = t, ratio) of src/main.cpp:509: knnMatch(k = 2) + Lowe's ratio test."""
        qa, qs, ty = self._l2_rows(q, "query")
        ta, ts, ty2 = self._l2_rows(t, "train")
        if ty != ty2:
            raise ValueError("query and train descriptors must have the same dtype")
        n = len(qa)
        idx, dist, sec, good = np.zeros(n, np.int32), np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.uint8)
        ng = C.c_int32()
        self._ck(self._lib.lc_l2_match_pair(self._h, _ptr(qa), n, qs, _ptr(ta), len(ta), ts, ty, float(ratio), _ptr(idx),
                                            _ptr(dist), _ptr(sec), _ptr(good), C.byref(ng)))
        return dict(train_idx=idx, distance=dist, second=sec, good=good.astype(bool), n_good=ng.value)

    def l2_loop_check(self, q, current_frame_id, loop_gap, ratio=0.7, min_rows=100):
        qa, qs, ty = self._l2_rows(q, "query")
        nloc = self.l2_db_size()["local_keyframes"]
        res = np.zeros(nloc, L2_RESULT)
        n = C.c_int64()
        self._ck(self._lib.lc_l2_loop_check(self._h, _ptr(qa), len(qa), qs, ty, int(current_frame_id), int(loop_gap),
                                            float(ratio), int(min_rows), _ptr(res), nloc, C.byref(n)))
        return res[:n.value]

    def l2_loop_search(self, first_view, n_query_views, loop_gap, ratio=0.7, min_rows=100):
        """The shipped loop search's matching step for stored views first_view .. first_view + n_query_views - 1 against the
        whole resident store: L2_RESULT array [n_query_views, n_views]."""
        n_views = self.l2_db_size()["keyframes"]
        res = np.zeros((int(n_query_views), n_views), L2_RESULT)
        n = C.c_int64()
        self._ck(self._lib.lc_l2_loop_search(self._h, int(first_view), int(n_query_views), int(loop_gap), float(ratio), int(min_rows), _ptr(res),
                                             res.size, C.byref(n)))
        return res

    # ---- detectLoops
    def loop_check(self, q, current_frame_id, want_best=True):
        qa, qs = _rows(q, "query")
        nloc = self.db_size()["local_keyframes"]
        res = np.zeros(nloc, KF_RESULT)
        best = np.full(len(qa), UINT64_MAX, np.uint64) if want_best else None
        n = C.c_int64()
        self._ck(self._lib.lc_loop_check(self._h, _ptr(qa), len(qa), qs, int(current_frame_id), _ptr(res), nloc,
                                         C.byref(n), _ptr(best) if want_best else None))
        return res[:n.value], best

    def loop_check_begin(self, q, current_frame_id, want_best=True):
        """Queue a loop check and return at once (the library has copied q); collect with loop_check_end()."""
        qa, qs = _rows(q, "query")
        self._ck(self._lib.lc_loop_check_begin(self._h, _ptr(qa), len(qa), qs, int(current_frame_id), int(bool(want_best))))
        self._pending = (len(qa), bool(want_best))

    def loop_check_end(self):
        nq, want_best = self._pending
        nloc = self.db_size()["local_keyframes"]
        res = np.zeros(nloc, KF_RESULT)
        best = np.full(nq, UINT64_MAX, np.uint64) if want_best else None
        n = C.c_int64()
        self._ck(self._lib.lc_loop_check_end(self._h, _ptr(res), nloc, C.byref(n), _ptr(best) if want_best else None))
        return res[:n.value], best

    def loop_check_batch(self, queries, current_frame_ids, want_best=True):
        qs = [np.ascontiguousarray(_rows(q, "query")[0]) for q in queries]
        nf = len(qs)
        nq = np.array([len(q) for q in qs], np.int32)
        cur = np.ascontiguousarray(current_frame_ids, np.int32)
        ptrs = (C.c_void_p * max(nf, 1))(*[_ptr(q) for q in qs])
        nloc = self.db_size()["local_keyframes"]
        res = np.zeros((nf, nloc), KF_RESULT)
        best = np.full(int(nq.sum()), UINT64_MAX, np.uint64) if want_best else None
        self._ck(self._lib.lc_loop_check_batch(self._h, nf, ptrs, nq.ctypes.data, cur.ctypes.data, _ptr(res), res.size,
                                               _ptr(best) if want_best else None))
        if want_best:
            offs = np.concatenate([[0], np.cumsum(nq)])
            best = [best[offs[i]:offs[i + 1]] for i in range(nf)]
        return res, best

    def select_candidates(self, results, nq, current_frame_id):
        results = np.ascontiguousarray(results, KF_RESULT)
        out = np.zeros(len(results), LOOP_CANDIDATE)
        n = C.c_int64()
        self._ck(self._lib.lc_select_candidates(self._h, _ptr(results), len(results), int(nq), int(current_frame_id),
                                                _ptr(out), len(out), C.byref(n)))
        return out[:n.value]

    def detect_loops(self, q, current_frame_id):
        qa, qs = _rows(q, "query")
        cap = max(1, self.db_size()["local_keyframes"])
        out = np.zeros(cap, LOOP_CANDIDATE)
        n = C.c_int64()
        self._ck(self._lib.lc_detect_loops(self._h, _ptr(qa), len(qa), qs, int(current_frame_id), _ptr(out), cap, C.byref(n)))
        return out[:n.value]

    # ---- processFrame: pair match vs the previous frame + append + loop check in one call
    def _rows_of_last_keyframe(self):
        k = self.db_size()["keyframes"]
        if k == 0:
            return 0
        n, fid = C.c_int32(), C.c_int32()
        self._ck(self._lib.lc_db_read(self._h, k - 1, None, 0, C.byref(n), C.byref(fid)))
        return n.value

    def process_frame(self, frame_id, desc):
        """lc_process_frame.  Returns (pair, candidates, kf_index): pair as match_pair returns it (query = the previous
        frame, train = this one; empty for the first frame), candidates a LOOP_CANDIDATE array."""
        a, stride = _rows(desc)
        n_prev = self._rows_of_last_keyframe()
        idx, dist, good, nm, mn, ng = self._pair_out(n_prev)
        cap = self.db_size()["local_keyframes"] + 1
        out = np.zeros(cap, LOOP_CANDIDATE)
        n_out, k = C.c_int64(), C.c_int64(-1)
        self._ck(self._lib.lc_process_frame(self._h, int(frame_id), _ptr(a), len(a), stride, _ptr(idx), _ptr(dist), _ptr(good),
                                            n_prev, C.byref(nm), C.byref(mn), C.byref(ng), _ptr(out), cap, C.byref(n_out), C.byref(k)))
        return self._pair_ret(idx, dist, good, nm, mn, ng), out[:n_out.value], k.value

    def process_frame_begin(self, frame_id, desc):
        """Queue pair match + append + loop check of this frame and return at once; collect with process_frame_matches()
        (optional, waits for the pair match only) and process_frame_end()."""
        a, stride = _rows(desc)
        n_prev = self._rows_of_last_keyframe()
        k = C.c_int64(-1)
        self._ck(self._lib.lc_process_frame_begin(self._h, int(frame_id), _ptr(a), len(a), stride, C.byref(k)))
        self._frame_n_prev = n_prev  # rows of the previous frame = entries of the pair-match outputs of the frame in flight
        return k.value

    def process_frame_matches(self):
        n_prev = getattr(self, "_frame_n_prev", 0)
        idx, dist, good, nm, mn, ng = self._pair_out(n_prev)
        self._ck(self._lib.lc_process_frame_matches(self._h, _ptr(idx), _ptr(dist), _ptr(good), n_prev, C.byref(nm), C.byref(mn), C.byref(ng)))
        return self._pair_ret(idx, dist, good, nm, mn, ng)

    def process_frame_end(self, want_records=False):
        nloc = self.db_size()["local_keyframes"]
        out = np.zeros(max(1, nloc), LOOP_CANDIDATE)
        res = np.zeros(nloc, KF_RESULT) if want_records else None
        n_out, n_res = C.c_int64(), C.c_int64()
        self._ck(self._lib.lc_process_frame_end(self._h, _ptr(out), len(out), C.byref(n_out), _ptr(res) if want_records else None,
                                                nloc, C.byref(n_res)))
        return (out[:n_out.value], res[:n_res.value]) if want_records else out[:n_out.value]

    # ---- device-pointer variants (raw addresses; see sharded.py for the torch plumbing)
    def loop_check_device(self, d_q, nq, current_frame_id, d_results, d_best=None, stream=None):
        self._ck(self._lib.lc_loop_check_device(self._h, d_q, int(nq), int(current_frame_id), d_results, d_best, stream))

    def loop_check_batch_device(self, d_q, nq, current_frame_ids, d_results, d_best=None, stream=None):
        nq = np.ascontiguousarray(nq, np.int32)
        cur = np.ascontiguousarray(current_frame_ids, np.int32)
        self._ck(self._lib.lc_loop_check_batch_device(self._h, len(nq), d_q, nq.ctypes.data, cur.ctypes.data,
                                                      d_results, d_best, stream))

    def merge_shards_device(self, d_gathered, contrib_bytes, world, n_frames, slots, rows, d_records_out, d_best_out, stream=None):
        self._ck(self._lib.lc_merge_shards_device(self._h, d_gathered, int(contrib_bytes), int(world), int(n_frames), int(slots),
                                                  int(rows), d_records_out, d_best_out, stream))

    # ---- peer-memory exchange of a sharded check (no collective)
    def exchange_create(self, slot_bytes):
        """Returns (64-byte CUDA IPC handle, device address) of this rank's exchange buffer."""
        handle = (C.c_ubyte * 64)()
        base = _vp()
        self._ck(self._lib.lc_exchange_create(self._h, int(slot_bytes), handle, C.byref(base)))
        return bytes(handle), base.value

    def exchange_connect(self, world, rank, handles=None, bases=None):
        """handles: list of `world` 64-byte IPC handles in rank order (ranks in other processes); bases: list of device
        addresses (ranks of this process)."""
        h = b"".join(handles) if handles is not None else None
        b = (C.c_void_p * world)(*bases) if bases is not None else None
        self._ck(self._lib.lc_exchange_connect(self._h, int(world), int(rank), h, b))

    def exchange_slot(self, seq):
        p, n = _vp(), C.c_int64()
        self._ck(self._lib.lc_exchange_slot(self._h, int(seq), C.byref(p), C.byref(n)))
        return p.value, n.value

    def exchange_close(self):
        self._ck(self._lib.lc_exchange_close(self._h))

    def merge_shards_peer_device(self, seq, n_frames, slots, rows, d_records_out, d_best_out, phase=3, stream=None):
        self._ck(self._lib.lc_merge_shards_peer_device(self._h, int(seq), int(n_frames), int(slots), int(rows), d_records_out, d_best_out,
                                                       int(phase), stream))

    def match_pair_device(self, d_q, nq, d_t, nt, d_idx, d_dist, d_good, d_summary, stream=None):
        self._ck(self._lib.lc_match_pair_device(self._h, d_q, int(nq), d_t, int(nt), d_idx, d_dist, d_good, d_summary, stream))

    # ---- introspection
    def probe_mxf4_peak(self, milliseconds=100):
        """(MAC/s of the whole GPU, SM MHz held) of a pure tcgen05.mma kind::mxf4 issue loop: the tensor roofline's denominator."""
        macs, mhz = C.c_double(), C.c_double()
        self._ck(self._lib.lc_probe_mxf4_peak(self._h, int(milliseconds), C.byref(macs), C.byref(mhz)))
        return macs.value, mhz.value

    def last_timing(self) -> Timing:
        t = Timing()
        self._ck(self._lib.lc_get_last_timing(self._h, C.byref(t)))
        return t

    def launch_count(self):
        return int(self._lib.lc_launch_count(self._h))

    def synchronize(self):
        self._ck(self._lib.lc_synchronize(self._h))
