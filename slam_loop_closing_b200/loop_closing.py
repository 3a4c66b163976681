"""Host-side mirror of the reference's loop-closing interface with the matcher swapped for liblc_b200.

Mirrors /root/reference/include/loop_closing.hpp:9-82 (namespace loop_closing): `Frame` (:12-19),
`LoopCandidate` (:22-27) and `LoopClosingSystem` (:29-80) keep their names, members, argument meaning
and defaults.  Only the bodies behind matchFeatures (:40) and detectLoops (:48) - the
cv::BFMatcher(NORM_HAMMING) of :73 - run on the GPU through the C ABI (include/lc_abi.h); ORB
detection, essential-matrix RANSAC, recoverPose, triangulation and drawing stay on the host in OpenCV
exactly as BASELINE.json's north_star says.  The reference's src/loop_closing.cpp is absent, so the
workflow follows README.md:92-146.

The reference is C++; its toolchain's OpenCV dev files are not in this image, so this executable twin
is Python over the same C ABI that include/loop_closing_b200.hpp binds for C++.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import _native

try:  # OpenCV is only needed for the host-side stages (ORB, E-matrix, triangulation, drawing)
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


@dataclass
class Frame:  # include/loop_closing.hpp:12-19
    id: int
    image: Optional[np.ndarray] = None
    keypoints: list = field(default_factory=list)
    descriptors: np.ndarray = field(default_factory=lambda: np.zeros((0, 32), np.uint8))
    pose: np.ndarray = field(default_factory=lambda: np.eye(4))
    points3D: list = field(default_factory=list)


@dataclass
class LoopCandidate:  # include/loop_closing.hpp:22-27
    current_frame_id: int
    matched_frame_id: int
    num_matches: int
    similarity_score: float


@dataclass
class DMatch:
    """cv::DMatch as BFMatcher::match returns it (imgIdx is always 0 for a two-frame match)."""
    queryIdx: int
    trainIdx: int
    imgIdx: int
    distance: float


class B200Matcher:
    """The GPU matcher + device-resident keyframe store behind matchFeatures / detectLoops."""

    def __init__(self, **cfg):
        self.ctx = _native.Context(**cfg)

    def match_pair(self, q, t):
        return self.ctx.match_pair(q, t)

    def append(self, frame_id, descriptors):
        return self.ctx.db_append(frame_id, descriptors)

    def loop_check(self, q, current_frame_id):
        res, _ = self.ctx.loop_check(q, current_frame_id, want_best=False)
        return res

    def match_keyframes(self, q, kf_indices):
        return self.ctx.match_keyframes(q, kf_indices)

    # the matcher side of a whole processFrame in one GPU submission (lc_process_frame_begin / _matches / _end)
    def frame_begin(self, frame_id, descriptors):
        return self.ctx.process_frame_begin(frame_id, descriptors)

    def frame_matches(self):
        return self.ctx.process_frame_matches()

    def frame_end(self):
        return self.ctx.process_frame_end()

    def close(self):
        self.ctx.close()


class LoopClosingSystem:
    """loop_closing::LoopClosingSystem (include/loop_closing.hpp:29-80)."""

    def __init__(self, loop_threshold: float = 0.7, min_loop_gap: int = 30, *, matcher=None, nfeatures: int = 2000,
                 min_good: int = 50, filter_floor: int = 0, filter_strict: bool = False, min_init: int = -1,
                 gap_strict: bool = False, thr_inclusive: bool = False, device: int = 0, fused: bool = True):
        self.loop_threshold_ = float(loop_threshold)
        self.min_loop_gap_ = int(min_loop_gap)
        self.min_good_ = int(min_good)
        self.thr_inclusive_ = bool(thr_inclusive)
        self.frames_: List[Frame] = []
        self.loop_closures_: List[LoopCandidate] = []
        self.consecutive_matches_: List[int] = []  # good-match count of every (prev, cur) pair, for tests
        self.nfeatures_ = nfeatures
        self.fused_ = bool(fused)  # one GPU submission per frame (lc_process_frame_*) when the matcher offers it
        self.feature_detector_ = None  # cv::ORB, created on first use (include/loop_closing.hpp:72)
        # the matcher_ of include/loop_closing.hpp:73; anything with match_pair/append/loop_check fits
        self.matcher_ = matcher if matcher is not None else B200Matcher(
            device=device, loop_threshold=loop_threshold, min_loop_gap=min_loop_gap, min_good=min_good,
            filter_floor=filter_floor, filter_strict=filter_strict, min_init=min_init, gap_strict=gap_strict,
            thr_inclusive=thr_inclusive)
        # README.md:136 "Camera intrinsics assumed (fx=fy=800, cx=640, cy=360)"
        self.K_ = np.array([[800.0, 0, 640.0], [0, 800.0, 360.0], [0, 0, 1.0]])

    # ------------------------------------------------------------------ per-frame driver (hpp:34)
    def processFrame(self, image: np.ndarray, frame_id: int) -> None:
        frame = Frame(id=int(frame_id), image=image)
        self.detectFeatures(frame)
        self._process(frame, geometry=True)

    def processDescriptors(self, descriptors: np.ndarray, frame_id: int) -> List[LoopCandidate]:
        """processFrame for callers that already hold the ORB descriptors (no image): the hot path only."""
        frame = Frame(id=int(frame_id), descriptors=np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 32))
        return self._process(frame, geometry=False)

    def _process(self, frame: Frame, geometry: bool) -> List[LoopCandidate]:
        fused = self.fused_ and hasattr(self.matcher_, "frame_begin")
        if fused:
            # one GPU submission for the frame: pair match vs the previous frame, append, loop check.  The pose stages
            # below run on the host while the GPU is still scanning the store.
            self.matcher_.frame_begin(frame.id, frame.descriptors)
        if self.frames_:  # README.md:96-100
            prev = self.frames_[-1]
            matches = self._dmatches(self.matcher_.frame_matches()) if fused else self.matchFeatures(prev, frame)
            self.consecutive_matches_.append(len(matches))
            if geometry and len(matches) >= 8:  # README.md:132
                ok, R, t = self.estimatePose(prev, frame, matches)
                if ok:
                    T = np.eye(4)
                    T[:3, :3], T[:3, 3] = R, t.ravel()
                    frame.pose = prev.pose @ np.linalg.inv(T)
                    frame.points3D = self.triangulatePoints(prev, frame, matches)
        self.frames_.append(frame)  # hpp:69 - and the device store gets the descriptors
        if fused:
            loops = [LoopCandidate(int(c["current_frame_id"]), int(c["matched_frame_id"]), int(c["num_matches"]),
                                   float(c["similarity_score"])) for c in self.matcher_.frame_end()]
        else:
            self.matcher_.append(frame.id, frame.descriptors)
            loops = self.detectLoops(frame.id)
        self.loop_closures_.extend(loops)
        return loops

    # ------------------------------------------------------------------ host stages that stay in OpenCV
    def detectFeatures(self, frame: Frame) -> None:  # hpp:37, README.md:114
        if cv2 is None:
            raise RuntimeError("OpenCV (cv2) is required for ORB detection")
        if self.feature_detector_ is None:
            self.feature_detector_ = cv2.ORB_create(self.nfeatures_)
        img = frame.image if frame.image.ndim == 2 else cv2.cvtColor(frame.image, cv2.COLOR_BGR2GRAY)
        kps, desc = self.feature_detector_.detectAndCompute(img, None)
        frame.keypoints = list(kps)
        frame.descriptors = desc if desc is not None else np.zeros((0, 32), np.uint8)

    def estimatePose(self, frame1: Frame, frame2: Frame, matches: List[DMatch]):  # hpp:43, README.md:128-132
        if len(matches) < 8 or not frame1.keypoints or not frame2.keypoints:
            return False, None, None
        p1 = np.float64([frame1.keypoints[m.queryIdx].pt for m in matches])
        p2 = np.float64([frame2.keypoints[m.trainIdx].pt for m in matches])
        E, mask = cv2.findEssentialMat(p1, p2, self.K_, method=cv2.RANSAC, prob=0.999, threshold=1.0)
        if E is None or E.shape != (3, 3):
            return False, None, None
        _, R, t, _ = cv2.recoverPose(E, p1, p2, self.K_, mask=mask)
        return True, R, t

    def triangulatePoints(self, frame1: Frame, frame2: Frame, matches: List[DMatch]):  # hpp:51, README.md:134-138
        if not matches or not frame1.keypoints or not frame2.keypoints:
            return []
        P1 = self.K_ @ np.linalg.inv(frame1.pose)[:3]
        P2 = self.K_ @ np.linalg.inv(frame2.pose)[:3]
        p1 = np.float64([frame1.keypoints[m.queryIdx].pt for m in matches]).T
        p2 = np.float64([frame2.keypoints[m.trainIdx].pt for m in matches]).T
        X = cv2.triangulatePoints(P1, P2, p1, p2)
        X = (X[:3] / np.where(np.abs(X[3]) < 1e-12, 1e-12, X[3])).T
        cam = (np.linalg.inv(frame2.pose) @ np.c_[X, np.ones(len(X))].T).T
        keep = (cam[:, 2] > 0) & (np.linalg.norm(X, axis=1) < 100.0)  # README.md:138
        return [tuple(p) for p in X[keep]]

    def visualizeMatches(self, frame1: Frame, frame2: Frame, matches: List[DMatch]):  # hpp:56
        cvm = [cv2.DMatch(m.queryIdx, m.trainIdx, m.imgIdx, m.distance) for m in matches]
        return cv2.drawMatches(frame1.image, frame1.keypoints, frame2.image, frame2.keypoints, cvm, None)

    # ------------------------------------------------------------------ the hot path (GPU)
    @staticmethod
    def _dmatches(r) -> List[DMatch]:
        return [DMatch(int(i), int(r["train_idx"][i]), 0, float(r["dist"][i])) for i in np.flatnonzero(r["good"])]

    def matchFeatures(self, frame1: Frame, frame2: Frame) -> List[DMatch]:  # hpp:40, README.md:116-117
        return self._dmatches(self.matcher_.match_pair(frame1.descriptors, frame2.descriptors))

    def detectLoops(self, current_frame_id: int) -> List[LoopCandidate]:  # hpp:48, README.md:121-126
        cur = next((f for f in reversed(self.frames_) if f.id == current_frame_id), None)
        if cur is None:
            return []
        nq = len(cur.descriptors)
        out = []
        for r in self.matcher_.loop_check(cur.descriptors, current_frame_id):
            if r["n_good"] < 0:
                continue
            denom = min(nq, int(r["n_kf"]))
            if denom == 0:
                continue
            sim = float(int(r["n_good"])) / float(denom)  # README.md:126
            over = sim >= self.loop_threshold_ if self.thr_inclusive_ else sim > self.loop_threshold_
            if over and int(r["n_good"]) >= self.min_good_:  # README.md:123-124
                out.append(LoopCandidate(int(current_frame_id), int(r["frame_id"]), int(r["n_good"]), sim))
        return out

    def loopMatches(self, current_frame_id: int, loops: List[LoopCandidate]) -> dict:
        """matchFeatures(current, matched frame) for every loop candidate in ONE GPU call (lc_match_keyframes): the
        DMatch lists estimatePose / triangulatePoints (hpp:43-53) need to verify a loop geometrically, without
        matching again on the CPU.  Returns {matched_frame_id: [DMatch]}."""
        cur = next((f for f in reversed(self.frames_) if f.id == current_frame_id), None)
        if cur is None or not loops:
            return {}
        index_of = {f.id: i for i, f in enumerate(self.frames_)}  # frames_ and the device store share insertion order
        sel = [index_of[c.matched_frame_id] for c in loops]
        out = {}
        for c, (m, _) in zip(loops, self.matcher_.match_keyframes(cur.descriptors, sel)):
            out[c.matched_frame_id] = [DMatch(int(a), int(b), 0, float(d)) for a, b, d in
                                       zip(m["query_idx"], m["train_idx"], m["distance"])]
        return out

    # ------------------------------------------------------------------ accessors / output
    def getFrames(self) -> List[Frame]:  # hpp:60
        return self.frames_

    def getLoopClosures(self) -> List[LoopCandidate]:  # hpp:63
        return self.loop_closures_

    def saveResults(self, output_dir: str) -> None:  # hpp:66, README.md:142-165
        os.makedirs(output_dir, exist_ok=True)
        with open(os.path.join(output_dir, "loop_closures.txt"), "w") as f:
            f.write(f"Total frames processed: {len(self.frames_)}\n")
            f.write(f"Loop closures detected: {len(self.loop_closures_)}\n\n")
            f.write("Loop Closures Detected:\n======================\n\n")
            for lc in self.loop_closures_:
                f.write(f"Frame {lc.current_frame_id} <-> Frame {lc.matched_frame_id}\n")
                f.write(f"  Matches: {lc.num_matches}\n  Similarity: {lc.similarity_score:.6g}\n\n")
