"""Host-side mirror of the reference's matcher interface over the C ABI.

Same names, argument meaning and error behaviour as the reference
(paths relative to /root/reference):

* ``Matching.Options`` / ``Matching.Result``      src/mve/sfm/matching.h:29-62
* ``MatchingBase.Options``                        src/mve/sfm/matching_base.h:25-31
* ``ExhaustiveMatching.init / pairwise_match / pairwise_match_lowres``
                                                  src/mve/sfm/exhaustive_matching.cc:56,115,147
* ``FeatureSet`` / ``Viewport``                   src/mve/sfm/feature_set.h:70-72,
                                                  src/mve/sfm/bundler_common.h:37-62

All matching work happens in libosfm_match.so (CUDA, sm_100a).  Nothing here falls back
to the CPU: if the library or the GPU is missing, construction raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import KIND_SIFT_U8, KIND_SURF_S8, MatcherError

FLT_MAX = float(np.finfo(np.float32).max)


class Matching:
    """sfm::Matching -- options / result types and the small list utilities."""

    @dataclass
    class Options:  # matching.h:29-50
        descriptor_length: int
        lowe_ratio_threshold: float
        distance_threshold: float

    @dataclass
    class Result:  # matching.h:56-62
        matches_1_2: np.ndarray = field(default_factory=lambda: np.empty(0, np.int32))
        matches_2_1: np.ndarray = field(default_factory=lambda: np.empty(0, np.int32))

    @staticmethod
    def count_consistent_matches(matches: "Matching.Result") -> int:
        """matching.cc:39-47 (host utility for callers that already hold a Result)."""
        m12 = np.asarray(matches.matches_1_2)
        m21 = np.asarray(matches.matches_2_1)
        idx = np.nonzero(m12 != -1)[0]
        if idx.size == 0:
            return 0
        return int(np.count_nonzero(m21[m12[idx]] == idx))


class MatchingBase:
    @dataclass
    class Options:  # matching_base.h:25-31
        sift_matching_opts: Matching.Options = field(
            default_factory=lambda: Matching.Options(128, 0.8, FLT_MAX))
        surf_matching_opts: Matching.Options = field(
            default_factory=lambda: Matching.Options(64, 0.7, FLT_MAX))


@dataclass
class FeatureSet:
    """The slice of sfm::FeatureSet the matcher reads (feature_set.h:64-72).

    Descriptors are either float arrays as Sift/Surf produce them (n x 128 in [0,1],
    n x 64 in [-1,1]) or already quantised (uint8 n x 128, int8 n x 64)."""
    sift_descriptors: Optional[np.ndarray] = None
    surf_descriptors: Optional[np.ndarray] = None
    positions: Optional[np.ndarray] = None

    def clear_descriptors(self) -> None:  # feature_set.h:58
        self.sift_descriptors = None
        self.surf_descriptors = None


@dataclass
class Viewport:  # bundler_common.h:37-62
    features: FeatureSet = field(default_factory=FeatureSet)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None or a.size == 0 else a.ctypes.data_as(C.c_void_p)


TWO_VIEW_OK, TWO_VIEW_SKIPPED, TWO_VIEW_LOWRES_REJECTED, TWO_VIEW_TOO_FEW_MATCHES = 0, 1, 2, 3
TWO_VIEW_TOO_FEW_INLIERS = 4


class TwoViewOptions:
    """The matcher-side fields of sfm::bundler::Matching::Options (bundler_matching.h:58-76)."""

    def __init__(self, use_lowres_matching=False, num_lowres_features=500, min_lowres_matches=5,
                 min_feature_matches=24, match_num_previous_frames=0, min_matching_inliers=12,
                 ransac_max_iterations=1000, ransac_threshold=0.0015):
        self.min_matching_inliers = min_matching_inliers
        self.ransac_max_iterations = ransac_max_iterations
        self.ransac_threshold = ransac_threshold
        self.use_lowres_matching = use_lowres_matching
        self.num_lowres_features = num_lowres_features
        self.min_lowres_matches = min_lowres_matches
        self.min_feature_matches = min_feature_matches
        self.match_num_previous_frames = match_num_previous_frames


class PackedViews:
    """Quantised descriptors of a list of viewports with the pointer tables the C ABI takes,
    built once (ExhaustiveMatching.init accepts it in place of the viewports and then stages
    every view with a single call)."""

    def __init__(self, viewports: Sequence["Viewport"]):
        self.sift, self.surf, self.sizes = [], [], []
        for vp in viewports:
            fs = vp.features if hasattr(vp, "features") else vp
            s = None if fs.sift_descriptors is None else np.asarray(fs.sift_descriptors)
            f = None if fs.surf_descriptors is None else np.asarray(fs.surf_descriptors)
            if (s is not None and s.dtype != np.uint8) or (f is not None and f.dtype != np.int8):
                raise ValueError("PackedViews takes quantised descriptors (uint8 SIFT, int8 SURF)")
            s = None if s is None else np.ascontiguousarray(s).reshape(-1, 128)
            f = None if f is None else np.ascontiguousarray(f).reshape(-1, 64)
            self.sift.append(s)
            self.surf.append(f)
            self.sizes.append((0 if s is None else s.shape[0], 0 if f is None else f.shape[0]))
        n = len(self.sizes)
        self.p_sift = (C.c_void_p * n)(*[None if a is None or a.shape[0] == 0 else a.ctypes.data for a in self.sift])
        self.p_surf = (C.c_void_p * n)(*[None if a is None or a.shape[0] == 0 else a.ctypes.data for a in self.surf])
        self.n_sift = (C.c_int32 * n)(*[a for a, _ in self.sizes])
        self.n_surf = (C.c_int32 * n)(*[b for _, b in self.sizes])

    def __len__(self):
        return len(self.sizes)


class ExhaustiveMatching:
    """Drop-in for sfm::ExhaustiveMatching, computing on a B200.

    ``init`` stages every view's descriptors into HBM once; ``pairwise_match`` and
    ``pairwise_match_lowres`` return exactly what the reference returns.  ``match_pairs``
    is the batched form the pipeline should prefer (all pairs in one persistent kernel).
    """

    def __init__(self, opts: Optional[MatchingBase.Options] = None, device: int = 0,
                 devices: Optional[Sequence[int]] = None):
        """``devices``: several GPUs of this box behind one handle (osfm_match_create_multi): the
        pool is replicated with an NCCL broadcast at init, the batched calls are sharded."""
        self.opts = opts if opts is not None else MatchingBase.Options()
        if self.opts.sift_matching_opts.descriptor_length != 128 or \
                self.opts.surf_matching_opts.descriptor_length != 64:
            raise ValueError("descriptor lengths are fixed: 128 (SIFT) and 64 (SURF)")
        self._L = _lib.load()
        cfg = _lib.Config()
        self._L.osfm_match_default_config(C.byref(cfg))
        cfg.device = device
        cfg.sift_lowe_ratio = self.opts.sift_matching_opts.lowe_ratio_threshold
        cfg.sift_distance_threshold = self.opts.sift_matching_opts.distance_threshold
        cfg.surf_lowe_ratio = self.opts.surf_matching_opts.lowe_ratio_threshold
        cfg.surf_distance_threshold = self.opts.surf_matching_opts.distance_threshold
        self._h = C.c_void_p()
        if devices is not None and len(devices) > 0:
            devs = (C.c_int * len(devices))(*[int(d) for d in devices])
            rc = self._L.osfm_match_create_multi(C.byref(cfg), devs, len(devices), C.byref(self._h))
        else:
            rc = self._L.osfm_match_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            msg = self._L.osfm_match_last_error(self._h).decode() if self._h else "allocation failed"
            if self._h:
                self._L.osfm_match_destroy(self._h)
                self._h = C.c_void_p()
            raise MatcherError(rc, msg)
        self._sizes: List[tuple] = []
        self._keepalive = None
        self._staged_sources = None

    # -- plumbing ---------------------------------------------------------------------
    def _check(self, rc: int) -> None:
        if rc != 0:
            raise MatcherError(rc, self._L.osfm_match_last_error(self._h).decode())

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            self._L.osfm_match_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- ExhaustiveMatching::init ----------------------------------------------------------
    def init(self, viewports: Sequence[Viewport], overlap_copies: bool = False) -> None:
        """``overlap_copies``: stage through osfm_match_begin_overlapped -- commit does not wait
        for the host-to-device copies, the first batched call matches the early pairs while the
        later views arrive.  The descriptor arrays (quantised, uint8 / int8) are kept alive
        here until the next init / wait_staged / close."""
        if viewports is None:
            raise ValueError("Viewports must not be null")  # bundler_matching.cc:47-48
        begin = self._L.osfm_match_begin_overlapped if overlap_copies else self._L.osfm_match_begin
        self._check(begin(self._h, len(viewports)))
        if isinstance(viewports, PackedViews):
            pk = viewports
            self._check(self._L.osfm_match_set_views_q8(self._h, 0, len(pk), pk.p_sift, pk.n_sift, pk.p_surf, pk.n_surf))
            self._sizes = list(pk.sizes)
            self._check(self._L.osfm_match_commit(self._h))
            self._staged_sources = pk if overlap_copies else None
            return
        self._sizes = []
        keep = []   # staging copies are asynchronous: sources must outlive osfm_match_commit
        for v, vp in enumerate(viewports):
            fs = vp.features if hasattr(vp, "features") else vp
            sift = None if fs.sift_descriptors is None else np.asarray(fs.sift_descriptors)
            surf = None if fs.surf_descriptors is None else np.asarray(fs.surf_descriptors)
            n_sift = 0 if sift is None else sift.reshape(-1, 128).shape[0]
            n_surf = 0 if surf is None else surf.reshape(-1, 64).shape[0]
            quantised = (sift is None or sift.dtype == np.uint8) and (surf is None or surf.dtype == np.int8)
            if quantised:
                s = None if sift is None else np.ascontiguousarray(sift, np.uint8).reshape(-1, 128)
                f = None if surf is None else np.ascontiguousarray(surf, np.int8).reshape(-1, 64)
                self._check(self._L.osfm_match_set_view_q8(self._h, v, _ptr(s), n_sift, _ptr(f), n_surf))
            else:
                s = None if sift is None else np.ascontiguousarray(sift, np.float32).reshape(-1, 128)
                f = None if surf is None else np.ascontiguousarray(surf, np.float32).reshape(-1, 64)
                self._check(self._L.osfm_match_set_view_f32(self._h, v, _ptr(s), n_sift, 128,
                                                            _ptr(f), n_surf, 64))
            keep.append((s, f))
            self._sizes.append((n_sift, n_surf))
        self._check(self._L.osfm_match_commit(self._h))
        self._staged_sources = keep if overlap_copies else None

    def wait_staged(self) -> None:
        """After an overlapped init: returns when every view has arrived on the device."""
        self._check(self._L.osfm_match_wait_staged(self._h))
        self._staged_sources = None

    def init_device_pool(self, sift_pool, row_offsets, sizes) -> None:
        """Adopts a SIFT descriptor pool already resident on this GPU (a torch uint8
        tensor of shape [rows + >=256 padding rows, 128]); used by the multi-GPU path after
        the NCCL broadcast.  ``row_offsets[v]`` / ``sizes[v]`` locate view v."""
        off = np.ascontiguousarray(row_offsets, np.int64)
        n = np.ascontiguousarray(sizes, np.int32)
        rows = int(sift_pool.shape[0]) - 256
        if rows < 0 or (off.size and int((off + n).max()) > rows):
            raise ValueError("pool needs 256 padding rows after the last view")
        self._check(self._L.osfm_match_commit_device(
            self._h, int(n.size), C.c_void_p(int(sift_pool.data_ptr())),
            off.ctypes.data_as(C.POINTER(C.c_int64)), n.ctypes.data_as(C.POINTER(C.c_int32)),
            C.c_int64(rows), None, None, None, C.c_int64(0)))
        self._sizes = [(int(x), 0) for x in n]
        self._keepalive = sift_pool

    @property
    def num_devices(self) -> int:
        return int(self._L.osfm_match_num_devices(self._h))

    @property
    def num_views(self) -> int:
        return len(self._sizes)

    def view_size(self, view_id: int) -> tuple:
        return self._sizes[view_id]

    # -- ExhaustiveMatching::pairwise_match ------------------------------------------------
    def pairwise_match(self, view_1_id: int, view_2_id: int,
                       result: Optional[Matching.Result] = None) -> Matching.Result:
        n1 = sum(self._sizes[view_1_id]) if 0 <= view_1_id < len(self._sizes) else 0
        n2 = sum(self._sizes[view_2_id]) if 0 <= view_2_id < len(self._sizes) else 0
        m12 = np.empty(max(n1, 1), np.int32)
        m21 = np.empty(max(n2, 1), np.int32)
        l12, l21, cnt = C.c_int(0), C.c_int(0), C.c_int(0)
        i32p = C.POINTER(C.c_int32)
        self._check(self._L.osfm_match_pair(self._h, view_1_id, view_2_id,
                                            m12.ctypes.data_as(i32p), C.byref(l12),
                                            m21.ctypes.data_as(i32p), C.byref(l21), C.byref(cnt)))
        if result is None:
            result = Matching.Result()
        result.matches_1_2 = m12[:l12.value].copy()
        result.matches_2_1 = m21[:l21.value].copy()
        self.last_consistent = cnt.value
        return result

    # -- ExhaustiveMatching::pairwise_match_lowres -----------------------------------------
    def pairwise_match_lowres(self, view_1_id: int, view_2_id: int, num_features: int) -> int:
        cnt = C.c_int(0)
        self._check(self._L.osfm_match_pair_lowres(self._h, view_1_id, view_2_id,
                                                   C.c_size_t(num_features), C.byref(cnt)))
        return cnt.value

    # -- Matching::twoway_match<T> ---------------------------------------------------------
    def twoway_match(self, kind: int, view_1_id: int, view_2_id: int) -> Matching.Result:
        k = 0 if kind == KIND_SIFT_U8 else 1
        n1 = self._sizes[view_1_id][k]
        n2 = self._sizes[view_2_id][k]
        m12 = np.empty(max(n1, 1), np.int32)
        m21 = np.empty(max(n2, 1), np.int32)
        i32p = C.POINTER(C.c_int32)
        self._check(self._L.osfm_match_pair_twoway(self._h, kind, view_1_id, view_2_id,
                                                   m12.ctypes.data_as(i32p), m21.ctypes.data_as(i32p)))
        return Matching.Result(m12[:n1].copy(), m21[:n2].copy())

    # -- Matching::twoway_match<float> ------------------------------------------------------
    def twoway_match_f32(self, options: Matching.Options, set_1, set_2) -> Matching.Result:
        """The float descriptor path (matching.h:148-159 with T = float): bit-identical to
        the reference's SSE3 build.  ``set_1`` / ``set_2`` are n x descriptor_length floats."""
        dim = int(options.descriptor_length)
        a = np.ascontiguousarray(set_1, np.float32).reshape(-1, dim)
        b = np.ascontiguousarray(set_2, np.float32).reshape(-1, dim)
        m12 = np.empty(max(a.shape[0], 1), np.int32)
        m21 = np.empty(max(b.shape[0], 1), np.int32)
        i32p = C.POINTER(C.c_int32)
        self._check(self._L.osfm_match_twoway_f32(
            self._h, _ptr(a), a.shape[0], _ptr(b), b.shape[0], dim,
            C.c_float(options.lowe_ratio_threshold), C.c_float(options.distance_threshold),
            m12.ctypes.data_as(i32p), m21.ctypes.data_as(i32p)))
        return Matching.Result(m12[:a.shape[0]].copy(), m21[:b.shape[0]].copy())

    def debug_set_float_path(self, mode: int) -> None:
        """Float path: 0 = tensor-core filter first for large pairs, 1 = exact kernel only,
        2 = always the filter first.  Same match vectors in every mode."""
        self._check(self._L.osfm_match_debug_set_float_path(self._h, mode))

    def debug_float_filter(self, set_1, set_2, dim: int = 128):
        """What the float path's tensor-core filter sees: (s1, s2, j1), each n1 + n2 long."""
        a = np.ascontiguousarray(set_1, np.float32).reshape(-1, dim)
        b = np.ascontiguousarray(set_2, np.float32).reshape(-1, dim)
        n = a.shape[0] + b.shape[0]
        s1, s2, j1 = np.empty(n, np.float32), np.empty(n, np.float32), np.empty(n, np.int32)
        f32p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_int32)
        self._check(self._L.osfm_match_debug_float_filter(
            self._h, _ptr(a), a.shape[0], _ptr(b), b.shape[0], dim,
            s1.ctypes.data_as(f32p), s2.ctypes.data_as(f32p), j1.ctypes.data_as(i32p)))
        return s1, s2, j1

    # -- batched ------------------------------------------------------------------------------
    def match_pairs(self, pairs, out: Optional[np.ndarray] = None) -> tuple:
        """All pairs in one pass.  Returns (results, n_consistent): a list of
        Matching.Result (views into one dense buffer) and an int32 array of
        count_consistent_matches per pair.  ``out`` may be a caller-owned int32 buffer
        (e.g. page-locked memory) that receives the dense results."""
        pr = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        npairs = pr.shape[0]
        i32p, i64p = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        total = self._L.osfm_match_pairs_result_size(self._h, pr.ctypes.data_as(i32p), npairs)
        if total < 0:
            self._check(int(total))
        if out is not None:
            if out.dtype != np.int32 or out.size < total or not out.flags.c_contiguous:
                raise ValueError(f"out must be a contiguous int32 buffer of at least {total} elements")
            dense = out.reshape(-1)
        else:
            dense = np.empty(max(int(total), 1), np.int32)
        offsets = np.zeros(2 * npairs + 1, np.int64)
        counts = np.zeros(max(npairs, 1), np.int32)
        self._check(self._L.osfm_match_pairs(self._h, pr.ctypes.data_as(i32p), npairs,
                                             dense.ctypes.data_as(i32p), offsets.ctypes.data_as(i64p),
                                             counts.ctypes.data_as(i32p)))
        off = offsets.tolist()
        out_list = [Matching.Result(dense[off[2 * p]:off[2 * p + 1]], dense[off[2 * p + 1]:off[2 * p + 2]])
                    for p in range(npairs)]
        return out_list, counts[:npairs]

    def pairs_result_size(self, pairs) -> int:
        pr = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        total = self._L.osfm_match_pairs_result_size(self._h, pr.ctypes.data_as(C.POINTER(C.c_int32)), pr.shape[0])
        if total < 0:
            self._check(int(total))
        return int(total)

    def match_pairs_compact(self, pairs, out_ij) -> np.ndarray:
        """Device-resident batched matching (SIFT): surviving (i, j) index pairs of every
        image pair, ordered by i, written to the int32 torch tensor ``out_ij`` of shape
        [capacity, 2] on this GPU.  Returns the int64 list offsets (npairs + 1)."""
        pr = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        npairs = pr.shape[0]
        loff = np.zeros(npairs + 1, np.int64)
        rc = self._L.osfm_match_pairs_compact_device(
            self._h, pr.ctypes.data_as(C.POINTER(C.c_int32)), npairs,
            C.c_void_p(int(out_ij.data_ptr())), C.c_int64(int(out_ij.shape[0])),
            loff.ctypes.data_as(C.POINTER(C.c_int64)))
        self._check(rc)
        return loff

    def match_pairs_lists(self, pairs, out_ij: np.ndarray) -> np.ndarray:
        """Batched matching (SIFT) with a HOST result: the surviving (i, j) index pairs of every
        image pair, ordered by i (what bundler::Matching::two_view_matching builds from the
        Matching::Result, bundler_matching.cc:178-192), written to the int32 array ``out_ij`` of
        shape [capacity, 2] (pinned memory makes the copy asynchronous).  Returns the int64 list
        offsets (npairs + 1)."""
        pr = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        npairs = pr.shape[0]
        loff = np.zeros(npairs + 1, np.int64)
        assert out_ij.dtype == np.int32 and out_ij.flags.c_contiguous
        rc = self._L.osfm_match_pairs_compact(
            self._h, pr.ctypes.data_as(C.POINTER(C.c_int32)), npairs,
            out_ij.ctypes.data_as(C.c_void_p), C.c_int64(int(out_ij.shape[0])),
            loff.ctypes.data_as(C.POINTER(C.c_int64)))
        self._check(rc)
        return loff

    def two_view_candidates(self, pairs, opts: "TwoViewOptions" = None) -> list:
        """bundler::Matching::two_view_matching up to RANSAC (bundler_matching.cc:139-192) for a
        list of (view_1, view_2) pairs: pair rules, low-res gate, full match, match-count
        threshold, correspondence list.  Returns one ``(status, count, ij)`` per pair, ``ij`` an
        int32 array of shape [k, 2] (empty unless status == TWO_VIEW_OK)."""
        pr = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        npairs = pr.shape[0]
        o = _lib.TwoViewOptions()
        self._L.osfm_match_two_view_default_options(C.byref(o))
        if opts is not None:
            for name in ("use_lowres_matching", "num_lowres_features", "min_lowres_matches",
                         "min_feature_matches", "match_num_previous_frames"):
                setattr(o, name, int(getattr(opts, name)))
        cap = int(sum(min(sum(self._sizes[a]), sum(self._sizes[b])) for a, b in pr)) + 1
        ij = np.empty((cap, 2), np.int32)
        loff = np.zeros(npairs + 1, np.int64)
        status = np.zeros(npairs, np.int32)
        count = np.zeros(npairs, np.int32)
        i32p = C.POINTER(C.c_int32)
        self._check(self._L.osfm_match_two_view_candidates(
            self._h, C.byref(o), pr.ctypes.data_as(i32p), npairs, ij.ctypes.data_as(C.c_void_p), C.c_int64(cap),
            loff.ctypes.data_as(C.POINTER(C.c_int64)), status.ctypes.data_as(i32p), count.ctypes.data_as(i32p)))
        return [(int(status[p]), int(count[p]), ij[loff[p]:loff[p + 1]].copy()) for p in range(npairs)]

    def two_view_matching(self, pairs, positions, opts: "TwoViewOptions" = None) -> list:
        """bundler::Matching::compute's loop body for a list of pairs (bundler_matching.cc:92-220):
        candidates, RANSAC-F (samples from std::rand() in pair order), inlier threshold.
        ``positions``: [sum of features, 2] float32, view after view.  Returns one
        ``(status, count, ij)`` per pair; the TWO_VIEW_OK ones are the PairwiseMatching."""
        pr = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        npairs = pr.shape[0]
        opts = opts if opts is not None else TwoViewOptions()
        o = _lib.TwoViewOptions()
        self._L.osfm_match_two_view_default_options(C.byref(o))
        for name in ("use_lowres_matching", "num_lowres_features", "min_lowres_matches",
                     "min_feature_matches", "match_num_previous_frames"):
            setattr(o, name, int(getattr(opts, name)))
        r = _lib.RansacOptions()
        self._L.osfm_match_ransac_default_options(C.byref(r))
        r.max_iterations = int(opts.ransac_max_iterations)
        r.min_matching_inliers = int(opts.min_matching_inliers)
        r.threshold = float(opts.ransac_threshold)
        pos = np.ascontiguousarray(positions, np.float32).reshape(-1, 2)
        if len(pos) != sum(sum(sz) for sz in self._sizes):
            raise ValueError("positions must hold one (x, y) per feature of every view")
        cap = int(sum(min(sum(self._sizes[a]), sum(self._sizes[b])) for a, b in pr)) + 1
        ij = np.empty((cap, 2), np.int32)
        loff = np.zeros(npairs + 1, np.int64)
        status = np.zeros(npairs, np.int32)
        count = np.zeros(npairs, np.int32)
        i32p = C.POINTER(C.c_int32)
        self._check(self._L.osfm_match_two_view(
            self._h, C.byref(o), C.byref(r), pos.ctypes.data_as(C.POINTER(C.c_float)), pr.ctypes.data_as(i32p), npairs,
            ij.ctypes.data_as(C.c_void_p), C.c_int64(cap), loff.ctypes.data_as(C.POINTER(C.c_int64)),
            status.ctypes.data_as(i32p), count.ctypes.data_as(i32p)))
        return [(int(status[p]), int(count[p]), ij[loff[p]:loff[p + 1]].copy()) for p in range(npairs)]

    def tracks_compute(self, features_per_view, pair_views, offsets, ij) -> tuple:
        """sfm::bundler::Tracks::compute (bundler_tracks.cc:47-203) on the device: returns
        (track id of every feature or -1, view after view; number of tracks; number of
        components dropped for holding two features of one view).  Tracks are numbered in
        ascending order of their first feature."""
        f = np.ascontiguousarray(features_per_view, np.int32)
        pv = np.ascontiguousarray(np.asarray(pair_views, np.int32).reshape(-1, 2))
        off = np.ascontiguousarray(offsets, np.int64)
        m = np.ascontiguousarray(np.asarray(ij, np.int32).reshape(-1, 2))
        out = np.full(int(f.sum()), -9, np.int32)
        nt, nc = C.c_int32(0), C.c_int32(0)
        i32p = C.POINTER(C.c_int32)
        self._check(self._L.osfm_tracks_compute(
            self._h, len(f), f.ctypes.data_as(i32p), pv.ctypes.data_as(i32p), off.ctypes.data_as(C.POINTER(C.c_int64)),
            m.ctypes.data_as(i32p), len(pv), out.ctypes.data_as(i32p), C.byref(nt), C.byref(nc)))
        return out, int(nt.value), int(nc.value)

    def ransac_fundamental(self, features_per_view, positions, pair_views, offsets, ij, samples=None,
                           max_iterations: int = 1000, threshold: float = 0.0015) -> tuple:
        """sfm::RansacFundamental::estimate (ransac_fundamental.cc:26-105) for every pair at
        once.  samples=None: the library draws them from std::rand(), in pair order, as the
        reference does, overlapped with the device work.  Returns (inlier offsets [npairs + 1], inlier (i, j) lists,
        fundamental matrices [npairs, 3, 3])."""
        f = np.ascontiguousarray(features_per_view, np.int32)
        pos = np.ascontiguousarray(positions, np.float32).reshape(-1, 2)
        pv = np.ascontiguousarray(np.asarray(pair_views, np.int32).reshape(-1, 2))
        off = np.ascontiguousarray(offsets, np.int64)
        m = np.ascontiguousarray(np.asarray(ij, np.int32).reshape(-1, 2))
        if len(pos) != int(f.sum()):
            raise ValueError("positions must hold one (x, y) per feature")
        smp = None
        if samples is not None:
            smp = np.ascontiguousarray(samples, np.int32)
            if smp.size != len(pv) * max_iterations * 8:
                raise ValueError("samples must hold 8 indices per pair and iteration")
        out = np.empty((max(len(m), 1), 2), np.int32)
        out_off = np.zeros(len(pv) + 1, np.int64)
        F = np.zeros((len(pv), 3, 3), np.float64)
        i32p, i64p = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        self._check(self._L.osfm_ransac_fundamental(
            self._h, len(f), f.ctypes.data_as(i32p), pos.ctypes.data_as(C.POINTER(C.c_float)), pv.ctypes.data_as(i32p),
            off.ctypes.data_as(i64p), m.ctypes.data_as(i32p), len(pv),
            None if smp is None else smp.ctypes.data_as(i32p), int(max_iterations),
            float(threshold), out.ctypes.data_as(i32p), out_off.ctypes.data_as(i64p),
            F.ctypes.data_as(C.POINTER(C.c_double))))
        return out_off, out[:out_off[-1]], F

    def set_lookahead(self, max_pairs: int) -> None:
        """osfm_match_set_lookahead: pairwise_match / pairwise_match_lowres called pair by pair in
        the reference's order are served from batched passes over the next ``max_pairs`` pairs."""
        self._check(self._L.osfm_match_set_lookahead(self._h, int(max_pairs)))

    # -- introspection --------------------------------------------------------------------------
    def stats(self) -> dict:
        s = _lib.Stats()
        self._check(self._L.osfm_match_get_stats(self._h, C.byref(s)))
        return s.asdict()

    def debug_set_scan_mode(self, mode: int) -> None:
        self._check(self._L.osfm_match_debug_set_scan_mode(self._h, mode))

    def debug_set_both_directions(self, on: bool) -> None:
        """A/B switch: both directions of every pair through the filter pass (as the reference
        executes Matching::twoway_match) instead of one direction + the claimed rows of the other."""
        self._check(self._L.osfm_match_debug_set_both_directions(self._h, int(on)))

    def debug_set_exact_path(self, mode: int) -> None:
        """A/B switch for the rows that reach 2^16: 0 = CUDA-core inner products + warp-per-row
        replay when they fit the scratch buffer, 1 = always the tensor-core scan pass."""
        self._check(self._L.osfm_match_debug_set_exact_path(self._h, mode))

    def debug_dump_similarity(self, kind: int, view_q: int, view_c: int) -> np.ndarray:
        k = 0 if kind == KIND_SIFT_U8 else 1
        nq = self._sizes[view_q][k]
        nc = self._sizes[view_c][k]
        ld = 256 * ((nc + 255) // 256)
        out = np.zeros((nq, ld), np.int32)
        self._check(self._L.osfm_match_debug_dump_similarity(
            self._h, kind, view_q, view_c, out.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int64(out.size)))
        return out[:, :nc]

    def debug_dump_packed(self, kind: int, view_q: int, view_c: int) -> np.ndarray:
        """The similarity matrix as the filter epilogue reads it from tensor memory
        (tcgen05.ld .pack::16b): uint16, truncated to 16 bits, columns in order."""
        k = 0 if kind == KIND_SIFT_U8 else 1
        nq = self._sizes[view_q][k]
        nc = self._sizes[view_c][k]
        ld = 256 * ((nc + 255) // 256)
        out = np.zeros((nq, ld // 2), np.uint32)
        self._check(self._L.osfm_match_debug_dump_packed(
            self._h, kind, view_q, view_c, out.ctypes.data_as(C.c_void_p), C.c_int64(out.size)))
        return out.view(np.uint16)[:, :nc]   # little endian: low half = even column

    def debug_trace(self, pairs) -> np.ndarray:
        """clock64() stamps of CTA 0's pipeline events (see osfm_match_debug_trace)."""
        pr = np.ascontiguousarray(np.asarray(pairs, np.int32).reshape(-1, 2))
        out = np.zeros((20, 256, 4), np.int64)
        self._check(self._L.osfm_match_debug_trace(
            self._h, pr.ctypes.data_as(C.POINTER(C.c_int32)), len(pr),
            out.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int64(out.size)))
        return out


def ransac_draw_samples(offsets, max_iterations: int = 1000) -> np.ndarray:
    """The 8-match samples RansacFundamental::estimate_8_point would draw for these pairs, in
    order, from std::rand() (ransac_fundamental.cc:70-76): [npairs, max_iterations, 8],
    ascending inside a sample.  Consumes the C library's rand() sequence, as the reference."""
    off = np.ascontiguousarray(offsets, np.int64)
    npairs = len(off) - 1
    out = np.empty((npairs, max_iterations, 8), np.int32)
    rc = _lib.load().osfm_ransac_draw_samples(npairs, off.ctypes.data_as(C.POINTER(C.c_int64)), int(max_iterations),
                                              out.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc != 0:
        raise MatcherError(rc, "every pair needs at least 8 matches")
    return out
