"""ctypes bindings for the CPU checker.  TEST INFRASTRUCTURE ONLY.

Two libraries live here:

* ``liboracle.so``  -- the C restatement of the reference matcher
  (``osfm_oracle.c``; each function cites the reference file:line it follows).
* ``_ref/libosfm_ref.so`` -- the UNMODIFIED reference sources compiled in place
  from ``/root/reference`` plus a thin ``extern "C"`` driver (``ref_driver.cc``).
  It can only be (re)built where ``/root/reference`` exists; the prebuilt file
  travels to the GPU box.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product
(``orthosfm_b200``) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "liboracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libosfm_ref.so")
REFERENCE_ROOT = "/root/reference"

FLT_MAX = float(np.finfo(np.float32).max)


def build(ref: bool = True) -> None:
    """Compile the restatement and, where the reference tree exists, the reference."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    if ref and os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "mve", "sfm")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


def have_ref() -> bool:
    return os.path.exists(_REF_SO)


class NNResult(C.Structure):
    _fields_ = [("dist_1st_best", C.c_float), ("dist_2nd_best", C.c_float),
                ("index_1st_best", C.c_int), ("index_2nd_best", C.c_int)]

    def astuple(self):
        return (self.dist_1st_best, self.dist_2nd_best,
                self.index_1st_best, self.index_2nd_best)


def _ptr(a: np.ndarray, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def _c(a, dtype) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


_KIND = {
    "u8": (np.uint8, C.c_uint8, 128),
    "s8": (np.int8, C.c_int8, 64),
    "f32": (np.float32, C.c_float, 128),
}


class _Impl:
    """Shared front-end over either library (prefix 'osfm_oracle_' or 'osfm_ref_')."""

    def __init__(self, path: str, prefix: str):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} missing: run `make -C oracle` (and `make -C oracle ref` "
                f"where /root/reference exists)")
        self.lib = C.CDLL(path)
        self.prefix = prefix
        self.is_ref = prefix == "osfm_ref_"
        f = self._f
        f("num_threads").restype = C.c_int
        f("count_consistent").restype = C.c_int

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    def num_threads(self) -> int:
        return int(self._f("num_threads")())

    def set_num_threads(self, n: int) -> None:
        self._f("set_num_threads")(C.c_int(int(n)))

    def use_all_cores(self) -> int:
        """torchrun exports OMP_NUM_THREADS=1; use every core this process may run on."""
        try:
            n = len(os.sched_getaffinity(0))
        except AttributeError:
            n = os.cpu_count() or 1
        self.set_num_threads(n)
        return self.num_threads()

    # -- NearestNeighbor<T>::find ------------------------------------
    def nn(self, kind: str, query, elements, sse3_order: bool = True):
        npdt, ct, _ = _KIND[kind]
        q = _c(query, npdt)
        e = _c(elements, npdt).reshape(-1, q.shape[-1])
        r = NNResult()
        args = [_ptr(q, ct), _ptr(e, ct), C.c_int(e.shape[0]), C.c_int(q.shape[-1])]
        if kind == "f32" and not self.is_ref:
            args.append(C.c_int(1 if sse3_order else 0))
        self._f("nn_" + kind)(*args, C.byref(r))
        return r.astuple()

    # -- Matching::twoway_match<T> -----------------------------------
    def twoway(self, kind: str, set_1, set_2, ratio: float, dist: float = FLT_MAX,
               sse3_order: bool = True):
        npdt, ct, dim = _KIND[kind]
        a = _c(set_1, npdt)
        b = _c(set_2, npdt)
        if a.ndim == 2:
            dim = a.shape[1]
        elif b.ndim == 2:
            dim = b.shape[1]
        a = a.reshape(-1, dim)
        b = b.reshape(-1, dim)
        n1, n2 = a.shape[0], b.shape[0]
        m12 = np.full(max(n1, 1), -7, dtype=np.int32)
        m21 = np.full(max(n2, 1), -7, dtype=np.int32)
        args = [_ptr(a, ct), C.c_int(n1), _ptr(b, ct), C.c_int(n2), C.c_int(dim),
                C.c_float(ratio), C.c_float(dist)]
        if kind == "f32" and not self.is_ref:
            args.append(C.c_int(1 if sse3_order else 0))
        self._f("twoway_" + kind)(*args, _ptr(m12, C.c_int), _ptr(m21, C.c_int))
        return m12[:n1].copy(), m21[:n2].copy()

    # -- filters -----------------------------------------------------
    def remove_inconsistent(self, m12, m21):
        a = _c(m12, np.int32).copy()
        b = _c(m21, np.int32).copy()
        self._f("remove_inconsistent")(_ptr(a, C.c_int), C.c_int(a.size),
                                       _ptr(b, C.c_int), C.c_int(b.size))
        return a, b

    def count_consistent(self, m12, m21) -> int:
        a = _c(m12, np.int32)
        b = _c(m21, np.int32)
        return int(self._f("count_consistent")(_ptr(a, C.c_int), C.c_int(a.size),
                                               _ptr(b, C.c_int), C.c_int(b.size)))

    def combine_results(self, sift_12, sift_21, surf_12, surf_21):
        s12, s21 = _c(sift_12, np.int32), _c(sift_21, np.int32)
        f12, f21 = _c(surf_12, np.int32), _c(surf_21, np.int32)
        o12 = np.full(max(s12.size + f12.size, 1), -7, dtype=np.int32)
        o21 = np.full(max(s21.size + f21.size, 1), -7, dtype=np.int32)
        self._f("combine_results")(
            _ptr(s12, C.c_int), C.c_int(s12.size), _ptr(s21, C.c_int), C.c_int(s21.size),
            _ptr(f12, C.c_int), C.c_int(f12.size), _ptr(f21, C.c_int), C.c_int(f21.size),
            _ptr(o12, C.c_int), _ptr(o21, C.c_int))
        return o12[:s12.size + f12.size].copy(), o21[:s21.size + f21.size].copy()

    def match_filtered(self, kind: str, set_1, set_2, ratio: float):
        """twoway_match + remove_inconsistent_matches (the reference's timed unit)."""
        m12, m21 = self.twoway(kind, set_1, set_2, ratio)
        return self.remove_inconsistent(m12, m21)


class Oracle(_Impl):
    """The C restatement (liboracle.so)."""

    def __init__(self):
        super().__init__(_ORACLE_SO, "osfm_oracle_")
        self.lib.osfm_oracle_pairwise_match_lowres.restype = C.c_int

    def quantize_sift(self, desc) -> np.ndarray:
        d = _c(desc, np.float32).reshape(-1, 128)
        out = np.empty(d.shape, dtype=np.uint8)
        self.lib.osfm_oracle_quantize_sift(_ptr(d, C.c_float), C.c_int(d.shape[0]),
                                           _ptr(out, C.c_uint8))
        return out

    def quantize_surf(self, desc) -> np.ndarray:
        d = _c(desc, np.float32).reshape(-1, 64)
        out = np.empty(d.shape, dtype=np.int8)
        self.lib.osfm_oracle_quantize_surf(_ptr(d, C.c_float), C.c_int(d.shape[0]),
                                           _ptr(out, C.c_int8))
        return out

    @staticmethod
    def _sets(sift_1, sift_2, surf_1, surf_2):
        s1 = _c(sift_1, np.uint8).reshape(-1, 128)
        s2 = _c(sift_2, np.uint8).reshape(-1, 128)
        f1 = _c(surf_1, np.int8).reshape(-1, 64)
        f2 = _c(surf_2, np.int8).reshape(-1, 64)
        return s1, s2, f1, f2

    def pairwise_match(self, sift_1, sift_2, surf_1, surf_2):
        """ExhaustiveMatching::pairwise_match on already-quantised descriptors."""
        s1, s2, f1, f2 = self._sets(sift_1, sift_2, surf_1, surf_2)
        n12 = (s1.shape[0] if s1.shape[0] else 0) + (f1.shape[0] if f1.shape[0] else 0)
        n21 = (s2.shape[0] if s1.shape[0] else 0) + (f2.shape[0] if f1.shape[0] else 0)
        m12 = np.full(max(n12, 1), -7, dtype=np.int32)
        m21 = np.full(max(n21, 1), -7, dtype=np.int32)
        self.lib.osfm_oracle_pairwise_match(
            _ptr(s1, C.c_uint8), C.c_int(s1.shape[0]), _ptr(s2, C.c_uint8), C.c_int(s2.shape[0]),
            _ptr(f1, C.c_int8), C.c_int(f1.shape[0]), _ptr(f2, C.c_int8), C.c_int(f2.shape[0]),
            _ptr(m12, C.c_int), _ptr(m21, C.c_int))
        return m12[:n12].copy(), m21[:n21].copy()

    def pairwise_match_lowres(self, sift_1, sift_2, surf_1, surf_2, num_features: int) -> int:
        s1, s2, f1, f2 = self._sets(sift_1, sift_2, surf_1, surf_2)
        return int(self.lib.osfm_oracle_pairwise_match_lowres(
            _ptr(s1, C.c_uint8), C.c_int(s1.shape[0]), _ptr(s2, C.c_uint8), C.c_int(s2.shape[0]),
            _ptr(f1, C.c_int8), C.c_int(f1.shape[0]), _ptr(f2, C.c_int8), C.c_int(f2.shape[0]),
            C.c_int(num_features)))


def two_view_candidates(impl: "Oracle", sift, surf, pairs, use_lowres_matching=False, num_lowres_features=500,
                        min_lowres_matches=5, min_feature_matches=24, match_num_previous_frames=0):
    """bundler::Matching::compute's pair rules and two_view_matching up to RANSAC
    (src/mve/sfm/bundler_matching.cc:92-100, 139-192), restated over the oracle's
    pairwise_match / pairwise_match_lowres.  Returns (status, count, ij) per pair with the
    status codes of include/osfm_match.h."""
    out = []
    e8 = np.zeros((0, 128), np.uint8)
    e64 = np.zeros((0, 64), np.int8)
    for v1, v2 in pairs:
        s1 = sift[v1] if sift is not None and sift[v1] is not None else e8
        s2 = sift[v2] if sift is not None and sift[v2] is not None else e8
        f1 = surf[v1] if surf is not None and surf[v1] is not None else e64
        f2 = surf[v2] if surf is not None and surf[v2] is not None else e64
        n1, n2 = len(s1) + len(f1), len(s2) + len(f2)
        none = np.zeros((0, 2), np.int32)
        if match_num_previous_frames != 0 and v2 + match_num_previous_frames < v1:
            out.append((1, 0, none)); continue
        if n1 == 0 or n2 == 0:
            out.append((1, 0, none)); continue
        if use_lowres_matching and n1 * n2 > 1000000:
            k = impl.pairwise_match_lowres(s1, s2, f1, f2, num_lowres_features)
            if k < min_lowres_matches:
                out.append((2, k, none)); continue
        m12, m21 = impl.pairwise_match(s1, s2, f1, f2)
        k = impl.count_consistent(m12, m21)
        if k < max(8, min_feature_matches):
            out.append((3, k, none)); continue
        i = np.nonzero(m12 >= 0)[0]
        out.append((0, k, np.stack([i, m12[i]], axis=1).astype(np.int32)))
    return out


def canonical_track_ids(track_ids: np.ndarray) -> np.ndarray:
    """Relabels per-feature track ids (-1 = none) in order of first appearance, so that two
    labelings of the same partition become identical arrays."""
    t = np.asarray(track_ids, np.int64)
    out = np.full(t.shape, -1, np.int32)
    used = t >= 0
    if used.any():
        _, first = np.unique(t[used], return_index=True)
        order = np.argsort(first)                       # old labels by first appearance
        remap = np.empty(order.size, np.int64)
        remap[order] = np.arange(order.size)
        labels = np.unique(t[used])
        out[used] = remap[np.searchsorted(labels, t[used])]
    return out


def tracks_compute(features, pair_views, offsets, ij):
    """sfm::bundler::Tracks::compute (src/mve/sfm/bundler_tracks.cc:47-203) restated: sequential
    track-id propagation over the match lists in order, unify into the larger track, then
    tracks with two features of one view are dropped.  Returns Viewport::track_ids of all
    views, concatenated (ids as the reference numbers them), and the number of tracks."""
    features = np.asarray(features, np.int64)
    base = np.concatenate([[0], np.cumsum(features)])
    tid = np.full(int(base[-1]), -1, np.int64)
    tracks = []                                             # lists of global feature ids
    ij = np.asarray(ij).reshape(-1, 2)
    for p in range(len(pair_views)):
        v1, v2 = pair_views[p]
        for k in range(offsets[p], offsets[p + 1]):
            a, b = base[v1] + ij[k, 0], base[v2] + ij[k, 1]
            t1, t2 = tid[a], tid[b]
            if t1 == -1 and t2 == -1:
                tid[a] = tid[b] = len(tracks)
                tracks.append([a, b])
            elif t1 == -1:
                tid[a] = t2
                tracks[t2].append(a)
            elif t2 == -1:
                tid[b] = t1
                tracks[t1].append(b)
            elif t1 != t2:
                if len(tracks[t1]) < len(tracks[t2]):       # unify_tracks, :23-43
                    t1, t2 = t2, t1
                for g in tracks[t2]:
                    tid[g] = t1
                tracks[t1].extend(tracks[t2])
                tracks[t2] = []
    view_of = np.repeat(np.arange(len(features)), features)
    keep = []
    for t in tracks:                                        # remove_invalid_tracks, :148-203
        keep.append(bool(t) and len(set(view_of[t].tolist())) == len(t))
    remap = np.full(len(tracks), -1, np.int64)
    remap[np.nonzero(keep)[0]] = np.arange(int(np.sum(keep)))
    out = np.where(tid >= 0, remap[np.maximum(tid, 0)], -1).astype(np.int32)
    return out, int(np.sum(keep))


# ---- tracks.txt / AAA_BBB.txt -------------------------------------------------------------------
# The functions below restate orthosfm's writer / reader line by line.  Parity is pinned: the
# reference's own src/matching/matching_io.cpp is compiled, unmodified, into
# _ref/libmatching_io_ref.so against stand-in third-party headers (oracle/stubs/, class
# ReferenceTrackIO below); the restatement, the product's osfm_io_* and the committed golden files
# (tests/golden/tracks_golden.txt, tests/golden/pairwise_golden/) are held against it.

def _g(x) -> str:
    """operator<<(ostream&, float) with the default precision 6 -- printf's %g."""
    return "%g" % float(np.float32(x))


def tracks_from_ids(features, track_ids, positions, image_width, colors=None):
    """The std::vector<Track> of the MVE bridge (src/matching/matching_mve.cpp:455-466) for
    the given Viewport::track_ids: per track a list of
    (viewID, localFeatureID, globalFeatureID, x, y, r, g, b).  Features inside a track in
    ascending (view, feature) -- the reference's order inside a track follows its walk over
    the pairs and is not part of the format (loadTracksFromFile takes any)."""
    track_ids = np.asarray(track_ids)
    tracks = [[] for _ in range(int(track_ids.max()) + 1 if len(track_ids) else 0)]
    at = 0
    for v, n in enumerate(features):
        for f in range(int(n)):
            t = int(track_ids[at])
            if t >= 0:
                x = np.float32(float(image_width) * (float(np.float32(positions[at][0])) + 0.5))
                y = np.float32(float(image_width) * (float(np.float32(positions[at][1])) + 0.5))
                rgb = (0, 0, 0) if colors is None else tuple(int(c) for c in colors[at])
                tracks[t].append((v, f, 32768 * v + f, x, y) + rgb)
            at += 1
    return tracks


def save_tracks_text(tracks) -> str:
    """orthosfm::saveTracksToFile (src/matching/matching_io.cpp:16-50)."""
    out = []
    for tr in tracks:
        line = "%d;" % len(tr)
        line += ";".join("%d;%d;%d;%s;%s;%d;%d;%d" % (f[0], f[1], f[2], _g(f[3]), _g(f[4]), f[5], f[6], f[7])
                         for f in tr)
        out.append(line + "\n")
    return "".join(out)


def load_tracks_text(text: str):
    """orthosfm::loadTracksFromFile (src/matching/matching_io.cpp:52-95)."""
    tracks = []
    for line in text.splitlines():
        tok = line.split(";")
        n = int(tok[0])
        tr = []
        for k in range(n):
            q = tok[1 + 8 * k: 9 + 8 * k]
            tr.append((int(q[0]), int(q[1]), int(q[2]), np.float32(q[3]), np.float32(q[4]), int(q[5]), int(q[6]),
                       int(q[7])))
        tracks.append(tr)
    return tracks


def save_pairwise_tracks_text(tracks, view_ids) -> dict:
    """orthosfm::saveTracksToPairwiseFiles (src/matching/matching_io.cpp:97-140) with
    filterTracksToAvailableCameras(ids, tracks, true, false) (src/util/common.cpp:85-130):
    {file name: content}."""
    files = {}
    for a in range(len(view_ids)):
        for b in range(a + 1, len(view_ids)):
            ids = [view_ids[a], view_ids[b]]
            filtered = []
            for tr in tracks:
                cur = [f for f in tr if f[0] in ids]
                if len(cur) == len(ids):
                    filtered.append(cur)
            if not filtered:
                continue
            text = ""
            for tr in filtered:
                for k, vid in enumerate(ids):
                    for f in tr:
                        if f[0] == vid:
                            text += _g(f[3]) + " " + _g(f[4]) + (" " if k == 0 else "\n")
            files["%03d_%03d.txt" % (ids[0], ids[1])] = text
    return files


class RansacHostCheck:
    """orthosfm_b200/csrc/ransac_math.cuh (the device arithmetic of RANSAC-F) compiled for the
    host (oracle/ransac_hostcheck.cc): lets CPU-only tests hold it against the reference."""

    def __init__(self):
        self.lib = C.CDLL(os.path.join(_HERE, "_ref", "libransac_hostcheck.so"))
        self.lib.osfm_hostcheck_sampson.restype = C.c_double
        self.lib.osfm_hostcheck_ransac.restype = C.c_int

    @staticmethod
    def available() -> bool:
        return os.path.exists(os.path.join(_HERE, "_ref", "libransac_hostcheck.so"))

    def fundamental(self, p1, p2) -> np.ndarray:
        p1, p2 = _c(p1, np.float64), _c(p2, np.float64)
        F = np.zeros(9, np.float64)
        self.lib.osfm_hostcheck_fundamental(_ptr(p1, C.c_double), _ptr(p2, C.c_double), _ptr(F, C.c_double))
        return F

    def fundamental_staged(self, p1, p2):
        """The device's route: bidiagonalise / iterate with the fixed-point stop in a strided
        buffer / finish / rank 2.  Returns (F, trips of the 9 x 9 loop)."""
        p1, p2 = _c(p1, np.float64), _c(p2, np.float64)
        F = np.zeros(9, np.float64)
        it = C.c_int(0)
        self.lib.osfm_hostcheck_fundamental_staged(_ptr(p1, C.c_double), _ptr(p2, C.c_double), _ptr(F, C.c_double),
                                                   C.byref(it))
        return F, it.value

    def sampson(self, F, match) -> float:
        F, match = _c(F, np.float64), _c(match, np.float64)
        return self.lib.osfm_hostcheck_sampson(_ptr(F, C.c_double), _ptr(match, C.c_double))

    def svd(self, a):
        a = _c(a, np.float64)
        n = a.shape[0]
        s, v, u = np.zeros(n), np.zeros((n, n)), np.zeros((n, n))
        if n == 9:
            self.lib.osfm_hostcheck_svd9(_ptr(a, C.c_double), _ptr(s, C.c_double), _ptr(v, C.c_double))
            return s, v
        self.lib.osfm_hostcheck_svd3(_ptr(a, C.c_double), _ptr(u, C.c_double), _ptr(s, C.c_double),
                                     _ptr(v, C.c_double))
        return u, s, v

    def ransac(self, matches_xy, samples, threshold: float = 0.0015):
        m = _c(matches_xy, np.float64)
        smp = _c(samples, np.int32).reshape(-1, 8)
        inl = np.zeros(len(m), np.int32)
        F = np.zeros(9, np.float64)
        n = self.lib.osfm_hostcheck_ransac(_ptr(m, C.c_double), C.c_int(len(m)), _ptr(smp, C.c_int),
                                           C.c_int(len(smp)), C.c_double(threshold), _ptr(inl, C.c_int),
                                           _ptr(F, C.c_double))
        return inl[:n].copy(), F


def srand(seed: int) -> None:
    """std::srand on the C library every library in this process shares."""
    C.CDLL(None).srand(C.c_uint(seed))


class Reference(_Impl):
    """The reference itself, compiled from /root/reference (oracle/_ref)."""

    def __init__(self):
        super().__init__(_REF_SO, "osfm_ref_")
        L = self.lib
        L.osfm_ref_match_pairs_u8.restype = C.c_long
        L.osfm_ref_exhaustive_create.restype = C.c_void_p
        L.osfm_ref_exhaustive_pairwise_match_lowres.restype = C.c_int
        L.osfm_ref_sift_gray8.restype = C.c_int

    def match_pairs_u8(self, views, pairs, ratio: float = 0.8):
        """twoway_match + remove_inconsistent over a pair list, OpenMP over pairs
        (bundler_matching.cc:74).  Returns per-pair consistent-match counts."""
        views = [_c(v, np.uint8).reshape(-1, 128) for v in views]
        ptrs = (C.POINTER(C.c_uint8) * len(views))(*[_ptr(v, C.c_uint8) for v in views])
        sizes = np.array([v.shape[0] for v in views], dtype=np.int32)
        pr = _c(pairs, np.int32).reshape(-1, 2)
        counts = np.zeros(max(pr.shape[0], 1), dtype=np.int32)
        self.lib.osfm_ref_match_pairs_u8(ptrs, _ptr(sizes, C.c_int), C.c_int(len(views)),
                                         _ptr(pr, C.c_int), C.c_int(pr.shape[0]),
                                         C.c_float(ratio), _ptr(counts, C.c_int))
        return counts[:pr.shape[0]].copy()

    def match_pairs_u8_digest(self, views, pairs, ratio: float = 0.8):
        """Like match_pairs_u8, plus a 64-bit FNV-1a digest of every pair's correspondence list
        ((i, j) int32 pairs in ascending i): returns (counts, digests)."""
        views = [_c(v, np.uint8).reshape(-1, 128) for v in views]
        ptrs = (C.POINTER(C.c_uint8) * len(views))(*[_ptr(v, C.c_uint8) for v in views])
        sizes = np.array([v.shape[0] for v in views], dtype=np.int32)
        pr = _c(pairs, np.int32).reshape(-1, 2)
        counts = np.zeros(max(pr.shape[0], 1), dtype=np.int32)
        digest = np.zeros(max(pr.shape[0], 1), dtype=np.uint64)
        self.lib.osfm_ref_match_pairs_u8_digest.restype = C.c_long
        self.lib.osfm_ref_match_pairs_u8_digest(ptrs, _ptr(sizes, C.c_int), C.c_int(len(views)),
                                                _ptr(pr, C.c_int), C.c_int(pr.shape[0]), C.c_float(ratio),
                                                _ptr(counts, C.c_int), _ptr(digest, C.c_ulonglong))
        return counts[:pr.shape[0]], digest[:pr.shape[0]]

    def match_large_pair_u8(self, set_1, set_2, ratio: float = 0.8):
        """One large pair on all cores (the reference's oneway_match on chunks of the query rows,
        then remove_inconsistent_matches): returns (count, digest, m12, m21)."""
        a = _c(set_1, np.uint8).reshape(-1, 128)
        b = _c(set_2, np.uint8).reshape(-1, 128)
        m12 = np.full(max(a.shape[0], 1), -7, dtype=np.int32)
        m21 = np.full(max(b.shape[0], 1), -7, dtype=np.int32)
        digest = C.c_ulonglong(0)
        f = self.lib.osfm_ref_match_large_pair_u8
        f.restype = C.c_long
        n = f(_ptr(a, C.c_uint8), C.c_int(a.shape[0]), _ptr(b, C.c_uint8), C.c_int(b.shape[0]), C.c_float(ratio),
              _ptr(m12, C.c_int), _ptr(m21, C.c_int), C.byref(digest))
        return int(n), int(digest.value), m12[:a.shape[0]], m21[:b.shape[0]]

    def tracks_compute(self, features, pair_views, offsets, ij):
        """The reference's own Tracks::compute through ref_driver.cc."""
        features = _c(features, np.int32)
        pv = _c(np.asarray(pair_views).reshape(-1, 2), np.int32)
        off = _c(offsets, np.int64)
        ijc = _c(np.asarray(ij).reshape(-1, 2), np.int32)
        out = np.full(int(features.sum()), -9, np.int32)
        f = self.lib.osfm_ref_tracks_compute
        f.restype = C.c_int
        n = f(C.c_int(len(features)), _ptr(features, C.c_int), C.c_int(len(pv)), _ptr(pv, C.c_int),
              off.ctypes.data_as(C.POINTER(C.c_longlong)), _ptr(ijc, C.c_int), _ptr(out, C.c_int))
        return out, int(n)

    def save_prebundle(self, path, features, positions, colors, pair_views, offsets, ij) -> None:
        """The reference's own save_prebundle_to_file (bundler_common.cc:180-188)."""
        features = _c(features, np.int32)
        pv = _c(np.asarray(pair_views).reshape(-1, 2), np.int32)
        off = _c(offsets, np.int64)
        ijc = _c(np.asarray(ij).reshape(-1, 2), np.int32)
        pos, col = _c(positions, np.float32), _c(colors, np.uint8)
        f = self.lib.osfm_ref_save_prebundle
        f.restype = C.c_int
        rc = f(path.encode(), C.c_int(len(features)), _ptr(features, C.c_int), _ptr(pos, C.c_float),
               _ptr(col, C.c_ubyte), C.c_int(len(pv)), _ptr(pv, C.c_int),
               off.ctypes.data_as(C.POINTER(C.c_longlong)), _ptr(ijc, C.c_int))
        if rc != 0:
            raise OSError(f"reference could not write {path}")

    def load_prebundle_digest(self, path):
        """Counts and checksums of a prebundle file as the reference's load_prebundle_from_file
        reads it (bundler_common.cc:110-178, 190-198)."""
        counts = np.zeros(4, np.int64)
        sums = np.zeros(3, np.float64)
        f = self.lib.osfm_ref_load_prebundle_digest
        f.restype = C.c_int
        rc = f(path.encode(), counts.ctypes.data_as(C.POINTER(C.c_longlong)), sums.ctypes.data_as(C.POINTER(C.c_double)))
        if rc != 0:
            raise OSError(f"reference could not read {path}")
        return counts, sums

    def fundamental(self, p1, p2) -> np.ndarray:
        """fundamental_8_point + enforce_fundamental_constraints (fundamental.cc:78-126) on
        eight correspondences ([8, 2] each)."""
        p1, p2 = _c(p1, np.float64), _c(p2, np.float64)
        F = np.zeros(9, np.float64)
        self.lib.osfm_ref_fundamental(_ptr(p1, C.c_double), _ptr(p2, C.c_double), _ptr(F, C.c_double))
        return F

    def sampson(self, F, match) -> float:
        F, match = _c(F, np.float64), _c(match, np.float64)
        f = self.lib.osfm_ref_sampson
        f.restype = C.c_double
        return f(_ptr(F, C.c_double), _ptr(match, C.c_double))

    def svd(self, a):
        """math::matrix_svd (matrix_svd.h) of a 9 x 9 (-> s, V) or 3 x 3 (-> U, s, V) matrix."""
        a = _c(a, np.float64)
        n = a.shape[0]
        s, v, u = np.zeros(n), np.zeros((n, n)), np.zeros((n, n))
        if n == 9:
            self.lib.osfm_ref_svd9(_ptr(a, C.c_double), _ptr(s, C.c_double), _ptr(v, C.c_double))
            return s, v
        self.lib.osfm_ref_svd3(_ptr(a, C.c_double), _ptr(u, C.c_double), _ptr(s, C.c_double), _ptr(v, C.c_double))
        return u, s, v

    def ransac(self, matches_xy, iterations: int = 1000, threshold: float = 0.0015, seed: int = -1):
        """RansacFundamental::estimate (ransac_fundamental.cc:26-60) on [n, 4] matches
        (x1 y1 x2 y2); seed >= 0 calls std::srand(seed) first.  Returns (inlier indices, F)."""
        m = _c(matches_xy, np.float64)
        inl = np.zeros(len(m), np.int32)
        F = np.zeros(9, np.float64)
        f = self.lib.osfm_ref_ransac
        f.restype = C.c_int
        n = f(_ptr(m, C.c_double), C.c_int(len(m)), C.c_int(iterations), C.c_double(threshold), C.c_int(seed),
              _ptr(inl, C.c_int), _ptr(F, C.c_double))
        return inl[:n].copy(), F

    def exhaustive(self, views_float):
        """views_float: list of (sift n x 128 float32, surf n x 64 float32)."""
        return _RefExhaustive(self, views_float)

    def sift_gray8(self, pixels: np.ndarray, max_desc: int = 1 << 16) -> np.ndarray:
        img = _c(pixels, np.uint8)
        h, w = img.shape
        out = np.zeros((max_desc, 128), dtype=np.float32)
        n = self.lib.osfm_ref_sift_gray8(_ptr(img, C.c_uint8), C.c_int(w), C.c_int(h),
                                         _ptr(out, C.c_float), C.c_int(max_desc))
        return out[:n].copy()


class _RefExhaustive:
    def __init__(self, ref: Reference, views_float):
        self.L = ref.lib
        self.sizes = []
        self.h = C.c_void_p(self.L.osfm_ref_exhaustive_create(C.c_int(len(views_float))))
        for v, (sift, surf) in enumerate(views_float):
            s = _c(sift, np.float32).reshape(-1, 128)
            f = _c(surf, np.float32).reshape(-1, 64)
            self.sizes.append(s.shape[0] + f.shape[0])
            self.L.osfm_ref_exhaustive_set_view(self.h, C.c_int(v),
                                                _ptr(s, C.c_float), C.c_int(s.shape[0]),
                                                _ptr(f, C.c_float), C.c_int(f.shape[0]))
        self.L.osfm_ref_exhaustive_init(self.h)

    def pairwise_match(self, v1: int, v2: int):
        m12 = np.full(self.sizes[v1] + 1, -7, dtype=np.int32)
        m21 = np.full(self.sizes[v2] + 1, -7, dtype=np.int32)
        n12, n21 = C.c_int(0), C.c_int(0)
        self.L.osfm_ref_exhaustive_pairwise_match(self.h, C.c_int(v1), C.c_int(v2),
                                                  _ptr(m12, C.c_int), C.byref(n12),
                                                  _ptr(m21, C.c_int), C.byref(n21))
        return m12[:n12.value].copy(), m21[:n21.value].copy()

    def bundler_compute(self, positions, use_lowres_matching=False, num_lowres_features=500, min_lowres_matches=5,
                        min_feature_matches=24, min_matching_inliers=12, match_num_previous_frames=0,
                        ransac_max_iterations=1000, ransac_threshold=0.0015, seed=-1):
        """bundler::Matching::init + compute (bundler_matching.cc:45-133) over these viewports
        with the given positions; consumes the descriptors (init frees them), so call once.
        Returns [(view_1, view_2, ij)] in the order compute() accepted the pairs."""
        pos = _c(positions, np.float32).reshape(-1, 2)
        assert len(pos) == sum(self.sizes)
        opts = np.array([int(use_lowres_matching), num_lowres_features, min_lowres_matches, min_feature_matches,
                         min_matching_inliers, match_num_previous_frames, ransac_max_iterations], np.int32)
        nv = len(self.sizes)
        cap_pairs = nv * (nv - 1) // 2 + 1
        cap_ij = cap_pairs * max(self.sizes) + 1
        pairs = np.zeros((cap_pairs, 2), np.int32)
        off = np.zeros(cap_pairs + 1, np.int64)
        ij = np.zeros((cap_ij, 2), np.int32)
        f = self.L.osfm_ref_bundler_compute
        f.restype = C.c_int
        n = f(self.h, _ptr(pos, C.c_float), _ptr(opts, C.c_int), C.c_double(ransac_threshold), C.c_int(seed),
              _ptr(pairs, C.c_int), C.c_int(cap_pairs), off.ctypes.data_as(C.POINTER(C.c_longlong)),
              _ptr(ij, C.c_int), C.c_longlong(cap_ij))
        assert n >= 0
        return [(int(pairs[p, 0]), int(pairs[p, 1]), ij[off[p]:off[p + 1]].copy()) for p in range(n)]

    def pairwise_match_lowres(self, v1: int, v2: int, num_features: int) -> int:
        return int(self.L.osfm_ref_exhaustive_pairwise_match_lowres(
            self.h, C.c_int(v1), C.c_int(v2), C.c_int(num_features)))

    def __del__(self):
        try:
            self.L.osfm_ref_exhaustive_destroy(self.h)
        except Exception:
            pass


_CUDASIFT_SO = os.path.join(_HERE, "_ref", "libcudasift_ref.so")


def have_cudasift() -> bool:
    return os.path.exists(_CUDASIFT_SO)


class CudaSiftReference:
    """The reference's own GPU matcher -- CudaSift's MatchSiftData / FindMaxCorr10
    (/root/reference/src/cuda_sift/matching.cu:1090-1206, 301-397), the unmodified sources
    compiled for sm_100 (oracle/Makefile, cudasift_driver.cu).  One-way FP32 nearest neighbour
    with score and ambiguity: a baseline to time on the same GPU, not a parity target."""

    def __init__(self):
        self.lib = C.CDLL(_CUDASIFT_SO)
        self.lib.osfm_cudasift_match.restype = C.c_int

    def match(self, set_1, set_2, reps: int = 5):
        """Returns (mean ms, min ms, match index per row of set_1, score, ambiguity)."""
        a = _c(set_1, np.uint8).reshape(-1, 128)
        b = _c(set_2, np.uint8).reshape(-1, 128)
        ms = np.zeros(2, np.float64)
        match = np.zeros(a.shape[0], np.int32)
        score = np.zeros(a.shape[0], np.float32)
        amb = np.zeros(a.shape[0], np.float32)
        rc = self.lib.osfm_cudasift_match(_ptr(a, C.c_uint8), C.c_int(a.shape[0]), _ptr(b, C.c_uint8),
                                          C.c_int(b.shape[0]), C.c_int(reps), _ptr(ms, C.c_double),
                                          _ptr(match, C.c_int32), _ptr(score, C.c_float), _ptr(amb, C.c_float))
        if rc != 0:
            raise RuntimeError(f"osfm_cudasift_match failed ({rc})")
        return float(ms[0]), float(ms[1]), match, score, amb


def list_digest(ij: np.ndarray) -> int:
    """FNV-1a (64 bit) over an (i, j) int32 correspondence list, as osfm_ref_match_pairs_u8_digest."""
    h = np.uint64(1469598103934665603)
    prime = np.uint64(1099511628211)
    with np.errstate(over="ignore"):
        for b in np.ascontiguousarray(ij, np.int32).view(np.uint8).reshape(-1):
            h = (h ^ np.uint64(b)) * prime
    return int(h)


_MATCHING_IO_SO = os.path.join(_HERE, "_ref", "libmatching_io_ref.so")


def have_ref_io() -> bool:
    return os.path.exists(_MATCHING_IO_SO)


class ReferenceTrackIO:
    """The reference's own orthosfm::saveTracksToFile / loadTracksFromFile /
    saveTracksToPairwiseFiles (src/matching/matching_io.cpp:16-140), compiled unmodified
    (oracle/Makefile, matching_io_driver.cc).  Tracks are lists of
    (viewID, localFeatureID, globalFeatureID, x, y, r, g, b) as tracks_from_ids builds them."""

    def __init__(self):
        self.lib = C.CDLL(_MATCHING_IO_SO)

    @staticmethod
    def _table(tracks):
        offsets = np.concatenate([[0], np.cumsum([len(t) for t in tracks])]).astype(np.int64)
        flat = [f for t in tracks for f in t]
        ids = np.array([[f[0], f[1], f[2]] for f in flat], np.uint32).reshape(-1, 3)
        xy = np.array([[f[3], f[4]] for f in flat], np.float32).reshape(-1, 2)
        rgb = np.array([[f[5], f[6], f[7]] for f in flat], np.uint32).reshape(-1, 3)
        return offsets, np.ascontiguousarray(ids), np.ascontiguousarray(xy), np.ascontiguousarray(rgb)

    def save_tracks(self, path: str, tracks) -> None:
        off, ids, xy, rgb = self._table(tracks)
        self._quiet(lambda: self.lib.osfm_refio_save_tracks(
            path.encode(), C.c_int64(len(tracks)), _ptr(off, C.c_int64), _ptr(ids, C.c_uint32), _ptr(xy, C.c_float),
            _ptr(rgb, C.c_uint32)))

    def save_pairwise(self, folder: str, tracks, num_views: int) -> None:
        off, ids, xy, rgb = self._table(tracks)
        self.lib.osfm_refio_save_pairwise(folder.encode(), C.c_int(num_views), C.c_int64(len(tracks)),
                                          _ptr(off, C.c_int64), _ptr(ids, C.c_uint32), _ptr(xy, C.c_float),
                                          _ptr(rgb, C.c_uint32))

    def load_tracks(self, path: str):
        nt, nf = C.c_int64(0), C.c_int64(0)
        self.lib.osfm_refio_load_tracks(path.encode(), C.byref(nt), C.byref(nf), None, None, None, None)
        off = np.zeros(nt.value + 1, np.int64)
        ids = np.zeros((max(nf.value, 1), 3), np.uint32)
        xy = np.zeros((max(nf.value, 1), 2), np.float32)
        rgb = np.zeros((max(nf.value, 1), 3), np.uint32)
        self.lib.osfm_refio_load_tracks(path.encode(), None, None, _ptr(off, C.c_int64), _ptr(ids, C.c_uint32),
                                        _ptr(xy, C.c_float), _ptr(rgb, C.c_uint32))
        return {"offsets": off, "ids": ids[:nf.value], "xy": xy[:nf.value], "rgb": rgb[:nf.value]}

    @staticmethod
    def _quiet(fn):
        """saveTracksToFile announces itself on std::cout: keep test output clean."""
        import sys
        sys.stdout.flush()
        saved = os.dup(1)
        devnull = os.open(os.devnull, os.O_WRONLY)
        try:
            os.dup2(devnull, 1)
            return fn()
        finally:
            C.CDLL(None).fflush(None)
            os.dup2(saved, 1)
            os.close(devnull)
            os.close(saved)
