"""On-disk formats of the reference the matching result travels in: the MVE prebundle file
(src/mve/sfm/bundler_common.cc:56-190: feature positions / colors per view plus the pairwise
match lists), tracks.txt and the AAA_BBB.txt pair files (src/matching/matching_io.cpp:16-140).
Thin mirror of the C ABI (osfm_io_*); host code only."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def save_prebundle(path: str, features_per_view, positions, colors, pair_views, offsets, ij) -> None:
    """positions: [sum(features), 2] float32 or None; colors: [sum(features), 3] uint8 or None;
    pair p joins views pair_views[p] with the matches ij[offsets[p]:offsets[p + 1]]."""
    L = _lib.load()
    f = np.ascontiguousarray(features_per_view, np.int32)
    pv = np.ascontiguousarray(np.asarray(pair_views, np.int32).reshape(-1, 2))
    off = np.ascontiguousarray(offsets, np.int64)
    m = np.ascontiguousarray(np.asarray(ij, np.int32).reshape(-1, 2))
    pos = None if positions is None else np.ascontiguousarray(positions, np.float32)
    col = None if colors is None else np.ascontiguousarray(colors, np.uint8)
    i32p, i64p = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    rc = L.osfm_io_save_prebundle(
        path.encode(), len(f), f.ctypes.data_as(i32p),
        None if pos is None else pos.ctypes.data_as(C.POINTER(C.c_float)),
        None if col is None else col.ctypes.data_as(C.POINTER(C.c_uint8)),
        len(pv), pv.ctypes.data_as(i32p), off.ctypes.data_as(i64p), m.ctypes.data_as(i32p))
    if rc != 0:
        raise _lib.MatcherError(rc, f"cannot write {path}")


def load_prebundle(path: str) -> dict:
    L = _lib.load()
    h = C.c_void_p()
    nv, npairs = C.c_int(0), C.c_int(0)
    npos, ncol, nm = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    rc = L.osfm_io_load_prebundle(path.encode(), C.byref(h), C.byref(nv), C.byref(npos), C.byref(ncol),
                                  C.byref(npairs), C.byref(nm))
    if rc != 0:
        raise _lib.MatcherError(rc, f"cannot read {path}")
    try:
        out = {
            "n_positions": np.zeros(nv.value, np.int32), "n_colors": np.zeros(nv.value, np.int32),
            "positions": np.zeros((npos.value, 2), np.float32), "colors": np.zeros((ncol.value, 3), np.uint8),
            "pair_views": np.zeros((npairs.value, 2), np.int32), "offsets": np.zeros(npairs.value + 1, np.int64),
            "ij": np.zeros((nm.value, 2), np.int32),
        }
        i32p, i64p = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        L.osfm_io_prebundle_get(h, out["n_positions"].ctypes.data_as(i32p), out["n_colors"].ctypes.data_as(i32p),
                                out["positions"].ctypes.data_as(C.POINTER(C.c_float)),
                                out["colors"].ctypes.data_as(C.POINTER(C.c_uint8)),
                                out["pair_views"].ctypes.data_as(i32p), out["offsets"].ctypes.data_as(i64p),
                                out["ij"].ctypes.data_as(i32p))
        return out
    finally:
        L.osfm_io_prebundle_free(h)


def _tracks_args(features_per_view, track_of_feature, positions):
    f = np.ascontiguousarray(features_per_view, np.int32)
    t = np.ascontiguousarray(track_of_feature, np.int32)
    if len(t) != int(f.sum()):
        raise ValueError("track_of_feature must hold one id per feature")
    pos = None if positions is None else np.ascontiguousarray(positions, np.float32)
    return f, t, pos


def save_tracks(path: str, features_per_view, track_of_feature, num_tracks: int, positions, image_width: float,
                colors=None) -> None:
    """tracks.txt from Matching.tracks_compute's result (see osfm_io_save_tracks)."""
    L = _lib.load()
    f, t, pos = _tracks_args(features_per_view, track_of_feature, positions)
    col = None if colors is None else np.ascontiguousarray(colors, np.uint8)
    i32p = C.POINTER(C.c_int32)
    rc = L.osfm_io_save_tracks(path.encode(), len(f), f.ctypes.data_as(i32p), t.ctypes.data_as(i32p), int(num_tracks),
                               None if pos is None else pos.ctypes.data_as(C.POINTER(C.c_float)), float(image_width),
                               None if col is None else col.ctypes.data_as(C.POINTER(C.c_uint8)))
    if rc != 0:
        raise _lib.MatcherError(rc, f"cannot write {path}")


def save_pairwise_tracks(folder: str, features_per_view, track_of_feature, num_tracks: int, positions,
                         image_width: float) -> int:
    """The AAA_BBB.txt files; returns how many were written."""
    L = _lib.load()
    f, t, pos = _tracks_args(features_per_view, track_of_feature, positions)
    n = C.c_int(0)
    i32p = C.POINTER(C.c_int32)
    rc = L.osfm_io_save_pairwise_tracks(folder.encode(), len(f), f.ctypes.data_as(i32p), t.ctypes.data_as(i32p),
                                        int(num_tracks),
                                        None if pos is None else pos.ctypes.data_as(C.POINTER(C.c_float)),
                                        float(image_width), C.byref(n))
    if rc != 0:
        raise _lib.MatcherError(rc, f"cannot write into {folder}")
    return n.value


def load_tracks(path: str) -> dict:
    """Reads a tracks.txt: offsets [num_tracks + 1], ids [n, 3] (view, local id, global id),
    xy [n, 2], rgb [n, 3]."""
    L = _lib.load()
    h = C.c_void_p()
    nt, nf = C.c_int64(0), C.c_int64(0)
    rc = L.osfm_io_load_tracks(path.encode(), C.byref(h), C.byref(nt), C.byref(nf))
    if rc != 0:
        raise _lib.MatcherError(rc, f"cannot read {path}")
    try:
        out = {"offsets": np.zeros(nt.value + 1, np.int64), "ids": np.zeros((nf.value, 3), np.uint32),
               "xy": np.zeros((nf.value, 2), np.float32), "rgb": np.zeros((nf.value, 3), np.uint32)}
        u32p = C.POINTER(C.c_uint32)
        L.osfm_io_track_table_get(h, out["offsets"].ctypes.data_as(C.POINTER(C.c_int64)),
                                  out["ids"].ctypes.data_as(u32p), out["xy"].ctypes.data_as(C.POINTER(C.c_float)),
                                  out["rgb"].ctypes.data_as(u32p))
        return out
    finally:
        L.osfm_io_track_table_free(h)
