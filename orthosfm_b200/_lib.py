"""ctypes binding of libosfm_match.so (include/osfm_match.h).  Fails loudly when the
CUDA library is missing or no B200 is present -- there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libosfm_match.so")

OSFM_OK = 0
ERR_NAMES = {
    -1: "OSFM_ERR_INVALID_ARGUMENT", -2: "OSFM_ERR_NO_DEVICE", -3: "OSFM_ERR_CUDA",
    -4: "OSFM_ERR_STATE", -5: "OSFM_ERR_OUT_OF_MEMORY", -6: "OSFM_ERR_INTERNAL", -7: "OSFM_ERR_IO",
}
KIND_SIFT_U8 = 0
KIND_SURF_S8 = 1

# every symbol include/osfm_match.h declares
EXPORTS = [
    "osfm_match_default_config", "osfm_match_abi_version", "osfm_match_create",
    "osfm_match_destroy", "osfm_match_last_error", "osfm_match_begin",
    "osfm_match_set_view_f32", "osfm_match_set_view_q8", "osfm_match_commit",
    "osfm_match_commit_device", "osfm_match_num_views", "osfm_match_view_size",
    "osfm_match_pair", "osfm_match_pair_twoway", "osfm_match_twoway_f32", "osfm_match_pair_lowres",
    "osfm_match_pairs_result_size", "osfm_match_pairs", "osfm_match_pairs_compact_device", "osfm_match_pairs_compact",
    "osfm_match_two_view_default_options", "osfm_match_two_view_candidates", "osfm_tracks_compute",
    "osfm_io_save_prebundle", "osfm_io_load_prebundle", "osfm_io_prebundle_get", "osfm_io_prebundle_free",
    "osfm_io_save_tracks", "osfm_io_save_pairwise_tracks", "osfm_io_load_tracks", "osfm_io_track_table_get",
    "osfm_io_track_table_free",
    "osfm_match_begin_overlapped", "osfm_match_wait_staged", "osfm_match_set_views_q8",
    "osfm_ransac_draw_samples", "osfm_ransac_fundamental", "osfm_match_two_view",
    "osfm_match_ransac_default_options",
    "osfm_match_get_stats", "osfm_match_debug_set_scan_mode", "osfm_match_debug_dump_similarity", "osfm_match_debug_dump_packed", "osfm_match_debug_trace",
    "osfm_match_debug_set_both_directions", "osfm_match_debug_set_exact_path", "osfm_match_debug_set_float_path", "osfm_match_debug_float_filter", "osfm_match_set_lookahead",
    "osfm_match_create_multi", "osfm_match_num_devices",
]


class Config(C.Structure):
    _fields_ = [("device", C.c_int), ("sift_lowe_ratio", C.c_float),
                ("sift_distance_threshold", C.c_float), ("surf_lowe_ratio", C.c_float),
                ("surf_distance_threshold", C.c_float), ("reserved", C.c_int * 4)]


class TwoViewOptions(C.Structure):
    _fields_ = [("use_lowres_matching", C.c_int), ("num_lowres_features", C.c_int),
                ("min_lowres_matches", C.c_int), ("min_feature_matches", C.c_int),
                ("match_num_previous_frames", C.c_int), ("reserved", C.c_int * 3)]


class RansacOptions(C.Structure):
    _fields_ = [("max_iterations", C.c_int), ("min_matching_inliers", C.c_int), ("threshold", C.c_double),
                ("reserved", C.c_int * 4)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_int64), ("scan_items", C.c_int64),
                ("candidate_rows", C.c_int64), ("slow_rows", C.c_int64),
                ("self_check_failures", C.c_int64), ("last_scan_ms", C.c_double),
                ("last_total_ms", C.c_double), ("last_comparisons", C.c_int64),
                ("exact_rows", C.c_int64), ("last_scan_sm_cycles", C.c_int64), ("last_scan_ns", C.c_int64),
                ("claimed_rows", C.c_int64), ("last_phase_ms", C.c_double * 8),
                ("reverse_restricted_pairs", C.c_int64), ("reverse_candidate_rows", C.c_int64),
                ("exact_wide_rows", C.c_int64), ("float_filter_rows", C.c_int64), ("float_exact_rows", C.c_int64)]

    PHASES = ("filter", "classify", "resolve_fwd", "claim", "resolve_rev", "mutual", "compact")

    def asdict(self):
        d = {n: getattr(self, n) for n, _ in self._fields_ if n != "last_phase_ms"}
        d["last_phase_ms"] = {name: float(self.last_phase_ms[i]) for i, name in enumerate(self.PHASES)}
        return d


class MatcherError(RuntimeError):
    """Mirrors the std::runtime_error / std::invalid_argument the MVE code throws
    (src/mve/sfm/bundler_matching.cc:40,48,62)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {message}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m orthosfm_b200.csrc.build` "
            f"(or __graft_entry__.build()).  orthosfm_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, ip, i32p, i64p = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    L.osfm_match_default_config.argtypes = [C.POINTER(Config)]
    L.osfm_match_default_config.restype = None
    L.osfm_match_abi_version.restype = C.c_int
    L.osfm_match_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.osfm_match_create_multi.argtypes = [C.POINTER(Config), ip, C.c_int, C.POINTER(vp)]
    L.osfm_match_num_devices.argtypes = [vp]
    L.osfm_match_destroy.argtypes = [vp]
    L.osfm_match_destroy.restype = None
    L.osfm_match_last_error.argtypes = [vp]
    L.osfm_match_last_error.restype = C.c_char_p
    L.osfm_match_begin.argtypes = [vp, C.c_int]
    L.osfm_match_set_view_f32.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, vp, C.c_int, C.c_int]
    L.osfm_match_set_view_q8.argtypes = [vp, C.c_int, vp, C.c_int, vp, C.c_int]
    L.osfm_match_commit.argtypes = [vp]
    L.osfm_match_commit_device.argtypes = [vp, C.c_int, vp, i64p, i32p, C.c_int64, vp, i64p, i32p, C.c_int64]
    L.osfm_match_num_views.argtypes = [vp]
    L.osfm_match_view_size.argtypes = [vp, C.c_int, ip, ip]
    L.osfm_match_pair.argtypes = [vp, C.c_int, C.c_int, i32p, ip, i32p, ip, ip]
    L.osfm_match_pair_twoway.argtypes = [vp, C.c_int, C.c_int, C.c_int, i32p, i32p]
    L.osfm_match_pair_lowres.argtypes = [vp, C.c_int, C.c_int, C.c_size_t, ip]
    L.osfm_match_twoway_f32.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_int, C.c_float, C.c_float, i32p, i32p]
    L.osfm_match_pairs_result_size.argtypes = [vp, i32p, C.c_int]
    L.osfm_match_pairs_result_size.restype = C.c_int64
    L.osfm_match_pairs.argtypes = [vp, i32p, C.c_int, i32p, i64p, i32p]
    L.osfm_match_pairs_compact_device.argtypes = [vp, i32p, C.c_int, vp, C.c_int64, i64p]
    L.osfm_match_pairs_compact.argtypes = [vp, i32p, C.c_int, vp, C.c_int64, i64p]
    L.osfm_match_two_view_default_options.argtypes = [C.POINTER(TwoViewOptions)]
    L.osfm_match_two_view_default_options.restype = None
    L.osfm_match_two_view_candidates.argtypes = [vp, C.POINTER(TwoViewOptions), i32p, C.c_int, vp, C.c_int64,
                                                 i64p, i32p, i32p]
    L.osfm_tracks_compute.argtypes = [vp, C.c_int, i32p, i32p, i64p, i32p, C.c_int, i32p, i32p, i32p]
    f32p, u8p = C.POINTER(C.c_float), C.POINTER(C.c_uint8)
    L.osfm_io_save_prebundle.argtypes = [C.c_char_p, C.c_int, i32p, f32p, u8p, C.c_int, i32p, i64p, i32p]
    L.osfm_io_load_prebundle.argtypes = [C.c_char_p, C.POINTER(vp), ip, i64p, i64p, ip, i64p]
    L.osfm_io_prebundle_get.argtypes = [vp, i32p, i32p, f32p, u8p, i32p, i64p, i32p]
    L.osfm_io_prebundle_free.argtypes = [vp]
    L.osfm_io_prebundle_free.restype = None
    u32p = C.POINTER(C.c_uint32)
    L.osfm_io_save_tracks.argtypes = [C.c_char_p, C.c_int, i32p, i32p, C.c_int, f32p, C.c_double, u8p]
    L.osfm_io_save_pairwise_tracks.argtypes = [C.c_char_p, C.c_int, i32p, i32p, C.c_int, f32p, C.c_double, ip]
    L.osfm_io_load_tracks.argtypes = [C.c_char_p, C.POINTER(vp), i64p, i64p]
    L.osfm_io_track_table_get.argtypes = [vp, i64p, u32p, f32p, u32p]
    L.osfm_io_track_table_free.argtypes = [vp]
    L.osfm_io_track_table_free.restype = None
    f64p = C.POINTER(C.c_double)
    L.osfm_match_begin_overlapped.argtypes = [vp, C.c_int]
    L.osfm_match_wait_staged.argtypes = [vp]
    L.osfm_match_set_views_q8.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp), i32p, C.POINTER(vp), i32p]
    L.osfm_match_ransac_default_options.argtypes = [C.POINTER(RansacOptions)]
    L.osfm_match_ransac_default_options.restype = None
    L.osfm_match_two_view.argtypes = [vp, C.POINTER(TwoViewOptions), C.POINTER(RansacOptions), f32p, i32p, C.c_int,
                                      vp, C.c_int64, i64p, i32p, i32p]
    L.osfm_ransac_draw_samples.argtypes = [C.c_int, i64p, C.c_int, i32p]
    L.osfm_ransac_fundamental.argtypes = [vp, C.c_int, i32p, f32p, i32p, i64p, i32p, C.c_int, i32p, C.c_int,
                                          C.c_double, i32p, i64p, f64p]
    L.osfm_match_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.osfm_match_debug_set_scan_mode.argtypes = [vp, C.c_int]
    L.osfm_match_debug_set_both_directions.argtypes = [vp, C.c_int]
    L.osfm_match_debug_set_exact_path.argtypes = [vp, C.c_int]
    L.osfm_match_debug_set_float_path.argtypes = [vp, C.c_int]
    L.osfm_match_debug_float_filter.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_int, f32p, f32p, i32p]
    L.osfm_match_set_lookahead.argtypes = [vp, C.c_int]
    L.osfm_match_debug_dump_similarity.argtypes = [vp, C.c_int, C.c_int, C.c_int, i32p, C.c_int64]
    L.osfm_match_debug_trace.argtypes = [vp, i32p, C.c_int, i64p, C.c_int64]
    L.osfm_match_debug_dump_packed.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int64]
    _lib = L
    return L
