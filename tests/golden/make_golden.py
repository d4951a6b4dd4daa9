"""Generates the golden fixtures in this directory from the REFERENCE itself
(oracle/_ref/libosfm_ref.so = the unmodified sources under /root/reference compiled in
place, see oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

Outputs (committed):
  real_pair.npz    quantised SIFT descriptors of the reference's own test images
                   src/cuda_sift/data/left.pgm / righ.pgm, produced by the reference's
                   sfm::Sift, with the reference ExhaustiveMatching results.
  real_triple.npz  the same for the three images the reference ships (left / righ / rimg_pts.pgm):
                   all 3 pairs, BASELINE config 1 restated (3-image set, exhaustive matching).
  tracks_golden.txt, pairwise_golden/   the OrthoSfM track files of a seeded track table, written by
                   the reference's own src/matching/matching_io.cpp (compiled unmodified).
  cases.npz        seeded synthetic and adversarial descriptor sets with the reference's
                   twoway_match / remove_inconsistent / count results (u8, s8, f32).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from orthosfm_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF = oracle.Reference()
ORA = oracle.Oracle()


def read_pgm(path):
    with open(path, "rb") as f:
        data = f.read()
    # P5 <w> <h> <maxval> then raw bytes; header tokens may be separated by any whitespace
    tokens, pos = [], 0
    while len(tokens) < 4:
        while data[pos:pos + 1].isspace():
            pos += 1
        if data[pos:pos + 1] == b"#":
            while data[pos:pos + 1] != b"\n":
                pos += 1
            continue
        start = pos
        while not data[pos:pos + 1].isspace():
            pos += 1
        tokens.append(data[start:pos])
    pos += 1
    assert tokens[0] == b"P5" and int(tokens[3]) == 255
    w, h = int(tokens[1]), int(tokens[2])
    return np.frombuffer(data, dtype=np.uint8, count=w * h, offset=pos).reshape(h, w)


def real_pair():
    d = "/root/reference/src/cuda_sift/data"
    fl = [REF.sift_gray8(read_pgm(os.path.join(d, n))) for n in ("left.pgm", "righ.pgm")]
    print("SIFT descriptors:", [x.shape for x in fl])
    empty = np.zeros((0, 64), np.float32)
    ex = REF.exhaustive([(fl[0], empty), (fl[1], empty)])
    m12, m21 = ex.pairwise_match(1, 0)
    lowres = ex.pairwise_match_lowres(1, 0, 500)
    q = [ORA.quantize_sift(x) for x in fl]
    # the quantised route through Matching::twoway_match must agree with the float route
    # through ExhaustiveMatching::init (convert_descriptor)
    t12, t21 = REF.twoway("u8", q[1], q[0], 0.8)
    f12, f21 = REF.remove_inconsistent(t12, t21)
    assert np.array_equal(f12, m12) and np.array_equal(f21, m21)
    print("consistent:", int((m12 >= 0).sum()), "lowres:", lowres)
    # a small float sample pins the quantiser itself (convert_descriptor)
    np.savez_compressed(os.path.join(HERE, "real_pair.npz"),
                        sift_0=q[0], sift_1=q[1], twoway_12=t12, twoway_21=t21,
                        match_12=m12, match_21=m21, lowres_500=np.int32(lowres),
                        float_sample=fl[0][:64], float_sample_q=q[0][:64])


def real_triple():
    """BASELINE config 1 restated: a 3-image set through the reference's own
    feature extractor and its ExhaustiveMatching, all 3 pairs in bundler::Matching::compute's
    order (view_1 > view_2).  The images are the three the reference ships
    (src/cuda_sift/data): the Suzanne renders of the testbench are an external download."""
    d = "/root/reference/src/cuda_sift/data"
    fl = [REF.sift_gray8(read_pgm(os.path.join(d, n))) for n in ("left.pgm", "righ.pgm", "rimg_pts.pgm")]
    print("SIFT descriptors:", [x.shape for x in fl])
    empty = np.zeros((0, 64), np.float32)
    ex = REF.exhaustive([(f, empty) for f in fl])
    q = [ORA.quantize_sift(x) for x in fl]
    out = {"sift_%d" % v: q[v] for v in range(3)}
    for v1 in range(1, 3):
        for v2 in range(v1):
            m12, m21 = ex.pairwise_match(v1, v2)
            out[f"match_{v1}{v2}_12"] = m12
            out[f"match_{v1}{v2}_21"] = m21
            out[f"lowres_{v1}{v2}"] = np.int32(ex.pairwise_match_lowres(v1, v2, 500))
            t12, t21 = REF.twoway("u8", q[v1], q[v2], 0.8)
            f12, f21 = REF.remove_inconsistent(t12, t21)
            assert np.array_equal(f12, m12) and np.array_equal(f21, m21)
            print((v1, v2), "consistent:", int((m12 >= 0).sum()), "lowres:", int(out[f"lowres_{v1}{v2}"]))
    np.savez_compressed(os.path.join(HERE, "real_triple.npz"), **out)


def track_files_case():
    """Deterministic input of the golden track files (the tests rebuild it): 5 views, tracks of
    2-5 features, positions and colours."""
    rng = np.random.default_rng(424242)
    feats = [40, 55, 33, 61, 48]
    n = sum(feats)
    base = np.concatenate([[0], np.cumsum(feats)])
    ids = np.full(n, -1, np.int32)
    free = [list(range(base[v], base[v + 1])) for v in range(len(feats))]
    members = []
    for _ in range(45):
        k = int(rng.integers(2, len(feats) + 1))
        views = sorted(rng.choice(len(feats), k, replace=False).tolist())
        if any(len(free[v]) == 0 for v in views):
            continue
        members.append([free[v].pop(int(rng.integers(len(free[v])))) for v in views])
    members.sort(key=lambda m: min(m))
    for t, m in enumerate(members):
        ids[m] = t
    pos = ((rng.random((n, 2)) - 0.5) * 0.97).astype(np.float32)
    pos[3] = (0.25, -0.125)            # values that print short ...
    pos[7] = (1e-7, 0.49999997)        # ... and in exponent form
    col = rng.integers(0, 256, (n, 3), dtype=np.uint8)
    return feats, ids, len(members), pos, col, 3000.0


def track_files():
    """tracks.txt and the AAA_BBB.txt pair files as the reference's OWN writer produces them
    (src/matching/matching_io.cpp:16-50, 97-140, compiled unmodified: oracle/_ref/libmatching_io_ref.so)."""
    import shutil
    io = oracle.ReferenceTrackIO()
    feats, ids, nt, pos, col, width = track_files_case()
    tracks = oracle.tracks_from_ids(feats, ids, pos, width, col)
    io.save_tracks(os.path.join(HERE, "tracks_golden.txt"), tracks)
    folder = os.path.join(HERE, "pairwise_golden")
    shutil.rmtree(folder, ignore_errors=True)
    os.makedirs(folder)
    io.save_pairwise(folder, tracks, len(feats))
    print("tracks_golden.txt:", os.path.getsize(os.path.join(HERE, "tracks_golden.txt")), "bytes;",
          len(os.listdir(folder)), "pair files")


def cases():
    rng = np.random.default_rng(20261018)
    out = {}

    def add(name, kind, a, b, ratio):
        t12, t21 = REF.twoway(kind, a, b, ratio)
        f12, f21 = REF.remove_inconsistent(t12, t21)
        out[name + ".a"] = a
        out[name + ".b"] = b
        out[name + ".ratio"] = np.float32(ratio)
        out[name + ".t12"] = t12
        out[name + ".t21"] = t21
        out[name + ".f12"] = f12
        out[name + ".f21"] = f21
        out[name + ".count"] = np.int32(REF.count_consistent(t12, t21))
        print(f"{name:28s} {kind} {a.shape[0]:5d} x {b.shape[0]:5d}  oneway {int((t12 >= 0).sum()):5d}"
              f"  consistent {int((f12 >= 0).sum()):5d}")

    v = synth.sift_views(1, 3, 700)
    add("u8.synth_700x700", "u8", v[0], v[1], 0.8)
    add("u8.synth_300x650", "u8", v[2][:300], v[1][:650], 0.8)
    add("u8.synth_ratio1", "u8", v[0][:257], v[1][:513], 1.0)
    # exact duplicates: ties (highest index wins), 0/0 ratio accepted, ip >= 65536 wraps
    dup = np.concatenate([v[0][:100], v[0][:100], v[1][:60]])
    add("u8.duplicates", "u8", v[0][:150], dup, 0.8)
    # arbitrary bytes: every inner product overflows the 16-bit lanes / stores
    add("u8.random_bytes", "u8", rng.integers(0, 256, (150, 128), dtype=np.uint8),
        rng.integers(0, 256, (210, 128), dtype=np.uint8), 0.8)
    # inner products straddling 65536
    mid_a = rng.integers(0, 64, (200, 128), dtype=np.uint8)
    mid_b = rng.integers(0, 64, (260, 128), dtype=np.uint8)
    mid_b[::4] = mid_a[:65]
    add("u8.straddle_65536", "u8", mid_a, mid_b, 0.8)
    add("u8.single_candidate", "u8", v[0][:40], v[1][:1], 0.8)
    add("u8.single_query", "u8", v[0][:1], v[1][:300], 0.8)
    add("u8.zeros", "u8", np.zeros((5, 128), np.uint8), np.zeros((7, 128), np.uint8), 0.8)
    # one lane carries everything: 16 x (64*64) = 65536 in lane 0 -> seen as 0
    lane = np.zeros((3, 128), np.uint8)
    lane[0, 0::8] = 64
    lane[1, 0::8] = 63
    lane[2, 1::8] = 64
    add("u8.lane_wrap", "u8", lane, lane, 0.8)

    sp = synth.surf_pool(1, 300)
    s0, s1 = synth.surf_view(1, 0, 500, sp), synth.surf_view(1, 1, 640, sp)
    add("s8.synth_500x640", "s8", s0, s1, 0.7)
    sa = rng.integers(-127, 128, (120, 64), dtype=np.int8)
    sb = rng.integers(-127, 128, (200, 64), dtype=np.int8)
    sa[::2] = sb[:120:2]
    add("s8.random_bytes", "s8", sa, sb, 0.7)
    add("s8.big_positive", "s8", rng.integers(60, 128, (90, 64), dtype=np.int8),
        rng.integers(60, 128, (130, 64), dtype=np.int8), 1.0)
    add("s8.all_negative", "s8", rng.integers(1, 100, (20, 64), dtype=np.int8),
        -rng.integers(1, 100, (30, 64), dtype=np.int8), 1.0)

    fa = np.abs(rng.standard_normal((200, 128))).astype(np.float32)
    fa /= np.linalg.norm(fa, axis=1, keepdims=True)
    fb = np.abs(rng.standard_normal((260, 128))).astype(np.float32)
    fb /= np.linalg.norm(fb, axis=1, keepdims=True)
    fb[:80] = fa[:80] + 0.01 * rng.standard_normal((80, 128)).astype(np.float32)
    add("f32.synth_200x260", "f32", fa, fb, 0.8)

    # combine_results
    s12 = np.array([1, -1, 0, 2], np.int32); s21 = np.array([2, 0, 3], np.int32)
    f12 = np.array([-1, 1], np.int32); f21 = np.array([-1, 1, -1], np.int32)
    c12, c21 = REF.combine_results(s12, s21, f12, f21)
    out["combine.s12"], out["combine.s21"], out["combine.f12"], out["combine.f21"] = s12, s21, f12, f21
    out["combine.c12"], out["combine.c21"] = c12, c21
    np.savez_compressed(os.path.join(HERE, "cases.npz"), **out)


if __name__ == "__main__":
    real_pair()
    real_triple()
    track_files()
    cases()
    for f in ("real_pair.npz", "real_triple.npz", "cases.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
