"""CPU tests of the host side: the C-ABI library loads and exports every symbol the
header declares, the package fails loudly without a GPU, and the host-side helpers
(pair enumeration, synthetic generator, list utilities) behave like the reference."""
import math
import os
import re
import subprocess

import numpy as np
import pytest

from orthosfm_b200 import Matching, _lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "osfm_match.h")


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(osfm_(?:match|tracks|io|ransac)_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib_path():
    if not os.path.exists(_lib.LIB_PATH):
        from orthosfm_b200.csrc import build
        build.build()
    return _lib.LIB_PATH


def test_header_symbols_all_exported(lib_path):
    declared = _declared_symbols()
    assert len(declared) >= 20
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib_path], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [s for s in declared if s not in exported]
    assert not missing, f"declared in include/osfm_match.h but not exported: {missing}"
    assert sorted(_lib.EXPORTS) == declared


def test_library_loads_and_reports_abi(lib_path):
    L = _lib.load()
    assert L.osfm_match_abi_version() == 2
    cfg = _lib.Config()
    L.osfm_match_default_config(cfg)
    assert cfg.device == 0
    assert cfg.sift_lowe_ratio == pytest.approx(0.8) and cfg.surf_lowe_ratio == pytest.approx(0.7)
    assert cfg.sift_distance_threshold == pytest.approx(np.finfo(np.float32).max)


def test_library_contains_blackwell_tensor_core_code(lib_path):
    """The scan kernel must be real tcgen05/TMA code, not a CUDA-core fallback."""
    sass = subprocess.run(["cuobjdump", "-sass", lib_path], capture_output=True, text=True)
    if sass.returncode != 0:
        pytest.skip("cuobjdump not available")
    assert "UTCIMMA" in sass.stdout      # tcgen05.mma kind::i8
    assert "LDTM" in sass.stdout         # tcgen05.ld
    assert "UTMALDG" in sass.stdout      # TMA tensor load
    assert "sm_100a" in sass.stdout


def test_no_cpu_fallback_without_gpu(lib_path):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    from orthosfm_b200 import ExhaustiveMatching, MatcherError
    with pytest.raises(MatcherError) as ei:
        ExhaustiveMatching()
    assert ei.value.code == -2  # OSFM_ERR_NO_DEVICE


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "orthosfm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "osfm_oracle" not in text and "liboracle" not in text, f


def test_all_pairs_is_the_reference_enumeration():
    # bundler_matching.cc:92-93
    n = 23
    pairs = synth.all_pairs(n)
    assert pairs.shape == (n * (n - 1) // 2, 2)
    for i, (v1, v2) in enumerate(pairs):
        r1 = int(0.5 + math.sqrt(0.25 + 2.0 * i))
        r2 = i - r1 * (r1 - 1) // 2
        assert (v1, v2) == (r1, r2) and v1 > v2


def test_synth_is_deterministic_and_sift_like():
    a = synth.sift_views(2, 3, 256)
    b = synth.sift_views(2, 3, 256)
    for x, y in zip(a, b):
        assert x.dtype == np.uint8 and x.shape == (256, 128) and np.array_equal(x, y)
    n2 = (a[0].astype(np.int64) ** 2).sum(axis=1)
    assert abs(n2.mean() - 65025) < 600  # unit norm * 255
    s = synth.surf_view(2, 0, 128)
    assert s.dtype == np.int8 and abs((s.astype(np.int64) ** 2).sum(axis=1).mean() - 16129) < 300


def test_count_consistent_matches_mirror(ora):
    rng = np.random.default_rng(3)
    m12 = rng.integers(-1, 50, 80).astype(np.int32)
    m21 = rng.integers(-1, 80, 50).astype(np.int32)
    assert Matching.count_consistent_matches(Matching.Result(m12, m21)) == ora.count_consistent(m12, m21)


def _prebundle_case(seed=0):
    rng = np.random.default_rng(seed)
    feats = np.array([40, 0, 25, 33], np.int32)
    n = int(feats.sum())
    pos = rng.random((n, 2), dtype=np.float32) * 2 - 1
    col = rng.integers(0, 256, (n, 3), dtype=np.uint8)
    pairs = [(2, 0), (3, 0), (3, 2)]
    lists = [np.stack([np.sort(rng.choice(feats[a], 12, replace=False)), rng.choice(feats[b], 12, replace=False)], 1)
             for a, b in pairs]
    lists[1] = lists[1][:0]          # a pair with an empty list
    off = np.concatenate([[0], np.cumsum([len(x) for x in lists])]).astype(np.int64)
    return feats, pos, col, pairs, off, np.concatenate(lists).astype(np.int32)


def test_prebundle_round_trip(tmp_path):
    """osfm_io_save_prebundle / osfm_io_load_prebundle (the MVE prebundle file,
    bundler_common.cc:56-190): what is written is read back."""
    from orthosfm_b200 import io as osio
    feats, pos, col, pairs, off, ij = _prebundle_case()
    path = str(tmp_path / "prebundle.sfm")
    osio.save_prebundle(path, feats, pos, col, pairs, off, ij)
    d = osio.load_prebundle(path)
    assert np.array_equal(d["n_positions"], feats) and np.array_equal(d["n_colors"], feats)
    assert np.array_equal(d["positions"], pos) and np.array_equal(d["colors"], col)
    assert np.array_equal(d["pair_views"], np.asarray(pairs)) and np.array_equal(d["offsets"], off)
    assert np.array_equal(d["ij"], ij)
    assert open(path, "rb").read(14) == b"MVE_PREBUNDLE\n"
    with pytest.raises(_lib.MatcherError):
        osio.load_prebundle(str(tmp_path / "missing.sfm"))
    bad = tmp_path / "bad.sfm"
    bad.write_bytes(b"NOT_A_PREBUNDLE_FILE")
    with pytest.raises(_lib.MatcherError):
        osio.load_prebundle(str(bad))


def test_prebundle_is_byte_identical_to_the_reference_writer(tmp_path):
    """The same data through the reference's own save_prebundle_to_file gives the same bytes,
    and the reference's load_prebundle_from_file reads our file."""
    import oracle
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    from orthosfm_b200 import io as osio
    ref = oracle.Reference()
    feats, pos, col, pairs, off, ij = _prebundle_case(3)
    ours, theirs = str(tmp_path / "ours.sfm"), str(tmp_path / "theirs.sfm")
    osio.save_prebundle(ours, feats, pos, col, pairs, off, ij)
    ref.save_prebundle(theirs, feats, pos, col, pairs, off, ij)
    assert open(ours, "rb").read() == open(theirs, "rb").read()
    counts, sums = ref.load_prebundle_digest(ours)
    assert counts.tolist() == [len(feats), int(feats.sum()), len(pairs), len(ij)]
    assert np.isclose(sums[0], float(pos[:, 0].astype(np.float64).sum() + 2 * pos[:, 1].astype(np.float64).sum()))
    assert sums[1] == float((col.astype(np.float64) * [1, 3, 5]).sum())
    pv = np.asarray(pairs, np.float64)
    assert sums[2] == float((1000 * pv[:, 0] + 7 * pv[:, 1]).sum() + (31.0 * ij[:, 0] + 17.0 * ij[:, 1]).sum())


def _tracks_case(seed=0):
    rng = np.random.default_rng(seed)
    feats = np.array([30, 0, 22, 27, 19], np.int32)
    n = int(feats.sum())
    view_of = np.repeat(np.arange(len(feats)), feats)
    ids = np.full(n, -1, np.int32)
    num_tracks = 0
    # tracks of 2..4 views, at most one feature per view, ids ascending by first feature
    free = [list(np.flatnonzero(view_of == v)) for v in range(len(feats))]
    members = []
    for _ in range(14):
        views = sorted(rng.choice([0, 2, 3, 4], size=int(rng.integers(2, 5)), replace=False))
        if any(not free[v] for v in views):
            continue
        members.append([free[v].pop(int(rng.integers(len(free[v])))) for v in views])
    members.sort(key=lambda m: min(m))
    for t, m in enumerate(members):
        ids[m] = t
    num_tracks = len(members)
    pos = (rng.random((n, 2), dtype=np.float32) - 0.5) * np.float32(0.97)
    col = rng.integers(0, 256, (n, 3), dtype=np.uint8)
    return feats, ids, num_tracks, pos, col


@pytest.mark.parametrize("with_colors", [False, True])
def test_tracks_file_matches_the_restated_writer_and_reads_back(tmp_path, with_colors):
    """osfm_io_save_tracks / osfm_io_load_tracks against the restatement of
    orthosfm::saveTracksToFile / loadTracksFromFile (matching_io.cpp:16-95)."""
    import oracle
    from orthosfm_b200 import io as osio
    feats, ids, nt, pos, col = _tracks_case(1)
    width = 3000.0
    path = str(tmp_path / "tracks.txt")
    osio.save_tracks(path, feats, ids, nt, pos, width, col if with_colors else None)
    want = oracle.tracks_from_ids(feats, ids, pos, width, col if with_colors else None)
    text = open(path).read()
    assert text == oracle.save_tracks_text(want)
    back = osio.load_tracks(path)
    parsed = oracle.load_tracks_text(text)
    assert back["offsets"].tolist() == np.concatenate([[0], np.cumsum([len(t) for t in parsed])]).tolist()
    flat = [f for t in parsed for f in t]
    assert back["ids"].tolist() == [list(f[:3]) for f in flat]
    assert np.array_equal(back["xy"], np.array([[f[3], f[4]] for f in flat], np.float32))
    assert back["rgb"].tolist() == [list(f[5:]) for f in flat]
    with pytest.raises(_lib.MatcherError):
        osio.load_tracks(str(tmp_path / "missing.txt"))
    with pytest.raises(_lib.MatcherError):
        osio.save_tracks(path, feats, np.where(ids >= 0, ids + nt, ids), nt, pos, width)     # id out of range


def test_pairwise_track_files_match_the_restated_writer(tmp_path):
    """osfm_io_save_pairwise_tracks against the restatement of saveTracksToPairwiseFiles
    (matching_io.cpp:97-140)."""
    import oracle
    from orthosfm_b200 import io as osio
    feats, ids, nt, pos, _ = _tracks_case(2)
    width = 1234.0
    n = osio.save_pairwise_tracks(str(tmp_path), feats, ids, nt, pos, width)
    want = oracle.save_pairwise_tracks_text(oracle.tracks_from_ids(feats, ids, pos, width), list(range(len(feats))))
    assert n == len(want) and n > 0
    assert sorted(os.listdir(tmp_path)) == sorted(want)
    for name, text in want.items():
        assert open(tmp_path / name).read() == text


def _golden_track_case():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_tracks", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    src = open(spec.origin).read()
    # only the seeded input builder (the module itself loads the compiled reference at import)
    start = src.index("def track_files_case():")
    end = src.index("def track_files():")
    ns = {"np": np}
    exec(src[start:end], ns)
    return ns["track_files_case"]()


def test_track_files_equal_the_reference_writers_golden_files(tmp_path):
    """osfm_io_save_tracks / osfm_io_save_pairwise_tracks against files written by the
    reference's OWN writer (src/matching/matching_io.cpp:16-50, 97-140, compiled unmodified into
    oracle/_ref/libmatching_io_ref.so; tests/golden/make_golden.py::track_files): same bytes."""
    import oracle
    from orthosfm_b200 import io as osio
    golden = os.path.join(ROOT, "tests", "golden")
    feats, ids, nt, pos, col, width = _golden_track_case()
    path = str(tmp_path / "tracks.txt")
    osio.save_tracks(path, feats, ids, nt, pos, width, col)
    assert open(path, "rb").read() == open(os.path.join(golden, "tracks_golden.txt"), "rb").read()
    # the restatement in oracle/ agrees with the reference's file as well
    want = oracle.tracks_from_ids(feats, ids, pos, width, col)
    assert oracle.save_tracks_text(want) == open(os.path.join(golden, "tracks_golden.txt")).read()
    folder = tmp_path / "pairs"
    folder.mkdir()
    n = osio.save_pairwise_tracks(str(folder), feats, ids, nt, pos, width)
    names = sorted(os.listdir(os.path.join(golden, "pairwise_golden")))
    assert n == len(names) and sorted(os.listdir(folder)) == names
    for name in names:
        assert open(folder / name, "rb").read() == open(os.path.join(golden, "pairwise_golden", name), "rb").read(), name
    # and the reference's file reads back through osfm_io_load_tracks
    back = osio.load_tracks(os.path.join(golden, "tracks_golden.txt"))
    flat = [f for t in want for f in t]
    assert back["ids"].tolist() == [list(f[:3]) for f in flat]
    assert np.array_equal(back["xy"], np.array([[np.float32("%g" % f[3]), np.float32("%g" % f[4])] for f in flat], np.float32))
    assert back["rgb"].tolist() == [list(f[5:]) for f in flat]


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_track_files_equal_the_compiled_reference_writer(tmp_path, seed):
    """Random track tables through the reference's own writer and reader (where oracle/_ref was
    built) and through osfm_io_*: same files, same parsed tables."""
    import oracle
    from orthosfm_b200 import io as osio
    if not oracle.have_ref_io():
        pytest.skip("oracle/_ref/libmatching_io_ref.so not built (needs /root/reference)")
    ref = oracle.ReferenceTrackIO()
    feats, ids, nt, pos, col = _tracks_case(seed)
    width = [3000.0, 1234.0, 640.0][seed - 1]
    tracks = oracle.tracks_from_ids(feats, ids, pos, width, col)
    ours, theirs = str(tmp_path / "ours.txt"), str(tmp_path / "theirs.txt")
    osio.save_tracks(ours, feats, ids, nt, pos, width, col)
    ref.save_tracks(theirs, tracks)
    assert open(ours, "rb").read() == open(theirs, "rb").read()
    a, b = osio.load_tracks(theirs), ref.load_tracks(ours)
    for key in ("offsets", "ids", "xy", "rgb"):
        assert np.array_equal(a[key], b[key]), key
    fo, ft = tmp_path / "po", tmp_path / "pt"
    fo.mkdir()
    ft.mkdir()
    osio.save_pairwise_tracks(str(fo), feats, ids, nt, pos, width)
    ref.save_pairwise(str(ft), tracks, len(feats))
    assert sorted(os.listdir(fo)) == sorted(os.listdir(ft)) and len(os.listdir(fo)) > 0
    for name in os.listdir(ft):
        assert open(fo / name, "rb").read() == open(ft / name, "rb").read(), name
