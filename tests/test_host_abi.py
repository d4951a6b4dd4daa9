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
    return sorted(set(re.findall(r"\b(osfm_(?:match|tracks)_\w+)\s*\(", text)))


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
    assert L.osfm_match_abi_version() == 1
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
