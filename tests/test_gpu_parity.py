"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI,
against the CPU oracle on the same inputs.  Integer work: the bar is bit-exact."""
import numpy as np
import pytest

from orthosfm_b200 import (ExhaustiveMatching, FeatureSet, KIND_SIFT_U8, KIND_SURF_S8,
                           MatcherError, Matching, MatchingBase, Viewport, synth)
from orthosfm_b200 import TwoViewOptions, TWO_VIEW_OK, TWO_VIEW_SKIPPED, TWO_VIEW_LOWRES_REJECTED, TWO_VIEW_TOO_FEW_MATCHES  # noqa: E402,F811

pytestmark = pytest.mark.gpu

from test_oracle import CASE_NAMES  # noqa: E402

U8_CASES = [n for n in CASE_NAMES if n.startswith("u8.")]
S8_CASES = [n for n in CASE_NAMES if n.startswith("s8.")]


def vps(sift=None, surf=None):
    n = len(sift) if sift is not None else len(surf)
    return [Viewport(FeatureSet(sift_descriptors=None if sift is None else sift[i],
                                surf_descriptors=None if surf is None else surf[i])) for i in range(n)]


def matcher(sift=None, surf=None, opts=None):
    m = ExhaustiveMatching(opts)
    m.init(vps(sift, surf))
    return m


def assert_clean(m):
    st = m.stats()
    assert st["self_check_failures"] == 0
    assert st["kernel_launches"] > 0


# ------------------------------------------------------------------ tensor-core product

@pytest.mark.parametrize("n1,n2", [(1, 1), (128, 256), (100, 300), (129, 257), (300, 1000)])
def test_similarity_matrix_is_exact(n1, n2):
    rng = np.random.default_rng(n1 * 1000 + n2)
    a = rng.integers(0, 256, (n1, 128), dtype=np.uint8)
    b = rng.integers(0, 256, (n2, 128), dtype=np.uint8)
    with matcher([a, b]) as m:
        s = m.debug_dump_similarity(KIND_SIFT_U8, 0, 1)
    assert np.array_equal(s, a.astype(np.int64) @ b.astype(np.int64).T)


def test_similarity_matrix_signed_is_exact():
    rng = np.random.default_rng(5)
    a = rng.integers(-127, 128, (200, 64), dtype=np.int8)
    b = rng.integers(-127, 128, (333, 64), dtype=np.int8)
    with matcher(surf=[a, b]) as m:
        s = m.debug_dump_similarity(KIND_SURF_S8, 0, 1)
    assert np.array_equal(s, a.astype(np.int64) @ b.astype(np.int64).T)


@pytest.mark.parametrize("n1,n2", [(5, 7), (128, 256), (200, 1000)])
def test_packed_tmem_layout(n1, n2):
    """The filter epilogue reads the accumulator with tcgen05.ld .pack::16b: it must see the
    low 16 bits of every similarity, columns in order (low half-word = even column)."""
    rng = np.random.default_rng(n1 + n2)
    a = rng.integers(0, 256, (n1, 128), dtype=np.uint8)     # products far beyond 16 bits
    b = rng.integers(0, 256, (n2, 128), dtype=np.uint8)
    with matcher([a, b]) as m:
        s = m.debug_dump_packed(KIND_SIFT_U8, 0, 1)
    assert np.array_equal(s, ((a.astype(np.int64) @ b.astype(np.int64).T) & 0xffff).astype(np.uint16))


def test_packed_tmem_layout_signed():
    rng = np.random.default_rng(11)
    a = rng.integers(-127, 128, (150, 64), dtype=np.int8)
    b = rng.integers(-127, 128, (300, 64), dtype=np.int8)
    with matcher(surf=[a, b]) as m:
        s = m.debug_dump_packed(KIND_SURF_S8, 0, 1)
    assert np.array_equal(s, ((a.astype(np.int64) @ b.astype(np.int64).T) & 0xffff).astype(np.uint16))


# ------------------------------------------------------------------ golden vectors

@pytest.mark.parametrize("name", U8_CASES)
def test_golden_u8(golden_cases, name):
    z, _ = golden_cases
    a, b, ratio = z[name + ".a"], z[name + ".b"], float(z[name + ".ratio"])
    opts = MatchingBase.Options()
    opts.sift_matching_opts.lowe_ratio_threshold = ratio
    with matcher([a, b], opts=opts) as m:
        tw = m.twoway_match(KIND_SIFT_U8, 0, 1)
        assert np.array_equal(tw.matches_1_2, z[name + ".t12"])
        assert np.array_equal(tw.matches_2_1, z[name + ".t21"])
        res = m.pairwise_match(0, 1)
        assert np.array_equal(res.matches_1_2, z[name + ".f12"])
        assert np.array_equal(res.matches_2_1, z[name + ".f21"])
        assert m.last_consistent == int(z[name + ".count"])
        assert_clean(m)


@pytest.mark.parametrize("name", S8_CASES)
def test_golden_s8(golden_cases, name):
    z, _ = golden_cases
    a, b, ratio = z[name + ".a"], z[name + ".b"], float(z[name + ".ratio"])
    opts = MatchingBase.Options()
    opts.surf_matching_opts.lowe_ratio_threshold = ratio
    with matcher(surf=[a, b], opts=opts) as m:
        tw = m.twoway_match(KIND_SURF_S8, 0, 1)
        assert np.array_equal(tw.matches_1_2, z[name + ".t12"])
        assert np.array_equal(tw.matches_2_1, z[name + ".t21"])
        res = m.pairwise_match(0, 1)
        assert np.array_equal(res.matches_1_2, z[name + ".f12"])
        assert np.array_equal(res.matches_2_1, z[name + ".f21"])
        assert_clean(m)


def test_golden_real_image_pair(golden_real):
    g = golden_real
    with matcher([g["sift_0"], g["sift_1"]]) as m:
        tw = m.twoway_match(KIND_SIFT_U8, 1, 0)
        assert np.array_equal(tw.matches_1_2, g["twoway_12"])
        assert np.array_equal(tw.matches_2_1, g["twoway_21"])
        res = m.pairwise_match(1, 0)
        assert np.array_equal(res.matches_1_2, g["match_12"])
        assert np.array_equal(res.matches_2_1, g["match_21"])
        assert m.last_consistent == 810
        assert m.pairwise_match_lowres(1, 0, 500) == int(g["lowres_500"])
        assert_clean(m)


def test_golden_three_image_set(golden_triple):
    """BASELINE config 1 restated: the three images the reference ships, all 3 pairs in
    bundler::Matching::compute's order, pair by pair (with and without look-ahead) and batched,
    against the reference's own ExhaustiveMatching results."""
    g = golden_triple
    pairs = [(1, 0), (2, 0), (2, 1)]
    with matcher([g["sift_0"], g["sift_1"], g["sift_2"]]) as m:
        for window in (0, 8):
            m.set_lookahead(window)
            for v1, v2 in pairs:
                assert m.pairwise_match_lowres(v1, v2, 500) == int(g[f"lowres_{v1}{v2}"])
                res = m.pairwise_match(v1, v2)
                assert np.array_equal(res.matches_1_2, g[f"match_{v1}{v2}_12"])
                assert np.array_equal(res.matches_2_1, g[f"match_{v1}{v2}_21"])
        m.set_lookahead(0)
        results, counts = m.match_pairs(pairs)
        for (v1, v2), r in zip(pairs, results):
            assert np.array_equal(r.matches_1_2, g[f"match_{v1}{v2}_12"])
            assert np.array_equal(r.matches_2_1, g[f"match_{v1}{v2}_21"])
        assert counts.tolist() == [810, 810, 2376]
        assert_clean(m)


def test_quantiser_matches_convert_descriptor(ora, golden_real):
    """set_view_f32 quantises on the device exactly like convert_descriptor."""
    g = golden_real
    rng = np.random.default_rng(11)
    f = np.concatenate([g["float_sample"], rng.uniform(-0.2, 1.3, (200, 128)).astype(np.float32)])
    f[64, :6] = [0.5 / 255, 1.5 / 255, 2.5 / 255, 254.5 / 255, 0.49999 / 255, 1.0]
    s = rng.uniform(-1.2, 1.2, (150, 64)).astype(np.float32)
    s[0, :4] = [0.5 / 127, -0.5 / 127, 1.5 / 127, -126.5 / 127]
    qs, qf = ora.quantize_sift(f), ora.quantize_surf(s)
    with ExhaustiveMatching() as m:
        m.init([Viewport(FeatureSet(f, s)), Viewport(FeatureSet(qs, qf))])
        # identical quantised sets -> the similarity of view 0 with view 1 has the squared
        # norms on its diagonal and is symmetric; compare it with the oracle's bytes
        got = m.debug_dump_similarity(KIND_SIFT_U8, 0, 1)
        want = qs.astype(np.int64) @ qs.astype(np.int64).T
        assert np.array_equal(got, want)
        got = m.debug_dump_similarity(KIND_SURF_S8, 0, 1)
        assert np.array_equal(got, qf.astype(np.int64) @ qf.astype(np.int64).T)


# ------------------------------------------------------------------ seeded random shapes

@pytest.mark.parametrize("n1,n2", [(1, 1), (1, 1000), (1000, 1), (127, 255), (128, 256), (129, 257),
                                   (500, 700), (2000, 3000), (4096, 4096), (5000, 333)])
def test_twoway_and_filtered_match_oracle(ora, n1, n2):
    vs = synth.sift_views(7, 2, max(n1, n2))
    a, b = vs[0][:n1], vs[1][:n2]
    with matcher([a, b]) as m:
        tw = m.twoway_match(KIND_SIFT_U8, 0, 1)
        res = m.pairwise_match(0, 1)
        assert_clean(m)
    o12, o21 = ora.twoway("u8", a, b, 0.8)
    assert np.array_equal(tw.matches_1_2, o12) and np.array_equal(tw.matches_2_1, o21)
    f12, f21 = ora.remove_inconsistent(o12, o21)
    assert np.array_equal(res.matches_1_2, f12) and np.array_equal(res.matches_2_1, f21)


@pytest.mark.parametrize("n1,n2", [(1, 1), (300, 1000), (129, 257), (2000, 3000), (4096, 4100)])
def test_unit_norm_descriptors_take_the_filter(ora, n1, n2):
    """Quantised unit vectors (what SIFT produces, synth noise="renorm"): almost every row
    carries the 16-bit norm certificate, so this exercises the packed filter epilogue and the
    EXACT pass over its survivors rather than the uncertified route."""
    vs = synth.sift_views(17, 2, max(n1, n2), noise="renorm")
    a, b = vs[0][:n1], vs[1][:n2]
    with matcher([a, b]) as m:
        tw = m.twoway_match(KIND_SIFT_U8, 0, 1)
        res = m.pairwise_match(0, 1)
        assert_clean(m)
        st = m.stats()
    rows = 2 * (n1 + n2)
    assert st["slow_rows"] <= 0.05 * rows + 4, st      # few rows without certificate ...
    assert st["candidate_rows"] < 0.6 * rows + 8, st   # ... and the filter rejects most rows
    o12, o21 = ora.twoway("u8", a, b, 0.8)
    assert np.array_equal(tw.matches_1_2, o12) and np.array_equal(tw.matches_2_1, o21)
    f12, f21 = ora.remove_inconsistent(o12, o21)
    assert np.array_equal(res.matches_1_2, f12) and np.array_equal(res.matches_2_1, f21)
    if min(n1, n2) >= 1000:
        assert (f12 >= 0).sum() > 20


def test_duplicate_descriptors_and_ties(ora):
    """Exact duplicates inside and across views: best == second best (ratio 1, or 0/0 = NaN
    which accepts), ties broken towards the highest index (nearest_neighbor.cc:89)."""
    vs = synth.sift_views(18, 2, 1200, noise="renorm")
    a, b = vs[0].copy(), vs[1].copy()
    b[100:140] = a[100:140]          # identical descriptor in both views
    b[700:720] = a[100:120]          # ... and a second copy of some of them (tie, later index)
    a[300:310] = a[100:110]          # duplicates inside a view
    b[1100] = b[3]
    for ratio in (0.8, 1.0):
        opts = MatchingBase.Options()
        opts.sift_matching_opts.lowe_ratio_threshold = ratio
        with matcher([a, b], opts=opts) as m:
            tw = m.twoway_match(KIND_SIFT_U8, 0, 1)
            assert_clean(m)
        o12, o21 = ora.twoway("u8", a, b, ratio)
        assert np.array_equal(tw.matches_1_2, o12) and np.array_equal(tw.matches_2_1, o21)


def test_degenerate_rows(ora):
    """All-zero descriptors (every similarity 0: the best is the reference's initial value and
    the LAST candidate wins), zero rows mixed into real ones, and a candidate count that leaves
    a ragged last tile holding other views' rows."""
    vs = synth.sift_views(19, 3, 700, noise="renorm")
    a, b, c = vs[0].copy(), vs[1][:300].copy(), vs[2][:513].copy()
    a[5] = 0
    a[6] = 0
    b[7] = 0
    z = np.zeros((259, 128), np.uint8)
    with matcher([a, b, c, z]) as m:
        got = {(i, j): m.twoway_match(KIND_SIFT_U8, i, j) for i, j in [(0, 1), (0, 2), (1, 2), (3, 0), (3, 3), (1, 3)]}
        assert_clean(m)
    sets = [a, b, c, z]
    for (i, j), tw in got.items():
        o12, o21 = ora.twoway("u8", sets[i], sets[j], 0.8)
        assert np.array_equal(tw.matches_1_2, o12) and np.array_equal(tw.matches_2_1, o21), (i, j)


def test_mixed_certified_wrapping_and_doubtful_rows(ora):
    """One view pair containing every route through the pipeline: unit-norm rows (norm
    certificate), rows with inflated norms that no candidate takes to 2^16 (certified after the
    fact), and rows that really reach 2^16 (EXACT pass, wrapped 16-bit stores)."""
    vs = synth.sift_views(20, 2, 2000, noise="renorm")
    a, b = vs[0].copy(), vs[1].copy()
    rng = np.random.default_rng(3)
    hot = rng.choice(2000, 60, replace=False)
    a[hot[:30]] = np.minimum(a[hot[:30]].astype(np.int32) + 3, 255).astype(np.uint8)   # norms up by ~10 %
    b[hot[30:]] = np.minimum(b[hot[30:]].astype(np.int32) + 3, 255).astype(np.uint8)
    b[hot[:10]] = a[hot[:10]]                                                            # ... some of them twins
    with matcher([a, b]) as m:
        tw = m.twoway_match(KIND_SIFT_U8, 0, 1)
        assert_clean(m)
        st = m.stats()
    o12, o21 = ora.twoway("u8", a, b, 0.8)
    assert np.array_equal(tw.matches_1_2, o12) and np.array_equal(tw.matches_2_1, o21)
    s = a.astype(np.int64) @ b.astype(np.int64).T
    assert (s >= 65536).any()                      # the EXACT pass had work ...
    assert 0 < st["exact_rows"] < 400, st           # ... but only on the rows that reach 2^16


@pytest.mark.parametrize("mode", [0, 1])
def test_exact_rows_both_paths(ora, mode):
    """The rows that reach 2^16 are replayed either from CUDA-core inner products spread over the
    device (mode 0: they fit the scratch buffer) or by the tensor-core scan pass (mode 1): same
    results on inflated norms, on arbitrary bytes (every 16-bit lane wraps, the verify / replay
    route) and through the batched entry point with its reverse pass."""
    vs = synth.sift_views(20, 2, 2000, noise="renorm")
    a, b = vs[0].copy(), vs[1].copy()
    rng = np.random.default_rng(3)
    hot = rng.choice(2000, 60, replace=False)
    a[hot[:30]] = np.minimum(a[hot[:30]].astype(np.int32) + 3, 255).astype(np.uint8)
    b[hot[30:]] = np.minimum(b[hot[30:]].astype(np.int32) + 3, 255).astype(np.uint8)
    b[hot[:10]] = a[hot[:10]]
    c = rng.integers(0, 256, (333, 128), dtype=np.uint8)
    d = rng.integers(0, 256, (517, 128), dtype=np.uint8)
    d[:100] = c[:100]
    lsb = synth.sift_views(32, 2, 900)                    # norms drift: many rows reach 2^16
    sets = [a, b, c, d, lsb[0][:700], lsb[1]]
    with matcher(sets) as m:
        m.debug_set_exact_path(mode)
        for i, j in [(0, 1), (2, 3), (4, 5), (3, 0)]:
            tw = m.twoway_match(KIND_SIFT_U8, i, j)
            res = m.pairwise_match(i, j)
            o12, o21 = ora.twoway("u8", sets[i], sets[j], 0.8)
            assert np.array_equal(tw.matches_1_2, o12) and np.array_equal(tw.matches_2_1, o21), (i, j)
            f12, f21 = ora.remove_inconsistent(o12, o21)
            assert np.array_equal(res.matches_1_2, f12) and np.array_equal(res.matches_2_1, f21), (i, j)
        assert_clean(m)
        st = m.stats()
    assert st["exact_rows"] > 100, st
    if mode == 0:     # (the lsb pair's rows exceed the scratch buffer sized for views this small)
        assert 100 < st["exact_wide_rows"] <= st["exact_rows"], st
    else:
        assert st["exact_wide_rows"] == 0, st


def test_surf_degenerate_rows(ora):
    """Signed kind: rows whose best similarity is negative (index stays 0), exactly zero, and
    duplicates; ragged sizes."""
    pool = synth.surf_pool(6, 300)
    a, b = synth.surf_view(6, 0, 700, pool).copy(), synth.surf_view(6, 1, 515, pool).copy()
    b[:40] = -a[:40]              # anti-parallel twins: large negative similarities
    a[50] = 0                     # similarity 0 with everything
    b[60:64] = a[100]             # four copies of one row
    only_neg_b = -np.abs(b[:20])
    only_pos_a = np.abs(a[:33])   # every similarity of this pair of sets is <= 0
    with matcher(surf=[a, b, only_pos_a, only_neg_b]) as m:
        got = {(i, j): m.twoway_match(KIND_SURF_S8, i, j) for i, j in [(0, 1), (2, 3), (3, 2), (0, 3)]}
        assert_clean(m)
    sets = [a, b, only_pos_a, only_neg_b]
    for (i, j), tw in got.items():
        o12, o21 = ora.twoway("s8", sets[i], sets[j], 0.7)
        assert np.array_equal(tw.matches_1_2, o12) and np.array_equal(tw.matches_2_1, o21), (i, j)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_adversarial_bytes_match_oracle(ora, seed):
    """Arbitrary bytes: inner products up to 8.3e6, every 16-bit lane wraps."""
    rng = np.random.default_rng(seed)
    hi = [256, 96, 48, 200][seed]
    n1, n2 = int(rng.integers(1, 600)), int(rng.integers(1, 600))
    a = rng.integers(0, hi, (n1, 128), dtype=np.uint8)
    b = rng.integers(0, hi, (n2, 128), dtype=np.uint8)
    k = min(n1, n2) // 3
    b[:k] = a[:k]                       # exact duplicates
    with matcher([a, b]) as m:
        tw = m.twoway_match(KIND_SIFT_U8, 0, 1)
        assert_clean(m)
    o12, o21 = ora.twoway("u8", a, b, 0.8)
    assert np.array_equal(tw.matches_1_2, o12) and np.array_equal(tw.matches_2_1, o21)


def _one_product_cases():
    """(name, kind, a, b, ratio): inputs that reach every route of the reverse-direction pass."""
    cases = []
    vs = synth.sift_views(31, 2, 1500, noise="renorm")
    cases.append(("renorm", "u8", vs[0], vs[1][:1300], 0.8))
    vs = synth.sift_views(32, 2, 900)                       # "lsb": norms drift, many rows reach 2^16
    cases.append(("lsb", "u8", vs[0][:700], vs[1], 0.8))
    a, b = synth.sift_views(33, 2, 1200, noise="renorm")
    a, b = a.copy(), b.copy()
    b[100:140] = a[100:140]                                 # twins across the views
    b[700:720] = a[100:120]                                 # ... twice: several claimants / ties
    a[300:310] = a[100:110]                                 # two rows of a claim the same row of b
    b[1100] = b[3]
    a[5] = 0
    b[7] = 0
    cases.append(("ties", "u8", a, b, 0.8))
    cases.append(("ties-ratio-1", "u8", a, b, 1.0))
    rng = np.random.default_rng(34)
    a = rng.integers(0, 256, (333, 128), dtype=np.uint8)    # arbitrary bytes: every 16-bit lane wraps
    b = rng.integers(0, 256, (517, 128), dtype=np.uint8)
    b[:100] = a[:100]
    cases.append(("bytes", "u8", a, b, 0.8))
    a = rng.integers(0, 48, (400, 128), dtype=np.uint8)     # small bytes: nothing certified, nothing wraps
    b = rng.integers(0, 48, (300, 128), dtype=np.uint8)
    b[:50] = a[:50]
    cases.append(("small-bytes", "u8", a, b, 0.8))
    z = np.zeros((259, 128), np.uint8)
    cases.append(("zeros", "u8", z, vs[1][:300], 1.0))
    pool = synth.surf_pool(35, 300)
    a, b = synth.surf_view(35, 0, 700, pool).copy(), synth.surf_view(35, 1, 515, pool).copy()
    b[:40] = -a[:40]
    a[50] = 0
    b[60:64] = a[100]
    cases.append(("surf", "s8", a, b, 0.7))
    cases.append(("surf-ratio-1", "s8", a, b, 1.0))
    cases.append(("surf-negative", "s8", np.abs(a[:33]), -np.abs(b[:20]), 1.0))   # every similarity <= 0
    pool = synth.surf_pool(36, 700)
    cases.append(("surf-large", "s8", synth.surf_view(36, 0, 1600, pool), synth.surf_view(36, 1, 1400, pool), 0.7))
    a = rng.integers(-127, 128, (300, 64), dtype=np.int8)   # arbitrary signed bytes: no norm certificate
    b = rng.integers(-127, 128, (280, 64), dtype=np.int8)
    b[:60] = a[:60]
    cases.append(("surf-bytes", "s8", a, b, 0.7))
    return cases


@pytest.mark.parametrize("case", _one_product_cases(), ids=lambda c: c[0])
def test_one_product_per_pair_equals_both_directions(ora, case):
    """pairwise_match runs ONE direction of a pair through the filter pass and evaluates the
    other direction only for the rows the first one claims (post_kernels.cuh, claim_kernel).
    The result must be what the reference's two full scans + remove_inconsistent_matches give,
    and what the same kernels give with both directions scanned."""
    _, kind, a, b, ratio = case
    opts = MatchingBase.Options()
    opts.sift_matching_opts.lowe_ratio_threshold = ratio
    opts.surf_matching_opts.lowe_ratio_threshold = ratio
    kw = {"sift": [a, b]} if kind == "u8" else {"surf": [a, b]}
    with matcher(opts=opts, **kw) as m:
        one = [m.pairwise_match(0, 1), m.pairwise_match(1, 0)]
        lowres_one = m.pairwise_match_lowres(0, 1, 200)
        claimed = m.stats()["claimed_rows"]
        restricted = m.stats()["reverse_restricted_pairs"]
        m.debug_set_both_directions(2)                    # claimed rows against the whole other view
        whole = [m.pairwise_match(0, 1), m.pairwise_match(1, 0)]
        claimed_after_whole = m.stats()["claimed_rows"]
        m.debug_set_both_directions(1)
        both = [m.pairwise_match(0, 1), m.pairwise_match(1, 0)]
        lowres_both = m.pairwise_match_lowres(0, 1, 200)
        assert m.stats()["claimed_rows"] == claimed_after_whole   # nothing is claimed when both directions are scanned
        assert_clean(m)
    o12, o21 = ora.twoway(kind, a, b, ratio)
    f12, f21 = ora.remove_inconsistent(o12, o21)
    for r in (one[0], whole[0], both[0]):
        assert np.array_equal(r.matches_1_2, f12) and np.array_equal(r.matches_2_1, f21)
    for r in (one[1], whole[1], both[1]):
        assert np.array_equal(r.matches_1_2, f21) and np.array_equal(r.matches_2_1, f12)
    assert lowres_one == lowres_both
    if case[0] in ("renorm", "surf-large"):
        assert restricted == 2, restricted      # both calls held their claimed rows against a subset of the other view
    assert 0 <= claimed <= (o12 >= 0).sum() + (o21 >= 0).sum() + 200   # at most one row per forward match


def test_replay_list_holds_each_row_once(ora):
    """Saturated / arbitrary bytes at a size where the EXACT pass's replay list would overflow
    if a row entered it once per big candidate: every similarity reaches 2^16 and every 16-bit
    lane wraps, so every row is replayed on CUDA cores -- once."""
    rng = np.random.default_rng(77)
    a = rng.integers(128, 256, (10240, 128), dtype=np.uint8)
    a[::7] = 255
    b = rng.integers(128, 256, (1031, 128), dtype=np.uint8)
    b[::5] = 255
    with matcher([a, b]) as m:
        tw = m.twoway_match(KIND_SIFT_U8, 0, 1)
        res = m.pairwise_match(0, 1)
        assert_clean(m)
    o12, o21 = ora.twoway("u8", a, b, 0.8)
    assert np.array_equal(tw.matches_1_2, o12) and np.array_equal(tw.matches_2_1, o21)
    f12, f21 = ora.remove_inconsistent(o12, o21)
    assert np.array_equal(res.matches_1_2, f12) and np.array_equal(res.matches_2_1, f21)


@pytest.mark.parametrize("ratio,dist", [(0.8, None), (1.0, None), (0.6, None), (0.8, 150.0), (0.95, 60.0)])
def test_thresholds(ora, ratio, dist):
    vs = synth.sift_views(9, 2, 900)
    opts = MatchingBase.Options()
    opts.sift_matching_opts.lowe_ratio_threshold = ratio
    if dist is not None:
        opts.sift_matching_opts.distance_threshold = dist
    with matcher(vs, opts=opts) as m:
        tw = m.twoway_match(KIND_SIFT_U8, 0, 1)
    kw = {} if dist is None else {"dist": dist}
    o12, o21 = ora.twoway("u8", vs[0], vs[1], ratio, **kw)
    assert np.array_equal(tw.matches_1_2, o12) and np.array_equal(tw.matches_2_1, o21)


def test_surf_synthetic_matches_oracle(ora):
    pool = synth.surf_pool(4, 400)
    a, b = synth.surf_view(4, 0, 1500, pool), synth.surf_view(4, 1, 1100, pool)
    with matcher(surf=[a, b]) as m:
        tw = m.twoway_match(KIND_SURF_S8, 0, 1)
        assert_clean(m)
    o12, o21 = ora.twoway("s8", a, b, 0.7)
    assert np.array_equal(tw.matches_1_2, o12) and np.array_equal(tw.matches_2_1, o21)
    assert (o12 >= 0).sum() > 50


# ------------------------------------------------------------------ float path

def _unit(x):
    return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)


def test_golden_f32(golden_cases):
    z, _ = golden_cases
    name = "f32.synth_200x260"
    with ExhaustiveMatching() as m:
        r = m.twoway_match_f32(Matching.Options(128, float(z[name + ".ratio"]), 3.402823466e+38),
                               z[name + ".a"], z[name + ".b"])
    assert np.array_equal(r.matches_1_2, z[name + ".t12"])
    assert np.array_equal(r.matches_2_1, z[name + ".t21"])


@pytest.mark.parametrize("n1,n2,dim,ratio,dist", [(1, 1, 128, 0.8, None), (300, 500, 128, 0.8, None),
                                                  (777, 65, 128, 0.9, 0.5), (200, 1000, 64, 0.7, None),
                                                  (1500, 1500, 128, 0.8, None), (64, 129, 64, 1.0, 0.05)])
def test_float_path_is_bit_identical_to_reference_order(ora, n1, n2, dim, ratio, dist):
    """twoway_match<float>: the GPU forms every inner product in the reference's SSE3
    summation order, so the bar is equality, not the 1e-5 tie tolerance."""
    rng = np.random.default_rng(n1 * 7 + n2)
    signed = dim == 64
    a = rng.standard_normal((n1, dim)) if signed else np.abs(rng.standard_normal((n1, dim)))
    b = rng.standard_normal((n2, dim)) if signed else np.abs(rng.standard_normal((n2, dim)))
    a, b = _unit(a), _unit(b)
    k = min(n1, n2) // 3
    b[:k] = _unit(a[:k] + 0.02 * rng.standard_normal((k, dim)))
    if k > 4:
        b[k:k + 2] = a[:2]              # exact duplicates: ties and d = 0
    thr = 3.402823466e+38 if dist is None else dist
    with ExhaustiveMatching() as m:
        r = m.twoway_match_f32(Matching.Options(dim, ratio, thr), a, b)
    o12, o21 = ora.twoway("f32", a, b, ratio, dist=thr, sse3_order=True)
    assert np.array_equal(r.matches_1_2, o12) and np.array_equal(r.matches_2_1, o21)
    if n1 > 100 and ratio < 1.0:
        assert (o12 >= 0).sum() > 10


@pytest.mark.parametrize("n1,n2,dim,ratio,dist", [(1, 1, 128, 0.8, None), (300, 500, 128, 0.8, None),
                                                  (777, 65, 128, 0.9, 0.5), (200, 1000, 64, 0.7, None),
                                                  (1500, 1500, 128, 0.8, None), (64, 129, 64, 1.0, 0.05),
                                                  (2500, 2100, 128, 0.8, None), (257, 3000, 128, 0.8, 0.3)])
def test_float_path_filter_first_gives_the_same_vectors(ora, n1, n2, dim, ratio, dist):
    """The tensor-core filter (tf32 hi/lo split) decides the rows whose outcome is clear within its
    error bound and leaves the rest to the exact kernel: the vectors stay the reference's bit for
    bit -- near duplicates, exact duplicates (ties, d = 0), signed 64-d input, ragged tiles."""
    rng = np.random.default_rng(n1 * 7 + n2)
    signed = dim == 64
    a = rng.standard_normal((n1, dim)) if signed else np.abs(rng.standard_normal((n1, dim)))
    b = rng.standard_normal((n2, dim)) if signed else np.abs(rng.standard_normal((n2, dim)))
    a, b = _unit(a), _unit(b)
    k = min(n1, n2) // 3
    b[:k] = _unit(a[:k] + 0.02 * rng.standard_normal((k, dim)))
    if k > 4:
        b[k:k + 2] = a[:2]
    thr = 3.402823466e+38 if dist is None else dist
    with ExhaustiveMatching() as m:
        m.debug_set_float_path(2)
        r = m.twoway_match_f32(Matching.Options(dim, ratio, thr), a, b)
        st = m.stats()
        m.debug_set_float_path(1)
        r1 = m.twoway_match_f32(Matching.Options(dim, ratio, thr), a, b)
    o12, o21 = ora.twoway("f32", a, b, ratio, dist=thr, sse3_order=True)
    assert np.array_equal(r.matches_1_2, o12) and np.array_equal(r.matches_2_1, o21)
    assert np.array_equal(r1.matches_1_2, o12) and np.array_equal(r1.matches_2_1, o21)
    assert st["float_filter_rows"] == n1 + n2
    if n1 >= 1500 and not signed:       # the filter decides nearly everything
        assert st["float_exact_rows"] < 0.05 * (n1 + n2), st


def test_float_path_filter_unusual_input(ora):
    """Rows that are not unit vectors (the bound scales with the norms), zero rows, a huge row and
    non-finite values (everything goes to the exact kernel): same vectors as the exact path."""
    rng = np.random.default_rng(77)
    a = np.abs(rng.standard_normal((1100, 128))).astype(np.float32) * rng.uniform(0.01, 30.0, (1100, 1)).astype(np.float32)
    b = np.abs(rng.standard_normal((1300, 128))).astype(np.float32) * rng.uniform(0.01, 30.0, (1300, 1)).astype(np.float32)
    b[:300] = a[:300] * np.float32(1.5)
    a[5] = 0
    b[7] = 0
    opts = Matching.Options(128, 0.8, 3.402823466e+38)
    with ExhaustiveMatching() as m:
        for variant in range(3):
            aa, bb = a.copy(), b.copy()
            if variant == 1:
                aa[11, 3] = 1e30
            if variant == 2:
                bb[13, 5] = np.inf
                aa[17, 1] = np.nan
            m.debug_set_float_path(1)
            want = m.twoway_match_f32(opts, aa, bb)
            m.debug_set_float_path(2)
            got = m.twoway_match_f32(opts, aa, bb)
            assert np.array_equal(got.matches_1_2, want.matches_1_2), variant
            assert np.array_equal(got.matches_2_1, want.matches_2_1), variant
            if variant == 0:
                o12, o21 = ora.twoway("f32", aa, bb, 0.8, dist=3.402823466e+38, sse3_order=True)
                assert np.array_equal(got.matches_1_2, o12) and np.array_equal(got.matches_2_1, o21)


def test_float_filter_error_is_far_below_the_bound():
    """The similarities the tensor cores produce from the tf32 hi/lo split against float64: the
    decision rule allows 6e-5 |a| |b|; what is seen must be a small fraction of it."""
    rng = np.random.default_rng(5)
    a = _unit(np.abs(rng.standard_normal((700, 128))))
    b = _unit(np.abs(rng.standard_normal((900, 128))))
    b[:200] = _unit(a[:200] + 0.01 * rng.standard_normal((200, 128)))
    a[300:] *= np.float32(7.0)                      # the error scales with the norms
    with ExhaustiveMatching() as m:
        s1, s2, j1 = m.debug_float_filter(a, b)
    S = a.astype(np.float64) @ b.astype(np.float64).T
    na, nb = np.linalg.norm(a.astype(np.float64), axis=1), np.linalg.norm(b.astype(np.float64), axis=1)
    for top1, top2, idx, M, nq, ncmax in ((s1[:700], s2[:700], j1[:700], S, na, nb.max()),
                                          (s1[700:], s2[700:], j1[700:], S.T, nb, na.max())):
        srt = np.sort(M, axis=1)
        err1 = np.abs(top1 - srt[:, -1]) / (nq * ncmax)
        err2 = np.abs(top2 - srt[:, -2]) / (nq * ncmax)
        assert err1.max() < 3e-6 and err2.max() < 3e-6, (err1.max(), err2.max())
        picked = M[np.arange(M.shape[0]), idx]
        assert (np.abs(picked - srt[:, -1]) / (nq * ncmax)).max() < 3e-6


def test_float_path_empty_sets():
    a = _unit(np.abs(np.random.default_rng(0).standard_normal((5, 128))))
    with ExhaustiveMatching() as m:
        r = m.twoway_match_f32(Matching.Options(128, 0.8, 3.402823466e+38), a[:0], a)
    assert r.matches_1_2.size == 0 and (r.matches_2_1 == -1).all()


# ------------------------------------------------------------------ the plugin surface

def test_pairwise_match_sift_plus_surf_combined(ora):
    """ExhaustiveMatching::pairwise_match incl. combine_results and the rule that a feature
    type only takes part if view_1 has descriptors of it."""
    sp, fp = synth.scene_pool(5, 300), synth.surf_pool(5, 200)
    sizes = [(600, 300), (700, 250), (0, 200), (400, 0), (0, 0)]
    sift = [synth.sift_view(5, v, s, sp) for v, (s, _) in enumerate(sizes)]
    surf = [synth.surf_view(5, v, f, fp) for v, (_, f) in enumerate(sizes)]
    with matcher(sift, surf) as m:
        for v1 in range(len(sizes)):
            for v2 in range(len(sizes)):
                if v1 == v2:
                    continue
                res = m.pairwise_match(v1, v2)
                o12, o21 = ora.pairwise_match(sift[v1], sift[v2], surf[v1], surf[v2])
                assert np.array_equal(res.matches_1_2, o12), (v1, v2)
                assert np.array_equal(res.matches_2_1, o21), (v1, v2)
                assert m.pairwise_match_lowres(v1, v2, 150) == \
                    ora.pairwise_match_lowres(sift[v1], sift[v2], surf[v1], surf[v2], 150)
        assert_clean(m)


def test_batched_equals_single_pair_and_compact_lists(ora):
    import torch
    sizes = [700, 1, 0, 333, 1024, 129]
    pool = synth.scene_pool(6, 400)
    views = [synth.sift_view(6, v, n, pool) for v, n in enumerate(sizes)]
    pairs = synth.all_pairs(len(sizes))
    with matcher(views) as m:
        results, counts = m.match_pairs(pairs)
        out_ij = torch.empty((4096, 2), dtype=torch.int32, device="cuda")
        loff = m.match_pairs_compact(pairs, out_ij)
        ij = out_ij.cpu().numpy()
        no_surf = np.zeros((0, 64), np.int8)
        for p, (v1, v2) in enumerate(pairs):
            # the batched call has pairwise_match semantics: a pair whose view_1 has no SIFT
            # features yields two EMPTY vectors (exhaustive_matching.cc:123)
            o12, o21 = ora.pairwise_match(views[v1], views[v2], no_surf, no_surf)
            assert np.array_equal(results[p].matches_1_2, o12), (v1, v2)
            assert np.array_equal(results[p].matches_2_1, o21), (v1, v2)
            assert counts[p] == int((o12 >= 0).sum())
            lst = ij[loff[p]:loff[p + 1]]
            idx = np.nonzero(o12 >= 0)[0]
            assert np.array_equal(lst[:, 0], idx) and np.array_equal(lst[:, 1], o12[idx])
        assert_clean(m)
        # too small a list buffer is an error, not a silent truncation
        with pytest.raises(MatcherError):
            m.match_pairs_compact(pairs, out_ij[:3])


def test_error_behaviour():
    vs = synth.sift_views(8, 2, 64)
    m = ExhaustiveMatching()
    with pytest.raises(MatcherError) as ei:      # match before init/commit
        m.pairwise_match_lowres(0, 1, 10)
    assert ei.value.code == -4
    m.init(vps(vs))
    with pytest.raises(MatcherError) as ei:
        m.pairwise_match_lowres(0, 5, 10)
    assert ei.value.code == -1
    with pytest.raises(MatcherError):
        m.match_pairs(np.array([[0, -1]], np.int32))
    # the handle stays usable after an error
    assert isinstance(m.pairwise_match(0, 1), Matching.Result)
    m.close()


# ------------------------------------------------------------------ full-size properties

def test_full_size_pair_properties(ora):
    """BASELINE config 2 sizes (8192 x 8192): checked through size-independent properties
    plus a sample of rows against the oracle's nearest-neighbour search."""
    pool = synth.scene_pool(2, 4096)
    a, b = synth.sift_view(2, 0, 8192, pool), synth.sift_view(2, 1, 8192, pool)
    with matcher([a, b]) as m:
        tw = m.twoway_match(KIND_SIFT_U8, 0, 1)
        res = m.pairwise_match(0, 1)
        swapped = m.pairwise_match(1, 0)
        assert_clean(m)
    m12, m21 = res.matches_1_2, res.matches_2_1
    # mutual filter: both vectors hold exactly the same pairs
    i = np.nonzero(m12 >= 0)[0]
    assert np.array_equal(m21[m12[i]], i)
    j = np.nonzero(m21 >= 0)[0]
    assert np.array_equal(m12[m21[j]], j) and i.size == j.size
    # filtered is a subset of the one-way result
    assert np.array_equal(m12[i], tw.matches_1_2[i])
    # swapping the views swaps the vectors
    assert np.array_equal(swapped.matches_1_2, m21) and np.array_equal(swapped.matches_2_1, m12)
    assert i.size > 800    # planted near-duplicates are found
    # sample rows against the oracle, both directions
    rng = np.random.default_rng(0)
    for r in rng.integers(0, 8192, 48):
        o = ora.twoway("u8", a[r:r + 1], b, 0.8)[0][0]
        assert tw.matches_1_2[r] == o
        o = ora.twoway("u8", b[r:r + 1], a, 0.8)[0][0]
        assert tw.matches_2_1[r] == o


def test_baseline_config_2_lists_equal_the_reference():
    """BASELINE config 2 at full size (36 x 8192, all 630 pairs in one batched call): the
    correspondence lists of a spread of pairs against the reference matcher itself (compiled
    from its sources, oracle/_ref) or, where that is not built, the C restatement."""
    import oracle
    views = synth.sift_views(2, 36, 8192, noise="renorm")
    pairs = synth.all_pairs(36)
    with matcher(views) as m:
        out = np.empty((630 * 2048, 2), np.int32)
        loff = m.match_pairs_lists(pairs, out)
        assert_clean(m)
        st = m.stats()
    assert st["exact_rows"] < 64                     # unit-norm data: next to nothing reaches 2^16
    assert st["reverse_restricted_pairs"] >= 600     # (nearly) every pair's claimed rows met a subset of the other view ...
    assert st["reverse_candidate_rows"] < 630 * 8192 // 4    # ... a small one
    sample = list(range(0, 630, 53))                 # 12 pairs
    impl = oracle.Reference() if oracle.have_ref() else oracle.Oracle()
    from concurrent.futures import ThreadPoolExecutor      # (ctypes releases the GIL)
    with ThreadPoolExecutor(len(sample)) as ex:
        want = list(ex.map(lambda p: impl.match_filtered("u8", views[pairs[p][0]], views[pairs[p][1]], 0.8)[0], sample))
    for p, o12 in zip(sample, want):
        v1, v2 = pairs[p]
        i = np.nonzero(o12 >= 0)[0]
        got = out[loff[p]:loff[p + 1]]
        assert np.array_equal(got[:, 0], i) and np.array_equal(got[:, 1], o12[i]), (v1, v2)
        assert i.size > 500


def _digests_of_lists(out, loff):
    import oracle
    return np.array([oracle.list_digest(out[loff[p]:loff[p + 1]]) for p in range(len(loff) - 1)], np.uint64)


def test_baseline_config_2_all_630_lists_equal_the_reference():
    """BASELINE config 2 in full: every one of the 630 correspondence lists against the
    reference matcher itself (oracle/_ref, OpenMP over the pairs: about a minute of host time on
    16 cores), compared through per-pair counts and 64-bit digests of the (i, j) lists."""
    import oracle
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    views = synth.sift_views(2, 36, 8192, noise="renorm")
    pairs = synth.all_pairs(36)
    with matcher(views) as m:
        out = np.empty((630 * 2048, 2), np.int32)
        loff = m.match_pairs_lists(pairs, out)
        assert_clean(m)
    ref = oracle.Reference()
    ref.use_all_cores()
    counts, digests = ref.match_pairs_u8_digest(views, pairs, 0.8)
    assert np.array_equal(np.diff(loff), counts)
    assert np.array_equal(_digests_of_lists(out, loff), digests)
    assert counts.min() > 500


@pytest.mark.parametrize("nv,n", [(4, 16384), (3, 32768)])
def test_config_3_and_4_sized_views_lists_equal_the_reference(nv, n):
    """Views of BASELINE config 3 / 4 size (16 384 / 32 768 descriptors, i.e. MAX_FEATURES of
    src/matching/matching.h:24): the complete lists of all pairs of a few such views against the
    reference matcher (counts + digests)."""
    import oracle
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    views = synth.sift_views(3 if n == 16384 else 4, nv, n, noise="renorm")
    pairs = synth.all_pairs(nv)
    with matcher(views) as m:
        out = np.empty((len(pairs) * n // 4, 2), np.int32)
        loff = m.match_pairs_lists(pairs, out)
        assert_clean(m)
    ref = oracle.Reference()
    ref.use_all_cores()
    counts, digests = ref.match_pairs_u8_digest(views, pairs, 0.8)
    assert np.array_equal(np.diff(loff), counts)
    assert np.array_equal(_digests_of_lists(out, loff), digests)
    assert counts.min() > n // 16


def test_config_5_single_large_pair(ora):
    """BASELINE config 5: one pair of 200 000 x 200 000 descriptors (782 work items per
    direction, ragged last tile): list ordering, mutual consistency through a second call with
    the views swapped, and sampled rows against the oracle."""
    import torch
    n = 200000
    dev = torch.device("cuda", 0)
    pool = synth.torch_sift_views(5, 2, n, dev, noise="renorm")
    pool = torch.cat([pool, torch.zeros((256, 128), dtype=torch.uint8, device=dev)])
    m = ExhaustiveMatching(device=0)
    m.init_device_pool(pool, np.array([0, n], np.int64), np.array([n, n], np.int32))
    out = torch.empty((n, 2), dtype=torch.int32, device=dev)
    loff = m.match_pairs_compact(np.array([[1, 0]], np.int32), out)
    a = out[:int(loff[1])].cpu().numpy()
    loff2 = m.match_pairs_compact(np.array([[0, 1]], np.int32), out)
    b = out[:int(loff2[1])].cpu().numpy()
    st = m.stats()
    m.close()
    assert st["self_check_failures"] == 0
    assert st["exact_wide_rows"] == st["exact_rows"] > 0     # a few rows against 200 000: spread over the device
    assert a.shape[0] > 10000 and np.all(np.diff(a[:, 0]) > 0)
    # swapping the views transposes the list
    bs = b[np.argsort(b[:, 1], kind="stable")]
    assert np.array_equal(bs[:, ::-1], a)
    v1 = pool[n:2 * n].cpu().numpy()
    v0 = pool[:n].cpu().numpy()
    got = dict(zip(a[:, 0].tolist(), a[:, 1].tolist()))
    rng = np.random.default_rng(5)
    for r in list(rng.integers(0, n, 10)) + a[:6, 0].tolist():
        o = int(ora.twoway("u8", v1[r:r + 1], v0, 0.8)[0][0])
        back = int(ora.twoway("u8", v0[o:o + 1], v1, 0.8)[0][0]) if o >= 0 else -1
        assert got.get(int(r), -1) == (o if (o >= 0 and back == r) else -1), r


def test_config_5_full_list_equals_the_reference():
    """BASELINE config 5 in full: the whole correspondence list of the 200 000 x 200 000 pair
    against the reference matcher itself (oracle/_ref: its oneway_match on chunks of the query rows
    over all host cores, 8 x 10^10 executed comparisons: under a minute on 16 cores), by count and
    64-bit digest."""
    import os
    import torch
    import oracle
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    if len(os.sched_getaffinity(0)) < 8:
        pytest.skip("fewer than 8 host cores: the reference would take minutes")
    n = 200000
    dev = torch.device("cuda", 0)
    pool = synth.torch_sift_views(5, 2, n, dev, noise="renorm")
    v0, v1 = pool[:n].cpu().numpy(), pool[n:2 * n].cpu().numpy()
    pool = torch.cat([pool, torch.zeros((256, 128), dtype=torch.uint8, device=dev)])
    m = ExhaustiveMatching(device=0)
    m.init_device_pool(pool, np.array([0, n], np.int64), np.array([n, n], np.int32))
    out = torch.empty((n, 2), dtype=torch.int32, device=dev)
    loff = m.match_pairs_compact(np.array([[1, 0]], np.int32), out)
    got = out[:int(loff[1])].cpu().numpy()
    st = m.stats()
    m.close()
    assert st["self_check_failures"] == 0
    ref = oracle.Reference()
    ref.use_all_cores()
    count, digest, m12, _ = ref.match_large_pair_u8(v1, v0, 0.8)
    assert got.shape[0] == count
    assert oracle.list_digest(got) == digest
    rows = np.flatnonzero(m12 >= 0)
    assert np.array_equal(got[:, 0], rows) and np.array_equal(got[:, 1], m12[rows])


def test_more_rows_than_one_batch(ora):
    """40 views x 16384, all 780 pairs: 25.6 M job rows, i.e. more than the 16 M rows one batch
    of scratch memory holds; lists of pairs from every batch against the oracle (sampled rows),
    and the list of a pair matched alone."""
    import torch
    nv, n = 40, 16384
    dev = torch.device("cuda", 0)
    pool = synth.torch_sift_views(13, nv, n, dev, noise="renorm")
    pool = torch.cat([pool, torch.zeros((256, 128), dtype=torch.uint8, device=dev)])
    m = ExhaustiveMatching(device=0)
    m.init_device_pool(pool, np.arange(nv, dtype=np.int64) * n, np.full(nv, n, np.int32))
    pairs = synth.all_pairs(nv)
    out = torch.empty((len(pairs) * n // 4, 2), dtype=torch.int32, device=dev)
    loff = m.match_pairs_compact(pairs, out)
    lists = out[:int(loff[-1])].cpu().numpy()
    st = m.stats()
    assert st["self_check_failures"] == 0
    single = torch.empty((n, 2), dtype=torch.int32, device=dev)
    rng = np.random.default_rng(1)
    for p in (0, 399, 400, 779):           # first / last pairs of both batches
        v1, v2 = pairs[p]
        got = lists[loff[p]:loff[p + 1]]
        assert got.shape[0] > 1000 and np.all(np.diff(got[:, 0]) > 0)
        lo = m.match_pairs_compact(pairs[p:p + 1], single)
        assert np.array_equal(single[:int(lo[1])].cpu().numpy(), got), p
        a = pool[v1 * n:(v1 + 1) * n].cpu().numpy()
        b = pool[v2 * n:(v2 + 1) * n].cpu().numpy()
        d = dict(zip(got[:, 0].tolist(), got[:, 1].tolist()))
        for r in rng.integers(0, n, 6):
            o = int(ora.twoway("u8", a[r:r + 1], b, 0.8)[0][0])
            back = int(ora.twoway("u8", b[o:o + 1], a, 0.8)[0][0]) if o >= 0 else -1
            assert d.get(int(r), -1) == (o if (o >= 0 and back == r) else -1), (p, r)
    m.close()


def test_concurrent_callers_are_serialised(ora):
    """pairwise_match is const and called from an OpenMP loop in the reference
    (bundler_matching.cc:74): the handle serialises concurrent callers."""
    from concurrent.futures import ThreadPoolExecutor
    vs = synth.sift_views(14, 5, 600, noise="renorm")
    pairs = [tuple(p) for p in synth.all_pairs(5)]
    with matcher(vs) as m:
        with ThreadPoolExecutor(8) as ex:
            res = list(ex.map(lambda p: m.pairwise_match(int(p[0]), int(p[1])), pairs * 3))
        assert_clean(m)
    for (v1, v2), r in zip(pairs * 3, res):
        o12, o21 = ora.match_filtered("u8", vs[v1], vs[v2], 0.8)
        assert np.array_equal(r.matches_1_2, o12) and np.array_equal(r.matches_2_1, o21)


def test_many_pairs_one_launch(ora):
    """12 views x 2048: the persistent kernel over 66 pairs; every pair checked."""
    views = synth.sift_views(12, 12, 2048)
    pairs = synth.all_pairs(12)
    with matcher(views) as m:
        results, counts = m.match_pairs(pairs)
        assert_clean(m)
    for p, (v1, v2) in enumerate(pairs):
        o12, o21 = ora.match_filtered("u8", views[v1], views[v2], 0.8)
        assert np.array_equal(results[p].matches_1_2, o12), (v1, v2)
        assert np.array_equal(results[p].matches_2_1, o21), (v1, v2)


# ------------------------------------------------------------------ two-view gates

@pytest.mark.parametrize("lowres", [False, True])
def test_two_view_candidates_match_the_reference_gates(ora, lowres):
    """bundler::Matching::two_view_matching up to RANSAC for a batch of pairs: pair rules,
    low-res gate (only pairs with more than 10^6 feature products), match-count threshold and
    the correspondence lists, SIFT + SURF in the combined index space."""
    import oracle
    from orthosfm_b200 import TwoViewOptions
    sp, fp = synth.scene_pool(21, 500), synth.surf_pool(21, 200)
    other = synth.scene_pool(22, 500)
    sizes = [(1500, 300), (1400, 0), (900, 250), (1300, 200), (0, 0), (700, 0), (1200, 100)]
    sift = [synth.sift_view(21, v, s, sp if v not in (3, 5) else other, noise="renorm") for v, (s, _) in enumerate(sizes)]
    surf = [synth.surf_view(21, v, f, fp) if f else None for v, (_, f) in enumerate(sizes)]
    sift = [x if len(x) else None for x in sift]
    pairs = synth.all_pairs(len(sizes))
    opts = TwoViewOptions(use_lowres_matching=lowres, num_lowres_features=400, min_lowres_matches=12,
                          min_feature_matches=50, match_num_previous_frames=0)
    with matcher(sift, surf) as m:
        got = m.two_view_candidates(pairs, opts)
        prev = m.two_view_candidates(pairs, TwoViewOptions(match_num_previous_frames=2))
        assert_clean(m)
    want = oracle.two_view_candidates(ora, sift, surf, pairs, use_lowres_matching=lowres, num_lowres_features=400,
                                      min_lowres_matches=12, min_feature_matches=50)
    seen = set()
    for p, ((gs, gc, gij), (ws, wc, wij)) in enumerate(zip(got, want)):
        assert (gs, gc) == (ws, wc), (tuple(pairs[p]), gs, gc, ws, wc)
        assert np.array_equal(gij, wij), tuple(pairs[p])
        seen.add(gs)
    assert seen >= ({0, 1, 2, 3} if lowres else {0, 1, 3}), seen       # every outcome occurs
    want_prev = oracle.two_view_candidates(ora, sift, surf, pairs, match_num_previous_frames=2)
    assert [(a, b) for a, b, _ in prev] == [(a, b) for a, b, _ in want_prev]


# ------------------------------------------------------------------ tracks

def _random_match_lists(rng, feats, density):
    pairs, lists, off = [], [], [0]
    for v1 in range(1, len(feats)):
        for v2 in range(v1):
            if feats[v1] == 0 or feats[v2] == 0 or rng.random() > density:
                continue
            k = int(rng.integers(1, max(2, min(feats[v1], feats[v2]) // 3)))
            i = np.sort(rng.choice(feats[v1], k, replace=False))
            j = rng.choice(feats[v2], k, replace=False)
            pairs.append((v1, v2))
            lists.append(np.stack([i, j], 1))
            off.append(off[-1] + k)
    ij = np.concatenate(lists).astype(np.int32) if lists else np.zeros((0, 2), np.int32)
    return pairs, np.array(off, np.int64), ij


@pytest.mark.parametrize("seed,nviews,density", [(0, 6, 0.8), (1, 12, 0.5), (2, 30, 0.3), (3, 3, 1.0)])
def test_tracks_equal_the_reference_partition(seed, nviews, density):
    """Connected components of the match graph minus the components with two features of one
    view = bundler::Tracks::compute, up to the numbering of the tracks (compared after
    relabelling both in order of first appearance)."""
    import oracle
    rng = np.random.default_rng(seed)
    feats = rng.integers(0, 400, nviews)
    feats[rng.integers(0, nviews)] = 0
    pairs, off, ij = _random_match_lists(rng, feats, density)
    with ExhaustiveMatching() as m:
        got, nt, nconf = m.tracks_compute(feats, pairs, off, ij)
    want, nw = (oracle.Reference() if oracle.have_ref() else oracle).tracks_compute(feats, pairs, off, ij)
    assert nt == nw
    assert np.array_equal(got, oracle.canonical_track_ids(want))
    assert np.array_equal(got, oracle.canonical_track_ids(got))       # numbered by first feature
    if nviews >= 12:
        assert nconf > 0                                              # conflicts occurred and were dropped


def test_tracks_heavy_merging_is_deterministic():
    """Long chains and heavily contended unions (each feature matched in most pairs), three
    times over: the partition is the reference's every time."""
    import oracle
    rng = np.random.default_rng(11)
    nviews, nf = 40, 2500
    feats = np.full(nviews, nf)
    pairs, lists, off = [], [], [0]
    for v1 in range(1, nviews):
        for v2 in range(v1):
            k = 1500
            i = np.sort(rng.choice(nf, k, replace=False))
            # mostly "the same physical point" (same index) plus a few wrong matches that fuse
            # tracks and create conflicts
            j = i.copy()
            wrong = rng.random(k) < 0.002
            j[wrong] = rng.integers(0, nf, int(wrong.sum()))
            pairs.append((v1, v2))
            lists.append(np.stack([i, j], 1))
            off.append(off[-1] + k)
    ij = np.concatenate(lists).astype(np.int32)
    off = np.array(off, np.int64)
    want, nw = (oracle.Reference() if oracle.have_ref() else oracle).tracks_compute(feats, pairs, off, ij)
    want = oracle.canonical_track_ids(want)
    with ExhaustiveMatching() as m:
        for _ in range(3):
            got, nt, nconf = m.tracks_compute(feats, pairs, off, ij)
            assert nt == nw and nconf > 0
            assert np.array_equal(got, want)


def test_tracks_from_matcher_output(ora):
    """The whole chain on the device side of the seam: match lists of all pairs of six views
    (the matcher) -> tracks, against the oracle's tracks from the oracle's match lists."""
    import oracle
    vs = synth.sift_views(23, 6, 1500, noise="renorm")
    pairs = synth.all_pairs(6)
    with matcher(vs) as m:
        out = np.empty((len(pairs) * 1500, 2), np.int32)
        loff = m.match_pairs_lists(pairs, out)
        got, nt, _ = m.tracks_compute([1500] * 6, pairs, loff, out[:loff[-1]])
    lists = []
    for v1, v2 in pairs:
        o12, _ = ora.match_filtered("u8", vs[v1], vs[v2], 0.8)
        i = np.nonzero(o12 >= 0)[0]
        lists.append(np.stack([i, o12[i]], 1))
    off = np.concatenate([[0], np.cumsum([len(x) for x in lists])]).astype(np.int64)
    want, nw = oracle.tracks_compute([1500] * 6, pairs, off, np.concatenate(lists))
    assert nt == nw and nt > 100
    assert np.array_equal(got, oracle.canonical_track_ids(want))


def test_matcher_output_through_the_file_formats(ora, tmp_path):
    """Match lists -> prebundle.sfm (read back by the reference's own loader when it is built)
    and match lists -> tracks -> tracks.txt (equal to the restated writer on the oracle's
    tracks)."""
    import oracle
    from orthosfm_b200 import io as osio
    nv, n = 5, 1200
    vs = synth.sift_views(31, nv, n, noise="renorm")
    pairs = synth.all_pairs(nv)
    rng = np.random.default_rng(5)
    pos = (rng.random((nv * n, 2), dtype=np.float32) - 0.5)
    col = rng.integers(0, 256, (nv * n, 3), dtype=np.uint8)
    with matcher(vs) as m:
        out = np.empty((len(pairs) * n, 2), np.int32)
        loff = m.match_pairs_lists(pairs, out)
        ij = out[:loff[-1]]
        ids, nt, _ = m.tracks_compute([n] * nv, pairs, loff, ij)
    pre = str(tmp_path / "prebundle.sfm")
    osio.save_prebundle(pre, [n] * nv, pos, col, pairs, loff, ij)
    d = osio.load_prebundle(pre)
    assert np.array_equal(d["ij"], ij) and np.array_equal(d["offsets"], loff)
    if oracle.have_ref():
        counts, sums = oracle.Reference().load_prebundle_digest(pre)
        assert counts.tolist() == [nv, nv * n, len(pairs), len(ij)]
        pv = np.asarray(pairs, np.float64)
        assert sums[2] == float((1000 * pv[:, 0] + 7 * pv[:, 1]).sum() + (31.0 * ij[:, 0] + 17.0 * ij[:, 1]).sum())
    lists = []
    for v1, v2 in pairs:
        o12, _ = ora.match_filtered("u8", vs[v1], vs[v2], 0.8)
        i = np.nonzero(o12 >= 0)[0]
        lists.append(np.stack([i, o12[i]], 1))
    off = np.concatenate([[0], np.cumsum([len(x) for x in lists])]).astype(np.int64)
    want, nw = oracle.tracks_compute([n] * nv, pairs, off, np.concatenate(lists))
    txt = str(tmp_path / "tracks.txt")
    osio.save_tracks(txt, [n] * nv, ids, nt, pos, 2048.0, col)
    assert open(txt).read() == oracle.save_tracks_text(
        oracle.tracks_from_ids([n] * nv, oracle.canonical_track_ids(want), pos, 2048.0, col))


# ------------------------------------------------------------------ overlapped staging

def test_overlapped_staging_gives_the_same_results():
    """osfm_match_begin_overlapped: commit does not wait for the copies and the pair list is
    matched in phases as the views arrive.  Same lists, dense results, single pairs and gates
    as the plain cycle -- with SIFT and SURF, pairs in the reference's order, in reverse order
    (the first pair needs the last view) and over repeated cycles on one handle."""
    nv, n = 13, 1500
    sift = synth.sift_views(41, nv, n, noise="renorm")
    pool = synth.surf_pool(41, 300)
    surf = [synth.surf_view(41, v, 300 + 10 * v, pool) for v in range(nv)]
    surf[4] = surf[4][:0]
    pairs = [(a, b) for a in range(nv) for b in range(a)]
    cap = len(pairs) * (n + 500)
    with matcher(sift, surf) as m:
        out = np.empty((cap, 2), np.int32)
        want_off = m.match_pairs_lists(pairs, out).copy()
        want = out[:want_off[-1]].copy()
        want_rev_out = np.empty((cap, 2), np.int32)
        want_rev_off = m.match_pairs_lists(pairs[::-1], want_rev_out).copy()
        want_single = m.pairwise_match(nv - 1, 2)
        want_gates = m.two_view_candidates(pairs, TwoViewOptions(use_lowres_matching=True, num_lowres_features=300,
                                                                 min_lowres_matches=10, min_feature_matches=30))
    m = ExhaustiveMatching()
    try:
        for cycle in range(3):
            m.init(vps(sift, surf), overlap_copies=True)
            got = np.empty((cap, 2), np.int32)
            if cycle == 0:
                off = m.match_pairs_lists(pairs, got)
                assert np.array_equal(off, want_off) and np.array_equal(got[:off[-1]], want)
            elif cycle == 1:
                off = m.match_pairs_lists(pairs[::-1], got)
                assert np.array_equal(off, want_rev_off) and np.array_equal(got[:off[-1]], want_rev_out[:off[-1]])
            else:
                r = m.pairwise_match(nv - 1, 2)          # right after commit: needs the last view
                assert np.array_equal(r.matches_1_2, want_single.matches_1_2)
                assert np.array_equal(r.matches_2_1, want_single.matches_2_1)
                gates = m.two_view_candidates(pairs, TwoViewOptions(use_lowres_matching=True, num_lowres_features=300,
                                                                    min_lowres_matches=10, min_feature_matches=30))
                for (s1, c1, l1), (s2, c2, l2) in zip(want_gates, gates):
                    assert s1 == s2 and c1 == c2 and np.array_equal(l1, l2)
            m.wait_staged()
            assert_clean(m)
        m.init(vps(sift, surf))                          # and back to the plain cycle
        off = m.match_pairs_lists(pairs, got)
        assert np.array_equal(off, want_off) and np.array_equal(got[:off[-1]], want)
        from orthosfm_b200 import PackedViews            # every view staged by one osfm_match_set_views_q8 call
        packed = PackedViews(vps(sift, surf))
        for overlap in (False, True):
            m.init(packed, overlap_copies=overlap)
            off = m.match_pairs_lists(pairs, got)
            assert np.array_equal(off, want_off) and np.array_equal(got[:off[-1]], want)
        with pytest.raises(MatcherError):                # float descriptors are quantised on the device: plain cycle only
            m.init(vps([s.astype(np.float32) / 255 for s in sift]), overlap_copies=True)
    finally:
        m.close()


def test_lookahead_serves_the_pair_by_pair_loop(ora):
    """osfm_match_set_lookahead: the unchanged loop of bundler::Matching::compute
    (bundler_matching.cc:74-132: pairwise_match_lowres, then pairwise_match, pair by pair in the
    order of :92-93) gets the same results from batched passes over the pairs that follow."""
    nv, n = 9, 700
    sift = synth.sift_views(51, nv, n, noise="renorm")
    sift[3] = sift[3][:0]                                    # an empty view in the middle
    pool = synth.surf_pool(51, 200)
    surf = [synth.surf_view(51, v, 150 + 7 * v, pool) for v in range(nv)]
    pairs = [(a, b) for a in range(1, nv) for b in range(a)]
    with matcher(sift, surf) as m:
        want = [(m.pairwise_match_lowres(a, b, 300), m.pairwise_match(a, b)) for a, b in pairs]
        launches0 = m.stats()["kernel_launches"]
        for window in (5, 1000):
            m.set_lookahead(window)
            for k, (a, b) in enumerate(pairs):
                if k % 7 == 3:
                    continue                                 # the caller may skip pairs ...
                assert m.pairwise_match_lowres(a, b, 300) == want[k][0]
                r = m.pairwise_match(a, b)
                assert np.array_equal(r.matches_1_2, want[k][1].matches_1_2), (window, a, b)
                assert np.array_equal(r.matches_2_1, want[k][1].matches_2_1), (window, a, b)
            r = m.pairwise_match(2, 1)                       # ... or go back
            assert np.array_equal(r.matches_1_2, want[2][1].matches_1_2)
            r = m.pairwise_match(1, 2)                       # view_1 < view_2: not in the enumeration, direct
            assert np.array_equal(r.matches_1_2, want[2][1].matches_2_1)
        batched = m.stats()["kernel_launches"] - launches0
        m.set_lookahead(0)
        r = m.pairwise_match(5, 2)
        assert np.array_equal(r.matches_1_2, want[pairs.index((5, 2))][1].matches_1_2)
        assert_clean(m)
    assert batched < launches0                               # far fewer launches than pair by pair


def test_phase_times_add_up():
    views = synth.sift_views(52, 6, 2048, noise="renorm")
    with matcher(views) as m:
        out = np.empty((15 * 2048, 2), np.int32)
        m.match_pairs_lists(synth.all_pairs(6), out)
        st = m.stats()
    ph = st["last_phase_ms"]
    assert set(ph) == {"filter", "classify", "resolve_fwd", "claim", "resolve_rev", "mutual", "compact"}
    assert all(v >= 0 for v in ph.values()) and ph["filter"] > 0 and ph["resolve_fwd"] > 0 and ph["resolve_rev"] > 0
    assert abs(st["last_scan_ms"] - ph["filter"]) < 1e-9
    assert sum(ph.values()) <= st["last_total_ms"] * 1.02 + 0.05


# ------------------------------------------------------------------ RANSAC for the fundamental matrix

def _ransac_case(counts, seed, outliers=0.3):
    """Pairs of views with the given match counts: positions per view, pair lists."""
    rng = np.random.default_rng(seed)
    npairs = len(counts)
    feats, pos, pairs, lists = [], [], [], []
    for p, n in enumerate(counts):
        xy = synth.two_view_scene(seed * 100 + p, n, outliers)
        if p % 5 == 3:
            xy[:, 1] = xy[:, 0]; xy[:, 3] = xy[:, 2]            # a degenerate pair: every point on one line
        extra = int(rng.integers(0, 20))
        pa, pb = rng.permutation(n + extra), rng.permutation(n + extra)      # match k joins features pa[k], pb[k]
        va = np.zeros((n + extra, 2), np.float32); vb = np.zeros((n + extra, 2), np.float32)
        va[pa[:n]] = xy[:, :2]; vb[pb[:n]] = xy[:, 2:]
        order = np.argsort(pa[:n])                                # lists come sorted by the first feature
        lists.append(np.stack([pa[:n][order], pb[:n][order]], 1))
        feats += [n + extra, n + extra]
        pos += [va, vb]
        pairs.append((2 * p, 2 * p + 1))
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return np.array(feats, np.int32), np.concatenate(pos), np.array(pairs, np.int32), off, \
        np.concatenate(lists).astype(np.int32), npairs


@pytest.mark.parametrize("counts,iters", [([8, 9, 30, 200, 64, 1000], 300), ([2500, 12, 700], 1000)])
def test_ransac_fundamental_equals_reference(counts, iters):
    """osfm_ransac_draw_samples + osfm_ransac_fundamental against the reference's own
    RansacFundamental::estimate run pair after pair on one std::rand() sequence
    (ransac_fundamental.cc:26-105 as called from bundler_matching.cc:176-220): identical
    inlier lists and fundamental matrices."""
    import oracle
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    ref = oracle.Reference()
    feats, pos, pairs, off, ij, npairs = _ransac_case(counts, 3)
    base = np.concatenate([[0], np.cumsum(feats)])
    oracle.srand(11)
    want = []
    for p in range(npairs):
        l = ij[off[p]:off[p + 1]]
        xy = np.concatenate([pos[base[pairs[p, 0]] + l[:, 0]], pos[base[pairs[p, 1]] + l[:, 1]]], 1)
        want.append(ref.ransac(xy, iters, 0.0015))
    oracle.srand(11)
    with matcher(synth.sift_views(1, 2, 64)) as m:
        ooff, oij, F = m.ransac_fundamental(feats, pos, pairs, off, ij, max_iterations=iters, threshold=0.0015)
    assert sum(len(w[0]) for w in want) > 0
    for p in range(npairs):
        inl, wF = want[p]
        assert np.array_equal(oij[ooff[p]:ooff[p + 1]], ij[off[p]:off[p + 1]][inl]), p
        if len(inl):
            assert np.array_equal(F[p].ravel(), wF), p
        else:
            assert not F[p].any()


@pytest.mark.parametrize("chunk_pairs,own_samples", [(1, False), (2, False), (4, True)])
def test_ransac_fundamental_in_chunks_equals_reference(chunk_pairs, own_samples, monkeypatch):
    """The call works through the pairs in chunks (draws of one chunk on the host while the
    device runs the one before; per-fit scratch reused): forced to 1, 2 and 4 pairs per chunk
    over 7 pairs, samples drawn inside or handed in."""
    import oracle
    from orthosfm_b200 import ransac_draw_samples
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    ref = oracle.Reference()
    feats, pos, pairs, off, ij, npairs = _ransac_case([30, 9, 120, 64, 8, 300, 45], 13)
    base = np.concatenate([[0], np.cumsum(feats)])
    iters = 120
    oracle.srand(4)
    want = []
    for p in range(npairs):
        l = ij[off[p]:off[p + 1]]
        xy = np.concatenate([pos[base[pairs[p, 0]] + l[:, 0]], pos[base[pairs[p, 1]] + l[:, 1]]], 1)
        want.append(ref.ransac(xy, iters, 0.0015))
    monkeypatch.setenv("OSFM_RANSAC_CHUNK_PAIRS", str(chunk_pairs))
    oracle.srand(4)
    smp = ransac_draw_samples(off, iters) if own_samples else None
    with matcher(synth.sift_views(1, 2, 64)) as m:
        ooff, oij, F = m.ransac_fundamental(feats, pos, pairs, off, ij, samples=smp, max_iterations=iters)
    for p in range(npairs):
        inl, wF = want[p]
        assert np.array_equal(oij[ooff[p]:ooff[p + 1]], ij[off[p]:off[p + 1]][inl]), p
        if len(inl):
            assert np.array_equal(F[p].ravel(), wF), p


def test_ransac_fundamental_with_non_finite_positions_ends_and_equals_reference():
    """A NaN, an infinite and a huge position: the iteration kernel must end (its loops are
    bounded whatever the data) and the result is still the reference's."""
    import oracle
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    ref = oracle.Reference()
    feats, pos, pairs, off, ij, npairs = _ransac_case([60, 40, 90], 7)
    pos = pos.copy()
    base = np.concatenate([[0], np.cumsum(feats)])
    pos[base[0] + ij[3, 0]] = np.nan
    pos[base[3] + ij[off[1] + 5, 1], 0] = np.inf
    pos[base[4] + ij[off[2] + 7, 0]] = 3e30
    oracle.srand(2)
    want = []
    for p in range(npairs):
        l = ij[off[p]:off[p + 1]]
        xy = np.concatenate([pos[base[pairs[p, 0]] + l[:, 0]], pos[base[pairs[p, 1]] + l[:, 1]]], 1)
        want.append(ref.ransac(xy, 150, 0.0015))
    oracle.srand(2)
    with matcher(synth.sift_views(1, 2, 64)) as m:
        ooff, oij, F = m.ransac_fundamental(feats, pos, pairs, off, ij, max_iterations=150)
    for p in range(npairs):
        assert np.array_equal(oij[ooff[p]:ooff[p + 1]], ij[off[p]:off[p + 1]][want[p][0]]), p


def test_ransac_fundamental_argument_errors():
    feats, pos, pairs, off, ij, npairs = _ransac_case([20, 7], 5)
    with matcher(synth.sift_views(1, 2, 64)) as m:
        with pytest.raises(MatcherError):          # a pair with fewer than 8 matches (the reference throws)
            m.ransac_fundamental(feats, pos, pairs, off, ij, max_iterations=10)
        feats, pos, pairs, off, ij, npairs = _ransac_case([20, 9], 5)
        smp = np.zeros((2, 10, 8), np.int32)       # not eight distinct ascending indices
        with pytest.raises(MatcherError):
            m.ransac_fundamental(feats, pos, pairs, off, ij, samples=smp, max_iterations=10)
        bad = ij.copy(); bad[3, 1] = 10 ** 6       # a match outside its view
        with pytest.raises(MatcherError):
            m.ransac_fundamental(feats, pos, pairs, off, bad, max_iterations=10)
        ooff, oij, F = m.ransac_fundamental(feats, pos, pairs[:0], off[:1], ij[:0], max_iterations=10)
        assert ooff.tolist() == [0] and len(oij) == 0


@pytest.mark.parametrize("lowres", [False, True])
def test_two_view_stage_equals_reference_bundler_matching(lowres):
    """osfm_match_two_view over all pairs in the order of bundler::Matching::compute against
    the reference's own bundler::Matching (init + compute, bundler_matching.cc:45-220) on the
    same viewports and the same std::rand() seed: the same pairs survive, with the same
    inlier matches."""
    import oracle
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    nv, n = 7, 1200
    descs, poss = synth.sfm_scene(4, nv, n, 900, visible=0.45)
    descs[5] = descs[5][:0]; poss[5] = poss[5][:0]              # a view without features
    descs[6] = _unrelated(descs[6])                              # a view that matches nothing
    empty = np.zeros((0, 64), np.float32)
    rx = oracle.Reference().exhaustive([(d.astype(np.float32) / 255.0, empty) for d in descs])
    kw = dict(use_lowres_matching=lowres, num_lowres_features=400, min_lowres_matches=12, min_feature_matches=50,
              min_matching_inliers=30, ransac_max_iterations=250, ransac_threshold=0.0015)
    want = rx.bundler_compute(np.concatenate(poss), seed=5, **kw)
    pairs = [(a, b) for a in range(nv) for b in range(a)]       # i -> (view_1, view_2), bundler_matching.cc:92-93
    opts = TwoViewOptions(**kw)
    oracle.srand(5)
    with matcher(descs) as m:
        got = m.two_view_matching(pairs, np.concatenate(poss), opts)
    accepted = [(a, b, ij) for (a, b), (st, cnt, ij) in zip(pairs, got) if st == TWO_VIEW_OK]
    assert len(want) >= 8 and len(accepted) == len(want)
    for (wa, wb, wij), (a, b, ij) in zip(want, accepted):
        assert (wa, wb) == (a, b)
        assert np.array_equal(wij, ij), (a, b)
    statuses = {st for st, _, _ in got}
    assert TWO_VIEW_SKIPPED in statuses and (TWO_VIEW_TOO_FEW_MATCHES in statuses or TWO_VIEW_LOWRES_REJECTED in statuses)


def test_two_view_stage_with_sift_and_surf_equals_reference():
    """The same with SIFT and SURF features: the lists, the positions and the RANSAC samples
    live in the combined [sift..., surf...] index space (matching.cc:50-89)."""
    import oracle
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    nv = 5
    sift, surf, poss = synth.sfm_scene(8, nv, 900, 600, visible=0.5, surf_n=400)
    rx = oracle.Reference().exhaustive([(d.astype(np.float32) / 255.0, s.astype(np.float32) / np.float32(127.0))
                                        for d, s in zip(sift, surf)])
    kw = dict(min_feature_matches=40, min_matching_inliers=25, ransac_max_iterations=200)
    want = rx.bundler_compute(np.concatenate(poss), seed=9, **kw)
    pairs = [(a, b) for a in range(nv) for b in range(a)]
    oracle.srand(9)
    with matcher(sift, surf) as m:
        got = m.two_view_matching(pairs, np.concatenate(poss), TwoViewOptions(**kw))
    accepted = [(a, b, ij) for (a, b), (st, cnt, ij) in zip(pairs, got) if st == TWO_VIEW_OK]
    assert len(want) == len(pairs) == len(accepted)
    for (wa, wb, wij), (a, b, ij) in zip(want, accepted):
        assert (wa, wb) == (a, b) and np.array_equal(wij, ij), (a, b)
        assert (ij[:, 0] >= 900).any() and (ij[:, 0] < 900).any()      # SURF and SIFT matches among the inliers


def test_whole_chain_descriptors_to_tracks_equals_reference():
    """Descriptors and positions in, feature tracks out: osfm_match_two_view ->
    osfm_tracks_compute against the reference's bundler::Matching::compute ->
    bundler::Tracks::compute on the same std::rand() seed (what calculateTracksUsingMVE runs,
    src/matching/matching_mve.cpp:405-452): the same tracks."""
    import oracle
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    nv, n = 6, 1000
    descs, poss = synth.sfm_scene(21, nv, n, 800, visible=0.5)
    empty = np.zeros((0, 64), np.float32)
    ref = oracle.Reference()
    rx = ref.exhaustive([(d.astype(np.float32) / 255.0, empty) for d in descs])
    kw = dict(min_feature_matches=50, min_matching_inliers=30, ransac_max_iterations=200)
    want_pairs = rx.bundler_compute(np.concatenate(poss), seed=3, **kw)
    wp = np.array([(a, b) for a, b, _ in want_pairs], np.int32)
    woff = np.concatenate([[0], np.cumsum([len(ij) for _, _, ij in want_pairs])]).astype(np.int64)
    want_ids, want_nt = ref.tracks_compute([n] * nv, wp, woff, np.concatenate([ij for _, _, ij in want_pairs]))
    pairs = [(a, b) for a in range(nv) for b in range(a)]
    oracle.srand(3)
    with matcher(descs) as m:
        got = m.two_view_matching(pairs, np.concatenate(poss), TwoViewOptions(**kw))
        acc = [(p, ij) for p, (st, _, ij) in zip(pairs, got) if st == TWO_VIEW_OK]
        off = np.concatenate([[0], np.cumsum([len(ij) for _, ij in acc])]).astype(np.int64)
        ids, nt, _ = m.tracks_compute([n] * nv, [p for p, _ in acc], off, np.concatenate([ij for _, ij in acc]))
    assert nt == want_nt and nt > 300
    assert np.array_equal(ids, oracle.canonical_track_ids(want_ids))


def _unrelated(desc):
    rng = np.random.default_rng(99)
    return synth._normalise_clamp_quantise(np.abs(rng.standard_normal(desc.shape, dtype=np.float32)))


# ------------------------------------------------------------------ the reference-side binding

def test_reference_side_binding():
    """oracle/_ref/shim_check is the C++ subclass of sfm::MatchingBase
    (orthosfm_b200/csrc/gpu_exhaustive_matching.h) compiled against the reference's own
    headers and run, in one process, against the reference's own sfm::ExhaustiveMatching on
    the same bundler::ViewportList (float SIFT + SURF descriptors, views without SIFT or
    without SURF included).  Built only where /root/reference exists; the binary travels."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "shim_check")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/shim_check not built (needs /root/reference)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "SHIM_CHECK PASS" in r.stdout, r.stdout
    # the batched drop-in for bundler::Matching (gpu_bundler_matching.h) against the reference's own
    assert "BUNDLER_CHECK PASS" in r.stderr, r.stderr


# ------------------------------------------------------------------ several GPUs behind one handle

def _device_count():
    import torch
    return torch.cuda.device_count()


def test_multi_device_handle_equals_single_device():
    """osfm_match_create_multi: the pool replicated by ncclBroadcast, every batched call cut into
    one range of pairs per device.  Same dense results, lists, gates and two-view stage as one
    device.  Skipped on a one-GPU box."""
    if _device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    import oracle
    devs = list(range(min(_device_count(), 4)))
    nv = 9
    sift, surf, pos = synth.sfm_scene(61, nv, 1400, 900, surf_n=260)
    sift[2] = sift[2][:300]                                      # unequal costs: the ranges differ in length
    pos[2] = np.concatenate([pos[2][:300], pos[2][1400:]])
    surf[5] = surf[5][:0]
    pos[5] = pos[5][:1400]
    positions = np.concatenate(pos)
    pairs = synth.all_pairs(nv)
    cap = len(pairs) * 2000
    opts = TwoViewOptions(use_lowres_matching=True, num_lowres_features=300, min_lowres_matches=8,
                          min_feature_matches=30, min_matching_inliers=20, ransac_max_iterations=100)

    def run(m):
        out = {}
        res, counts = m.match_pairs(pairs)
        out["dense"] = [(r.matches_1_2.copy(), r.matches_2_1.copy()) for r in res]
        out["counts"] = counts.copy()
        lists = np.empty((cap, 2), np.int32)
        off = m.match_pairs_lists(pairs, lists)
        out["lists"] = (off.copy(), lists[:off[-1]].copy())
        out["gates"] = m.two_view_candidates(pairs, opts)
        oracle.srand(3)
        out["two_view"] = m.two_view_matching(pairs, positions, opts)
        out["single"] = m.pairwise_match(7, 3)
        return out

    with ExhaustiveMatching() as m1:
        m1.init(vps(sift, surf))
        want = run(m1)
    with ExhaustiveMatching(devices=devs) as mm:
        assert mm.num_devices == len(devs)
        for cycle in range(2):                                   # re-staging on a multi-device handle
            mm.init(vps(sift, surf), overlap_copies=(cycle == 1))
            got = run(mm)
            assert_clean(mm)
            for (a12, a21), (b12, b21) in zip(want["dense"], got["dense"]):
                assert np.array_equal(a12, b12) and np.array_equal(a21, b21)
            assert np.array_equal(want["counts"], got["counts"])
            assert np.array_equal(want["lists"][0], got["lists"][0]) and np.array_equal(want["lists"][1], got["lists"][1])
            for key in ("gates", "two_view"):
                for (s1, c1, l1), (s2, c2, l2) in zip(want[key], got[key]):
                    assert s1 == s2 and c1 == c2 and np.array_equal(l1, l2), key
            assert np.array_equal(want["single"].matches_1_2, got["single"].matches_1_2)
    with pytest.raises(MatcherError):
        ExhaustiveMatching(devices=[0, 0])


def test_reference_side_binding_on_two_devices():
    """shim_check again with OSFM_SHIM_DEVICES=0,1: sfm::GpuExhaustiveMatching and
    sfm::bundler::GpuMatching over a two-device matcher against the reference's own classes."""
    import os
    import subprocess
    if _device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "shim_check")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/shim_check not built (needs /root/reference)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600, env=dict(os.environ, OSFM_SHIM_DEVICES="0,1"))
    assert r.returncode == 0, r.stdout + r.stderr
    assert "SHIM_CHECK PASS" in r.stdout, r.stdout
    assert "BUNDLER_CHECK PASS" in r.stderr, r.stderr


def test_reference_gpu_matcher_baseline_runs_and_agrees_on_clear_matches():
    """The reference's own GPU matcher (CudaSift MatchSiftData / FindMaxCorr10, unmodified
    sources compiled for sm_100 into oracle/_ref/libcudasift_ref.so) is timed beside the product by
    bench.py.  Here: it runs on this GPU, and where the product accepts a match (clear ratio) its
    float arg-max names the same candidate."""
    import oracle
    if not oracle.have_cudasift():
        pytest.skip("oracle/_ref/libcudasift_ref.so not built (needs /root/reference)")
    vs = synth.sift_views(71, 2, 2048, noise="renorm")       # a multiple of 32: its tail handling plays no part
    mean_ms, min_ms, match, score, amb = oracle.CudaSiftReference().match(vs[1], vs[0], reps=2)
    assert 0 < min_ms <= mean_ms
    with matcher(vs) as m:
        ours = m.twoway_match(KIND_SIFT_U8, 1, 0).matches_1_2
    ok = ours >= 0
    assert ok.sum() > 200
    assert (match[ok] == ours[ok]).mean() > 0.99
