"""CPU tests: the C restatement (oracle/) against the golden vectors generated from the
reference itself, and -- where oracle/_ref exists -- directly against the reference on
fresh seeded inputs.  This is what pins the oracle (the reference
ships no golden vectors of its own)."""
import numpy as np
import pytest

from orthosfm_b200 import synth

CASE_NAMES = [
    "u8.synth_700x700", "u8.synth_300x650", "u8.synth_ratio1", "u8.duplicates",
    "u8.random_bytes", "u8.straddle_65536", "u8.single_candidate", "u8.single_query",
    "u8.zeros", "u8.lane_wrap", "s8.synth_500x640", "s8.random_bytes", "s8.big_positive",
    "s8.all_negative", "f32.synth_200x260",
]


def test_golden_case_list_is_complete(golden_cases):
    _, names = golden_cases
    assert sorted(names) == sorted(CASE_NAMES)


@pytest.mark.parametrize("name", CASE_NAMES)
def test_oracle_matches_golden(ora, golden_cases, name):
    z, _ = golden_cases
    kind = name.split(".")[0]
    a, b, ratio = z[name + ".a"], z[name + ".b"], float(z[name + ".ratio"])
    t12, t21 = ora.twoway(kind, a, b, ratio)
    assert np.array_equal(t12, z[name + ".t12"])
    assert np.array_equal(t21, z[name + ".t21"])
    f12, f21 = ora.remove_inconsistent(t12, t21)
    assert np.array_equal(f12, z[name + ".f12"])
    assert np.array_equal(f21, z[name + ".f21"])
    assert ora.count_consistent(t12, t21) == int(z[name + ".count"])
    # count of the unfiltered result == survivors of the mutual filter
    assert int((f12 >= 0).sum()) == int(z[name + ".count"])


def test_oracle_combine_results_golden(ora, golden_cases):
    z, _ = golden_cases
    c12, c21 = ora.combine_results(z["combine.s12"], z["combine.s21"], z["combine.f12"], z["combine.f21"])
    assert np.array_equal(c12, z["combine.c12"])
    assert np.array_equal(c21, z["combine.c21"])


def test_oracle_real_image_pair(ora, golden_real):
    g = golden_real
    t12, t21 = ora.twoway("u8", g["sift_1"], g["sift_0"], 0.8)
    assert np.array_equal(t12, g["twoway_12"]) and np.array_equal(t21, g["twoway_21"])
    empty = np.zeros((0, 64), np.int8)
    m12, m21 = ora.pairwise_match(g["sift_1"], g["sift_0"], empty, empty)
    assert np.array_equal(m12, g["match_12"]) and np.array_equal(m21, g["match_21"])
    assert int((m12 >= 0).sum()) == 810
    assert ora.pairwise_match_lowres(g["sift_1"], g["sift_0"], empty, empty, 500) == int(g["lowres_500"]) == 187


def test_oracle_three_image_set(ora, golden_triple):
    """BASELINE config 1 restated: 3 real images, all 3 pairs, against the reference's results."""
    g = golden_triple
    empty = np.zeros((0, 64), np.int8)
    counts = {}
    for v1 in range(1, 3):
        for v2 in range(v1):
            m12, m21 = ora.pairwise_match(g[f"sift_{v1}"], g[f"sift_{v2}"], empty, empty)
            assert np.array_equal(m12, g[f"match_{v1}{v2}_12"]) and np.array_equal(m21, g[f"match_{v1}{v2}_21"])
            assert ora.pairwise_match_lowres(g[f"sift_{v1}"], g[f"sift_{v2}"], empty, empty, 500) == int(g[f"lowres_{v1}{v2}"])
            counts[(v1, v2)] = int((m12 >= 0).sum())
    assert counts == {(1, 0): 810, (2, 0): 810, (2, 1): 2376}


def test_oracle_quantiser_golden(ora, golden_real):
    g = golden_real
    assert np.array_equal(ora.quantize_sift(g["float_sample"]), g["float_sample_q"])


def test_quantiser_rounding_rule(ora):
    # convert_descriptor: clamp, *255, round half away from zero, truncate to uchar
    x = np.zeros((1, 128), np.float32)
    x[0, :8] = [0.0, 1.0, 2.0, -0.3, 0.5 / 255, 1.5 / 255, 0.49999 / 255, 254.5 / 255]
    q = ora.quantize_sift(x)[0, :8]
    assert q.tolist() == [0, 255, 255, 0, 1, 2, 0, 255]
    s = np.zeros((1, 64), np.float32)
    s[0, :6] = [-1.0, 1.0, -2.0, 0.5 / 127, -0.5 / 127, -1.5 / 127]
    assert ora.quantize_surf(s)[0, :6].tolist() == [-127, 127, -127, 1, -1, -2]


def test_oracle_edge_semantics(ora):
    """Appendix A of SURVEY.md, each verified against the compiled reference there."""
    v = synth.sift_views(3, 1, 64)[0]
    # two identical perfect candidates: the later index wins, d1 = d2 -> ratio 1 > 0.64 rejects
    # unless d = 0 (0/0 = NaN accepts)
    q = np.full((1, 128), 0, np.uint8)
    q[0, :16] = 63                      # |q|^2 = 63504 < 65536: no wrap
    cands = np.concatenate([v[:5], q, v[5:9], q])
    d1, d2, i1, i2 = ora.nn("u8", q[0], cands)
    assert (i1, i2) == (10, 5) and d1 == d2 == 2 * (65025 - 63504)
    # single candidate: second best stays 0 -> d2 = 65534
    d1, d2, i1, _ = ora.nn("u8", q[0], q)
    assert d2 == 65534.0 and i1 == 0
    # lane wrap: 16 x 64*64 in one lane is seen as 0
    w = np.zeros((1, 128), np.uint8)
    w[0, 0::8] = 64
    d1, _, _, _ = ora.nn("u8", w[0], w)
    assert d1 == 65534.0
    # empty sets
    m12, m21 = ora.twoway("u8", v[:0], v, 0.8)
    assert m12.size == 0 and (m21 == -1).all()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_oracle_vs_reference_fresh_inputs(ora, ref, seed):
    rng = np.random.default_rng(100 + seed)
    n1, n2 = int(rng.integers(1, 400)), int(rng.integers(1, 400))
    vs = synth.sift_views(10 + seed, 2, max(n1, n2))
    cases = [("u8", vs[0][:n1], vs[1][:n2], 0.8),
             ("u8", rng.integers(0, 256, (n1, 128), dtype=np.uint8),
              rng.integers(0, 256, (n2, 128), dtype=np.uint8), 0.8),
             ("u8", rng.integers(0, 48, (n1, 128), dtype=np.uint8),
              rng.integers(0, 48, (n2, 128), dtype=np.uint8), 0.9),
             ("s8", rng.integers(-127, 128, (n1, 64), dtype=np.int8),
              rng.integers(-127, 128, (n2, 64), dtype=np.int8), 0.7),
             ("s8", synth.surf_view(seed, 0, n1), synth.surf_view(seed, 1, n2), 0.7)]
    fa = rng.standard_normal((n1, 128)).astype(np.float32)
    fb = rng.standard_normal((n2, 128)).astype(np.float32)
    cases.append(("f32", fa / np.linalg.norm(fa, axis=1, keepdims=True),
                  fb / np.linalg.norm(fb, axis=1, keepdims=True), 0.8))
    for kind, a, b, ratio in cases:
        o = ora.twoway(kind, a, b, ratio)
        r = ref.twoway(kind, a, b, ratio)
        assert np.array_equal(o[0], r[0]) and np.array_equal(o[1], r[1]), kind
        assert ora.nn(kind, a[0], b) == ref.nn(kind, a[0], b)
        of, rf = ora.remove_inconsistent(*o), ref.remove_inconsistent(*r)
        assert np.array_equal(of[0], rf[0]) and np.array_equal(of[1], rf[1])


def test_oracle_pairwise_match_vs_reference_plugin(ora, ref):
    """ExhaustiveMatching::{init,pairwise_match,pairwise_match_lowres} on float input,
    including a view without SIFT features (the combine_results size quirk)."""
    rng = np.random.default_rng(7)

    def sift_f(n):
        x = np.abs(rng.standard_normal((n, 128))).astype(np.float32)
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        return np.minimum(x, 0.2) / np.linalg.norm(np.minimum(x, 0.2), axis=1, keepdims=True)

    def surf_f(n):
        x = rng.standard_normal((n, 64)).astype(np.float32)
        return x / np.linalg.norm(x, axis=1, keepdims=True)

    views = [(sift_f(300), surf_f(120)), (sift_f(350), surf_f(90)), (sift_f(0), surf_f(80)),
             (sift_f(200), surf_f(0))]
    views[1][0][:100] = views[0][0][:100]
    views[1][1][:40] = views[0][1][:40]
    views[2][1][:30] = views[0][1][:30]
    ex = ref.exhaustive(views)
    q = [(ora.quantize_sift(s), ora.quantize_surf(f)) for s, f in views]
    for v1, v2 in [(1, 0), (0, 1), (2, 0), (0, 2), (3, 0), (0, 3), (2, 3), (3, 2)]:
        r12, r21 = ex.pairwise_match(v1, v2)
        o12, o21 = ora.pairwise_match(q[v1][0], q[v2][0], q[v1][1], q[v2][1])
        assert np.array_equal(o12, r12) and np.array_equal(o21, r21), (v1, v2)
        assert ora.pairwise_match_lowres(q[v1][0], q[v2][0], q[v1][1], q[v2][1], 150) == \
            ex.pairwise_match_lowres(v1, v2, 150)


def test_tracks_restatement_equals_reference(ref):
    """oracle.tracks_compute (the restatement of bundler::Tracks::compute) against the
    reference's own code through ref_driver.cc: same track ids, feature by feature."""
    import oracle
    rng = np.random.default_rng(7)
    feats = np.array([60, 45, 70, 0, 38, 52])
    pairs, lists, off = [], [], [0]
    for v1 in range(1, len(feats)):
        for v2 in range(v1):
            if feats[v1] == 0 or feats[v2] == 0 or rng.random() < 0.2:
                continue
            k = int(rng.integers(3, 30))
            i = np.sort(rng.choice(feats[v1], k, replace=False))
            j = rng.choice(feats[v2], k, replace=False)
            pairs.append((v1, v2))
            lists.append(np.stack([i, j], 1))
            off.append(off[-1] + k)
    ij = np.concatenate(lists).astype(np.int32)
    off = np.array(off, np.int64)
    a, na = ref.tracks_compute(feats, pairs, off, ij)
    b, nb = oracle.tracks_compute(feats, pairs, off, ij)
    assert na == nb and np.array_equal(a, b)
    assert na > 10 and (a >= 0).sum() > 2 * na - 1
    c = oracle.canonical_track_ids(a)
    assert c.max() + 1 == na and np.array_equal(oracle.canonical_track_ids(c), c)


# ------------------------------------------------------------------ RANSAC-F arithmetic (host build of the device code)

def _hostcheck():
    import oracle
    if not (oracle.have_ref() and oracle.RansacHostCheck.available()):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return oracle.Reference(), oracle.RansacHostCheck()


def test_svd_restatement_equals_reference_svd():
    """orthosfm_b200/csrc/ransac_math.cuh compiled for the host gives the doubles of the
    reference's math::matrix_svd (matrix_svd.h), 9 x 9 and 3 x 3, random, sparse, rank
    deficient and repeated-value inputs."""
    ref, hc = _hostcheck()
    rng = np.random.default_rng(0)
    for t in range(600):
        a = rng.standard_normal((9, 9))
        if t % 3 == 0:
            a[8] = 0                                  # the padded 8 x 9 case
        if t % 7 == 0:
            a.flat[rng.integers(0, 81, 30)] = 0
        if t % 11 == 0:
            a[3] = a[1]; a[:, 5] = a[:, 2]            # rank deficient
        if t % 13 == 0:
            a = np.round(a)                           # many exact zeros and ties
        s1, v1 = ref.svd(a)
        s2, v2 = hc.svd(a)
        assert np.array_equal(s1, s2) and np.array_equal(v1, v2), t
    for t in range(600):
        a = rng.standard_normal((3, 3))
        if t % 5 == 0:
            a.flat[rng.integers(0, 9, 3)] = 0
        if t % 7 == 0:
            a[2] = a[0]
        if t % 11 == 0:
            a = np.diag(rng.standard_normal(3))
        for x, y in zip(ref.svd(a), hc.svd(a)):
            assert np.array_equal(x, y), t
    for a in (np.zeros((3, 3)), np.eye(3), np.zeros((9, 9)), np.eye(9), np.ones((9, 9))):
        for x, y in zip(ref.svd(a), hc.svd(a)):
            assert np.array_equal(x, y)


def test_fundamental_and_sampson_equal_reference():
    """fundamental_8_point + enforce_fundamental_constraints and sampson_distance
    (fundamental.cc:78-126, 225-247): same doubles, degenerate samples included."""
    ref, hc = _hostcheck()
    rng = np.random.default_rng(1)
    stopped_early = 0
    for t in range(1500):
        m = synth.two_view_scene(100 + t, 8, outlier_fraction=0.0 if t % 2 else 0.4).astype(np.float64)
        if t % 9 == 0:
            m[5] = m[2]                               # a repeated correspondence
        if t % 10 == 0:
            m[:, 1] = m[:, 0]; m[:, 3] = m[:, 2]      # all points on a line
        if t % 25 == 0:
            m[:] = m[0]                               # one point eight times
        F1 = ref.fundamental(m[:, :2], m[:, 2:])
        F2 = hc.fundamental(m[:, :2], m[:, 2:])
        assert np.array_equal(F1, F2, equal_nan=True), t
        F3, trips = hc.fundamental_staged(m[:, :2], m[:, 2:])          # the device's route, fixed-point stop
        assert np.array_equal(F1, F3, equal_nan=True), t
        stopped_early += trips < 81
        q = rng.uniform(-0.5, 0.5, 4)
        d1, d2 = ref.sampson(F1, q), hc.sampson(F2, q)
        assert d1 == d2 or (np.isnan(d1) and np.isnan(d2)), t
    assert stopped_early > 1200       # the stop matters: without it two thirds of the samples run 81 trips


@pytest.mark.parametrize("n,outliers", [(8, 0.0), (9, 0.5), (40, 0.3), (400, 0.3), (1500, 0.6)])
def test_ransac_restatement_and_sample_draws_equal_reference(n, outliers):
    """srand(s); the reference's RansacFundamental::estimate  ==  srand(s);
    osfm_ransac_draw_samples + the device arithmetic (host build): same inliers, same F.
    Two pairs back to back share the one rand() sequence."""
    import oracle
    from orthosfm_b200 import ransac_draw_samples
    ref, hc = _hostcheck()
    a = synth.two_view_scene(n, n, outliers).astype(np.float64)
    b = synth.two_view_scene(n + 1, n + 3, outliers).astype(np.float64)
    iters = 200
    oracle.srand(7)
    want = [ref.ransac(a, iters, 0.0015), ref.ransac(b, iters, 0.0015)]
    oracle.srand(7)
    smp = ransac_draw_samples(np.array([0, len(a), len(a) + len(b)], np.int64), iters)
    assert (np.diff(smp, axis=2) > 0).all() and smp.min() >= 0
    assert smp[0].max() < len(a) and smp[1].max() < len(b)
    got = [hc.ransac(a, smp[0], 0.0015), hc.ransac(b, smp[1], 0.0015)]
    for (wi, wF), (gi, gF) in zip(want, got):
        assert np.array_equal(wi, gi)
        if len(wi):
            assert np.array_equal(wF, gF)
    if n >= 40 and outliers <= 0.3:
        assert len(want[0][0]) >= 0.5 * (1 - outliers) * n      # the scene is a real two-view geometry


def test_sample_draws_leave_the_rand_sequence_where_the_reference_would():
    """osfm_ransac_draw_samples reads the process-wide rand() sequence without glibc's lock
    (setstate / random_r); whoever draws next must continue as if rand() had been called:
    reference(A), reference(B)  ==  ours(A), reference(B), and plain rand() agrees as well."""
    import ctypes
    import oracle
    from orthosfm_b200 import ransac_draw_samples
    ref, hc = _hostcheck()
    libc = ctypes.CDLL(None)
    a = synth.two_view_scene(1, 300, 0.3).astype(np.float64)
    b = synth.two_view_scene(2, 200, 0.3).astype(np.float64)
    for seed in (0, 1, 12345):
        oracle.srand(seed)
        ref.ransac(a, 50, 0.0015)
        want_b = ref.ransac(b, 50, 0.0015)
        want_next = [libc.rand() for _ in range(5)]
        oracle.srand(seed)
        smp = ransac_draw_samples(np.array([0, len(a)], np.int64), 50)
        got_b = ref.ransac(b, 50, 0.0015)
        got_next = [libc.rand() for _ in range(5)]
        assert np.array_equal(want_b[0], got_b[0]) and np.array_equal(want_b[1], got_b[1])
        assert want_next == got_next
        # the first sample is the first eight distinct values of rand() % n, sorted
        oracle.srand(seed)
        seen = []
        while len(seen) < 8:
            v = libc.rand() % len(a)
            if v not in seen:
                seen.append(v)
        assert sorted(seen) == smp[0, 0].tolist()


def test_fundamental_on_non_finite_and_extreme_input_equals_reference():
    """NaN, infinity, 1e150 and 1e-160 scales: the loop must end and the doubles (NaNs
    included) must be the reference's, on the plain and on the staged / early-stop route."""
    ref, hc = _hostcheck()
    rng = np.random.default_rng(0)
    for t in range(200):
        m = synth.two_view_scene(t, 8, 0.0).astype(np.float64)
        kind = t % 4
        if kind == 0:
            m[rng.integers(8), rng.integers(4)] = np.nan
        elif kind == 1:
            m[rng.integers(8), rng.integers(4)] = np.inf
        elif kind == 2:
            m *= 1e150
        else:
            m *= 1e-160
        F1 = ref.fundamental(m[:, :2], m[:, 2:])
        F2 = hc.fundamental(m[:, :2], m[:, 2:])
        F3, _ = hc.fundamental_staged(m[:, :2], m[:, 2:])
        assert np.array_equal(F1, F2, equal_nan=True) and np.array_equal(F1, F3, equal_nan=True), (t, kind)


def test_reference_large_pair_entry_equals_the_pairwise_entry():
    """osfm_ref_match_large_pair_u8 (the reference's oneway_match on chunks of the query rows, over
    OpenMP threads) against osfm_ref_match_pairs_u8_digest (its twoway_match on the whole sets):
    same count, digest and vectors, with set sizes that are no multiple of the chunk."""
    import oracle
    from orthosfm_b200 import synth
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    ref = oracle.Reference()
    v = synth.sift_views(41, 2, 1400, noise="renorm")
    a, b = v[0][:1333], v[1][:1100]
    counts, digests = ref.match_pairs_u8_digest([a, b], np.array([[0, 1]], np.int32), 0.8)
    count, digest, m12, m21 = ref.match_large_pair_u8(a, b, 0.8)
    assert count == counts[0] > 50 and digest == digests[0]
    t12, t21 = ref.twoway("u8", a, b, 0.8)
    f12, f21 = ref.remove_inconsistent(t12, t21)
    assert np.array_equal(m12, f12) and np.array_equal(m21, f21)
