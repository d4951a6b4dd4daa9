"""Track building at BASELINE config 2 scale: device (osfm_tracks_compute) against the
reference's Tracks::compute on one host core, same match lists."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from orthosfm_b200 import ExhaustiveMatching, FeatureSet, Viewport, synth
nv, n = 36, 8192
views = synth.sift_views(2, nv, n, noise="renorm")
pairs = synth.all_pairs(nv)
with ExhaustiveMatching() as m:
    m.init([Viewport(FeatureSet(sift_descriptors=v)) for v in views])
    out = np.empty((len(pairs) * 2048, 2), np.int32)
    loff = m.match_pairs_lists(pairs, out)
    ij = out[:loff[-1]]
    for it in range(3):
        t = time.perf_counter()
        got, nt, nc = m.tracks_compute([n] * nv, pairs, loff, ij)
        dt = time.perf_counter() - t
    print(f"device: {len(ij)} matches, {nv * n} features -> {nt} tracks ({nc} dropped for conflicts), {dt * 1e3:.2f} ms incl. H2D/D2H")
if oracle.have_ref():
    r = oracle.Reference()
    t = time.perf_counter()
    want, nw = r.tracks_compute([n] * nv, pairs, loff, ij)
    dt = time.perf_counter() - t
    print(f"reference (1 core): {nw} tracks, {dt * 1e3:.1f} ms; partitions equal: {np.array_equal(got, oracle.canonical_track_ids(want))}")
