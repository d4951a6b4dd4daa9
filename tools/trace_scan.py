"""Pipeline time line of the scan kernel's CTA 0 (MODE 5): where do the cycles of a tile go?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from orthosfm_b200 import ExhaustiveMatching, FeatureSet, Viewport, synth
views = synth.sift_views(2, 12, 8192, noise="renorm")
pairs = synth.all_pairs(12)
with ExhaustiveMatching() as m:
    m.init([Viewport(FeatureSet(sift_descriptors=v)) for v in views])
    m.match_pairs(pairs[:4])
    tr = m.debug_trace(pairs)
np.save("gpurun_out/trace.npy", tr)
t0 = tr[0][32, 2]
for w in (0, 1):
    mma = tr[w]
    print(f"MMA issuer warp {w}: groups 64..88: acc_empty wait start, accumulator free, issued | period | accumulator")
    for e in range(32, 48):
        print(e, mma[e, 0] - t0, mma[e, 1] - t0, mma[e, 2] - t0, "|", mma[e, 2] - mma[e - 1, 2], "|", mma[e, 3])
    d = np.diff(mma[16:250, 2])
    print("  mean issue period (cycles per group of 128x128x128):", d.mean(), "min", d.min(), "max", d.max())
for w in (4, 8, 13, 18):
    print(f"epilogue warp {w} (half {(w - 4) >> 3}, column half {((w - 4) >> 2) & 1}): wait start, acc ready, handed back, maxima done | ready->back, fold")
    for e in range(32, 44):
        print(e, *(tr[w, e] - t0), "|", tr[w, e, 2] - tr[w, e, 1], tr[w, e, 3] - tr[w, e, 2])
