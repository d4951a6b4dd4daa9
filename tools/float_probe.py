"""The float path (osfm_match_twoway_f32) on one pair: tensor-core filter first against the exact
kernel alone, host buffers in and out, plus the device-only share seen by CUDA events.
    python tools/float_probe.py [n]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from orthosfm_b200 import ExhaustiveMatching, Matching, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
v = synth.sift_views(2, 2, n, noise="renorm")
fa = v[1].astype(np.float32) / 255.0
fb = v[0].astype(np.float32) / 255.0
fa /= np.linalg.norm(fa, axis=1, keepdims=True)
fb /= np.linalg.norm(fb, axis=1, keepdims=True)
opts = Matching.Options(128, 0.8, float(np.finfo(np.float32).max))
with ExhaustiveMatching() as m:
    res = {}
    for mode, name in ((2, "filter first"), (1, "exact kernel only")):
        m.debug_set_float_path(mode)
        st0 = m.stats()
        ts = []
        for _ in range(5):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = m.twoway_match_f32(opts, fa, fb)
            ts.append(time.perf_counter() - t0)
        st1 = m.stats()
        res[mode] = r
        print(f"{name:18s}: {1e3 * min(ts[1:]):8.3f} ms per pair ({n} x {n}), matches {(r.matches_1_2 >= 0).sum()}, "
              f"rows left to the exact kernel {(st1['float_exact_rows'] - st0['float_exact_rows']) // 5} of {2 * n}", flush=True)
    print("same vectors:", bool(np.array_equal(res[1].matches_1_2, res[2].matches_1_2) and np.array_equal(res[1].matches_2_1, res[2].matches_2_1)))
