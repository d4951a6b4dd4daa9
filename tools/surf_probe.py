"""SURF (signed 64-d) throughput probe: 36 views x 8192, all pairs, dense results on the device."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from orthosfm_b200 import ExhaustiveMatching, FeatureSet, Viewport, synth
nv, n = 24, 8192
pool = synth.surf_pool(3, n // 2)
views = [synth.surf_view(3, v, n, pool) for v in range(nv)]
pairs = synth.all_pairs(nv)
with ExhaustiveMatching() as m:
    m.init([Viewport(FeatureSet(surf_descriptors=v)) for v in views])
    for it in range(3):
        t = time.perf_counter()
        res, counts = m.match_pairs(pairs)
        dt = time.perf_counter() - t
        st = m.stats()
        print(f"run {it}: wall {dt*1e3:.1f} ms, device {st['last_total_ms']:.2f} ms, scan {st['last_scan_ms']:.2f} ms, "
              f"{len(pairs)*n*n/st['last_total_ms']/1e9:.2f} Tcmp/s, consistent {int(counts.sum())}, "
              f"cand {st['candidate_rows']} slow {st['slow_rows']}")
