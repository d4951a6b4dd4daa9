"""Small fixed workload for ncu: 12 views x 8192 (66 pairs), three batched calls."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from orthosfm_b200 import ExhaustiveMatching, FeatureSet, Viewport, synth
nv = int(sys.argv[1]) if len(sys.argv) > 1 else 12
noise = sys.argv[2] if len(sys.argv) > 2 else "renorm"
views = synth.sift_views(2, nv, 8192, noise=noise)
pairs = synth.all_pairs(nv)
with ExhaustiveMatching() as m:
    m.init([Viewport(FeatureSet(sift_descriptors=v)) for v in views])
    out = torch.empty((len(pairs) * 4096, 2), dtype=torch.int32, device="cuda")
    for _ in range(3):
        loff = m.match_pairs_compact(pairs, out)
    print(m.stats(), int(loff[-1]))
