"""Config 2 (36 x 8192, 630 pairs), device-resident: a few batched calls for an ncu launch list.
    python tools/step_probe.py [views] [noise] [steps] [both]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from orthosfm_b200 import ExhaustiveMatching, FeatureSet, Viewport, synth
nv = int(sys.argv[1]) if len(sys.argv) > 1 else 36
noise = sys.argv[2] if len(sys.argv) > 2 else "renorm"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
both = len(sys.argv) > 4 and sys.argv[4] == "both"
views = synth.sift_views(2, nv, 8192, noise=noise)
pairs = synth.all_pairs(nv)
with ExhaustiveMatching() as m:
    m.init([Viewport(FeatureSet(sift_descriptors=v)) for v in views])
    m.debug_set_both_directions(both)
    out = torch.empty((len(pairs) * 4096, 2), dtype=torch.int32, device="cuda")
    for _ in range(steps):
        loff = m.match_pairs_compact(pairs, out)
        st = m.stats()
        print(f"device {st['last_total_ms']:.3f} ms  " + "  ".join(f"{k} {v:.3f}" for k, v in st["last_phase_ms"].items()), flush=True)
    print({k: v for k, v in st.items() if "rows" in k or k in ("kernel_launches", "self_check_failures")}, int(loff[-1]))
