"""BASELINE config 5 (one pair 200 000 x 200 000), device-resident: per-phase device times.
    python tools/config5_probe.py [n] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from orthosfm_b200 import ExhaustiveMatching, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda", 0)
pool = synth.torch_sift_views(5, 2, n, dev, noise="renorm")
pool = torch.cat([pool, torch.zeros((256, 128), dtype=torch.uint8, device=dev)])
with ExhaustiveMatching(device=0) as m:
    m.init_device_pool(pool, np.array([0, n], np.int64), np.array([n, n], np.int32))
    out = torch.empty((n, 2), dtype=torch.int32, device=dev)
    for _ in range(steps):
        loff = m.match_pairs_compact(np.array([[1, 0]], np.int32), out)
        st = m.stats()
        print(f"device {st['last_total_ms']:.3f} ms  " + "  ".join(f"{k} {v:.3f}" for k, v in st["last_phase_ms"].items()), flush=True)
    print({k: v for k, v in st.items() if "rows" in k or "restricted" in k or k in ("kernel_launches", "self_check_failures")}, int(loff[-1]))
