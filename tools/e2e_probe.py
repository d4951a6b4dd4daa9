"""Where the end-to-end step of config 2 spends its time: staging call, matching call, device span."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from orthosfm_b200 import ExhaustiveMatching, FeatureSet, PackedViews, Viewport, synth
nv, n = 36, 8192
views = synth.sift_views(2, nv, n, noise="renorm")
host = [torch.from_numpy(v).pin_memory().numpy() for v in views]
vps = [Viewport(FeatureSet(sift_descriptors=v)) for v in host]
packed = PackedViews(vps)
pairs = synth.all_pairs(nv)
lists = torch.empty((len(pairs) * n // 4 + 4096, 2), dtype=torch.int32).pin_memory().numpy()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
me = ExhaustiveMatching()
modes = sys.argv[1].split(":") if len(sys.argv) > 1 else ["plain", "overlap", "overlap+packed"]
for mode in modes:
    rows = []
    for it in range(8):
        flush.fill_(1); torch.cuda.synchronize()
        t0 = time.perf_counter()
        if mode == "plain": me.init(vps)
        elif mode == "overlap": me.init(vps, overlap_copies=True)
        else: me.init(packed, overlap_copies=True)
        t1 = time.perf_counter()
        me.match_pairs_lists(pairs, lists)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        st = me.stats()
        rows.append((1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t2 - t0), st["last_total_ms"], st["last_scan_ms"]))
    r = np.median(np.array(rows[2:]), axis=0)
    print(f"{mode:15s} init {r[0]:.3f} ms  match {r[1]:.3f} ms  total {r[2]:.3f} ms  device span of match {r[3]:.3f} ms  filter {r[4]:.3f} ms", flush=True)
me.close()
