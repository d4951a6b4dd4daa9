"""Development probe run on the GPU box: raw similarity dump vs numpy, parity vs the
oracle on a few shapes, and scan-kernel ablation timings.  Not part of the test-suite."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from orthosfm_b200 import ExhaustiveMatching, FeatureSet, Viewport, synth, KIND_SIFT_U8  # noqa: E402

O = oracle.Oracle()
out = {}


def views_to_vp(views):
    return [Viewport(FeatureSet(sift_descriptors=v)) for v in views]


def step(name, fn):
    t = time.time()
    try:
        r = fn()
        out[name] = r
        print(f"[{name}] {r}  ({time.time() - t:.2f}s)", flush=True)
    except Exception as e:  # noqa: BLE001
        out[name] = f"EXC {type(e).__name__}: {e}"
        print(f"[{name}] EXC {type(e).__name__}: {e}", flush=True)


def dump_check(n1, n2):
    rng = np.random.default_rng(n1 * 131 + n2)
    a = rng.integers(0, 256, (n1, 128), dtype=np.uint8)
    b = rng.integers(0, 256, (n2, 128), dtype=np.uint8)
    with ExhaustiveMatching() as m:
        m.init(views_to_vp([a, b]))
        s = m.debug_dump_similarity(KIND_SIFT_U8, 0, 1)
    ref = a.astype(np.int64) @ b.astype(np.int64).T
    bad = np.argwhere(s != ref)
    return {"shape": list(s.shape), "mismatch": int(bad.shape[0]),
            "first_bad": bad[:4].tolist(), "got": [int(s[i, j]) for i, j in bad[:4]],
            "want": [int(ref[i, j]) for i, j in bad[:4]]}


def parity(n1, n2, cfg=2, ratio=0.8):
    vs = synth.sift_views(cfg, 2, max(n1, n2))
    a, b = vs[0][:n1], vs[1][:n2]
    with ExhaustiveMatching() as m:
        m.init(views_to_vp([a, b]))
        tw = m.twoway_match(KIND_SIFT_U8, 0, 1)
        res = m.pairwise_match(0, 1)
        st = m.stats()
    o12, o21 = O.twoway("u8", a, b, ratio)
    f12, f21 = O.remove_inconsistent(o12, o21)
    return {"twoway_ok": bool(np.array_equal(tw.matches_1_2, o12) and np.array_equal(tw.matches_2_1, o21)),
            "filtered_ok": bool(np.array_equal(res.matches_1_2, f12) and np.array_equal(res.matches_2_1, f21)),
            "n_oneway": int((o12 >= 0).sum()), "n_consistent": int((f12 >= 0).sum()),
            "bad12": int((tw.matches_1_2 != o12).sum()), "bad21": int((tw.matches_2_1 != o21).sum()),
            "cand": st["candidate_rows"], "slow": st["slow_rows"], "selfcheck": st["self_check_failures"]}


def timing(num_views, n, modes=(0, 1, 2), noise="renorm"):
    import torch
    views = synth.sift_views(2, num_views, n, noise=noise)
    pairs = synth.all_pairs(num_views)
    r = {}
    with ExhaustiveMatching() as m:
        m.init(views_to_vp(views))
        cap = int(len(pairs) * n * 0.5) + 1024
        out_ij = torch.empty((cap, 2), dtype=torch.int32, device="cuda")
        for mode in modes:
            m.debug_set_scan_mode(mode)
            best = None
            for it in range(3):
                try:
                    loff = m.match_pairs_compact(pairs, out_ij)
                except Exception as e:  # noqa: BLE001
                    if mode == 0:
                        raise
                    loff = None
                st = m.stats()
                if best is None or st["last_scan_ms"] < best["last_scan_ms"]:
                    best = st
            cmp_ = best["last_comparisons"]
            r[f"mode{mode}"] = {"scan_ms": round(best["last_scan_ms"], 3), "total_ms": round(best["last_total_ms"], 3),
                                "Tcmp/s_scan": round(cmp_ / best["last_scan_ms"] / 1e9, 3),
                                "TOPs_alg": round(cmp_ * 256 / best["last_scan_ms"] / 1e9, 1),
                                "matches": None if loff is None else int(loff[-1]),
                                "cand": st["candidate_rows"], "slow": st["slow_rows"],
                                "sm_ghz": round(best["last_scan_sm_cycles"] / max(best["last_scan_ns"], 1), 3),
                                "Mcycles": round(best["last_scan_sm_cycles"] / 1e6, 3)}
        m.debug_set_scan_mode(0)
    return r


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "dump"):
        step("dump_128x256", lambda: dump_check(128, 256))
        step("dump_100x300", lambda: dump_check(100, 300))
        step("dump_300x1000", lambda: dump_check(300, 1000))
    if which in ("all", "parity"):
        for n1, n2 in [(128, 256), (500, 700), (1000, 1), (1, 1000), (2000, 3000), (4096, 4096)]:
            step(f"parity_{n1}x{n2}", lambda n1=n1, n2=n2: parity(n1, n2))
    if which in ("all", "timing"):
        step("timing_36x8192", lambda: timing(36, 8192, modes=(0, 1, 2)))
        step("timing_36x8192_lsb", lambda: timing(36, 8192, modes=(0,), noise="lsb"))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/probe.json", "w") as f:
        json.dump(out, f, indent=1)
