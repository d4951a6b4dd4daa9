"""BASELINE configs 3 (a slice), 4 (a slice) and 5 on one GPU: do the larger shapes run, how
fast, and do size-independent properties hold?  Prints one JSON object per config."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import oracle
from orthosfm_b200 import ExhaustiveMatching, synth

dev = torch.device("cuda", 0)
O = oracle.Oracle()


def run(cfg, num_views, n, npairs_limit, name):
    pool = synth.torch_sift_views(cfg, num_views, n, dev, noise="renorm")
    pool = torch.cat([pool, torch.zeros((256, 128), dtype=torch.uint8, device=dev)])
    sizes = np.full(num_views, n, np.int32)
    offsets = np.arange(num_views, dtype=np.int64) * n
    pairs = synth.all_pairs(num_views)[:npairs_limit]
    m = ExhaustiveMatching(device=0)
    m.init_device_pool(pool, offsets, sizes)
    out = torch.empty((int(len(pairs) * n * 0.3) + 4096, 2), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    t = time.perf_counter()
    m.match_pairs_compact(pairs, out)     # first call: sizes the scratch buffers (cudaMalloc)
    torch.cuda.synchronize()
    dt_first = time.perf_counter() - t
    t = time.perf_counter()
    loff = m.match_pairs_compact(pairs, out)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    st = m.stats()
    res = {"config": name, "views": num_views, "n": n, "pairs": len(pairs), "seconds": round(dt, 4), "first_call_seconds": round(dt_first, 4), "device_ms": round(m.stats()["last_total_ms"], 2),
           "Tcmp_per_s": round(len(pairs) * n * n / dt / 1e12, 3), "matches": int(loff[-1]),
           "stats": {k: st[k] for k in ("candidate_rows", "slow_rows", "exact_rows", "self_check_failures")}}
    # spot check a few rows of the first and the last pair against the oracle
    lists = out[: int(loff[-1])].cpu().numpy()
    ok = True
    for p in (0, len(pairs) - 1):
        v1, v2 = pairs[p]
        a = pool[v1 * n:(v1 + 1) * n].cpu().numpy()
        b = pool[v2 * n:(v2 + 1) * n].cpu().numpy()
        lst = lists[loff[p]:loff[p + 1]]
        got = dict(zip(lst[:, 0].tolist(), lst[:, 1].tolist()))
        rng = np.random.default_rng(p)
        for r in rng.integers(0, n, 12):
            o = int(O.twoway("u8", a[r:r + 1], b, 0.8)[0][0])
            back = int(O.twoway("u8", b[o:o + 1], a, 0.8)[0][0]) if o >= 0 else -1
            want = o if (o >= 0 and back == r) else -1
            ok = ok and got.get(int(r), -1) == want
        ok = ok and bool(np.all(np.diff(lst[:, 0]) > 0))
    res["spot_check_vs_oracle"] = ok
    print(json.dumps(res), flush=True)
    m.close()
    del pool, out
    torch.cuda.empty_cache()


which = sys.argv[1:] or ["5", "3", "4"]
if "5" in which:
    run(5, 2, 200000, 1, "config 5: one pair 200k x 200k")
if "3" in which:
    run(3, 200, 16384, 2000, "config 3 slice: 200 x 16384, first 2000 of 19900 pairs")
if "4" in which:
    run(4, 64, 32768, 400, "config 4 slice: 64 of the 1000 views x 32768, first 400 pairs")
