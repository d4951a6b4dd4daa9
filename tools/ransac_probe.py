"""Times osfm_ransac_draw_samples / osfm_ransac_fundamental on a config-2 sized job (630 pairs,
~1000 matches each, 1000 iterations) and the reference's RansacFundamental on a few pairs."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import oracle
from orthosfm_b200 import synth, ransac_draw_samples, Matching
from orthosfm_b200.matcher import ExhaustiveMatching
sys.path.insert(0, "tests")

npairs, n, iters = int(sys.argv[1]) if len(sys.argv) > 1 else 630, 1024, 1000
feats, pos, pairs, lists = [], [], [], []
for p in range(npairs):
    xy = synth.two_view_scene(p, n, 0.3)
    feats += [n, n]; pos += [xy[:, :2], xy[:, 2:]]; pairs.append((2 * p, 2 * p + 1))
    lists.append(np.stack([np.arange(n), np.arange(n)], 1))
feats = np.array(feats, np.int32); pos = np.concatenate(pos); pairs = np.array(pairs, np.int32)
off = (np.arange(npairs + 1) * n).astype(np.int64); ij = np.concatenate(lists).astype(np.int32)
from test_gpu_parity import matcher
with matcher(synth.sift_views(1, 2, 64)) as m:
    for rep in range(3):
        oracle.srand(1)
        t0 = time.perf_counter(); smp = ransac_draw_samples(off, iters); t1 = time.perf_counter()
        ooff, oij, F = m.ransac_fundamental(feats, pos, pairs, off, ij, samples=smp, max_iterations=iters)
        t2 = time.perf_counter()
        print(f"draw {1e3*(t1-t0):.1f} ms, device call {1e3*(t2-t1):.1f} ms, inliers {ooff[-1]}", flush=True)
        oracle.srand(1)
        t0 = time.perf_counter()
        ooff2, oij2, F2 = m.ransac_fundamental(feats, pos, pairs, off, ij, max_iterations=iters)
        t1 = time.perf_counter()
        assert np.array_equal(ooff, ooff2) and np.array_equal(oij, oij2) and np.array_equal(F, F2)
        print(f"draws inside, overlapped: {1e3*(t1-t0):.1f} ms", flush=True)
if oracle.have_ref():
    ref = oracle.Reference()
    oracle.srand(1)
    t0 = time.perf_counter()
    k = 4
    tot = 0
    for p in range(k):
        inl, _ = ref.ransac(np.concatenate([pos[2*p*n:(2*p+1)*n], pos[(2*p+1)*n:(2*p+2)*n]], 1), iters, 0.0015)
        tot += len(inl)
        assert np.array_equal(oij[ooff[p]:ooff[p+1]], ij[off[p]:off[p+1]][inl])
    dt = time.perf_counter() - t0
    print(f"reference: {1e3*dt/k:.1f} ms per pair -> {dt/k*npairs:.2f} s for {npairs} pairs (1 core); first {k} pairs equal")
