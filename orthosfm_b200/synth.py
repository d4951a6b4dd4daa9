"""Deterministic synthetic SIFT-like descriptor sets.

Mirrors what the reference feeds its matcher: unit-norm non-negative 128-d
vectors, clamped at 0.2 and re-normalised (src/mve/sfm/sift.cc:832-839), then
quantised like convert_descriptor (src/mve/sfm/exhaustive_matching.cc:18-27):
``q = floor(255 * clamp(x, 0, 1) + 0.5)`` stored as one byte.

Pure noise produces no consistent matches, so a fraction of every image's
descriptors is a perturbed copy of a shared "scene" pool; those are the rows that
survive the ratio test and the mutual filter.  Two perturbations:

``noise="renorm"``  what a second photograph does: Gaussian noise on the float
    descriptor, then SIFT's normalise / clamp / normalise and the quantiser again, so
    every row is a quantised *unit* vector (squared norm 65025 +- a few hundred) and an
    inner product of 2^16 -- where the reference's 16-bit arithmetic wraps -- is rare.
    This is what the benchmark uses.
``noise="lsb"``     +-1 LSB on the quantised bytes without re-normalising.  Not what
    SIFT produces (norms drift by +-400, every tenth planted pair reaches 2^16), which
    makes it a good stress test of the wrap handling; the parity tests default to it.
"""
from __future__ import annotations

import numpy as np

SIFT_DIM = 128
SURF_DIM = 64


def _normalise_clamp_quantise(x: np.ndarray) -> np.ndarray:
    x = x / np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-12)
    x = np.minimum(x, 0.2)
    x = x / np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-12)
    return np.floor(255.0 * np.clip(x, 0.0, 1.0) + 0.5).astype(np.uint8)


def scene_pool(cfg: int, size: int) -> np.ndarray:
    """Shared pool of quantised scene descriptors for configuration ``cfg``."""
    rng = np.random.Generator(np.random.MT19937(1000 * cfg + 999))
    return _normalise_clamp_quantise(np.abs(rng.standard_normal((size, SIFT_DIM), dtype=np.float32)))


RENORM_SIGMA = 0.004   # per-dimension noise of a planted copy, in units of the unit vector


def sift_view(cfg: int, view: int, n: int, pool: np.ndarray | None = None,
              planted_fraction: float = 0.25, noise: str = "lsb") -> np.ndarray:
    """``n x 128`` uint8 descriptors of image ``view`` (seed = 1000*cfg + view)."""
    rng = np.random.Generator(np.random.MT19937(1000 * cfg + view))
    desc = _normalise_clamp_quantise(np.abs(rng.standard_normal((n, SIFT_DIM), dtype=np.float32)))
    if pool is not None and planted_fraction > 0 and n > 0:
        k = min(int(n * planted_fraction), pool.shape[0])
        rows = rng.permutation(n)[:k]
        picks = rng.permutation(pool.shape[0])[:k]
        if noise == "lsb":
            d = rng.integers(-1, 2, size=(k, SIFT_DIM), dtype=np.int16)
            desc[rows] = np.clip(pool[picks].astype(np.int16) + d, 0, 255).astype(np.uint8)
        elif noise == "renorm":
            x = pool[picks].astype(np.float32) / 255.0
            x = np.abs(x + RENORM_SIGMA * rng.standard_normal((k, SIFT_DIM), dtype=np.float32))
            desc[rows] = _normalise_clamp_quantise(x)
        else:
            raise ValueError(f"unknown noise model {noise!r}")
    return desc


def sift_views(cfg: int, num_views: int, n: int, planted_fraction: float = 0.25,
               pool_size: int | None = None, noise: str = "lsb") -> list[np.ndarray]:
    pool = scene_pool(cfg, pool_size if pool_size is not None else max(n // 2, 1))
    return [sift_view(cfg, v, n, pool, planted_fraction, noise) for v in range(num_views)]


def surf_view(cfg: int, view: int, n: int, pool: np.ndarray | None = None,
              planted_fraction: float = 0.25) -> np.ndarray:
    """``n x 64`` int8 SURF-like descriptors: signed, unit norm scaled to 127
    (convert_descriptor, exhaustive_matching.cc:30-39)."""
    rng = np.random.Generator(np.random.MT19937(1000 * cfg + 500 + view))
    x = rng.standard_normal((n, SURF_DIM), dtype=np.float32)
    x = x / np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-12)
    v = np.clip(x, -1.0, 1.0) * 127.0
    q = np.where(v > 0, np.floor(v + 0.5), np.ceil(v - 0.5)).astype(np.int8)
    if pool is not None and planted_fraction > 0 and n > 0:
        k = min(int(n * planted_fraction), pool.shape[0])
        rows = rng.permutation(n)[:k]
        picks = rng.permutation(pool.shape[0])[:k]
        noise = rng.integers(-1, 2, size=(k, SURF_DIM), dtype=np.int16)
        q[rows] = np.clip(pool[picks].astype(np.int16) + noise, -127, 127).astype(np.int8)
    return q


def surf_pool(cfg: int, size: int) -> np.ndarray:
    return surf_view(cfg, 498, size)


def all_pairs(num_views: int) -> np.ndarray:
    """The reference's pair enumeration (src/mve/sfm/bundler_matching.cc:92-93):
    flat index i -> (view_1, view_2) with view_1 > view_2."""
    out = np.empty((num_views * (num_views - 1) // 2, 2), dtype=np.int32)
    k = 0
    for v1 in range(1, num_views):
        for v2 in range(v1):
            out[k] = (v1, v2)
            k += 1
    return out


def torch_sift_views(cfg: int, num_views: int, n: int, device, planted_fraction: float = 0.25,
                     noise: str = "lsb"):
    """Same distribution generated on the device with torch (for the large bench
    configurations where 4 GB of numpy randoms would dominate start-up).  Returns
    a ``num_views*n x 128`` uint8 tensor; seeded, but not bit-identical to the
    numpy generator."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(1000 * cfg + 7)

    def quantise(x: "torch.Tensor") -> "torch.Tensor":
        x = x / x.norm(dim=1, keepdim=True).clamp_min(1e-12)
        x = x.clamp_max(0.2)
        x = x / x.norm(dim=1, keepdim=True).clamp_min(1e-12)
        return torch.floor(255.0 * x.clamp(0.0, 1.0) + 0.5).to(torch.uint8)

    def make(rows: int) -> "torch.Tensor":
        return quantise(torch.randn((rows, SIFT_DIM), generator=g, device=device, dtype=torch.float32).abs_())

    pool = make(max(n // 2, 1))
    k = min(int(n * planted_fraction), pool.shape[0])
    out = torch.empty((num_views * n, SIFT_DIM), dtype=torch.uint8, device=device)
    chunk = max(1, (1 << 22) // max(n, 1))
    for v0 in range(0, num_views, chunk):
        v1 = min(num_views, v0 + chunk)
        out[v0 * n:v1 * n] = make((v1 - v0) * n)
    if k > 0:
        for v in range(num_views):
            rows = torch.randperm(n, generator=g, device=device)[:k] + v * n
            picks = torch.randperm(pool.shape[0], generator=g, device=device)[:k]
            if noise == "lsb":
                d = torch.randint(-1, 2, (k, SIFT_DIM), generator=g, device=device, dtype=torch.int16)
                out[rows] = (pool[picks].to(torch.int16) + d).clamp_(0, 255).to(torch.uint8)
            else:
                x = pool[picks].to(torch.float32) / 255.0
                x = (x + RENORM_SIGMA * torch.randn((k, SIFT_DIM), generator=g, device=device)).abs_()
                out[rows] = quantise(x)
    return out


def two_view_scene(seed: int, n: int, outlier_fraction: float = 0.3, noise: float = 3e-4):
    """Feature positions of n matches between two views of a random rigid scene, in MVE's
    normalised image coordinates (roughly [-0.5, 0.5]): [n, 4] float32 (x1 y1 x2 y2).  A
    fraction of the matches is wrong (the second point is random)."""
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1, 1, (n, 3)) + np.array([0, 0, 5.0])
    a = rng.uniform(-0.2, 0.2, 3)
    Rx = np.array([[1, 0, 0], [0, np.cos(a[0]), -np.sin(a[0])], [0, np.sin(a[0]), np.cos(a[0])]])
    Ry = np.array([[np.cos(a[1]), 0, np.sin(a[1])], [0, 1, 0], [-np.sin(a[1]), 0, np.cos(a[1])]])
    Rz = np.array([[np.cos(a[2]), -np.sin(a[2]), 0], [np.sin(a[2]), np.cos(a[2]), 0], [0, 0, 1]])
    R = Rz @ Ry @ Rx
    t = rng.uniform(-0.5, 0.5, 3)
    focal = 1.2
    x1 = focal * X[:, :2] / X[:, 2:3]
    Y = X @ R.T + t
    x2 = focal * Y[:, :2] / Y[:, 2:3]
    x1 += rng.normal(0, noise, x1.shape)
    x2 += rng.normal(0, noise, x2.shape)
    wrong = rng.random(n) < outlier_fraction
    x2[wrong] = rng.uniform(-0.5, 0.5, (int(wrong.sum()), 2))
    return np.concatenate([x1, x2], 1).astype(np.float32)


def sfm_scene(seed: int, num_views: int, n: int, scene_points: int, visible: float = 0.5, noise: float = 2e-4,
              surf_n: int = 0):
    """A multi-view scene for the whole two-view stage: ``scene_points`` 3-D points, each with
    a SIFT descriptor; every view sees a random ``visible`` fraction of them from its own pose
    (descriptor re-quantised with a little noise, position = projection + noise) and fills the
    rest of its ``n`` features with private descriptors at random positions.  Returns
    (list of [n, 128] uint8 descriptors, list of [n, 2] float32 positions).  With ``surf_n``
    every view also gets that many SURF features (half of them of a second set of scene
    points): returns (sift, surf [surf_n, 64] int8, positions [n + surf_n, 2], SIFT rows first
    as in FeatureSet::positions)."""
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1, 1, (scene_points, 3)) + np.array([0, 0, 5.0])
    pool = _normalise_clamp_quantise(np.abs(rng.standard_normal((scene_points, SIFT_DIM), dtype=np.float32)))
    ks = surf_n // 2
    Xs = rng.uniform(-1, 1, (ks, 3)) + np.array([0, 0, 5.0])
    spool = surf_view(seed, 10 ** 6, ks) if ks else None
    descs, poss, surfs = [], [], []
    for v in range(num_views):
        a = rng.uniform(-0.25, 0.25, 3)
        Rx = np.array([[1, 0, 0], [0, np.cos(a[0]), -np.sin(a[0])], [0, np.sin(a[0]), np.cos(a[0])]])
        Ry = np.array([[np.cos(a[1]), 0, np.sin(a[1])], [0, 1, 0], [-np.sin(a[1]), 0, np.cos(a[1])]])
        Rz = np.array([[np.cos(a[2]), -np.sin(a[2]), 0], [np.sin(a[2]), np.cos(a[2]), 0], [0, 0, 1]])
        Y = X @ (Rz @ Ry @ Rx).T + rng.uniform(-0.5, 0.5, 3)
        proj = 1.2 * Y[:, :2] / Y[:, 2:3]
        k = min(int(scene_points * visible), n)
        seen = rng.permutation(scene_points)[:k]
        desc = _normalise_clamp_quantise(np.abs(rng.standard_normal((n, SIFT_DIM), dtype=np.float32)))
        pos = rng.uniform(-0.5, 0.5, (n, 2))
        rows = rng.permutation(n)[:k]
        x = pool[seen].astype(np.float32) / 255.0
        desc[rows] = _normalise_clamp_quantise(np.abs(x + RENORM_SIGMA * rng.standard_normal(x.shape, dtype=np.float32)))
        pos[rows] = proj[seen] + rng.normal(0, noise, (k, 2))
        descs.append(desc)
        if surf_n:
            q = surf_view(seed, v, surf_n)
            spos = rng.uniform(-0.5, 0.5, (surf_n, 2))
            srows = rng.permutation(surf_n)[:ks]
            d = rng.integers(-1, 2, size=(ks, SURF_DIM), dtype=np.int16)
            q[srows] = np.clip(spool.astype(np.int16) + d, -127, 127).astype(np.int8)
            Ys = Xs @ (Rz @ Ry @ Rx).T + (Y[0] - X[0] @ (Rz @ Ry @ Rx).T)
            spos[srows] = 1.2 * Ys[:, :2] / Ys[:, 2:3] + rng.normal(0, noise, (ks, 2))
            surfs.append(q)
            pos = np.concatenate([pos, spos])
        poss.append(pos.astype(np.float32))
    if surf_n:
        return descs, surfs, poss
    return descs, poss
