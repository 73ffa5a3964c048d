"""Counter-based synthetic clouds (BASELINE configs C4/C5).

N = side*side points on a smooth height field: jittered side x side grid of pitch h,
  x = (i + 0.5 + 0.8 (u1 - 0.5)) h,  y = (j + 0.5 + 0.8 (u2 - 0.5)) h,
  z = 1.5 + 0.10 sin(2 pi x / 0.9) cos(2 pi y / 1.3) + 0.03 sin(2 pi (x + y) / 0.21) + 0.0005 n3,
u1, u2 ~ U(0,1), n3 ~ U(-1,1) from SplitMix64(seed, point index, stream); the points are then
permuted by the argsort of a fourth SplitMix64 stream, so memory order is random like the bundled
clouds.  Everything is a pure function of (seed, side, pitch): CPU oracle and GPU see identical bits.
"""
import numpy as np

SEED = 20240601


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x.copy()
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _uniform(seed, idx, stream):
    with np.errstate(over="ignore"):
        key = _splitmix64(np.uint64(seed) * np.uint64(0x100000001B3) + np.uint64(stream))
        bits = _splitmix64(idx.astype(np.uint64) ^ key)
    return (bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0), bits


def sheet_cloud(side=1024, pitch=0.004, seed=SEED, shuffle=True):
    """Returns float32 [side*side, 3]."""
    n = side * side
    idx = np.arange(n, dtype=np.uint64)
    i = (idx % np.uint64(side)).astype(np.float64)
    j = (idx // np.uint64(side)).astype(np.float64)
    u1, _ = _uniform(seed, idx, 1)
    u2, _ = _uniform(seed, idx, 2)
    u3, _ = _uniform(seed, idx, 3)
    x = (i + 0.5 + 0.8 * (u1 - 0.5)) * pitch
    y = (j + 0.5 + 0.8 * (u2 - 0.5)) * pitch
    z = (1.5 + 0.10 * np.sin(2 * np.pi * x / 0.9) * np.cos(2 * np.pi * y / 1.3)
         + 0.03 * np.sin(2 * np.pi * (x + y) / 0.21) + 0.0005 * (2.0 * u3 - 1.0))
    pts = np.stack([x, y, z], axis=1).astype(np.float32)
    if shuffle:
        _, bits = _uniform(seed, idx, 4)
        pts = pts[np.argsort(bits, kind="stable")]
    return np.ascontiguousarray(pts)
