"""Development diagnostic: the statistics behind the tolerances of tests/test_gpu_full_size.py (sampled rows of the
2^20-point sheet against the oracle on the full surface)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pcl_feature_extraction_b200 as pfx
from oracle import binding as orc
from pcl_feature_extraction_b200.synth import sheet_cloud

K, R = 32, 0.0128
pts = sheet_cloud(side=1024, pitch=0.004, seed=20240601)
ctx = pfx.Context(0)
ctx.set_viewpoint(0, 0, 0)
ctx.set_surface(pts)
nr = ctx.normals(k=K)
f = ctx.fpfh(k=K)
s, rf = ctx.shot352(R)
rows = np.random.default_rng(3).choice(len(pts), 600, replace=False)
q = np.ascontiguousarray(pts[rows])
of = orc.fpfh(pts, nr, q, k=K)
d = np.abs(of - f[rows]).max(1)
print("fpfh: frac<=1e-3", (d <= 1e-3).mean(), "frac<=1e-2", (d <= 1e-2).mean(), "median", np.median(d), "max", d.max(),
      "rows above 1e-3:", np.sort(d)[-8:])
os_, orf = orc.shot352(pts, nr, q, R)
ok = ~np.isnan(os_[:, 0])
_, gap2 = orc.shot_lrf(pts, q, R)
for thr in (1e-2, 1e-3, 1e-4, 0):
    clear = ok & (gap2.min(1) > thr)
    e = np.abs(s[rows][clear] - os_[clear]).max(1)
    fr = np.abs(rf[rows][clear] - orf[clear]).max(1)
    print(f"gap>{thr}: clear frac {clear.mean():.4f}; rows within 1e-4: {(e <= 1e-4).mean():.4f}; frames within 1e-5: {(fr <= 1e-5).mean():.4f}; "
          f"frames within 1e-3: {(fr <= 1e-3).mean():.4f}")
# the descriptor in the kernel's OWN frame, every row
og, _ = orc.shot352(pts, nr, q, R, lrf_in=rf[rows])
e = np.abs(s[rows][ok] - og[ok]).max(1)
print("descriptor given the kernel's own frames: max", e.max(), "frac<=1e-4", (e <= 1e-4).mean(), "frac<=1e-6", (e <= 1e-6).mean())
# how the frames differ where they differ
bad = ok & (np.abs(rf[rows] - orf).max(1) > 1e-3)
print("rows with frames differing by > 1e-3:", bad.sum(), "their gaps:", np.sort(gap2[bad].min(1))[:10], "...", np.sort(gap2[bad].min(1))[-5:] if bad.any() else "")
if bad.any():
    a, b = rf[rows][bad].reshape(-1, 3, 3), orf[bad].reshape(-1, 3, 3)
    dots = np.abs(np.einsum("nij,nij->ni", a, b))
    print("  |dot| of corresponding axes (x, y, z), first rows:\n", np.round(dots[:8], 4))
ctx.close()
