import sys, os, time
sys.path.insert(0, "/root/repo")
import numpy as np
import pcl_feature_extraction_b200 as pfx
Z = np.load("/root/repo/tests/golden/clouds.npz")
ctx = pfx.Context(0); ctx.set_viewpoint(0,0,0)
pts = Z["underwater_source"]
for rep in range(2):
    ctx.set_surface(pts)
    h = ctx.harris3d(0.01, 1e-6)
    snapped = h["snapped_idx"][h["snapped_idx"] >= 0]
    ctx.normals(radius=0.03, want_output=False)
    ctx.set_queries(pts[snapped])
    ctx.profile_begin(None)
    t0=time.perf_counter(); s,_ = ctx.shot352(0.05); dt=time.perf_counter()-t0
    prof = ctx.profile_end()
    ctx.set_queries(None)
print("shot352 wall ms", dt*1e3, "queries", len(snapped))
for nm,(c,ms) in sorted(prof.items(), key=lambda kv:-kv[1][1])[:12]: print(f"  {nm:45s} x{c:2d} {ms:8.3f} ms")
ctx.set_surface(pts)
ctx.profile_begin(None); t0=time.perf_counter(); h = ctx.harris3d(0.01, 1e-6); dt=time.perf_counter()-t0; prof=ctx.profile_end()
print("harris wall ms", dt*1e3)
for nm,(c,ms) in sorted(prof.items(), key=lambda kv:-kv[1][1])[:10]: print(f"  {nm:45s} x{c:2d} {ms:8.3f} ms")
