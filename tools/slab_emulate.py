"""Development diagnostic (one GPU): emulate one interior slab of an 8-way slab run inside one process and find the
first stage whose owned rows differ from the whole-cloud run."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pcl_feature_extraction_b200 as pfx
from pcl_feature_extraction_b200.synth import sheet_cloud

K, PITCH = 32, 0.004
R = 3.2 * PITCH
side = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
world, rank = 8, int(sys.argv[2]) if len(sys.argv) > 2 else 3
halo = float(sys.argv[3]) if len(sys.argv) > 3 else 0.1264
pts = sheet_cloud(side=side, pitch=PITCH, seed=20240601)
n = len(pts)
axis = int(np.argmax(pts.max(0) - pts.min(0)))
c = pts[:, axis]
lo_all, hi_all = c.min(), c.max()
cuts = np.quantile(c, np.arange(1, world) / world).astype(np.float32)
lo = -np.inf if rank == 0 else cuts[rank - 1]
hi = np.inf if rank == world - 1 else cuts[rank]
gid = np.arange(n)
owned = np.where((c >= lo) & (c < hi))[0]
halo_ids = np.where(((c >= lo - halo) & (c < hi + halo)) & ~((c >= lo) & (c < hi)))[0]
key = lambda ids: ids[np.lexsort((ids, ids % world))]     # (source rank, position) order of the slab run
local = np.concatenate([key(owned), key(halo_ids)])
ctx = pfx.Context(0)
ctx.set_viewpoint(0, 0, 0)

def run(cloud):
    ctx.set_surface(cloud)
    ctx.set_queries(None)
    nr = ctx.normals(k=K)
    idx, d2 = ctx.knn(K)
    f = ctx.fpfh(k=K)
    s, rf = ctx.shot352(R)
    return nr, idx, d2, f, s, rf

full = run(pts)
part = run(np.ascontiguousarray(pts[local]))
no = len(owned)
g_own = local[:no]
out = {"n_owned": int(no), "n_local": int(len(local)), "halo": halo}
names = ["normals", "knn_idx", "knn_d2", "fpfh", "shot", "lrf"]
for nm, a, b in zip(names, full, part):
    A, B = a[g_own], b[:no]
    if nm == "knn_idx":
        B = local[B]                       # local indices -> global ids
        bad = np.where((A != B).any(1))[0]
    else:
        bad = np.where(~((A.view(np.uint32) == B.view(np.uint32)) | (np.isnan(A) & np.isnan(B))).all(1))[0]
    dcut = np.minimum(np.abs(c[g_own[bad]] - lo), np.abs(c[g_own[bad]] - hi)) if len(bad) else np.zeros(0)
    out[nm] = {"rows_differ": int(len(bad)), "dist_to_cut": [float(v) for v in np.sort(dcut)[:6]],
               "ids": [int(v) for v in g_own[bad][:6]]}
# the k-th neighbour distances near the cut and the worst chain reach
dk = np.sqrt(full[2][:, -1])
near = (np.abs(c - lo) < 0.15) | (np.abs(c - hi) < 0.15)
out["dk_near_max"] = float(dk[near].max())
out["dk_all_max"] = float(dk.max())
print(json.dumps(out))
ctx.close()
