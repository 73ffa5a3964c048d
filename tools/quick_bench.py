"""Development timing of the C4/C5 stages on one GPU (not the contract bench; see bench.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pcl_feature_extraction_b200 as pfx
from pcl_feature_extraction_b200.synth import sheet_cloud

side = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
pts = sheet_cloud(side=side)
n = len(pts)
p4 = np.zeros((n, 4), np.float32); p4[:, :3] = pts
dev = torch.device("cuda:0")
d_pts = torch.from_numpy(p4).to(dev)
ctx = pfx.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
if os.environ.get("PFX_OCC"):
    ctx.set_knn_occupancy(float(os.environ["PFX_OCC"]))
d_f = torch.empty((n, 33), dtype=torch.float32, device=dev)
d_s = torch.empty((n, 361), dtype=torch.float32, device=dev)

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

for it in range(3):
    t = [ev()]
    ctx.set_surface_dev(d_pts.data_ptr(), n, 16)
    if os.environ.get("PFX_HINT"):  # only useful with PFX_SHOT_ROWS=0 (SHOT on its own radius grid)
        ctx.prepare_radius(0.0128)
    t.append(ev())
    ctx.normals_dev(0.0, 32, None); t.append(ev())
    ctx.fpfh_dev(0.0, 32, d_f.data_ptr()); t.append(ev())
    ctx.shot352_dev(0.0128, d_s.data_ptr()); t.append(ev())
    torch.cuda.synchronize()
    names = ["set_surface", "grid+knn+normals", "spfh+fpfh", "grid+lrf+shot"]
    ms = [t[i].elapsed_time(t[i + 1]) for i in range(4)]
    print(it, " ".join(f"{nm}={m:.3f}ms" for nm, m in zip(names, ms)), f"total={sum(ms):.3f}ms launches={ctx.launches}")
ctx.profile_begin(None)
ctx.set_surface_dev(d_pts.data_ptr(), n, 16)
if os.environ.get("PFX_HINT"):
    ctx.prepare_radius(0.0128)
ctx.normals_dev(0.0, 32, None)
ctx.fpfh_dev(0.0, 32, d_f.data_ptr())
ctx.shot352_dev(0.0128, d_s.data_ptr())
prof = ctx.profile_end()
ctx.normals_dev(0.0, 32, None)
print("grid info (kNN grid):", ctx.grid_info())
for nm, (c, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"  {nm:45s} x{c:2d} {ms:8.3f} ms")
print("  sum", sum(ms for _, ms in prof.values()))
print("fpfh sample", d_f[12345 % n, :6].cpu().numpy(), "shot norm", float(torch.linalg.norm(d_s[777 % n, :352])))
