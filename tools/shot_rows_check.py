"""Dense SHOT352 from the resident k-search rows (default) against the stencil walk on a radius grid
(PFX_SHOT_ROWS=0): rows must agree bit for bit; timings of both (development check, not the contract bench)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pcl_feature_extraction_b200 as pfx
from pcl_feature_extraction_b200.synth import sheet_cloud

side = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
radius = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0128
pts = sheet_cloud(side=side)
n = len(pts)
p4 = np.zeros((n, 4), np.float32); p4[:, :3] = pts
dev = torch.device("cuda:0")
d_pts = torch.from_numpy(p4).to(dev)
outs = {}
for mode in ("0", "1"):
    os.environ["PFX_SHOT_ROWS"] = mode
    ctx = pfx.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    d_s = torch.full((n, 361), -7.0, dtype=torch.float32, device=dev)
    for it in range(3):
        ctx.set_surface_dev(d_pts.data_ptr(), n, 16)
        ctx.normals_dev(0.0, 32, None)
        l0 = ctx.launches
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.shot352_dev(radius, d_s.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        print(f"rows={mode} pass {it}: shot352 {e0.elapsed_time(e1):.3f} ms, {ctx.launches - l0} launches")
    outs[mode] = d_s.cpu().numpy().copy()
    ctx.close()
a, b = outs["0"], outs["1"]
same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
rows_same = same.all(axis=1)
diff = np.nan_to_num(np.abs(a - b), nan=0.0)
print(f"n={n} r={radius}: rows bit-identical {rows_same.mean():.6f}, max abs diff {diff.max():.3e}, "
      f"NaN rows {np.isnan(a[:, 0]).sum()} / {np.isnan(b[:, 0]).sum()}, untouched cells {(b == -7.0).sum()}")
