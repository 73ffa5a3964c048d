"""Does keeping two clouds in flight (two contexts, two streams) raise the dense-step throughput on one GPU?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pcl_feature_extraction_b200 as pfx
from pcl_feature_extraction_b200.synth import sheet_cloud

side = 1024
dev = torch.device("cuda:0")
n = side * side
clouds = []
for c in range(2):
    p4 = np.zeros((n, 4), np.float32); p4[:, :3] = sheet_cloud(side=side, seed=20240601 + c)
    clouds.append(torch.from_numpy(p4).to(dev))
NCTX = int(sys.argv[1]) if len(sys.argv) > 1 else 2
streams = [torch.cuda.Stream(device=dev) for _ in range(NCTX)]
ctxs, outs = [], []
for s in streams:
    c = pfx.Context(0)
    c.set_stream(s.cuda_stream)
    ctxs.append(c)
    outs.append((torch.empty((n, 33), dtype=torch.float32, device=dev), torch.empty((n, 361), dtype=torch.float32, device=dev)))

def step(i):
    c = ctxs[i % NCTX]; f, s = outs[i % NCTX]
    c.set_surface_dev(clouds[i & 1].data_ptr(), n, 16)
    c.normals_dev(0.0, 32, None)
    c.fpfh_dev(0.0, 32, f.data_ptr())
    c.shot352_dev(0.0128, s.data_ptr())

for i in range(6):
    step(i)
torch.cuda.synchronize()
K = 40
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(torch.cuda.current_stream())
for s in streams:
    s.wait_event(e0)
for i in range(K):
    step(i)
ends = []
for s in streams:
    e = torch.cuda.Event(); e.record(s); ends.append(e)
for e in ends:
    torch.cuda.current_stream().wait_event(e)
e1.record(torch.cuda.current_stream())
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(f"contexts in flight: {NCTX}  {ms:.3f} ms per cloud  {2*n/ms/1e3:.1f} M descriptors/s")
# same bits as a single context
ref = pfx.Context(0); ref.set_stream(torch.cuda.current_stream().cuda_stream)
f = torch.empty((n, 33), dtype=torch.float32, device=dev); s = torch.empty((n, 361), dtype=torch.float32, device=dev)
ref.set_surface_dev(clouds[(K - 1) & 1].data_ptr(), n, 16); ref.normals_dev(0.0, 32, None); ref.fpfh_dev(0.0, 32, f.data_ptr()); ref.shot352_dev(0.0128, s.data_ptr())
torch.cuda.synchronize()
fo, so = outs[(K - 1) % NCTX]
print("equal to a single context:", bool(torch.equal(f, fo)), bool(torch.equal(torch.nan_to_num(s), torch.nan_to_num(so))))
