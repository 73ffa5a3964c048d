"""Development probe of the HOST-buffer path: raw pinned-copy bandwidth of the box beside the per-call times of one
e2e step (not the contract bench; see bench.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pcl_feature_extraction_b200 as pfx
from pcl_feature_extraction_b200.synth import sheet_cloud

dev = torch.device("cuda:0")
n = 1 << 20
K = 32
R = 0.0128
# raw copies
hb = torch.empty(n * 361, dtype=torch.float32).pin_memory()
db = torch.empty(n * 361, dtype=torch.float32, device=dev)
for name, fn in (("d2h", lambda: hb.copy_(db, non_blocking=True)), ("h2d", lambda: db.copy_(hb, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"raw pinned {name}: {hb.numel() * 4 / dt / 1e9:.1f} GB/s ({dt * 1e3:.1f} ms for 1.51 GB)")

hosts = []
for c in range(2):
    p = sheet_cloud(side=1024, pitch=0.004, seed=20240601 + 1000 * c)
    p4 = torch.zeros((n, 4), dtype=torch.float32).pin_memory()
    p4[:, :3] = torch.from_numpy(p)
    hosts.append(p4)
h_fpfh = torch.empty((n, 33), dtype=torch.float32).pin_memory()
h_shot = torch.empty((n, 361), dtype=torch.float32).pin_memory()
ctx = pfx.Context(0)
ctx.set_viewpoint(0.0, 0.0, 0.0)
HOST, ASYNC = pfx.capi.HOST, pfx.capi.HOST_ASYNC
P = pfx.capi._ptr

def step(i, mode, sync_each):
    t = [time.perf_counter()]
    def mark():
        if sync_each:
            ctx._chk(ctx.lib.pfx_sync(ctx.h))
        t.append(time.perf_counter())
    ctx._chk(ctx.lib.pfx_set_surface(ctx.h, P(hosts[i & 1]), n, 16, HOST)); mark()
    ctx._chk(ctx.lib.pfx_normals(ctx.h, 0.0, K, None, 16, 3, HOST)); mark()
    ctx._chk(ctx.lib.pfx_fpfh(ctx.h, 0.0, K, P(h_fpfh), 132, mode)); mark()
    ctx._chk(ctx.lib.pfx_shot352(ctx.h, R, None, P(h_shot), 1444, mode)); mark()
    return [(t[j + 1] - t[j]) * 1e3 for j in range(4)]

if "after" in sys.argv[1:]:
    # what bench.py runs on the same context before its e2e leg
    import bench
    t0 = time.perf_counter()
    if "nomatch" not in sys.argv[1:]:
        bench.matching_record(pfx, ctx, torch, dev, {"bf16_tflops": 1668.4})
    if "nobundled" not in sys.argv[1:]:
        bench.bundled_record(pfx, ctx)
    ctx.set_queries(None)
    torch.cuda.empty_cache()
    print(f"sub-records took {time.perf_counter() - t0:.1f} s")
    for i in range(12):
        t0 = time.perf_counter()
        step(i, ASYNC, False)
        t1 = time.perf_counter()
        ctx._chk(ctx.lib.pfx_sync(ctx.h))
        print(f"  step {i}: enqueue {1e3 * (t1 - t0):7.2f} ms, with sync {1e3 * (time.perf_counter() - t0):7.2f} ms, launches {ctx.launches}")

for mode, nm in ((ASYNC, "HOST_ASYNC"), (HOST, "HOST")):
    for sync_each in (True, False):
        for i in range(2):
            step(i, mode, sync_each)
        ctx._chk(ctx.lib.pfx_sync(ctx.h))
        t0 = time.perf_counter()
        acc = np.zeros(4)
        for i in range(5):
            acc += step(i, mode, sync_each)
        ctx._chk(ctx.lib.pfx_sync(ctx.h))
        dt = (time.perf_counter() - t0) / 5
        print(f"{nm:10s} sync_each={sync_each}: {dt * 1e3:7.2f} ms/step = {2 * n / dt / 1e6:6.1f} M desc/s; host ms per call "
              f"[set_surface, normals, fpfh, shot] = {np.round(acc / 5, 2)}")
ctx.close()
