"""Development timing of descriptor matching (tensor-core engine vs exact scan) on one GPU."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pcl_feature_extraction_b200 as pfx

dev = torch.device("cuda:0")
ctx = pfx.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
cases = [(16384, 16384, 352), (16384, 16384, 33), (65536, 65536, 352)]
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in sys.argv[1].split("x"))]
for na, nb, dim in cases:
    g = torch.Generator(device=dev).manual_seed(na + dim)
    a = torch.rand((na, dim), device=dev, generator=g) * 100
    b = torch.rand((nb, dim), device=dev, generator=g) * 100
    m = min(na, nb) // 2
    b[:m] = a[:m] + torch.randn((m, dim), device=dev, generator=g)
    res = {}
    for eng, nm in ((1, "tcgen05"), (0, "exact")):
        if eng == 0 and na * nb * dim > 2e11:
            continue
        ctx.set_match_engine(eng)
        idx = torch.empty(na, dtype=torch.int32, device=dev)
        d2 = torch.empty(na, dtype=torch.float32, device=dev)
        i0 = ctx.match_info()
        for _ in range(2):
            ctx.match_nn_dev(a.data_ptr(), na, b.data_ptr(), nb, dim, idx.data_ptr(), d2.data_ptr())
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            ctx.match_nn_dev(a.data_ptr(), na, b.data_ptr(), nb, dim, idx.data_ptr(), d2.data_ptr())
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        i1 = ctx.match_info()
        res[nm] = (idx.clone(), d2.clone())
        redo = (i1["redone_exact"] - i0["redone_exact"]) / max(1, i1["rows"] - i0["rows"])
        print(f"{na}x{nb}x{dim} {nm:8s} {ms:9.3f} ms  {2.0*na*nb*dim/ms/1e9:9.1f} TFLOP/s(algorithmic)  redo={redo:.4f}")
        ctx.profile_begin(None)
        ctx.match_nn_dev(a.data_ptr(), na, b.data_ptr(), nb, dim, idx.data_ptr(), d2.data_ptr())
        for k, (c, t) in sorted(ctx.profile_end().items(), key=lambda kv: -kv[1][1])[:5]:
            print(f"      {k:40s} x{c} {t:9.3f} ms")
    if len(res) == 2:
        same = bool((res["tcgen05"][0] == res["exact"][0]).all()) and bool((res["tcgen05"][1] == res["exact"][1]).all())
        print("   engines bit-identical:", same)
ctx.set_match_engine(-1)
