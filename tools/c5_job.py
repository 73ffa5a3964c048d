"""BASELINE.json config 5 as ONE job: 64 synthetic 2^20-point clouds, dense normals (k = 32) + FPFH33 (k = 32) +
SHOT352 (r = 12.8 mm) for every point, then reciprocal-free 1-NN matching of the cloud pairs (2c, 2c + 1) on every
64th descriptor (16 384 x 16 384 per pair, D = 33 and D = 352), cloud pairs dealt over the GPUs (SURVEY.md §8d C5,
§8e partitioning 1: no data-path collective).

  python tools/c5_job.py [n_clouds]                                  # one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/c5_job.py [n_clouds]

Inputs are generated on the host beforehand and are resident in HBM when the clock starts; descriptors stay on the
device (the matcher reads them in place through a 64-row stride).  Rank 0 prints one JSON line: device time of the
whole job, max over ranks."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import pcl_feature_extraction_b200 as pfx
from pcl_feature_extraction_b200.synth import sheet_cloud

SIDE, K, R_SHOT, STEP = 1024, 32, 0.0128, 64


def main():
    n_clouds = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    n = SIDE * SIDE
    pairs = list(range(rank, n_clouds // 2, world))            # pair p = clouds (2p, 2p + 1), dealt round-robin
    clouds = {}
    for p in pairs:
        for c in (2 * p, 2 * p + 1):
            p4 = np.zeros((n, 4), np.float32)
            p4[:, :3] = sheet_cloud(side=SIDE, pitch=0.004, seed=20240601 + c)
            clouds[c] = torch.from_numpy(p4).to(dev)
    # two cloud pairs in flight per GPU: pair p runs on context / stream p % 2 (index builds and kernel tails of one
    # pair fill the gaps of the other; contexts own all of their state)
    NCTX = 2
    streams = [torch.cuda.current_stream()] + [torch.cuda.Stream(device=dev) for _ in range(NCTX - 1)]
    lanes = []
    m = n // STEP
    for st in streams:
        c = pfx.Context(local)
        c.set_stream(st.cuda_stream)
        lanes.append(dict(ctx=c, stream=st,
                          f=[torch.empty((n, 33), dtype=torch.float32, device=dev) for _ in range(2)],
                          s=[torch.empty((n, 361), dtype=torch.float32, device=dev) for _ in range(2)],
                          nn_i=torch.empty(m, dtype=torch.int32, device=dev), nn_d=torch.empty(m, dtype=torch.float32, device=dev),
                          checksum=torch.zeros(2, dtype=torch.int64, device=dev)))
    ctx = lanes[0]["ctx"]

    def describe(L, c, slot):
        k = L["ctx"]
        k.set_surface_dev(clouds[c].data_ptr(), n, 16)
        k.normals_dev(0.0, K, None)
        k.fpfh_dev(0.0, K, L["f"][slot].data_ptr())
        k.shot352_dev(R_SHOT, L["s"][slot].data_ptr())

    def job():
        for L in lanes:
            with torch.cuda.stream(L["stream"]):
                L["checksum"].zero_()
        # the matcher synchronises its stream (it reads back the count of rows to redo), so the descriptions of a
        # group of NCTX pairs are enqueued first, on their own streams, and the matches follow
        for g0 in range(0, len(pairs), NCTX):
            group = list(enumerate(pairs[g0:g0 + NCTX]))
            for t, p in group:
                L = lanes[t]
                with torch.cuda.stream(L["stream"]):
                    describe(L, 2 * p, 0)
                    describe(L, 2 * p + 1, 1)
            for t, p in group:
                L = lanes[t]
                with torch.cuda.stream(L["stream"]):
                    # every 64th descriptor of cloud 2p against every 64th of cloud 2p + 1, read in place (row stride 64 rows)
                    L["ctx"].match_nn_dev(L["f"][0].data_ptr(), m, L["f"][1].data_ptr(), m, 33, L["nn_i"].data_ptr(),
                                          L["nn_d"].data_ptr(), stride_a=STEP * 132, stride_b=STEP * 132)
                    L["checksum"][0] += L["nn_i"].to(torch.int64).sum()
                    L["ctx"].match_nn_dev(L["s"][0].data_ptr(), m, L["s"][1].data_ptr(), m, 352, L["nn_i"].data_ptr(),
                                          L["nn_d"].data_ptr(), stride_a=STEP * 1444, stride_b=STEP * 1444)
                    L["checksum"][1] += L["nn_i"].to(torch.int64).sum()
        for L in lanes[1:]:  # join the side streams
            ev = torch.cuda.Event()
            ev.record(L["stream"])
            torch.cuda.current_stream().wait_event(ev)

    job()  # warm-up: buffers, grids, operand tiles
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for L in lanes[1:]:
        L["stream"].wait_event(e0)
    job()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    cs = sum(L["checksum"] for L in lanes)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(cs, op=dist.ReduceOp.SUM)
    if rank == 0:
        sec = float(ms.item()) * 1e-3
        info = ctx.match_info()
        print(json.dumps({
            "tool": "c5_job", "n_gpus": world, "clouds": 2 * (n_clouds // 2), "points_per_cloud": n,
            "job_ms": sec * 1e3, "clouds_per_s": 2 * (n_clouds // 2) / sec,
            "descriptors_per_s": 2.0 * n * 2 * (n_clouds // 2) / sec,
            "matches": {"pairs": n_clouds // 2, "rows_per_side": m, "dims": [33, 352]},
            "index_checksums": [int(cs[0].item()), int(cs[1].item())],
            "rank0_matcher": info, "pairs_in_flight_per_gpu": NCTX, "scaling": "strong (fixed 64-cloud job)",
            "note": "device time of the whole job, max over ranks; inputs resident, descriptors stay on the device"}), flush=True)
    for L in lanes:
        L["ctx"].close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
