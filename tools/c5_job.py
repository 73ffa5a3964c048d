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
    ctx = pfx.Context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    d_f = [torch.empty((n, 33), dtype=torch.float32, device=dev) for _ in range(2)]
    d_s = [torch.empty((n, 361), dtype=torch.float32, device=dev) for _ in range(2)]
    m = n // STEP
    nn_i = torch.empty(m, dtype=torch.int32, device=dev)
    nn_d = torch.empty(m, dtype=torch.float32, device=dev)
    checksum = torch.zeros(2, dtype=torch.int64, device=dev)

    def describe(c, slot):
        ctx.set_surface_dev(clouds[c].data_ptr(), n, 16)
        ctx.prepare_radius(R_SHOT)
        ctx.normals_dev(0.0, K, None)
        ctx.fpfh_dev(0.0, K, d_f[slot].data_ptr())
        ctx.shot352_dev(R_SHOT, d_s[slot].data_ptr())

    def job():
        checksum.zero_()
        for p in pairs:
            describe(2 * p, 0)
            describe(2 * p + 1, 1)
            # every 64th descriptor of cloud 2p against every 64th of cloud 2p + 1, read in place (row stride 64 rows)
            ctx.match_nn_dev(d_f[0].data_ptr(), m, d_f[1].data_ptr(), m, 33, nn_i.data_ptr(), nn_d.data_ptr(),
                             stride_a=STEP * 132, stride_b=STEP * 132)
            checksum[0] += nn_i.to(torch.int64).sum()
            ctx.match_nn_dev(d_s[0].data_ptr(), m, d_s[1].data_ptr(), m, 352, nn_i.data_ptr(), nn_d.data_ptr(),
                             stride_a=STEP * 1444, stride_b=STEP * 1444)
            checksum[1] += nn_i.to(torch.int64).sum()

    job()  # warm-up: buffers, grids, operand tiles
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    job()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    cs = checksum.clone()
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
            "rank0_matcher": info, "scaling": "strong (fixed 64-cloud job)",
            "note": "device time of the whole job, max over ranks; inputs resident, descriptors stay on the device"}), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
