"""Slab-sharded dense FPFH33 + SHOT352 of ONE cloud over N GPUs (SURVEY.md §8e partitioning 2).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/slab_bench.py [--side 1024] [--steps 5] [--verify]

Each rank owns one slab (equal-count cut along the longest axis), receives its neighbours' halo points once over
NCCL P2P, runs the unchanged dense path on (owned + halo) points and keeps its owned rows.  Prints one JSON line
(rank 0): whole-cloud descriptors/s with the MAX over ranks of the device-timed step, the halo share, and with
--verify the comparison of the gathered rows against a single-GPU run of the whole cloud on rank 0."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import pcl_feature_extraction_b200 as pfx
from pcl_feature_extraction_b200 import sharding
from pcl_feature_extraction_b200.synth import sheet_cloud

K_NN, PITCH = 32, 0.004
SHOT_RADIUS = 3.2 * PITCH


def dense_step(ctx, d_pts, n, d_f, d_s):
    ctx.set_surface_dev(d_pts.data_ptr(), n, 16)
    ctx.prepare_radius(SHOT_RADIUS)
    ctx.normals_dev(0.0, K_NN, None)
    ctx.fpfh_dev(0.0, K_NN, d_f.data_ptr())
    ctx.shot352_dev(SHOT_RADIUS, d_s.data_ptr())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--verify", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pts = sheet_cloud(side=args.side, pitch=PITCH, seed=20240601)  # every rank can read the cloud; it keeps its slab
    n_total = len(pts)
    ctx = pfx.Context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_viewpoint(0.0, 0.0, 0.0)

    axis = sharding.longest_axis(pts.min(0), pts.max(0))
    cuts = sharding.slab_cuts(pts[:, axis], world)
    mine = np.where(sharding.slab_of(pts[:, axis], cuts) == rank)[0]
    # support of the k-search chain: 3 x the largest k-th neighbour distance (all-reduce MAX), and SHOT's radius chain
    ctx.set_surface(pts[mine])
    _, d2 = ctx.knn(K_NN)
    dk = float(np.sqrt(d2[:, -1].max()))
    halo = max(sharding.knn_support_radius(dk * 1.05, 3, device=dev), SHOT_RADIUS + sharding.knn_support_radius(dk * 1.05, 1, device=dev))
    owned = np.concatenate([pts[mine], mine[:, None].astype(np.float32)], 1).astype(np.float32)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    local_rows, n_owned = sharding.exchange_halo(owned, axis, cuts, rank, world, halo, device=dev)
    e1.record(); torch.cuda.synchronize()
    halo_ms = e0.elapsed_time(e1)
    n_local = len(local_rows)
    p4 = np.zeros((n_local, 4), np.float32); p4[:, :3] = local_rows[:, :3]
    d_pts = torch.from_numpy(p4).to(dev)
    d_f = torch.empty((n_local, 33), dtype=torch.float32, device=dev)
    d_s = torch.empty((n_local, 361), dtype=torch.float32, device=dev)
    for _ in range(args.warmup):
        dense_step(ctx, d_pts, n_local, d_f, d_s)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        dense_step(ctx, d_pts, n_local, d_f, d_s)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = sharding.max_over_ranks(e0.elapsed_time(e1) / args.steps, device=dev)
    out = {"metric": "FPFH+SHOT descriptors/sec, one 1M-pt cloud slab-sharded", "n_gpus": world, "points": n_total,
           "value": 2.0 * n_total / (ms * 1e-3), "unit": "descriptors/s", "ms_per_step": ms, "halo_m": halo,
           "halo_points_share": (n_local - n_owned) / max(n_owned, 1), "halo_exchange_ms": halo_ms, "scaling": "strong"}
    if args.verify:
        rows = torch.cat([d_f[:n_owned], d_s[:n_owned]], 1).cpu().numpy()
        full = sharding.gather_rows(rows, n_total, mine, rank, world, device=dev)
        if rank == 0:
            p4 = np.zeros((n_total, 4), np.float32); p4[:, :3] = pts
            dp = torch.from_numpy(p4).to(dev)
            f1 = torch.empty((n_total, 33), dtype=torch.float32, device=dev)
            s1 = torch.empty((n_total, 361), dtype=torch.float32, device=dev)
            dense_step(ctx, dp, n_total, f1, s1)
            torch.cuda.synchronize()
            ref = torch.cat([f1, s1], 1).cpu().numpy()
            same = (full.view(np.uint32) == ref.view(np.uint32)) | (np.isnan(full) & np.isnan(ref))
            diff = np.nan_to_num(np.abs(full - ref), nan=0.0)
            out["verify"] = {"rows_bit_identical": float(same.all(1).mean()), "fpfh_max_abs_diff": float(diff[:, :33].max()),
                             "shot_max_abs_diff": float(diff[:, 33:385].max())}
    if rank == 0:
        print(json.dumps(out))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
