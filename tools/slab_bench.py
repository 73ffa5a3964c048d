"""Slab-sharded dense FPFH33 + SHOT352 of ONE cloud over N GPUs through the C ABI's group API (SURVEY.md §8e
partitioning 2): pfx_group_join (NCCL inside the library), pfx_slab_distribute (device-resident all-to-all of slab +
halo points, grouped ncclSend / ncclRecv), the unchanged dense stages on owned + halo points, owned rows kept.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/slab_bench.py [--side 1024] [--steps 5] [--verify]

torch.distributed is only the launcher-side plumbing here (broadcast of the 128-byte group id, the barrier and the
gather of rows for --verify); every byte of the data path moves inside libpfx_b200.so.  Prints one JSON line (rank 0):
whole-cloud descriptors/s with the MAX over ranks of the device-timed step, the halo share, and with --verify the
comparison of the gathered rows against a single-GPU run of the whole cloud on rank 0.  `slab_record` is also what
bench.py reports as its "slab" sub-record at N > 1."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

K_NN, PITCH = 32, 0.004
SHOT_RADIUS = 3.2 * PITCH


def dense_step(ctx, d_f, d_s):
    ctx.normals_dev(0.0, K_NN, None)
    ctx.fpfh_dev(0.0, K_NN, d_f.data_ptr())
    ctx.shot352_dev(SHOT_RADIUS, d_s.data_ptr())


def join_group(ctx, dist, rank, world):
    """rank 0 makes the NCCL id, the launcher's process group carries it to the others"""
    box = [ctx.group_unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    ctx.group_join(rank, world, box[0])


def slab_record(pfx, ctx, torch, dist, dev, rank, world, side=1024, steps=5, warmup=2, verify=False, seed=20240601):
    """one cloud, every rank starting from an arbitrary disjoint part of it (points rank, rank + world, ... of the
    shuffled cloud) resident in ITS device memory.  A step = pfx_slab_distribute (device to device) + the dense
    stages on owned + halo points; the rows pfx_slab_owned_rows lists are this rank's results."""
    from pcl_feature_extraction_b200.synth import sheet_cloud
    pts = sheet_cloud(side=side, pitch=PITCH, seed=seed)
    n_total = len(pts)
    ids = np.arange(rank, n_total, world, dtype=np.int32)
    part = np.zeros((len(ids), 4), np.float32)
    part[:, :3] = pts[ids]
    d_part = torch.from_numpy(part).to(dev)
    d_ids = torch.from_numpy(ids).to(dev)
    ctx.set_viewpoint(0.0, 0.0, 0.0)

    # support of the k-search chain (normals -> SPFH -> FPFH: three k-th neighbour distances) and of SHOT (radius + one
    # k-th distance for its normals).  Measured once on the owned points (halo 0): the largest k-th neighbour distance
    # over all ranks bounds everything; only points within three of those of a cut can start a chain that crosses it,
    # so the halo is three times the largest k-th distance AMONG THEM (the global maximum sits at the corners of the
    # sheet, where a neighbourhood is a quarter disc)
    n_owned0, _ = ctx.slab_distribute((d_part.data_ptr(), len(ids)), 0.0, global_ids=d_ids.data_ptr(), mem=pfx.capi.DEVICE)
    info = ctx.slab_info()
    _, d2 = ctx.knn(K_NN)
    dkk = np.sqrt(d2[:n_owned0, -1].astype(np.float64)) if n_owned0 else np.zeros(0)
    dk_all = ctx.group_allreduce([float(dkk.max()) if len(dkk) else 0.0], "max")[0]
    coord = pts[ctx.slab_global_ids()[:n_owned0], info["axis"]].astype(np.float64)
    near = np.zeros(len(coord), bool)
    if np.isfinite(info["lo"]):
        near |= coord < info["lo"] + 3.0 * dk_all
    if np.isfinite(info["hi"]):
        near |= coord >= info["hi"] - 3.0 * dk_all
    dk = ctx.group_allreduce([float(dkk[near].max()) if near.any() else 0.0], "max")[0] * 1.05
    halo = max(3.0 * dk, SHOT_RADIUS + dk)
    # second look: on the halo-0 surface a point next to a cut misses the neighbours beyond it, so its k-th distance is
    # overestimated (a half disc).  With the provisional halo in place the k-th distances of the owned points near the
    # cuts are the true ones: measure them again and shrink the halo to three of those
    n_own1, n_loc1 = ctx.slab_distribute((d_part.data_ptr(), len(ids)), halo, global_ids=d_ids.data_ptr(), mem=pfx.capi.DEVICE)
    own1 = ctx.slab_owned_rows()
    _, d2 = ctx.knn(K_NN)
    dkk = np.sqrt(d2[own1, -1].astype(np.float64)) if n_own1 else np.zeros(0)
    coord = pts[ctx.slab_global_ids()[own1], info["axis"]].astype(np.float64)
    near = np.zeros(len(coord), bool)
    if np.isfinite(info["lo"]):
        near |= coord < info["lo"] + halo
    if np.isfinite(info["hi"]):
        near |= coord >= info["hi"] - halo
    dk_true = ctx.group_allreduce([float(dkk[near].max()) if near.any() else 0.0], "max")[0] * 1.05
    halo = min(halo, max(3.0 * dk_true, SHOT_RADIUS + dk_true))

    def distribute():
        return ctx.slab_distribute((d_part.data_ptr(), len(ids)), halo, global_ids=d_ids.data_ptr(), mem=pfx.capi.DEVICE)

    n_owned, n_local = distribute()
    d_f = torch.empty((n_local, 33), dtype=torch.float32, device=dev)
    d_s = torch.empty((n_local, 361), dtype=torch.float32, device=dev)
    # ---- (1) the distribution alone (ingest: the analogue of the H2D upload of a single-GPU cloud)
    for _ in range(warmup):
        distribute()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        distribute()
    e1.record()
    torch.cuda.synchronize()
    ms_dist = ctx.group_allreduce([e0.elapsed_time(e1) / steps], "max")[0]
    # ---- (2) the dense stages on the resident slab (owned + halo points): what `value` times on one GPU.  The surface
    # is re-declared every step (device pointer of the slab rows) so that every step rebuilds its indices
    surf_rows = torch.empty((n_local, 4), dtype=torch.float32, device=dev)
    ctx._chk(ctx.lib.pfx_get_surface(ctx.h, pfx.capi._ptr(surf_rows), 16, pfx.capi.DEVICE))
    gids_local = ctx.slab_global_ids()
    own_rows = ctx.slab_owned_rows()      # local rows (ascending global id) of this rank's own points

    def dense():
        ctx.set_surface_dev(surf_rows.data_ptr(), n_local, 16)
        dense_step(ctx, d_f, d_s)

    for _ in range(warmup):
        dense()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0.record()
    for _ in range(steps):
        dense()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = ctx.group_allreduce([e0.elapsed_time(e1) / steps], "max")[0]
    halo_share = ctx.group_allreduce([(n_local - n_owned) / max(n_owned, 1)], "max")[0]
    out = {"workload": f"one {side}x{side}-point cloud slab-sharded over {world} GPUs: pfx_slab_distribute (NCCL send/recv of "
                       "slab + halo points, device to device), then dense normals k=32 + FPFH33 k=32 + SHOT352 r=12.8 mm on the "
                       "resident owned + halo points of every rank",
           "n_gpus": world, "points": n_total, "value": 2.0 * n_total / (ms * 1e-3), "unit": "descriptors/s", "ms_per_cloud": ms,
           "distribute_ms": ms_dist, "ms_per_cloud_with_distribution": ms + ms_dist, "halo_m": halo,
           "halo_points_share_max": halo_share, "scaling": "strong",
           "timed": "CUDA events, max over ranks; ms_per_cloud = dense stages on the resident slab, distribute_ms = the "
                    "device-to-device all-to-all that builds it (several small host round trips: counts, cuts)"}
    if verify:
        # the cuts of all ranks (for the diagnosis of differing rows): rank r contributes its upper bound at slot r
        his = [0.0] * world
        his[rank] = info["hi"] if np.isfinite(info["hi"]) else 0.0
        cut_pos = np.array(ctx.group_allreduce(his, "sum")[: world - 1])
        own_t = torch.from_numpy(own_rows.astype(np.int64)).to(dev)
        gid = torch.from_numpy(gids_local[own_rows].astype(np.int64)).to(dev)
        rows = torch.cat([d_f[own_t], d_s[own_t]], 1)
        if world > 1:
            sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
            dist.all_gather(sizes, torch.tensor([n_owned], dtype=torch.int64, device=dev))
            m = int(max(int(s.item()) for s in sizes))
            pad_rows = torch.zeros((m, rows.shape[1]), dtype=torch.float32, device=dev)
            pad_gid = torch.full((m,), -1, dtype=torch.int64, device=dev)
            pad_rows[:n_owned] = rows
            pad_gid[:n_owned] = gid
            all_rows = [torch.empty_like(pad_rows) for _ in range(world)] if rank == 0 else None
            all_gid = [torch.empty_like(pad_gid) for _ in range(world)] if rank == 0 else None
            dist.gather(pad_rows, all_rows, dst=0)
            dist.gather(pad_gid, all_gid, dst=0)
        else:
            all_rows, all_gid = [rows], [gid]
        if rank == 0:
            full = torch.zeros((n_total, 394), dtype=torch.float32, device=dev)
            seen = torch.zeros(n_total, dtype=torch.int32, device=dev)
            for r, g in zip(all_rows, all_gid):
                keep = g >= 0
                full[g[keep]] = r[keep]
                seen[g[keep]] += 1
            p4 = np.zeros((n_total, 4), np.float32)
            p4[:, :3] = pts
            dp = torch.from_numpy(p4).to(dev)
            f1 = torch.empty((n_total, 33), dtype=torch.float32, device=dev)
            s1 = torch.empty((n_total, 361), dtype=torch.float32, device=dev)
            ctx.set_surface_dev(dp.data_ptr(), n_total, 16)
            dense_step(ctx, f1, s1)
            torch.cuda.synchronize()
            ref = torch.cat([f1, s1], 1)
            same = (full.view(torch.int32) == ref.view(torch.int32)) | (torch.isnan(full) & torch.isnan(ref))
            diff = torch.nan_to_num((full - ref).abs(), nan=0.0)
            out["verify"] = {"every_point_owned_once": bool((seen == 1).all().item()),
                             "rows_bit_identical": float(same.all(1).float().mean().item()),
                             "fpfh_rows_bit_identical": float(same[:, :33].all(1).float().mean().item()),
                             "shot_rows_bit_identical": float(same[:, 33:].all(1).float().mean().item()),
                             "fpfh_max_abs_diff": float(diff[:, :33].max().item()),
                             "shot_max_abs_diff": float(diff[:, 33:385].max().item())}
            bad = torch.nonzero(~same.all(1)).flatten().cpu().numpy()
            if len(bad) and len(cut_pos):
                c = pts[bad, info["axis"]].astype(np.float64)
                dcut = np.abs(c[:, None] - cut_pos[None, :]).min(1)
                out["verify"]["differing_rows"] = {"count": int(len(bad)), "dist_to_cut_min": float(dcut.min()),
                                                   "dist_to_cut_median": float(np.median(dcut)), "dist_to_cut_max": float(dcut.max()),
                                                   "beyond_halo": int((dcut > halo).sum()), "first_ids": [int(v) for v in bad[:8]],
                                                   "first_dist": [float(v) for v in dcut[:8]]}
    return out


def main():
    import torch
    import torch.distributed as dist
    import pcl_feature_extraction_b200 as pfx
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--verify", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    ctx = pfx.Context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    join_group(ctx, dist, rank, world)
    out = slab_record(pfx, ctx, torch, dist, dev, rank, world, side=args.side, steps=args.steps, warmup=args.warmup, verify=args.verify)
    if rank == 0:
        print(json.dumps(out), flush=True)
    ctx.group_leave()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
