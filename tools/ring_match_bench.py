"""Ring-rotated exact matching with BOTH descriptor sets sharded over the GPUs of one box, through
the C ABI: pfx_match_ring rotates the target blocks with ncclSend / ncclRecv on a second stream while the tensor-core
matcher works on the current block, and keeps the packed (d2, global row) minimum on the device.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/ring_match_bench.py [rows_per_rank] [dim]
Rank 0 prints one JSON line; results are checked against a single-GPU match of the gathered matrices when they are
small enough.  `ring_record` is also what bench.py reports inside its "slab" sub-record at N > 1."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def ring_record(pfx, ctx, torch, dist, dev, rank, world, rows=65536, dim=352, reps=3, check_rows=131072):
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    a = torch.rand((rows, dim), generator=g, device=dev)
    b = torch.rand((rows, dim), generator=g, device=dev)
    idx = torch.empty(rows, dtype=torch.int32, device=dev)
    d2 = torch.empty(rows, dtype=torch.float32, device=dev)
    ctx.set_match_engine(1)
    try:
        def run():
            ctx.match_ring_dev(a.data_ptr(), rows, b.data_ptr(), rows, dim, rank * rows, idx.data_ptr(), d2.data_ptr())
        run()  # warm-up (buffers, NCCL channels)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = ctx.group_allreduce([e0.elapsed_time(e1) / reps], "max")[0]
        ok = None
        if world * rows <= check_rows and world > 1:  # check against one GPU matching everything
            allb = [torch.empty_like(b) for _ in range(world)]
            dist.all_gather(allb, b)
            B = torch.cat(allb)
            i2 = torch.empty(rows, dtype=torch.int32, device=dev)
            dd = torch.empty(rows, dtype=torch.float32, device=dev)
            ctx.match_nn_dev(a.data_ptr(), rows, B.data_ptr(), len(B), dim, i2.data_ptr(), dd.data_ptr())
            torch.cuda.synchronize()
            flag = float(bool((i2 == idx).all().item()) and bool((dd == d2).all().item()))
            ok = bool(ctx.group_allreduce([flag], "min")[0] == 1.0)
    finally:
        ctx.set_match_engine(-1)
    flops = 2.0 * (world * rows) * (world * rows) * dim
    return {"workload": f"{world * rows} x {world * rows} x {dim} exact 1-NN, both sides sharded over {world} GPUs (pfx_match_ring)",
            "n_gpus": world, "rows_per_rank": rows, "dim": dim, "ms": ms, "tflops_aggregate": flops / (ms * 1e-3) / 1e12,
            "equals_single_gpu": ok, "timed": "CUDA events on the launching stream, max over ranks"}


def main():
    import torch
    import torch.distributed as dist
    import pcl_feature_extraction_b200 as pfx
    from tools.slab_bench import join_group
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    dim = int(sys.argv[2]) if len(sys.argv) > 2 else 352
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    ctx = pfx.Context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    join_group(ctx, dist, rank, world)
    out = ring_record(pfx, ctx, torch, dist, dev, rank, world, rows=rows, dim=dim)
    if rank == 0:
        print(json.dumps(out), flush=True)
    ctx.group_leave()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
