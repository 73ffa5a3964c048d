"""Ring-rotated exact matching with BOTH descriptor sets sharded over the GPUs of one box:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/ring_match_bench.py [rows_per_rank] [dim]
Every rank holds rows_per_rank queries and rows_per_rank targets (device resident); target blocks travel around
the ring over NCCL P2P while the tensor-core matcher works on the current block.  Rank 0 prints one JSON line;
results are checked against a single-GPU match of the gathered matrices when they are small enough."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import pcl_feature_extraction_b200 as pfx
from pcl_feature_extraction_b200 import sharding

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 352
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
a = torch.rand((rows, dim), generator=g, device=dev)
b = torch.rand((rows, dim), generator=g, device=dev)
ctx = pfx.Context(local)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ctx.set_match_engine(1)
idx_d = torch.empty(rows, dtype=torch.int32, device=dev)
d2_d = torch.empty(rows, dtype=torch.float32, device=dev)

def match_fn(qa, tb):
    ctx.match_nn_dev(qa.data_ptr(), len(qa), tb.data_ptr(), len(tb), dim, idx_d.data_ptr(), d2_d.data_ptr())
    torch.cuda.synchronize()
    return idx_d.cpu().numpy(), d2_d.cpu().numpy()

def run():
    return sharding.ring_match_nn(match_fn, a, b, rank * rows, world * rows, rank, world, device=dev)

run()  # warm-up (buffers, NCCL channels)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
idx, d2 = run()
torch.cuda.synchronize()
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
ok = None
if world * rows <= 131072 and world > 1:   # check against one GPU matching everything
    allb = [torch.empty_like(b) for _ in range(world)]
    dist.all_gather(allb, b)
    B = torch.cat(allb)
    i2 = torch.empty(rows, dtype=torch.int32, device=dev); dd = torch.empty(rows, dtype=torch.float32, device=dev)
    ctx.match_nn_dev(a.data_ptr(), rows, B.data_ptr(), len(B), dim, i2.data_ptr(), dd.data_ptr())
    torch.cuda.synchronize()
    flag = torch.tensor([int(np.array_equal(i2.cpu().numpy(), idx) and np.array_equal(dd.cpu().numpy(), d2))], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ok = bool(flag.item())
if rank == 0:
    secs = float(dt.item())
    flops = 2.0 * (world * rows) * (world * rows) * dim
    print(json.dumps({"tool": "ring_match_bench", "n_gpus": world, "rows_per_rank": rows, "dim": dim, "seconds": secs,
                      "algorithmic_TFLOPs": flops / secs / 1e12, "equals_single_gpu": ok,
                      "note": "wall clock, max over ranks, includes host-side key merging and one D2H of (idx, d2) per block"}), flush=True)
ctx.close()
if world > 1:
    dist.destroy_process_group()
