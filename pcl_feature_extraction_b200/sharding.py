"""Multi-GPU partitioning of the feature-extraction path: one process per GPU (torch.distributed), no
collective on the data path except the one-off halo exchange.

The reference has nothing distributed; every stage of the path is per-point with a bounded
spatial support, so it shards two ways:

* cloud-level  - independent clouds are dealt round-robin to the ranks (`assign_clouds`); no communication.
  This is what `bench.py --gpus N` measures (config C5's 64-cloud batch).
* slab-level   - ONE cloud is cut into `world` slabs along its longest axis at equal-count quantiles
  (`slab_cuts`).  A rank owns the points of its slab and receives, once, the points of its two neighbours
  that lie within `halo` of the cut (`exchange_halo`: pairwise isend / irecv, NCCL on GPUs, gloo on CPUs).
  Every stage then runs unchanged on (owned + halo) points and only the owned rows are kept: with
  halo >= the support of the stage chain, each owned point sees exactly the neighbourhood it sees on one GPU.
  Support of a chain = sum of its radii: normals r_n; SPFH needs normals of points within r_f -> r_f + r_n;
  FPFH needs SPFH rows of points within r_f -> 2 r_f + r_n; SHOT (frame + descriptor at r_s on normals) ->
  r_s + r_n.  For k-searches the radius is data dependent: `knn_support_radius` takes the largest k-th
  neighbour distance over all ranks (all-reduce MAX).

Host-side logic only (numpy + torch.distributed): the CUDA library is called by the caller on the returned
arrays, so this module is exercised on CPU with the gloo backend in tests/test_sharding_gloo.py.
"""
import numpy as np

try:  # torch is plumbing here (process groups); the pure partitioning functions work without it
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


def assign_clouds(n_clouds, world, rank):
    """cloud ids of `rank` (round-robin), e.g. 64 clouds on 8 ranks -> 8 each"""
    return list(range(rank, n_clouds, world))


def longest_axis(lo, hi):
    return int(np.argmax(np.asarray(hi, np.float64) - np.asarray(lo, np.float64)))


def slab_cuts(coord, world):
    """world-1 cut positions at equal-count quantiles of the 1-D coordinates (finite values only)"""
    c = np.sort(coord[np.isfinite(coord)].astype(np.float64))
    if world <= 1 or len(c) == 0:
        return np.zeros(0, np.float64)
    idx = (np.arange(1, world) * len(c)) // world
    return c[np.minimum(idx, len(c) - 1)]


def slab_of(coord, cuts):
    """slab index of every coordinate: slab r = [cuts[r-1], cuts[r]) (non-finite points go to slab 0)"""
    s = np.searchsorted(cuts, coord, side="right").astype(np.int32)
    s[~np.isfinite(coord)] = 0
    return s


def chain_support(normal_radius=0.0, feature_radius=0.0, descriptor="fpfh"):
    """halo width that makes a sharded run see the single-GPU neighbourhoods (see the module docstring)"""
    if descriptor == "normals":
        return normal_radius
    if descriptor == "fpfh":
        return 2.0 * feature_radius + normal_radius
    if descriptor == "shot":
        return feature_radius + normal_radius
    if descriptor == "fpfh+shot":
        return max(2.0 * feature_radius + normal_radius, feature_radius + normal_radius)
    raise ValueError(descriptor)


def halo_masks(coord, cuts, rank, halo):
    """(to_left, to_right): which of THIS rank's owned points its left / right neighbour needs"""
    world = len(cuts) + 1
    to_left = np.zeros(len(coord), bool)
    to_right = np.zeros(len(coord), bool)
    if rank > 0:
        to_left = coord < cuts[rank - 1] + halo
    if rank < world - 1:
        to_right = coord >= cuts[rank] - halo
    return to_left, to_right


def halo_mask_for(coord, cuts, peer, halo):
    """which of the given (owned) points rank `peer` needs: those within `halo` of its slab
    [cuts[peer-1], cuts[peer]).  For peer = rank +- 1 this is `halo_masks`; a farther peer only gets points when
    the slabs between are narrower than the halo (equal-count cuts through a dense region)."""
    world = len(cuts) + 1
    m = np.isfinite(coord)
    if peer > 0:
        m &= coord >= cuts[peer - 1] - halo
    if peer < world - 1:
        m &= coord < cuts[peer] + halo
    return m


def exchange_halo(owned, axis, cuts, rank, world, halo, device=None, group=None):
    """One-off halo exchange with every slab that lies within `halo` of this one.

    owned: [n, C] float32 rows of this rank (xyz first).  Returns (local, n_owned): local = owned rows followed
    by the halo rows received from the other ranks in ascending rank order.  Usually only the two neighbours
    trade points; equal-count cuts through a dense region can make a slab narrower than the halo, and then an
    owned point needs neighbours from rank r +- 2 and beyond, so every pair of ranks whose slabs come within
    `halo` of each other trades.  Works on any backend: the count matrix travels first (all-gather), then the
    payload as pairwise isend / irecv (NCCL P2P over NVLink when the tensors are on GPUs)."""
    owned = np.ascontiguousarray(owned, np.float32)
    if world == 1:
        return owned, len(owned)
    coord = owned[:, axis].astype(np.float64)
    dev = device if device is not None else torch.device("cpu")
    peers = [p for p in range(world) if p != rank]
    send = {p: owned[halo_mask_for(coord, cuts, p, halo)] for p in peers}
    width = owned.shape[1]
    # counts: row r of the matrix = what rank r sends to each rank
    mine = torch.zeros(world, dtype=torch.int64, device=dev)
    for p in peers:
        mine[p] = len(send[p])
    rows = [torch.zeros(world, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(rows, mine, group=group)
    counts = torch.stack(rows).cpu().numpy()
    # payload
    bufs_out = {p: torch.from_numpy(np.ascontiguousarray(send[p])).to(dev) for p in peers if len(send[p])}
    bufs_in = {p: torch.empty((int(counts[p, rank]), width), dtype=torch.float32, device=dev) for p in peers
               if counts[p, rank] > 0}
    ops = []
    for p in peers:
        if p in bufs_out:
            ops.append(dist.P2POp(dist.isend, bufs_out[p], p, group))
        if p in bufs_in:
            ops.append(dist.P2POp(dist.irecv, bufs_in[p], p, group))
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    parts = [owned] + [bufs_in[p].cpu().numpy() for p in peers if p in bufs_in]
    return np.concatenate(parts, 0), len(owned)


def knn_support_radius(kth_dist_local_max, chain_len, device=None, group=None):
    """support of a chain of `chain_len` k-searches: chain_len x the largest k-th neighbour distance over all
    ranks (all-reduce MAX of one float)"""
    t = torch.tensor([float(kth_dist_local_max)], dtype=torch.float64, device=device or torch.device("cpu"))
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return chain_len * float(t.item())


def cloud_resolution_over_ranks(nn_dist_owned, device=None, group=None):
    """computeCloudResolution (keypoints.h:401-428) of a slab-sharded cloud: every rank passes the distances of its
    OWNED points to their nearest other point (searched among owned + halo points); one all-reduce of (sum, count).
    Non-finite distances (points without a neighbour) are left out, as in the reference."""
    d = np.asarray(nn_dist_owned, np.float64)
    d = d[np.isfinite(d)]
    t = torch.tensor([float(d.sum()), float(len(d))], dtype=torch.float64, device=device or torch.device("cpu"))
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t[0].item() / t[1].item()) if t[1].item() > 0 else 0.0


def max_over_ranks(seconds, device=None, group=None):
    """the time every multi-GPU number is quoted with: MAX over ranks of a device-measured duration"""
    t = torch.tensor([float(seconds)], dtype=torch.float64, device=device or torch.device("cpu"))
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def gather_rows(rows, n_total, owner_index, rank, world, device=None, group=None):
    """collect the owned rows of every rank on rank 0 into an [n_total, C] array in the ORIGINAL point order.
    rows: [n_owned, C] float32; owner_index: [n_owned] original indices.  (verification / output assembly; the
    sharded compute itself never needs it)"""
    rows = np.ascontiguousarray(rows, np.float32)
    owner_index = np.ascontiguousarray(owner_index, np.int64)
    if world == 1:
        out = np.zeros((n_total, rows.shape[1]), np.float32)
        out[owner_index] = rows
        return out
    dev = device or torch.device("cpu")
    n_local = torch.tensor([len(rows)], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    m = int(max(int(s.item()) for s in sizes))
    pad_rows = torch.zeros((m, rows.shape[1]), dtype=torch.float32, device=dev)
    pad_idx = torch.full((m,), -1, dtype=torch.int64, device=dev)
    pad_rows[: len(rows)] = torch.from_numpy(rows).to(dev)
    pad_idx[: len(rows)] = torch.from_numpy(owner_index).to(dev)
    all_rows = [torch.empty_like(pad_rows) for _ in range(world)]
    all_idx = [torch.empty_like(pad_idx) for _ in range(world)]
    dist.all_gather(all_rows, pad_rows, group=group)
    dist.all_gather(all_idx, pad_idx, group=group)
    if rank != 0:
        return None
    out = np.zeros((n_total, rows.shape[1]), np.float32)
    for r, i in zip(all_rows, all_idx):
        i = i.cpu().numpy()
        keep = i >= 0
        out[i[keep]] = r.cpu().numpy()[keep]
    return out


# ------------------------------------------------------------------ descriptor matching over ranks
def pack_nn(d2, idx, offset=0):
    """(d2 float32 >= 0, idx int32 >= 0 or -1) -> int64 keys whose integer order is (d2, global index) order.
    d2 >= 0, so the float's bit pattern orders like its value; a missing match (-1) packs to the largest key."""
    d2 = np.ascontiguousarray(d2, np.float32)
    idx = np.ascontiguousarray(idx, np.int64)
    # sanitise what a user-supplied match_fn may return: -0.0 has its sign bit set (a negative key that would win
    # every MIN), a NaN distance is not a match
    none = (idx < 0) | np.isnan(d2)
    d2 = np.abs(np.where(none, np.float32(0), d2)).astype(np.float32)
    key = (d2.view(np.uint32).astype(np.int64) << 32) | (idx + offset)
    key[none] = np.iinfo(np.int64).max
    return key


def unpack_nn(key):
    key = np.asarray(key, np.int64)
    none = key == np.iinfo(np.int64).max
    idx = (key & 0xFFFFFFFF).astype(np.int32)
    d2 = (key >> 32).astype(np.uint32).view(np.float32).copy()
    idx[none] = -1
    d2[none] = np.inf
    return idx, d2


def sharded_match_nn(match_fn, a, b_local, b_offset, device=None, group=None):
    """Exact 1-NN of every row of `a` among target rows that are SHARDED over the ranks.

    Every rank holds all query rows `a` and its block `b_local` of the targets (global index of its first row:
    b_offset).  match_fn(a, b_local) -> (idx int32, d2 float32) is the single-GPU matcher (Context.match_nn).  The
    one collective of the path: an all-reduce MIN over packed (d2, global index) keys, which is exactly the
    "smallest distance, then lowest index" rule of the single-GPU result.  Returns (idx, d2) on every rank."""
    idx, d2 = match_fn(a, b_local) if len(b_local) else (np.full(len(a), -1, np.int32), np.full(len(a), np.inf, np.float32))
    key = pack_nn(d2, idx, b_offset)
    t = torch.from_numpy(key).to(device or torch.device("cpu"))
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return unpack_nn(t.cpu().numpy())


def ring_match_nn(match_fn, a_local, b_local, b_offset, b_total, rank, world, device=None, group=None):
    """Exact 1-NN when BOTH descriptor sets are sharded (ring rotation of target blocks).

    Rank r holds its queries `a_local` and the target block `b_local` (global index of its first row: b_offset;
    b_total rows over all ranks).  In step s it matches its queries against the block that started on rank
    (r + s) % world while that block is already on its way to rank r - 1 (isend / irecv posted BEFORE the match,
    so the transfer - NCCL P2P over NVLink on GPUs - overlaps the GEMM of the current block), and keeps the
    smaller packed (d2, global index) key.  No all-reduce: a rank ends with the final answer for its own queries.

    a_local / b_local: numpy arrays (CPU ranks, gloo) or torch tensors that already live on `device` (GPU ranks:
    blocks then travel device to device and are handed to match_fn as tensors).  match_fn(a, b) -> (idx int32
    [-1 = none], d2 float32) as numpy arrays is the single-GPU matcher.  Returns (idx, d2) of a_local."""
    as_tensor = torch is not None and isinstance(b_local, torch.Tensor)
    n_a = len(a_local)
    dim = a_local.shape[1]
    best = np.full(n_a, np.iinfo(np.int64).max, np.int64)
    dev = device or (b_local.device if as_tensor else torch.device("cpu"))
    if as_tensor:
        cur = b_local.reshape(-1, dim).contiguous()
    else:
        a_local = np.ascontiguousarray(a_local, np.float32)
        cur = torch.from_numpy(np.ascontiguousarray(b_local, np.float32).reshape(-1, dim)).to(dev)
    cur_off = int(b_offset)

    def match(block, off):
        if len(block) == 0 or n_a == 0:
            return
        idx, d2 = match_fn(a_local, block if as_tensor else block.cpu().numpy())
        np.minimum(best, pack_nn(d2, idx, off), out=best)

    if world == 1:
        match(cur, cur_off)
        return unpack_nn(best)
    left, right = (rank - 1) % world, (rank + 1) % world
    for step in range(world):
        nxt, hdr_in, works = None, None, []
        if step < world - 1:
            # header (rows, offset) and payload of the block travel to the left neighbour; ours arrives from the right
            hdr_out = torch.tensor([len(cur), cur_off], dtype=torch.int64, device=dev)
            hdr_in = torch.zeros(2, dtype=torch.int64, device=dev)
            for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, hdr_out, left, group),
                                             dist.P2POp(dist.irecv, hdr_in, right, group)]):
                w.wait()
            nxt = torch.empty((int(hdr_in[0].item()), dim), dtype=torch.float32, device=dev)
            ops = []
            if cur.numel():
                ops.append(dist.P2POp(dist.isend, cur, left, group))
            if nxt.numel():
                ops.append(dist.P2POp(dist.irecv, nxt, right, group))
            works = dist.batch_isend_irecv(ops) if ops else []
        match(cur, cur_off)
        for w in works:
            w.wait()
        if nxt is not None:
            cur, cur_off = nxt, int(hdr_in[1].item())
    idx, d2 = unpack_nn(best)
    assert idx.max(initial=-1) < b_total
    return idx, d2
