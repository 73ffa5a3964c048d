"""Host-side logic of the multi-GPU path (pcl_feature_extraction_b200/sharding.py) on CPU: two processes, gloo
backend.  Each rank owns one slab of a cloud, exchanges halos with its neighbour, runs the CPU ORACLE on
(owned + halo) and keeps its owned rows; gathered on rank 0 they must equal the single-process result bit for
bit - the halo makes every owned point's neighbourhood identical.  (On GPUs the same partitioning feeds the
CUDA library; that path is exercised by tools/slab_bench.py under torchrun.)"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from pcl_feature_extraction_b200 import sharding  # noqa: E402


def test_assign_clouds_and_cuts():
    assert sharding.assign_clouds(64, 8, 3) == [3, 11, 19, 27, 35, 43, 51, 59]
    assert sum(len(sharding.assign_clouds(10, 4, r)) for r in range(4)) == 10
    rng = np.random.default_rng(0)
    x = rng.normal(size=10001)
    cuts = sharding.slab_cuts(x, 4)
    s = sharding.slab_of(x, cuts)
    counts = np.bincount(s, minlength=4)
    assert len(cuts) == 3 and counts.min() >= 2499 and counts.max() <= 2502
    x[5] = np.nan
    assert sharding.slab_of(x, cuts)[5] == 0
    assert sharding.chain_support(0.03, 0.05, "fpfh") == pytest.approx(0.13)
    assert sharding.chain_support(0.03, 0.05, "shot") == pytest.approx(0.08)
    tl, tr = sharding.halo_masks(np.array([0.0, 0.5, 0.9]), np.array([1.0]), 0, 0.15)  # rank 0 owns x < 1.0
    assert not tl.any() and tr.tolist() == [False, False, True]                         # 0.9 is within the halo
    tl, tr = sharding.halo_masks(np.array([1.0, 1.1, 2.0]), np.array([1.0]), 1, 0.15)
    assert tl.tolist() == [True, True, False] and not tr.any()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, cloud_path, out_path):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import binding as orc
        orc.set_num_threads(2)
        pts = np.load(cloud_path)
        n = len(pts)
        r_n, r_f = 0.03, 0.05
        lo, hi = pts.min(0), pts.max(0)
        axis = sharding.longest_axis(lo, hi)
        cuts = sharding.slab_cuts(pts[:, axis], world)
        slab = sharding.slab_of(pts[:, axis], cuts)
        mine = np.where(slab == rank)[0]
        # rows carry (x, y, z, original index) so that results can be put back in order
        owned = np.concatenate([pts[mine], mine[:, None].astype(np.float32)], 1).astype(np.float32)
        halo = sharding.chain_support(r_n, r_f, "fpfh+shot")
        local, n_owned = sharding.exchange_halo(owned, axis, cuts, rank, world, halo)
        xyz = np.ascontiguousarray(local[:, :3])
        # every point within `halo` of an owned point must be present locally
        x = pts[:, axis].astype(np.float64)
        lo_r = cuts[rank - 1] if rank > 0 else -np.inf
        hi_r = cuts[rank] if rank < world - 1 else np.inf
        need = (slab == rank) | ((x >= lo_r - halo * 0.999) & (x < hi_r + halo * 0.999))
        have = np.zeros(n, bool)
        have[local[:, 3].astype(np.int64)] = True
        assert have[need].all()
        # the path on (owned + halo): normals everywhere, FPFH + SHOT for the owned points only
        nr, _, _ = orc.normals(xyz, radius=r_n)
        f = orc.fpfh(xyz, nr, q=xyz[:n_owned], radius=r_f)
        s, rf = orc.shot352(xyz, nr, xyz[:n_owned], r_f)
        rows = np.concatenate([nr[:n_owned], f, s, rf], 1).astype(np.float32)
        full = sharding.gather_rows(rows, n, mine, rank, world)
        t = sharding.max_over_ranks(1.0 + rank)
        assert t == float(world)
        # cloud resolution of the whole cloud from the owned points of every rank (2-NN among owned + halo points)
        _, d2nn = orc.knn(xyz, xyz[:n_owned], 2)
        res = sharding.cloud_resolution_over_ranks(np.sqrt(d2nn[:, 1].astype(np.float64)))
        assert abs(res - orc.cloud_resolution(pts)) < 1e-9
        assert sharding.knn_support_radius(0.01 * (rank + 1), 3) == pytest.approx(0.03 * world)
        if rank == 0:
            np.save(out_path, full)
    finally:
        dist.destroy_process_group()


def _match_worker(rank, world, port, path, out_path):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import binding as orc
        orc.set_num_threads(2)
        z = np.load(path)
        a, b = z["a"], z["b"]
        bounds = np.linspace(0, len(b), world + 1).astype(int)
        lo, hi = bounds[rank], bounds[rank + 1]
        idx, d2 = sharding.sharded_match_nn(orc.match_nn, a, np.ascontiguousarray(b[lo:hi]), lo)
        if rank == 0:
            np.savez(out_path, idx=idx, d2=d2)
    finally:
        dist.destroy_process_group()


def _ring_worker(rank, world, port, path, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import binding as orc
        orc.set_num_threads(2)
        z = np.load(path)
        a, b = z["a"], z["b"]
        ab = np.linspace(0, len(a), world + 1).astype(int)
        bb = np.array(z["b_bounds"])
        idx, d2 = sharding.ring_match_nn(orc.match_nn, a[ab[rank]:ab[rank + 1]], np.ascontiguousarray(b[bb[rank]:bb[rank + 1]]),
                                         bb[rank], len(b), rank, world)
        np.savez(os.path.join(out_dir, f"ring_{rank}.npz"), idx=idx, d2=d2, lo=ab[rank])
    finally:
        dist.destroy_process_group()


def test_pack_nn_orders_like_distance_then_index():
    d2 = np.array([0.5, 0.5, 0.25, 3.0, 0.0], np.float32)
    idx = np.array([7, 3, 9, -1, 2], np.int32)
    key = sharding.pack_nn(d2, idx, offset=100)
    order = np.argsort(key, kind="stable")
    assert order.tolist() == [4, 2, 1, 0, 3]  # smaller d2 first, ties by lower index, missing last
    i2, dd = sharding.unpack_nn(key)
    assert i2.tolist() == [107, 103, 109, -1, 102] and dd[3] == np.inf and np.array_equal(dd[[0, 1, 2, 4]], d2[[0, 1, 2, 4]])


def test_target_sharded_matching_equals_single_process(tmp_path, orc):
    """descriptor matching with the targets sharded over 2 ranks and a packed-MIN all-reduce (gloo)"""
    import torch.multiprocessing as mp
    rng = np.random.default_rng(5)
    a = rng.integers(0, 4, (700, 33)).astype(np.float32)   # coarse values: many exact distance ties across shards
    b = rng.integers(0, 4, (901, 33)).astype(np.float32)
    a[3, 0] = np.nan
    path, out_path = str(tmp_path / "ab.npz"), str(tmp_path / "out.npz")
    np.savez(path, a=a, b=b)
    mp.spawn(_match_worker, args=(2, _free_port(), path, out_path), nprocs=2, join=True)
    got = np.load(out_path)
    idx, d2 = orc.match_nn(a, b)
    assert np.array_equal(got["idx"], idx) and got["idx"][3] == -1
    assert np.array_equal(got["d2"][idx >= 0], d2[idx >= 0])


def test_ring_matching_with_both_sides_sharded_equals_single_process(tmp_path, orc):
    """queries AND targets sharded over 3 ranks; target blocks rotate around the ring (one of them empty)"""
    import torch.multiprocessing as mp
    rng = np.random.default_rng(9)
    a = rng.integers(0, 4, (500, 36)).astype(np.float32)
    b = rng.integers(0, 4, (777, 36)).astype(np.float32)
    a[11, 5] = np.nan
    world = 3
    path = str(tmp_path / "ab.npz")
    np.savez(path, a=a, b=b, b_bounds=np.array([0, 400, 400, 777]))   # rank 1 holds no targets
    mp.spawn(_ring_worker, args=(world, _free_port(), path, str(tmp_path)), nprocs=world, join=True)
    idx, d2 = orc.match_nn(a, b)
    for r in range(world):
        got = np.load(tmp_path / f"ring_{r}.npz")
        lo = int(got["lo"])
        sl = slice(lo, lo + len(got["idx"]))
        assert np.array_equal(got["idx"], idx[sl])
        ok = idx[sl] >= 0
        assert np.array_equal(got["d2"][ok], d2[sl][ok])
    assert idx[11] == -1


@pytest.mark.parametrize("case", ["wide_slabs", "clustered"])
def test_slab_sharded_equals_single_process(tmp_path, orc, clouds, case):
    import torch.multiprocessing as mp
    if case == "wide_slabs":
        pts = np.ascontiguousarray(clouds["underwater_source"][:6000])
        world = 2
    else:
        # equal-count cuts through a dense cluster: the two middle slabs of four are far narrower than the halo
        # (0.13 m), so owned points need neighbours from ranks r +- 2 and r +- 3 (sharding.py:82)
        rng = np.random.default_rng(5)
        base = clouds["underwater_source"]
        ax = int(np.argmax(base.max(0) - base.min(0)))
        mid = np.median(base[:, ax])
        dense = base[np.abs(base[:, ax] - mid) < 0.03][:4000]
        sparse = base[rng.choice(len(base), 2000, replace=False)]
        pts = np.unique(np.concatenate([dense, sparse]), axis=0).astype(np.float32)
        world = 4
        cuts = sharding.slab_cuts(pts[:, ax], world)
        assert np.diff(cuts).min() < 0.13
    cloud_path, out_path = str(tmp_path / "cloud.npy"), str(tmp_path / "out.npy")
    np.save(cloud_path, pts)
    mp.spawn(_worker, args=(world, _free_port(), cloud_path, out_path), nprocs=world, join=True)
    got = np.load(out_path)
    r_n, r_f = 0.03, 0.05
    nr, _, _ = orc.normals(pts, radius=r_n)
    f = orc.fpfh(pts, nr, radius=r_f)
    s, rf = orc.shot352(pts, nr, None, r_f)
    want = np.concatenate([nr, f, s, rf], 1).astype(np.float32)
    assert got.shape == want.shape
    # identical neighbourhoods, identical arithmetic: bit for bit (NaN rows included)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
