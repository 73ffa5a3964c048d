"""PFH125 and PrincipalCurvatures (reference evaluation.cpp:676-715): the CPU oracle against
closed-form cases, the sequential-float-sum emulation against brute force, and the CUDA path (through the C ABI)
against the oracle.

Tolerances (GPU vs oracle, identical input normals):
 * PFH125: a row sums to 100; the CUDA pair features use a polynomial atan2 (3e-7 from libm), so a vote moves only
   when f1 sits that close to a bin edge: at most 2 moved votes per row are accepted, i.e.
   max-abs <= 2 * 100 / (n (n - 1) / 2) + 1e-4; rows with identical votes must be bit-identical (the float value of
   a bin is the sequential sum PCL builds, reproduced exactly from the vote count).
 * PrincipalCurvatures: pc1 / pc2 within 1e-6 absolute + 1e-5 relative; the principal direction within 1e-4
   (same sign - both sides build it like pcl::computeCorrespondingEigenVector) where the relative eigen gap exceeds 5 %."""
import numpy as np
import pytest


def bumpy(n, seed):
    rng = np.random.default_rng(seed)
    u = rng.uniform(0, 1, (n, 2))
    z = 0.08 * np.sin(7 * u[:, 0]) * np.cos(5 * u[:, 1]) + 0.03 * np.sin(23 * u[:, 0] + 11 * u[:, 1])
    return np.c_[u, z].astype(np.float32)


def test_seq_float_sum_equals_the_loop():
    import pcl_feature_extraction_b200 as pfx
    lib = pfx.capi.load()
    rng = np.random.default_rng(5)
    for t in range(300):
        mode = t % 3
        if mode == 0:
            n = int(rng.integers(2, 3000))
            incr = np.float32(100.0) / np.float32(n * (n - 1) // 2)
        elif mode == 1:
            incr = np.float32(rng.integers(1, 1000)) * np.float32(0.5)   # ties
        else:
            incr = np.float32(10.0 ** rng.uniform(-6, 3))
        c = int(rng.integers(0, 60000))
        s = np.float32(0)
        for _ in range(c):
            s = np.float32(s + incr)
        got = np.float32(lib.pfx_seq_float_sum(float(incr), c))
        assert got.tobytes() == s.tobytes(), (incr, c, got, s)


def test_oracle_closed_forms(orc):
    rng = np.random.default_rng(0)
    n = 20000
    # cylinder of radius 0.3: the normal map varies only around the axis -> one principal curvature is 0 and the
    # principal direction is tangential (perpendicular to the axis and to the normal)
    th, z = rng.uniform(0, 2 * np.pi, n), rng.uniform(0, 1, n)
    pts = np.c_[0.3 * np.cos(th), 0.3 * np.sin(th), z].astype(np.float32)
    nr = np.c_[np.cos(th), np.sin(th), 0 * th, 0 * th].astype(np.float32)
    out, gap = orc.principal_curvatures(pts, nr, pts[:200], radius=0.05)
    assert np.all(np.abs(out[:, 4]) < 1e-9) and np.all(out[:, 3] > 1e-4)
    assert np.all(np.abs(out[:, 2]) < 1e-6)                               # no axial component
    assert np.all(np.abs(np.sum(out[:, :3] * nr[:200, :3], 1)) < 1e-5)    # in the tangent plane
    assert np.allclose(np.linalg.norm(out[:, :3], axis=1), 1, atol=1e-6)
    # sphere: both curvatures equal up to sampling noise, PFH of a perfectly symmetric neighbourhood is one bin
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    pts = (0.5 * v).astype(np.float32)
    nr = np.c_[v, np.zeros(n)].astype(np.float32)
    out, gap = orc.principal_curvatures(pts, nr, pts[:100], radius=0.06)
    assert np.all(out[:, 3] >= out[:, 4]) and np.all(out[:, 4] > 0.3 * out[:, 3])
    h, c = orc.pfh125(pts, nr, pts[:20], radius=0.06, want_counts=True)
    m = c.sum(1)
    assert np.all(np.abs(h.sum(1) - 100) < 0.05)
    # every pair votes once: n (n - 1) / 2 votes
    cnt = orc.radius_count(pts, pts[:20], 0.06)
    assert np.array_equal(m, cnt * (cnt - 1) // 2)
    # a query without neighbours -> NaN row
    far = np.array([[5, 5, 5]], np.float32)
    assert np.isnan(orc.pfh125(pts, nr, far, radius=0.05)).all()
    assert np.isnan(orc.principal_curvatures(pts, nr, far, radius=0.05)[0]).all()


def gpu_normals(ctx, pts, radius=None, k=0):
    ctx.set_surface(pts)
    ctx.set_queries(None)
    ctx.set_viewpoint(0, 0, 10)
    nr = ctx.normals(radius=radius or 0.0, k=k)
    ctx.set_viewpoint(0, 0, 0)   # shared session context: restore the default
    return nr


@pytest.mark.gpu
@pytest.mark.parametrize("n,radius,k,dense", [(20000, 0.03, 0, False), (20000, 0.0, 24, False), (3000, 0.06, 0, True),
                                               (60000, 0.05, 0, False), (4000, 0.0, 16, True)])
def test_gpu_pfh125_equals_oracle(ctx, orc, n, radius, k, dense):
    pts = bumpy(n, 3)
    nr = gpu_normals(ctx, pts, radius=0.03)
    rng = np.random.default_rng(1)
    q = pts if dense else pts[rng.choice(n, 150, replace=False)]
    ctx.set_queries(None if dense else q)
    g = ctx.pfh125(radius=radius, k=k)
    ctx.set_queries(None)
    o, cnt = orc.pfh125(pts, nr, q, radius=radius, k=k, want_counts=True)
    assert g.shape == o.shape and not np.isnan(g).any()
    votes = cnt.sum(1).astype(np.float64)
    tol = 2 * 100.0 / np.maximum(votes, 1) + 1e-4
    assert np.all(np.abs(g - o).max(1) <= tol)
    assert (np.abs(g - o).max(1) == 0).mean() > 0.9   # almost every row bit-identical
    assert np.all(np.abs(g.sum(1) - 100) < 0.2)


@pytest.mark.gpu
def test_gpu_pfh125_edge_cases(ctx, orc):
    pts = bumpy(4000, 4)
    nr = gpu_normals(ctx, pts, radius=0.05)
    # far query -> NaN row; NaN query -> NaN row; query with a single neighbour -> zeros
    lone = np.array([[3.0, 3.0, 3.0]], np.float32)
    pts2 = np.vstack([pts, lone])
    nr2 = np.vstack([nr, [[0, 0, 1, 0]]]).astype(np.float32)
    ctx.set_surface(pts2)
    ctx.set_surface_normals(nr2)
    q = np.array([[9, 9, 9], [np.nan, 0, 0], [3.0, 3.0, 3.0], pts[5]], np.float32)
    ctx.set_queries(q)
    g = ctx.pfh125(radius=0.05)
    o = orc.pfh125(pts2, nr2, q, radius=0.05)
    ctx.set_queries(None)
    assert np.isnan(g[0]).all() and np.isnan(g[1]).all() and np.all(g[2] == 0)
    assert np.array_equal(np.isnan(g), np.isnan(o))
    assert np.abs(g[3] - o[3]).max() < 1e-2
    # a neighbourhood larger than the small stage (> 1024 points) goes through the second launch
    dense = bumpy(30000, 6)
    nrd = gpu_normals(ctx, dense, radius=0.02)
    qd = dense[:3]
    ctx.set_queries(qd)
    g = ctx.pfh125(radius=0.15)
    ctx.set_queries(None)
    cntn = orc.radius_count(dense, qd, 0.15)
    assert cntn.max() > 1024
    o = orc.pfh125(dense, nrd, qd, radius=0.15)
    assert np.abs(g - o).max() < 1e-2
    # normals missing -> error state
    ctx.set_surface(pts)
    with pytest.raises(RuntimeError):
        ctx.pfh125(radius=0.05)


@pytest.mark.gpu
@pytest.mark.parametrize("n,radius,k,dense", [(20000, 0.03, 0, True), (20000, 0.0, 20, True), (20000, 0.04, 0, False),
                                               (20000, 0.0, 12, False)])
def test_gpu_principal_curvatures_equal_oracle(ctx, orc, n, radius, k, dense):
    pts = bumpy(n, 7)
    nr = gpu_normals(ctx, pts, radius=0.03)
    q = pts if dense else pts[:500]   # upstream pairs query i with surface normal i
    ctx.set_queries(None if dense else q)
    g = ctx.principal_curvatures(radius=radius, k=k)
    ctx.set_queries(None)
    o, gap = orc.principal_curvatures(pts, nr, q, radius=radius, k=k)
    assert np.array_equal(np.isnan(g), np.isnan(o))
    ok = ~np.isnan(o[:, 0])
    assert ok.mean() > 0.99
    assert np.all(np.abs(g[ok, 3:] - o[ok, 3:]) <= 1e-6 + 1e-5 * np.abs(o[ok, 3:]))
    clear = ok & (gap > 0.05)
    assert clear.mean() > 0.8
    assert np.abs(g[clear, :3] - o[clear, :3]).max() < 1e-4
    assert np.all(g[ok, 3] >= g[ok, 4])


@pytest.mark.gpu
def test_gpu_curvatures_on_a_cylinder(ctx):
    rng = np.random.default_rng(0)
    n = 30000
    th, z = rng.uniform(0, 2 * np.pi, n), rng.uniform(0, 1, n)
    pts = np.c_[0.3 * np.cos(th), 0.3 * np.sin(th), z].astype(np.float32)
    nr = np.c_[np.cos(th), np.sin(th), 0 * th, 0 * th].astype(np.float32)
    ctx.set_surface(pts)
    ctx.set_surface_normals(nr)
    g = ctx.principal_curvatures(radius=0.05)
    assert np.all(np.abs(g[:, 4]) < 1e-7) and np.all(g[:, 3] > 1e-4)
    assert np.all(np.abs(g[:, 2]) < 1e-5)


@pytest.mark.gpu
def test_device_resident_buffers_give_the_same_rows(ctx):
    """PFX_DEVICE outputs (and a padded record stride) == PFX_HOST outputs for the widened descriptors and ICP"""
    import torch
    import ctypes as C
    import pcl_feature_extraction_b200 as pfx
    from pcl_feature_extraction_b200.capi import _ptr, DEVICE, IcpParams, IcpResult
    pts = bumpy(12000, 9)
    rng = np.random.default_rng(2)
    rgb = rng.integers(0, 1 << 24, len(pts)).astype(np.uint32)
    ctx.set_surface(pts)
    ctx.set_viewpoint(0, 0, 10)
    ctx.normals(radius=0.03, want_output=False)
    ctx.set_viewpoint(0, 0, 0)
    ctx.set_surface_colors(rgb)
    q = pts[:257]
    ctx.set_queries(q)
    ctx.set_query_colors(rgb[:257])
    dev = torch.device("cuda:0")
    h_pfh, h_pc = ctx.pfh125(radius=0.04), ctx.principal_curvatures(radius=0.04)
    h_sc, h_rf = ctx.shot1344(0.04)
    d_pfh = torch.full((257, 128), -7.0, dtype=torch.float32, device=dev)      # stride 512 B > 500 B
    d_pc = torch.full((257, 8), -7.0, dtype=torch.float32, device=dev)         # stride 32 B > 20 B
    d_sc = torch.full((257, 1353), -7.0, dtype=torch.float32, device=dev)
    ctx._chk(ctx.lib.pfx_pfh125(ctx.h, 0.04, 0, _ptr(d_pfh), 512, DEVICE))
    ctx._chk(ctx.lib.pfx_principal_curvatures(ctx.h, 0.04, 0, _ptr(d_pc), 32, DEVICE))
    ctx._chk(ctx.lib.pfx_shot1344(ctx.h, 0.04, None, _ptr(d_sc), 5412, DEVICE))
    ctx.sync()
    torch.cuda.synchronize()
    assert np.array_equal(d_pfh.cpu().numpy()[:, :125], h_pfh, equal_nan=True) and np.all(d_pfh.cpu().numpy()[:, 125:] == -7.0)
    assert np.array_equal(d_pc.cpu().numpy()[:, :5], h_pc, equal_nan=True) and np.all(d_pc.cpu().numpy()[:, 5:] == -7.0)
    assert np.array_equal(d_sc.cpu().numpy()[:, :1344], h_sc, equal_nan=True)
    assert np.array_equal(d_sc.cpu().numpy()[:, 1344:], h_rf, equal_nan=True)
    ctx.set_queries(None)
    # ICP with the source and the aligned cloud on the device (PointXYZRGB-like 32-byte records)
    src = (pts[::3] + np.float32([0.004, -0.003, 0.002])).astype(np.float32)
    h = ctx.icp_align(src, want_aligned=True)
    rec = np.zeros((len(src), 8), np.float32)
    rec[:, :3] = src
    d_src = torch.from_numpy(rec).to(dev)
    d_al = torch.zeros((len(src), 8), dtype=torch.float32, device=dev)
    prm, res = IcpParams(0.07, 100, 1e-6, 1e-4), IcpResult()
    ctx._chk(ctx.lib.pfx_icp_align(ctx.h, _ptr(d_src), len(src), 32, C.byref(prm), None, C.byref(res), _ptr(d_al), 32, DEVICE))
    torch.cuda.synchronize()
    assert np.array_equal(np.array(res.transform, np.float32).reshape(4, 4), h["T"]) and res.iterations == h["iterations"]
    assert res.fitness == h["fitness"]
    assert np.array_equal(d_al.cpu().numpy()[:, :3], h["aligned"]) and np.all(d_al.cpu().numpy()[:, 3:] == 0)


def test_oracle_moment_invariants_known_answer(orc):
    # a flat disc of radius r sampled densely: second moments about the centroid are n r^2 / 4 on the two in-plane
    # axes and 0 across -> j1 = n r^2 / 2, j2 = (n r^2 / 4)^2, j3 = 0
    rng = np.random.default_rng(1)
    n = 200000
    u = rng.uniform(-1, 1, (n, 2))
    pts = np.c_[u, np.zeros(n)].astype(np.float32)
    r = 0.2
    out = orc.moment_invariants(pts, np.zeros((1, 3), np.float32), radius=r)[0]
    m = orc.radius_count(pts, np.zeros((1, 3), np.float32), r)[0]
    assert abs(out[0] / (m * r * r / 2) - 1) < 0.02
    assert abs(out[1] / (m * r * r / 4) ** 2 - 1) < 0.04
    assert abs(out[2]) < 1e-6 * out[0] ** 3
    assert np.isnan(orc.moment_invariants(pts, np.array([[9, 9, 9]], np.float32), radius=r)).all()


@pytest.mark.gpu
@pytest.mark.parametrize("radius,k,dense", [(0.03, 0, True), (0.0, 24, True), (0.05, 0, False), (0.0, 16, False)])
def test_gpu_moment_invariants_equal_oracle(ctx, orc, radius, k, dense):
    """tolerance: both sides sum in double (about the query on the GPU, about the centroid on the CPU):
    1e-5 of the natural scale j1^p of each invariant"""
    pts = bumpy(20000, 12)
    q = pts if dense else (pts[:400] + np.float32(0.001))
    ctx.set_surface(pts)
    ctx.set_queries(None if dense else q)
    g = ctx.moment_invariants(radius=radius, k=k)
    ctx.set_queries(None)
    o = orc.moment_invariants(pts, q, radius=radius, k=k)
    assert np.array_equal(np.isnan(g), np.isnan(o)) and not np.isnan(o).all()
    ok = ~np.isnan(o[:, 0])
    j1 = np.abs(o[ok, 0]).astype(np.float64) + 1e-30
    assert np.all(np.abs(g[ok, 0] - o[ok, 0]) <= 1e-5 * j1)
    assert np.all(np.abs(g[ok, 1] - o[ok, 1]) <= 1e-5 * j1 ** 2)
    assert np.all(np.abs(g[ok, 2] - o[ok, 2]) <= 1e-5 * j1 ** 3)
    # far / non-finite queries -> NaN rows
    ctx.set_queries(np.array([[9, 9, 9], [np.nan, 0, 0]], np.float32))
    assert np.isnan(ctx.moment_invariants(radius=0.05)).all()
    ctx.set_queries(None)
