"""Spin images (reference evaluation.cpp:515-554): the CPU oracle against a known answer, and
the CUDA path (through the C ABI) against the oracle.  Tolerance: both sides vote in double (the CPU sums in
neighbour order, the kernel in 2^-40 fixed point) and round the normalised image to float: 2e-7 absolute on cells
that sum to 1, except where a neighbour sits within double round-off of a bin edge (none observed)."""
import numpy as np
import pytest


def bumpy(n, seed):
    rng = np.random.default_rng(seed)
    u = rng.uniform(0, 1, (n, 2))
    z = 0.08 * np.sin(7 * u[:, 0]) * np.cos(5 * u[:, 1]) + 0.03 * np.sin(23 * u[:, 0] + 11 * u[:, 1])
    return np.c_[u, z].astype(np.float32)


def test_oracle_spin_image_on_a_plane(orc):
    rng = np.random.default_rng(0)
    g = np.stack(np.meshgrid(np.arange(200), np.arange(200)), -1).reshape(-1, 2) * 0.005 + rng.uniform(-.001, .001, (40000, 2))
    pl = np.c_[g, np.zeros(40000)].astype(np.float32)
    q = pl[20100:20101]
    o = orc.spin_image153(pl, q, np.array([[0, 0, 1, 0]], np.float32), 0.08).reshape(9, 17)
    # axis = plane normal: beta = 0 for every neighbour -> all weight in the centre column; the alpha profile of a
    # uniformly sampled disc grows linearly (ring area) up to the cylinder radius r / sqrt(2)
    assert abs(o.sum() - 1) < 1e-6 and abs(o[:, 8].sum() - 1) < 1e-6
    rows = o[:, 8]
    assert np.all(np.diff(rows[:8]) > 0)
    # axis in the plane: the image is symmetric in beta
    o2 = orc.spin_image153(pl, q, np.array([[1, 0, 0, 0]], np.float32), 0.08).reshape(9, 17)
    assert np.abs(o2 - o2[:, ::-1]).max() < 0.01
    # NaN normal -> NaN row; a lone point -> a zero image (one neighbour: itself, no normalisation)
    assert np.isnan(orc.spin_image153(pl, q, np.array([[np.nan, 0, 1, 0]], np.float32), 0.08)).all()
    lone = np.array([[5, 5, 5]], np.float32)
    assert np.all(orc.spin_image153(np.vstack([pl, lone]), lone, np.array([[0, 0, 1, 0]], np.float32), 0.08) == 0)


@pytest.mark.gpu
@pytest.mark.parametrize("n,r,dense,seed", [(20000, 0.08, False, 1), (5000, 0.1, True, 2), (60000, 0.04, False, 3)])
def test_gpu_spin_image_equals_oracle(ctx, orc, n, r, dense, seed):
    pts = bumpy(n, seed)
    nr_all, _, _ = orc.normals(pts, radius=0.04)
    sel = np.random.default_rng(seed).choice(n, 300, replace=False)
    q, qn = (pts, nr_all) if dense else (pts[sel], nr_all[sel])
    qn = qn.copy()
    qn[3] = np.nan                                   # a query without a normal
    ref = orc.spin_image153(pts, q, qn, r)
    ctx.set_surface(pts)
    ctx.set_queries(None if dense else q)
    g = ctx.spin_image153(r, qn)
    ctx.set_queries(None)
    assert np.array_equal(np.isnan(g), np.isnan(ref)) and np.isnan(g[3]).all()
    ok = ~np.isnan(ref[:, 0])
    assert np.abs(g[ok] - ref[ok]).max() <= 2e-7
    assert np.abs(g[ok].sum(1) - 1).max() < 1e-5


@pytest.mark.gpu
def test_gpu_spin_image_preconditions(ctx):
    import pcl_feature_extraction_b200 as pfx
    pts = bumpy(2000, 4)
    ctx.set_surface(pts)
    ctx.set_queries(pts[:5])
    with pytest.raises(pfx.PfxError) as e:          # wrong number of normals
        ctx.spin_image153(0.05, np.zeros((4, 4), np.float32))
    assert e.value.code == pfx.capi.E_PRECOND
    with pytest.raises(pfx.PfxError):               # k-search is not implemented for spin images
        ctx.spin_image153(0.0, np.zeros((5, 4), np.float32))
    far = np.array([[9, 9, 9]], np.float32)
    ctx.set_queries(far)
    assert np.all(ctx.spin_image153(0.05, np.array([[0, 0, 1, 0]], np.float32)) == 0)   # no neighbours: zero image
    ctx.set_queries(None)
