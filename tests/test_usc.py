"""Unique Shape Context (reference evaluation.cpp:344-371): the CPU oracle against known
answers, and the CUDA path (through the C ABI) against the oracle.

Tolerance (GPU vs oracle, identical frames): the oracle adds the float weights of a bin sequentially in the
distance order of the neighbours, the kernel accumulates them exactly (64-bit fixed point) and rounds once, so
bins agree to the round-off of a float sum: 1e-5 relative (+1e-6 of the row maximum).  CUDA's atan2f / acosf differ
from libm by an ulp, which can move a neighbour that sits on a bin boundary into the adjacent bin: at most 2 % of
the rows may carry such a moved weight."""
import numpy as np
import pytest


def bumpy(n, seed):
    rng = np.random.default_rng(seed)
    u = rng.uniform(0, 1, (n, 2))
    z = 0.08 * np.sin(7 * u[:, 0]) * np.cos(5 * u[:, 1]) + 0.03 * np.sin(23 * u[:, 0] + 11 * u[:, 1])
    return np.c_[u, z].astype(np.float32)


def test_oracle_usc_known_answers(orc):
    pts = bumpy(20000, 1)
    q = pts[:60]
    r = 0.08
    out, rf, dens = orc.usc1980(pts, q, r, local_radius=0.15)
    assert not np.isnan(out).any() and (out >= 0).all()
    cnt = orc.radius_count(pts, pts, r / 5.0)
    assert np.array_equal(dens, cnt)                                     # density = neighbours at r / 5, itself included
    # every neighbour (but the query itself) lands in exactly one bin with weight 1 / density / cbrt(volume): the
    # number of occupied bins cannot exceed the neighbour count, and scaling the cloud by s scales every weight by 1 / s
    nb = orc.radius_count(pts, q, r)
    assert np.all((out > 0).sum(1) <= nb - 1)
    out2, rf2, _ = orc.usc1980(pts * np.float32(2), q * np.float32(2), 2 * r, local_radius=0.3)
    ratio = out2[out > 0] / out[out > 0]
    assert np.abs(ratio - 0.5).max() < 1e-3
    # frames are SHOT's at the local radius; a far query has a NaN frame -> NaN row, zero frame
    rf_shot, _ = orc.shot_lrf(pts, q, 0.15)
    assert np.array_equal(rf, rf_shot)
    far = np.array([[9, 9, 9]], np.float32)
    o, f, _ = orc.usc1980(pts, far, r, local_radius=0.15)
    assert np.isnan(o).all() and np.all(f == 0)


@pytest.mark.gpu
@pytest.mark.parametrize("n,r,local,dense,seed", [(20000, 0.08, 0.15, False, 2), (4000, 0.1, 2.5, True, 3),
                                                    (40000, 0.05, 0.1, False, 4)])
def test_gpu_usc_equals_oracle(ctx, orc, n, r, local, dense, seed):
    pts = bumpy(n, seed)
    sel = np.random.default_rng(seed).choice(n, 200, replace=False)
    q = pts if dense else pts[sel]
    ref, rf_ref, _ = orc.usc1980(pts, q, r, local_radius=local)
    ctx.set_surface(pts)
    ctx.set_queries(None if dense else q)
    g, rf = ctx.usc1980(r, local_radius=local, lrf_in=rf_ref)       # identical frames: the descriptor stage alone
    assert np.array_equal(np.isnan(g[:, 0]), np.isnan(ref[:, 0]))
    ok = ~np.isnan(ref[:, 0])
    assert ok.mean() > 0.9 and np.array_equal(rf[ok], rf_ref[ok])
    tol = 1e-5 * np.abs(ref[ok]) + 1e-6 * ref[ok].max(1, keepdims=True)
    row_ok = np.all(np.abs(g[ok] - ref[ok]) <= tol, axis=1)
    assert row_ok.mean() >= 0.98, row_ok.mean()
    assert np.abs(g[ok].sum(1) - ref[ok].sum(1)).max() <= 1e-4 * ref[ok].sum(1).max()   # a moved weight keeps the row sum
    # end to end (frames estimated on the GPU at the local radius)
    g2, rf2 = ctx.usc1980(r, local_radius=local)
    _, gap = orc.shot_lrf(pts, q, local)
    clear = ok & (gap.min(1) > 1e-2)
    assert clear.mean() > 0.5
    assert np.abs(rf2[clear] - rf_ref[clear]).max() < 1e-5
    ctx.set_queries(None)


@pytest.mark.gpu
def test_gpu_usc_preconditions_and_nan_rows(ctx):
    import pcl_feature_extraction_b200 as pfx
    pts = bumpy(3000, 5)
    ctx.set_surface(pts)
    ctx.set_queries(np.array([[9, 9, 9], pts[0]], np.float32))
    g, rf = ctx.usc1980(0.08, local_radius=0.2)
    assert np.isnan(g[0]).all() and np.all(rf[0] == 0) and not np.isnan(g[1]).any()
    with pytest.raises(pfx.PfxError) as e:      # search radius below the minimal radius
        ctx.usc1980(0.01, min_radius=0.02)
    assert e.value.code == pfx.capi.E_PRECOND
    with pytest.raises(pfx.PfxError):
        ctx.usc1980(0.0)
    ctx.set_queries(None)
