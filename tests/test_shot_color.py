"""SHOT1344 = SHOT shape + CIELab colour (reference evaluation.cpp:786-805): the CPU oracle
against known answers, and the CUDA path (through the C ABI) against the oracle.

Tolerances (GPU vs oracle, identical normals and frames): unit-L2 descriptor rows within 1e-4 max-abs (int32
fixed-point accumulation + float trigonometry in the continuous interpolation weights, as for SHOT352); the
normalised Lab triplets are bit-identical (same lookup tables, float arithmetic without FMA)."""
import numpy as np
import pytest


def textured_sheet(n, seed):
    rng = np.random.default_rng(seed)
    u = rng.uniform(0, 1, (n, 2))
    z = 0.06 * np.sin(8 * u[:, 0]) * np.cos(6 * u[:, 1])
    pts = np.c_[u, z].astype(np.float32)
    r = (127 + 120 * np.sin(25 * u[:, 0])).astype(np.uint32)
    g = (127 + 120 * np.cos(19 * u[:, 1])).astype(np.uint32)
    b = rng.integers(0, 256, n).astype(np.uint32)
    return pts, (r << 16) | (g << 8) | b


def test_oracle_lab_and_layout(orc):
    pts, rgb = textured_sheet(6000, 1)
    nr, _, _ = orc.normals(pts, radius=0.05)
    primaries = np.array([0xffffff, 0x000000, 0xff0000, 0x00ff00, 0x0000ff], np.uint32)
    _, _, lab = orc.shot1344(pts[:5], primaries, nr[:5], pts[:1], primaries[:1], 0.08, want_lab=True)
    lab = lab * [100, 120, 120]
    # sRGB -> CIELab (D65) of white, black and the primaries, to the resolution of PCL's 4000-entry table
    exp = np.array([[100, 0, 0], [0, 0, 0], [53.24, 80.09, 67.20], [87.73, -86.18, 83.18], [32.30, 79.19, -107.86]])
    assert np.abs(lab - exp).max() < 0.15
    q, qrgb = pts[:80], rgb[:80]
    out, rf = orc.shot1344(pts, rgb, nr, q, qrgb, 0.08)
    ok = ~np.isnan(out[:, 0])
    assert ok.mean() > 0.9
    assert np.abs(np.linalg.norm(out[ok], axis=1) - 1).max() < 1e-5 and (out[ok] >= 0).all()
    # the first 352 slots are SHOT352's histogram before normalisation; both channels carry the same total weight
    s352, rf352 = orc.shot352(pts, nr, q, 0.08)
    a = out[ok, :352]
    assert np.abs(a / np.linalg.norm(a, axis=1, keepdims=True) - s352[ok]).max() < 1e-6
    assert np.array_equal(rf[ok], rf352[ok])
    assert np.abs(out[ok, :352].sum(1) - out[ok, 352:].sum(1)).max() < 1e-4
    # uniform colour: every neighbour has colour distance 0 -> only colour slot 0 (and its interpolation partner) of
    # each volume is used
    flat = np.full(len(pts), 0x336699, np.uint32)
    o2, _ = orc.shot1344(pts, flat, nr, q, flat[:80], 0.08)
    col = o2[ok, 352:].reshape(-1, 32, 31)
    assert np.all(col[:, :, 1:] == 0) and np.all(col[:, :, 0].sum(1) > 0)


@pytest.mark.gpu
@pytest.mark.parametrize("n,radius,dense,seed", [(20000, 0.05, False, 2), (5000, 0.08, True, 3), (40000, 0.03, False, 4)])
def test_gpu_shot1344_equals_oracle(ctx, orc, n, radius, dense, seed):
    pts, rgb = textured_sheet(n, seed)
    nr, _, _ = orc.normals(pts, radius=0.04)
    sel = np.random.default_rng(seed).choice(n, 300, replace=False)
    q, qrgb = (pts, rgb) if dense else (pts[sel], rgb[sel])
    ref, rf_ref = orc.shot1344(pts, rgb, nr, q, qrgb, radius)
    ctx.set_surface(pts)
    ctx.set_surface_normals(nr)
    ctx.set_surface_colors(rgb)
    ctx.set_queries(None if dense else q)
    if not dense:
        ctx.set_query_colors(qrgb)
    # identical frames on both sides: the descriptor stage alone
    s, rf = ctx.shot1344(radius, lrf_in=rf_ref)
    assert np.array_equal(np.isnan(s[:, 0]), np.isnan(ref[:, 0]))
    ok = ~np.isnan(ref[:, 0])
    assert ok.mean() > 0.9
    d = np.abs(s[ok] - ref[ok]).max(1)
    assert d.max() <= 1e-4, (d.max(), np.quantile(d, 0.99))
    assert np.abs(np.linalg.norm(s[ok], axis=1) - 1).max() < 1e-5
    # end to end (frames estimated on the GPU): rows with a clear frame agree
    s2, rf2 = ctx.shot1344(radius)
    _, gap = orc.shot_lrf(pts, q, radius)
    good = ok & (gap.min(1) > 1e-2)
    assert good.mean() > 0.7
    assert (np.abs(s2[good] - ref[good]).max(1) <= 1e-4).mean() > 0.99
    ctx.set_queries(None)


@pytest.mark.gpu
def test_gpu_shot1344_preconditions(ctx):
    import pcl_feature_extraction_b200 as pfx
    pts, rgb = textured_sheet(3000, 5)
    ctx.set_surface(pts)
    nr = ctx.normals(radius=0.05)
    with pytest.raises(pfx.PfxError) as e:      # no colours yet
        ctx.shot1344(0.05)
    assert e.value.code == pfx.capi.E_STATE
    with pytest.raises(pfx.PfxError):           # wrong number of colours
        ctx.set_surface_colors(rgb[:10])
    ctx.set_surface_colors(rgb)
    with pytest.raises(pfx.PfxError) as e:      # SHOT needs a radius
        ctx.shot1344(0.0)
    assert e.value.code == pfx.capi.E_PRECOND
    ctx.set_queries(pts[:7])
    with pytest.raises(pfx.PfxError) as e:      # query colours missing
        ctx.shot1344(0.05)
    assert e.value.code == pfx.capi.E_STATE
    ctx.set_query_colors(rgb[:7])
    s, rf = ctx.shot1344(0.05)
    assert s.shape == (7, 1344) and not np.isnan(s).any()
    # a new surface invalidates the colours
    ctx.set_surface(pts)
    ctx.set_surface_normals(nr)
    with pytest.raises(pfx.PfxError):
        ctx.shot1344(0.05)
    # a far query -> NaN descriptor and frame
    ctx.set_surface_colors(rgb)
    ctx.set_queries(np.array([[9, 9, 9]], np.float32))
    ctx.set_query_colors(np.array([0], np.uint32))
    s, rf = ctx.shot1344(0.05)
    assert np.isnan(s).all() and np.isnan(rf).all()
    ctx.set_queries(None)


@pytest.mark.gpu
def test_gpu_binary_pcd_payload_is_the_record_layout(ctx, tmp_path):
    """a binary PCD payload (x, y, z, rgb records of 16 bytes) goes to pfx_set_surface / pfx_set_surface_colors as it is"""
    import ctypes as C
    from pcl_feature_extraction_b200.pcd import write_pcd, read_pcd
    from pcl_feature_extraction_b200.capi import HOST
    pts, rgb = textured_sheet(8000, 8)
    write_pcd(tmp_path / "c.pcd", pts, rgb)
    xyz2, rgb2, hdr = read_pcd(tmp_path / "c.pcd")
    assert np.array_equal(xyz2, pts) and np.array_equal(rgb2, rgb)
    raw = open(tmp_path / "c.pcd", "rb").read()
    payload = np.frombuffer(raw[raw.index(b"DATA binary\n") + len(b"DATA binary\n"):], dtype=np.uint8).copy()
    assert len(payload) == 16 * len(pts)
    # reference result from separate arrays
    ctx.set_surface(pts)
    ctx.set_viewpoint(0, 0, 0)
    ctx.normals(radius=0.04, want_output=False)
    ctx.set_surface_colors(rgb)
    s_ref, rf_ref = ctx.shot1344(0.05)
    # the same from the file payload, stride 16
    base = payload.ctypes.data
    ctx._chk(ctx.lib.pfx_set_surface(ctx.h, C.c_void_p(base), len(pts), 16, HOST))
    ctx.normals(radius=0.04, want_output=False)
    ctx._chk(ctx.lib.pfx_set_surface_colors(ctx.h, C.c_void_p(base + 12), len(pts), 16, HOST))
    s, rf = ctx.shot1344(0.05)
    assert np.array_equal(s, s_ref, equal_nan=True) and np.array_equal(rf, rf_ref, equal_nan=True)
