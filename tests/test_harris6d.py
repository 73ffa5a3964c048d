"""Harris 6D keypoints (reference keypoints.h:166-179; in its active detector list, evaluation.cpp:63-65): the CPU
oracle against known answers, and the CUDA path (through the C ABI) against the oracle.

Contract: intensity, gradient estimation, the 6x6 covariance and its 4th eigenvalue follow written-out
rules where upstream leans on Eigen internals (float column-pivoting QR, 6x6 eigen solver); both sides follow them
operation for operation in reference order, so the response is BIT-identical and the keypoint indices, refined
corners and snapped cloud indices are equal."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def colours():
    z = np.load(os.path.join(ROOT, "tests", "golden", "clouds_rgb.npz"))
    return {k: z[k] for k in z.files}


def textured_plane(n=6000, seed=0):
    """a plane z = 0.1 x with a sharp intensity edge along x = 0.5: the gradient is along +x in the plane"""
    rng = np.random.default_rng(seed)
    xy = rng.uniform(0, 1, (n, 2)) * np.array([1.0, 0.3])
    pts = np.c_[xy, 0.1 * xy[:, 0]].astype(np.float32)
    level = np.where(pts[:, 0] > 0.5, 230, 20).astype(np.uint32)
    rgb = (level << 16) | (level << 8) | level
    return pts, rgb


def test_oracle_harris6d_known_answers(orc):
    pts, rgb = textured_plane()
    r = 0.03
    nr, _, _ = orc.normals(pts, radius=r)
    resp, grad, inten = orc.harris6d_response(pts, rgb, nr, r)
    # grey level g -> g / 256 (0.114 + 0.587 + 0.2989 = 0.9999)
    assert np.abs(inten - 0.9999 * (rgb & 255) / 256.0).max() < 1e-6
    # gradients: unit vectors along the in-plane x direction at the edge, zero (below upstream's length 200) far from it
    at_edge = np.abs(pts[:, 0] - 0.5) < 0.01
    far = np.abs(pts[:, 0] - 0.5) > 0.05
    assert np.all(grad[far] == 0)
    ge = grad[at_edge]
    nz = np.linalg.norm(ge, axis=1) > 0
    assert nz.mean() > 0.9
    ge = ge[nz]
    assert np.abs(np.linalg.norm(ge, axis=1) - 1).max() < 1e-5
    tangent = np.array([1.0, 0.0, 0.1]) / np.hypot(1.0, 0.1)
    assert (ge @ tangent).min() > 0.85 and np.median(ge @ tangent) > 0.98
    assert np.abs((ge * nr[at_edge, :3][nz]).sum(1)).max() < 1e-4    # projected onto the tangent plane
    # away from the edge every (normal, gradient) vector is (n, 0): rank one, the 4th smallest eigenvalue vanishes; along
    # the straight edge the gradients share one direction up to sampling scatter: a small response
    assert np.abs(resp[np.abs(pts[:, 0] - 0.5) > 0.08]).max() < 1e-9
    assert np.median(resp[at_edge]) < 0.01
    # crossing a second edge (a checker corner) adds a third direction near the corner only
    level = np.where((pts[:, 0] > 0.5) ^ (pts[:, 1] > 0.15), 230, 20).astype(np.uint32)
    rgb2 = (level << 16) | (level << 8) | level
    resp2, _, _ = orc.harris6d_response(pts, rgb2, nr, r)
    corner = np.hypot(pts[:, 0] - 0.5, pts[:, 1] - 0.15) < 0.02
    away = np.hypot(pts[:, 0] - 0.5, pts[:, 1] - 0.15) > 0.08
    assert np.median(resp2[corner]) > 5 * max(resp2[away].max(), 1e-9) and np.median(resp2[corner]) > 0.1
    kp = orc.harris_nms(pts, resp2, r, 1e-4)
    best = kp[np.argmax(resp2[kp])]
    assert np.hypot(pts[best, 0] - 0.5, pts[best, 1] - 0.15) < 0.03      # the strongest keypoint is the checker corner


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["underwater_source", "underwater_target"])
def test_gpu_harris6d_equals_oracle_on_the_bundled_clouds(ctx, orc, clouds, colours, name):
    pts, rgb = clouds[name], colours[name]
    ctx.set_surface(pts)
    ctx.set_queries(None)
    ctx.set_viewpoint(0, 0, 0)
    ctx.set_surface_colors(rgb)
    h = ctx.harris6d(0.01, 1e-6)
    nr, _, _ = orc.normals(pts, radius=0.01)
    resp, grad, inten = orc.harris6d_response(pts, rgb, nr, 0.01)
    kp = orc.harris_nms(pts, resp, 0.01, 1e-6)
    corners = orc.harris_refine(pts, nr, 0.01, pts[kp].copy())
    sn = orc.snap_to_cloud(pts, corners, 1e-4)
    assert np.array_equal(h["response"].view(np.uint32), resp.view(np.uint32)), "Harris 6D response must be bit-identical"
    assert np.array_equal(h["kp_idx"], kp)
    assert np.array_equal(h["kp_xyz"].view(np.uint32), corners.view(np.uint32))
    assert np.array_equal(h["snapped_idx"], sn)
    assert len(kp) > 50 and (resp > 1e-6).mean() > 0.001
    # the detector's private normals must not leak into the context (FeatureFromNormals stages still need theirs)
    import pcl_feature_extraction_b200 as pfx
    with pytest.raises(pfx.PfxError):
        ctx.fpfh(radius=0.02)


@pytest.mark.gpu
def test_gpu_harris6d_synthetic_corner_and_preconditions(ctx, orc):
    import pcl_feature_extraction_b200 as pfx
    pts, _ = textured_plane(8000, 3)
    level = np.where((pts[:, 0] > 0.5) ^ (pts[:, 1] > 0.15), 230, 20).astype(np.uint32)
    rgb = (level << 16) | (level << 8) | level
    ctx.set_surface(pts)
    ctx.set_queries(None)
    with pytest.raises(pfx.PfxError) as e:  # no colours yet
        ctx.harris6d(0.03, 1e-6)
    assert e.value.code == pfx.capi.E_STATE
    ctx.set_surface_colors(rgb)
    h = ctx.harris6d(0.03, 1e-4)
    nr, _, _ = orc.normals(pts, radius=0.03)
    resp, _, _ = orc.harris6d_response(pts, rgb, nr, 0.03)
    assert np.array_equal(h["response"].view(np.uint32), resp.view(np.uint32))
    kp = orc.harris_nms(pts, resp, 0.03, 1e-4)
    assert np.array_equal(h["kp_idx"], kp)
    # the strongest keypoint sits at the checker corner
    best = h["kp_idx"][np.argmax(resp[h["kp_idx"]])]
    assert np.hypot(pts[best, 0] - 0.5, pts[best, 1] - 0.15) < 0.03
    # caller's normals survive the call
    ctx.set_surface_normals(nr)
    ctx.harris6d(0.03, 1e-4)
    f = ctx.fpfh(radius=0.03)
    assert not np.isnan(f).all()
