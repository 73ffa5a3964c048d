"""CPU tests of the NARF part of the oracle (range image, borders, keypoints, Narf36): closed-form known
answers on a ray-cast synthetic scene.  The oracle is test infrastructure; parity with real PCL is unpinned
(oracle/narf.cpp header)."""
import numpy as np
import pytest

W, H, F = 160, 120, 131.25


def raycast_scene(box=True):
    """points on the pixel-centre rays of a W x H pinhole camera: back plane z = 2, box front face z = 1.2"""
    u, v = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    dx, dy = (u - W / 2) / F, (v - H / 2) / F
    z = np.full_like(dx, 2.0)
    if box:
        hit = (np.abs(dx * 1.2) < 0.3) & (np.abs(dy * 1.2) < 0.2)
        z[hit] = 1.2
    pts = np.stack([dx * z, dy * z, z], -1).reshape(-1, 3).astype(np.float32)
    return pts, z


def test_planar_range_image_known_answer(orc):
    pts, z = raycast_scene(box=False)
    img, d = orc.range_image_planar(pts, W, H, W / 2, H / 2, F, F)
    u, v = np.meshgrid(np.arange(W), np.arange(H))
    expect = 2.0 * np.sqrt(((u - W / 2) / F) ** 2 + ((v - H / 2) / F) ** 2 + 1.0)
    assert np.isfinite(img[..., 3]).all()
    assert np.abs(img[..., 3] - expect).max() < 1e-5
    assert np.abs(img[..., 2] - 2.0).max() < 1e-5  # re-derived 3-D points lie on the plane


def test_zbuffer_keeps_the_nearest_and_is_order_independent(orc):
    pts, _ = raycast_scene(box=True)
    far = pts * 1.5  # same rays, farther: must lose every pixel
    both = np.concatenate([far, pts])
    a, _ = orc.range_image_planar(both, W, H, W / 2, H / 2, F, F)
    b, _ = orc.range_image_planar(both[::-1].copy(), W, H, W / 2, H / 2, F, F)
    c, _ = orc.range_image_planar(pts, W, H, W / 2, H / 2, F, F)
    assert np.array_equal(a[..., 3], b[..., 3]) and np.array_equal(a[..., 3], c[..., 3])


def test_spherical_image_is_cropped_to_the_observed_box(orc):
    pts, _ = raycast_scene(box=False)
    res = np.deg2rad(0.5)
    img, d = orc.range_image_spherical(pts, res)
    # field of view: 2 atan(80 / 131.25) = 62.7 deg x 2 atan(60 / 131.25) = 49.1 deg at 0.5 deg per pixel
    assert 120 <= d.width <= 130 and 95 <= d.height <= 102
    assert d.off_x > 0 and d.off_y > 0
    ok = np.isfinite(img[..., 3])
    assert ok.mean() > 0.8
    assert np.abs(img[ok][:, 2] - 2.0).max() < 0.02  # pixel-centre re-projection of points on the plane z = 2


def test_borders_of_a_box_in_front_of_a_plane(orc):
    pts, z = raycast_scene(box=True)
    img, d = orc.range_image_planar(pts, W, H, W / 2, H / 2, F, F)
    traits, scores, cs, cd = orc.narf_borders(img, d)
    traits = traits.reshape(H, W)
    box = z < 1.5
    obstacle, shadow, veil = (traits & 1) > 0, (traits & 2) > 0, (traits & 4) > 0
    assert obstacle.sum() > 50
    assert not (obstacle & ~box).any()      # obstacle borders lie on the nearer surface
    assert not (shadow & box).any()         # shadow borders on the farther one
    # obstacle borders hug the silhouette: within 2 pixels of the background, and they trace most of it (the
    # 3-pixel 1-D averaging of the border score puts the maximum one pixel inside the edge)
    def dilate(m, r):
        out = m.copy()
        for dy in range(-r, r + 1):
            for dx in range(-r, r + 1):
                out |= np.roll(np.roll(m, dy, 0), dx, 1)
        return out
    assert not (obstacle & ~dilate(~box, 2)).any()
    sil = box & dilate(~box, 1)
    assert (dilate(obstacle, 1) & sil).sum() >= 0.9 * sil.sum()
    assert not (veil & ~dilate(obstacle, 3)).any()   # veil points only between an obstacle border and its shadow
    # surface change: 1 on borders, ~0 in the flat interior of the plane
    cs = cs.reshape(H, W)
    assert cs[obstacle].min() > 0.4
    assert cs[10:20, 10:30].max() < 0.05


def test_narf_keypoints_sit_at_the_box_corners(orc):
    pts, z = raycast_scene(box=True)
    img, d = orc.range_image_planar(pts, W, H, W / 2, H / 2, F, F)
    kp, val, interest = orc.narf_keypoints(img, d, 0.2)
    assert 2 <= len(kp) <= 12 and (val >= 0.45).all()
    ys, xs = kp // W, kp % W
    box = z < 1.5
    by, bx = np.where(box)
    corners = np.array([[by.min(), bx.min()], [by.min(), bx.max()], [by.max(), bx.min()], [by.max(), bx.max()]])
    dist = np.abs(np.stack([ys, xs], 1)[:, None, :] - corners[None]).max(-1).min(1)
    assert (dist <= 6).all(), dist          # within a few pixels of a corner of the box
    assert np.array_equal(kp, np.sort(kp))  # ascending pixel index


def test_narf36_of_a_plane_is_flat_and_rotation_invariance_adds_rows(orc):
    pts, _ = raycast_scene(box=False)
    img, d = orc.range_image_planar(pts, W, H, W / 2, H / 2, F, F)
    kp = np.array([60 * W + 80, 40 * W + 50], np.int32)
    f = orc.narf36(img, d, kp, 0.2, rotation_invariant=False)
    assert f.shape == (2, 42)
    assert np.abs(f[:, 6:]).max() < 2e-2    # a plane seen along its normal: every beam is flat
    assert np.abs(f[:, 2] - 2.0).max() < 1e-3
    pts2, _ = raycast_scene(box=True)
    img2, d2 = orc.range_image_planar(pts2, W, H, W / 2, H / 2, F, F)
    kp2, _, _ = orc.narf_keypoints(img2, d2, 0.2)
    f1 = orc.narf36(img2, d2, kp2, 0.2, rotation_invariant=False)
    f2 = orc.narf36(img2, d2, kp2, 0.2, rotation_invariant=True)
    assert len(f1) == len(kp2) and len(f2) >= len(f1)
    assert np.isfinite(f2).all() and np.abs(f2[:, 6:]).max() <= 0.5 + 1e-6
    assert np.abs(f1[:, 6:]).max() > 0.05   # a corner is not flat
