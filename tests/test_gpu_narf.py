"""GPU parity of the range image / NARF path (narf.cu) against the CPU oracle (oracle/narf.cpp), through the
C ABI.  Stage-wise: every stage after the projection is fed the ORACLE's range image, so that a difference is
attributable to that stage.  Index outputs (border traits, keypoint pixels, descriptor counts) must be
identical; floats within the tolerance written at the assert."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W, H, F = 320, 240, 262.5


@pytest.fixture(scope="module")
def indoor(clouds):
    return clouds["indoor_source"]


@pytest.fixture(scope="module")
def oracle_planar(orc, indoor):
    return orc.range_image_planar(indoor, W, H, W / 2, H / 2, F, F)


def to_pfx_desc(pfx, d):
    return pfx.capi.RangeImageDesc(d.width, d.height, d.planar, d.cx, d.cy, d.fx, d.fy, d.ang_res, d.off_x, d.off_y)


def test_range_image_planar_bit_exact(ctx, orc, indoor, oracle_planar):
    oimg, od = oracle_planar
    ctx.set_surface(indoor)
    d = ctx.range_image_planar(W, H, W / 2, H / 2, F, F)
    gd, img = ctx.range_image_get()
    assert (gd.width, gd.height, gd.planar) == (W, H, 1)
    # ranges come from IEEE sqrt / div only: bit-identical, including which pixels are unobserved (-inf)
    assert np.array_equal(img[..., 3].view(np.uint32), oimg[..., 3].view(np.uint32))
    ok = np.isfinite(oimg[..., 3])
    assert ok.sum() > 30000
    assert np.abs(img[ok][:, :3] - oimg[ok][:, :3]).max() <= 1e-6
    assert np.isnan(img[~ok][:, 0]).all()


def test_range_image_planar_reference_size_and_order_independence(ctx, orc, indoor):
    """the reference's own geometry (640 x 480, f = 525: keypoints.h:204-216); shuffled input, same image"""
    oimg, _ = orc.range_image_planar(indoor, 640, 480, 320, 240, 525, 525)
    rng = np.random.default_rng(3)
    ctx.set_surface(indoor[rng.permutation(len(indoor))])
    ctx.range_image_planar(640, 480, 320.0, 240.0, 525.0, 525.0)
    _, img = ctx.range_image_get()
    assert np.array_equal(img[..., 3].view(np.uint32), oimg[..., 3].view(np.uint32))


def test_range_image_spherical_c3(ctx, orc, indoor):
    """config C3: 0.5 degree spherical image, cropped"""
    res = float(np.deg2rad(0.5))
    oimg, od = orc.range_image_spherical(indoor, res)
    ctx.set_surface(indoor)
    d = ctx.range_image_spherical(res)
    gd, img = ctx.range_image_get()
    assert (gd.width, gd.height, gd.off_x, gd.off_y) == (od.width, od.height, od.off_x, od.off_y)
    # atan2 / asin / cos differ by an ulp between libm and CUDA: a point on a pixel boundary may move
    same = img[..., 3].view(np.uint32) == oimg[..., 3].view(np.uint32)
    assert same.mean() > 0.995, same.mean()
    both = np.isfinite(img[..., 3]) & np.isfinite(oimg[..., 3]) & same
    assert np.abs(img[both][:, :3] - oimg[both][:, :3]).max() <= 1e-5


def test_empty_and_degenerate_inputs(ctx, orc):
    ctx.set_surface(np.zeros((0, 3), np.float32))
    d = ctx.range_image_spherical(float(np.deg2rad(0.5)))
    assert d.width == 0 and d.height == 0
    kp, xyz, val, _ = ctx.narf_keypoints(0.2)
    assert len(kp) == 0
    pts = np.array([[0, 0, 1], [np.nan, 0, 1], [0.1, 0, 2]], np.float32)
    ctx.set_surface(pts)
    ctx.range_image_planar(64, 48, 32.0, 24.0, 50.0, 50.0)
    _, img = ctx.range_image_get()
    oimg, _ = orc.range_image_planar(pts, 64, 48, 32, 24, 50, 50)
    assert np.array_equal(img[..., 3].view(np.uint32), oimg[..., 3].view(np.uint32))
    assert np.isfinite(img[..., 3]).sum() == 3   # two direct hits + one floor/ceil splat of the half-pixel hit
    kp, _, _, _ = ctx.narf_keypoints(0.2)
    assert len(kp) == 0


def test_narf_borders_given_image(ctx, orc, oracle_planar):
    import pcl_feature_extraction_b200 as pfx
    oimg, od = oracle_planar
    ctx.range_image_set(to_pfx_desc(pfx, od), oimg)
    traits, scores, cs, cd = ctx.narf_borders()
    otraits, oscores, ocs, ocd = orc.narf_borders(oimg, od)
    assert (otraits & 1).sum() > 500
    assert np.array_equal(traits, otraits)                      # border classification: identical bit sets
    assert np.abs(scores - oscores).max() <= 1e-5               # border scores
    assert np.abs(cs - ocs).max() <= 1e-4                       # surface-change score (sqrt of a PCA eigenvalue)
    strong = ocs > 0.2                                           # direction = principal axis: sign-free, defined
    dots = np.abs((cd[strong] * ocd[strong]).sum(1))            # where the score is not ~0
    assert (dots > 1 - 1e-4).mean() > 0.995


def test_narf_keypoints_given_image(ctx, orc, oracle_planar):
    import pcl_feature_extraction_b200 as pfx
    oimg, od = oracle_planar
    ctx.range_image_set(to_pfx_desc(pfx, od), oimg)
    kp, xyz, val, interest = ctx.narf_keypoints(0.2)
    okp, oval, ointerest = orc.narf_keypoints(oimg, od, 0.2)
    assert np.abs(interest - ointerest).max() <= 1e-4           # interest image
    assert len(okp) >= 3
    assert np.array_equal(kp, okp)                              # keypoint pixels: identical
    assert np.abs(val - oval).max() <= 1e-4
    assert np.allclose(xyz, oimg.reshape(-1, 4)[okp, :3])


def test_narf36_given_image_and_keypoints(ctx, orc, oracle_planar):
    import pcl_feature_extraction_b200 as pfx
    oimg, od = oracle_planar
    ctx.range_image_set(to_pfx_desc(pfx, od), oimg)
    okp, _, _ = orc.narf_keypoints(oimg, od, 0.2)
    # keypoints plus a spread of ordinary pixels (flat regions, edges, unobserved ones)
    extra = np.arange(0, W * H, 997, dtype=np.int32)
    kps = np.concatenate([okp, extra]).astype(np.int32)
    for rot in (False, True):
        f = ctx.narf36(kps, 0.2, rotation_invariant=rot)
        of = orc.narf36(oimg, od, kps, 0.2, rotation_invariant=rot)
        assert len(of) > 20
        assert f.shape == of.shape                              # same keypoints rejected, same rotation counts
        assert np.abs(f[:, :3] - of[:, :3]).max() <= 1e-5       # position
        assert np.abs(f[:, 6:] - of[:, 6:]).max() <= 1e-4       # descriptor in [-0.5, 0.5]
        ang = np.abs(np.angle(np.exp(1j * (f[:, 3:6] - of[:, 3:6]))))
        assert ang.max() <= 1e-3                                # roll / pitch / yaw


def test_narf_end_to_end_c3(ctx, orc, indoor):
    """config C3 end to end on the GPU: spherical image -> keypoints -> rotation-invariant Narf36, against the
    oracle run on its own image (the two images may differ in a few boundary pixels, see above)"""
    res = float(np.deg2rad(0.5))
    ctx.set_surface(indoor)
    ctx.range_image_spherical(res)
    gd, img = ctx.range_image_get()
    kp, xyz, val, _ = ctx.narf_keypoints(0.2)
    f = ctx.narf36(kp, 0.2, True)
    # the oracle on the GPU's image must agree exactly in indices
    from oracle.binding import RiDesc
    od = RiDesc(gd.width, gd.height, 0, 0.0, 0.0, 1.0, 1.0, gd.ang_res, gd.off_x, gd.off_y)
    okp, oval, _ = orc.narf_keypoints(img, od, 0.2)
    assert np.array_equal(kp, okp)
    of = orc.narf36(img, od, okp, 0.2, True)
    assert f.shape == of.shape
    if len(of):
        assert np.abs(f[:, 6:] - of[:, 6:]).max() <= 1e-4


def _pose(axis, angle, t):
    axis = np.asarray(axis, np.float64) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)
    P = np.eye(4)
    P[:3, :3], P[:3, 3] = R, t
    return P


@pytest.mark.parametrize("planar", [True, False])
def test_sensor_pose_moves_the_world_not_the_image(ctx, indoor, planar):
    """keypoints.h:207-210 hands RangeImage the pose translation(sensor_origin_) * rotation(sensor_orientation_).  A
    cloud moved rigidly together with its sensor must give the image of the unmoved cloud under the identity pose:
    the same ranges (up to the float rounding of the transform), the image points, keypoint positions and Narf36
    poses moved by the pose, the same descriptor values."""
    P = _pose([0.3, -1.0, 0.5], 0.7, [0.4, -1.1, 2.0])
    R, t = P[:3, :3], P[:3, 3]
    moved = (indoor.astype(np.float64) @ R.T + t).astype(np.float32)

    def run(cloud, pose):
        ctx.set_surface(cloud)
        ctx.range_image_set_pose(pose)
        if planar:
            ctx.range_image_planar(W, H, W / 2, H / 2, F, F)
        else:
            ctx.range_image_spherical(float(np.deg2rad(0.5)))
        d, img = ctx.range_image_get()
        kp, xyz, val, _ = ctx.narf_keypoints(0.2)
        f = ctx.narf36(kp, 0.2, True)
        return d, img, kp, xyz, f

    try:
        d0, img0, kp0, xyz0, f0 = run(indoor, None)
        d1, img1, kp1, xyz1, f1 = run(moved, P)
    finally:
        ctx.range_image_set_pose(None)
    assert (d0.width, d0.height, d0.off_x, d0.off_y) == (d1.width, d1.height, d1.off_x, d1.off_y)
    r0, r1 = img0[..., 3], img1[..., 3]
    both = np.isfinite(r0) & np.isfinite(r1)
    assert (np.isfinite(r0) == np.isfinite(r1)).mean() > 0.999
    # the moved cloud lands in the sensor frame within ~1e-6 m of the original: ranges agree to that, except where a
    # point sat on a pixel boundary and now feeds the neighbouring pixel
    close = np.abs(r0[both] - r1[both]) <= 2e-5
    assert close.mean() > 0.995, close.mean()
    # image points are world coordinates: those of the unmoved image, moved
    want = img0[..., :3][both].astype(np.float64) @ R.T + t
    good = np.abs(r0[both] - r1[both]) <= 2e-5
    assert np.abs(img1[..., :3][both][good] - want[good]).max() < 1e-4
    # keypoints: (almost) the same pixels, positions moved by the pose, descriptors unchanged
    common = np.intersect1d(kp0, kp1)
    assert len(common) >= 0.8 * max(len(kp0), 1) and len(kp0) > 5
    i0 = np.searchsorted(kp0, common)
    i1 = np.searchsorted(kp1, common)
    assert np.abs(xyz1[i1] - (xyz0[i0].astype(np.float64) @ R.T + t)).max() < 1e-4
    assert abs(len(f0) - len(f1)) <= max(2, 0.2 * len(f0))
    # descriptors of keypoints that yield one orientation each in both runs: same values, positions moved
    if len(f0) == len(f1) and np.array_equal(kp0, kp1):
        assert np.abs(f1[:, :3] - (f0[:, :3].astype(np.float64) @ R.T + t)).max() < 1e-3
        assert np.median(np.abs(f1[:, 6:] - f0[:, 6:]).max(1)) < 1e-3
