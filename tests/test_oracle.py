"""CPU tests of the oracle itself (no GPU): closed-form known answers, brute-force search parity,
invariances and the committed fixtures.  The oracle is a restatement of PCL 1.7 (PCL cannot be
built here; parity with real PCL is UNPINNED) - these tests are what pins it."""
import os

import numpy as np
import pytest


def plane(n=60, jitter=0.003, z=0.0, seed=0):
    rng = np.random.default_rng(seed)
    g = np.stack(np.meshgrid(np.arange(n), np.arange(n)), -1).reshape(-1, 2) * 0.01 + rng.uniform(-jitter, jitter, (n * n, 2))
    return np.concatenate([g, np.full((n * n, 1), z)], 1).astype(np.float32)


def rot(seed=3):
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q


def test_search_grid_equals_bruteforce(orc, clouds):
    xyz = clouds["indoor_target"][:15000]
    q = np.ascontiguousarray(clouds["indoor_target"][15000:15200])
    for r in (0.01, 0.05):
        a = orc.radius_search(xyz, q, r)
        b = orc.radius_search(xyz, q, r, brute=True)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    for k in (1, 2, 32):
        a = orc.knn(xyz, q, k)
        b = orc.knn(xyz, q, k, brute=True)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_search_semantics(orc):
    # strict '<' on (float)(r*r), self included, ties by index, sorted ascending
    surf = np.array([[0, 0, 0], [0.1, 0, 0], [0, 0.1, 0], [0.1, 0, 0], [0.2, 0, 0]], np.float32)
    off, idx, d2 = orc.radius_search(surf, surf[:1], 0.1)
    assert list(idx) == [0]  # 0.1^2 in float is not < (float)(0.1*0.1)
    off, idx, d2 = orc.radius_search(surf, surf[:1], 0.1000001)
    assert list(idx) == [0, 1, 2, 3]
    kidx, kd2 = orc.knn(surf, surf[:1], 3)
    assert list(kidx[0]) == [0, 1, 2]
    kidx, _ = orc.knn(surf, surf[:1], 8)
    assert list(kidx[0]) == [0, 1, 2, 3, 4, -1, -1, -1]


def test_normals_plane_and_sphere(orc):
    pl = plane(z=2.0)
    nr, cnt, gap = orc.normals(pl, radius=0.03, vp=(0, 0, 0))
    assert np.abs(nr[:, :3] - np.array([0, 0, -1.0])).max() < 1e-6
    assert np.abs(nr[:, 3]).max() < 1e-9
    rng = np.random.default_rng(1)
    v = rng.normal(size=(20000, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    sph = (0.5 * v + np.array([0, 0, 3.0])).astype(np.float32)
    nr, cnt, gap = orc.normals(sph, radius=0.05, vp=(0, 0, 3.0))  # viewpoint at the centre: inward normals
    assert np.abs(nr[:, :3] + v).max() < 6e-2  # finite-sample tilt of a 5 cm patch on a 0.5 m sphere
    nr_k, _, _ = orc.normals(sph, k=20, vp=(0, 0, 3.0))
    assert np.abs(nr_k[:, :3] + v).max() < 1e-1


def test_normals_pcl_float_variant_is_close_but_not_equal(orc, clouds):
    xyz = clouds["indoor_source"]
    q = np.ascontiguousarray(xyz[:2000])
    a, _, gap = orc.normals(xyz, q=q, radius=0.03, mode=0)
    b, _, _ = orc.normals(xyz, q=q, radius=0.03, mode=1)
    d = np.abs(a[:, :3] - b[:, :3]).max(1)
    assert 1e-4 < np.median(d) < 2e-2  # PCL's own float sums are ~3e-3 off


def test_normals_preconditions(orc):
    pl = plane(20)
    with pytest.raises(RuntimeError):
        orc.normals(pl, radius=0.03, k=5)
    with pytest.raises(RuntimeError):
        orc.normals(pl)


def test_fpfh_plane_known_answer(orc):
    pl = plane()
    nr = np.zeros((len(pl), 4), np.float32)
    nr[:, 2] = 1
    f = orc.fpfh(pl, nr, radius=0.05)
    expect = np.zeros(33, np.float32)
    expect[[5, 16, 27]] = 100
    assert np.abs(f - expect).max() < 1e-3
    fk = orc.fpfh(pl, nr, k=16)
    assert np.abs(fk - expect).max() < 1e-3


def test_fpfh_dihedral_edge(orc):
    # two perpendicular half planes: pairs across the edge have n1.n2 = 0 -> f1 = +-pi/2 -> bins 2 / 8
    a = plane(30, jitter=0.002, seed=1)
    b = a[:, [0, 2, 1]].copy()  # rotate the plane into x-z
    b[:, 2] = a[:, 1] + 0.0
    a[:, 1] = -a[:, 1] - 0.001
    pts = np.concatenate([a, b]).astype(np.float32)
    nr = np.zeros((len(pts), 4), np.float32)
    nr[: len(a), 2] = 1
    nr[len(a):, 1] = 1
    s = orc.spfh(pts, nr, np.arange(len(pts), dtype=np.int32), radius=0.03)
    near_edge = (np.abs(pts[:, 1]) < 0.004) & (np.abs(pts[:, 2]) < 0.004) & (pts[:, 0] > 0.05) & (pts[:, 0] < 0.24)
    h1 = s[near_edge][:, :11]
    assert (h1[:, [2, 8]].sum(1) > 5).all()  # cross-edge pairs
    assert (h1[:, [0, 1, 3, 4, 6, 7, 9, 10]].sum(1) < 1e-3).all()


def test_fpfh_shot_rigid_motion_invariance(orc, clouds):
    xyz = clouds["underwater_source"][:12000].astype(np.float64)
    R, t = rot(), np.array([0.3, -0.2, 0.5])
    xyz2 = (xyz @ R.T + t).astype(np.float32)
    xyz1 = xyz.astype(np.float32)
    n1, _, g1 = orc.normals(xyz1, radius=0.02)
    n2 = n1.copy()
    n2[:, :3] = n1[:, :3] @ R.T.astype(np.float32)
    q1, q2 = np.ascontiguousarray(xyz1[::60]), np.ascontiguousarray(xyz2[::60])
    f1, f2 = orc.fpfh(xyz1, n1, q1, radius=0.03), orc.fpfh(xyz2, n2, q2, radius=0.03)
    d = np.abs(f1 - f2).max(1)
    assert np.median(d) < 0.5 and (d < 5).mean() > 0.9  # float32 rotation moves a few votes across bins
    s1, r1 = orc.shot352(xyz1, n1, q1, 0.03)
    s2, r2 = orc.shot352(xyz2, n2, q2, 0.03)
    ok = ~np.isnan(s1[:, 0]) & ~np.isnan(s2[:, 0])
    ds = np.abs(s1[ok] - s2[ok]).max(1)
    assert np.median(ds) < 0.05


def test_shot_invariants(orc, clouds):
    xyz = clouds["indoor_source"][:30000]
    nr, _, _ = orc.normals(xyz, radius=0.03)
    q = np.ascontiguousarray(xyz[::150])
    s, rf = orc.shot352(xyz, nr, q, 0.05)
    ok = ~np.isnan(s[:, 0])
    assert ok.mean() > 0.9
    assert np.abs(np.linalg.norm(s[ok], axis=1) - 1).max() < 1e-6
    assert (s[ok] >= -1e-7).all()
    x, y, z = rf[ok, 0:3], rf[ok, 3:6], rf[ok, 6:9]
    assert np.abs((x * z).sum(1)).max() < 1e-5 and np.abs(np.linalg.norm(x, axis=1) - 1).max() < 1e-5
    assert np.abs(np.cross(z, x) - y).max() < 1e-6
    with pytest.raises(RuntimeError):
        orc.shot352(xyz, nr, q, 0.0)
    # isolated query -> NaN row
    s2, rf2 = orc.shot352(xyz, nr, np.array([[50, 50, 50]], np.float32), 0.05)
    assert np.isnan(s2).all() and np.isnan(rf2).all()


def test_iss_and_harris_properties(orc, clouds):
    xyz = orc.voxel_grid(clouds["indoor_target"], 0.01)
    assert len(xyz) == 33116  # 1 cm voxels of indoor/target
    res = orc.cloud_resolution(xyz)
    kp, sal = orc.iss(xyz, 6 * res, 4 * res)
    assert 300 < len(kp) < 3000 and (np.diff(kp) > 0).all()
    # every keypoint is a local maximum of the saliency among its non-max-radius neighbours
    off, idx, _ = orc.radius_search(xyz, xyz[kp], 4 * res)
    for i in range(len(kp)):
        nb = idx[off[i]:off[i + 1]]
        assert len(nb) >= 5 and sal[kp[i]] >= sal[nb].max()
    # plane: no ISS keypoints, Harris response = 0.04 + det - 0.04 tr^2 with identical normals = 0
    pl = plane(40)
    kp2, _ = orc.iss(pl, 0.03, 0.02)
    assert len(kp2) == 0
    nr = np.zeros((len(pl), 4), np.float32)
    nr[:, 2] = 1
    resp = orc.harris_response(pl, nr, 0.02)
    assert np.abs(resp).max() < 1e-7
    assert len(orc.harris_nms(pl, resp, 0.02, 1e-6)) == 0


def test_match_semantics(orc):
    a = np.array([[0, 0], [1, 0], [5, 5], [np.nan, 0]], np.float32)
    b = np.array([[1, 0], [0, 0], [0, 0], [np.nan, 1], [9, 9]], np.float32)
    idx, d2 = orc.match_nn(a, b)
    assert list(idx) == [1, 0, 4, -1]  # tie (b1 == b2) -> lowest index; NaN query never matches; NaN target skipped
    q, m, d = orc.match_reciprocal(a, b)
    assert list(q) == [0, 1, 2] and list(m) == [1, 0, 4]
    assert np.allclose(d, [0, 0, 32])


def test_cloud_facts_from_survey(orc, clouds):
    # SURVEY.md §6 table (measured independently with numpy/scipy while surveying)
    assert {k: len(v) for k, v in clouds.items()} == {"indoor_source": 101127, "indoor_target": 88555,
                                                      "underwater_source": 50759, "underwater_target": 53823}
    assert abs(orc.cloud_resolution(clouds["indoor_source"]) - 4.32e-3) < 2e-5
    assert abs(orc.cloud_resolution(clouds["underwater_source"]) - 3.02e-3) < 2e-5
    assert len(orc.voxel_grid(clouds["indoor_source"], 0.01)) == 41884
    off, _, _ = orc.radius_search(clouds["indoor_source"], np.ascontiguousarray(clouds["indoor_source"][::20]), 0.03)
    assert abs(np.diff(off).mean() - 91) < 3


def test_golden_fixture_regression(orc):
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_kat.npz"))
    crop, q = z["crop"], z["q"]
    idx, d2 = orc.knn(crop, q, 16)
    assert np.array_equal(idx, z["knn_idx"]) and np.array_equal(d2, z["knn_d2"])
    off, ridx, rd2 = orc.radius_search(crop, q, 0.02)
    assert np.array_equal(off, z["rad_off"]) and np.array_equal(ridx, z["rad_idx"])
    nr, _, _ = orc.normals(crop, radius=0.03)
    assert np.array_equal(nr, z["normals"])
    assert np.array_equal(orc.fpfh(crop, nr, q, radius=0.05), z["fpfh"])
    s, rf = orc.shot352(crop, nr, q, 0.05)
    assert np.array_equal(s, z["shot"], equal_nan=True)
    kp, _ = orc.iss(crop, 6 * z["resolution"][0], 4 * z["resolution"][0])
    assert np.array_equal(kp, z["iss_kp"])
    qq, mm, dd = orc.match_reciprocal(z["match_a"], z["match_b"])
    assert np.array_equal(qq, z["match_q"]) and np.array_equal(mm, z["match_m"])
