"""Point-to-point ICP (reference evaluation.cpp:863-885): the CPU oracle against closed-form
cases, and the CUDA path (through the C ABI) against the oracle.

Tolerances (GPU vs oracle): both run the same float iteration (correspondences, in-place float transform of the
source, float 4x4 products) and differ only in how the double-precision Umeyama moments are summed (sequential on
the CPU, a fixed tree on the GPU) and in the 3x3 rotation solve (Horn quaternion vs Kabsch / eigen(H^T H)), i.e.
at the 1e-15 level before each transform is rounded to float.  Hence: iteration count, convergence flag and state
equal; final transform within 2e-6 absolute; fitness within 1e-5 relative (+1e-12 absolute)."""
import numpy as np
import pytest

T_ATOL = 2e-6
FIT_RTOL = 1e-5


def sheet(n, seed):
    rng = np.random.default_rng(seed)
    u = rng.uniform(0, 1, (n, 2))
    z = 0.1 * np.sin(6 * u[:, 0]) * np.cos(5 * u[:, 1]) + 0.05 * np.sin(17 * u[:, 0] + 3 * u[:, 1])
    return np.c_[u, z].astype(np.float32)


def rigid(ax, ay, az, t):
    cx, sx, cy, sy, cz, sz = np.cos(ax), np.sin(ax), np.cos(ay), np.sin(ay), np.cos(az), np.sin(az)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = t
    return T


def unmoved(tgt, T, every=2):
    """source = T^-1 (every-th target point): ICP should find T"""
    p = tgt[::every].astype(np.float64)
    return ((p - T[:3, 3]) @ T[:3, :3]).astype(np.float32)


def test_oracle_recovers_a_small_rigid_motion(orc):
    tgt = sheet(20000, 1)
    T = rigid(0.01, -0.02, 0.03, [0.01, -0.008, 0.005])
    r = orc.icp(unmoved(tgt, T), tgt)
    assert r["converged"] and r["state"] in (2, 3, 4)
    assert np.abs(r["T"] - T).max() < 1e-5
    assert r["fitness"] < 1e-10
    assert 2 <= r["iterations"] < 100
    R = r["T"][:3, :3].astype(np.float64)
    assert np.abs(R @ R.T - np.eye(3)).max() < 1e-6 and abs(np.linalg.det(R) - 1) < 1e-6


def test_oracle_edge_cases(orc):
    tgt = sheet(5000, 2)
    # identical clouds: first iteration finds the identity -> TRANSFORM criterion
    r = orc.icp(tgt, tgt)
    assert r["converged"] and r["state"] == 2 and r["iterations"] == 1
    assert np.array_equal(r["T"], np.eye(4, dtype=np.float32)) and r["fitness"] == 0.0
    # source out of reach of the correspondence distance: NO_CORRESPONDENCES, not converged, transform = guess
    far = tgt + np.float32(10.0)
    r = orc.icp(far, tgt)
    assert not r["converged"] and r["state"] == 5 and r["iterations"] == 0
    assert np.array_equal(r["T"], np.eye(4, dtype=np.float32))
    assert r["fitness"] > 100.0  # getFitnessScore has no distance bound
    # iteration limit: do-while runs max_iterations iterations and reports ITERATIONS
    T = rigid(0.02, 0.02, 0.02, [0.02, 0.01, -0.01])
    r = orc.icp(unmoved(tgt, T), tgt, max_iterations=3)
    assert r["converged"] and r["state"] == 1 and r["iterations"] == 3
    # a guess that is already the answer converges at once
    r = orc.icp(unmoved(tgt, T), tgt, guess=T.astype(np.float32))
    assert r["converged"] and r["iterations"] <= 2 and np.abs(r["T"] - T).max() < 1e-5
    # non-finite source points are ignored
    s = unmoved(tgt, T)
    s2 = s.copy()
    s2[::7] = np.nan
    r2 = orc.icp(s2, tgt)
    assert r2["converged"] and np.abs(r2["T"] - T).max() < 1e-4


def check_equal(g, o):
    assert (g["converged"], g["state"], g["iterations"]) == (o["converged"], o["state"], o["iterations"])
    assert np.abs(g["T"] - o["T"]).max() <= T_ATOL
    if o["fitness"] < 1e300:
        assert abs(g["fitness"] - o["fitness"]) <= FIT_RTOL * o["fitness"] + 1e-12
    else:
        assert g["fitness"] > 1e300


@pytest.mark.gpu
@pytest.mark.parametrize("n,angles,t,seed", [
    (20000, (0.01, -0.02, 0.03), (0.01, -0.008, 0.005), 1),
    (6000, (0.0, 0.0, 0.05), (0.02, 0.02, 0.0), 2),
    (30000, (-0.03, 0.01, 0.0), (0.0, 0.0, 0.03), 3),
])
def test_gpu_icp_equals_oracle(ctx, orc, n, angles, t, seed):
    tgt = sheet(n, seed)
    T = rigid(*angles, t)
    rng = np.random.default_rng(seed)
    src = unmoved(tgt, T) + rng.normal(0, 0.0005, (len(tgt[::2]), 3)).astype(np.float32)
    ctx.set_surface(tgt)
    g = ctx.icp_align(src, want_aligned=True)
    o = orc.icp(src, tgt)
    check_equal(g, o)
    assert np.abs(g["T"] - T).max() < 5e-3
    # the aligned cloud is the source moved by the final transform (float, ((m0 x + m1 y) + m2 z) + m3)
    M = g["T"]
    exp = np.empty_like(src)
    for r in range(3):
        v = M[r, 0] * src[:, 0]
        v = v + M[r, 1] * src[:, 1]
        v = v + M[r, 2] * src[:, 2]
        exp[:, r] = v + M[r, 3]
    assert np.array_equal(g["aligned"], exp)


@pytest.mark.gpu
def test_gpu_icp_edge_cases(ctx, orc):
    tgt = sheet(5000, 2)
    ctx.set_surface(tgt)
    check_equal(ctx.icp_align(tgt), orc.icp(tgt, tgt))
    far = tgt + np.float32(10.0)
    g = ctx.icp_align(far)
    check_equal(g, orc.icp(far, tgt))
    assert g["state"] == 5 and g["correspondences"] == 0
    T = rigid(0.02, 0.02, 0.02, [0.02, 0.01, -0.01])
    s = unmoved(tgt, T)
    check_equal(ctx.icp_align(s, max_iterations=3), orc.icp(s, tgt, max_iterations=3))
    check_equal(ctx.icp_align(s, guess=T.astype(np.float32)), orc.icp(s, tgt, guess=T.astype(np.float32)))
    s2 = s.copy()
    s2[::7] = np.nan
    check_equal(ctx.icp_align(s2), orc.icp(s2, tgt))
    # half of the source out of reach: only the other half corresponds
    s3 = s.copy()
    s3[: len(s3) // 2] += np.float32(5.0)
    g = ctx.icp_align(s3)
    check_equal(g, orc.icp(s3, tgt))
    assert 0 < g["correspondences"] <= len(s3) - len(s3) // 2
    # empty source
    g = ctx.icp_align(np.zeros((0, 3), np.float32))
    assert not g["converged"] and g["state"] == 5
    # run-to-run reproducible (fixed reduction order)
    a, b = ctx.icp_align(s), ctx.icp_align(s)
    assert np.array_equal(a["T"], b["T"]) and a["fitness"] == b["fitness"]


@pytest.mark.gpu
def test_gpu_icp_on_the_indoor_pair(ctx, orc, clouds):
    """the reference's use: direct ICP of the (voxel-filtered) indoor source cloud onto the target cloud"""
    xyz = []
    for name in ("indoor_source", "indoor_target"):
        ctx.set_surface(clouds[name])
        xyz.append(ctx.voxel_grid(0.02))  # (leaf 0.02 keeps the CPU oracle to a few seconds)
    ctx.set_surface(xyz[1])
    g = ctx.icp_align(xyz[0])
    o = orc.icp(xyz[0], xyz[1])
    # 30 - 100 iterations of float round-off on real data: iteration counts may differ by the last criterion check
    assert g["converged"] == o["converged"]
    assert abs(g["iterations"] - o["iterations"]) <= 1
    assert np.abs(g["T"] - o["T"]).max() < 1e-4
    assert abs(g["fitness"] - o["fitness"]) <= 1e-3 * o["fitness"]
    R = g["T"][:3, :3].astype(np.float64)
    assert np.abs(R @ R.T - np.eye(3)).max() < 1e-5 and abs(np.linalg.det(R) - 1) < 1e-5
