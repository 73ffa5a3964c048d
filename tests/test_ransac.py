"""RANSAC correspondence rejection (reference features.h:282-297): the CPU oracle against
closed-form cases, and the CUDA path (through the C ABI) against the oracle."""
import numpy as np
import pytest


def planted(n=400, outlier_share=0.6, seed=1, noise=0.0):
    rng = np.random.default_rng(seed)
    src = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    ang = 0.4
    R = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
    t = np.array([0.3, -0.2, 0.5])
    moved = src @ R.T + t + rng.normal(0, noise, (n, 3))
    out = rng.random(n) < outlier_share
    moved[out] = rng.uniform(-2, 2, (out.sum(), 3))
    perm = rng.permutation(n).astype(np.int32)   # correspondence i: source i <-> target perm[i]
    tgt = np.zeros((n, 3), np.float32)
    tgt[perm] = moved.astype(np.float32)
    return src, tgt, np.arange(n, dtype=np.int32), perm, ~out, R, t


def test_oracle_recovers_a_planted_transform(orc):
    src, tgt, q, m, inl, R, t = planted()
    keep, T, it, bh = orc.ransac_reject(src, tgt, q, m)
    assert np.array_equal(keep, inl)                    # exactly the planted inliers survive
    assert np.abs(T[:3, :3] - R).max() < 1e-6 and np.abs(T[:3, 3] - t).max() < 1e-6
    assert np.allclose(T[3], [0, 0, 0, 1])
    assert 1 <= it <= 1001 and 0 <= bh < it
    # adaptive bound: with inlier share w the loop stops at k = log(1 - 0.99) / log(1 - w^3) iterations
    w = inl.mean()
    k = np.log(0.01) / np.log(1 - w ** 3)
    assert it == max(bh + 1, int(np.ceil(k)))


def test_oracle_edge_cases(orc):
    src, tgt, q, m, inl, R, t = planted(n=50)
    keep, T, it, bh = orc.ransac_reject(src, tgt, q[:2], m[:2])  # PCL: < 3 correspondences -> all kept, identity
    assert keep.all() and np.array_equal(T, np.eye(4, dtype=np.float32)) and it == 0
    # pure outliers: no consensus; the loop runs to max_iterations + 1 and only a handful survive
    rng = np.random.default_rng(3)
    a = rng.uniform(-1, 1, (60, 3)).astype(np.float32)
    b = rng.uniform(-1, 1, (60, 3)).astype(np.float32)
    keep, T, it, bh = orc.ransac_reject(a, b, np.arange(60, dtype=np.int32), np.arange(60, dtype=np.int32))
    assert keep.sum() <= 8 and it == 1001
    # coincident points only: every sample is degenerate, nothing survives
    z = np.zeros((10, 3), np.float32)
    keep, T, it, bh = orc.ransac_reject(z, z, np.arange(10, dtype=np.int32), np.arange(10, dtype=np.int32))
    assert keep.sum() == 0 and np.array_equal(T, np.eye(4, dtype=np.float32))
    # the seed is part of the contract: same seed, same answer; the planted answer is found whatever the seed
    k1 = orc.ransac_reject(src, tgt, q, m, seed=7)
    k2 = orc.ransac_reject(src, tgt, q, m, seed=7)
    k3 = orc.ransac_reject(src, tgt, q, m, seed=8)
    assert np.array_equal(k1[0], k2[0]) and k1[2:] == k2[2:] and np.array_equal(k1[0], k3[0])


@pytest.mark.gpu
@pytest.mark.parametrize("n,share,noise,seed", [(400, 0.6, 0.0, 1), (2000, 0.8, 0.002, 2), (37, 0.3, 0.001, 3), (3, 0.0, 0.0, 4)])
def test_gpu_ransac_equals_oracle(ctx, orc, n, share, noise, seed):
    import pcl_feature_extraction_b200 as pfx
    src, tgt, q, m, inl, R, t = planted(n, share, seed, noise)
    corr = np.zeros(n, pfx.capi.CORR_DTYPE)
    corr["index_query"], corr["index_match"] = q, m
    corr["distance"] = np.arange(n, dtype=np.float32)   # payload must travel with the survivors
    out, T, it, bh = ctx.ransac_reject(src, tgt, corr, seed=seed)
    keep, oT, oit, obh = orc.ransac_reject(src, tgt, q, m, seed=seed)
    assert (it, bh) == (oit, obh)                       # same sequential decisions
    assert np.array_equal(out["index_query"], q[keep]) and np.array_equal(out["index_match"], m[keep])
    assert np.array_equal(out["distance"], corr["distance"][keep])
    assert np.abs(T - oT).max() < 1e-5                  # two different rotation solvers, both in double
    if noise == 0.0 and n > 3:
        assert np.array_equal(keep, inl)


@pytest.mark.gpu
def test_gpu_ransac_edge_cases(ctx, orc):
    import pcl_feature_extraction_b200 as pfx
    src, tgt, q, m, inl, R, t = planted(50)
    corr = np.zeros(2, pfx.capi.CORR_DTYPE)
    corr["index_query"], corr["index_match"] = q[:2], m[:2]
    out, T, it, bh = ctx.ransac_reject(src, tgt, corr)
    assert len(out) == 2 and np.array_equal(T, np.eye(4, dtype=np.float32))
    out, T, it, bh = ctx.ransac_reject(src, tgt, corr[:0])
    assert len(out) == 0
    z = np.zeros((10, 3), np.float32)
    c10 = np.zeros(10, pfx.capi.CORR_DTYPE)
    c10["index_query"] = c10["index_match"] = np.arange(10)
    out, T, it, bh = ctx.ransac_reject(z, z, c10)
    assert len(out) == 0 and np.array_equal(T, np.eye(4, dtype=np.float32))
    bad = c10.copy()
    bad["index_match"][3] = 99
    with pytest.raises(RuntimeError):
        ctx.ransac_reject(z, z, bad)


@pytest.mark.gpu
def test_gpu_ransac_device_buffers_validate_indices(ctx, orc):
    """PFX_DEVICE correspondences (e.g. straight from pfx_match) with a -1 / out-of-range index: a clean error,
    not an out-of-bounds read; valid device buffers give the host-buffer result."""
    import ctypes as C
    import torch
    import pcl_feature_extraction_b200 as pfx
    src, tgt, q, m, inl, R, t = planted(200)
    corr = np.zeros(len(q), pfx.capi.CORR_DTYPE)
    corr["index_query"], corr["index_match"] = q, m
    ref, Tref, it_ref, bh_ref = ctx.ransac_reject(src, tgt, corr)

    def run(c):
        d_src = torch.from_numpy(np.ascontiguousarray(src, np.float32)).cuda()
        d_tgt = torch.from_numpy(np.ascontiguousarray(tgt, np.float32)).cuda()
        d_c = torch.from_numpy(c.view(np.int32).reshape(-1, 3).copy()).cuda()
        d_out = torch.zeros_like(d_c)
        T = np.zeros(16, np.float32)
        n_out = C.c_size_t(0)
        it, bh = C.c_int(0), C.c_int(0)
        torch.cuda.synchronize()
        ctx._chk(ctx.lib.pfx_ransac_reject(ctx.h, d_src.data_ptr(), len(src), 12, d_tgt.data_ptr(), len(tgt), 12,
                                           d_c.data_ptr(), len(c), 0.015, 1000, 12345, d_out.data_ptr(), len(c),
                                           C.byref(n_out), pfx.capi._ptr(T), C.byref(it), C.byref(bh), pfx.capi.DEVICE))
        return d_out[: n_out.value].cpu().numpy(), T.reshape(4, 4), it.value, bh.value

    out, T, it, bh = run(corr)
    assert (it, bh) == (it_ref, bh_ref) and np.array_equal(out[:, 0], ref["index_query"]) and np.array_equal(T, Tref)
    for col, val in (("index_match", -1), ("index_query", len(src)), ("index_match", 1 << 30)):
        bad = corr.copy()
        bad[col][7] = val
        with pytest.raises(RuntimeError):
            run(bad)
    out, T, it, bh = run(corr)  # the context is still usable
    assert np.array_equal(out[:, 0], ref["index_query"])


@pytest.mark.gpu
def test_gpu_ransac_on_config_c1_correspondences(ctx, orc, clouds):
    """the reference's use: filterCorrespondences on the ISS keypoints + FPFH matches of the indoor pair"""
    kps, feats = [], []
    for name in ("indoor_source", "indoor_target"):
        ctx.set_surface(clouds[name])
        xyz = ctx.voxel_grid(0.01)
        ctx.set_surface(xyz)
        ctx.set_viewpoint(0, 0, 0)
        ctx.normals(radius=0.03, want_output=False)
        res = ctx.cloud_resolution()
        kp, _ = ctx.iss(6 * res, 4 * res)
        ctx.set_queries(xyz[kp])
        feats.append(ctx.fpfh(radius=0.05))
        ctx.set_queries(None)
        kps.append(xyz[kp])
    corr = ctx.match(feats[0], feats[1], reciprocal=True)
    out, T, it, bh = ctx.ransac_reject(kps[0], kps[1], corr, threshold=0.015, max_iterations=1000)
    keep, oT, oit, obh = orc.ransac_reject(kps[0], kps[1], corr["index_query"], corr["index_match"], 0.015, 1000)
    assert (it, bh) == (oit, obh)
    assert np.array_equal(out["index_query"], corr["index_query"][keep])
    assert np.abs(T - oT).max() < 1e-5
    # (FPFH matches of this pair are mostly wrong: RANSAC may find no consensus in 1000 draws; a least-squares fit
    # of three wrong pairs need not even reproduce its own samples)
    assert len(out) == keep.sum() <= len(corr)
    # a rigid transform: R orthonormal, det +1
    Rm = T[:3, :3].astype(np.float64)
    assert np.abs(Rm @ Rm.T - np.eye(3)).max() < 1e-5 and abs(np.linalg.det(Rm) - 1) < 1e-5
