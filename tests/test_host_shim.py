"""The C++ host side (pcl_compat.hpp + feature_pipeline.hpp + evaluation_b200.cpp): the reference's
evaluate() sequencing through the pcl::Feature-style shim, compared with the Python-bound C ABI and
the oracle.  Needs a GPU (the driver calls the CUDA library; there is no CPU path)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "pcl_feature_extraction_b200", "lib", "evaluation_b200")


def test_shim_compiles_against_the_c_abi():
    """CPU: the driver binary is built by __graft_entry__.build() and links the C-ABI library."""
    assert os.path.exists(BIN), "run __graft_entry__.build()"
    out = subprocess.run(["ldd", BIN], capture_output=True, text=True).stdout
    assert "libpfx_b200.so" in out


@pytest.mark.gpu
def test_evaluation_driver_matches_c_abi_and_oracle(tmp_path, clouds, ctx, orc):
    from pcl_feature_extraction_b200.pcd import write_pcd
    src = np.ascontiguousarray(clouds["underwater_source"][:30000])
    tgt = np.ascontiguousarray(clouds["underwater_target"][:30000])
    rng = np.random.default_rng(11)
    src_rgb = rng.integers(0, 1 << 24, len(src)).astype(np.uint32)
    tgt_rgb = rng.integers(0, 1 << 24, len(tgt)).astype(np.uint32)
    write_pcd(tmp_path / "s.pcd", src, src_rgb)
    write_pcd(tmp_path / "t.pcd", tgt, tgt_rgb)
    # (last argument 0: the throughput kernels, which is what the Python context of this test runs too; the driver's
    # default, reference-order arithmetic, is covered by test_evaluation_driver_strict_and_reuse)
    r = subprocess.run([BIN, str(tmp_path / "s.pcd"), str(tmp_path / "t.pcd"), "0.05", "0.03", str(tmp_path), "0"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    rows = [l.split(", ") for l in lines[1:] if not l.startswith("#")]
    assert sum(l.startswith("# direct ICP") for l in lines) == 1
    byname = {(x[0], x[1]): x for x in rows}
    for kp_name in ("Harris3D", "Iss"):
        for d_name in ("ShapeContext", "FPFH", "SHOT", "SHOTColor", "SpinImage", "USC", "MomentInvariants", "PFH",
                       "PrincipalCurvatures"):
            assert (kp_name, d_name) in byname
    # NARF row (present when both clouds yield keypoints): the shim's RangeImagePlanar / NarfKeypoint /
    # NarfDescriptor objects against the C ABI called from Python on the same cloud
    ctx.set_surface(src)
    ctx.range_image_planar(640, 480, 320.0, 240.0, 525.0, 525.0)
    kp_px, kp_xyz_abi, _, _ = ctx.narf_keypoints(0.2)
    if ("Narf", "NARF") in byname:
        px = np.fromfile(tmp_path / "Narf_src_px.bin", dtype=np.int32)
        assert np.array_equal(px, kp_px)
        f36 = np.fromfile(tmp_path / "Narf_NARF_src.bin", dtype=np.float32).reshape(-1, 42)
        assert np.array_equal(f36, ctx.narf36(kp_px, 0.2, True))
        assert int(byname[("Narf", "NARF")][6]) == len(f36)

    def kp_xyz(name):
        a = np.fromfile(tmp_path / name, dtype=np.float32).reshape(-1, 8)
        return np.ascontiguousarray(a[:, :3])

    # ISS keypoints: same as the C ABI called from Python, and as the oracle NMS on the GPU saliency
    ctx.set_viewpoint(0, 0, 0)   # the clouds' sensor origin (VIEWPOINT 0 0 0 ...), what the shim passes
    ctx.set_surface(src)
    res = ctx.cloud_resolution()
    kp, sal = ctx.iss(6 * res, 4 * res)
    assert np.array_equal(kp_xyz("Iss_src_kp.bin"), src[kp])
    assert np.array_equal(kp, orc.iss_nms(src, sal, 4 * res))
    # Harris keypoints after the snap
    h = ctx.harris3d(0.01, 1e-6)
    snapped = h["snapped_idx"][h["snapped_idx"] >= 0]
    assert np.array_equal(kp_xyz("Harris3D_src_kp.bin"), src[snapped])
    # FPFH at the ISS keypoints: shim == C ABI bit for bit; vs oracle within the FPFH tolerance
    f_shim = np.fromfile(tmp_path / "Iss_FPFH_src.bin", dtype=np.float32).reshape(-1, 33)
    ctx.set_surface(src)
    ctx.set_queries(None)
    nr = ctx.normals(radius=0.03)
    ctx.set_queries(src[kp])
    f_abi = ctx.fpfh(radius=0.05)
    assert np.array_equal(f_shim, f_abi)
    f_or = orc.fpfh(src, nr, src[kp], radius=0.05)
    assert (np.abs(f_shim - f_or).max(1) <= 1e-3).mean() >= 0.99
    # SHOT rows carry descriptor[352] + rf[9]
    s_shim = np.fromfile(tmp_path / "Iss_SHOT_src.bin", dtype=np.float32).reshape(-1, 361)
    s_abi, rf_abi = ctx.shot352(0.05)
    assert np.array_equal(s_shim[:, :352], s_abi, equal_nan=True) and np.array_equal(s_shim[:, 352:], rf_abi, equal_nan=True)
    # correspondences: reciprocal exact matching of the dumped descriptors == oracle, bit-exact indices
    f_tgt = np.fromfile(tmp_path / "Iss_FPFH_tgt.bin", dtype=np.float32).reshape(-1, 33)
    corr = np.fromfile(tmp_path / "Iss_FPFH_corr.bin", dtype=np.int32).reshape(-1, 3)
    q, m, d = orc.match_reciprocal(f_shim, f_tgt)
    assert np.array_equal(corr[:, 0], q) and np.array_equal(corr[:, 1], m)
    assert np.array_equal(corr[:, 2].view(np.float32), d)
    assert int(byname[("Iss", "FPFH")][8]) == len(q)
    # RANSAC rejection through the shim == the C ABI called from Python == the oracle
    kp_s, kp_t = kp_xyz("Iss_src_kp.bin"), kp_xyz("Iss_tgt_kp.bin")
    filt = np.fromfile(tmp_path / "Iss_FPFH_filtered.bin", dtype=np.int32).reshape(-1, 3)
    tf = np.fromfile(tmp_path / "Iss_FPFH_tf.bin", dtype=np.float32).reshape(4, 4)
    keep, oT, _, _ = orc.ransac_reject(kp_s, kp_t, q, m, 0.015, 1000)
    assert np.array_equal(filt[:, 0], q[keep]) and np.array_equal(filt[:, 1], m[keep])
    assert np.abs(tf - oT).max() < 1e-5
    assert int(byname[("Iss", "FPFH")][9]) == int(keep.sum())
    # ICP through the shim's IterativeClosestPoint == the C ABI called from Python (bit for bit); the keypoint
    # clouds' ICP also against the oracle
    icp = np.fromfile(tmp_path / "direct_icp.bin", dtype=np.float32)
    ctx.set_surface(tgt)
    g = ctx.icp_align(src)
    assert np.array_equal(icp[:16].reshape(4, 4), g["T"]) and icp[16] == np.float32(g["fitness"]) and bool(icp[17]) == g["converged"]
    icp_kp = np.fromfile(tmp_path / "Iss_icp_kp.bin", dtype=np.float32)
    o = orc.icp(kp_s, kp_t)
    assert bool(icp_kp[17]) == o["converged"]
    assert np.abs(icp_kp[:16].reshape(4, 4) - o["T"]).max() < 1e-5
    if o["fitness"] < 1e300:
        assert abs(icp_kp[16] - o["fitness"]) <= 1e-4 * o["fitness"] + 1e-12
    # PFH125 / PrincipalCurvatures rows through the shim == the C ABI called from Python (bit for bit)
    ctx.set_surface(src)
    ctx.set_queries(None)
    ctx.normals(radius=0.03, want_output=False)
    ctx.set_queries(src[kp])
    p_shim = np.fromfile(tmp_path / "Iss_PFH_src.bin", dtype=np.float32).reshape(-1, 125)
    assert np.array_equal(p_shim, ctx.pfh125(radius=0.05), equal_nan=True)
    c_shim = np.fromfile(tmp_path / "Iss_PrincipalCurvatures_src.bin", dtype=np.float32).reshape(-1, 5)
    assert np.array_equal(c_shim, ctx.principal_curvatures(radius=0.05), equal_nan=True)
    ctx.set_queries(None)
    # SHOT1344 through the shim (colours from the PointXYZRGB records) == the C ABI called from Python
    ctx.set_surface(src)
    ctx.set_queries(None)
    ctx.normals(radius=0.03, want_output=False)
    ctx.set_surface_colors(src_rgb)
    for kp_name in ("Iss", "Harris3D"):
        rec = np.fromfile(tmp_path / (kp_name + "_src_kp.bin"), dtype=np.float32).reshape(-1, 8)
        q_rgb = np.ascontiguousarray(rec[:, 4]).view(np.uint32) & 0xffffff
        # ISSKeypoint3D copies only x, y, z into its output points (colour stays PointXYZRGB's default black), the
        # Harris keypoints are cloud points found by the 1 cm snap and keep their colour
        if kp_name == "Iss":
            assert np.all(q_rgb == 0)
        else:
            assert np.array_equal(q_rgb, src_rgb[snapped])
        ctx.set_queries(np.ascontiguousarray(rec[:, :3]))
        ctx.set_query_colors(q_rgb)
        sc_shim = np.fromfile(tmp_path / (kp_name + "_SHOTColor_src.bin"), dtype=np.float32).reshape(-1, 1353)
        sc_abi, sc_rf = ctx.shot1344(0.05)
        assert np.array_equal(sc_shim[:, :1344], sc_abi, equal_nan=True) and np.array_equal(sc_shim[:, 1344:], sc_rf, equal_nan=True)
    ctx.set_queries(None)
    # USC through the shim (r / 10, r / 5, PCL's local radius 2.5) == the C ABI called from Python
    rec = np.fromfile(tmp_path / "Iss_src_kp.bin", dtype=np.float32).reshape(-1, 8)
    ctx.set_surface(src)
    ctx.set_queries(np.ascontiguousarray(rec[:, :3]))
    u_shim = np.fromfile(tmp_path / "Iss_USC_src.bin", dtype=np.float32).reshape(-1, 1989)
    u_abi, u_rf = ctx.usc1980(0.05)
    assert np.array_equal(u_shim[:, :1980], u_abi, equal_nan=True) and np.array_equal(u_shim[:, 1980:], u_rf, equal_nan=True)
    ctx.set_queries(None)
    # spin images through the shim (normals estimated on the keypoint cloud, as the reference does) == C ABI from Python
    sp_shim = np.fromfile(tmp_path / "Iss_SpinImage_src.bin", dtype=np.float32).reshape(-1, 153)
    sp_nrm = np.fromfile(tmp_path / "Iss_SpinImage_src_normals.bin", dtype=np.float32).reshape(-1, 8)
    kq = np.ascontiguousarray(rec[:, :3])
    ctx.set_surface(kq)                      # Tools::estimateNormals(keypoints): the keypoint cloud is its own surface
    ctx.set_queries(None)
    assert np.array_equal(ctx.normals(radius=0.03)[:, :3], sp_nrm[:, :3], equal_nan=True)
    ctx.set_surface(src)
    ctx.set_queries(kq)
    assert np.array_equal(sp_shim, ctx.spin_image153(0.05, np.ascontiguousarray(sp_nrm[:, :4])), equal_nan=True)
    ctx.set_queries(None)


@pytest.mark.gpu
def test_evaluation_driver_strict_and_reuse(tmp_path, clouds, orc):
    """the driver's default mode on the bundled underwater pair WITH its colours: reference-order arithmetic
    (PFX_PARITY_STRICT), so the Harris3D / Harris6D keypoints and the FPFH rows equal the oracle's bit for bit, the
    3DSC row is there, and the reference's redundancy (the same cloud and its normals re-submitted for every
    descriptor type, features.h:186-193) costs one upload per cloud and one normals pass per cloud and detector"""
    from pcl_feature_extraction_b200.pcd import write_pcd
    rgbz = np.load(os.path.join(ROOT, "tests", "golden", "clouds_rgb.npz"))
    src = np.ascontiguousarray(clouds["underwater_source"])
    tgt = np.ascontiguousarray(clouds["underwater_target"])
    write_pcd(tmp_path / "s.pcd", src, rgbz["underwater_source"])
    write_pcd(tmp_path / "t.pcd", tgt, rgbz["underwater_target"])
    r = subprocess.run([BIN, str(tmp_path / "s.pcd"), str(tmp_path / "t.pcd"), "0.05", "0.03", str(tmp_path)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    rows = {(x[0], x[1]): x for x in (l.split(", ") for l in lines[1:] if not l.startswith("#"))}
    for kp_name in ("Harris3D", "Harris6D", "Iss"):
        for d_name in ("ShapeContext", "FPFH", "SHOT", "USC"):
            assert (kp_name, d_name) in rows, (kp_name, d_name, r.stdout[-2000:])

    def kp_xyz(name):
        return np.ascontiguousarray(np.fromfile(tmp_path / name, dtype=np.float32).reshape(-1, 8)[:, :3])

    # Harris3D: the oracle's whole chain (normals r = 1 cm, response, NMS, refinement, 1 cm snap)
    nr1, _, _ = orc.normals(src, radius=0.01)
    resp = orc.harris_response(src, nr1, 0.01)
    kp = orc.harris_nms(src, resp, 0.01, 1e-6)
    sn = orc.snap_to_cloud(src, orc.harris_refine(src, nr1, 0.01, src[kp].copy()), 1e-4)
    assert np.array_equal(kp_xyz("Harris3D_src_kp.bin"), src[sn[sn >= 0]])
    # Harris6D likewise, with the intensity gradient
    resp6, _, _ = orc.harris6d_response(src, rgbz["underwater_source"], nr1, 0.01)
    kp6 = orc.harris_nms(src, resp6, 0.01, 1e-6)
    sn6 = orc.snap_to_cloud(src, orc.harris_refine(src, nr1, 0.01, src[kp6].copy()), 1e-4)
    assert np.array_equal(kp_xyz("Harris6D_src_kp.bin"), src[sn6[sn6 >= 0]])
    assert int(rows[("Harris6D", "FPFH")][4]) == int((sn6 >= 0).sum()) > 20
    # FPFH at the Harris3D keypoints: bit for bit the oracle's rows, hence the same correspondences
    nr, _, _ = orc.normals(src, radius=0.03)
    f_or = orc.fpfh(src, nr, src[sn[sn >= 0]], radius=0.05)
    f_shim = np.fromfile(tmp_path / "Harris3D_FPFH_src.bin", dtype=np.float32).reshape(-1, 33)
    assert np.array_equal(f_shim.view(np.uint32), f_or.view(np.uint32))
    # reuse: 2 big clouds -> 2 uploads of them; everything else is keypoint clouds (spin-image normals, ICP)
    reuse = np.fromfile(tmp_path / "reuse.bin", dtype=np.uint64)
    up, reused, passes, p_reused, n_up, n_skip = (int(v) for v in reuse)
    print(r.stdout.strip().splitlines()[-1])
    assert reused >= 3 * 2 * 8          # 3 detectors x 2 clouds x (>= 8 descriptor types announce the cloud again)
    assert p_reused >= 3 * 2 * 4        # ... and at least 4 of them ask for its normals again
    assert passes <= 3 * 2 + 3 * 2      # one pass per big cloud and detector, one per keypoint cloud (spin images)


@pytest.mark.gpu
def test_group_api_from_one_cpp_process_with_a_thread_per_gpu():
    """the multi-GPU group API (pfx_group_join, pfx_slab_distribute, pfx_slab_owned_rows, pfx_group_allreduce) driven
    from C++ threads of ONE process, no launcher: as many ranks as the box has GPUs (a group of one on a single-GPU box
    still goes through NCCL and the whole distribution path); the gathered owned rows must equal the single-GPU rows"""
    import torch
    demo = os.path.join(ROOT, "pcl_feature_extraction_b200", "lib", "group_threads_demo")
    assert os.path.exists(demo), "run __graft_entry__.build()"
    n = min(torch.cuda.device_count(), 4)
    r = subprocess.run([demo, str(n), "192"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"n_gpus={n}" in r.stdout and "rows_bit_identical=1" in r.stdout
