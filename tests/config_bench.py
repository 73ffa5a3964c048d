"""Configs C1-C3 of BASELINE.json on the bundled clouds (tests/golden/clouds.npz): the GPU path through the C ABI
(host buffers in, host buffers out: what a user of the reference gets) beside the CPU oracle on all host threads.
One JSON line per config: per-stage milliseconds for both, sizes, and whether the index outputs agree.
Not the contract bench (bench.py); the numbers land in profiles/ and DESIGN.md."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pcl_feature_extraction_b200 as pfx
from oracle import binding as orc

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "clouds.npz"))
ctx = pfx.Context(0)
ctx.set_viewpoint(0, 0, 0)


class Timer:
    def __init__(self):
        self.ms = {}
    def run(self, name, fn, sync=None):
        t0 = time.perf_counter()
        r = fn()
        if sync:
            sync()
        self.ms[name] = self.ms.get(name, 0.0) + 1e3 * (time.perf_counter() - t0)
        return r


def c1(gpu):
    T = Timer()
    feats, kps = [], []
    for name in ("indoor_source", "indoor_target"):
        pts = Z[name]
        if gpu:
            ctx.set_surface(pts)
            xyz = T.run("voxel_grid", lambda: ctx.voxel_grid(0.01))
            T.run("set_surface", lambda: ctx.set_surface(xyz))
            nr = T.run("normals", lambda: ctx.normals(radius=0.03))
            res = T.run("resolution", ctx.cloud_resolution)
            kp, _ = T.run("iss", lambda: ctx.iss(6 * res, 4 * res))
            ctx.set_queries(xyz[kp])
            f = T.run("fpfh", lambda: ctx.fpfh(radius=0.05))
            ctx.set_queries(None)
        else:
            xyz = T.run("voxel_grid", lambda: orc.voxel_grid(pts, 0.01))
            nr, _, _ = T.run("normals", lambda: orc.normals(xyz, radius=0.03))
            res = T.run("resolution", lambda: orc.cloud_resolution(xyz))
            kp, _ = T.run("iss", lambda: orc.iss(xyz, 6 * res, 4 * res))
            f = T.run("fpfh", lambda: orc.fpfh(xyz, nr, q=xyz[kp], radius=0.05))
        feats.append(f); kps.append(kp)
    if gpu:
        c = T.run("match", lambda: ctx.match(feats[0], feats[1], reciprocal=True))
        corr = np.stack([c["index_query"], c["index_match"]], 1)
    else:
        q, m, _ = T.run("match", lambda: orc.match_reciprocal(feats[0], feats[1]))
        corr = np.stack([q, m], 1)
    return T.ms, {"keypoints": [len(k) for k in kps], "correspondences": len(corr)}, (kps, corr)


def c2(gpu, normal_r=0.03):
    T = Timer()
    feats, kps = [], []
    for name in ("underwater_source", "underwater_target"):
        pts = Z[name]
        if gpu:
            T.run("set_surface", lambda: ctx.set_surface(pts))
            h = T.run("harris3d", lambda: ctx.harris3d(0.01, 1e-6))
            snapped = h["snapped_idx"][h["snapped_idx"] >= 0]
            T.run("normals", lambda: ctx.normals(radius=normal_r, want_output=False))
            ctx.set_queries(pts[snapped])
            s, _ = T.run("shot352", lambda: ctx.shot352(0.05))
            ctx.set_queries(None)
        else:
            nr1, _, _ = T.run("harris3d", lambda: orc.normals(pts, radius=0.01))
            resp = T.run("harris3d", lambda: orc.harris_response(pts, nr1, 0.01))
            kp = T.run("harris3d", lambda: orc.harris_nms(pts, resp, 0.01, 1e-6))
            corners = T.run("harris3d", lambda: orc.harris_refine(pts, nr1, 0.01, pts[kp].copy()))
            sn = T.run("harris3d", lambda: orc.snap_to_cloud(pts, corners, 1e-4))
            snapped = sn[sn >= 0]
            nr, _, _ = T.run("normals", lambda: orc.normals(pts, radius=normal_r))
            s, _ = T.run("shot352", lambda: orc.shot352(pts, nr, pts[snapped], 0.05))
        ok = ~np.isnan(s[:, 0])
        feats.append(np.ascontiguousarray(s[ok])); kps.append(snapped)
    if gpu:
        c = T.run("match", lambda: ctx.match(feats[0], feats[1], reciprocal=True))
        n = len(c)
    else:
        q, m, _ = T.run("match", lambda: orc.match_reciprocal(feats[0], feats[1]))
        n = len(q)
    return T.ms, {"keypoints": [len(k) for k in kps], "descriptors": [len(f) for f in feats], "correspondences": n}, kps


def c3(gpu, planar):
    T = Timer()
    pts = Z["indoor_source"]
    res = float(np.deg2rad(0.5))
    if gpu:
        ctx.set_surface(pts)
        if planar:
            d = T.run("range_image", lambda: ctx.range_image_planar(640, 480, 320.0, 240.0, 525.0, 525.0), ctx.sync)
        else:
            d = T.run("range_image", lambda: ctx.range_image_spherical(res), ctx.sync)
        kp, _, _, _ = T.run("narf_keypoints", lambda: ctx.narf_keypoints(0.2))
        f = T.run("narf36", lambda: ctx.narf36(kp, 0.2, True))
        shape = (d.width, d.height)
    else:
        if planar:
            img, d = T.run("range_image", lambda: orc.range_image_planar(pts, 640, 480, 320, 240, 525, 525))
        else:
            img, d = T.run("range_image", lambda: orc.range_image_spherical(pts, res))
        kp, _, _ = T.run("narf_keypoints", lambda: orc.narf_keypoints(img, d, 0.2))
        f = T.run("narf36", lambda: orc.narf36(img, d, kp, 0.2, True))
        shape = (d.width, d.height)
    return T.ms, {"image": shape, "keypoints": len(kp), "descriptors": len(f)}, kp


def report(name, fn, *args, strict_too=False):
    fn(True, *args)  # warm-up (allocations, first-use costs)
    g_ms, g_info, g_idx = fn(True, *args)
    c_ms, c_info, c_idx = fn(False, *args)
    strict = None
    if strict_too:  # PFX_PARITY_STRICT: reference-order arithmetic, index outputs equal to the CPU path's
        ctx.set_parity_mode(True)
        try:
            fn(True, *args)
            s_ms, s_info, s_idx = fn(True, *args)
        finally:
            ctx.set_parity_mode(False)
        strict = {"gpu_ms": {k: round(v, 3) for k, v in s_ms.items()}, "gpu_total_ms": round(sum(s_ms.values()), 3), "gpu": s_info,
                  "index_outputs_equal_cpu": s_info == c_info}
    out = {"config": name, "gpu_strict": strict, "gpu_ms": {k: round(v, 3) for k, v in g_ms.items()}, "gpu_total_ms": round(sum(g_ms.values()), 3),
           "cpu_ms": {k: round(v, 1) for k, v in c_ms.items()}, "cpu_total_ms": round(sum(c_ms.values()), 1),
           "cpu_threads": orc.num_threads(), "gpu": g_info, "cpu": c_info,
           "note": "GPU: C ABI with host buffers (H2D/D2H included), wall clock after a warm-up pass; CPU: restated-PCL oracle, OpenMP"}
    print(json.dumps(out))


if __name__ == "__main__":
    report("C1 indoor pair: VoxelGrid 1cm + normals r=3cm + ISS + FPFH33 r=5cm + reciprocal matching", c1, strict_too=True)
    report("C2 underwater pair: Harris3D + SHOT352 r=5cm (normals r=3cm) + reciprocal matching", c2, strict_too=True)
    report("C3 indoor source: spherical range image 0.5 deg + NARF keypoints (support 0.2) + Narf36", c3, False)
    report("C3' indoor source: planar 640x480 f=525 range image (the reference's geometry) + NARF + Narf36", c3, True)
    ctx.close()
