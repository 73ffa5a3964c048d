"""Development diagnostic (GPU box): where do the GPU path and the CPU oracle part ways?

  python tools/parity_diag.py [spfh] [harris] [c1]

spfh   - SPFH rows from identical normals: share of bit-identical rows, and which sub-histogram carries the moved votes
harris - config C2's detector, stage by stage: normals, response, NMS, refine, snap
c1     - config C1: keypoints and correspondences, with the distance margins of the differing matches
One JSON line per section.  Uses the oracle: test infrastructure, never part of the product."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pcl_feature_extraction_b200 as pfx
from oracle import binding as orc
from pcl_feature_extraction_b200.synth import sheet_cloud

Z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "clouds.npz"))
ctx = pfx.Context(0)
ctx.set_viewpoint(0, 0, 0)
what = sys.argv[1:] or ["spfh", "harris", "c1"]


def spfh_diag():
    out = {}
    for name, xyz, kw, nkw in (("sheet_k32", sheet_cloud(side=256, pitch=0.004), dict(k=32), dict(k=32)),
                               ("underwater_r3cm", Z["underwater_source"][:30000], dict(radius=0.03), dict(radius=0.02))):
        xyz = np.ascontiguousarray(xyz)
        nr, _, _ = orc.normals(xyz, **nkw)
        ctx.set_surface(xyz)
        ctx.set_queries(None)
        ctx.set_surface_normals(nr)
        s = ctx.spfh(**kw)
        ref = orc.spfh(xyz, nr, np.arange(len(xyz), dtype=np.int32), **kw)
        d = np.abs(s - ref)
        rows = d.max(1) > 0
        blk = [(d[:, 11 * b:11 * b + 11].max(1) > 0).mean() for b in range(3)]
        f = ctx.fpfh(**kw)
        fr = orc.fpfh(xyz, nr, **kw)
        fd = np.nan_to_num(np.abs(f - fr), nan=0.0).max(1)
        out[name] = {"rows": len(xyz), "spfh_rows_differ": float(rows.mean()), "by_block_f1_f2_f3": [float(b) for b in blk],
                     "fpfh_rows_gt_1e-2": float((fd > 1e-2).mean()), "fpfh_rows_gt_1e-3": float((fd > 1e-3).mean()),
                     "fpfh_max": float(fd.max()), "fpfh_median": float(np.median(fd))}
    print(json.dumps({"spfh": out}))


def harris_diag():
    out = {}
    for name in ("underwater_source", "underwater_target"):
        pts = Z[name]
        ctx.set_surface(pts)
        ctx.set_queries(None)
        h = ctx.harris3d(0.01, 1e-6)
        nr1, _, gap = orc.normals(pts, radius=0.01)
        resp = orc.harris_response(pts, nr1, 0.01)
        kp = orc.harris_nms(pts, resp, 0.01, 1e-6)
        corners = orc.harris_refine(pts, nr1, 0.01, pts[kp].copy())
        sn = orc.snap_to_cloud(pts, corners, 1e-4)
        g_resp = h["response"]
        rec = {"n": len(pts), "gpu_kp": int(len(h["kp_idx"])), "cpu_kp": int(len(kp)),
               "kp_sym_diff": int(len(np.setxor1d(h["kp_idx"], kp))),
               "resp_bit_identical": float((g_resp.view(np.uint32) == resp.view(np.uint32)).mean()),
               "resp_max_abs_diff": float(np.abs(g_resp - resp).max()),
               "gpu_snapped": int((h["snapped_idx"] >= 0).sum()), "cpu_snapped": int((sn >= 0).sum())}
        if "normals" in h:
            rec["normals_bit_identical"] = float((h["normals"].view(np.uint32) == nr1.view(np.uint32)).all(1).mean())
        if len(h["kp_idx"]) == len(kp) and np.array_equal(h["kp_idx"], kp):
            rec["corner_bit_identical"] = float((h["kp_xyz"].view(np.uint32) == corners.view(np.uint32)).all(1).mean())
            rec["snapped_equal"] = bool(np.array_equal(h["snapped_idx"], sn))
        out[name] = rec
    print(json.dumps({"harris": out}))


def c1_diag():
    feats = {True: [], False: []}
    kps = {True: [], False: []}
    for name in ("indoor_source", "indoor_target"):
        pts = Z[name]
        ctx.set_surface(pts)
        xyz = ctx.voxel_grid(0.01)
        oxyz = orc.voxel_grid(pts, 0.01)
        same_vox = np.array_equal(xyz, oxyz)
        ctx.set_surface(xyz)
        ctx.set_queries(None)
        nr = ctx.normals(radius=0.03)
        onr, _, _ = orc.normals(xyz, radius=0.03)
        res = ctx.cloud_resolution()
        kp, _ = ctx.iss(6 * res, 4 * res)
        okp, _ = orc.iss(xyz, 6 * orc.cloud_resolution(xyz), 4 * orc.cloud_resolution(xyz))
        ctx.set_queries(xyz[kp])
        f = ctx.fpfh(radius=0.05)
        ctx.set_queries(None)
        of = orc.fpfh(xyz, onr, q=xyz[okp], radius=0.05)
        feats[True].append(f); feats[False].append(of)
        kps[True].append(kp); kps[False].append(okp)
        print(json.dumps({"c1_cloud": name, "voxel_equal": bool(same_vox), "kp_equal": bool(np.array_equal(kp, okp)),
                          "normals_bit_identical": float((nr.view(np.uint32) == onr.view(np.uint32)).all(1).mean()),
                          "normals_max_abs": float(np.nanmax(np.abs(nr - onr))),
                          "fpfh_bit_identical_rows": float((f.view(np.uint32) == of.view(np.uint32)).all(1).mean()) if f.shape == of.shape else None,
                          "fpfh_max_abs": float(np.abs(f - of).max()) if f.shape == of.shape else None}))
    c = ctx.match(feats[True][0], feats[True][1], reciprocal=True)
    q, m, dist = orc.match_reciprocal(feats[False][0], feats[False][1])
    g = set(zip(c["index_query"].tolist(), c["index_match"].tolist()))
    o = set(zip(q.tolist(), m.tolist()))
    diff = sorted(g ^ o)
    # margins of the differing pairs in the ORACLE's descriptors: best vs second-best distance, both directions
    a, b = feats[False]
    marg = []
    for (i, j) in diff[:20]:
        d_row = ((a[i][None] - b) ** 2).sum(1)
        d_col = ((a - b[j][None]) ** 2).sum(1)
        s_row, s_col = np.sort(d_row)[:2], np.sort(d_col)[:2]
        marg.append({"pair": [int(i), int(j)], "row_best2": [float(v) for v in s_row], "col_best2": [float(v) for v in s_col]})
    print(json.dumps({"c1": {"gpu_corr": len(g), "cpu_corr": len(o), "sym_diff": len(diff), "margins": marg}}))


if "spfh" in what:
    spfh_diag()
if "harris" in what:
    harris_diag()
if "c1" in what:
    c1_diag()
ctx.close()
