"""Generates the committed fixtures under tests/golden/ (run in the build container only; the GPU
box has no /root/reference).

clouds.npz      : xyz of the reference's four bundled PCD clouds (data/indoor, data/underwater),
                  needed because BASELINE configs C1-C3 are defined on them.
clouds_rgb.npz  : the packed colours of the underwater pair (Harris 6D works on intensities).
oracle_kat.npz  : outputs of the CPU oracle (oracle/) on a fixed crop; they pin the oracle against
                  silent drift.  They are NOT PCL outputs: PCL cannot be built here (parity unpinned).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding as orc  # noqa: E402
from pcl_feature_extraction_b200.pcd import read_pcd  # noqa: E402

REF = "/root/reference/data"
out = os.path.dirname(os.path.abspath(__file__))

clouds = {}
for name, rel in [("indoor_source", "indoor/source.pcd"), ("indoor_target", "indoor/target.pcd"),
                  ("underwater_source", "underwater/source.pcd"), ("underwater_target", "underwater/target.pcd")]:
    xyz, _, _ = read_pcd(os.path.join(REF, rel))
    clouds[name] = xyz
np.savez_compressed(os.path.join(out, "clouds.npz"), **clouds)

# packed 0x00RRGGBB colours of the underwater pair (the launch file's clouds): HarrisKeypoint6D needs intensities
colours = {}
for name, rel in [("underwater_source", "underwater/source.pcd"), ("underwater_target", "underwater/target.pcd")]:
    _, rgb, _ = read_pcd(os.path.join(REF, rel))
    colours[name] = (rgb & 0x00FFFFFF).astype(np.uint32)
np.savez_compressed(os.path.join(out, "clouds_rgb.npz"), **colours)

# fixed crop: a box of the indoor source cloud (keeps the real sampling pattern)
src = clouds["indoor_source"]
c = src[(src[:, 0] > -0.5) & (src[:, 0] < 0.1) & (src[:, 1] > -0.6) & (src[:, 1] < 0.0)]
crop = np.ascontiguousarray(c[:8000])
q = crop[::40].copy()
kat = {"crop": crop, "q": q}
kat["knn_idx"], kat["knn_d2"] = orc.knn(crop, q, 16)
kat["rad_off"], kat["rad_idx"], kat["rad_d2"] = orc.radius_search(crop, q, 0.02)
nr, cnt, gap = orc.normals(crop, radius=0.03)
kat["normals"], kat["normals_cnt"], kat["normals_gap"] = nr, cnt, gap
kat["resolution"] = np.array([orc.cloud_resolution(crop)])
res = float(kat["resolution"][0])
kat["iss_kp"], kat["iss_sal"] = orc.iss(crop, 6 * res, 4 * res)
kat["fpfh"] = orc.fpfh(crop, nr, q, radius=0.05)
kat["shot"], kat["shot_rf"] = orc.shot352(crop, nr, q, 0.05)
nr1, _, _ = orc.normals(crop, radius=0.01)
kat["harris_resp"] = orc.harris_response(crop, nr1, 0.01)
kat["harris_kp"] = orc.harris_nms(crop, kat["harris_resp"], 0.01, 1e-6)
rng = np.random.default_rng(7)
a = kat["fpfh"][:120]
b = kat["fpfh"][60:] + rng.normal(0, 0.5, kat["fpfh"][60:].shape).astype(np.float32)
kat["match_a"], kat["match_b"] = a, b
kat["match_q"], kat["match_m"], kat["match_d"] = orc.match_reciprocal(a, b)
np.savez_compressed(os.path.join(out, "oracle_kat.npz"), **kat)
print({k: v.shape for k, v in kat.items()})
