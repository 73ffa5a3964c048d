"""PCD v0.7 reader/writer for the reference's bundled clouds (`FIELDS x y z rgb`, binary, 16 B/point).

Mirrors what pcl::io::loadPCDFile / savePCDFile do for the reference at evaluation.cpp:226,231,258
(SURVEY.md A.10): ASCII header up to the `DATA binary` line, then POINTS records; the bundled files
carry trailing zero padding after the point data, which is ignored.
"""
import numpy as np


def read_pcd(path):
    """Returns (xyz float32 [n,3], rgb uint32 [n] or None, header dict)."""
    with open(path, "rb") as f:
        raw = f.read()
    hdr = {}
    pos = 0
    while True:
        end = raw.index(b"\n", pos)
        line = raw[pos:end].decode("ascii", "replace").strip()
        pos = end + 1
        if not line or line.startswith("#"):
            continue
        key, _, val = line.partition(" ")
        hdr[key.upper()] = val
        if key.upper() == "DATA":
            break
    fields = hdr["FIELDS"].split()
    sizes = [int(s) for s in hdr["SIZE"].split()]
    types = hdr["TYPE"].split()
    counts = [int(c) for c in hdr.get("COUNT", " ".join(["1"] * len(fields))).split()]
    n = int(hdr["POINTS"])
    tmap = {("F", 4): "<f4", ("F", 8): "<f8", ("U", 4): "<u4", ("U", 1): "u1", ("U", 2): "<u2",
            ("I", 4): "<i4", ("I", 2): "<i2", ("I", 1): "i1"}
    if hdr["DATA"] == "binary":
        dt = np.dtype([(fld, tmap[(t, s)], (c,)) if c > 1 else (fld, tmap[(t, s)])
                       for fld, t, s, c in zip(fields, types, sizes, counts)])
        rec = np.frombuffer(raw, dtype=dt, count=n, offset=pos)
        xyz = np.stack([rec["x"], rec["y"], rec["z"]], axis=1).astype(np.float32)
        rgb = rec["rgb"].view(np.uint32).copy() if "rgb" in fields else None
    elif hdr["DATA"] == "ascii":
        arr = np.loadtxt(raw[pos:].decode().splitlines()[:n], dtype=np.float64).reshape(n, -1)
        xyz = arr[:, [fields.index(c) for c in "xyz"]].astype(np.float32)
        rgb = None
    else:
        raise ValueError("unsupported PCD DATA mode " + hdr["DATA"])
    return np.ascontiguousarray(xyz), rgb, hdr


def write_pcd(path, xyz, rgb=None):
    xyz = np.ascontiguousarray(xyz, np.float32)
    n = len(xyz)
    if rgb is None:
        rgb = np.zeros(n, np.uint32)
    rec = np.zeros(n, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("rgb", "<u4")])
    rec["x"], rec["y"], rec["z"], rec["rgb"] = xyz[:, 0], xyz[:, 1], xyz[:, 2], rgb
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z rgb\nSIZE 4 4 4 4\n"
           "TYPE F F F F\nCOUNT 1 1 1 1\nWIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA binary\n" % (n, n))
    with open(path, "wb") as f:
        f.write(hdr.encode("ascii"))
        f.write(rec.tobytes())
