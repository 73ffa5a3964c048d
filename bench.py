#!/usr/bin/env python
"""bench.py — headline benchmark of the feature-extraction hot path (BASELINE.json metric:
"FPFH+SHOT descriptors/sec, 1M-pt cloud").

A step = one pass of the hot path over one synthetic 1M-point cloud (BASELINE config C4 widened with
C5's SHOT stage): voxel-hash build -> kNN(32) -> normals -> SPFH -> FPFH33 -> SHOT LRF
-> SHOT352 (neighbourhoods from the resident k-search rows), i.e. 2 descriptors (one FPFH33 row + one SHOT352 row) per point.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--side S]

N > 1 is launched by torchrun (one rank per GPU, NCCL); clouds are sharded over ranks with no
data-path collective (weak scaling).  `value` times the C-ABI with inputs resident in HBM and
outputs left in HBM; `e2e` times the same calls with HOST buffers (H2D of the cloud and D2H of every
descriptor inside the timed region); `e2e_resident` describes a pair of host clouds, leaves the rows in HBM,
matches them on the device and copies only the correspondences back (BASELINE config 5's shape).  `roofline` is measured live with CUDA events around the
dominant kernel; `cpu_baseline` is the CPU oracle (a restatement of PCL: kind "port") on a bounded
sample.  `--impl reference` times that CPU path alone (the reference itself - ROS + PCL - cannot be
built here: DESIGN.md §3).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K_NN = 32
PITCH = 0.004
SHOT_RADIUS = 3.2 * PITCH
METRIC = "FPFH+SHOT descriptors/sec, 1M-pt cloud"
UNIT = "descriptors/s"

# algorithmic bytes per point of each stage (BASELINE.md §4 / SURVEY.md §8d); nb = mean neighbours
ALG_BYTES = {
    "knn_kernel": lambda nb: 16 + 8 * nb,
    "normals_kernel": lambda nb: 32 + 4 * nb,
    "spfh_kernel": lambda nb: 164 + 4 * nb,
    "fpfh_kernel": lambda nb: 264 + 8 * nb,
    "lrf_kernel": lambda nb: 52 + 8 * nb,
    "shot_kernel": lambda nb: 1512 + 8 * nb,
    # fused frame + descriptor kernel: S5 + S6 of SURVEY.md §8d
    "shot_fused_kernel": lambda nb: (52 + 8 * nb) + (1512 + 8 * nb),
    "knn_tile_kernel": lambda nb: (16 + 8 * nb) + (32 + 4 * nb),  # kNN sets + fused normals (S1 + S2)
    "fpfh_list_kernel": lambda nb: 264 + 8 * nb,
    "fpfh_list32_kernel": lambda nb: 264 + 8 * nb,
    "spfh_list32_kernel": lambda nb: 164 + 4 * nb,
}


def workload_name(side):
    return (f"synthetic {side}x{side}-point height-field sheet (pitch 4 mm, seed 20240601+rank, shuffled): "
            f"dense normals k=32 + FPFH33 k=32 + SHOT352 r=12.8 mm")


def make_config(side):
    """the workload both arms (--impl ours / reference) run, word for word the same"""
    return {"workload": workload_name(side), "points_per_cloud": side * side, "clouds_per_step_per_gpu": 1,
            "descriptors_per_point": 2,
            "l2": "working set per step ~1.9 GB (1.5 GB SHOT output) >> 126 MB L2; input alternates between 2 clouds"}


def whole_step_bytes(nbar):
    """hash build 38 + k-search list (16 + 8k) + normals (32 + 4k) + SPFH (164 + 4k) + FPFH
    (264 + 8k) at k = 32, + radius list (24 + 8n) + SHOT frame (52 + 8n) + SHOT352 (1512 + 8n) at the MEASURED mean
    neighbour count n of the radius stages (3 638 B/pt at n = 32)"""
    return 1282.0 + 1588.0 + 24.0 * nbar


def ncu_traffic(kernel_key):
    """DRAM bytes per launch of a kernel from the committed ncu --set full capture (profiles/), or None"""
    try:
        for name in ("r02_traffic.json", "r01_traffic.json"):
            path = os.path.join(ROOT, "profiles", name)
            if os.path.exists(path):
                with open(path) as f:
                    return json.load(f).get(kernel_key)
        return None
    except Exception:
        return None


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_pipeline(orc, pts):
    nr, _, _ = orc.normals(pts, k=K_NN)
    f = orc.fpfh(pts, nr, k=K_NN)
    s, rf = orc.shot352(pts, nr, None, SHOT_RADIUS)
    return f, s


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, rank, world):
    """--impl reference: the CPU path (oracle port of PCL) on ALL host threads, the full cloud per step.
    torchrun exports OMP_NUM_THREADS=1 to its workers: the thread count is set explicitly."""
    if rank != 0:
        return
    from oracle import binding as orc
    from pcl_feature_extraction_b200.synth import sheet_cloud
    cores = host_threads()
    orc.set_num_threads(cores)
    side = args.ref_side or args.side
    pts = sheet_cloud(side=side, pitch=PITCH, seed=20240601)
    t0 = time.perf_counter()
    cpu_pipeline(orc, pts)  # warm-up (one pass is enough for a CPU code; it also sizes the timed leg)
    t1 = time.perf_counter() - t0
    steps = max(1, min(args.steps, int(150.0 / max(t1, 1e-3))))  # the run ends within a few minutes on any box
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_pipeline(orc, pts)
    dt = time.perf_counter() - t0
    val = 2.0 * len(pts) * steps / dt
    sample = (f"the full {side}x{side}-point cloud per step ({len(pts)} points), {steps} timed steps after one warm-up pass"
              if side == args.side else f"{side}x{side}-point sheet per step ({len(pts)} points), same stages and parameters")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "steps_timed": steps, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": make_config(args.side),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": f"restated-PCL CPU oracle with OpenMP on {cores} host threads (set explicitly); PCL/ROS cannot be built in this "
                "image; under torchrun rank 0 alone runs the CPU arm",
    }), flush=True)


def matching_record(pfx, ctx, torch, dev, peaks, rows=65536, dim=352, reps=5):
    """descriptor matching, the one dense contraction (features.h:224-273): rows x rows x dim exact 1-NN through the
    tcgen05 engine (candidate GEMM + fp32 rescore + exact redo), device-resident; TFLOP/s = 2 * na * nb * dim / time"""
    g = torch.Generator(device=dev).manual_seed(rows + dim)
    a = torch.rand((rows, dim), device=dev, generator=g)
    b = torch.rand((rows, dim), device=dev, generator=g)
    m = rows // 2
    b[:m] = a[:m] + 0.01 * torch.randn((m, dim), device=dev, generator=g)
    idx = torch.empty(rows, dtype=torch.int32, device=dev)
    d2 = torch.empty(rows, dtype=torch.float32, device=dev)
    ctx.set_match_engine(1)
    try:
        for _ in range(2):
            ctx.match_nn_dev(a.data_ptr(), rows, b.data_ptr(), rows, dim, idx.data_ptr(), d2.data_ptr())
        torch.cuda.synchronize()
        i0 = ctx.match_info()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            ctx.match_nn_dev(a.data_ptr(), rows, b.data_ptr(), rows, dim, idx.data_ptr(), d2.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        i1 = ctx.match_info()
        ms = e0.elapsed_time(e1) / reps
        ctx.profile_begin("tc_candidates")
        ctx.match_nn_dev(a.data_ptr(), rows, b.data_ptr(), rows, dim, idx.data_ptr(), d2.data_ptr())
        prof = ctx.profile_end()
        cand_ms = sum(t for _, t in prof.values())
    finally:
        ctx.set_match_engine(-1)
    tf = 2.0 * rows * rows * dim / (ms * 1e-3) / 1e12
    planted_found = float((idx[:m].long() == torch.arange(m, device=dev)).float().mean().item())
    return {"workload": f"{rows} x {rows} x {dim} exact 1-NN (random rows, half of the targets planted near a query)",
            "ms": ms, "tflops": tf, "frac_of_bf16_burst": tf / peaks["bf16_tflops"], "bf16_tflops_peak": peaks["bf16_tflops"],
            "candidates_kernel_ms": cand_ms, "candidates_kernel_tflops": 2.0 * rows * rows * dim / (cand_ms * 1e-3) / 1e12 if cand_ms else None,
            "redo_rate": (i1["redone_exact"] - i0["redone_exact"]) / max(1, i1["rows"] - i0["rows"]),
            "planted_matches_found": planted_found, "bound": "tensor"}


def bundled_record(pfx, ctx):
    """configs C1, C2, C3 of BASELINE.json on the bundled clouds (the copies under tests/golden/, taken from the
    reference's data/ directory), GPU only, host buffers in and out (what a user of the reference gets), wall clock
    of the second pass.  C1 and C2 run in PFX_PARITY_STRICT (index-for-index equal to the CPU path:
    tests/test_gpu_end_to_end.py) and, for comparison, with the fast kernels."""
    path = os.path.join(ROOT, "tests", "golden", "clouds.npz")
    if not os.path.exists(path):
        return None
    with np.load(path) as npz:  # (an NpzFile decompresses an array on every access: read them once)
        Z = {k: np.ascontiguousarray(npz[k]) for k in npz.files}

    def c1():
        feats, n_in, n_kp = [], 0, []
        for name in ("indoor_source", "indoor_target"):
            pts = Z[name]
            n_in += len(pts)
            ctx.set_surface(pts)
            xyz = ctx.voxel_grid(0.01)
            ctx.set_surface(xyz)
            ctx.normals(radius=0.03, want_output=False)
            res = ctx.cloud_resolution()
            kp, _ = ctx.iss(6 * res, 4 * res)
            ctx.set_queries(xyz[kp])
            feats.append(ctx.fpfh(radius=0.05))
            ctx.set_queries(None)
            n_kp.append(int(len(kp)))
        c = ctx.match(feats[0], feats[1], reciprocal=True)
        return n_in, {"keypoints": n_kp, "correspondences": int(len(c))}

    def c2():
        feats, n_in, n_kp = [], 0, []
        for name in ("underwater_source", "underwater_target"):
            pts = Z[name]
            n_in += len(pts)
            ctx.set_surface(pts)
            h = ctx.harris3d(0.01, 1e-6)
            snapped = h["snapped_idx"][h["snapped_idx"] >= 0]
            ctx.normals(radius=0.03, want_output=False)
            ctx.set_queries(pts[snapped])
            sdesc, _ = ctx.shot352(0.05)
            ctx.set_queries(None)
            feats.append(np.ascontiguousarray(sdesc[~np.isnan(sdesc[:, 0])]))
            n_kp.append(int(len(snapped)))
        c = ctx.match(feats[0], feats[1], reciprocal=True)
        return n_in, {"keypoints": n_kp, "correspondences": int(len(c))}

    def c3():
        pts = Z["indoor_source"]
        ctx.set_surface(pts)
        ctx.range_image_spherical(float(np.deg2rad(0.5)))
        kp, _, _, _ = ctx.narf_keypoints(0.2)
        f = ctx.narf36(kp, 0.2, True)
        return len(pts), {"keypoints": int(len(kp)), "descriptors": int(len(f))}

    out = {}
    for name, fn, modes in (("c1_indoor_iss_fpfh_match", c1, (True, False)), ("c2_underwater_harris_shot_match", c2, (True, False)),
                            ("c3_indoor_narf_narf36", c3, (False,))):
        rec = {}
        for strict in modes:
            ctx.set_parity_mode(strict)
            try:
                fn()  # warm-up: allocations, first-use costs
                t0 = time.perf_counter()
                n_in, info = fn()
                dt = time.perf_counter() - t0
            finally:
                ctx.set_parity_mode(False)
            key = "strict" if strict else "fast"
            rec[key] = {"ms": 1e3 * dt, "points_per_s": n_in / dt, **info}
        out[name] = rec
    return out


def bind_to_gpu_numa(local):
    """Pin this rank's host threads to the CPUs NVML reports as local to its GPU, so that the page-locked
    staging buffers (first touch) and the copy threads sit on the GPU's NUMA node.  Returns the previous CPU set
    (restored for the CPU baseline leg), or None when nothing was changed."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {i for i in range(ncpu) if (mask[i // 64] >> (i % 64)) & 1}
        before = os.sched_getaffinity(0)
        cpus &= before
        if cpus and cpus != before:
            os.sched_setaffinity(0, cpus)
            return before
    except Exception:
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--side", type=int, default=1024, help="sheet is side x side points (1024 -> 2^20)")
    ap.add_argument("--ref-side", type=int, default=0, help="reference arm: sheet side (0 = the full --side cloud)")
    ap.add_argument("--cpu-side", type=int, default=448)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the matching / bundled-cloud / slab sub-records")
    ap.add_argument("--in-flight", type=int, default=2, help="clouds in flight per GPU (contexts on separate streams)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # stdout carries exactly ONE JSON line: anything a library prints to fd 1 on the way (NCCL's version banner when a
    # communicator is created) is sent to stderr; the result line goes to the saved descriptor
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    import pcl_feature_extraction_b200 as pfx
    from pcl_feature_extraction_b200.synth import sheet_cloud

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cpus_before = bind_to_gpu_numa(local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own log lines (version banner, NCCL_DEBUG output) go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- synthetic input: two clouds per rank, alternated between steps
    n = args.side * args.side
    hosts = []
    for c in range(2):
        p = sheet_cloud(side=args.side, pitch=PITCH, seed=20240601 + 1000 * c + rank)
        p4 = torch.zeros((n, 4), dtype=torch.float32).pin_memory()
        p4[:, :3] = torch.from_numpy(p)
        hosts.append(p4)
    devs = [h.to(dev) for h in hosts]
    d_fpfh = torch.empty((n, 33), dtype=torch.float32, device=dev)
    d_shot = torch.empty((n, 361), dtype=torch.float32, device=dev)

    ctx = pfx.Context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_viewpoint(0.0, 0.0, 0.0)

    def step_device(i):
        ctx.set_surface_dev(devs[i & 1].data_ptr(), n, 16)
        # (no radius index: dense SHOT takes its neighbourhoods from the k-search rows that normals / FPFH left resident)
        ctx.normals_dev(0.0, K_NN, None)
        ctx.fpfh_dev(0.0, K_NN, d_fpfh.data_ptr())
        ctx.shot352_dev(SHOT_RADIUS, d_shot.data_ptr())

    # ---- warm-up, with every kernel timed once to find the dominant one
    for i in range(max(args.warmup - 1, 0)):
        step_device(i)
    torch.cuda.synchronize()
    ctx.profile_begin(None)
    step_device(0)
    prof = ctx.profile_end()
    step_ms_prof = sum(ms for _, ms in prof.values())
    dom = max(prof.items(), key=lambda kv: kv[1][1])[0]
    dom_key = next((k for k in ALG_BYTES if k in dom), None)
    if dom_key is None:  # a build kernel dominates: report the heaviest of the modelled stages instead
        dom_key = max(ALG_BYTES, key=lambda k: sum(ms for nm, (_, ms) in prof.items() if k in nm))
    shares = {nm: round(ms / step_ms_prof, 4) for nm, (_, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1])[:8]}

    # mean neighbour count of the radius stages (for their algorithmic bytes), outside the timed region
    cnt = torch.empty(n, dtype=torch.int32, device=dev)
    ctx._chk(ctx.lib.pfx_radius_count(ctx.h, SHOT_RADIUS, pfx.capi._ptr(cnt), None, pfx.capi.DEVICE))
    torch.cuda.synchronize()
    nbar_radius = float(cnt.float().mean().item())
    del cnt
    nbar = nbar_radius if dom_key in ("lrf_kernel", "shot_kernel", "shot_fused_kernel") else float(K_NN)

    # ---- second context on its own stream: two clouds in flight per GPU.  The index builds and the tails of one
    # cloud's kernels run in the gaps of the other's (contexts own all of their state; results are bit-identical
    # to one context, tests/test_gpu_degenerate.py::test_two_contexts_side_by_side)
    IN_FLIGHT = max(1, args.in_flight)
    side_streams = [torch.cuda.Stream(device=dev) for _ in range(IN_FLIGHT - 1)]
    ctxs, outs = [ctx], [(d_fpfh, d_shot)]
    for st in side_streams:
        c2 = pfx.Context(local)
        c2.set_stream(st.cuda_stream)
        c2.set_viewpoint(0.0, 0.0, 0.0)
        ctxs.append(c2)
        outs.append((torch.empty((n, 33), dtype=torch.float32, device=dev), torch.empty((n, 361), dtype=torch.float32, device=dev)))

    def step_on(c, bufs, i):
        c.set_surface_dev(devs[i & 1].data_ptr(), n, 16)
        c.normals_dev(0.0, K_NN, None)
        c.fpfh_dev(0.0, K_NN, bufs[0].data_ptr())
        c.shot352_dev(SHOT_RADIUS, bufs[1].data_ptr())

    for i in range(2 * IN_FLIGHT):  # warm the extra contexts (buffers, grids, shared-memory attributes)
        step_on(ctxs[i % IN_FLIGHT], outs[i % IN_FLIGHT], i)
    torch.cuda.synchronize()

    # ---- timed region 1 (one cloud at a time): step latency and the dominant kernel's launch duration, taken with
    # CUDA events on the stream the kernel is launched on
    barrier()
    ctx.profile_begin(dom_key)
    l0 = torch.cuda.Event(enable_timing=True)
    l1 = torch.cuda.Event(enable_timing=True)
    l0.record()
    for i in range(args.steps):
        step_device(i)
    l1.record()
    barrier()
    dom_prof = ctx.profile_end()
    latency_ms = l0.elapsed_time(l1) / args.steps

    # ---- timed region 2 (the reported one): inputs resident in HBM, outputs left in HBM, IN_FLIGHT clouds in flight
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = sum(c.launches for c in ctxs)
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream()
    e0.record(cur)
    for st in side_streams:
        st.wait_event(e0)
    for i in range(args.steps):
        step_on(ctxs[i % IN_FLIGHT], outs[i % IN_FLIGHT], i)
    for st in side_streams:
        ev = torch.cuda.Event()
        ev.record(st)
        cur.wait_event(ev)
    e1.record(cur)
    barrier()
    launches = sum(c.launches for c in ctxs) - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1), latency_ms], dtype=torch.float64, device=dev)
    lt = torch.tensor([float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
    total_ms = float(ms[0].item())
    latency_ms = float(ms[1].item())
    value = 2.0 * n * world * args.steps / (total_ms * 1e-3)
    for c2 in ctxs[1:]:
        c2.close()
    del outs[1:]
    torch.cuda.empty_cache()

    # ---- the "pre-sorted input" variant of SURVEY section 8(d) C4: the same cloud in generation (row-major grid) order
    # instead of shuffled; one cloud at a time, outside the reported timed region
    presorted = None
    if rank == 0 and not args.no_extras:
        try:
            ps = sheet_cloud(side=args.side, pitch=PITCH, seed=20240601 + rank, shuffle=False)
            ps4 = torch.zeros((n, 4), dtype=torch.float32)
            ps4[:, :3] = torch.from_numpy(ps)
            d_ps = ps4.to(dev)

            def step_sorted():
                ctx.set_surface_dev(d_ps.data_ptr(), n, 16)
                ctx.normals_dev(0.0, K_NN, None)
                ctx.fpfh_dev(0.0, K_NN, d_fpfh.data_ptr())
                ctx.shot352_dev(SHOT_RADIUS, d_shot.data_ptr())

            for _ in range(3):
                step_sorted()
            torch.cuda.synchronize()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record()
            for _ in range(args.steps):
                step_sorted()
            p1.record()
            torch.cuda.synchronize()
            ps_ms = p0.elapsed_time(p1) / args.steps
            presorted = {"ms_per_step_single_cloud": ps_ms, "descriptors_per_s": 2.0 * n / (ps_ms * 1e-3),
                         "shuffled_ms_per_step_single_cloud": latency_ms,
                         "note": "the same synthetic sheet in generation order (no shuffle): the voxel hash sorts the points "
                                 "either way, so only the ingest / key kernels see the difference"}
            del d_ps
        except Exception as e:
            presorted = {"error": str(e)}

    # ---- sub-records of the one JSON line: matching GEMM and the bundled clouds (rank 0), outside the timed regions
    matching = bundled = slab = None
    if rank == 0 and not args.no_extras:
        peaks_x, _ = measured_peaks()
        try:
            matching = matching_record(pfx, ctx, torch, dev, peaks_x)
        except Exception as e:  # a sub-record must never take the headline down
            matching = {"error": str(e)}
        try:
            bundled = bundled_record(pfx, ctx)
        except Exception as e:
            bundled = {"error": str(e)}
        ctx.set_queries(None)
        torch.cuda.empty_cache()
    barrier()
    if world > 1 and not args.no_extras:
        # ---- one cloud slab-sharded over the ranks and both-sides-sharded matching, through the C ABI's group API
        # (NCCL inside the library; torch.distributed only carries the 128-byte group id).  Collective: every rank.
        try:
            from tools.slab_bench import join_group, slab_record
            from tools.ring_match_bench import ring_record
            join_group(ctx, dist, rank, world)
            slab = slab_record(pfx, ctx, torch, dist, dev, rank, world, side=args.side, steps=3, warmup=1, verify=True)
            slab["ring_matching"] = ring_record(pfx, ctx, torch, dist, dev, rank, world, rows=32768, dim=352, reps=2,
                                                check_rows=32768 * 8)
            ctx.group_leave()
        except Exception as e:
            slab = {"error": str(e)}
        ctx.set_queries(None)
        torch.cuda.empty_cache()
        barrier()

    # ---- e2e: the same C-ABI calls with HOST buffers (pinned), H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        # page-locked result buffers: the rows of step i are still being copied out (PFX_HOST_ASYNC, the context's
        # copy stream) while step i + 1 uploads its cloud and computes.  One set is enough: the copies of
        # consecutive steps are serialised on the copy stream (a consumer would drain step i before i + 1 lands).
        h_fpfh = [torch.empty((n, 33), dtype=torch.float32).pin_memory()] * 2
        h_shot = [torch.empty((n, 361), dtype=torch.float32).pin_memory()] * 2
        HOST, ASYNC = pfx.capi.HOST, pfx.capi.HOST_ASYNC

        def step_host(i):
            h = hosts[i & 1]
            ctx._chk(ctx.lib.pfx_set_surface(ctx.h, pfx.capi._ptr(h), n, 16, HOST))
            ctx._chk(ctx.lib.pfx_normals(ctx.h, 0.0, K_NN, None, 16, 3, HOST))
            ctx._chk(ctx.lib.pfx_fpfh(ctx.h, 0.0, K_NN, pfx.capi._ptr(h_fpfh[i & 1]), 132, ASYNC))
            ctx._chk(ctx.lib.pfx_shot352(ctx.h, SHOT_RADIUS, None, pfx.capi._ptr(h_shot[i & 1]), 1444, ASYNC))

        k2 = max(3, min(args.steps, 10))
        if os.environ.get("PFX_BENCH_DEBUG"):  # development aid: host time of every call of a step, synchronised
            for i in range(6):
                h, tt = hosts[i & 1], [time.perf_counter()]
                for call in (lambda: ctx.lib.pfx_set_surface(ctx.h, pfx.capi._ptr(h), n, 16, HOST),
                             lambda: ctx.lib.pfx_normals(ctx.h, 0.0, K_NN, None, 16, 3, HOST),
                             lambda: ctx.lib.pfx_fpfh(ctx.h, 0.0, K_NN, pfx.capi._ptr(h_fpfh[i & 1]), 132, ASYNC),
                             lambda: ctx.lib.pfx_shot352(ctx.h, SHOT_RADIUS, None, pfx.capi._ptr(h_shot[i & 1]), 1444, ASYNC)):
                    ctx._chk(call())
                    tt.append(time.perf_counter())
                    if os.environ["PFX_BENCH_DEBUG"] == "1":
                        ctx._chk(ctx.lib.pfx_sync(ctx.h))
                    tt.append(time.perf_counter())
                print("e2e debug step", i, [round(1e3 * (tt[j + 1] - tt[j]), 2) for j in range(8)], file=sys.stderr)
            ctx.profile_begin(None)
            step_host(0)
            for nm, (c, ms_k) in sorted(ctx.profile_end().items(), key=lambda kv: -kv[1][1])[:8]:
                print(f"e2e debug   {nm:45s} x{c:2d} {ms_k:8.3f} ms", file=sys.stderr)
        for i in range(4):  # warm-up: staging slots, and one pass over every cached voxel hash of the context
            step_host(i)
        ctx._chk(ctx.lib.pfx_sync(ctx.h))
        barrier()
        t0 = time.perf_counter()
        for i in range(k2):
            step_host(i)
        ctx._chk(ctx.lib.pfx_sync(ctx.h))  # every row of every step has landed in host memory
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        barrier()
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": 2.0 * n * world * k2 / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": n * 16,
               "d2h_bytes_per_step": n * (132 + 1444), "steps": k2,
               "note": "C-ABI with pinned HOST buffers: cloud uploaded and every FPFH33 and SHOT352 row copied back inside the timed "
                       "region (asynchronous delivery on the copy stream, pfx_sync at the end; PCIe-bound: 1.65 GB per step)"}

    # ---- e2e with the descriptors left in HBM (BASELINE config 5's shape: describe two clouds, match them): host
    # clouds in, only the 1-NN indices / distances of the matched rows come back - what a caller pays when the rows
    # feed the matcher instead of host code (the plain e2e above is bound by copying 1.5 GB of SHOT rows per cloud)
    e2e_resident = None
    if rank == 0 and not args.no_e2e and not args.no_extras:
        try:
            STEP = 64
            m = n // STEP
            d_f = [d_fpfh, torch.empty((n, 33), dtype=torch.float32, device=dev)]
            d_s = [d_shot, torch.empty((n, 361), dtype=torch.float32, device=dev)]
            nn_i = torch.empty(m, dtype=torch.int32, device=dev)
            nn_d = torch.empty(m, dtype=torch.float32, device=dev)
            h_i = [torch.empty(m, dtype=torch.int32).pin_memory() for _ in range(2)]
            h_d = [torch.empty(m, dtype=torch.float32).pin_memory() for _ in range(2)]
            HOSTM = pfx.capi.HOST

            def pair_step():
                for c in range(2):
                    ctx._chk(ctx.lib.pfx_set_surface(ctx.h, pfx.capi._ptr(hosts[c]), n, 16, HOSTM))
                    ctx.normals_dev(0.0, K_NN, None)
                    ctx.fpfh_dev(0.0, K_NN, d_f[c].data_ptr())
                    ctx.shot352_dev(SHOT_RADIUS, d_s[c].data_ptr())
                for t, (rows, dim, stride) in enumerate(((d_f, 33, 132), (d_s, 352, 1444))):
                    # every 64th descriptor of cloud 0 against every 64th of cloud 1, read in place
                    ctx.match_nn_dev(rows[0].data_ptr(), m, rows[1].data_ptr(), m, dim, nn_i.data_ptr(), nn_d.data_ptr(),
                                     stride_a=STEP * stride, stride_b=STEP * stride)
                    h_i[t].copy_(nn_i, non_blocking=True)
                    h_d[t].copy_(nn_d, non_blocking=True)
                torch.cuda.synchronize()

            for _ in range(3):
                pair_step()
            reps = 5
            t0 = time.perf_counter()
            for _ in range(reps):
                pair_step()
            dtp = (time.perf_counter() - t0) / reps
            e2e_resident = {"value": 4.0 * n / dtp, "unit": UNIT, "ms_per_cloud_pair": 1e3 * dtp,
                            "h2d_bytes_per_pair": 2 * n * 16, "d2h_bytes_per_pair": 2 * m * 8,
                            "workload": f"two {n}-point clouds from pinned host memory: dense normals + FPFH33 + SHOT352 each (rows left in HBM), "
                                        f"then exact 1-NN of every {STEP}th FPFH33 and SHOT352 row of cloud 0 among those of cloud 1 "
                                        f"({m} x {m}); indices and distances copied to the host; wall clock, synchronised per pair"}
            del d_f, d_s
        except Exception as e:
            e2e_resident = {"error": str(e)}
        ctx.set_queries(None)
        torch.cuda.empty_cache()
    barrier()  # (the other ranks wait here for rank 0's extra record instead of tearing the process group down)

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        # the dominant kernel may run as more than one launch per cloud (shot_fused_kernel: the pass over the k-search
        # rows + the pass over the rows it could not close); together they produce the n rows the bytes are counted for
        cnt_dom = sum(c for nm, (c, _) in dom_prof.items())
        ms_dom = sum(m for nm, (_, m) in dom_prof.items())
        launches_per_cloud = max(1, round(cnt_dom / max(args.steps, 1)))
        per_launch_s = (ms_dom / max(args.steps, 1)) * 1e-3
        alg = ALG_BYTES[dom_key](nbar) * n
        wsb = whole_step_bytes(nbar_radius)
        achieved = alg / per_launch_s / 1e9 if per_launch_s > 0 else 0.0
        roofline = {"bound": "hbm", "kernel": dom_key, "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": achieved / peaks["hbm_gbs"], "peak_kind": peak_kind + " (burst copy)", "traffic": ncu_traffic(dom_key) if n == (1 << 20) else None,
                    "alg_bytes_per_point": ALG_BYTES[dom_key](nbar), "points_per_launch": n, "mean_neighbours": nbar,
                    "launch_ms": per_launch_s * 1e3, "launches_per_cloud": launches_per_cloud,
                    "timed": "CUDA events on the launching stream over the single-cloud timed region (kernel alone on the GPU); "
                             "launch_ms = all launches of this kernel for one cloud",
                    "kernel_shares_of_step": shares,
                    "whole_step": {"alg_bytes_per_point": wsb, "mean_neighbours_radius_stages": nbar_radius,
                                   "achieved_GBps": wsb * n * args.steps / (total_ms * 1e-3) / 1e9,
                                   "frac": wsb * n * args.steps / (total_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                   "timed": f"the reported region ({IN_FLIGHT} clouds in flight)",
                                   "single_cloud": {"ms_per_step": latency_ms, "achieved_GBps": wsb * n / (latency_ms * 1e-3) / 1e9,
                                                    "frac": wsb * n / (latency_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}}}
        cpu = None
        if world == 1 and not args.no_cpu:
            if cpus_before:
                os.sched_setaffinity(0, cpus_before)  # the CPU baseline gets every host core
            from oracle import binding as orc
            cores = host_threads()
            orc.set_num_threads(cores)
            cs = args.cpu_side
            cp = sheet_cloud(side=cs, pitch=PITCH, seed=20240601)
            t0 = time.perf_counter()
            cpu_pipeline(orc, cp)
            dtc = time.perf_counter() - t0
            # the reference's own thread settings: NormalEstimationOMP and SHOTEstimationOMP on
            # all cores (tools.h:26, evaluation.cpp:770), FPFHEstimation single-threaded (evaluation.cpp:597)
            fs = min(cs, 224)
            fp = sheet_cloud(side=fs, pitch=PITCH, seed=20240601)
            t0 = time.perf_counter()
            nr_f, _, _ = orc.normals(fp, k=K_NN)
            t_n = time.perf_counter() - t0
            orc.set_num_threads(1)
            t0 = time.perf_counter()
            orc.fpfh(fp, nr_f, k=K_NN)
            t_f = time.perf_counter() - t0
            orc.set_num_threads(cores)
            t0 = time.perf_counter()
            orc.shot352(fp, nr_f, None, SHOT_RADIUS)
            t_s = time.perf_counter() - t0
            cpu = {"value": 2.0 * len(cp) / dtc, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{cs}x{cs}-point sheet ({len(cp)} points), same stages and parameters, one pass, "
                             f"{dtc:.1f} s; restated-PCL oracle with OpenMP on all host threads (PCL itself cannot be built here)",
                   "reference_faithful": {
                       "value": 2.0 * len(fp) / (t_n + t_f + t_s), "unit": UNIT, "cores": cores,
                       "threads": {"normals": cores, "fpfh": 1, "shot": cores},
                       "stage_s": {"normals": round(t_n, 3), "fpfh": round(t_f, 3), "shot": round(t_s, 3)},
                       "sample": f"{fs}x{fs}-point sheet ({len(fp)} points); thread settings of the reference: NormalEstimationOMP / "
                                 "SHOTEstimationOMP all cores, FPFHEstimation 1 thread (evaluation.cpp:597)"}}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": make_config(args.side),
            "run": {"clouds_in_flight_per_gpu": IN_FLIGHT,
                    "parallelism": f"cloud-sharded x{world}, no data-path collective; {IN_FLIGHT} clouds in flight per GPU "
                                   "(one context and stream each)"},
            "points_per_s": value / 2.0, "single_cloud_latency_ms": latency_ms, "ms_per_step_single_cloud": latency_ms,
            "presorted_input": presorted, "matching": matching, "bundled": bundled, "slab": slab,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_resident": e2e_resident, "gpu_launches": int(lt.item()), "clocks": clocks,
            "host_binding": "GPU-local CPUs (NVML affinity)" if cpus_before else "none",
        }
        os.write(result_fd, (json.dumps(out) + "\n").encode())
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
