"""Development probe: SM clock and power while the tensor-core matcher runs back to back (65536^2 x 352)."""
import sys, os, time, subprocess, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pcl_feature_extraction_b200 as pfx

dev = torch.device("cuda:0")
ctx = pfx.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
na = nb = 65536; dim = 352
g = torch.Generator(device=dev).manual_seed(1)
a = torch.rand((na, dim), device=dev, generator=g); b = torch.rand((nb, dim), device=dev, generator=g)
idx = torch.empty(na, dtype=torch.int32, device=dev); d2 = torch.empty(na, dtype=torch.float32, device=dev)
ctx.set_match_engine(1)
rows = []
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,temperature.gpu", "--format=csv,noheader,nounits", "-lms", "50", "-i", "0"],
                     stdout=subprocess.PIPE, text=True)
def rd():
    for line in p.stdout: rows.append((time.perf_counter(), line.strip()))
threading.Thread(target=rd, daemon=True).start()
time.sleep(0.5)
t0 = time.perf_counter()
n = 0
while time.perf_counter() - t0 < 4.0:
    for _ in range(20):
        ctx.match_nn_dev(a.data_ptr(), na, b.data_ptr(), nb, dim, idx.data_ptr(), d2.data_ptr())
    torch.cuda.synchronize(); n += 20
t1 = time.perf_counter()
time.sleep(0.3); p.terminate()
print(f"{n} matches in {t1 - t0:.2f} s = {(t1 - t0) / n * 1e3:.3f} ms each")
for t, r in rows:
    if t0 <= t <= t1: print(f"  t={t - t0:5.2f}s  sm_mhz, W, power_cap, temp = {r}")
