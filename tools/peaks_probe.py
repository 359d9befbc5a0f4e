"""Yard-stick probe run on the GPU box: cuBLAS DGEMM (torch.matmul fp64), HBM copy, pinned H2D/D2H, host cores.
Not product code; output is committed under profiles/."""
import json, os, time, subprocess
import torch
out = {}
out["cpu_count"] = os.cpu_count()
try:
    out["cpu_model"] = [l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
except Exception as e:
    out["cpu_model"] = str(e)
out["mem_total_kb"] = [l.split()[1] for l in open("/proc/meminfo") if l.startswith("MemTotal")][0]
dev = torch.device("cuda:0")
out["gpu"] = torch.cuda.get_device_name(0)
def ev_time(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b); ms = ev_time(lambda: torch.matmul(a, b), 5)
    out[f"dgemm_{n}_tflops"] = 2 * n**3 / ms * 1e-9
    del a, b
# syrk-shaped: (256 x 2^20) @ (2^20 x 256)
X = torch.randn(1 << 20, 256, dtype=torch.float64, device=dev)
Xt = X.t().contiguous()
torch.matmul(Xt, X); ms = ev_time(lambda: torch.matmul(Xt, X), 5)
out["dgemm_XtX_N2e20_p256_ms"] = ms
out["dgemm_XtX_tflops_full"] = 2 * (1 << 20) * 256 * 256 / ms * 1e-9
ms = ev_time(lambda: torch.matmul(X.t(), X), 5)
out["dgemm_XtX_viewT_ms"] = ms
v = torch.randn(256, dtype=torch.float64, device=dev)
torch.mv(X, v); ms = ev_time(lambda: torch.mv(X, v), 10)
out["dgemv_Xv_ms"] = ms; out["dgemv_GBs"] = X.numel() * 8 / ms * 1e-6
del X, Xt
src = torch.empty(1 << 28, dtype=torch.float64, device=dev); dst = torch.empty_like(src)
dst.copy_(src); ms = ev_time(lambda: dst.copy_(src), 10)
out["hbm_copy_GBs"] = 2 * src.numel() * 8 / ms * 1e-6
ms = ev_time(lambda: src.sum(), 10)
out["hbm_read_sum_GBs"] = src.numel() * 8 / ms * 1e-6
del dst
h = torch.empty(1 << 27, dtype=torch.float64).pin_memory()
d = src[: 1 << 27]
d.copy_(h, non_blocking=True); ms = ev_time(lambda: d.copy_(h, non_blocking=True), 5)
out["h2d_pinned_GBs"] = h.numel() * 8 / ms * 1e-6
ms = ev_time(lambda: h.copy_(d, non_blocking=True), 5)
out["d2h_pinned_GBs"] = h.numel() * 8 / ms * 1e-6
# small-transfer latency
hs = torch.empty(259, dtype=torch.float64).pin_memory(); ds = torch.empty(259, dtype=torch.float64, device=dev)
t = time.perf_counter()
for _ in range(1000):
    ds.copy_(hs, non_blocking=True); hs.copy_(ds, non_blocking=True); torch.cuda.synchronize()
out["small_roundtrip_us"] = (time.perf_counter() - t) * 1e3
# potrf 256 latency (cuSOLVER yard-stick)
A = torch.randn(256, 256, dtype=torch.float64, device=dev); A = A @ A.t() + 256 * torch.eye(256, dtype=torch.float64, device=dev)
torch.linalg.cholesky(A); ms = ev_time(lambda: torch.linalg.cholesky(A), 10)
out["cusolver_potrf256_ms"] = ms
import numpy as np
t = time.perf_counter(); a = np.random.rand(2048, 2048); b = a @ a; out["numpy_dgemm2048_gflops"] = 2 * 2048**3 / (time.perf_counter() - t) * 1e-9
try:
    out["smi"] = subprocess.run(["nvidia-smi", "--query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit,memory.total", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
except Exception as e:
    out["smi"] = str(e)
print(json.dumps(out, indent=1))
