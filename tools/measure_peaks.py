"""Measure library GEMM throughput on this B200 the way MEASURED_PEAKS.json does (torch.matmul, best of 10),
for TF32 and for plain fp32 — the denominators the fp32 configs are quoted against.  Writes gpurun_out/peaks_tf32.json."""
import json, os, time
import torch

def bench(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

out = {"gpu": torch.cuda.get_device_name(0)}
n = 8192
a = torch.randn(n, n, device="cuda"); b = torch.randn(n, n, device="cuda")
torch.backends.cuda.matmul.allow_tf32 = True
ms = bench(lambda: a @ b); out["tf32_tflops_8192"] = 2 * n**3 / ms / 1e9
torch.backends.cuda.matmul.allow_tf32 = False
ms = bench(lambda: a @ b, 3); out["fp32_tflops_8192"] = 2 * n**3 / ms / 1e9
ab, bb = a.bfloat16(), b.bfloat16()
ms = bench(lambda: ab @ bb); out["bf16_tflops_8192"] = 2 * n**3 / ms / 1e9
# config-2 shapes through cuBLAS TF32 (library baseline for the three GEMMs)
torch.backends.cuda.matmul.allow_tf32 = True
B, i, o = 65536, 1024, 1024
X = torch.randn(B, i, device="cuda"); W = torch.randn(o, i, device="cuda"); dZ = torch.randn(B, o, device="cuda")
for name, fn in [("fwd", lambda: X @ W.t()), ("dX", lambda: dZ @ W), ("dW", lambda: dZ.t() @ X)]:
    ms = bench(fn); out[f"cfg2_{name}_tf32_ms"] = ms; out[f"cfg2_{name}_tf32_tflops"] = 2 * B * i * o / ms / 1e9
# copy bandwidth
src = torch.empty(1 << 30, dtype=torch.bfloat16, device="cuda"); dst = torch.empty_like(src)
ms = bench(lambda: dst.copy_(src)); out["hbm_copy_gbs"] = 2 * src.numel() * 2 / ms / 1e6
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/peaks_tf32.json", "w"), indent=1)
print(json.dumps(out))
