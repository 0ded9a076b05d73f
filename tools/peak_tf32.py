"""Measured cuBLAS TF32 / BF16 dense GEMM throughput on this GPU (roofline context for 3xTF32)."""
import json, torch
torch.backends.cuda.matmul.allow_tf32 = True
res = {}
for name, dt in (("tf32", torch.float32), ("bf16", torch.bfloat16)):
    for N in (4096, 8192):
        a = torch.randn(N, N, device="cuda", dtype=dt); b = torch.randn(N, N, device="cuda", dtype=dt)
        for _ in range(3): a @ b
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        # sustained: 2 s back to back
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(10, int(2000 / best))
        e0.record()
        for _ in range(reps): c = a @ b
        e1.record(); torch.cuda.synchronize()
        res[f"{name}_{N}"] = dict(burst_tflops=2 * N**3 / (best * 1e-3) / 1e12, sustained_tflops=2 * N**3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12)
print(json.dumps(res, indent=1))
