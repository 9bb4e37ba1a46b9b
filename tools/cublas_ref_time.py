"""Library-GEMM reference time for the shapes tools/tcgen05_gemm_probe.cu runs (C = A @ B^T, bf16)."""
import sys
import torch

shapes = [(8192, 8192, 896), (16384, 8192, 3584), (16384, 151936, 896)]
for M, N, K in shapes:
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
    c = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        torch.matmul(a, b.t(), out=c)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20 if N < 100000 else 5
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b.t(), out=c)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"cuBLAS M={M} N={N} K={K}: {ms:.3f} ms, {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
    del a, b, c
