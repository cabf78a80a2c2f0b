"""Calibration: what do the ncu tensor-pipe metrics read for cuBLAS on this GPU? (fp16 8192^3 and the 384->384 conv's GEMM shape)"""
import torch
for (M, N, K) in [(8192, 8192, 8192), (73728, 384, 3456), (73728, 768, 6912)]:
    a = torch.randn(M, K, device="cuda", dtype=torch.float16)
    b = torch.randn(K, N, device="cuda", dtype=torch.float16)
    for _ in range(3):
        c = a @ b
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        c = a @ b
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    print(f"cuBLAS fp16 {M}x{N}x{K}: {ms * 1e3:.1f} us {2 * M * N * K / ms / 1e9:.0f} TFLOP/s", flush=True)
torch.cuda.profiler.start()
a = torch.randn(73728, 3456, device="cuda", dtype=torch.float16)
b = torch.randn(3456, 384, device="cuda", dtype=torch.float16)
c = a @ b
a2 = torch.randn(73728, 6912, device="cuda", dtype=torch.float16)
b2 = torch.randn(6912, 768, device="cuda", dtype=torch.float16)
c2 = a2 @ b2
torch.cuda.synchronize()
torch.cuda.profiler.stop()
