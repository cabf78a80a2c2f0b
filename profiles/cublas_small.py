import torch
for (M, N, K, lab) in [(1152, 1536, 13824, "L3 1536->1536"), (1152, 1536, 27648, "L3 3072->1536"), (4608, 1152, 10368, "L2 1152->1152"),
                       (4608, 1152, 20736, "L2 2304->1152"), (18432, 768, 6912, "L1 768->768"), (18432, 2304, 768, "L1 qkv"), (18432, 768, 768, "L1 proj")]:
    a = torch.randn(M, K, device="cuda", dtype=torch.float16)
    b = torch.randn(N, K, device="cuda", dtype=torch.float16)
    for _ in range(3):
        c = a @ b.t()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        c = a @ b.t()
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) / 20 * 1e3
    print(f"cuBLAS {lab}: {M}x{N}x{K} {us:.1f} us {2 * M * N * K / us / 1e6:.0f} TFLOP/s", flush=True)
