"""Small-K (attention qkv / proj 1x1) GEMM timings vs cuBLAS, with the epilogue options toggled.

The K loop of these launches is short (12..24 chunks of 64), so the per-tile epilogue is what is measured."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kandinsky-2_b200"))
import torch  # noqa: E402

from kandinsky2 import ops  # noqa: E402


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3


g = torch.Generator(device="cuda").manual_seed(0)
for (M, K, N, name) in [(18432, 768, 2304, "L1 qkv"), (18432, 768, 768, "L1 proj"), (4608, 1152, 3456, "L2 qkv"),
                        (4608, 1152, 1152, "L2 proj"), (1152, 1536, 4608, "L3 qkv"), (1152, 1536, 1536, "L3 proj")]:
    x = torch.randn(M, K, device="cuda", generator=g).half()
    w = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    r = torch.randn(M, N, device="cuda", generator=g).half()
    wp = ops.pack_conv_weight(w)
    wt = w.half().t().contiguous()
    y = torch.empty(M, N, device="cuda", dtype=torch.float16)
    gf = 2.0 * M * K * N / 1e9
    us = timeit(lambda: torch.matmul(x, wt, out=y))
    print(f"# {name}: M={M} K={K} N={N} {gf:.1f} GFLOP   cuBLAS {us:.1f} us {gf / us * 1e-3 * 1e3:.0f} TF/s", flush=True)
    for tune, label in [({}, "auto"), ({0: 128}, "BN=128"), ({0: 192}, "BN=192"), ({2: 1}, "1-CTA"), ({2: 1, 0: 128}, "1-CTA BN=128")]:
        for k in (0, 1, 2):
            ops.set_tuning(k, 0)
        for k, v in tune.items():
            ops.set_tuning(k, v)
        for bias, res, tag in [(None, None, "plain"), (b, None, "bias"), (b, r, "bias+res")]:
            info = [0] * 7
            fn = lambda: ops.gemm_rows(x, wp, N, bias=bias, residual=res, out=y)  # noqa: E731
            us = timeit(fn)
            print(f"  {label:14s} {tag:9s}: {us:7.1f} us {gf / us:6.0f} TF/s", flush=True)
    for k in (0, 1, 2):
        ops.set_tuning(k, 0)
