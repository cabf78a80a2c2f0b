"""Tail split (stream-K over the last partial wave of the CTA-pair conv kernel, tuning key 12) on / off: sustained-state timing
of the UNet's conv shapes at a given launch configuration (default: the library's own choice).
    python profiles/tail_probe.py [N,H,W,Cin,Cout,taps[,bn,pair,splits,es] ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kandinsky-2_b200"))
import torch  # noqa: E402

from kandinsky2 import ops  # noqa: E402


def sustained(fn, ms=20.0):
    s0, s1, s2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    fn()
    torch.cuda.synchronize()
    s0.record()
    fn()
    s1.record()
    torch.cuda.synchronize()
    n = max(4, int(ms / max(s0.elapsed_time(s1), 1e-3)))
    for _ in range(n):
        fn()
    s1.record()
    for _ in range(n):
        fn()
    s2.record()
    torch.cuda.synchronize()
    return s1.elapsed_time(s2) / n * 1e3


g = torch.Generator(device="cuda").manual_seed(0)
shapes = [(8, 24, 24, 1152, 1152, 9), (8, 24, 24, 2304, 1152, 9), (8, 96, 96, 384, 384, 9), (8, 96, 96, 768, 384, 9),
          (8, 12, 12, 1536, 1536, 9, 128, 2, 1, 1), (8, 48, 48, 768, 768, 9), (8, 24, 24, 1152, 3456, 1), (8, 24, 24, 1152, 1152, 1),
          (4, 32, 32, 1152, 1152, 9), (4, 64, 64, 768, 768, 9)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for sh in shapes:
    N, H, W, Cin, Cout, taps = sh[:6]
    cfg = tuple(sh[6:10]) if len(sh) >= 10 else None
    x = torch.randn(N, H, W, Cin, device="cuda", generator=g).half()
    k = 3 if taps == 9 else 1
    w = torch.randn(Cout, Cin, k, k, device="cuda", generator=g) / (k * Cin ** 0.5)
    b = torch.randn(Cout, device="cuda", generator=g)
    wp = ops.pack_conv_weight(w)
    y = torch.empty(N, H, W, Cout, device="cuda", dtype=torch.float16)
    part = torch.zeros(ops.gn_part_floats(N, H, W, Cout), device="cuda")
    flops = 2 * N * H * W * Cin * Cout * taps
    out = []
    for tail in (0, 1, 0, 1):
        ops.set_tuning(12, 2 * tail)  # 2 = wherever possible, regardless of the benefit model
        info = [0] * 7
        ops.conv_gemm([(x, taps)], wp, Cout, bias=b, out=y, gn_part=part, info=info, cfg=cfg)
        used = ops.conv_last_tail_split()
        us = sustained(lambda: ops.conv_gemm([(x, taps)], wp, Cout, bias=b, out=y, gn_part=part, cfg=cfg))
        out.append((tail, used, us))
    ops.set_tuning(12, 0)
    print(f"conv {taps} taps {N}x{H}x{W} {Cin}->{Cout} (N tile {info[0]}, pair {info[1]}, splits {info[2]}): " +
          "  ".join(f"tail {t} (parts {u}): {us:7.1f} us {flops / us / 1e6:5.0f} TF/s" for t, u, us in out), flush=True)
