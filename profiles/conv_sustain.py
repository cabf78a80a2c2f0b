"""Sustained-state (power-capped) timing of the conv kernel's launch configurations at the lower UNet levels.

The tuner's default timing is the minimum of a few cool-chip launches; the replayed step runs at the board's power cap, where
a level-3 conv that keeps 80 of 148 SMs busy draws the same 985 W as a full-machine conv (the governor raises the clock of
the busy SMs instead: 1830 MHz against 1410-1460 MHz) and pays 1.6x the energy per FLOP (profiles/energy_probe.py).  This
probe loops every candidate (N tile, CTA-pair mode, split-K factor, epilogue warp sets) for 20 ms and times the next 20 ms.

    python profiles/conv_sustain.py
"""
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
shapes = [(8, 24, 24, 1152, 1152, 9), (8, 12, 12, 1536, 1536, 9), (8, 12, 12, 3072, 1536, 9), (8, 24, 24, 2304, 1152, 9),
          (8, 48, 48, 768, 768, 9), (8, 96, 96, 384, 384, 9), (8, 48, 48, 768, 2304, 1), (8, 24, 24, 1152, 3456, 1)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for (N, H, W, Cin, Cout, taps) in shapes:
    x = torch.randn(N, H, W, Cin, device="cuda", generator=g).half()
    k = 3 if taps == 9 else 1
    w = torch.randn(Cout, Cin, k, k, device="cuda", generator=g) / (k * Cin ** 0.5)
    b = torch.randn(Cout, device="cuda", generator=g)
    wp = ops.pack_conv_weight(w)
    y = torch.empty(N, H, W, Cout, device="cuda", dtype=torch.float16)
    part = torch.zeros(ops.gn_part_floats(N, H, W, Cout), device="cuda")
    flops = 2 * N * H * W * Cin * Cout * taps
    info = [0] * 7
    ops.conv_gemm([(x, taps)], wp, Cout, bias=b, out=y, gn_part=part, info=info)
    print(f"== conv {taps} taps {N}x{H}x{W} {Cin}->{Cout}: library choice N tile {info[0]}, pair {info[1]}, splits {info[2]}", flush=True)
    res = []
    for bn in (128, 192, 256):
        if bn - 64 >= Cout:
            continue
        for pair in (2, 1):
            for sp in (1, 2, 3, 4):
                for es in ((1, 2) if pair == 2 else (1,)):
                    cfg = (bn, pair, sp, es)
                    inf = [0] * 7
                    try:
                        ops.conv_gemm([(x, taps)], wp, Cout, bias=b, out=y, gn_part=part, info=inf, cfg=cfg)
                    except Exception:
                        continue
                    if inf[2] != sp or inf[0] != bn:
                        continue  # the library refused this combination
                    us = sustained(lambda: ops.conv_gemm([(x, taps)], wp, Cout, bias=b, out=y, gn_part=part, cfg=cfg))
                    res.append((us, cfg))
    res.sort()
    for us, cfg in res[:8]:
        print(f"   N tile {cfg[0]:3d} pair {cfg[1]} splits {cfg[2]} epilogue sets {cfg[3]}: {us:7.1f} us  {flops / us / 1e6:6.0f} TFLOP/s", flush=True)
    base = [u for u, c in res if c == (info[0], 2 if info[1] else 1, info[2], 1)]
    if base:
        print(f"   library choice sustained: {base[0]:.1f} us; best {res[0][0]:.1f} us ({100 * (1 - res[0][0] / base[0]):.1f} % less)")
