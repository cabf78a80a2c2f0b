"""Warm, back-to-back timings of the GroupNorm glue kernels at the four UNet levels (the ncu launch list times them cold
and serialised): gn_apply (normalise + FiLM + SiLU), gn_finalize (from per-tile partials), and a conv + finalize + apply
chain as it occurs in a ResBlock.  Round-2 starting point for the gn_apply / gn_finalize work."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kandinsky-2_b200"))
import torch  # noqa: E402

from kandinsky2 import ops  # noqa: E402


def timeit(fn, reps=50):
    for _ in range(5):
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
for (NB, H, W, C) in [(8, 96, 96, 384), (8, 48, 48, 768), (8, 24, 24, 1152), (8, 12, 12, 1536)]:
    x = torch.randn(NB, H, W, C, device="cuda", generator=g).half()
    gamma = torch.randn(C, device="cuda", generator=g)
    beta = torch.randn(C, device="cuda", generator=g)
    film = torch.randn(NB, 2 * C, device="cuda", generator=g)
    y = torch.empty_like(x)
    st = ops.gn_stats(x)
    mb = 2 * x.numel() * 2 / 1e6
    us = timeit(lambda: ops.gn_apply(x, None, st, gamma, beta, film=film, act=1, y=y))
    print(f"gn_apply  {NB}x{H}x{W}x{C}: {us:6.1f} us  {mb / us:6.2f} TB/s ({mb:.0f} MB read+write)", flush=True)
    w = torch.randn(C, C, 3, 3, device="cuda", generator=g) / (3 * C ** 0.5)
    wp = ops.pack_conv_weight(w)
    part = torch.zeros(ops.gn_part_floats(NB, H, W, C), device="cuda")
    info = [0] * 7
    out = torch.empty_like(x)
    ops.conv_gemm([(x, 9)], wp, C, out=out, gn_part=part, info=info)
    st2 = torch.empty(NB, 32, 2, device="cuda")
    us_f = timeit(lambda: ops.gn_finalize(part, C, None, 0, NB, info[6] // NB, H * W, st2))
    print(f"gn_finalize (mode {info[5]}, {info[6] // NB} row groups per image): {us_f:6.1f} us", flush=True)

    def chain():
        ops.conv_gemm([(x, 9)], wp, C, out=out, gn_part=part)
        ops.gn_finalize(part, C, None, 0, NB, info[6] // NB, H * W, st2)
        ops.gn_apply(out, None, st2, gamma, beta, film=film, act=1, y=y)

    us_c = timeit(chain, reps=20)
    us_conv = timeit(lambda: ops.conv_gemm([(x, 9)], wp, C, out=out, gn_part=part), reps=20)
    print(f"conv {us_conv:6.1f} us; conv + finalize + apply chain {us_c:6.1f} us (glue = {us_c - us_conv:5.1f} us)", flush=True)
