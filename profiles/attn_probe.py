"""Attention kernel timing at the three UNet geometries for several warpgroup de-phasing delays (tuning key 5)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kandinsky-2_b200"))
import torch  # noqa: E402

from kandinsky2 import ops  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
for (B, heads, T, Tc) in [(8, 12, 2304, 32), (8, 18, 576, 32), (8, 24, 144, 32)]:
    qkv = torch.randn(B, T, heads * 192, device="cuda", generator=g).half()
    enc = torch.randn(B, Tc, heads * 128, device="cuda", generator=g).half()
    out = torch.empty(B, T, heads * 64, device="cuda", dtype=torch.float16)
    flops = 4 * B * heads * T * (T + Tc) * 64
    # tuning key 6: eighths of the exponentials on the FMA pipe; key 5: initial de-phasing of the two query tiles (cycles)
    # key 9: MMA issue order (0 fixed per key block, 1 event driven)
    for mode, delay, stag in ((0, 0, 300), (1, 0, 0), (1, 0, 300), (1, 0, 1000), (1, 0, 1700), (1, 0, 2400), (1, 2, 1700), (1, 3, 1700)):
        ops.set_tuning(9, mode)
        ops.set_tuning(6, delay)
        ops.set_tuning(5, stag)
        for _ in range(3):
            ops.attention_d64(qkv, heads, enc, out=out)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20):
            ops.attention_d64(qkv, heads, enc, out=out)
        e.record()
        torch.cuda.synchronize()
        us = s.elapsed_time(e) / 20 * 1e3
        print(f"T={T} heads={heads} issue={mode} poly={delay}/8 stagger={stag}: {us:.1f} us {flops / us / 1e6:.0f} TF/s", flush=True)

ops.set_tuning(9, 0)
ops.set_tuning(6, 0)
ops.set_tuning(5, 300)
