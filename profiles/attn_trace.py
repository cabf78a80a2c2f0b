"""Cycle-level hand-over trace of one attention CTA (tuning key 6 = 200 / 207, keys 7/8 = trace buffer).

Prints, for key blocks 0..15 of CTA (0,0,0): softmax warpgroup stamps (wait S, S arrived, S in registers, max done,
P written, arrive) and, per query tile, MMA-issuer stamps (wait s_free, S(j+1) issued, P(j) arrived, PV(j) issued)
relative to the first stamp."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kandinsky-2_b200"))
import torch  # noqa: E402

from kandinsky2 import ops  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
B, heads, T, Tc = 8, 12, 2304, 32
qkv = torch.randn(B, T, heads * 192, device="cuda", generator=g).half()
enc = torch.randn(B, Tc, heads * 128, device="cuda", generator=g).half()
out = torch.empty(B, T, heads * 64, device="cuda", dtype=torch.float16)
trace = torch.zeros(3 * 16 * 8, device="cuda", dtype=torch.int64)
addr = trace.data_ptr()


def s32(v):
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v >= (1 << 31) else v


ops.set_tuning(7, s32(addr))
ops.set_tuning(8, s32(addr >> 32))
for stagger in (0, 1000):
    mode = 200
    ops.set_tuning(6, mode)
    ops.set_tuning(5, stagger)
    for _ in range(3):
        ops.attention_d64(qkv, heads, enc, out=out)
    torch.cuda.synchronize()
    t = trace.cpu().view(3, 16, 8).tolist()
    base = min(v for r in t for b in r for v in b if v > 0)
    print(f"== mode {mode} stagger {stagger}")
    names = ["WG0", "WG1", "MMA"]
    for r in range(3):
        for j in range(16):
            print(f"{names[r]} j={j}: " + " ".join(f"{(v - base) if v else -1:7d}" for v in t[r][j]))
ops.set_tuning(6, 0)
ops.set_tuning(7, 0)
ops.set_tuning(8, 0)
