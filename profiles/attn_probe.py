"""Attention kernel (k2_attention_d64) at the three UNet geometries: checked against a torch fp32 softmax(QK^T)V on a
slice, then timed with CUDA events for each MUFU-free share of the exponentials (tuning key 6) and two de-phasing delays
(key 5); the last part prints the clock64 hand-over trace of CTA (0,0,0) (softmax points per key block: 0 block start,
1 rescale decision taken, 2 first 64 exponentials done, 3 96 done + PV(j-1) barrier passed, 4 all 128 done, 5 P handed
over / next maximum known; MMA points per tile: P arrived, PV issued, S buffer free, S(j+2) issued).
History of this probe's output: attn_probe_r2*.txt (a: first P-in-TMEM variants of the 16-warp kernel, b: row-per-thread
kernel, c: forced 8-deep MUFU bursts (slower), d: next-block maximum folded in + deferred PV wait, e: de-phasing sweep)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kandinsky-2_b200"))
import torch  # noqa: E402

from kandinsky2 import ops  # noqa: E402


def reference(qkv, enc, heads, b=0, nq=256):
    B, T, _ = qkv.shape
    x = qkv[b].float().view(T, heads, 3, 64)
    e = enc[b].float().view(-1, heads, 2, 64)
    q = x[:nq, :, 0]
    k = torch.cat([e[:, :, 0], x[:, :, 1]], 0)
    v = torch.cat([e[:, :, 1], x[:, :, 2]], 0)
    w = torch.softmax(torch.einsum("qhd,khd->hqk", q, k) * 0.125, -1)
    return torch.einsum("hqk,khd->qhd", w, v).reshape(nq, heads * 64)


# tuning key 6: eighths of the exponentials on the FMA pipe, + 10 = FFMA2 scale-and-subtract, + 30 = FFMA2 and FADD2 row sums
MODES = [int(v) for v in os.environ.get("K2_ATTN_MODES", "0,1,2,10,11,12,30,31").split(",")]
# tuning key 5: cycles the second query tile's MMA issuer starts late (the tiles have independent issuers and stay apart)
STAGGER = [int(v) for v in os.environ.get("K2_ATTN_STAGGER", "0").split(",")]
TRACE_STAGGER = int(os.environ.get("K2_ATTN_TRACE_STAGGER", "0"))
g = torch.Generator(device="cuda").manual_seed(0)
geoms = [(8, 12, 2304, 32), (8, 18, 576, 32), (8, 24, 144, 32)]
if len(sys.argv) > 1 and sys.argv[1] == "cfg3":
    geoms = [(4, 12, 4096, 32), (4, 18, 1024, 32), (4, 24, 256, 32)]
for (B, heads, T, Tc) in geoms:
    qkv = torch.randn(B, T, heads * 192, device="cuda", generator=g).half()
    enc = torch.randn(B, Tc, heads * 128, device="cuda", generator=g).half()
    out = torch.empty(B, T, heads * 64, device="cuda", dtype=torch.float16)
    flops = 4 * B * heads * T * (T + Tc) * 64
    nq = min(T, 384)
    ref = reference(qkv, enc, heads, b=B - 1, nq=nq)
    for half, poly, stag in [(h, m, st) for h in (1, 0) for m in MODES for st in STAGGER]:
        ops.set_tuning(9, half)
        ops.set_tuning(6, poly)
        ops.set_tuning(5, stag)
        out.zero_()
        for _ in range(3):
            ops.attention_d64(qkv, heads, enc, out=out)
        torch.cuda.synchronize()
        err = (out[B - 1, :nq].float() - ref).abs().max().item()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20):
            ops.attention_d64(qkv, heads, enc, out=out)
        e.record()
        torch.cuda.synchronize()
        us = s.elapsed_time(e) / 20 * 1e3
        print(f"T={T} heads={heads} {'half rows (16 warps)' if half else 'full rows (8 warps) '} mode={poly:2d} stagger={stag:4d}: {us:.1f} us {flops / us / 1e6:.0f} TF/s "
              f"max|err| vs fp32 {err:.2e}", flush=True)

# hand-over trace, ping-pong mode
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
for half in (1, 0):
    ops.set_tuning(9, half)
    ops.set_tuning(6, 200)
    ops.set_tuning(5, TRACE_STAGGER)
    trace.zero_()
    for _ in range(3):
        ops.attention_d64(qkv, heads, enc, out=out)
    torch.cuda.synchronize()
    t = trace.cpu().view(3, 16, 8).tolist()
    base = min(v for r in t for b in r for v in b if v > 0)
    print(f"== trace, key 9 = {half}, stagger {TRACE_STAGGER}")
    names = ["WG0", "WG1", "MMA"]
    for r in range(3):
        for j in range(16):
            print(f"{names[r]} j={j}: " + " ".join(f"{(v - base) if v else -1:7d}" for v in t[r][j]))
ops.set_tuning(6, 0)
ops.set_tuning(7, 0)
ops.set_tuning(8, 0)
ops.set_tuning(5, 1200)
ops.set_tuning(9, 1)
