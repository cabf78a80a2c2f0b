"""Correctness + speed probe of the halo 3x3 kernel layout variants (k2_set_tuning key 3) against the CTA-pair kernel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kandinsky-2_b200"))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from kandinsky2 import ops  # noqa: E402


def time_it(fn, reps=20):
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
for (N, H, W, Cin, Cout) in [(2, 32, 32, 64, 128), (8, 96, 96, 384, 384), (8, 96, 96, 768, 768), (8, 48, 48, 768, 768),
                             (8, 96, 96, 1152, 384)]:
    x = torch.randn(N, H, W, Cin, device="cuda", generator=g).half()
    w = torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / (3 * Cin ** 0.5)
    b = torch.randn(Cout, device="cuda", generator=g)
    wp = ops.pack_conv_weight(w)
    y = torch.empty(N, H, W, Cout, device="cuda", dtype=torch.float16)
    ref = F.conv2d(x[:1].float().permute(0, 3, 1, 2), w.half().float(), b, padding=1).permute(0, 2, 3, 1)
    gflop = 2 * N * H * W * Cout * 9 * Cin / 1e9
    print(f"# {N}x{H}x{W} {Cin}->{Cout} {gflop:.1f} GFLOP", flush=True)
    for mode in (0, 1, 2, 3, 4):
        ops.set_tuning(3, mode)
        try:
            y.zero_()
            us = time_it(lambda: ops.conv_gemm([(x, 9)], wp, Cout, bias=b, out=y))
            rel = ((y[:1].float() - ref).norm() / ref.norm()).item()
            print(f"  halo mode {mode}: {us:8.1f} us {gflop / us / 1e3:7.1f} TF/s rel={rel:.1e} {'OK' if rel < 2e-3 else 'WRONG'}", flush=True)
        except Exception as ex:
            print(f"  halo mode {mode}: FAILED {str(ex)[:100]}", flush=True)
            break
    ops.set_tuning(3, 0)
