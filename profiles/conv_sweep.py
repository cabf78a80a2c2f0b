"""Micro-benchmark of k2_conv_gemm on the UNet's conv shapes: {1-CTA, CTA-pair} x N tile x split-K, CUDA-event timed
(20 reps after 3 warm-ups, inputs re-used: weights + activations of one conv mostly fit L2, as inside a step).
    python profiles/conv_sweep.py > gpurun_out/conv_sweep.txt"""
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kandinsky-2_b200"))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from kandinsky2 import ops  # noqa: E402

SHAPES = [  # (N, H, W, Cin, Cout, label)
    (8, 96, 96, 384, 384, "L0 384->384"),
    (8, 96, 96, 768, 768, "L0 up 768->768"),
    (8, 96, 96, 1152, 384, "L0 1152->384"),
    (8, 48, 48, 768, 768, "L1 768->768"),
    (8, 48, 48, 1536, 768, "L1 1536->768"),
    (8, 24, 24, 1152, 1152, "L2 1152->1152"),
    (8, 24, 24, 2304, 1152, "L2 2304->1152"),
    (8, 12, 12, 1536, 1536, "L3 1536->1536"),
    (8, 12, 12, 3072, 1536, "L3 3072->1536"),
]


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
    return s.elapsed_time(e) / reps * 1e3  # us


TWO = (1, 2)
SPLITS = (1, 2, 3, 4, 6, 8)


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    check = os.environ.get("K2_SWEEP_CHECK", "1") == "1"
    global TWO, SPLITS
    shapes = SHAPES
    if os.environ.get("K2_SWEEP_BIG"):
        shapes = [s for s in SHAPES if s[1] >= 48]
        TWO, SPLITS = (1, 2), (1,)
    if os.environ.get("K2_SWEEP_SMALL"):
        shapes = [s for s in SHAPES if s[1] <= 24]
        TWO, SPLITS = (2,), (1, 2, 3, 4, 5, 6, 7, 8)
    for (N, H, W, Cin, Cout, label) in shapes:
        x = torch.randn(N, H, W, Cin, device="cuda", generator=g).half()
        w = torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / (3 * Cin ** 0.5)
        b = torch.randn(Cout, device="cuda", generator=g)
        wp = ops.pack_conv_weight(w)
        y = torch.empty(N, H, W, Cout, device="cuda", dtype=torch.float16)
        gflop = 2 * N * H * W * Cout * 9 * Cin / 1e9
        ref = None
        if check:
            ref = F.conv2d(x[:1].float().permute(0, 3, 1, 2), w.half().float(), b, padding=1).permute(0, 2, 3, 1)
        print(f"# {label}: M={N * H * W} K={9 * Cin} N={Cout} {gflop:.1f} GFLOP", flush=True)
        for two, bn, sp in itertools.product(TWO, (128, 192, 256), SPLITS):
            if sp > 1 and N * H * W > 8 * 48 * 48:
                continue
            ops.set_tuning(2, two); ops.set_tuning(0, bn); ops.set_tuning(1, sp)
            try:
                us = time_it(lambda: ops.conv_gemm([(x, 9)], wp, Cout, bias=b, out=y))
            except Exception as ex:  # e.g. forced split not feasible
                print(f"  2cta={two - 1} BN={bn} S={sp}: {str(ex)[:60]}")
                continue
            err = ""
            if ref is not None:
                rel = ((y[:1].float() - ref).norm() / ref.norm()).item()
                err = f" rel={rel:.1e}" + (" BAD" if rel > 2e-3 else "")
            print(f"  2cta={two - 1} BN={bn} S={sp}: {us:8.1f} us {gflop / us / 1e3:7.1f} TF/s{err}", flush=True)
        ops.set_tuning(2, 0); ops.set_tuning(0, 0); ops.set_tuning(1, 0)
        us = time_it(lambda: ops.conv_gemm([(x, 9)], wp, Cout, bias=b, out=y))
        print(f"  auto: {us:8.1f} us {gflop / us / 1e3:7.1f} TF/s", flush=True)


if __name__ == "__main__":
    main()
