"""MoVQ decode of 4 latents 96x96 -> 4 x 768x768 (the tail of BASELINE configs[1]): graph-replay time (CUDA events) and the
per-kernel-family event sums of one eager pass of the launch plan.
    python profiles/movq_time.py [B h w]"""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kandinsky-2_b200"))
import torch  # noqa: E402

from kandinsky2 import ops  # noqa: E402
from kandinsky2.configs import CONFIG_2_2  # noqa: E402
from kandinsky2.vqgan import MOVQ  # noqa: E402

B, h, w = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (4, 96, 96)
dev = torch.device("cuda", 0)
ops.set_tuning(4, 1)
m = MOVQ(**CONFIG_2_2["image_enc_params"]["params"], device=dev, param_dtype=torch.float16).init_synthetic_(1)
z = torch.randn(B, 4, h, w, device=dev)
for _ in range(3):
    u8 = m.decode_to_uint8(z)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    u8 = m.decode_to_uint8(z)
    e.record()
    torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
plan = m._plan("decode", B, h, w)
prof = plan.profile(reps=2)
flops = sum(v["flops"] for v in prof.values())
print(f"decode_to_uint8 {B} x {h}x{w} -> {tuple(u8.shape)}: median {statistics.median(ts):.2f} ms, min {min(ts):.2f} ms; "
      f"{len(plan.steps)} launches, reference-graph {flops / 1e12:.2f} TFLOP -> {flops / (statistics.median(ts) * 1e-3) / 1e12:.0f} TFLOP/s")
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"  {k:14s} {v['launches']:4d} launches {v['ms']:8.3f} ms  {v['flops'] / 1e12:7.3f} TFLOP")
