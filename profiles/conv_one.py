"""One conv shape, a few launches, for `ncu --set full -k regex:conv_gemm` (env K2_SHAPE = N,H,W,Cin,Cout)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kandinsky-2_b200"))
import torch  # noqa: E402

from kandinsky2 import ops  # noqa: E402

N, H, W, Cin, Cout = [int(v) for v in os.environ.get("K2_SHAPE", "8,12,12,1536,1536").split(",")]
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(N, H, W, Cin, device="cuda", generator=g).half()
w = torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / (3 * Cin ** 0.5)
b = torch.randn(Cout, device="cuda", generator=g)
wp = ops.pack_conv_weight(w)
y = torch.empty(N, H, W, Cout, device="cuda", dtype=torch.float16)
for k, v in [(int(a), int(b_)) for a, b_ in (kv.split("=") for kv in os.environ.get("K2_TUNE", "").split(",") if kv)]:
    ops.set_tuning(k, v)
for _ in range(3):
    ops.conv_gemm([(x, 9)], wp, Cout, bias=b, out=y)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ops.conv_gemm([(x, 9)], wp, Cout, bias=b, out=y)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
