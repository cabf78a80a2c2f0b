"""One MoVQ decode of 4 latents 96x96 -> 768x768 inside a cudaProfilerStart/Stop range (ncu launch list) + event timing."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kandinsky-2_b200"))
import torch  # noqa: E402

from kandinsky2.configs import CONFIG_2_1  # noqa: E402
from kandinsky2.vqgan import MOVQ  # noqa: E402

dev = torch.device("cuda", 0)
m = MOVQ(**CONFIG_2_1["image_enc_params"]["params"], device=dev, param_dtype=torch.float16).init_synthetic_(1)
z = torch.randn(4, 4, 96, 96, device=dev)
m.decode_to_uint8(z)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
m.decode_to_uint8(z)
e.record()
torch.cuda.synchronize()
print(f"decode 4x768x768: {s.elapsed_time(e):.1f} ms")
torch.cuda.profiler.start()
m.decode_to_uint8(z)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
