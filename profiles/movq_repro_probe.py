"""Does building / running a second MoVQ launch plan (another batch size) change the results of the first one?  (The last
assertion of tests/test_gpu_movq_sampler.py::test_movq_decode_full_size_vs_oracle started to fail on some boxes.)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kandinsky-2_b200"))
import torch  # noqa: E402

from kandinsky2 import launch_plan  # noqa: E402
from kandinsky2.vqgan import MOVQ  # noqa: E402
from oracle import movq_oracle as mo, synth  # noqa: E402

dd = dict(mo.DDCONFIG_2_1)
sd = synth.synth_state_dict(mo.movq_param_spec(dd, 4, 16384), seed=10)
m = MOVQ(dd, 16384, 4)
m.load_state_dict(sd)
m.to("cuda")
z = torch.randn(2, 4, 96, 96, generator=torch.Generator().manual_seed(2)).cuda()
m.use_cuda_graph = True


def diff(a, b):
    d = (a - b).abs()
    return f"{int((d > 0).sum())} elements differ, max abs {d.max().item():.3e}"


y = [m.decode(z) for _ in range(4)]
print("plan A, runs 1-3 vs run 0:", [diff(v, y[0]) for v in y[1:]], flush=True)
planA = m._plan("decode", 2, 96, 96)
cfgA = [k for k in launch_plan._tune_cache]
print("tuned shapes after plan A:", len(cfgA), flush=True)
y1 = m.decode(z[1:])
print("plan B built; tuned shapes now:", len(launch_plan._tune_cache), flush=True)
y2 = [m.decode(z) for _ in range(3)]
print("plan A after plan B, vs run 0:", [diff(v, y[0]) for v in y2], flush=True)
print("same plan object:", m._plan("decode", 2, 96, 96) is planA, flush=True)
m.use_cuda_graph = False
y3 = m.decode(z)
print("plan A eager after plan B vs run 0:", diff(y3, y[0]), flush=True)
u8 = m.decode_to_uint8(z, crop_h=760, crop_w=768)
print("u8 vs process_images(run 0):", int((u8 != mo.process_images(y[0])[:, :760, :768]).sum()), "bytes differ", flush=True)
