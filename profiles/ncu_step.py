"""One eager pass of the cfg-2 UNet launch plan (or of the MoVQ decode plan) between cudaProfilerStart / Stop, for
    ncu --profile-from-start off [--launch-skip I --launch-count 1 --set full ...] python profiles/ncu_step.py [unet|movq]
Prints the launch list (index inside the profiled range, kind, GFLOP of the reference graph) so that --launch-skip can pick a
kernel: the plan is built (and its conv configurations autotuned) BEFORE the profiled range."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kandinsky-2_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
from kandinsky2 import ops  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "unet"
dev = torch.device("cuda", 0)
ops.set_tuning(4, 0)  # no programmatic dependent launch under the profiler: kernels are serialised anyway
if what == "unet":
    from kandinsky2.model.gaussian_diffusion import FusedStep, create_ddpm_v22
    B, H, W = 4, 96, 96
    model = bench.build_unet(dev)
    coef, ts = create_ddpm_v22(50)._tables(dev)
    step = FusedStep(model, B, H, W, dict(image_emb=torch.randn(2 * B, 1280, device=dev)), 4.0, False, 2.0, 0)
    plan = step.plan
    plan.x_in.normal_()
    plan.t_in.fill_(500.0)
else:
    from kandinsky2.configs import CONFIG_2_2
    from kandinsky2.vqgan import MOVQ
    m = MOVQ(**CONFIG_2_2["image_enc_params"]["params"], device=dev, param_dtype=torch.float16).init_synthetic_(1)
    plan = m._plan("decode", 4, 96, 96)
    plan.x_in.normal_()
plan._serial = True   # forked branches in line: the launch order under the profiler = the order printed here
plan.launch()
torch.cuda.synchronize()
i = 0
for fn, kind, flops in plan.steps:  # index = kernels launched before this step (a split-K conv is 2 kernels, a join none)
    ops.reset_launch_count()
    fn()
    n = int(ops.launch_count())
    if n:
        print(f"launch {i:4d} {kind:16s} {flops / 1e9:9.2f} GFLOP" + (f"  ({n} kernels)" if n > 1 else ""), flush=True)
    i += n
torch.cuda.synchronize()
torch.cuda.profiler.start()
plan.launch()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
