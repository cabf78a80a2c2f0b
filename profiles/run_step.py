"""One eager (un-graphed) denoising step of the bench workload inside a cudaProfilerStart/Stop range, for
    ncu --profile-from-start off ... python profiles/run_step.py
(the launch list and the --set full capture of the conv kernel committed under profiles/)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kandinsky-2_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
from kandinsky2.model.gaussian_diffusion import FusedStep, create_ddpm_v22  # noqa: E402
from kandinsky2.model.unet import Text2ImUNet  # noqa: E402

B, H, W = 4, 96, 96
dev = torch.device("cuda", 0)
model = Text2ImUNet(**bench.UNET_CFG, device=dev, param_dtype=torch.float16)
model.init_synthetic_(0)
model.finalize(release_params=True)
model.use_cuda_graph = False
diff = create_ddpm_v22(50)
coef, ts = diff._tables(dev)
step = FusedStep(model, B, H, W, dict(image_emb=torch.randn(2 * B, 1280, device=dev)), 4.0, False, 2.0, 0)
x = torch.randn(B, 4, H, W, device=dev)
step.noise.normal_()
step.run(x, ts[49], coef[49])  # warm-up (sets function attributes, loads modules)
torch.cuda.synchronize()
torch.cuda.profiler.start()
step.run(x, ts[48], coef[48])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
