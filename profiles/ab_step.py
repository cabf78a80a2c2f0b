"""In-process A/B of launch-plan variants of the cfg-2 denoising step (graph replay, CUDA events, variants interleaved
round-robin so that clock / thermal drift hits all of them alike -- separate bench.py processes differ by ~2 % run to run).

    python profiles/ab_step.py "fold=0,tune=0" "fold=18,tune=0" "fold=18,tune=1" [--rounds 5 --steps 20 --hw 96x96 --batch 4]

Variant keys: fold = launch_plan.FOLD_MAX_RG, tune = launch_plan.TUNE, fork = launch_plan.FORK, smallm = launch_plan.TUNE_SMALL_M,
up2 = unet._UP2 (0/1), any k2_set_tuning key as
t<key>=<value>; drop=<kind>+<kind> removes every launch of those kinds from the captured graph (results are then WRONG: it measures
what a kernel family costs INSIDE the power-capped graph replay, which the eager per-kernel event sums cannot)."""
import argparse
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kandinsky-2_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
from kandinsky2 import launch_plan, ops  # noqa: E402
from kandinsky2.model import unet as unet_mod  # noqa: E402
from kandinsky2.model.gaussian_diffusion import FusedStep, create_ddpm_v22  # noqa: E402


TUNING_DEFAULTS = {4: 1, 5: 1200, 9: 1, 12: 0}  # k2_api.cu defaults that are not 0 (key 4 is switched on by this script)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("variants", nargs="+")
    ap.add_argument("--rounds", type=int, default=5)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--hw", default="96x96")
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--profile", action="store_true", help="also print the per-kind eager event sums of each variant")
    args = ap.parse_args()
    H, W = (int(v) for v in args.hw.split("x"))
    B = args.batch
    dev = torch.device("cuda", 0)
    ops.set_tuning(4, 1)
    model = unet_mod.Text2ImUNet(**bench.UNET_CFG, device=dev, param_dtype=torch.float16)
    model.init_synthetic_(seed=0)
    model.finalize(release_params=True)
    emb = torch.randn(2 * B, 1280, device=dev)
    diffusion = create_ddpm_v22(50)
    coef, ts = diffusion._tables(dev)
    steps = []
    for spec in args.variants:
        kv = dict(item.split("=") for item in spec.split(",") if item)
        launch_plan.FOLD_MAX_RG = int(kv.get("fold", launch_plan.FOLD_MAX_RG))
        launch_plan.TUNE = kv.get("tune", "1") != "0"
        launch_plan.FORK = kv.get("fork", "1") != "0"
        launch_plan.TUNE_SMALL_M = int(kv.get("smallm", "8192"))
        if hasattr(unet_mod, "_UP2"):
            unet_mod._UP2 = kv.get("up2", "1") != "0"
        tkeys = {int(k[1:]): int(v) for k, v in kv.items() if k[0] == "t" and k[1:].isdigit()}
        for k, v in tkeys.items():
            ops.set_tuning(k, v)
        model._plans = {}
        if kv.get("repack"):
            model.finalize()
        st = FusedStep(model, B, H, W, dict(image_emb=emb), guidance_scale=4.0, cond_first=False, clip_range=2.0,
                       threshold_mode=0)
        x = torch.randn(B, 4, H, W, device=dev)
        if kv.get("drop"):
            dropped = set(kv["drop"].split("+"))
            st.plan.steps = [s for s in st.plan.steps if s[1] not in dropped]
            st.plan.graph = None
        for n in range(3):  # builds + captures the graph with this variant's settings
            st.noise.normal_()
            st.run(x, ts[40], coef[40])
        for k in tkeys:
            ops.set_tuning(k, TUNING_DEFAULTS.get(k, 0))
        steps.append((spec, st, x))
    torch.cuda.synchronize()
    times = {spec: [] for spec, _, _ in steps}
    for r in range(args.rounds):
        for spec, st, x in steps:
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            s.record()
            for n in range(args.steps):
                st.noise.normal_()
                st.run(x, ts[49 - n % 50], coef[49 - n % 50])
            e.record()
            torch.cuda.synchronize()
            times[spec].append(s.elapsed_time(e) / args.steps)
    for spec, st, _ in steps:
        t = times[spec]
        line = f"{spec:32s} median {statistics.median(t):7.3f} ms/step  min {min(t):7.3f}  ({1e3 / statistics.median(t):6.2f} steps/s)  all {[round(v, 3) for v in t]}"
        if args.profile:
            prof = st.plan.profile(reps=2)
            line += "  " + str({k: round(v["ms"], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])})
        print(line, flush=True)


if __name__ == "__main__":
    main()
