"""Power / clock / energy per launch of the step's kernel families, each looped back to back for ~1.5 s while NVML is
sampled every 10 ms (instantaneous board power where the driver exposes it, else the averaged reading).

Why: the replayed denoising step sits at the board's power cap (sw_power_cap, ~1.70 of 1.965 GHz).  In that regime the
step time is (energy per step) / (power cap): making a latency-bound kernel faster without removing work does not move
the step (profiles/README.md, the attention A/B), removing Joules does.  This probe says where the Joules go.

    python profiles/energy_probe.py            # kernels + cuBLAS reference + the whole step graph
"""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kandinsky-2_b200"))
import pynvml  # noqa: E402
import torch  # noqa: E402

from kandinsky2 import ops  # noqa: E402

pynvml.nvmlInit()
H = pynvml.nvmlDeviceGetHandleByIndex(0)
FI_INSTANT = getattr(pynvml, "NVML_FI_DEV_POWER_INSTANT", 186)


def read_power():
    try:
        v = pynvml.nvmlDeviceGetFieldValues(H, [FI_INSTANT])[0]
        if v.nvmlReturn == 0:
            return v.value.uiVal / 1e3
    except Exception:
        pass
    return pynvml.nvmlDeviceGetPowerUsage(H) / 1e3


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.rows = []
        self.stop = False

    def run(self):
        while not self.stop:
            self.rows.append((time.time(), read_power(), pynvml.nvmlDeviceGetClockInfo(H, pynvml.NVML_CLOCK_SM)))
            time.sleep(0.01)


def measure(name, fn, per_step, seconds=1.5, batch=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    smp = Sampler()
    smp.start()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    n = 0
    s.record()
    while time.time() - t0 < seconds:
        for _ in range(batch):
            fn()
        n += batch
        if n % (4 * batch) == 0:
            torch.cuda.synchronize()  # bounds the queue; a few us of idle per 200 launches
    e.record()
    torch.cuda.synchronize()
    t1 = time.time()
    smp.stop = True
    smp.join()
    us = s.elapsed_time(e) / n * 1e3
    mid = [r for r in smp.rows if t0 + 0.4 * (t1 - t0) <= r[0] <= t1 - 0.05]
    pw = sum(r[1] for r in mid) / max(1, len(mid))
    mhz = sorted(r[2] for r in mid)[len(mid) // 2] if mid else 0
    mj = pw * us * 1e-3  # W * us = uJ -> mJ
    print(f"{name:46s} {us:9.1f} us  {pw:6.0f} W  {mhz:5d} MHz  {mj:8.2f} mJ/launch  x{per_step:3d} = {mj * per_step / 1e3:6.3f} J/step",
          flush=True)
    return us, pw, mhz


def attention_modes():
    """python profiles/energy_probe.py attn: energy per launch of the level-1 / level-2 attention geometry for each softmax
    arithmetic mode (tuning key 6), layout (key 9) and start-up offset (key 5), then the whole step with the attention mode baked
    into its graph.  In an energy-bound step the mode with the fewest joules wins, whatever its time alone."""
    g = torch.Generator(device="cuda").manual_seed(0)
    for (B, heads, T, Tc, cnt) in [(8, 12, 2304, 32, 7), (8, 18, 576, 32, 7)]:
        qkv = torch.randn(B, T, heads * 192, device="cuda", generator=g).half()
        enc = torch.randn(B, Tc, heads * 128, device="cuda", generator=g).half()
        out = torch.empty(B, T, heads * 64, device="cuda", dtype=torch.float16)
        for (half, mode, stag) in [(1, 0, 1200), (1, 0, 0), (1, 30, 1200), (1, 1, 1200), (1, 31, 1200), (0, 0, 1200), (0, 30, 1200)]:
            ops.set_tuning(9, half)
            ops.set_tuning(6, mode)
            ops.set_tuning(5, stag)
            measure(f"attention T={T} layout {half} mode {mode:2d} stagger {stag:4d}", lambda: ops.attention_d64(qkv, heads, enc, out=out), cnt,
                    seconds=1.2)
    import bench
    from kandinsky2.model import unet as unet_mod
    from kandinsky2.model.gaussian_diffusion import FusedStep, create_ddpm_v22
    dev = torch.device("cuda", 0)
    ops.set_tuning(4, 1)
    model = unet_mod.Text2ImUNet(**bench.UNET_CFG, device=dev, param_dtype=torch.float16)
    model.init_synthetic_(seed=0)
    model.finalize(release_params=True)
    Bn = 4
    emb = torch.randn(2 * Bn, 1280, device=dev)
    coef, ts = create_ddpm_v22(50)._tables(dev)
    x = torch.randn(Bn, 4, 96, 96, device=dev)
    for rep in range(2):
        for (half, mode, stag) in [(1, 0, 1200), (1, 30, 1200), (1, 0, 0), (0, 30, 1200)]:
            ops.set_tuning(9, half)
            ops.set_tuning(6, mode)
            ops.set_tuning(5, stag)
            model._plans = {}
            st = FusedStep(model, Bn, 96, 96, dict(image_emb=emb), guidance_scale=4.0, cond_first=False, clip_range=2.0, threshold_mode=0)
            for n in range(3):
                st.noise.normal_()
                st.run(x, ts[40], coef[40])
            measure(f"whole step, attention layout {half} mode {mode:2d} stagger {stag:4d}", lambda: st.run(x, ts[25], coef[25]), 1,
                    seconds=2.0, batch=10)
    ops.set_tuning(9, 1)
    ops.set_tuning(6, 0)
    ops.set_tuning(5, 1200)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "attn":
        return attention_modes()
    g = torch.Generator(device="cuda").manual_seed(0)
    time.sleep(0.5)
    print(f"idle: {read_power():.0f} W, {pynvml.nvmlDeviceGetClockInfo(H, pynvml.NVML_CLOCK_SM)} MHz; "
          f"limit {pynvml.nvmlDeviceGetEnforcedPowerLimit(H) / 1e3:.0f} W")
    a = torch.randn(8192, 8192, device="cuda", generator=g).half()
    b = torch.randn(8192, 8192, device="cuda", generator=g).half()
    c = torch.empty(8192, 8192, device="cuda", dtype=torch.float16)
    us, _, _ = measure("cuBLAS fp16 8192^3", lambda: torch.matmul(a, b, out=c), 0, batch=10)
    print(f"    -> {2 * 8192 ** 3 / us / 1e6:.0f} TFLOP/s sustained")
    del a, b, c
    # convolutions (3x3, NHWC), the three dominant geometries; launches per step from the plan (approximate families)
    for (N, Hh, W, Cin, Cout, cnt) in [(8, 96, 96, 384, 384, 14), (8, 48, 48, 768, 768, 16), (8, 24, 24, 1152, 1152, 16),
                                        (8, 12, 12, 1536, 1536, 18)]:
        x = torch.randn(N, Hh, W, Cin, device="cuda", generator=g).half()
        w = torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / (3 * Cin ** 0.5)
        bias = torch.randn(Cout, device="cuda", generator=g)
        wp = ops.pack_conv_weight(w)
        y = torch.empty(N, Hh, W, Cout, device="cuda", dtype=torch.float16)
        us, _, _ = measure(f"conv3x3 {N}x{Hh}x{W} {Cin}->{Cout}", lambda: ops.conv_gemm([(x, 9)], wp, Cout, bias=bias, out=y), cnt)
        print(f"    -> {2 * N * Hh * W * Cin * Cout * 9 / us / 1e6:.0f} TFLOP/s")
    for (B, heads, T, Tc, cnt) in [(8, 12, 2304, 32, 7), (8, 18, 576, 32, 7), (8, 24, 144, 32, 8)]:
        qkv = torch.randn(B, T, heads * 192, device="cuda", generator=g).half()
        enc = torch.randn(B, Tc, heads * 128, device="cuda", generator=g).half()
        out = torch.empty(B, T, heads * 64, device="cuda", dtype=torch.float16)
        measure(f"attention T={T} heads={heads}", lambda: ops.attention_d64(qkv, heads, enc, out=out), cnt)
    for (NB, Hh, W, C, cnt) in [(8, 96, 96, 384, 12), (8, 48, 48, 768, 26), (8, 24, 24, 1152, 27), (8, 12, 12, 1536, 30)]:
        x = torch.randn(NB, Hh, W, C, device="cuda", generator=g).half()
        gamma = torch.randn(C, device="cuda", generator=g)
        beta = torch.randn(C, device="cuda", generator=g)
        film = torch.randn(NB, 2 * C, device="cuda", generator=g)
        y = torch.empty_like(x)
        st = ops.gn_stats(x)
        measure(f"gn_apply {NB}x{Hh}x{W}x{C} (FiLM + SiLU)", lambda: ops.gn_apply(x, None, st, gamma, beta, film=film, act=1, y=y), cnt)
    # the whole step graph
    import bench
    from kandinsky2.model import unet as unet_mod
    from kandinsky2.model.gaussian_diffusion import FusedStep, create_ddpm_v22
    dev = torch.device("cuda", 0)
    ops.set_tuning(4, 1)
    model = unet_mod.Text2ImUNet(**bench.UNET_CFG, device=dev, param_dtype=torch.float16)
    model.init_synthetic_(seed=0)
    model.finalize(release_params=True)
    Bn = 4
    emb = torch.randn(2 * Bn, 1280, device=dev)
    diffusion = create_ddpm_v22(50)
    st = FusedStep(model, Bn, 96, 96, dict(image_emb=emb), guidance_scale=4.0, cond_first=False, clip_range=2.0, threshold_mode=0)
    x = torch.randn(Bn, 4, 96, 96, device=dev)
    coef, ts = diffusion._tables(dev)
    for n in range(3):
        st.noise.normal_()
        st.run(x, ts[40], coef[40])

    def step():
        st.run(x, ts[25], coef[25])

    try:
        measure("whole denoising step (one graph launch)", step, 1, seconds=2.5, batch=10)
    except Exception as ex:  # the step API is exercised by bench.py; this line is a convenience only
        print("step graph not measured:", repr(ex))


if __name__ == "__main__":
    main()
