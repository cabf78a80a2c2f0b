"""GPU parity of the diffusion prior (kandinsky2/model/prior.py, csrc/k2_prior.cu) against the outputs of the reference's own
PriorTransformer / PriorDiffusionModel (tests/golden/prior_tiny.pt, written by oracle/make_golden.py).  Tolerances: the
product keeps an fp16 residual stream (like the reference under use_fp16), the golden is fp32."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _setup():
    from kandinsky2.model.prior import PriorTransformer
    from oracle import prior_oracle as po, synth
    fx = torch.load(os.path.join(GOLD, "prior_tiny.pt"), weights_only=False)
    cfg = fx["cfg"]
    sd = synth.synth_state_dict(po.prior_param_spec(cfg), seed=fx["weight_seed"])
    m = PriorTransformer(**cfg, device="cuda")
    assert sorted(m.state_dict()) == sorted(sd)
    m.load_state_dict({k: v.cuda() for k, v in sd.items()}, strict=True)
    return fx, m.finalize()


def test_small_kernels_vs_torch():
    from kandinsky2 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(37, 128, device="cuda", generator=g).half()
    gam, bet = torch.randn(128, device="cuda", generator=g), torch.randn(128, device="cuda", generator=g)
    ref = torch.nn.functional.layer_norm(x.float(), (128,), gam, bet)
    assert (ops.layernorm_f16(x, gam, bet).float() - ref).abs().max().item() < 2e-2
    y = torch.randn(64, 96, device="cuda", generator=g).half()
    assert (ops.gelu_f16_(y.clone()).float() - torch.nn.functional.gelu(y.float())).abs().max().item() < 2e-3
    B, T, H = 3, 9, 2
    qkv = torch.randn(B, T, H * 192, device="cuda", generator=g).half()
    keep = torch.ones(B, T, dtype=torch.uint8, device="cuda")
    keep[1, 4:7] = 0
    out = ops.attention_small(qkv, H, keep_mask=keep, causal=True)
    q, k, v = qkv.float().view(B, T, H, 192).split(64, dim=-1)
    w = torch.einsum("bthc,bshc->bhts", q, k) * 0.125
    add = torch.where(keep.bool(), 0.0, float("-inf"))[:, None, None, :] + torch.full((T, T), float("-inf"), device="cuda").triu_(1)
    ref = torch.einsum("bhts,bshc->bthc", torch.softmax(w + add, dim=-1), v).reshape(B, T, H * 64)
    assert (out.float() - ref).abs().max().item() < 5e-3


def test_prior_forward_matches_reference_golden():
    fx, m = _setup()
    out = m(fx["x"].cuda(), fx["t"].cuda(), text_emb=fx["text_emb"].cuda(), text_enc=fx["text_enc"].cuda(), mask=fx["mask"].cuda())
    ref = fx["out"].cuda()
    rel = ((out - ref).norm() / ref.norm()).item()
    assert rel < 1e-2, rel  # fp16 residual stream against the fp32 reference


def test_prior_sampling_matches_reference_golden():
    from kandinsky2.model.prior import sample_prior
    fx, m = _setup()
    s = sample_prior(m, fx["text_emb"].cuda(), fx["text_enc"].cuda(), fx["mask"].cuda(), fx["use_steps"], fx["guidance"],
                     fx["clip_mean"].cuda(), fx["clip_std"].cuda(), fx["x_T"].cuda(), fx["step_noise"].cuda())
    ref = fx["sample"].cuda()
    rel = ((s - ref).norm() / ref.norm()).item()
    assert rel < 3e-2, rel
