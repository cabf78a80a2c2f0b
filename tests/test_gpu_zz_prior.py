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


def test_prior_embedder_drives_the_pipeline():
    """The prior wired in behind the pipelines' embedder protocol (kandinsky2_1_model.py:159-182, 300-344: generate_clip_emb
    -> generate_img): CLIP-text stand-in -> k2 prior sampling -> image embedding -> Kandinsky2_1.generate_text2img."""
    from kandinsky2 import get_kandinsky2
    from kandinsky2.model.prior import PriorEmbedder, PriorTransformer
    from kandinsky2.pipelines import SyntheticEmbedder
    from tests.test_gpu_movq_sampler import _tiny_overrides
    torch.manual_seed(0)
    prior = PriorTransformer(text_ctx=16, xf_width=128, xf_layers=2, xf_heads=2, xf_final_ln=True, xf_padding=False, clip_dim=768,
                             clip_xf_width=96, device="cuda")
    for p in prior.parameters():
        p.data.normal_(0.0, 0.05) if p.dim() > 1 else p.data.normal_(0.0, 0.02).add_(1.0 if p.dim() == 1 and p.numel() == 128 else 0.0)
    calls = []

    def clip_text(prompts):  # deterministic stand-in for tokenizer + CLIP text tower
        calls.append(list(prompts))
        outs = []
        for p in prompts:
            g = torch.Generator().manual_seed(len(p) + 17 * sum(map(ord, p)))
            outs.append((torch.randn(768, generator=g), torch.randn(16, 96, generator=g), torch.arange(16) < 3 + len(p) % 8))
        return tuple(torch.stack(t) for t in zip(*outs))

    syn = SyntheticEmbedder(768)
    emb = PriorEmbedder(prior, clip_text, clip_mean=torch.zeros(768, device="cuda"), clip_std=torch.ones(768, device="cuda"),
                        prior_steps="5", prior_cf_scale=4, text_encoder=syn.text_emb)
    e1, e2 = emb.image_emb("a red cat", 2), emb.image_emb("a red cat", 2)
    assert e1.shape == (2, 768) and torch.equal(e1, e2) and torch.isfinite(e1).all()
    assert calls[0] == ["a red cat", "a red cat", "", ""]      # [prompt x B | negative prior prompt x B]
    assert not torch.allclose(e1, emb.image_emb("a blue dog", 2))
    pipe = get_kandinsky2("cuda", task_type="text2img", model_version="2.1", cache_dir="/nonexistent", embedder=emb,
                          config_overrides=_tiny_overrides())
    imgs = pipe.generate_text2img("a red cat", num_steps=3, batch_size=2, guidance_scale=4, h=64, w=64, sampler="p_sampler")
    mixed = pipe.mix_images(["a cat", "a dog"], [0.4, 0.6], num_steps=3, batch_size=1, h=64, w=64, sampler="p_sampler")
    assert len(imgs) == 2 and imgs[0].size == (64, 64) and len(mixed) == 1
