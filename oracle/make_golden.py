"""Writes tests/golden/*.pt by EXECUTING THE REFERENCE (build container only: needs /root/reference).

    python -m oracle.make_golden

Every fixture holds the config, the weight seed (weights are re-synthesised from oracle/synth.py), the
inputs and the reference's outputs in fp32.  The same script asserts that the oracle restatement
(oracle/*_oracle.py) reproduces the reference on each fixture -- this is what pins the oracle.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim, synth  # noqa: E402
from oracle import unet_oracle as uo  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def build_ref_unet(cfg):
    with ref_shim.reference_modules() as R:
        m21 = R.load("model.text2im_model2_1")
        cls = m21.InpaintText2ImUNet if cfg.get("inpainting") else m21.Text2ImUNet
        model = cls(model_dim=cfg["model_dim"], image_encoder_in_dim=cfg["image_encoder_in_dim"],
                    text_encoder_in_dim1=cfg["text_encoder_in_dim1"], text_encoder_in_dim2=cfg["text_encoder_in_dim2"],
                    num_image_embs=cfg["num_image_embs"], pooling_type="from_model", in_channels=cfg["in_channels"],
                    model_channels=cfg["model_channels"], out_channels=cfg["out_channels"],
                    num_res_blocks=cfg["num_res_blocks"], attention_resolutions=tuple(cfg["attention_ds"]), dropout=0,
                    channel_mult=cfg["channel_mult"], use_fp16=False, num_heads=1,
                    num_head_channels=cfg["num_head_channels"], num_heads_upsample=-1, use_scale_shift_norm=True,
                    resblock_updown=True, cache_text_emb=True)
    return model.eval()


def unet_inputs(cfg, B, H, W, ntext, seed):
    g = torch.Generator().manual_seed(seed)
    d = dict(x=torch.randn(B, cfg["in_channels"], H, W, generator=g),
             t=torch.tensor([999.0, 500.0, 20.0, 0.0][:B]),
             full_emb=torch.randn(B, ntext, cfg["text_encoder_in_dim1"], generator=g),
             pooled_emb=torch.randn(B, cfg["text_encoder_in_dim2"], generator=g),
             image_emb=torch.randn(B, cfg["image_encoder_in_dim"], generator=g))
    if cfg.get("inpainting"):
        d["inpaint_image"] = torch.randn(B, cfg["in_channels"], H, W, generator=g)
        d["inpaint_mask"] = (torch.rand(B, 1, H, W, generator=g) > 0.5).float()
    return d


def golden_unet(name, cfg, B, H, W, ntext, wseed, iseed):
    model = build_ref_unet(cfg)
    spec = uo.unet_param_spec(cfg)
    ref_keys = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    assert ref_keys == [(k, tuple(s)) for k, s in spec], "oracle parameter spec != reference state_dict"
    sd = synth.synth_state_dict(spec, seed=wseed)
    model.load_state_dict(sd, strict=True)
    inp = unet_inputs(cfg, B, H, W, ntext, iseed)
    kw = {k: v for k, v in inp.items() if k not in ("x", "t")}
    with torch.no_grad():
        y_ref = model(inp["x"], inp["t"], **kw)
        y_orc = uo.unet_forward(sd, cfg, inp["x"], inp["t"], **kw)
    err = (y_ref - y_orc).abs().max().item()
    assert err <= 1e-5, f"{name}: oracle deviates from the reference by {err}"
    torch.save(dict(cfg=cfg, weight_seed=wseed, inputs=inp, out=y_ref, shape=(B, H, W), ntext=ntext,
                    weight_checksum=float(sum(v.double().sum() for v in sd.values()))),
               os.path.join(GOLD, name + ".pt"))
    print(f"{name}: reference out std {y_ref.std():.4f}, oracle-vs-reference max abs {err:.2e}")


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    golden_unet("unet_tiny", uo.CONFIG_TINY, 2, 16, 16, 7, wseed=1, iseed=5)
    golden_unet("unet_tiny_inpaint", dict(uo.CONFIG_TINY, inpainting=True), 2, 16, 16, 7, wseed=2, iseed=6)
    for extra in EXTRA:
        extra()


EXTRA = []

if __name__ == "__main__":
    main()
