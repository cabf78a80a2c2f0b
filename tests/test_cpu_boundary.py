"""CPU: the drop-in boundary -- every symbol of include/k2b200.h is exported by libk2b200.so, the product modules
expose the reference's state_dict keys / constructor surface, and the product refuses to run without a GPU."""
import os
import re
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol():
    from kandinsky2 import _native
    hdr = open(os.path.join(ROOT, "include", "k2b200.h")).read()
    declared = set(re.findall(r"\b(k2_[a-z0-9_]+)\s*\(", hdr))
    lib = _native.load()
    assert not _native.MISSING, _native.MISSING
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in k2b200.h but not exported"
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    assert lib.k2_version() >= 100


def test_unet_state_dict_keys_match_reference_spec():
    from kandinsky2.configs import CONFIG_2_1
    from kandinsky2.model.model_creation import create_model
    from oracle import unet_oracle as uo
    cfg = dict(CONFIG_2_1["model_config"], num_channels=64, num_res_blocks=1)  # same topology rules, small
    m = create_model(**cfg, up=False, inpainting=False)
    ocfg = dict(uo.CONFIG_2_1, model_channels=64, num_res_blocks=1)
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, tuple(s)) for k, s in uo.unet_param_spec(ocfg)]
    mi = create_model(**cfg, up=False, inpainting=True)
    assert mi.state_dict()["input_blocks.0.0.weight"].shape[1] == 9
    assert m.dtype == torch.float16 and m.model_channels == 64
    with pytest.raises(NotImplementedError):
        create_model(**dict(cfg, use_scale_shift_norm=False), up=False, inpainting=False)


def test_movq_state_dict_keys_match_reference_spec():
    from kandinsky2.vqgan import MOVQ
    from oracle import movq_oracle as mo
    m = MOVQ(mo.DDCONFIG_TINY, 64, 4)
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == \
        [(k, tuple(s)) for k, s in mo.movq_param_spec(dict(mo.DDCONFIG_TINY, double_z=False), 4, 64)]
    # training checkpoints may carry loss.* entries: dropped
    sd = dict(m.state_dict())
    sd["loss.discriminator.main.0.weight"] = torch.zeros(1)
    m.load_state_dict(sd, strict=True)


def test_no_cpu_fallback():
    from kandinsky2._native import K2Error
    from kandinsky2.model.unet import Text2ImUNet
    from oracle import unet_oracle as uo
    if torch.cuda.is_available():
        pytest.skip("checks the CPU-only failure mode")
    cfg = uo.CONFIG_TINY
    m = Text2ImUNet(model_dim=cfg["model_dim"], image_encoder_in_dim=48, text_encoder_in_dim1=96, text_encoder_in_dim2=48,
                    num_image_embs=3, pooling_type="from_model", in_channels=4, model_channels=64, out_channels=8,
                    num_res_blocks=1, attention_resolutions=(2,), channel_mult=(1, 2), use_fp16=True,
                    num_head_channels=64, use_scale_shift_norm=True, resblock_updown=True)
    with pytest.raises(K2Error):
        m(torch.zeros(2, 4, 16, 16), torch.zeros(2), full_emb=torch.zeros(2, 7, 96), pooled_emb=torch.zeros(2, 48),
          image_emb=torch.zeros(2, 48))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "kandinsky-2_b200", "kandinsky2")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"
