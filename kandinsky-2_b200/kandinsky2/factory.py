"""get_kandinsky2 with the reference's signature (kandinsky2/__init__.py:164-192).

The reference downloads checkpoints from the HF Hub (:100-152); this build has no network: if
`<cache_dir>/2_1/decoder_fp16.ckpt` (or inpainting_fp16.ckpt) and `movq_final.ckpt` exist they are loaded (the state
dict keys are the reference's), otherwise the networks get random weights of the architecture."""
import os
from copy import deepcopy

import torch

from .checkpoints import diffusers_unet_to_k2
from .configs import CONFIG_2_1, CONFIG_2_2
from .pipelines import Kandinsky2_1, Kandinsky2_2


def _maybe_load(path):
    return torch.load(path, map_location="cpu") if os.path.exists(path) else None


def get_kandinsky2(device, task_type="text2img", cache_dir="/tmp/kandinsky2", use_auth_token=None,
                   model_version="2.1", use_flash_attention=False, embedder=None, config_overrides=None):
    if model_version == "2.0":
        raise NotImplementedError("Kandinsky 2.0 is outside the hot path of this build (SURVEY.md section 2 rows 17-18)")
    if model_version not in ("2.1", "2.2"):
        raise ValueError("Only 2.0, 2.1 and 2.2 are available")
    if task_type == "controlnet" and model_version != "2.2":
        raise ValueError("task_type='controlnet' (Kandinsky 2.2 ControlNet-depth) needs model_version='2.2'")
    if task_type not in ("text2img", "img2img", "inpainting", "controlnet"):  # "controlnet": extension, see pipelines.py
        raise ValueError("Only text2img, img2img, inpainting is available")
    config = deepcopy(CONFIG_2_1 if model_version == "2.1" else CONFIG_2_2)
    for k, v in (config_overrides or {}).items():
        config[k].update(v)
    sub = os.path.join(cache_dir, "2_1" if model_version == "2.1" else "2_2")
    unet_sd = _maybe_load(os.path.join(sub, "inpainting_fp16.ckpt" if task_type == "inpainting" else "decoder_fp16.ckpt"))
    if unet_sd is not None and any(k.startswith(("down_blocks.", "mid_block.", "up_blocks.")) for k in unet_sd):
        # Kandinsky 2.2 ships its decoder as a diffusers UNet2DConditionModel state dict (kandinsky2_2_model.py:26-28): rename
        # and re-pack it into this package's (= the reference 2.1 backbone's) key layout
        mc = config["model_config"]
        unet_sd = diffusers_unet_to_k2(
            unet_sd, in_channels=9 if task_type == "inpainting" else 4, model_channels=mc["num_channels"],
            channel_mult=tuple(int(v) for v in mc["channel_mult"].split(",")) if mc.get("channel_mult") else (1, 2, 3, 4),
            num_res_blocks=mc["num_res_blocks"],
            attention_ds=tuple(mc["image_size"] // int(r) for r in mc["attention_resolutions"].split(",")))
    movq_sd = _maybe_load(os.path.join(sub, "movq_final.ckpt"))
    cls = Kandinsky2_1 if model_version == "2.1" else Kandinsky2_2
    return cls(config, device, task_type=task_type, embedder=embedder, unet_state_dict=unet_sd, movq_state_dict=movq_sd)
