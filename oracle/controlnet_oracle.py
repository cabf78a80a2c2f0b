"""Oracle: Kandinsky 2.2 ControlNet-depth denoiser (TEST INFRASTRUCTURE -- see oracle/__init__.py).

PARITY UNPINNED.  BASELINE.json configs[4] ("ControlNet-depth 768x768") has no entry point in the reference package: the
reference shows it only in notebooks/kandinsky2_2_controlnet.ipynb, which drives diffusers' KandinskyV22ControlnetPipeline
directly with `kandinsky-community/kandinsky-2-2-controlnet-depth`.  diffusers is not in /root/reference (setup.py:27), so this
restates the published algorithm:

  * the UNet is the Kandinsky 2.2 decoder backbone (= the pinned 2.1 backbone, oracle/unet_oracle.py) with in_channels = 8 and
    `addition_embed_type="image_hint"` (diffusers ImageHintTimeEmbedding):
        time_image_embeds = LayerNorm(Linear(image_embeds))                          -> added to the time embedding
        hint_features     = input_hint_block(hint)     hint [N, 3, 8h, 8w] in [0, 1] -> [N, 4, h, w]
        input_hint_block  = Conv3x3(3,16) SiLU Conv3x3(16,16) SiLU Conv3x3(16,32,stride 2) SiLU Conv3x3(32,32) SiLU
                            Conv3x3(32,96,stride 2) SiLU Conv3x3(96,96) SiLU Conv3x3(96,256,stride 2) SiLU Conv3x3(256,4)
        sample            = cat([sample, hint_features], dim=1)                      -> conv_in sees 8 channels
  * encoder_hid_proj (ImageProjection) and everything after conv_in are those of the plain 2.2 decoder.
"""
import torch
import torch.nn.functional as F

from . import unet_oracle as uo

HINT_CHANNELS = [(3, 16, 1), (16, 16, 1), (16, 32, 2), (32, 32, 1), (32, 96, 2), (96, 96, 1), (96, 256, 2), (256, 4, 1)]

CONFIG_2_2_HINT = dict(uo.CONFIG_2_2, in_channels=8)


def hint_param_spec():
    """[(key, shape)] of add_embedding.input_hint_block (nn.Sequential indices 0, 2, 4, ... : convs; odd: SiLU)."""
    spec = []
    for i, (cin, cout, _) in enumerate(HINT_CHANNELS):
        spec += [(f"add_embedding.input_hint_block.{2 * i}.weight", (cout, cin, 3, 3)),
                 (f"add_embedding.input_hint_block.{2 * i}.bias", (cout,))]
    return spec


def param_spec(cfg):
    return uo.unet_param_spec(cfg) + hint_param_spec()


def hint_features(sd, hint):
    h = hint
    for i, (_, _, stride) in enumerate(HINT_CHANNELS):
        h = F.conv2d(h, sd[f"add_embedding.input_hint_block.{2 * i}.weight"], sd[f"add_embedding.input_hint_block.{2 * i}.bias"],
                     stride=stride, padding=1)
        if i + 1 < len(HINT_CHANNELS):
            h = F.silu(h)
    return h


def unet_forward(sd, cfg, x, timesteps, image_emb, hint):
    """x [N, 4, h, w], hint [N, 3, 8h, 8w] -> [N, 8, h, w]: the 2.2 backbone on cat([x, input_hint_block(hint)])."""
    return uo.unet_forward(sd, cfg, torch.cat([x, hint_features(sd, hint).to(x.dtype)], dim=1), timesteps, image_emb=image_emb)
