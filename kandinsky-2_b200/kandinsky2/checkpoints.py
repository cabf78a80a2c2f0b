"""Checkpoint / wire formats -- pure host code.

* Kandinsky 2.1 files (`decoder_fp16.ckpt`, `movq_final.ckpt`; `kandinsky2_1_model.py:82-91`) are plain state dicts with
  the key names this package's modules already use: `Text2ImUNet.load_state_dict` / `MOVQ.load_state_dict` take them as
  they are.
* Kandinsky 2.2 ships the decoder as a diffusers `UNet2DConditionModel` (`kandinsky2_2_model.py:26-28`,
  `kandinsky-community/kandinsky-2-2-decoder`, subfolder `unet`).  It is the same backbone with a
  different parameter naming and separate q / k / v (and added k / v) Linear layers where the reference has one
  head-interleaved `qkv` (and `encoder_kv`) Conv1d.  `diffusers_unet_to_k2` renames and re-packs such a state dict into
  the key layout of `Text2ImUNet(cond_version="2.2")`; `k2_to_diffusers_unet` is its inverse.

**Parity unpinned:** diffusers is not part of /root/reference (`setup.py:27` lists it unpinned) and is not installed in the
build container, so the diffusers-side key names below are restated from the published `UNet2DConditionModel` layout
(`ResnetDownsampleBlock2D` / `SimpleCrossAttnDownBlock2D` / `UNetMidBlock2DSimpleCrossAttn` / `SimpleCrossAttnUpBlock2D` /
`ResnetUpsampleBlock2D`, attention with `added_kv_proj_dim`).  What IS tested (tests/test_cpu_boundary.py): the two maps
are inverse bijections onto the package's exact key set, and the head-interleaved packing reproduces separate
q / k / v projections numerically.
"""
import torch

from .model.unet import _topology

_RES = {"norm1": "in_layers.0", "conv1": "in_layers.2", "time_emb_proj": "emb_layers.1", "norm2": "out_layers.0",
        "conv2": "out_layers.3", "conv_shortcut": "skip_connection"}
_TOP = {"time_embedding.linear_1": "time_embed.0", "time_embedding.linear_2": "time_embed.2", "conv_in": "input_blocks.0.0",
        "conv_norm_out": "out.0", "conv_out": "out.2",
        "add_embedding.image_proj": "add_embedding.image_proj", "add_embedding.image_norm": "add_embedding.image_norm",
        "encoder_hid_proj.image_embeds": "encoder_hid_proj.image_embeds", "encoder_hid_proj.norm": "encoder_hid_proj.norm"}


def unet_block_map(in_channels, model_channels, channel_mult, num_res_blocks, attention_ds):
    """[(diffusers prefix, k2 prefix, kind)] for every ResBlock ('res') and attention block ('attn') of the UNet, in the
    reference's block order (`unet.py:421-557`): input_blocks.i.{0,1}, middle_block.{0,1,2}, output_blocks.i.{0,1,2}."""
    inp, mid, out = _topology(in_channels, model_channels, channel_mult, num_res_blocks, attention_ds)
    pairs = []
    level, j = 0, 0
    for i, blk in enumerate(inp[1:], start=1):
        if blk[0][0] == "res" and blk[0][3] == "down":
            pairs.append((f"down_blocks.{level}.downsamplers.0", f"input_blocks.{i}.0", "res"))
            level, j = level + 1, 0
            continue
        pairs.append((f"down_blocks.{level}.resnets.{j}", f"input_blocks.{i}.0", "res"))
        if len(blk) > 1:
            pairs.append((f"down_blocks.{level}.attentions.{j}", f"input_blocks.{i}.1", "attn"))
        j += 1
    pairs += [("mid_block.resnets.0", "middle_block.0", "res"), ("mid_block.attentions.0", "middle_block.1", "attn"),
              ("mid_block.resnets.1", "middle_block.2", "res")]
    level, j = 0, 0
    for i, blk in enumerate(out):
        pairs.append((f"up_blocks.{level}.resnets.{j}", f"output_blocks.{i}.0", "res"))
        pos = 1
        if len(blk) > 1 and blk[1][0] == "attn":
            pairs.append((f"up_blocks.{level}.attentions.{j}", f"output_blocks.{i}.1", "attn"))
            pos = 2
        j += 1
        if blk[-1][0] == "res" and blk[-1][3] == "up":
            pairs.append((f"up_blocks.{level}.upsamplers.0", f"output_blocks.{i}.{pos}", "res"))
            level, j = level + 1, 0
    return pairs


def pack_heads(parts, head_dim=64):
    """[W_a, W_b, ...] each [C, K] (rows = output channels, head-major) -> [len(parts)*C, K] with the reference's per-head
    interleave: head h owns rows [h*n*d, (h+1)*n*d) = [a_h | b_h | ...]  (`unet.py:296-299,304-307`: the Conv1d output is
    viewed as (heads, n*d, T) and split along dim 1)."""
    C = parts[0].shape[0]
    heads = C // head_dim
    stacked = torch.stack([p.reshape(heads, head_dim, *p.shape[1:]) for p in parts], dim=1)  # [heads, n, d, ...]
    return stacked.reshape(len(parts) * C, *parts[0].shape[1:])


def unpack_heads(w, n, head_dim=64):
    """Inverse of pack_heads: [n*C, ...] -> n tensors [C, ...]."""
    C = w.shape[0] // n
    heads = C // head_dim
    v = w.reshape(heads, n, head_dim, *w.shape[1:])
    return [v[:, i].reshape(C, *w.shape[1:]) for i in range(n)]


def diffusers_unet_to_k2(sd, in_channels=4, model_channels=384, channel_mult=(1, 2, 3, 4), num_res_blocks=3,
                         attention_ds=(2, 4, 8), head_dim=64):
    """diffusers `UNet2DConditionModel` state dict (Kandinsky 2.2 decoder) -> `Text2ImUNet(cond_version="2.2")` keys.
    Linear q/k/v/out weights [C, C] become k=1 Conv1d weights [.., C, 1]; q|k|v and add_k|add_v are head-interleaved."""
    out = {k: v for k, v in sd.items() if k.startswith("add_embedding.input_hint_block.")}  # ControlNet hint stem: same names
    for d, k in _TOP.items():
        for suffix in ("weight", "bias"):
            if f"{d}.{suffix}" in sd:
                out[f"{k}.{suffix}"] = sd[f"{d}.{suffix}"]
    for dp, kp, kind in unet_block_map(in_channels, model_channels, channel_mult, num_res_blocks, attention_ds):
        if kind == "res":
            for dn, kn in _RES.items():
                for suffix in ("weight", "bias"):
                    key = f"{dp}.{dn}.{suffix}"
                    if key in sd:  # conv_shortcut only where the channel count changes
                        out[f"{kp}.{kn}.{suffix}"] = sd[key]
        else:
            out[f"{kp}.norm.weight"] = sd[f"{dp}.group_norm.weight"]
            out[f"{kp}.norm.bias"] = sd[f"{dp}.group_norm.bias"]
            q, k_, v = (sd[f"{dp}.to_{n}.weight"] for n in "qkv")
            out[f"{kp}.qkv.weight"] = pack_heads([q, k_, v], head_dim).unsqueeze(-1)
            out[f"{kp}.qkv.bias"] = pack_heads([sd[f"{dp}.to_{n}.bias"] for n in "qkv"], head_dim)
            ak, av = sd[f"{dp}.add_k_proj.weight"], sd[f"{dp}.add_v_proj.weight"]
            out[f"{kp}.encoder_kv.weight"] = pack_heads([ak, av], head_dim).unsqueeze(-1)
            out[f"{kp}.encoder_kv.bias"] = pack_heads([sd[f"{dp}.add_k_proj.bias"], sd[f"{dp}.add_v_proj.bias"]], head_dim)
            out[f"{kp}.proj_out.weight"] = sd[f"{dp}.to_out.0.weight"].unsqueeze(-1)
            out[f"{kp}.proj_out.bias"] = sd[f"{dp}.to_out.0.bias"]
    return out


def k2_to_diffusers_unet(sd, in_channels=4, model_channels=384, channel_mult=(1, 2, 3, 4), num_res_blocks=3,
                         attention_ds=(2, 4, 8), head_dim=64):
    """Inverse of diffusers_unet_to_k2 (export, and the round-trip test)."""
    out = {k: v for k, v in sd.items() if k.startswith("add_embedding.input_hint_block.")}
    for d, k in _TOP.items():
        for suffix in ("weight", "bias"):
            if f"{k}.{suffix}" in sd:
                out[f"{d}.{suffix}"] = sd[f"{k}.{suffix}"]
    for dp, kp, kind in unet_block_map(in_channels, model_channels, channel_mult, num_res_blocks, attention_ds):
        if kind == "res":
            for dn, kn in _RES.items():
                for suffix in ("weight", "bias"):
                    key = f"{kp}.{kn}.{suffix}"
                    if key in sd:
                        out[f"{dp}.{dn}.{suffix}"] = sd[key]
        else:
            out[f"{dp}.group_norm.weight"] = sd[f"{kp}.norm.weight"]
            out[f"{dp}.group_norm.bias"] = sd[f"{kp}.norm.bias"]
            for n, w, b in zip("qkv", unpack_heads(sd[f"{kp}.qkv.weight"].squeeze(-1), 3, head_dim),
                               unpack_heads(sd[f"{kp}.qkv.bias"], 3, head_dim)):
                out[f"{dp}.to_{n}.weight"], out[f"{dp}.to_{n}.bias"] = w, b
            for n, w, b in zip(("add_k_proj", "add_v_proj"), unpack_heads(sd[f"{kp}.encoder_kv.weight"].squeeze(-1), 2, head_dim),
                               unpack_heads(sd[f"{kp}.encoder_kv.bias"], 2, head_dim)):
                out[f"{dp}.{n}.weight"], out[f"{dp}.{n}.bias"] = w, b
            out[f"{dp}.to_out.0.weight"] = sd[f"{kp}.proj_out.weight"].squeeze(-1)
            out[f"{dp}.to_out.0.bias"] = sd[f"{kp}.proj_out.bias"]
    return out
