"""Oracle: torch-fp32 restatement of MOVQ.decode (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Follows /root/reference/kandinsky2/vqgan:
  MOVQ.decode                 autoencoder.py:182-185  (post_quant_conv 1x1, then MOVQDecoder(quant2, quant))
  MOVQDecoder.__init__/forward movq_modules.py:228-357
  SpatialNorm.forward          movq_modules.py:61-68   (GroupNorm(32, eps 1e-6) * conv_y(zq) + conv_b(zq), zq nearest-resized)
  ResnetBlock.forward          movq_modules.py:159-179 (temb is None: temb_ch == 0)
  AttnBlock.forward            movq_modules.py:201-225 (single head, scale C^-0.5)
  Upsample.forward             movq_modules.py:93-97   (nearest 2x + conv3x3)
  VectorQuantizer distances    quntize.py:89-98        (argmin -> code indices)
Keys are the reference state_dict's (decoder.*, post_quant_conv.*, quantize.embedding.weight).
"""
import torch
import torch.nn.functional as F

DDCONFIG_2_1 = dict(z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=(1, 2, 2, 4),
                    num_res_blocks=2, attn_resolutions=(32,))  # configs.py:74-86
DDCONFIG_TINY = dict(z_channels=4, resolution=32, in_channels=3, out_ch=3, ch=32, ch_mult=(1, 2),
                     num_res_blocks=1, attn_resolutions=(16,))


def decoder_topology(dd):
    """-> (block_in at the lowest level, [levels from lowest res upward: dict(blocks=[(cin,cout)], attn=bool, up=bool)])"""
    ch, mult, nrb = dd["ch"], tuple(dd["ch_mult"]), dd["num_res_blocks"]
    nres = len(mult)
    block_in = ch * mult[-1]
    curr = dd["resolution"] // 2 ** (nres - 1)
    levels = []
    bi = block_in
    for i_level in reversed(range(nres)):
        bo = ch * mult[i_level]
        blocks = []
        for _ in range(nrb + 1):
            blocks.append((bi, bo))
            bi = bo
        levels.append(dict(level=i_level, blocks=blocks, attn=curr in dd["attn_resolutions"], up=i_level != 0, ch=bo))
        if i_level != 0:
            curr *= 2
    return block_in, levels


def _sn_spec(p, c, zc, spec):
    spec += [(p + "norm_layer.weight", (c,)), (p + "norm_layer.bias", (c,)),
             (p + "conv_y.weight", (c, zc, 1, 1)), (p + "conv_y.bias", (c,)),
             (p + "conv_b.weight", (c, zc, 1, 1)), (p + "conv_b.bias", (c,))]


def _res_spec(p, cin, cout, zc, spec):
    _sn_spec(p + "norm1.", cin, zc, spec)
    spec += [(p + "conv1.weight", (cout, cin, 3, 3)), (p + "conv1.bias", (cout,))]
    _sn_spec(p + "norm2.", cout, zc, spec)
    spec += [(p + "conv2.weight", (cout, cout, 3, 3)), (p + "conv2.bias", (cout,))]
    if cin != cout:
        spec += [(p + "nin_shortcut.weight", (cout, cin, 1, 1)), (p + "nin_shortcut.bias", (cout,))]


def _attn_spec(p, c, zc, spec):
    _sn_spec(p + "norm.", c, zc, spec)
    for n in ("q", "k", "v", "proj_out"):
        spec += [(p + n + ".weight", (c, c, 1, 1)), (p + n + ".bias", (c,))]


def movq_decoder_param_spec(dd, embed_dim=4, n_embed=None):
    """Decoder-side keys of the reference MOVQ state dict, in registration order."""
    zc = embed_dim
    block_in, levels = decoder_topology(dd)
    spec = [("decoder.conv_in.weight", (block_in, dd["z_channels"], 3, 3)), ("decoder.conv_in.bias", (block_in,))]
    _res_spec("decoder.mid.block_1.", block_in, block_in, zc, spec)
    _attn_spec("decoder.mid.attn_1.", block_in, zc, spec)
    _res_spec("decoder.mid.block_2.", block_in, block_in, zc, spec)
    # self.up is built lowest level first but stored with insert(0): state_dict order is up.0 (highest res) first
    for lv in sorted(levels, key=lambda l: l["level"]):
        p = f"decoder.up.{lv['level']}."
        for bi, (cin, cout) in enumerate(lv["blocks"]):
            _res_spec(p + f"block.{bi}.", cin, cout, zc, spec)
        if lv["attn"]:
            for bi in range(len(lv["blocks"])):
                _attn_spec(p + f"attn.{bi}.", lv["ch"], zc, spec)
        if lv["up"]:
            spec += [(p + "upsample.conv.weight", (lv["ch"], lv["ch"], 3, 3)), (p + "upsample.conv.bias", (lv["ch"],))]
    c_last = levels[-1]["ch"]
    _sn_spec("decoder.norm_out.", c_last, zc, spec)
    spec += [("decoder.conv_out.weight", (dd["out_ch"], c_last, 3, 3)), ("decoder.conv_out.bias", (dd["out_ch"],))]
    if n_embed:
        spec += [("quantize.embedding.weight", (n_embed, embed_dim))]
    spec += [("post_quant_conv.weight", (dd["z_channels"], embed_dim, 1, 1)), ("post_quant_conv.bias", (dd["z_channels"],))]
    return spec


# ---------------------------------------------------------------------------------------------
# encoder (image -> latent): vqgan_blocks.py:253-367 (Encoder), :87-91 (Normalize: GroupNorm 32, eps 1e-6),
# :109-126 (Downsample: pad (0,1,0,1) + conv3x3 stride 2), :129-193 (ResnetBlock), :196-239 (AttnBlock);
# MOVQ.encode = quant_conv(encoder(x)) without quantisation (autoencoder.py:176-180)
# ---------------------------------------------------------------------------------------------
def encoder_topology(dd):
    ch, mult, nrb = dd["ch"], tuple(dd["ch_mult"]), dd["num_res_blocks"]
    curr = dd["resolution"]
    in_mult = (1,) + mult
    levels = []
    for i in range(len(mult)):
        bi, bo = ch * in_mult[i], ch * mult[i]
        blocks = []
        for _ in range(nrb):
            blocks.append((bi, bo))
            bi = bo
        levels.append(dict(level=i, blocks=blocks, attn=curr in tuple(dd["attn_resolutions"]), down=i != len(mult) - 1, ch=bo))
        if i != len(mult) - 1:
            curr //= 2
    return levels


def _enc_res_spec(p, cin, cout, spec):
    spec += [(p + "norm1.weight", (cin,)), (p + "norm1.bias", (cin,)),
             (p + "conv1.weight", (cout, cin, 3, 3)), (p + "conv1.bias", (cout,)),
             (p + "norm2.weight", (cout,)), (p + "norm2.bias", (cout,)),
             (p + "conv2.weight", (cout, cout, 3, 3)), (p + "conv2.bias", (cout,))]
    if cin != cout:
        spec += [(p + "nin_shortcut.weight", (cout, cin, 1, 1)), (p + "nin_shortcut.bias", (cout,))]


def _enc_attn_spec(p, c, spec):
    spec += [(p + "norm.weight", (c,)), (p + "norm.bias", (c,))]
    for n in ("q", "k", "v", "proj_out"):
        spec += [(p + n + ".weight", (c, c, 1, 1)), (p + n + ".bias", (c,))]


def movq_encoder_param_spec(dd, embed_dim=4):
    spec = [("encoder.conv_in.weight", (dd["ch"], dd["in_channels"], 3, 3)), ("encoder.conv_in.bias", (dd["ch"],))]
    levels = encoder_topology(dd)
    for lv in levels:
        p = f"encoder.down.{lv['level']}."
        for bi, (cin, cout) in enumerate(lv["blocks"]):
            _enc_res_spec(p + f"block.{bi}.", cin, cout, spec)
        if lv["attn"]:
            for bi in range(len(lv["blocks"])):
                _enc_attn_spec(p + f"attn.{bi}.", lv["ch"], spec)
        if lv["down"]:
            spec += [(p + "downsample.conv.weight", (lv["ch"], lv["ch"], 3, 3)), (p + "downsample.conv.bias", (lv["ch"],))]
    c = levels[-1]["ch"]
    _enc_res_spec("encoder.mid.block_1.", c, c, spec)
    _enc_attn_spec("encoder.mid.attn_1.", c, spec)
    _enc_res_spec("encoder.mid.block_2.", c, c, spec)
    zc = dd["z_channels"] * (2 if dd.get("double_z") else 1)
    spec += [("encoder.norm_out.weight", (c,)), ("encoder.norm_out.bias", (c,)),
             ("encoder.conv_out.weight", (zc, c, 3, 3)), ("encoder.conv_out.bias", (zc,)),
             ("quant_conv.weight", (embed_dim, dd["z_channels"], 1, 1)), ("quant_conv.bias", (embed_dim,))]
    return spec


def _gn6(x, sd, p):
    return F.group_norm(x, 32, sd[p + "weight"], sd[p + "bias"], 1e-6)


def _enc_res(x, sd, p):
    h = F.conv2d(_swish(_gn6(x, sd, p + "norm1.")), sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1)
    h = F.conv2d(_swish(_gn6(h, sd, p + "norm2.")), sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1)
    if (p + "nin_shortcut.weight") in sd:
        x = F.conv2d(x, sd[p + "nin_shortcut.weight"], sd[p + "nin_shortcut.bias"])
    return x + h


def _enc_attn(x, sd, p):
    h = _gn6(x, sd, p + "norm.")
    q = F.conv2d(h, sd[p + "q.weight"], sd[p + "q.bias"])
    k = F.conv2d(h, sd[p + "k.weight"], sd[p + "k.bias"])
    v = F.conv2d(h, sd[p + "v.weight"], sd[p + "v.bias"])
    b, c, hh, ww = q.shape
    w_ = torch.bmm(q.reshape(b, c, -1).permute(0, 2, 1), k.reshape(b, c, -1)) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    h = torch.bmm(v.reshape(b, c, -1), w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + F.conv2d(h, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])


def movq_encode(sd, dd, x):
    """image fp32 [B, 3, H, W] in [-1, 1] -> latent [B, embed_dim, H/2^(levels-1), ...] (no quantisation)."""
    h = F.conv2d(x, sd["encoder.conv_in.weight"], sd["encoder.conv_in.bias"], padding=1)
    for lv in encoder_topology(dd):
        p = f"encoder.down.{lv['level']}."
        for bi in range(len(lv["blocks"])):
            h = _enc_res(h, sd, p + f"block.{bi}.")
            if lv["attn"]:
                h = _enc_attn(h, sd, p + f"attn.{bi}.")
        if lv["down"]:
            h = F.conv2d(F.pad(h, (0, 1, 0, 1)), sd[p + "downsample.conv.weight"], sd[p + "downsample.conv.bias"], stride=2)
    h = _enc_res(h, sd, "encoder.mid.block_1.")
    h = _enc_attn(h, sd, "encoder.mid.attn_1.")
    h = _enc_res(h, sd, "encoder.mid.block_2.")
    h = F.conv2d(_swish(_gn6(h, sd, "encoder.norm_out.")), sd["encoder.conv_out.weight"], sd["encoder.conv_out.bias"], padding=1)
    return F.conv2d(h, sd["quant_conv.weight"], sd["quant_conv.bias"])


def movq_param_spec(dd, embed_dim=4, n_embed=None):
    """Full MOVQ state dict in the reference's registration order (autoencoder.py:167-174)."""
    enc = movq_encoder_param_spec(dd, embed_dim)
    dec = movq_decoder_param_spec(dd, embed_dim, n_embed)
    qc = [e for e in enc if e[0].startswith("quant_conv.")]
    enc = [e for e in enc if not e[0].startswith("quant_conv.")]
    pq = [e for e in dec if e[0].startswith("post_quant_conv.")]
    dec = [e for e in dec if not e[0].startswith("post_quant_conv.")]
    return enc + dec + qc + pq


def _sn(f, zq, sd, p):
    z = F.interpolate(zq, size=f.shape[-2:], mode="nearest")
    nf = F.group_norm(f, 32, sd[p + "norm_layer.weight"], sd[p + "norm_layer.bias"], 1e-6)
    return nf * F.conv2d(z, sd[p + "conv_y.weight"], sd[p + "conv_y.bias"]) + F.conv2d(z, sd[p + "conv_b.weight"], sd[p + "conv_b.bias"])


def _swish(x):
    return x * torch.sigmoid(x)


def _res(x, zq, sd, p):
    h = F.conv2d(_swish(_sn(x, zq, sd, p + "norm1.")), sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1)
    h = F.conv2d(_swish(_sn(h, zq, sd, p + "norm2.")), sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1)
    if (p + "nin_shortcut.weight") in sd:
        x = F.conv2d(x, sd[p + "nin_shortcut.weight"], sd[p + "nin_shortcut.bias"])
    return x + h


def _attn(x, zq, sd, p):
    h = _sn(x, zq, sd, p + "norm.")
    q = F.conv2d(h, sd[p + "q.weight"], sd[p + "q.bias"])
    k = F.conv2d(h, sd[p + "k.weight"], sd[p + "k.bias"])
    v = F.conv2d(h, sd[p + "v.weight"], sd[p + "v.bias"])
    b, c, hh, ww = q.shape
    w_ = torch.bmm(q.reshape(b, c, -1).permute(0, 2, 1), k.reshape(b, c, -1)) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    h = torch.bmm(v.reshape(b, c, -1), w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + F.conv2d(h, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])


def movq_decode(sd, dd, quant, taps=None):
    """quant fp32 [B, 4, h, w] -> image fp32 [B, 3, 8h, 8w] (for the 4-level config)."""
    zq = quant
    h = F.conv2d(quant, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    h = F.conv2d(h, sd["decoder.conv_in.weight"], sd["decoder.conv_in.bias"], padding=1)
    h = _res(h, zq, sd, "decoder.mid.block_1.")
    h = _attn(h, zq, sd, "decoder.mid.attn_1.")
    h = _res(h, zq, sd, "decoder.mid.block_2.")
    if taps is not None:
        taps["mid"] = h
    _, levels = decoder_topology(dd)
    for lv in levels:
        p = f"decoder.up.{lv['level']}."
        for bi in range(len(lv["blocks"])):
            h = _res(h, zq, sd, p + f"block.{bi}.")
            if lv["attn"]:
                h = _attn(h, zq, sd, p + f"attn.{bi}.")
        if lv["up"]:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, sd[p + "upsample.conv.weight"], sd[p + "upsample.conv.bias"], padding=1)
        if taps is not None:
            taps[f"up{lv['level']}"] = h
    h = _swish(_sn(h, zq, sd, "decoder.norm_out."))
    return F.conv2d(h, sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"], padding=1)


def vq_indices(z, codebook):
    """quntize.py:89-98 on z [n, e_dim]: argmin_j ||z||^2 + ||e_j||^2 - 2 z.e_j  -> int64 [n]."""
    d = torch.sum(z ** 2, dim=1, keepdim=True) + torch.sum(codebook ** 2, dim=1) - 2 * torch.einsum("bd,dn->bn", z, codebook.t())
    return torch.argmin(d, dim=1)


def process_images(x):
    """utils.py:57-70 up to the uint8 tensor: ((x+1)*127.5).round().clamp(0,255) -> uint8 NHWC."""
    return ((x + 1) * 127.5).round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()


def decode_flops(dd, B, h, w):
    """2*MAC of one decode (convs + attention GEMMs)."""
    block_in, levels = decoder_topology(dd)
    f = 2 * B * h * w * (dd["z_channels"] * 4 + block_in * dd["z_channels"] * 9)

    def res(cin, cout, hh, ww):
        return 2 * B * hh * ww * cout * (9 * cin + 9 * cout + (cin if cin != cout else 0))

    def attn(c, hh, ww):
        T = hh * ww
        return 2 * B * T * c * 4 * c + 4 * B * T * T * c

    def sn(c, hh, ww):
        return 2 * B * hh * ww * c * 4 * 2

    f += res(block_in, block_in, h, w) * 2 + attn(block_in, h, w) + sn(block_in, h, w) * 5
    hh, ww = h, w
    for lv in levels:
        for cin, cout in lv["blocks"]:
            f += res(cin, cout, hh, ww) + sn(cin, hh, ww) + sn(cout, hh, ww)
            if lv["attn"]:
                f += attn(cout, hh, ww) + sn(cout, hh, ww)
        if lv["up"]:
            hh, ww = hh * 2, ww * 2
            f += 2 * B * hh * ww * lv["ch"] * lv["ch"] * 9
    c_last = levels[-1]["ch"]
    f += sn(c_last, hh, ww) + 2 * B * hh * ww * dd["out_ch"] * c_last * 9
    return f
