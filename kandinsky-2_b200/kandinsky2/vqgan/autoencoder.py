"""B200-native MOVQ: `decode` (latents -> image) through the C-ABI kernels; reference module boundary.

Drop-in for kandinsky2/vqgan/autoencoder.py:160-201 (class MOVQ; ctor (ddconfig, n_embed, embed_dim); decode :182-185)
with the decoder of kandinsky2/vqgan/movq_modules.py:228-357.  state_dict keys/shapes equal the reference's for
`decoder.*`, `post_quant_conv.*` and `quantize.embedding.weight` (the encoder / quant_conv halves of a reference
checkpoint are accepted and ignored: the image->latent direction is SURVEY.md 8f rank 1, not on this path).

Kernel program (activations NHWC fp16, fp32 accumulate):
  SpatialNorm + swish      gn_stats + gn_apply with the 4-channel latent modulation computed on the fly
                           (movq_modules.py:61-68: no [B,C,H,W] conv_y/conv_b tensors, no interpolate)
  conv3x3 / nin_shortcut   k2_conv_gemm (shortcut folded in as extra K segment / epilogue residual)
  AttnBlock (1 head, d=C)  q,k in one GEMM; scores = q k^T with the k rows as a strided B operand; row softmax;
                           P V with V^T produced directly by a GEMM (W_v as the A operand); v bias added after PV
                           (softmax rows sum to 1); proj GEMM + residual   (movq_modules.py:201-225)
  Upsample                 nearest 2x copy kernel + conv3x3 (movq_modules.py:93-97)
"""
import torch
import torch.nn as nn

from .. import ops
from .._native import K2Error


class _Node(nn.Module):
    pass


def _enc_topology(dd):
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


def _topology(dd):
    ch, mult, nrb = dd["ch"], tuple(dd["ch_mult"]), dd["num_res_blocks"]
    nres = len(mult)
    block_in = ch * mult[-1]
    curr = dd["resolution"] // 2 ** (nres - 1)
    levels, bi = [], block_in
    for lvl in reversed(range(nres)):
        bo = ch * mult[lvl]
        blocks = []
        for _ in range(nrb + 1):
            blocks.append((bi, bo))
            bi = bo
        levels.append(dict(level=lvl, blocks=blocks, attn=curr in tuple(dd["attn_resolutions"]), up=lvl != 0, ch=bo))
        if lvl != 0:
            curr *= 2
    return block_in, levels


class MOVQ(nn.Module):
    def __init__(self, ddconfig, n_embed, embed_dim, device=None, param_dtype=torch.float32):
        super().__init__()
        self.ddconfig = dict(ddconfig)
        self.n_embed, self.embed_dim = n_embed, embed_dim
        self._packed = None
        dd = self.ddconfig
        kw = dict(device=device, dtype=param_dtype)
        zc = embed_dim

        def P(path, *shape):
            node = self
            parts = path.split(".")
            for name in parts[:-1]:
                if name not in node._modules:
                    node.add_module(name, _Node())
                node = node._modules[name]
            node.register_parameter(parts[-1], nn.Parameter(torch.zeros(*shape, **kw), requires_grad=False))

        def SN(p, c):
            P(p + "norm_layer.weight", c); P(p + "norm_layer.bias", c)
            P(p + "conv_y.weight", c, zc, 1, 1); P(p + "conv_y.bias", c)
            P(p + "conv_b.weight", c, zc, 1, 1); P(p + "conv_b.bias", c)

        def RES(p, cin, cout):
            SN(p + "norm1.", cin)
            P(p + "conv1.weight", cout, cin, 3, 3); P(p + "conv1.bias", cout)
            SN(p + "norm2.", cout)
            P(p + "conv2.weight", cout, cout, 3, 3); P(p + "conv2.bias", cout)
            if cin != cout:
                P(p + "nin_shortcut.weight", cout, cin, 1, 1); P(p + "nin_shortcut.bias", cout)

        def ATT(p, c):
            SN(p + "norm.", c)
            for n in ("q", "k", "v", "proj_out"):
                P(p + n + ".weight", c, c, 1, 1); P(p + n + ".bias", c)

        # ---- encoder (image -> latent; vqgan_blocks.py:253-367) -- plain GroupNorm(32, eps 1e-6), no SpatialNorm
        def ERES(p, cin, cout):
            P(p + "norm1.weight", cin); P(p + "norm1.bias", cin)
            P(p + "conv1.weight", cout, cin, 3, 3); P(p + "conv1.bias", cout)
            P(p + "norm2.weight", cout); P(p + "norm2.bias", cout)
            P(p + "conv2.weight", cout, cout, 3, 3); P(p + "conv2.bias", cout)
            if cin != cout:
                P(p + "nin_shortcut.weight", cout, cin, 1, 1); P(p + "nin_shortcut.bias", cout)

        def EATT(p, c):
            P(p + "norm.weight", c); P(p + "norm.bias", c)
            for n in ("q", "k", "v", "proj_out"):
                P(p + n + ".weight", c, c, 1, 1); P(p + n + ".bias", c)

        self.enc_levels = _enc_topology(dd)
        P("encoder.conv_in.weight", dd["ch"], dd["in_channels"], 3, 3); P("encoder.conv_in.bias", dd["ch"])
        for lv in self.enc_levels:
            p = f"encoder.down.{lv['level']}."
            for bi, (cin, cout) in enumerate(lv["blocks"]):
                ERES(p + f"block.{bi}.", cin, cout)
            if lv["attn"]:
                for bi in range(len(lv["blocks"])):
                    EATT(p + f"attn.{bi}.", lv["ch"])
            if lv["down"]:
                P(p + "downsample.conv.weight", lv["ch"], lv["ch"], 3, 3); P(p + "downsample.conv.bias", lv["ch"])
        ce = self.enc_levels[-1]["ch"]
        ERES("encoder.mid.block_1.", ce, ce)
        EATT("encoder.mid.attn_1.", ce)
        ERES("encoder.mid.block_2.", ce, ce)
        zc_out = dd["z_channels"] * (2 if dd.get("double_z") else 1)
        P("encoder.norm_out.weight", ce); P("encoder.norm_out.bias", ce)
        P("encoder.conv_out.weight", zc_out, ce, 3, 3); P("encoder.conv_out.bias", zc_out)

        self.block_in, self.levels = _topology(dd)
        P("decoder.conv_in.weight", self.block_in, dd["z_channels"], 3, 3); P("decoder.conv_in.bias", self.block_in)
        RES("decoder.mid.block_1.", self.block_in, self.block_in)
        ATT("decoder.mid.attn_1.", self.block_in)
        RES("decoder.mid.block_2.", self.block_in, self.block_in)
        for lv in sorted(self.levels, key=lambda l: l["level"]):
            p = f"decoder.up.{lv['level']}."
            for bi, (cin, cout) in enumerate(lv["blocks"]):
                RES(p + f"block.{bi}.", cin, cout)
            if lv["attn"]:
                for bi in range(len(lv["blocks"])):
                    ATT(p + f"attn.{bi}.", lv["ch"])
            if lv["up"]:
                P(p + "upsample.conv.weight", lv["ch"], lv["ch"], 3, 3); P(p + "upsample.conv.bias", lv["ch"])
        c_last = self.levels[-1]["ch"]
        SN("decoder.norm_out.", c_last)
        P("decoder.conv_out.weight", dd["out_ch"], c_last, 3, 3); P("decoder.conv_out.bias", dd["out_ch"])
        P("quantize.embedding.weight", n_embed, embed_dim)
        P("quant_conv.weight", embed_dim, dd["z_channels"], 1, 1); P("quant_conv.bias", embed_dim)
        P("post_quant_conv.weight", dd["z_channels"], embed_dim, 1, 1); P("post_quant_conv.bias", dd["z_channels"])

    # ------------------------------------------------------------------------------------------
    def load_state_dict(self, state_dict, strict=True, assign=False):
        """Same keys as the reference's MOVQ (autoencoder.py:167-174); training-only `loss.*` entries are dropped."""
        sd = {k: v for k, v in state_dict.items() if not k.startswith("loss.")}
        self._packed = None
        return super().load_state_dict(sd, strict=strict, assign=assign)

    def _apply(self, fn, recurse=True):
        self._packed = None
        return super()._apply(fn, recurse)

    @torch.no_grad()
    def init_synthetic_(self, seed=0):
        dev = self._get("post_quant_conv.weight").device
        g = torch.Generator(device=dev).manual_seed(seed)
        for name, prm in self.named_parameters():
            if name.endswith("bias"):
                prm.normal_(0.0, 0.05, generator=g)
            elif prm.dim() == 1:
                prm.normal_(0.0, 0.1, generator=g).add_(1.0)
            elif name == "quantize.embedding.weight":
                prm.normal_(0.0, 1.0, generator=g)
            else:
                prm.normal_(0.0, prm[0].numel() ** -0.5, generator=g)
        self._packed = None
        return self

    def _get(self, key):
        node = self
        for name in key.split("."):
            node = node._modules[name] if name in node._modules else node._parameters[name]
        return node

    def finalize(self):
        dev = self._get("post_quant_conv.weight").device
        if dev.type != "cuda":
            raise K2Error("MOVQ must live on a CUDA sm_100 device; there is no CPU path")
        f32 = lambda k: self._get(k).detach().to(torch.float32).contiguous()
        pk = {}

        def sn(p):
            c = self._get(p + "norm_layer.weight").shape[0]
            w = torch.cat([f32(p + "conv_y.weight").reshape(c, -1), f32(p + "conv_y.bias")[:, None],
                           f32(p + "conv_b.weight").reshape(c, -1), f32(p + "conv_b.bias")[:, None]], 1).contiguous()
            return dict(g=f32(p + "norm_layer.weight"), b=f32(p + "norm_layer.bias"), w=w)

        def res(p, cin, cout):
            d = dict(n1=sn(p + "norm1."), n2=sn(p + "norm2."), w1=ops.pack_conv_weight(self._get(p + "conv1.weight")),
                     c1=f32(p + "conv1.bias"), c2=f32(p + "conv2.bias"))
            w2 = ops.pack_conv_weight(self._get(p + "conv2.weight"))
            if cin != cout:
                w2 = torch.cat([w2, ops.pack_conv_weight(self._get(p + "nin_shortcut.weight"))], 1).contiguous()
                d["c2"] = d["c2"] + f32(p + "nin_shortcut.bias")
            d["w2"] = w2
            return d

        def att_common(p):
            pc = ops.pack_conv_weight
            return dict(wqk=torch.cat([pc(self._get(p + "q.weight")), pc(self._get(p + "k.weight"))], 0).contiguous(),
                        bqk=torch.cat([f32(p + "q.bias"), f32(p + "k.bias")]).contiguous(),
                        wv=pc(self._get(p + "v.weight")), bv=f32(p + "v.bias"),
                        wp=pc(self._get(p + "proj_out.weight")), bp=f32(p + "proj_out.bias"))

        def att(p):
            d = att_common(p)
            d["n"] = sn(p + "norm.")
            return d

        pk["pq_w"] = f32("post_quant_conv.weight").reshape(self.ddconfig["z_channels"], self.embed_dim).contiguous()
        pk["pq_b"] = f32("post_quant_conv.bias")
        pk["in_w"] = ops.pack_stem_weight(self._get("decoder.conv_in.weight"))
        pk["in_b"] = f32("decoder.conv_in.bias")
        pk["mid1"] = res("decoder.mid.block_1.", self.block_in, self.block_in)
        pk["mida"] = att("decoder.mid.attn_1.")
        pk["mid2"] = res("decoder.mid.block_2.", self.block_in, self.block_in)
        for lv in self.levels:
            p = f"decoder.up.{lv['level']}."
            for bi, (cin, cout) in enumerate(lv["blocks"]):
                pk[p + f"block.{bi}"] = res(p + f"block.{bi}.", cin, cout)
                if lv["attn"]:
                    pk[p + f"attn.{bi}"] = att(p + f"attn.{bi}.")
            if lv["up"]:
                pk[p + "up_w"] = ops.pack_conv_weight(self._get(p + "upsample.conv.weight"))
                pk[p + "up_b"] = f32(p + "upsample.conv.bias")
        # ---- encoder
        def gn(p):
            return dict(g=f32(p + "weight"), b=f32(p + "bias"))

        def eres(p, cin, cout):
            d = dict(n1=gn(p + "norm1."), n2=gn(p + "norm2."), w1=ops.pack_conv_weight(self._get(p + "conv1.weight")),
                     c1=f32(p + "conv1.bias"), c2=f32(p + "conv2.bias"))
            w2 = ops.pack_conv_weight(self._get(p + "conv2.weight"))
            if cin != cout:
                w2 = torch.cat([w2, ops.pack_conv_weight(self._get(p + "nin_shortcut.weight"))], 1).contiguous()
                d["c2"] = d["c2"] + f32(p + "nin_shortcut.bias")
            d["w2"] = w2
            return d

        def eatt(p):
            d = att_common(p)
            d["n"] = gn(p + "norm.")
            return d

        pk["e_in_w"] = ops.pack_stem_weight(self._get("encoder.conv_in.weight"))
        pk["e_in_b"] = f32("encoder.conv_in.bias")
        for lv in self.enc_levels:
            p = f"encoder.down.{lv['level']}."
            for bi, (cin, cout) in enumerate(lv["blocks"]):
                pk[p + f"block.{bi}"] = eres(p + f"block.{bi}.", cin, cout)
                if lv["attn"]:
                    pk[p + f"attn.{bi}"] = eatt(p + f"attn.{bi}.")
            if lv["down"]:
                pk[p + "down_w"] = ops.pack_conv_weight(self._get(p + "downsample.conv.weight"))
                pk[p + "down_b"] = f32(p + "downsample.conv.bias")
        ce = self.enc_levels[-1]["ch"]
        pk["e_mid1"] = eres("encoder.mid.block_1.", ce, ce)
        pk["e_mida"] = eatt("encoder.mid.attn_1.")
        pk["e_mid2"] = eres("encoder.mid.block_2.", ce, ce)
        pk["e_out_n"] = gn("encoder.norm_out.")
        pk["e_out_w"] = ops.pad_rows(ops.pack_conv_weight(self._get("encoder.conv_out.weight")), 16)
        pk["e_out_b"] = f32("encoder.conv_out.bias")
        pk["qc_w"] = f32("quant_conv.weight").reshape(self.embed_dim, -1).contiguous()
        pk["qc_b"] = f32("quant_conv.bias")
        pk["out_n"] = sn("decoder.norm_out.")
        pk["out_w"] = ops.pad_rows(ops.pack_conv_weight(self._get("decoder.conv_out.weight")), 16)
        pk["out_b"] = f32("decoder.conv_out.bias")
        pk["codebook"] = f32("quantize.embedding.weight")
        self._packed = pk
        return self

    # ------------------------------------------------------------------------------------------
    def _sn_act(self, x, zq, n, act):
        """SpatialNorm (decoder: zq given) or plain GroupNorm(32, eps 1e-6) (encoder: zq None), optional swish."""
        st = ops.gn_stats(x, None, groups=32, eps=1e-6)
        if zq is None:
            return ops.gn_apply(x, None, st, n["g"], n["b"], act=act)
        return ops.gn_apply(x, None, st, n["g"], n["b"], act=act, zq=zq, sn_w=n["w"])

    def _res(self, x, zq, d):
        cout = d["c1"].shape[0]
        h = ops.conv_gemm([(self._sn_act(x, zq, d["n1"], 1), 9)], d["w1"], cout, bias=d["c1"])
        h = self._sn_act(h, zq, d["n2"], 1)
        if x.shape[-1] == cout:
            return ops.conv_gemm([(h, 9)], d["w2"], cout, bias=d["c2"], residual=x)
        return ops.conv_gemm([(h, 9), (x, 1)], d["w2"], cout, bias=d["c2"])

    def _attn(self, x, zq, d):
        B, H, W, C = x.shape
        T = H * W
        if T % 64:
            raise K2Error("MoVQ attention needs h*w to be a multiple of 64 (latents are multiples of 8 px)")
        hn = self._sn_act(x, zq, d["n"], 0).view(B, T, C)
        qk = ops.gemm_rows(hn, d["wqk"], 2 * C, bias=d["bqk"])
        o = torch.empty((B, T, C), dtype=torch.float16, device=x.device)
        scores = torch.empty((T, T), dtype=torch.float16, device=x.device)
        vT = torch.empty((C, T), dtype=torch.float16, device=x.device)
        for b in range(B):
            ops.gemm_rows(d["wv"], hn[b], T, out=vT)                      # V^T = W_v hn^T   [C, T]
            ops.gemm_rows(qk[b, :, :C], qk[b, :, C:], T, out=scores)      # q k^T            [T, T]
            ops.softmax_rows(scores, C ** -0.5, out=scores)
            ops.gemm_rows(scores, vT, C, bias=d["bv"], out=o[b])          # P V (+ b_v)      [T, C]
        return ops.gemm_rows(o, d["wp"], C, bias=d["bp"], residual=x.view(B, T, C)).view(B, H, W, C)

    @torch.no_grad()
    def decode(self, quant, out_dtype=None):
        """quant [B, z_channels, h, w] -> image [B, out_ch, H, W] (autoencoder.py:182-185). Output dtype follows
        the input (the reference decodes in fp16 when the pipeline is fp16) unless out_dtype is given."""
        if not quant.is_cuda:
            raise K2Error("k2b200 MOVQ.decode: input must be a CUDA tensor (no CPU fallback)")
        if self._packed is None:
            self.finalize()
        pk = self._packed
        q32 = quant.float().contiguous()
        zq = ops.nchw_to_nhwc_f32(q32)
        z2 = ops.pointwise_nchw_f32(q32, pk["pq_w"], pk["pq_b"])
        h = ops.gemm_rows(ops.stem_im2col(z2), pk["in_w"], self.block_in, bias=pk["in_b"])
        h = self._res(h, zq, pk["mid1"])
        h = self._attn(h, zq, pk["mida"])
        h = self._res(h, zq, pk["mid2"])
        for lv in self.levels:
            p = f"decoder.up.{lv['level']}."
            for bi in range(len(lv["blocks"])):
                h = self._res(h, zq, pk[p + f"block.{bi}"])
                if lv["attn"]:
                    h = self._attn(h, zq, pk[p + f"attn.{bi}"])
            if lv["up"]:
                h = ops.conv_gemm([(ops.upsample2x(h), 9)], pk[p + "up_w"], lv["ch"], bias=pk[p + "up_b"])
        h = self._sn_act(h, zq, pk["out_n"], 1)
        img = ops.conv_gemm([(h, 9)], pk["out_w"], self.ddconfig["out_ch"], bias=pk["out_b"], out_mode=1)
        dt = out_dtype or (quant.dtype if quant.is_floating_point() else torch.float32)
        return img if dt == torch.float32 else img.to(dt)

    @torch.no_grad()
    def encode(self, x):
        """image [B, 3, H, W] in [-1, 1] -> latent fp32 [B, embed_dim, H/8, W/8], no quantisation (autoencoder.py:176-180:
        quant_conv(Encoder(x))).  The stride-2 Downsample conv (pad (0,1,0,1), vqgan_blocks.py:109-126) is evaluated as
        the stride-1 'same' conv on tensor cores followed by taking the odd pixels."""
        if not x.is_cuda:
            raise K2Error("k2b200 MOVQ.encode: input must be a CUDA tensor (no CPU fallback)")
        if self._packed is None:
            self.finalize()
        pk = self._packed
        h = ops.gemm_rows(ops.stem_im2col(x.float().contiguous()), pk["e_in_w"], self.ddconfig["ch"], bias=pk["e_in_b"])
        for lv in self.enc_levels:
            p = f"encoder.down.{lv['level']}."
            for bi in range(len(lv["blocks"])):
                h = self._res(h, None, pk[p + f"block.{bi}"])
                if lv["attn"]:
                    h = self._attn(h, None, pk[p + f"attn.{bi}"])
            if lv["down"]:
                full = ops.conv_gemm([(h, 9)], pk[p + "down_w"], lv["ch"], bias=pk[p + "down_b"])
                h = ops.subsample2(full, 1, 1)
        h = self._res(h, None, pk["e_mid1"])
        h = self._attn(h, None, pk["e_mida"])
        h = self._res(h, None, pk["e_mid2"])
        h = self._sn_act(h, None, pk["e_out_n"], 1)
        zc_out = pk["e_out_b"].shape[0]
        z = ops.conv_gemm([(h, 9)], pk["e_out_w"], zc_out, bias=pk["e_out_b"], out_mode=1)
        return ops.pointwise_nchw_f32(z, pk["qc_w"], pk["qc_b"])

    @torch.no_grad()
    def decode_to_uint8(self, quant, crop_h=None, crop_w=None):
        """decode + process_images (kandinsky2/utils.py:57-70) fused on the device -> uint8 NHWC (cropped)."""
        img = self.decode(quant, out_dtype=torch.float32)
        return ops.images_to_u8(img, crop_h or img.shape[2], crop_w or img.shape[3])

    @torch.no_grad()
    def quantize_indices(self, z):
        """Nearest-codebook indices of z [B, e_dim, h, w] (quntize.py:80-99): int64 [B*h*w], ties -> lowest index."""
        if self._packed is None:
            self.finalize()
        zf = ops.nchw_to_nhwc_f32(z.float().contiguous()).reshape(-1, self.embed_dim)
        return ops.vq_argmin(zf, self._packed["codebook"])
