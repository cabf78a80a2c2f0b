"""Thin torch-tensor wrappers over the C ABI (include/k2b200.h).

torch is used for device memory, streams and one-off weight re-layout only; every arithmetic op on the
hot path is a kernel of libk2b200.so.  Activations are NHWC fp16 tensors of shape [NB, H, W, C] (a view
may be a channel slice of a wider buffer: stride(-2) is the row stride).
"""
import ctypes

import torch

from . import _native as nat
from ._native import K2ConvSrc, check, ptr, stream_ptr


def _row_stride(t):
    assert t.stride(-1) == 1, "channel dim must be contiguous"
    ld = t.stride(-2)
    # all leading dims must be row-contiguous w.r.t. ld
    n = t.shape[-2]
    for d in range(t.dim() - 3, -1, -1):
        assert t.shape[d] == 1 or t.stride(d) == ld * n, f"not a row-strided NHWC view: {t.shape} {t.stride()}"
        n *= t.shape[d]
    return ld


# ------------------------------------------------------------------------------------------------
# weight packing (host side, once per checkpoint load)
# ------------------------------------------------------------------------------------------------
def _pad64(c):
    return (c + 63) // 64 * 64


def pack_conv_weight(w, split=None):
    """[Cout, Cin, kh, kw] (kh=kw in {1,3}) or [Cout, Cin] / [Cout, Cin, 1] -> fp16 [Cout, taps*pad64(Cin)].

    k = tap * pad64(Cin) + c with tap = ky*3+kx (k2b200.h: k2_conv_gemm).  `split=(C0, C1)` packs a 1x1
    weight whose input is the channel concat of two sources as two independently padded segments.
    """
    w = w.detach()
    if w.dim() == 2:
        w = w[:, :, None, None]
    elif w.dim() == 3:
        w = w[:, :, :, None]
    cout, cin, kh, kw = w.shape
    taps = kh * kw
    assert taps in (1, 9)
    parts = [(0, cin)] if split is None else [(0, split[0]), (split[0], split[0] + split[1])]
    segs = []
    for lo, hi in parts:
        c = hi - lo
        ws = w[:, lo:hi].permute(0, 2, 3, 1).reshape(cout, taps, c)
        if _pad64(c) != c:
            ws = torch.nn.functional.pad(ws, (0, _pad64(c) - c))
        segs.append(ws.reshape(cout, taps * _pad64(c)))
    return torch.cat(segs, dim=1).to(torch.float16).contiguous()


_UP2_TAPS = {0: ((0,), (1, 2)), 1: ((0, 1), (2,))}  # output parity a -> 3x3 kernel rows feeding source offsets (a - 1, a)


def pack_conv_weight_up2(w):
    """[Cout, Cin, 3, 3] -> fp16 [Cout, 16 * pad64(Cin)] for a `taps = 4` source of k2_conv_gemm: the 3x3 convolution over the
    nearest-2x upsampled input, as four 2x2 convolutions over the input itself, one per output parity (a, b).  Output pixel
    (2y + a, 2x + b) sees upsampled rows 2y + a + ky - 1, i.e. source rows y + floor((a + ky - 1) / 2): the kernel rows that
    land on the same source row are summed (fp32) before the fp16 rounding.  k = ((a * 2 + b) * 4 + ty * 2 + tx) * pad64(Cin) + c,
    source offset (dy, dx) = (ty + a - 1, tx + b - 1).  2.25x fewer MACs than convolving the upsampled tensor, which is never
    materialised (unet.py:67-77 Upsample + :199-203 h_upd; movq_modules.py:93-97)."""
    w = w.detach().float()
    cout, cin = w.shape[:2]
    blocks = []
    for a in (0, 1):
        for b in (0, 1):
            for ty in (0, 1):
                for tx in (0, 1):
                    acc = torch.zeros(cout, cin, dtype=torch.float32, device=w.device)
                    for ky in _UP2_TAPS[a][ty]:
                        for kx in _UP2_TAPS[b][tx]:
                            acc += w[:, :, ky, kx]
                    if _pad64(cin) != cin:
                        acc = torch.nn.functional.pad(acc, (0, _pad64(cin) - cin))
                    blocks.append(acc)
    return torch.cat(blocks, dim=1).to(torch.float16).contiguous()


def pad_rows(wp, rows):
    if wp.shape[0] >= rows:
        return wp
    return torch.cat([wp, wp.new_zeros(rows - wp.shape[0], wp.shape[1])], 0).contiguous()


# ------------------------------------------------------------------------------------------------
# conv / GEMM
# ------------------------------------------------------------------------------------------------
_CONV_WS_BYTES = 256 << 20  # split-K partial sums (lower half), tail-split hand-over tiles + flags (upper half); fixed size and
# address (baked into captured CUDA graphs); conv launches that share it must be stream-ordered
_conv_ws = {}


def _workspace(dev):
    buf = _conv_ws.get(dev.index)
    if buf is None:
        buf = torch.empty(_CONV_WS_BYTES, dtype=torch.uint8, device=dev)
        buf[-(64 << 10):].zero_()  # tail-split hand-over flags (k2b200.h: zero before the first launch, kept zero by the library)
        _conv_ws[dev.index] = buf
    return buf


def conv_gemm(srcs, w_packed, cout, bias=None, residual=None, out=None, out_mode=0, geom=None, gn_part=None, info=None,
              cfg=None, w_batch_stride=0):
    """srcs: list of (tensor NHWC fp16 [NB,H,W,C], taps); taps = 9 (3x3), 1 (1x1) or 4 (single source: 3x3 over its nearest-2x
    upsampling, weights from pack_conv_weight_up2, output [NB,2H,2W,cout]).  Returns fp16 [NB,H,W,cout] (out_mode 0) or
    fp32 NCHW [NB,cout,H,W] (out_mode 1).  geom=(NB,H,W) overrides the geometry (GEMM on flat rows).
    cfg = (N tile, pair mode, splits, epilogue sets) overrides the library's choice (k2_conv_gemm_cfg; 0 = auto).
    w_batch_stride > 0: batched GEMM, image n uses the weight matrix at w_packed + n * w_batch_stride elements (w_packed is then
    any fp16 tensor whose data_ptr() is matrix 0, shape[0] / shape[1] / stride(0) = rows / K / row stride of ONE matrix)."""
    lib = nat.load()
    t0 = srcs[0][0]
    NB, H, W = geom if geom is not None else t0.shape[:3]
    if srcs[0][1] == 4:  # 3x3 conv over the nearest-2x upsampling of the source: geometry = the OUTPUT's
        H, W = 2 * H, 2 * W
    arr = (K2ConvSrc * len(srcs))()
    for i, (t, taps) in enumerate(srcs):
        assert t.dtype == torch.float16 and t.is_cuda
        arr[i].ptr = t.data_ptr()
        arr[i].C = t.shape[-1]
        arr[i].ld = _row_stride(t)
        arr[i].taps = taps
    if out is None:
        if out_mode == 0:
            out = torch.empty((NB, H, W, cout), dtype=torch.float16, device=t0.device)
        else:
            out = torch.empty((NB, cout, H, W), dtype=torch.float32, device=t0.device)
    ldo = _row_stride(out) if out_mode == 0 else 0
    ldr = _row_stride(residual) if residual is not None else 0
    assert w_packed.dtype == torch.float16 and w_packed.stride(1) == 1
    ws = _workspace(t0.device)
    _info = (ctypes.c_int * 7)()
    _cfg = (ctypes.c_int * 4)(*cfg) if cfg is not None else None
    check(lib.k2_conv_gemm_cfg(arr, len(srcs), NB, H, W, ptr(w_packed), w_packed.shape[0], w_packed.shape[1],
                               w_packed.stride(0), cout,
                               ptr(bias), ptr(residual), ldr, ptr(out), ldo, out_mode, ptr(ws), ws.numel(), ptr(gn_part),
                               _info, _cfg, int(w_batch_stride), stream_ptr()))
    if info is not None:
        info[:] = list(_info)
    return out


def conv_last_tail_split():
    """K parts of the tail split used by the most recent conv_gemm / conv_plan call of this thread (1 = none)."""
    return nat.load().k2_conv_last_tail_split()


def conv_plan(NB, H, W, taps, ktot, cout, out_mode=0, workspace_bytes=1 << 28, want_gn_partial=True):
    """What k2_conv_gemm would decide for this geometry (host arithmetic only, no GPU): dict of the info[7] fields."""
    lib = nat.load()
    info = (ctypes.c_int * 7)()
    check(lib.k2_conv_plan(NB, H, W, taps, ktot, cout, out_mode, workspace_bytes, int(want_gn_partial), info))
    keys = ("n_tile", "cta_pair", "splits", "m_tiles", "images_per_tile", "gn_partial_mode", "row_groups")
    return dict(zip(keys, list(info)))


def gemm_rows(x, w_packed, cout, bias=None, residual=None, out=None, cfg=None, info=None):
    """x: fp16 [..., K] rows -> fp16 [..., cout]; one 1x1 'conv' over M = prod(leading dims) rows."""
    lead = x.shape[:-1]
    M = 1
    for d in lead:
        M *= d
    x3 = x.reshape(1, 1, M, x.shape[-1]) if x.is_contiguous() else None
    if x3 is None:
        # row-strided view (channel slice): keep stride
        x3 = x.as_strided((1, 1, M, x.shape[-1]), (0, 0, x.stride(-2), 1))
    if out is None:
        out = torch.empty(tuple(lead) + (cout,), dtype=torch.float16, device=x.device)
    o3 = out.as_strided((1, 1, M, cout), (0, 0, out.stride(-2), 1))
    r3 = None
    if residual is not None:
        r3 = residual.as_strided((1, 1, M, cout), (0, 0, residual.stride(-2), 1))
    conv_gemm([(x3, 1)], w_packed, cout, bias=bias, residual=r3, out=o3, geom=(1, 1, M), cfg=cfg, info=info)
    return out


# ------------------------------------------------------------------------------------------------
# GroupNorm
# ------------------------------------------------------------------------------------------------
_gn_scratch = {}


_GN_SCRATCH_FLOATS = 1 << 23  # 32 MB, fixed: its address is baked into captured CUDA graphs


def _scratch(dev, nfloats):
    key = (dev.index, )
    buf = _gn_scratch.get(key)
    if buf is None:
        buf = torch.zeros(_GN_SCRATCH_FLOATS, dtype=torch.float32, device=dev)
        _gn_scratch[key] = buf
    if nfloats > buf.numel():
        raise nat.K2Error(f"gn_stats scratch of {buf.numel()} floats is too small for this geometry ({nfloats})")
    return buf


def gn_stats(x0, x1=None, groups=32, eps=1e-5, stats=None):
    """Per (image, group) [mean, rstd] of the channel concat [x0 | x1]; x*: fp16 NHWC."""
    lib = nat.load()
    NB, H, W, C0 = x0.shape
    C1 = x1.shape[-1] if x1 is not None else 0
    if stats is None:
        stats = torch.empty((NB, groups, 2), dtype=torch.float32, device=x0.device)
    need = lib.k2_gn_scratch_floats(NB, H * W, C0 + C1)
    scratch = _scratch(x0.device, need)
    check(lib.k2_gn_stats(ptr(x0), C0, _row_stride(x0), ptr(x1), C1, _row_stride(x1) if x1 is not None else 0,
                          NB, H * W, groups, eps, ptr(stats), ptr(scratch), stream_ptr()))
    return stats


def gn_part_floats(NB, H, W, cout):
    """Upper bound of the fp32 count of a conv's fused-statistics partial buffer ([row groups][cout][2], k2b200.h)."""
    tiles = NB * ((H * W + 63) // 64 + 2 * H)  # generous: boxes are >= 64 pixels except at ragged edges
    return max(tiles * 4, (NB * H * W + 15) // 16) * cout * 2


def gn_finalize(part0, C0, part1, C1, NB, rg_per_image, HW, stats, groups=32, eps=1e-5, rg1=None):
    """rg_per_image: row groups per image of source 0 (and of source 1 unless rg1 is given)."""
    lib = nat.load()
    check(lib.k2_gn_finalize(ptr(part0), C0, rg_per_image, ptr(part1), C1, rg1 if rg1 is not None else rg_per_image, NB, HW,
                             groups, eps, ptr(stats), stream_ptr()))
    return stats


def gn_apply(x0, x1, stats, gamma, beta, film=None, act=1, resample=0, groups=32, y=None, want_xres=False,
             zq=None, sn_w=None, xres=None):
    """Fused normalise (+FiLM) (+SiLU) (+2x up / 2x2 avg-pool) (+concat) -> fp16 NHWC. See k2b200.h."""
    lib = nat.load()
    NB, H, W, C0 = x0.shape
    C1 = x1.shape[-1] if x1 is not None else 0
    C = C0 + C1
    Ho, Wo = (H, W) if resample == 0 else ((H // 2, W // 2) if resample == 1 else (H * 2, W * 2))
    if y is None:
        y = torch.empty((NB, Ho, Wo, C), dtype=torch.float16, device=x0.device)
    if xres is not None:
        want_xres = True
    elif want_xres:
        xres = torch.empty((NB, Ho, Wo, C), dtype=torch.float16, device=x0.device)
    zh = zw = 0
    if zq is not None:
        zh, zw = zq.shape[1], zq.shape[2]
    check(lib.k2_gn_apply(ptr(x0), C0, _row_stride(x0), ptr(x1), C1, _row_stride(x1) if x1 is not None else 0,
                          NB, H, W, groups, ptr(stats), ptr(gamma), ptr(beta), ptr(film),
                          film.stride(0) if film is not None else 0, act, resample,
                          ptr(y), _row_stride(y), ptr(xres), _row_stride(xres) if xres is not None else 0,
                          ptr(zq), zh, zw, ptr(sn_w), stream_ptr()))
    return (y, xres) if want_xres else y


def gn_apply_fold(x0, x1, part0, rg0, part1, rg1, gamma, beta, film=None, act=1, resample=0, groups=32, eps=1e-5, y=None,
                  xres=None):
    """gn_apply that folds the producers' partial sums itself (no gn_finalize launch); round-2 candidate, see k2b200.h."""
    lib = nat.load()
    NB, H, W, C0 = x0.shape
    C1 = x1.shape[-1] if x1 is not None else 0
    Ho, Wo = (H, W) if resample == 0 else ((H // 2, W // 2) if resample == 1 else (H * 2, W * 2))
    if y is None:
        y = torch.empty((NB, Ho, Wo, C0 + C1), dtype=torch.float16, device=x0.device)
    check(lib.k2_gn_apply_fold(ptr(x0), C0, _row_stride(x0), ptr(x1), C1, _row_stride(x1) if x1 is not None else 0,
                               NB, H, W, groups, ptr(part0), rg0, ptr(part1), rg1 if part1 is not None else 0, eps,
                               ptr(gamma), ptr(beta), ptr(film), film.stride(0) if film is not None else 0, act, resample,
                               ptr(y), _row_stride(y), ptr(xres), _row_stride(xres) if xres is not None else 0,
                               stream_ptr()))
    return y


# ------------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------------
def attention_d64(qkv, heads, enc=None, scale=0.125, out=None, hs=192, q_off=0, k_off=64, v_off=128, ehs=128,
                  ek_off=0, ev_off=64):
    """qkv fp16 [B, T, >=heads*hs]; enc fp16 [B, Tc, heads*ehs] or None -> fp16 [B, T, heads*64]."""
    lib = nat.load()
    B, T = qkv.shape[:2]
    Tc = enc.shape[1] if enc is not None else 0
    if out is None:
        out = torch.empty((B, T, heads * 64), dtype=torch.float16, device=qkv.device)
    check(lib.k2_attention_d64(ptr(qkv), qkv.stride(1), hs, q_off, k_off, v_off, ptr(enc),
                               enc.stride(1) if enc is not None else 0, ehs, ek_off, ev_off, B, heads, T, Tc,
                               scale, ptr(out), out.stride(1), stream_ptr()))
    return out


def attention_d512(qkv, scale, out=None, q_off=0, k_off=512, v_off=1024):
    """qkv fp16 [B, T, >= 1536] (q | k | v of ONE head of width 512) -> fp16 [B, T, 512]; the MoVQ AttnBlock fused."""
    lib = nat.load()
    B, T = qkv.shape[:2]
    if out is None:
        out = torch.empty((B, T, 512), dtype=torch.float16, device=qkv.device)
    check(lib.k2_attention_d512(ptr(qkv), qkv.stride(1), q_off, k_off, v_off, B, T, float(scale), ptr(out), out.stride(1),
                                stream_ptr()))
    return out


# ------------------------------------------------------------------------------------------------
# small dense layers
# ------------------------------------------------------------------------------------------------
def linear(x, W, b=None, add=None, silu_in=False, silu_out=False, out=None):
    """fp32 x [M,K] (row-strided ok) @ W[N,K]^T (+b) (+add) -> fp32 [M,N]."""
    lib = nat.load()
    M, K = x.shape
    N = W.shape[0]
    assert W.shape[1] == K and W.is_contiguous() and x.dtype == torch.float32
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=x.device)
    check(lib.k2_linear(ptr(x), x.stride(0), ptr(W), 1 if W.dtype == torch.float16 else 0, ptr(b), ptr(add),
                        add.stride(0) if add is not None else 0, ptr(out), out.stride(0), M, N, K,
                        int(silu_in), int(silu_out), stream_ptr()))
    return out


def layernorm(x, gamma, beta, eps=1e-5):
    lib = nat.load()
    M, N = x.shape
    y = torch.empty_like(x)
    check(lib.k2_layernorm(ptr(x), ptr(gamma), ptr(beta), ptr(y), M, N, eps, stream_ptr()))
    return y


def timestep_embedding(t, dim, max_period=10000.0, out=None):
    lib = nat.load()
    B = t.shape[0]
    if out is None:
        out = torch.empty((B, dim), dtype=torch.float32, device=t.device)
    check(lib.k2_timestep_embedding(ptr(t), ptr(out), B, dim, max_period, stream_ptr()))
    return out


def f32_to_f16(x, out=None):
    lib = nat.load()
    assert x.is_contiguous() and x.dtype == torch.float32
    if out is None:
        out = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    check(lib.k2_f32_to_f16(ptr(x), ptr(out), x.numel(), stream_ptr()))
    return out


def silu_f16_(x):
    """SiLU in place on a contiguous fp16 tensor."""
    lib = nat.load()
    assert x.dtype == torch.float16 and x.is_contiguous()
    check(lib.k2_silu_f16(ptr(x), ptr(x), x.numel(), stream_ptr()))
    return x


def stem_im2col(x, x2=None, x3=None, mul23=False, kpad=None, out=None):
    """fp32 NCHW inputs -> fp16 [NB, H, W, kpad] 3x3 patches of cat([x, x2*(x3 if mul23), x3], 1)."""
    lib = nat.load()
    NB, Cx, H, W = x.shape
    C2 = x2.shape[1] if x2 is not None else 0
    C3 = x3.shape[1] if x3 is not None else 0
    cin = Cx + C2 + C3
    if kpad is None:
        kpad = _pad64(9 * cin)
    if out is None:
        out = torch.empty((NB, H, W, kpad), dtype=torch.float16, device=x.device)
    check(lib.k2_stem_im2col(ptr(x), Cx, ptr(x2), C2, ptr(x3), C3, int(mul23), NB, H, W, ptr(out), kpad,
                             stream_ptr()))
    return out


def pack_stem_weight(w):
    """[Cout, Cin, 3, 3] -> fp16 [Cout, pad64(9*Cin)] with k = tap*Cin + c (matches stem_im2col)."""
    cout, cin = w.shape[:2]
    wp = w.detach().permute(0, 2, 3, 1).reshape(cout, 9 * cin)
    k = _pad64(9 * cin)
    if k != 9 * cin:
        wp = torch.nn.functional.pad(wp, (0, k - 9 * cin))
    return wp.to(torch.float16).contiguous()


# ------------------------------------------------------------------------------------------------
# sampler / MoVQ helpers
# ------------------------------------------------------------------------------------------------
def sampler_step(model_out, x, noise, coef, guidance, cond_first, clip=2.0, threshold_mode=0, inpaint_init=None,
                 inpaint_mask=None, work=None, inpaint_noise=None):
    lib = nat.load()
    B, _, H, W = x.shape
    if work is None:
        work = torch.empty(B * 4 * H * W + 4096, dtype=torch.float32, device=x.device)
    check(lib.k2_sampler_step(ptr(model_out), ptr(x), ptr(noise), ptr(coef), B, H, W, float(guidance),
                              int(cond_first), float(clip), int(threshold_mode), ptr(inpaint_init),
                              ptr(inpaint_mask), ptr(inpaint_noise), ptr(work), stream_ptr()))
    return x


def step_begin(x, x_in, t_in, coef_out, ts_seq, coef_seq, noise_seq, noise, counter):
    """k2_step_begin: counter = device int32 [2] = (k, nsteps); see k2b200.h."""
    lib = nat.load()
    n = x.numel()
    assert x_in.numel() == 2 * n and x.is_contiguous() and x_in.is_contiguous()
    check(lib.k2_step_begin(ptr(x), ptr(x_in), n, ptr(t_in), t_in.numel(), ptr(coef_out), ptr(ts_seq), ptr(coef_seq),
                            ptr(noise_seq), ptr(noise), ptr(counter), stream_ptr()))


def step_end(counter):
    check(nat.load().k2_step_end(ptr(counter), stream_ptr()))


def plms_step(model_out, x, out, hist, store, coef, guidance, cond_first):
    """hist: list of up to 3 fp32 [B,4,H,W] tensors, newest first (None entries allowed)."""
    lib = nat.load()
    B, _, H, W = x.shape
    hh = list(hist) + [None] * (3 - len(hist))
    check(lib.k2_plms_step(ptr(model_out), model_out.shape[1], ptr(x), ptr(out), ptr(hh[0]), ptr(hh[1]), ptr(hh[2]), ptr(store),
                           ptr(coef), B, H, W, float(guidance), int(cond_first), stream_ptr()))
    return out


def vq_argmin(z, codebook):
    """z fp32 [n, dim], codebook fp32 [n_embed, dim] -> int64 [n] (ties -> lowest index)."""
    lib = nat.load()
    n, dim = z.shape
    idx = torch.empty((n,), dtype=torch.int64, device=z.device)
    check(lib.k2_vq_argmin(ptr(z), ptr(codebook), ptr(idx), n, codebook.shape[0], dim, stream_ptr()))
    return idx


def pointwise_nchw_f32(x, w, b, out=None):
    lib = nat.load()
    NB, Ci, H, W = x.shape
    Co = w.shape[0]
    y = out if out is not None else torch.empty((NB, Co, H, W), dtype=torch.float32, device=x.device)
    check(lib.k2_pointwise_nchw_f32(ptr(x), ptr(w), ptr(b), ptr(y), NB, Ci, Co, H * W, stream_ptr()))
    return y


def upsample2x(x, out=None):
    """fp16 NHWC [NB,H,W,C] -> nearest 2x [NB,2H,2W,C]."""
    lib = nat.load()
    NB, H, W, C = x.shape
    if out is None:
        out = torch.empty((NB, 2 * H, 2 * W, C), dtype=torch.float16, device=x.device)
    check(lib.k2_upsample2x_nhwc(ptr(x), _row_stride(x), ptr(out), _row_stride(out), NB, H, W, C, stream_ptr()))
    return out


def subsample2(x, oy=1, ox=1, out=None):
    """fp16 NHWC [NB,H,W,C] -> [NB,H/2,W/2,C] taking pixels (2y+oy, 2x+ox)."""
    lib = nat.load()
    NB, H, W, C = x.shape
    if out is None:
        out = torch.empty((NB, H // 2, W // 2, C), dtype=torch.float16, device=x.device)
    check(lib.k2_subsample2_nhwc(ptr(x), _row_stride(x), ptr(out), _row_stride(out), NB, H, W, C, oy, ox, stream_ptr()))
    return out


def softmax_rows(x, scale, out=None):
    """fp16 [rows, n] (row-strided) -> softmax(scale * x) fp16."""
    lib = nat.load()
    rows, n = x.shape
    if out is None:
        out = torch.empty((rows, n), dtype=torch.float16, device=x.device)
    check(lib.k2_softmax_rows(ptr(x), x.stride(0), ptr(out), out.stride(0), rows, n, float(scale), stream_ptr()))
    return out


def sn_apply(x, stats, gamma, beta, zq, sn_w, act=1, groups=32, y=None):
    """MoVQ SpatialNorm (+ swish): x fp16 NHWC [NB,H,W,C], zq fp32 [NB,zh,zw,4], sn_w fp32 [C,10] -> fp16 NHWC (k2_sn_apply)."""
    lib = nat.load()
    NB, H, W, C = x.shape
    if y is None:
        y = torch.empty((NB, H, W, C), dtype=torch.float16, device=x.device)
    check(lib.k2_sn_apply(ptr(x), C, _row_stride(x), NB, H, W, groups, ptr(stats), ptr(gamma), ptr(beta), ptr(zq),
                          zq.shape[1], zq.shape[2], ptr(sn_w), act, ptr(y), _row_stride(y), stream_ptr()))
    return y


def transpose_f16(x, out=None):
    """fp16 [B, T, C] (row-strided) -> contiguous [B, C, T]."""
    lib = nat.load()
    B, T, C = x.shape
    assert x.stride(-1) == 1 and x.stride(0) == T * x.stride(1)
    if out is None:
        out = torch.empty((B, C, T), dtype=torch.float16, device=x.device)
    check(lib.k2_transpose_f16(ptr(x), x.stride(1), ptr(out), B, T, C, stream_ptr()))
    return out


def nchw_to_nhwc_f32(x, out=None):
    lib = nat.load()
    NB, C, H, W = x.shape
    y = out if out is not None else torch.empty((NB, H, W, C), dtype=torch.float32, device=x.device)
    check(lib.k2_nchw_to_nhwc_f32(ptr(x), ptr(y), NB, C, H, W, stream_ptr()))
    return y


def images_to_u8(x, crop_h, crop_w, out=None):
    lib = nat.load()
    NB, C, H, W = x.shape
    if out is None:
        out = torch.empty((NB, crop_h, crop_w, C), dtype=torch.uint8, device=x.device)
    check(lib.k2_images_to_u8(ptr(x), ptr(out), NB, C, H, W, crop_h, crop_w, stream_ptr()))
    return out


def set_tuning(key, value):
    """k2_set_tuning: key 0 = force conv N tile, key 1 = split-K (0 auto, 1 off, n forced)."""
    check(nat.load().k2_set_tuning(int(key), int(value)))


def launch_count():
    return nat.load().k2_launch_count()


def reset_launch_count():
    nat.load().k2_reset_launch_count()


# ------------------------------------------------------------------------------------------------
# diffusion prior helpers (groundwork, see k2b200.h)
# ------------------------------------------------------------------------------------------------
def layernorm_f16(x, gamma, beta, eps=1e-5, out=None):
    """LayerNorm over the last dim of fp16 rows [..., N] (fp32 gain / bias) -> fp16."""
    lib = nat.load()
    N = x.shape[-1]
    M = x.numel() // N
    assert x.dtype == torch.float16 and x.stride(-1) == 1
    if out is None:
        out = torch.empty_like(x)
    check(lib.k2_layernorm_f16(ptr(x), _row_stride(x), ptr(gamma), ptr(beta), ptr(out), _row_stride(out), M, N, eps,
                               stream_ptr()))
    return out


def gelu_f16_(x):
    """Exact GELU in place on a contiguous fp16 tensor."""
    lib = nat.load()
    assert x.dtype == torch.float16 and x.is_contiguous()
    check(lib.k2_gelu_f16(ptr(x), ptr(x), x.numel(), stream_ptr()))
    return x


def attention_small(qkv, heads, keep_mask=None, causal=True, scale=0.125, out=None):
    """qkv fp16 [B, T, heads*192] (per head [q|k|v]), keep_mask uint8 [B, T] or None -> fp16 [B, T, heads*64]; T <= 128."""
    lib = nat.load()
    B, T = qkv.shape[:2]
    if out is None:
        out = torch.empty((B, T, heads * 64), dtype=torch.float16, device=qkv.device)
    check(lib.k2_attention_small(ptr(qkv), _row_stride(qkv), ptr(keep_mask), int(causal), ptr(out), _row_stride(out), B, T,
                                 heads, scale, stream_ptr()))
    return out
