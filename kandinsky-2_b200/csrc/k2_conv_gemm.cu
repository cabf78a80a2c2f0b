// k2_conv_gemm.cu -- im2col-free 3x3 / 1x1 convolution and plain GEMM on tcgen05 tensor cores.
//
// Replaces the cuDNN / cuBLAS call sites of the reference hot path:
//   nn.Conv2d 3x3   kandinsky2/model/unet.py:152,180,426,562 ; vqgan/movq_modules.py:139-148,268-270
//   nn.Conv2d 1x1   kandinsky2/model/unet.py:191 (skip_connection) ; movq_modules.py:150-157,188-199
//   nn.Conv1d k=1   kandinsky2/model/unet.py:251,257,258 (qkv / encoder_kv / proj_out)
//
// Formulation: D[M = pixels, N = Cout] = sum over (segment, tap, 64-channel chunk) A_tap[M, 64] * W[N, 64]^T.
//   * Activations are NHWC fp16. An M tile is a (TN x TH x TW) box of output pixels (<= 128 rows).
//     For tap (dy, dx) the A operand is the SAME box shifted by (dy, dx), fetched with ONE 4-D TMA
//     whose out-of-bounds elements are zero-filled by the hardware: conv padding costs nothing and no
//     im2col buffer exists in HBM or smem.
//   * Up to three A "segments" accumulate into the same TMEM tile: the 3x3 conv input plus 1x1 skip
//     inputs (raw x, optionally split in two for the un-materialised torch.cat of the up path). This is
//     how ResBlock's  skip_connection(x) + conv(h)  (unet.py:220) becomes a single kernel.
//   * Weights are pre-packed [Cout][K] fp16, K = concat over segments/taps/channels, loaded by 2-D TMA.
//   * Warp roles (256 threads): warp0 = TMA producer, warp1 = tcgen05.mma issuer, warp2 = TMEM alloc,
//     warps4-7 = epilogue (tcgen05.ld -> +bias (+residual) -> fp16 -> global). Persistent over tiles,
//     smem ring of STAGES (A 16 KB + B BN*128 B), TMEM accumulator double-buffered (2 x BN columns) so
//     the epilogue of tile i overlaps the mainloop of tile i+1.
#include <stdio.h>

#include "k2_common.cuh"
#include "k2_internal.h"

namespace k2 {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB

constexpr int EPI_STAGE_FLOATS = 32 * 33;              // per-warp staging: 4224 B (4096 used)
constexpr int EPI_BYTES = 4 * EPI_STAGE_FLOATS * 4 + 4 * 256 * 4;  // + per-warp bias copy (<= 256 columns)

template <int BN>
struct Cfg {
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 192 ? 5 : (BN >= 128 ? 6 : 8));
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  static constexpr int BAR_BYTES = 256;
  static constexpr int STAT_BYTES = EPI_BYTES;  // epilogue staging (output transpose, residual, statistics) + bias copies
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + STAT_BYTES + 1024;  // +1024 alignment slack
};

// up2 (3x3 conv over the nearest-2x upsampled source as four 2x2 phase convolutions): m_idx = phase * m_tiles_phase + box
// index; the box lives in SOURCE coordinates, `phase` = (a, b) = parity of the output pixel (2y + a, 2x + b).
__device__ __forceinline__ void decode_m_tile(const ConvGemmParams& p, int m_idx, int& n0, int& y0,
                                              int& x0, int& phase) {
  phase = 0;
  if (p.up2) {
    phase = m_idx / p.m_tiles_phase;
    m_idx -= phase * p.m_tiles_phase;
  }
  int tw_i = m_idx % p.tiles_w;
  int t = m_idx / p.tiles_w;
  int th_i = t % p.tiles_h;
  int tn_i = t / p.tiles_h;
  x0 = tw_i * p.TW;
  y0 = th_i * p.TH;
  n0 = tn_i * p.TN;
}

// Epilogue of one (128-row x BN-column) accumulator tile held in this CTA's TMEM at column `acc_col`:
// tcgen05.ld -> + bias (+ residual) -> fp16 rows (out_mode 0), fp32 split-K partials (out_mode 2) or fp32 NCHW
// (out_mode 1).
//
// A thread owns one accumulator ROW (TMEM lane), so storing straight from registers makes every warp store touch 32
// different cache lines with 16 bytes each (and every residual load likewise): with short K loops (the attention
// qkv / proj GEMMs, 12..24 chunks) that epilogue, not the tensor core, set the pace (profiles/conv_small_k.py).  The
// fp16 / split-K paths therefore go through a per-warp shared-memory transpose (32 rows x 128 B, 16-byte pieces XOR-
// swizzled by the row): registers -> smem by row, smem -> global with 8 lanes per row, i.e. 4 complete 128-byte lines
// per store instruction; the residual comes in the same way (coalesced load -> smem -> own row), the bias is read as
// broadcast LDS.128 from a per-warp copy, and the fused GroupNorm statistics are column sums over the staged fp16 tile.

// ES = number of epilogue warp SETS (each set = 4 warps covering the 4 TMEM lane quarters).  ES = 1 is the validated
// configuration; with ES = 2 (384-thread variant of the CTA-pair kernel, tuning key 10, round-2 candidate) set `es` handles
// the 64-column pairs jp with jp % 2 == es, so two warps per scheduler drain the accumulator.
template <int BN, int ES = 1, bool TAIL = false>
__device__ __forceinline__ void epilogue_tile(const ConvGemmParams& p, uint32_t tmem_base, int acc_col, int ew, int lane,
                                              int n0, int y0, int x0, int n_idx, int split, int m_idx, float* stat_smem,
                                              int es_arg = 0, int phase = 0, int tail_role = 0, int tail_slot = 0, int tail_np = 0) {
  // tail_role (CTA-pair kernel only): 0 = ordinary tile; 1 = a K part >= 1 of a tail-split tile: the raw fp32 accumulator goes
  // to p.tail_buf slot tail_slot, then the warp raises its flag; 2 = part 0 (the owner): waits for the flags of slots
  // tail_slot .. tail_slot + tail_np - 1, adds those partial tiles to its accumulator in a fixed order and carries on with the
  // normal epilogue (bias, residual, fp16 store, GroupNorm partials): the consumers cannot tell a tail-split tile from another.
  const int es = (ES == 1) ? 0 : es_arg;
  const int row = ew * 32 + lane;
  const int thw = p.TH * p.TW;
  constexpr int CH = (BN >= 32) ? 32 : 16;  // columns per tcgen05.ld
      const int tn = row / thw;
      const int rem = row - tn * thw;
      const int th = rem / p.TW;
      const int tw = rem - th * p.TW;
      const int n = n0 + tn, y = y0 + th, x = x0 + tw;
      const bool valid = (tn < p.TN) && (n < p.NB) && (y < p.H) && (x < p.W);
      const long long out_row = p.up2 ? (static_cast<long long>(n) * (2 * p.H) + (2 * y + (phase >> 1))) * (2 * p.W) + (2 * x + (phase & 1))
                                      : (static_cast<long long>(n) * p.H + y) * p.W + x;
      // GroupNorm partial row groups stay image-major under up2: (box index) * 4 + phase
      const int m_in = p.up2 ? m_idx - phase * p.m_tiles_phase : m_idx;
      auto part_index = [&](long long base) { return p.up2 ? base * 4 + phase : base; };
      if constexpr (BN % 64 == 0) {
        if ((p.out_mode == 0 && p.Cout % 64 == 0) || (p.out_mode == 2 && p.Cout % 32 == 0)) {
          const uint32_t stage = smem_u32(stat_smem + (es * 4 + ew) * EPI_STAGE_FLOATS);  // [32 rows][8 x 16 B], piece ^= row & 7
          float* bsm = stat_smem + 4 * ES * EPI_STAGE_FLOATS + (es * 4 + ew) * 256;
          const int my_pix = valid ? static_cast<int>(out_row) : -1;
          const int sub = lane >> 3, piece = lane & 7;
          int pix[8];  // pixel (output row) of the 8 staged rows this lane copies out: rows i*4 + sub
#pragma unroll
          for (int i = 0; i < 8; ++i) pix[i] = __shfl_sync(0xffffffffu, my_pix, i * 4 + sub);
          const uint32_t own = stage + lane * 128;
          const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(acc_col);
          if (p.out_mode == 2) {
            // split-K: raw fp32 partial sums [split][M][Cout] (bias / residual / statistics in the finalize pass)
            float* wsb = p.ws + static_cast<long long>(split) * p.M_total * p.Cout;
#pragma unroll 1
            for (int j = 0; j < BN / 32; ++j) {
              const int col0 = n_idx * BN + j * 32;
              if (col0 >= p.Cout) break;
              if constexpr (ES > 1) {
                if ((j % ES) != es) continue;
              }
              uint32_t r[32];
              tmem_ld_32x32b_x32(taddr0 + j * 32, r);
              tmem_ld_wait();
#pragma unroll
              for (int v = 0; v < 8; ++v)
                sts_v4(own + ((v ^ (lane & 7)) << 4), r[v * 4], r[v * 4 + 1], r[v * 4 + 2], r[v * 4 + 3]);
              __syncwarp();
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int rr = i * 4 + sub;
                const uint4 v4 = lds_v4(stage + rr * 128 + ((piece ^ (rr & 7)) << 4));
                if (pix[i] >= 0)
                  *reinterpret_cast<uint4*>(wsb + static_cast<long long>(pix[i]) * p.Cout + col0 + piece * 4) = v4;
              }
              __syncwarp();
            }
            return;
          }
          constexpr int NP = BN / 64;
          if (TAIL && tail_role == 1) {
            // accumulator order: float4 (jp, q) of row r at [(jp * 16 + q) * 128 + r] -- a warp stores / loads 512 contiguous bytes
            float4* buf = reinterpret_cast<float4*>(p.tail_buf) + static_cast<long long>(tail_slot) * (32 * BN);
#pragma unroll 1
            for (int jp = 0; jp < NP; ++jp) {
              if (n_idx * BN + jp * 64 >= p.Cout || (ES > 1 && (jp % ES) != es)) continue;
              uint32_t r[64];
              {
                uint32_t(&r0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[0]);
                uint32_t(&r1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[32]);
                tmem_ld_32x32b_x32(taddr0 + jp * 64, r0);
                tmem_ld_32x32b_x32(taddr0 + jp * 64 + 32, r1);
                tmem_ld_wait();
              }
#pragma unroll
              for (int q = 0; q < 16; ++q)
                __stcg(buf + (jp * 16 + q) * 128 + row, make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                                                   __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3])));
            }
            __threadfence();
            __syncwarp();
            if (lane == 0) st_release_gpu(p.tail_flags + tail_slot * 8 + es * 4 + ew, 1u);
            return;
          }
          if (TAIL && tail_role == 2) {
            if (lane == 0) {
              for (int pt = 0; pt < tail_np; ++pt) {
                const unsigned int* flag = p.tail_flags + (tail_slot + pt) * 8 + es * 4 + ew;
                const long long t0 = clock64();
                while (ld_acquire_gpu(flag) == 0u) {
                  if (clock64() - t0 > 4000000000LL) __trap();  // a scheduling bug traps instead of hanging the GPU
                }
              }
            }
            __syncwarp();
          }
          if (p.bias) {
#pragma unroll
            for (int c = lane; c < BN; c += 32) bsm[c] = (n_idx * BN + c < p.Cout) ? __ldg(p.bias + n_idx * BN + c) : 0.f;
            __syncwarp();
          }
          __half* outb = reinterpret_cast<__half*>(p.out);
          float4 st[NP];   // fused GroupNorm statistics of this warp's 32 rows: (sum, sumsq) of columns 2l, 2l+1 per pair
          uint4 rpre[8];   // residual of the NEXT 64-column pair, in flight while the current one is processed
          auto load_res = [&](int jp) {
            const int c0 = n_idx * BN + jp * 64;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              rpre[i] = make_uint4(0u, 0u, 0u, 0u);
              if (pix[i] >= 0 && c0 < p.Cout)
                rpre[i] = *reinterpret_cast<const uint4*>(p.residual + static_cast<long long>(pix[i]) * p.ldr + c0 + piece * 8);
            }
          };
          if constexpr (ES == 1) {
            if (p.residual) load_res(0);
          }
#pragma unroll
          for (int jp = 0; jp < NP; ++jp) {
            const int col0 = n_idx * BN + jp * 64;
            st[jp] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (col0 < p.Cout && (ES == 1 || (jp % ES) == es)) {
              if constexpr (ES > 1) {
                if (p.residual) load_res(jp);  // no prefetch registers: the second warp set hides the latency instead
              }
              if (p.residual) {  // coalesced: 8 lanes x 16 B per row, 4 rows per instruction -> staged by row
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const int rr = i * 4 + sub;
                  sts_v4(stage + rr * 128 + ((piece ^ (rr & 7)) << 4), rpre[i].x, rpre[i].y, rpre[i].z, rpre[i].w);
                }
                __syncwarp();
              }
              uint32_t r[64];
              {
                uint32_t(&r0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[0]);
                uint32_t(&r1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[32]);
                tmem_ld_32x32b_x32(taddr0 + jp * 64, r0);
                tmem_ld_32x32b_x32(taddr0 + jp * 64 + 32, r1);
                if constexpr (ES == 1) {
                  if (p.residual && jp + 1 < NP) load_res(jp + 1);
                }
                tmem_ld_wait();
              }
              if (TAIL && tail_role == 2) {
                for (int pt = 0; pt < tail_np; ++pt) {
                  const float4* buf = reinterpret_cast<const float4*>(p.tail_buf) + static_cast<long long>(tail_slot + pt) * (32 * BN);
#pragma unroll
                  for (int q0 = 0; q0 < 16; q0 += 4) {  // four loads in flight at a time: 16 registers, not 64
                    float4 t[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) t[u] = __ldcg(buf + (jp * 16 + q0 + u) * 128 + row);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                      const int q = q0 + u;
                      r[4 * q] = __float_as_uint(__uint_as_float(r[4 * q]) + t[u].x);
                      r[4 * q + 1] = __float_as_uint(__uint_as_float(r[4 * q + 1]) + t[u].y);
                      r[4 * q + 2] = __float_as_uint(__uint_as_float(r[4 * q + 2]) + t[u].z);
                      r[4 * q + 3] = __float_as_uint(__uint_as_float(r[4 * q + 3]) + t[u].w);
                    }
                    asm volatile("" ::: "memory");
                  }
                }
              }
#pragma unroll
              for (int v = 0; v < 8; ++v) {  // 8 columns = one 16-byte piece of the staged row
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(r[v * 8 + e]);
                if (p.bias) {
                  const float4 b0 = *reinterpret_cast<const float4*>(bsm + jp * 64 + v * 8);
                  const float4 b1 = *reinterpret_cast<const float4*>(bsm + jp * 64 + v * 8 + 4);
                  f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
                  f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
                }
                const uint32_t slot = own + ((v ^ (lane & 7)) << 4);
                if (p.residual) {
                  const uint4 rv = lds_v4(slot);
                  const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 t = __half22float2(rh[e]);
                    f[2 * e] += t.x;
                    f[2 * e + 1] += t.y;
                  }
                }
                uint32_t o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const __half2 hh = __floats2half2_rn(f[2 * e], f[2 * e + 1]);
                  o[e] = valid ? *reinterpret_cast<const uint32_t*>(&hh) : 0u;  // rows outside the image count as zeros
                }
                sts_v4(slot, o[0], o[1], o[2], o[3]);
              }
              __syncwarp();
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int rr = i * 4 + sub;
                const uint4 v4 = lds_v4(stage + rr * 128 + ((piece ^ (rr & 7)) << 4));
                if (pix[i] >= 0)
                  *reinterpret_cast<uint4*>(outb + static_cast<long long>(pix[i]) * p.ldo + col0 + piece * 8) = v4;
              }
              if (p.gn_part) {
                // statistics of the fp16-ROUNDED stored values: lane l sums columns 2l, 2l+1 over the warp's rows
                float4 h0 = make_float4(0.f, 0.f, 0.f, 0.f), h1 = make_float4(0.f, 0.f, 0.f, 0.f);  // rows 0-15 / 16-31
#pragma unroll
                for (int rr = 0; rr < 32; ++rr) {
                  const uint32_t w = lds_u32(stage + rr * 128 + (((lane >> 2) ^ (rr & 7)) << 4) + ((lane & 3) << 2));
                  const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w));
                  float4& hacc = (rr < 16) ? h0 : h1;
                  hacc.x += t.x;
                  hacc.y = fmaf(t.x, t.x, hacc.y);
                  hacc.z += t.y;
                  hacc.w = fmaf(t.y, t.y, hacc.w);
                }
                if (p.gn_mode == 2) {
                  // 16-pixel x 8-image tiles: half warp h of warp ew holds image n0 + 2*ew + h of spatial tile sp
                  const int per = p.tiles_h * p.tiles_w;
                  const int sp = m_in % per;
                  const int img = n0 + 2 * ew;
                  if (2 * ew < p.TN && img < p.NB)
                    *reinterpret_cast<float4*>(p.gn_part + part_index(static_cast<long long>(img) * per + sp) * p.Cout + col0 + 2 * lane) = h0;
                  if (2 * ew + 1 < p.TN && img + 1 < p.NB)
                    *reinterpret_cast<float4*>(p.gn_part + part_index(static_cast<long long>(img + 1) * per + sp) * p.Cout + col0 + 2 * lane) = h1;
                } else {
                  st[jp] = make_float4(h0.x + h1.x, h0.y + h1.y, h0.z + h1.z, h0.w + h1.w);
                }
              }
              __syncwarp();
            }
          }
          if (TAIL && tail_role == 2) {  // every partial of this warp's rows has been read: the flags are zero again for the next launch
            __syncwarp();
            if (lane == 0)
              for (int pt = 0; pt < tail_np; ++pt) p.tail_flags[(tail_slot + pt) * 8 + es * 4 + ew] = 0u;
          }
          if (p.gn_part && p.gn_mode == 1) {
            // one partial per (M tile, column): the four warps' sums are folded in a fixed order through the (now idle)
            // staging buffers, so k2_gn_finalize reads a quarter of what per-warp partials would cost
            if constexpr (ES == 1) {
              float4* mine = reinterpret_cast<float4*>(stat_smem + ew * EPI_STAGE_FLOATS);
#pragma unroll
              for (int jp = 0; jp < NP; ++jp) mine[jp * 32 + lane] = st[jp];
              named_bar_sync(1, 128);
              if (ew < NP) {
                const int col0 = n_idx * BN + ew * 64;
                if (col0 < p.Cout) {
                  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                  for (int w = 0; w < 4; ++w) {
                    const float4 t = reinterpret_cast<const float4*>(stat_smem + w * EPI_STAGE_FLOATS)[ew * 32 + lane];
                    acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
                  }
                  *reinterpret_cast<float4*>(p.gn_part + part_index(m_in) * p.Cout + col0 + 2 * lane) = acc;
                }
              }
              named_bar_sync(1, 128);
            } else {
              // each warp set folds the pairs it owns (jp % ES == es) among its own four warps: barrier 1 + es
              float4* mine = reinterpret_cast<float4*>(stat_smem + (es * 4 + ew) * EPI_STAGE_FLOATS);
#pragma unroll
              for (int jp = 0; jp < NP; ++jp) mine[jp * 32 + lane] = st[jp];
              named_bar_sync(1 + es, 128);
              const int jp_f = es + ew * ES;  // warp ew of the set folds the set's ew-th pair
              if (jp_f < NP) {
                const int col0 = n_idx * BN + jp_f * 64;
                if (col0 < p.Cout) {
                  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                  for (int w = 0; w < 4; ++w) {
                    const float4 t =
                        reinterpret_cast<const float4*>(stat_smem + (es * 4 + w) * EPI_STAGE_FLOATS)[jp_f * 32 + lane];
                    acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
                  }
                  *reinterpret_cast<float4*>(p.gn_part + part_index(m_in) * p.Cout + col0 + 2 * lane) = acc;
                }
              }
              named_bar_sync(1 + es, 128);
            }
          }
          return;
        }
      }
#pragma unroll 1
      for (int j = 0; j < BN / CH; ++j) {
        const uint32_t taddr =
            tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(acc_col + j * CH);
        uint32_t r[CH];
        if constexpr (CH == 32) {
          tmem_ld_32x32b_x32(taddr, r);
        } else {
          tmem_ld_32x32b_x16(taddr, r);
        }
        tmem_ld_wait();
        const int col0 = n_idx * BN + j * CH;
        if (valid && col0 < p.Cout) {
          if (p.out_mode == 2) {
            // split-K: raw fp32 partial sums [split][M][Cout] into the workspace (bias/residual in the finalize pass)
            float* w = p.ws + (static_cast<long long>(split) * p.M_total + out_row) * p.Cout + col0;
            if (col0 + CH <= p.Cout) {
#pragma unroll
              for (int v = 0; v < CH / 4; ++v)
                *reinterpret_cast<uint4*>(w + v * 4) = make_uint4(r[v * 4], r[v * 4 + 1], r[v * 4 + 2], r[v * 4 + 3]);
            } else {
#pragma unroll
              for (int e = 0; e < CH; ++e)
                if (col0 + e < p.Cout) w[e] = __uint_as_float(r[e]);
            }
          } else if (p.out_mode == 0) {
            __half* orow = reinterpret_cast<__half*>(p.out) + out_row * p.ldo + col0;
            const __half* rrow = p.residual ? p.residual + out_row * p.ldr + col0 : nullptr;
            if (col0 + CH <= p.Cout) {
#pragma unroll
              for (int v = 0; v < CH / 8; ++v) {
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  f[e] = __uint_as_float(r[v * 8 + e]);
                  if (p.bias) f[e] += __ldg(p.bias + col0 + v * 8 + e);
                }
                if (rrow) {
                  uint4 rv = *reinterpret_cast<const uint4*>(rrow + v * 8);
                  const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    float2 t = __half22float2(rh[e]);
                    f[2 * e] += t.x;
                    f[2 * e + 1] += t.y;
                  }
                }
                uint4 ov;
                __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
                for (int e = 0; e < 4; ++e) oh[e] = __floats2half2_rn(f[2 * e], f[2 * e + 1]);
                *reinterpret_cast<uint4*>(orow + v * 8) = ov;
              }
            } else {
#pragma unroll
              for (int e = 0; e < CH; ++e) {
                if (col0 + e < p.Cout) {
                  float f = __uint_as_float(r[e]);
                  if (p.bias) f += __ldg(p.bias + col0 + e);
                  if (rrow) f += __half2float(rrow[e]);
                  orow[e] = __float2half_rn(f);
                }
              }
            }
          } else {
            // fp32 NCHW (UNet / MoVQ output heads)
            float* o = reinterpret_cast<float*>(p.out);
#pragma unroll
            for (int e = 0; e < CH; ++e) {
              if (col0 + e < p.Cout) {
                float f = __uint_as_float(r[e]);
                if (p.bias) f += __ldg(p.bias + col0 + e);
                o[((static_cast<long long>(n) * p.Cout + (col0 + e)) * p.H + y) * p.W + x] = f;
              }
            }
          }
        }
      }
}

template <int BN>
__global__ void __launch_bounds__(256, 1) conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tmem_full = empty_bar + C::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* stat_smem = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES + C::BAR_BYTES);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmB);
    for (int s = 0; s < 3; ++s)
      if (p.seg_taps[s]) tma_prefetch_desc(&p.tmA[s]);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc(tmem_ptr, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();
  pdl_launch();

  const int total_tiles = p.m_tiles * p.n_tiles * p.splits;

  if (warp_idx == 0) {
    // ===================================== TMA producer =====================================
    if (elect_one()) {  // one lane, and ptxas KNOWS it is one: UTCHMMA / UTMALDG operands need no per-lane waterfall loop
      int stage = 0;
      uint32_t ring_phase = 0;
      const uint32_t tx_bytes = p.a_box_bytes + C::B_STAGE_BYTES;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_idx = tile % p.m_tiles;
        const int n_idx = (tile / p.m_tiles) % p.n_tiles;
        const int split = tile / (p.m_tiles * p.n_tiles);
        int n0, y0, x0, phase;
        decode_m_tile(p, m_idx, n0, y0, x0, phase);
        const int kb = phase * p.num_k_chunks;  // up2: each phase has its own 4-tap weight block
        const int k0 = split * p.k_per_split;
        const int k1 = min(p.num_k_chunks, k0 + p.k_per_split);
        // position (segment, tap, channel chunk) of flattened K chunk k0
        int s = 0, rem = k0;
        while (rem >= p.seg_taps[s] * p.seg_kchunks[s]) {
          rem -= p.seg_taps[s] * p.seg_kchunks[s];
          ++s;
        }
        int tap = rem / p.seg_kchunks[s];
        int c = rem - tap * p.seg_kchunks[s];
        for (int kc = k0; kc < k1; ++kc) {
          const int taps = p.seg_taps[s];
          const int dy = (taps == 9) ? (tap / 3 - 1) : (taps == 4 ? (tap >> 1) + (phase >> 1) - 1 : 0);
          const int dx = (taps == 9) ? (tap % 3 - 1) : (taps == 4 ? (tap & 1) + (phase & 1) - 1 : 0);
          mbar_wait(&empty_bar[stage], ring_phase ^ 1);
          uint8_t* sA = smem + stage * C::STAGE_BYTES;
          uint8_t* sB = sA + A_STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
          tma_load_4d(sA, &p.tmA[s], &full_bar[stage], c * BK, x0 + dx, y0 + dy, n0);
          tma_load_3d(sB, &p.tmB, &full_bar[stage], (kb + kc) * BK, n_idx * BN, p.w_batched ? n0 : 0);
          if (++stage == C::STAGES) {
            stage = 0;
            ring_phase ^= 1;
          }
          if (++c == p.seg_kchunks[s]) {
            c = 0;
            if (++tap == taps) {
              tap = 0;
              ++s;
            }
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================================== MMA issuer ========================================
    if (elect_one()) {  // one lane, and ptxas KNOWS it is one: UTCHMMA / UTMALDG operands need no per-lane waterfall loop
      constexpr uint32_t idesc = make_idesc_f16(BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        const int split = tile / (p.m_tiles * p.n_tiles);
        const int nk = min(p.num_k_chunks, (split + 1) * p.k_per_split) - split * p.k_per_split;
        for (int kc = 0; kc < nk; ++kc) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint64_t adesc = make_sw128_desc(a_addr);
          const uint64_t bdesc = make_sw128_desc(a_addr + A_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // +32 bytes (2 x 16 B units) per 16-element K step inside the 128 B swizzle row
            umma_f16(d_tmem, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2),
                     idesc, (kc | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // smem slot free once these MMAs retire
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);  // accumulator ready for the epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp_idx >= 4) {
    // ===================================== epilogue ==========================================
    const int ew = warp_idx - 4;  // == warp_idx % 4 -> TMEM lanes [32*ew, 32*ew+32)
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_idx = tile % p.m_tiles;
      const int n_idx = (tile / p.m_tiles) % p.n_tiles;
      const int split = tile / (p.m_tiles * p.n_tiles);
      int n0, y0, x0, phase;
      decode_m_tile(p, m_idx, n0, y0, x0, phase);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      epilogue_tile<BN>(p, tmem_base, acc * BN, ew, lane, n0, y0, x0, n_idx, split, m_idx, stat_smem, 0, phase);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// 2-CTA variant (tcgen05 cta_group::2): a CTA PAIR of one cluster computes a (256 pixel x BN channel) tile.
// CTA r of the pair owns M tile 2*pair+r: it loads ITS 128-row A box and HALF of the weight tile (BN/2 rows) per
// K chunk, so the pair pulls 32 KB + BN*128 B per chunk from L2 instead of 2 x (16 KB + BN*128 B) -- the 1-CTA
// kernel is L2->SM bandwidth bound at ~96 B/clk/SM.  The leader CTA (rank 0) issues the M=256 MMAs, which read both
// CTAs' shared memory and write each CTA's own TMEM; every TMA of the pair signals the LEADER's full barrier; the
// MMA commit multicasts the "stage free" / "accumulator ready" arrivals to both CTAs; the peer's epilogue warps arrive
// remotely on the leader's "accumulator drained" barrier.
// ------------------------------------------------------------------------------------------------
template <int BN, int ES = 1>
struct Cfg2 {
  static constexpr int B_STAGE_BYTES = (BN / 2) * BK * 2;   // this CTA's half of the weight tile
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  // the second epilogue warp set's staging / bias buffers (another EPI_BYTES) cost one pipeline stage
  static constexpr int STAGES = ((BN >= 256) ? 6 : (BN >= 192 ? 7 : 8)) - (ES - 1);
  static constexpr int TMEM_COLS = 512;                      // 2 accumulator buffers at columns 0 and 256
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + ES * EPI_BYTES + 1024;
  static constexpr int THREADS = 128 + 128 * ES;             // 4 control warps + ES sets of 4 epilogue warps
};

// TAIL: the launch uses the tail split (stream-K over the last partial wave); a separate instantiation, so that the default
// kernels stay instruction-for-instruction what they were before the feature existed.
template <int BN, int ES = 1, bool TAIL = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128 + 128 * ES, 1)
conv_gemm2_kernel(const __grid_constant__ ConvGemmParams p) {
  using C = Cfg2<BN, ES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tmem_full = empty_bar + C::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* stat_smem = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES + 256);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmB);
    for (int s = 0; s < 3; ++s)
      if (p.seg_taps[s]) tma_prefetch_desc(&p.tmA[s]);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);   // leader: one arrive.expect_tx covering BOTH CTAs' bytes
      mbar_init(&empty_bar[i], 1);  // one multicast commit per use
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8 * ES);  // 4 epilogue warps per set x 2 CTAs (only the leader's copy is waited on)
    }
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc2(tmem_ptr, C::TMEM_COLS);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs initialised before any remote arrival / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();
  pdl_launch();

  const int m_pairs = (p.m_tiles + 1) >> 1;
  // Work items of this launch: the tiles (x split-K parts), or -- tail split -- the tiles of the full waves followed by ONE
  // more wave in which the K loops of the remaining tail_count tiles, laid end to end (tail_count x num_k_chunks chunks), are
  // cut into equal spans of tail_kps chunks, one span per CTA pair (stream-K over the last wave only).  A span covers the
  // end of one tile and possibly the start of the next: up to two sub-items per pair (items tail_first + pair and
  // tail_first + num_pairs + pair).  The sub-item holding a tile's first chunk owns the tile (part 0); it is always the
  // LAST thing its pair does, and it only waits for sub-items that are the FIRST thing their pairs do in this wave: no cycles.
  const int total_tiles = TAIL ? p.tail_first + 2 * num_pairs : m_pairs * p.n_tiles * p.splits;
  // -> false: an empty sub-item (all three roles skip it); part < 0: ordinary item; nparts = K parts of the tile
  auto decode_item = [&](int item, int& tile, int& part, int& nparts, int& split, int& k0, int& k1) -> bool {
    if (TAIL && item >= p.tail_first) {
      const int idx = item - p.tail_first;
      const int sub = idx / num_pairs, i = idx - sub * num_pairs;
      const int nk = p.num_k_chunks, L = p.tail_kps;
      const int g0 = i * L, g1 = min(g0 + L, p.tail_count * nk);
      if (g0 >= g1) return false;
      int ta = g0 / nk;
      const int ka0 = g0 - ta * nk, ka1 = min(nk, ka0 + (g1 - g0));
      if (sub == 0) {
        k0 = ka0;
        k1 = ka1;
      } else {
        const int rest = (g1 - g0) - (ka1 - ka0);
        if (rest <= 0) return false;
        ++ta;
        k0 = 0;
        k1 = rest;
      }
      const int first_span = (ta * nk) / L;
      const int last_span = min(((ta + 1) * nk - 1) / L, (p.tail_count * nk - 1) / L);
      part = i - first_span;
      nparts = last_span - first_span + 1;
      tile = p.tail_first + ta;
      split = 0;
      return true;
    }
    tile = item;
    part = -1;
    nparts = 1;
    split = tile / (m_pairs * p.n_tiles);
    k0 = split * p.k_per_split;
    k1 = min(p.num_k_chunks, k0 + p.k_per_split);
    return true;
  };

  if (warp_idx == 0) {
    // ===================================== TMA producer (both CTAs) ==========================
    if (elect_one()) {  // one lane, and ptxas KNOWS it is one: UTCHMMA / UTMALDG operands need no per-lane waterfall loop
      int stage = 0;
      uint32_t ring_phase = 0;
      const uint32_t tx_bytes = 2u * (p.a_box_bytes + C::B_STAGE_BYTES);
      for (int item = pair; item < total_tiles; item += num_pairs) {
        int tile, part, nparts, split, k0, k1;
        if (!decode_item(item, tile, part, nparts, split, k0, k1)) continue;
        const int m_idx = (tile % m_pairs) * 2 + static_cast<int>(rank);
        const int n_idx = (tile / m_pairs) % p.n_tiles;
        int n0, y0, x0, phase;
        decode_m_tile(p, m_idx, n0, y0, x0, phase);  // m_idx == m_tiles (odd tail): n0 >= NB -> the box is all zero-fill
        const int kb = phase * p.num_k_chunks;  // up2: each phase has its own 4-tap weight block
        int s = 0, rem = k0;
        while (rem >= p.seg_taps[s] * p.seg_kchunks[s]) {
          rem -= p.seg_taps[s] * p.seg_kchunks[s];
          ++s;
        }
        int tap = rem / p.seg_kchunks[s];
        int c = rem - tap * p.seg_kchunks[s];
        for (int kc = k0; kc < k1; ++kc) {
          const int taps = p.seg_taps[s];
          const int dy = (taps == 9) ? (tap / 3 - 1) : (taps == 4 ? (tap >> 1) + (phase >> 1) - 1 : 0);
          const int dx = (taps == 9) ? (tap % 3 - 1) : (taps == 4 ? (tap & 1) + (phase & 1) - 1 : 0);
          mbar_wait(&empty_bar[stage], ring_phase ^ 1);
          uint8_t* sA = smem + stage * C::STAGE_BYTES;
          uint8_t* sB = sA + A_STAGE_BYTES;
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
          tma2_load_4d(sA, &p.tmA[s], &full_bar[stage], c * BK, x0 + dx, y0 + dy, n0);
          tma2_load_3d(sB, &p.tmB, &full_bar[stage], (kb + kc) * BK, n_idx * BN + static_cast<int>(rank) * (BN / 2),
                       p.w_batched ? n0 : 0);
          if (++stage == C::STAGES) {
            stage = 0;
            ring_phase ^= 1;
          }
          if (++c == p.seg_kchunks[s]) {
            c = 0;
            if (++tap == taps) {
              tap = 0;
              ++s;
            }
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================================== MMA issuer (leader CTA only) ======================
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = make_idesc_f16(256, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = pair; item < total_tiles; item += num_pairs) {
        int tile, part, nparts, split, k0, k1;
        if (!decode_item(item, tile, part, nparts, split, k0, k1)) continue;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * 256);
        const int nk = k1 - k0;
        for (int kc = 0; kc < nk; ++kc) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint64_t adesc = make_sw128_desc(a_addr);
          const uint64_t bdesc = make_sw128_desc(a_addr + A_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma2_f16(d_tmem, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2), idesc,
                      (kc | k) != 0 ? 1u : 0u);
          umma2_commit_mc(&empty_bar[stage], 0x3);  // the stage is free in BOTH CTAs once these MMAs retire
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma2_commit_mc(&tmem_full[acc], 0x3);  // each CTA's accumulator half is ready for its epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp_idx >= 4) {
    // ===================================== epilogue (both CTAs, own TMEM) ====================
    const int ew = (ES == 1) ? warp_idx - 4 : (warp_idx & 3);  // TMEM lane quarter
    const int es = (ES == 1) ? 0 : ((warp_idx - 4) >> 2);       // epilogue warp set
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t leader_empty0 = mapa_u32(smem_u32(&tmem_empty[0]), 0);
    const uint32_t leader_empty1 = mapa_u32(smem_u32(&tmem_empty[1]), 0);
    for (int item = pair; item < total_tiles; item += num_pairs) {
      int tile, part, nparts, split, k0, k1;
      if (!decode_item(item, tile, part, nparts, split, k0, k1)) continue;
      const int m_idx = (tile % m_pairs) * 2 + static_cast<int>(rank);
      const int n_idx = (tile / m_pairs) % p.n_tiles;
      int n0, y0, x0, phase;
      decode_m_tile(p, m_idx, n0, y0, x0, phase);
      // tail split: (tail_split - 1) hand-over slots per CTA half of a tail tile, one per K part >= 1
      const int tail_role = (part < 0 || nparts == 1) ? 0 : (part > 0 ? 1 : 2);
      const int tail_slot = (part < 0) ? 0 : ((tile - p.tail_first) * 2 + static_cast<int>(rank)) * (p.tail_split - 1) + max(part - 1, 0);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      epilogue_tile<BN, ES, TAIL>(p, tmem_base, acc * 256, ew, lane, n0, y0, x0, n_idx, split, m_idx, stat_smem, es, phase,
                                  tail_role, tail_slot, nparts - 1);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acc ? leader_empty1 : leader_empty0);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  cluster_sync_all();  // the peer's smem / TMEM must outlive every MMA of the pair
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, C::TMEM_COLS);
  }
}

template <int BN, int ES, bool TAIL>
int launch_pair(const ConvGemmParams& p, cudaStream_t stream) {
  using C = Cfg2<BN, ES>;
  static bool attr_set = false;
  if (!attr_set) {
    K2_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm2_kernel<BN, ES, TAIL>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const int m_pairs = (p.m_tiles + 1) / 2;
  const int total = m_pairs * p.n_tiles * p.splits;
  const int max_pairs = num_sms() / 2;
  const int pairs = (TAIL || total >= max_pairs) ? max_pairs : total;  // tail split: one K span per CTA pair of the device
  K2_CHECK_CUDA(launch_k(conv_gemm2_kernel<BN, ES, TAIL>, dim3(2 * pairs), dim3(C::THREADS), C::SMEM_BYTES, stream, p));
  return 0;
}
template <int BN>
int launch_bn2(const ConvGemmParams& p, cudaStream_t stream) {
  return p.tail_split > 1 ? launch_pair<BN, 1, true>(p, stream) : launch_pair<BN, 1, false>(p, stream);
}
// CTA-pair kernel with two epilogue warp sets (384 threads, one pipeline stage fewer); bit-identical to the one-set kernel
// (tests/test_gpu_conv_gemm.py::test_two_epilogue_sets_bit_identical).
template <int BN>
int launch_bn2e(const ConvGemmParams& p, cudaStream_t stream) {
  return p.tail_split > 1 ? launch_pair<BN, 2, true>(p, stream) : launch_pair<BN, 2, false>(p, stream);
}

// ------------------------------------------------------------------------------------------------
// Halo variant of the CTA-pair kernel for 3x3 convolutions: the 1-/2-CTA kernels above fetch the SAME activation
// pixels nine times (one shifted 128-pixel box per tap), and L2->SM operand bandwidth (~64 B/clk/SM) is what bounds
// them.  Here an M tile is 8 wide x 16 high and ONE (8+2) x (16+2) halo box per 64-channel chunk is staged in shared
// memory; the nine taps are nine tcgen05 A descriptors into that box: a row of 8 output pixels is 8 consecutive 128 B
// halo rows (one 8-row core group), the next output row starts `pitch` halo pixels later, so the stride between core
// groups is pitch*128 B and tap (dy, dx) only moves the start address by ((1+dy)*pitch + 1+dx)*128 B.  Activation
// traffic per chunk drops from 9 x 16 KB to one box; the weight tiles (one per tap) are unchanged.
// ------------------------------------------------------------------------------------------------
template <int BN>
struct Cfg3 {
  static constexpr int A_STAGE = 36864;                    // 18 rows x (up to) 16 pixels x 128 B
  static constexpr int A_STAGES = 3;
  static constexpr int B_STAGE = (BN / 2) * BK * 2;
  static constexpr int B_STAGES = (BN >= 256) ? 4 : 5;
  static constexpr int BAR_OFF = A_STAGES * A_STAGE + B_STAGES * B_STAGE;
  static constexpr int STAT_OFF = BAR_OFF + 256;
  static constexpr int SMEM_BYTES = STAT_OFF + EPI_BYTES + 1024;
};

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
conv_gemm3_kernel(const __grid_constant__ ConvGemmParams p) {
  using C = Cfg3<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smemB = smem + C::A_STAGES * C::A_STAGE;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
  uint64_t* a_empty = a_full + C::A_STAGES;
  uint64_t* b_full = a_empty + C::A_STAGES;
  uint64_t* b_empty = b_full + C::B_STAGES;
  uint64_t* tmem_full = b_empty + C::B_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* stat_smem = reinterpret_cast<float*>(smem + C::STAT_OFF);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int pitch = p.halo_pitch;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmB);
    for (int s = 0; s < 3; ++s)
      if (p.seg_taps[s]) tma_prefetch_desc(&p.tmA[s]);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < C::A_STAGES; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < C::B_STAGES; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc2(tmem_ptr, 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();
  pdl_launch();

  const int m_pairs = (p.m_tiles + 1) >> 1;
  const int total_tiles = m_pairs * p.n_tiles;

  if (warp_idx == 0) {
    // ===================================== TMA producer (both CTAs) ==========================
    if (elect_one()) {  // one lane, and ptxas KNOWS it is one: UTCHMMA / UTMALDG operands need no per-lane waterfall loop
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      const uint32_t halo_bytes = static_cast<uint32_t>(18 * pitch * 128);
      for (int tile = pair; tile < total_tiles; tile += num_pairs) {
        const int m_idx = (tile % m_pairs) * 2 + static_cast<int>(rank);
        const int n_idx = tile / m_pairs;
        int n0, y0, x0, phase_unused;
        decode_m_tile(p, m_idx, n0, y0, x0, phase_unused);
        int kc_base = 0;
        for (int s = 0; s < 3; ++s) {
          const int taps = p.seg_taps[s];
          if (taps == 0) break;
          const int kch = p.seg_kchunks[s];
          for (int c = 0; c < kch; ++c) {
            mbar_wait(&a_empty[as], aph ^ 1);
            uint8_t* sA = smem + as * C::A_STAGE;
            if (taps == 9) {
              if (rank == 0) mbar_arrive_expect_tx(&a_full[as], 2u * halo_bytes);
              tma2_load_4d(sA, &p.tmA[s], &a_full[as], c * BK, x0 - 1, y0 - 1, n0);
            } else {
              if (rank == 0) mbar_arrive_expect_tx(&a_full[as], 2u * A_STAGE_BYTES);
              tma2_load_4d(sA, &p.tmA[s], &a_full[as], c * BK, x0, y0, n0);
            }
            if (++as == C::A_STAGES) {
              as = 0;
              aph ^= 1;
            }
            for (int tap = 0; tap < taps; ++tap) {
              mbar_wait(&b_empty[bs], bph ^ 1);
              if (rank == 0) mbar_arrive_expect_tx(&b_full[bs], 2u * C::B_STAGE);
              tma2_load_3d(smemB + bs * C::B_STAGE, &p.tmB, &b_full[bs], (kc_base + tap * kch + c) * BK,
                           n_idx * BN + static_cast<int>(rank) * (BN / 2), 0);
              if (++bs == C::B_STAGES) {
                bs = 0;
                bph ^= 1;
              }
            }
          }
          kc_base += taps * kch;
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================================== MMA issuer (leader CTA only) ======================
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = make_idesc_f16(256, BN, 0, 0);
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < total_tiles; tile += num_pairs) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * 256);
        uint32_t first = 1;
        for (int s = 0; s < 3; ++s) {
          const int taps = p.seg_taps[s];
          if (taps == 0) break;
          const int kch = p.seg_kchunks[s];
          for (int c = 0; c < kch; ++c) {
            mbar_wait(&a_full[as], aph);
            tc_fence_after();
            const uint32_t a_base = smem_u32(smem + as * C::A_STAGE);
            for (int tap = 0; tap < taps; ++tap) {
              mbar_wait(&b_full[bs], bph);
              tc_fence_after();
              uint64_t adesc;
              if (taps == 9) {
                const uint32_t start = a_base + static_cast<uint32_t>(((tap / 3) * pitch + (tap % 3)) * 128);
                adesc = make_sw128_desc_ex(start, static_cast<uint32_t>(pitch * 128), p.halo_bo ? (start >> 7) & 7u : 0u);
              } else {
                adesc = make_sw128_desc(a_base);
              }
              const uint64_t bdesc = make_sw128_desc(smem_u32(smemB + bs * C::B_STAGE));
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                umma2_f16(d_tmem, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2), idesc,
                          (first && k == 0) ? 0u : 1u);
              }
              first = 0;
              umma2_commit_mc(&b_empty[bs], 0x3);
              if (++bs == C::B_STAGES) {
                bs = 0;
                bph ^= 1;
              }
            }
            umma2_commit_mc(&a_empty[as], 0x3);
            if (++as == C::A_STAGES) {
              as = 0;
              aph ^= 1;
            }
          }
        }
        umma2_commit_mc(&tmem_full[acc], 0x3);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp_idx >= 4) {
    // ===================================== epilogue (both CTAs, own TMEM) ====================
    const int ew = warp_idx - 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t leader_empty0 = mapa_u32(smem_u32(&tmem_empty[0]), 0);
    const uint32_t leader_empty1 = mapa_u32(smem_u32(&tmem_empty[1]), 0);
    for (int tile = pair; tile < total_tiles; tile += num_pairs) {
      const int m_idx = (tile % m_pairs) * 2 + static_cast<int>(rank);
      const int n_idx = tile / m_pairs;
      int n0, y0, x0, phase_unused;
      decode_m_tile(p, m_idx, n0, y0, x0, phase_unused);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      epilogue_tile<BN>(p, tmem_base, acc * 256, ew, lane, n0, y0, x0, n_idx, 0, m_idx, stat_smem);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acc ? leader_empty1 : leader_empty0);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

template <int BN>
int launch_bn3(const ConvGemmParams& p, cudaStream_t stream) {
  using C = Cfg3<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    K2_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm3_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const int m_pairs = (p.m_tiles + 1) / 2;
  const int total = m_pairs * p.n_tiles;
  const int max_pairs = num_sms() / 2;
  const int pairs = total < max_pairs ? total : max_pairs;
  K2_CHECK_CUDA(launch_k(conv_gemm3_kernel<BN>, dim3(2 * pairs), dim3(256), C::SMEM_BYTES, stream, p));
  return 0;
}

// split-K second pass: out[m, n] = fp16( sum_s ws[s][m][n] (fixed order) + bias[n] + residual[m, n] ).
// Block = 32 column vectors (256 channels) x 8 row lanes over 16 consecutive rows; optionally also emits the GroupNorm
// partial statistics of its 16 rows (same format as the conv epilogue's, 16-row groups) via a shared-memory fold.
__global__ void __launch_bounds__(256) splitk_finalize_kernel(const float* __restrict__ ws, int splits, long long M,
                                                              int Cout, const float* __restrict__ bias,
                                                              const __half* __restrict__ residual, int ldr,
                                                              __half* __restrict__ out, int ldo, float2* __restrict__ gn_part) {
  __shared__ float red[8][32][17];
  const int vx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c0 = (blockIdx.y * 32 + vx) * 8;
  const long long rg = blockIdx.x;
  pdl_wait();
  pdl_launch();
  float st[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) st[e] = 0.f;
  if (c0 < Cout) {
    float bv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) bv[e] = bias ? __ldg(bias + c0 + e) : 0.f;
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const long long m = rg * 16 + ry + rr * 8;
      if (m >= M) continue;
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = bv[e];
      for (int s = 0; s < splits; ++s) {
        const float4* w = reinterpret_cast<const float4*>(ws + (static_cast<long long>(s) * M + m) * Cout + c0);
        const float4 a = __ldcs(w), b = __ldcs(w + 1);
        f[0] += a.x; f[1] += a.y; f[2] += a.z; f[3] += a.w; f[4] += b.x; f[5] += b.y; f[6] += b.z; f[7] += b.w;
      }
      if (residual) {
        const uint4 rv = __ldg(reinterpret_cast<const uint4*>(residual + m * ldr + c0));
        const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 t = __half22float2(rh[e]);
          f[2 * e] += t.x;
          f[2 * e + 1] += t.y;
        }
      }
      uint4 ov;
      __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        oh[e] = __floats2half2_rn(f[2 * e], f[2 * e + 1]);
        const float2 t = __half22float2(oh[e]);  // statistics of the ROUNDED values
        st[2 * e] += t.x;
        st[2 * e + 1] += t.y;
        st[8 + 2 * e] = fmaf(t.x, t.x, st[8 + 2 * e]);
        st[8 + 2 * e + 1] = fmaf(t.y, t.y, st[8 + 2 * e + 1]);
      }
      *reinterpret_cast<uint4*>(out + m * ldo + c0) = ov;
    }
  }
  if (gn_part == nullptr) return;
#pragma unroll
  for (int e = 0; e < 16; ++e) red[ry][vx][e] = st[e];
  __syncthreads();
  // thread -> (column vector vx2, channel e2 of it): sums the 8 row lanes in order
  const int vx2 = threadIdx.x >> 3, e2 = threadIdx.x & 7;
  const int c = (blockIdx.y * 32 + vx2) * 8 + e2;
  if (c < Cout) {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      s1 += red[q][vx2][e2];
      s2 += red[q][vx2][8 + e2];
    }
    gn_part[rg * Cout + c] = make_float2(s1, s2);
  }
}

template <int BN>
int launch_bn(const ConvGemmParams& p, cudaStream_t stream) {
  using C = Cfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    K2_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       C::SMEM_BYTES));
    attr_set = true;
  }
  int total = p.m_tiles * p.n_tiles * p.splits;
  int grid = total < num_sms() ? total : num_sms();
  K2_CHECK_CUDA(launch_k(conv_gemm_kernel<BN>, dim3(grid), dim3(256), C::SMEM_BYTES, stream, p));
  return 0;
}

}  // namespace

int launch_splitk_finalize(const float* ws, int splits, long long M, int Cout, const float* bias, const __half* residual,
                           int ldr, __half* out, int ldo, float2* gn_part, cudaStream_t stream) {
  dim3 grid(static_cast<unsigned int>((M + 15) / 16), (Cout / 8 + 31) / 32);
  K2_CHECK_CUDA(launch_k(splitk_finalize_kernel, grid, dim3(256), 0, stream, ws, splits, M, Cout, bias, residual, ldr, out, ldo,
                         gn_part));
  return 0;
}

int launch_conv_gemm(const ConvGemmParams& p, int BN, int epilogue_sets, cudaStream_t stream) {
  if (p.halo_pitch) {
    switch (BN) {
      case 128: return launch_bn3<128>(p, stream);
      case 192: return launch_bn3<192>(p, stream);
      case 256: return launch_bn3<256>(p, stream);
      default: return fail("conv_gemm: unsupported BN for the halo kernel");
    }
  }
  if (p.two_cta && epilogue_sets == 2) {  // two epilogue warp sets (384 threads): short-K GEMMs are epilogue-paced
    switch (BN) {
      case 128: return launch_bn2e<128>(p, stream);
      case 192: return launch_bn2e<192>(p, stream);
      case 256: return launch_bn2e<256>(p, stream);
      default: return fail("conv_gemm: unsupported BN for the 2-CTA kernel");
    }
  }
  if (p.two_cta) {
    switch (BN) {
      case 128: return launch_bn2<128>(p, stream);
      case 192: return launch_bn2<192>(p, stream);
      case 256: return launch_bn2<256>(p, stream);
      default: return fail("conv_gemm: unsupported BN for the 2-CTA kernel");
    }
  }
  switch (BN) {
    case 16: return launch_bn<16>(p, stream);
    case 64: return launch_bn<64>(p, stream);
    case 128: return launch_bn<128>(p, stream);
    case 192: return launch_bn<192>(p, stream);
    case 256: return launch_bn<256>(p, stream);
    default: return fail("conv_gemm: unsupported BN");
  }
}

}  // namespace k2
