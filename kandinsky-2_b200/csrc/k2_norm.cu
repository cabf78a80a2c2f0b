// k2_norm.cu -- GroupNorm statistics + fused normalise / FiLM / SiLU / resample / concat / SpatialNorm.
//
// Replaces (reference file:line):
//   GroupNorm32.forward                 kandinsky2/model/nn.py:31-37   (fp32 statistics, eps 1e-5, 32 groups)
//   ResBlock FiLM + SiLU                kandinsky2/model/unet.py:209-216
//   Upsample / Downsample (h_upd,x_upd) kandinsky2/model/unet.py:67-77,105-107,198-203
//   torch.cat([h, hs.pop()], 1)         kandinsky2/model/text2im_model2_1.py:99  (read as two sources)
//   SpatialNorm.forward                 kandinsky2/vqgan/movq_modules.py:61-68   (eps 1e-6)
//
// Both kernels are HBM-bound: stats reads the tensor once, apply reads it once and writes it once
// (16-byte vector accesses, fully coalesced: a warp covers 512 contiguous bytes of one pixel row).
// The statistics are reduced in a fixed order (no float atomics), so results are bit-reproducible,
// which the multi-GPU == single-GPU test relies on.
#include "../../include/k2b200.h"
#include "k2_common.cuh"
#include "k2_internal.h"

namespace k2 {
namespace {

constexpr int ST_VX = 32;  // channel vectors (8 ch) per block row
constexpr int ST_PY = 8;   // pixel lanes

__device__ __forceinline__ const __half* src_ptr(const __half* s0, int C0, int ld0, const __half* s1, int ld1,
                                                 long long row, int c) {
  return (c < C0) ? (s0 + row * ld0 + c) : (s1 + row * ld1 + (c - C0));
}

__global__ void __launch_bounds__(256) gn_stats_kernel(const __half* __restrict__ s0, int C0, int ld0,
                                                       const __half* __restrict__ s1, int C1, int ld1, int HW,
                                                       int groups, float eps, int chunks, float* __restrict__ stats,
                                                       float* __restrict__ partial, unsigned int* __restrict__ counters) {
  extern __shared__ float sm[];
  // layout: red[ST_PY][ST_VX][16] | chan[ST_VX*8][2] | bins[groups][2]
  float* red = sm;
  float* chan = red + ST_PY * ST_VX * 16;
  float* bins = chan + ST_VX * 8 * 2;
  __shared__ bool is_last;

  const int C = C0 + C1;
  const int CV = C / 8;
  const int cpg = C / groups;
  const int n = blockIdx.y;
  const int chunk = blockIdx.x;
  const int vx = threadIdx.x % ST_VX;
  const int py = threadIdx.x / ST_VX;
  const int per = (HW + chunks - 1) / chunks;
  const int p0 = chunk * per;
  const int p1 = min(HW, p0 + per);

  for (int i = threadIdx.x; i < groups * 2; i += blockDim.x) bins[i] = 0.f;
  __syncthreads();

  for (int ct = 0; ct * ST_VX < CV; ++ct) {
    const int v = ct * ST_VX + vx;
    float a[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) a[e] = 0.f;
    if (v < CV) {
      const int c = v * 8;
      for (int p = p0 + py; p < p1; p += ST_PY) {
        const long long row = static_cast<long long>(n) * HW + p;
        uint4 raw = *reinterpret_cast<const uint4*>(src_ptr(s0, C0, ld0, s1, ld1, row, c));
        const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 f = __half22float2(h2[e]);
          a[2 * e] += f.x;
          a[2 * e + 1] += f.y;
          a[8 + 2 * e] += f.x * f.x;
          a[8 + 2 * e + 1] += f.y * f.y;
        }
      }
    }
    float* mine = red + (py * ST_VX + vx) * 16;
#pragma unroll
    for (int e = 0; e < 16; ++e) mine[e] = a[e];
    __syncthreads();
    if (py == 0) {
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        float s = 0.f;
        for (int q = 0; q < ST_PY; ++q) s += red[(q * ST_VX + vx) * 16 + e];
        // chan[(vx*8 + ch)][0/1]
        chan[(vx * 8 + (e & 7)) * 2 + (e >> 3)] = s;
      }
    }
    __syncthreads();
    // group owners: fold this tile's channels into the block bins in channel order
    const int c_lo = ct * ST_VX * 8;
    const int c_hi = min(C, c_lo + ST_VX * 8);
    const int g_first = c_lo / cpg;
    const int g_last = (c_hi - 1) / cpg;
    const int g = g_first + threadIdx.x;
    if (g <= g_last) {
      const int lo = max(c_lo, g * cpg);
      const int hi = min(c_hi, (g + 1) * cpg);
      float s = bins[g * 2], q = bins[g * 2 + 1];
      for (int c = lo; c < hi; ++c) {
        s += chan[(c - c_lo) * 2];
        q += chan[(c - c_lo) * 2 + 1];
      }
      bins[g * 2] = s;
      bins[g * 2 + 1] = q;
    }
    __syncthreads();
  }

  float* my_partial = partial + (static_cast<long long>(n) * chunks + chunk) * groups * 2;
  for (int i = threadIdx.x; i < groups * 2; i += blockDim.x) my_partial[i] = bins[i];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int prev = atomicAdd(&counters[n], 1u);
    is_last = (prev == static_cast<unsigned int>(chunks - 1));
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    const double cnt = static_cast<double>(HW) * cpg;
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
      double s = 0.0, q = 0.0;
      const float* pp = partial + static_cast<long long>(n) * chunks * groups * 2 + g * 2;
      for (int ch = 0; ch < chunks; ++ch) {
        s += static_cast<double>(__ldcg(pp + static_cast<long long>(ch) * groups * 2));
        q += static_cast<double>(__ldcg(pp + static_cast<long long>(ch) * groups * 2 + 1));
      }
      const double mean = s / cnt;
      double var = q / cnt - mean * mean;
      if (var < 0.0) var = 0.0;
      stats[(static_cast<long long>(n) * groups + g) * 2] = static_cast<float>(mean);
      stats[(static_cast<long long>(n) * groups + g) * 2 + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    }
    if (threadIdx.x == 0) counters[n] = 0;  // ready for the next launch
  }
}

struct ApplyParams {
  const __half* s0;
  const __half* s1;
  int C0, ld0, C1, ld1;
  int NB, H, W, groups;
  const float* stats;
  const float* gamma;
  const float* beta;
  const float* film;  // rows (scale | shift), stride film_ld, or null
  int film_ld;
  int act, resample;
  __half* y;
  int ldy;
  __half* xres;
  int ldx;
  const float* zq;  // [NB, zh, zw, 4] or null
  int zh, zw;
  const float* sn_w;  // [C, 10]
};

__device__ __forceinline__ void load8(const __half* p, float (&f)[8]) {
  uint4 raw = *reinterpret_cast<const uint4*>(p);
  const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float2 t = __half22float2(h2[e]);
    f[2 * e] = t.x;
    f[2 * e + 1] = t.y;
  }
}
__device__ __forceinline__ void store8(__half* p, const float (&f)[8]) {
  uint4 ov;
  __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
  for (int e = 0; e < 4; ++e) oh[e] = __floats2half2_rn(f[2 * e], f[2 * e + 1]);
  *reinterpret_cast<uint4*>(p) = ov;
}

// per-channel affine a*x+b that folds mean/rstd/gamma/beta/FiLM (and, for SpatialNorm, the zq modulation)
struct Affine8 {
  float a[8], b[8];
};

__device__ __forceinline__ void make_affine(const ApplyParams& p, int n, int c0, int y_in, int x_in, Affine8& af) {
  const int C = p.C0 + p.C1;
  const int cpg = C / p.groups;
  float zq[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.zq) {
    const int zy = (y_in * p.zh) / p.H;
    const int zx = (x_in * p.zw) / p.W;
    const float4 z = *reinterpret_cast<const float4*>(p.zq + ((static_cast<long long>(n) * p.zh + zy) * p.zw + zx) * 4);
    zq[0] = z.x; zq[1] = z.y; zq[2] = z.z; zq[3] = z.w;
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = c0 + e;
    const int g = c / cpg;
    const float mean = __ldg(p.stats + (static_cast<long long>(n) * p.groups + g) * 2);
    const float rstd = __ldg(p.stats + (static_cast<long long>(n) * p.groups + g) * 2 + 1);
    float ga = __ldg(p.gamma + c) * rstd;
    float be = __ldg(p.beta + c) - mean * ga;
    if (p.film) {
      const float sc = 1.f + __ldg(p.film + static_cast<long long>(n) * p.film_ld + c);
      const float sh = __ldg(p.film + static_cast<long long>(n) * p.film_ld + C + c);
      ga *= sc;
      be = be * sc + sh;
    }
    if (p.zq) {
      const float* w = p.sn_w + static_cast<long long>(c) * 10;
      const float my = __ldg(w + 0) * zq[0] + __ldg(w + 1) * zq[1] + __ldg(w + 2) * zq[2] + __ldg(w + 3) * zq[3] + __ldg(w + 4);
      const float mb = __ldg(w + 5) * zq[0] + __ldg(w + 6) * zq[1] + __ldg(w + 7) * zq[2] + __ldg(w + 8) * zq[3] + __ldg(w + 9);
      ga *= my;
      be = be * my + mb;
    }
    af.a[e] = ga;
    af.b[e] = be;
  }
}

__global__ void __launch_bounds__(256) gn_apply_kernel(const ApplyParams p) {
  const int C = p.C0 + p.C1;
  const int CV = C / 8;
  const long long item = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  // work items: resample 0/2 -> input pixels; resample 1 -> output pixels
  const int Hw = (p.resample == 1) ? p.H / 2 : p.H;
  const int Ww = (p.resample == 1) ? p.W / 2 : p.W;
  const long long total = static_cast<long long>(p.NB) * Hw * Ww * CV;
  if (item >= total) return;
  const int v = static_cast<int>(item % CV);
  const long long pix = item / CV;
  const int x = static_cast<int>(pix % Ww);
  const int y = static_cast<int>((pix / Ww) % Hw);
  const int n = static_cast<int>(pix / (static_cast<long long>(Ww) * Hw));
  const int c0 = v * 8;

  if (p.resample == 0 || p.resample == 2) {
    const long long row = (static_cast<long long>(n) * p.H + y) * p.W + x;
    float f[8], o[8];
    load8(src_ptr(p.s0, p.C0, p.ld0, p.s1, p.ld1, row, c0), f);
    Affine8 af;
    make_affine(p, n, c0, y, x, af);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float t = f[e] * af.a[e] + af.b[e];
      o[e] = p.act ? silu_f(t) : t;
    }
    if (p.resample == 0) {
      store8(p.y + row * p.ldy + c0, o);
      if (p.xres) store8(p.xres + row * p.ldx + c0, f);
    } else {
      const int Wo = p.W * 2;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const long long orow = (static_cast<long long>(n) * p.H * 2 + (2 * y + dy)) * Wo + (2 * x + dx);
          store8(p.y + orow * p.ldy + c0, o);
          if (p.xres) store8(p.xres + orow * p.ldx + c0, f);
        }
    }
  } else {
    // 2x2 average pool of act(norm(x)) and of raw x
    float acc[8], accx[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = accx[e] = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int yi = 2 * y + dy, xi = 2 * x + dx;
        const long long row = (static_cast<long long>(n) * p.H + yi) * p.W + xi;
        float f[8];
        load8(src_ptr(p.s0, p.C0, p.ld0, p.s1, p.ld1, row, c0), f);
        Affine8 af;
        make_affine(p, n, c0, yi, xi, af);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float t = f[e] * af.a[e] + af.b[e];
          // the reference rounds GN+SiLU to fp16 before AvgPool2d (nn.py:32, unet.py:199-200)
          t = p.act ? silu_f(t) : t;
          acc[e] += __half2float(__float2half_rn(t));
          accx[e] += f[e];
        }
      }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      acc[e] *= 0.25f;
      accx[e] *= 0.25f;
    }
    const long long orow = (static_cast<long long>(n) * Hw + y) * Ww + x;
    store8(p.y + orow * p.ldy + c0, acc);
    if (p.xres) store8(p.xres + orow * p.ldx + c0, accx);
  }
}

static int stats_chunks(int NB, int HW) {
  int chunks = (4 * num_sms() + NB - 1) / NB;
  int maxc = (HW + 63) / 64;
  if (chunks > maxc) chunks = maxc;
  if (chunks < 1) chunks = 1;
  return chunks;
}

}  // namespace
}  // namespace k2

using namespace k2;

extern "C" {

long long k2_gn_scratch_floats(int NB, int HW, int groups) {
  // [0, 1024): one arrival counter per image (fixed location: they must stay zero between launches
  // whatever geometry the previous launch had); then the partials for the worst-case chunk count
  int maxc = (HW + 63) / 64;
  long long c = 4LL * 160;
  if (c > maxc) c = maxc;
  if (c < 1) c = 1;
  return 1024 + static_cast<long long>(NB) * c * groups * 2;
}

int k2_gn_stats(const void* src0, int C0, int ld0, const void* src1, int C1, int ld1, int NB, int HW, int groups,
                float eps, float* stats, float* scratch, k2_stream_t stream) {
  const int C = C0 + C1;
  K2_REQUIRE(src0 && C0 > 0 && C0 % 8 == 0 && C1 % 8 == 0, "gn_stats: channels must be multiples of 8");
  K2_REQUIRE(C % groups == 0, "gn_stats: C % groups != 0");
  K2_REQUIRE(groups <= 256, "gn_stats: at most 256 groups");
  K2_REQUIRE(src1 || C1 == 0, "gn_stats: src1 null with C1 > 0");
  const int chunks = stats_chunks(NB, HW);
  K2_REQUIRE(NB <= 1024, "gn_stats: at most 1024 images per launch");
  float* partial = scratch + 1024;
  long long partial_floats = static_cast<long long>(NB) * chunks * groups * 2;
  K2_REQUIRE(1024 + partial_floats <= k2_gn_scratch_floats(NB, HW, groups), "gn_stats: scratch too small");
  unsigned int* counters = reinterpret_cast<unsigned int*>(scratch);  // zero-initialised by the caller, self-resetting
  dim3 grid(chunks, NB);
  size_t smem = (ST_PY * ST_VX * 16 + ST_VX * 8 * 2 + groups * 2) * sizeof(float);
  gn_stats_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(src0), C0, ld0, reinterpret_cast<const __half*>(src1), C1, ld1, HW, groups,
      eps, chunks, stats, partial, counters);
  K2_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int k2_gn_apply(const void* src0, int C0, int ld0, const void* src1, int C1, int ld1, int NB, int H, int W,
                int groups, const float* stats, const float* gamma, const float* beta, const float* film, int film_ld,
                int act, int resample, void* y, int ldy, void* xres, int ldx, const float* zq, int zh, int zw,
                const float* sn_w, k2_stream_t stream) {
  const int C = C0 + C1;
  K2_REQUIRE(src0 && y && stats && gamma && beta, "gn_apply: null pointer");
  K2_REQUIRE(C0 % 8 == 0 && C1 % 8 == 0 && C % groups == 0, "gn_apply: bad channel counts");
  K2_REQUIRE(resample >= 0 && resample <= 2, "gn_apply: resample in {0,1,2}");
  K2_REQUIRE(resample != 1 || (H % 2 == 0 && W % 2 == 0), "gn_apply: avg-pool needs even H, W");
  K2_REQUIRE(!zq || sn_w, "gn_apply: zq without sn_w");
  ApplyParams p;
  p.s0 = reinterpret_cast<const __half*>(src0);
  p.s1 = reinterpret_cast<const __half*>(src1);
  p.C0 = C0; p.ld0 = ld0; p.C1 = C1; p.ld1 = ld1;
  p.NB = NB; p.H = H; p.W = W; p.groups = groups;
  p.stats = stats; p.gamma = gamma; p.beta = beta; p.film = film; p.film_ld = film_ld;
  p.act = act; p.resample = resample;
  p.y = reinterpret_cast<__half*>(y); p.ldy = ldy;
  p.xres = reinterpret_cast<__half*>(xres); p.ldx = ldx;
  p.zq = zq; p.zh = zh; p.zw = zw; p.sn_w = sn_w;
  const int Hw = (resample == 1) ? H / 2 : H;
  const int Ww = (resample == 1) ? W / 2 : W;
  const long long total = static_cast<long long>(NB) * Hw * Ww * (C / 8);
  const long long blocks = (total + 255) / 256;
  gn_apply_kernel<<<static_cast<unsigned int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  K2_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // extern "C"
