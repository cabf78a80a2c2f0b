// pipe_probe.cu -- ground-truth pipe rates on the B200 for the attention softmax design (profiles/README.md):
//   * MUFU.EX2 cycles per warp-wide instruction with 1 / 2 / 4 warps per SM sub-partition and 1..16 independent chains,
//   * the same with the softmax's companion instructions (FFMA before, FADD + F2FP after) interleaved,
//   * tcgen05.ld (32x32b.x32) and tcgen05.st (32x32b.x16) bytes per clock per SM with 4 / 8 warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu ; run on one GPU.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ILP independent dependent-chains of ex2 per thread
template <int ILP>
__global__ void mufu_chain(float* out, long long* cyc, int iters) {
  float v[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) v[i] = -1.f - 0.001f * (threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = ex2(v[i]) - 1.5f;  // FADD keeps the argument in range (one extra FP32 op per ex2)
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// the softmax inner pattern on 32 values per iteration: fma -> ex2 -> two running sums -> fp16 pack.
// PACKED: the scale-and-subtract and the sums as fma.rn.f32x2 / add.f32x2 (FFMA2 / FADD2: two elements per issue slot)
template <bool PACKED>
__global__ void softmax_like(const float* in, uint32_t* out, long long* cyc, int iters, float c, float m) {
  float s[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) s[i] = in[(threadIdx.x * 32 + i) & 1023];
  float l0 = 0.f, l1 = 0.f;
  uint32_t acc = 0;
  uint64_t c2, nm2, l2;
  asm("mov.b64 %0, {%1, %1};" : "=l"(c2) : "f"(c));
  asm("mov.b64 %0, {%1, %1};" : "=l"(nm2) : "f"(-m));
  asm("mov.b64 %0, {%1, %1};" : "=l"(l2) : "f"(0.f));
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t packed[16];
#pragma unroll
    for (int e = 0; e < 32; e += 2) {
      float p0, p1;
      if constexpr (PACKED) {
        uint64_t a, d, pp;
        float a0, a1;
        asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(s[e]), "f"(s[e + 1]));
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(c2), "l"(nm2));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(d));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(a0));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(a1));
        asm("mov.b64 %0, {%1, %2};" : "=l"(pp) : "f"(p0), "f"(p1));
        asm("add.f32x2 %0, %1, %2;" : "=l"(l2) : "l"(l2), "l"(pp));
      } else {
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(fmaf(s[e], c, -m)));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(fmaf(s[e + 1], c, -m)));
        l0 += p0;
        l1 += p1;
      }
      __half2 hh = __floats2half2_rn(p0, p1);
      packed[e >> 1] = *reinterpret_cast<uint32_t*>(&hh);
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) acc ^= packed[e];
#pragma unroll
    for (int i = 0; i < 32; ++i) s[i] += 0.25f;  // new "scores" for the next round (one FADD per element)
  }
  const long long t1 = clock64();
  if constexpr (PACKED) asm("mov.b64 {%0, %1}, %2;" : "=f"(l0), "=f"(l1) : "l"(l2));
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __float_as_uint(l0 + l1);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void tmem_rate(uint32_t* out, long long* cyc, int iters, int mode) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
        static_cast<uint32_t>(__cvta_generic_to_shared(&tptr))));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = tptr + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (mode == 0) {
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(base + ch * 32));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc ^= r[0] ^ r[31];
      }
    } else {
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
            "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(base + ch * 16),
            "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
            "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]));
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tptr));
}

template <int ILP>
static void run_chain(float* out, long long* cyc, int threads) {
  const int iters = 2000;
  mufu_chain<ILP><<<148, threads>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  const double warps_per_smsp = threads / 32 / 4.0;
  const double per_inst = static_cast<double>(h) / (iters * ILP) / (warps_per_smsp < 1 ? 1 : warps_per_smsp);
  printf("ex2 chains: %2d warps/CTA (%.2g per sub-partition), ILP %2d: %.2f cycles per warp-wide ex2 per sub-partition\n",
         threads / 32, warps_per_smsp, ILP, per_inst);
}

int main() {
  float* out;
  long long* cyc;
  float* in;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  cudaMalloc(&in, 1024 * sizeof(float));
  cudaMemset(in, 0, 1024 * sizeof(float));
  for (int threads : {128, 256, 512}) {
    run_chain<1>(out, cyc, threads);
    run_chain<2>(out, cyc, threads);
    run_chain<4>(out, cyc, threads);
    run_chain<8>(out, cyc, threads);
    run_chain<16>(out, cyc, threads);
  }
  for (int packed = 0; packed < 2; ++packed)
    for (int threads : {128, 256, 512}) {
      const int iters = 500;
      if (packed) softmax_like<true><<<148, threads>>>(in, reinterpret_cast<uint32_t*>(out), cyc, iters, 0.18f, 3.f);
      else softmax_like<false><<<148, threads>>>(in, reinterpret_cast<uint32_t*>(out), cyc, iters, 0.18f, 3.f);
      cudaDeviceSynchronize();
      long long h;
      cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
      const double wps = threads / 32 / 4.0;
      printf("softmax pattern (%s, ex2, sum, pack; +1 FADD): %2d warps/CTA: %.2f cycles per ex2 per sub-partition\n",
             packed ? "FFMA2 / FADD2" : "fma", threads / 32, static_cast<double>(h) / (iters * 32) / wps);
    }
  for (int mode = 0; mode < 2; ++mode) {
    for (int threads : {128, 256}) {
      const int iters = 1000;
      tmem_rate<<<148, threads>>>(reinterpret_cast<uint32_t*>(out), cyc, iters, mode);
      cudaError_t e = cudaDeviceSynchronize();
      long long h;
      cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
      const double bytes = static_cast<double>(iters) * 4 * (mode == 0 ? 32 : 16) * 4 * threads;
      printf("%s: %d warps: %.1f bytes per clock per SM (%s)\n", mode == 0 ? "tcgen05.ld 32x32b.x32 + wait" : "tcgen05.st 32x32b.x16",
             threads / 32, bytes / h, cudaGetErrorString(e));
    }
  }
  return 0;
}
