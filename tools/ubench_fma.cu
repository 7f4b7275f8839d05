// ubench_fma.cu — measured issue rates of the CUDA-core FMA forms on sm_100a (the ceiling of the depthwise kernels).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench_fma tools/ubench_fma.cu && gpurun_out/ubench_fma
//
// Every variant keeps 16 independent accumulator chains per thread so the 4-cycle FMA latency never limits issue;
// 8 warps per scheduler.  Prints FMA / clk / SM for: FFMA (3 registers), FFMA2 (fma.rn.f32x2), HFMA2 (fp16x2),
// HFMA2.BF16, and a FFMA2 + ALU mix (one shift per FMA instruction, the depthwise kernel's bf16->fp32 conversion).
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

constexpr int kChains = 16;
constexpr int kIters = 4096;

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

template <int MODE>
__global__ void __launch_bounds__(1024) fma_kernel(float* out, float seed, long long* cycles) {
  const long long t0 = clock64();
  if (MODE == 0) {  // FFMA
    float acc[kChains], a = seed, b = seed * 0.5f;
#pragma unroll
    for (int i = 0; i < kChains; ++i) acc[i] = seed + i;
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
      for (int i = 0; i < kChains; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i]) : "f"(a), "f"(b));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += acc[i];
    if (s == 123.456f) out[threadIdx.x] = s;
  } else if (MODE == 1 || MODE == 4) {  // FFMA2 (+ one ALU op per FFMA2 in mode 4)
    unsigned long long acc[kChains];
    const float2 av = make_float2(seed, seed * 0.25f), bv = make_float2(seed * 0.5f, seed);
    const unsigned long long a = *reinterpret_cast<const unsigned long long*>(&av), b = *reinterpret_cast<const unsigned long long*>(&bv);
    unsigned sh[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
      const float2 v = make_float2(seed + i, seed - i);
      acc[i] = *reinterpret_cast<const unsigned long long*>(&v);
      sh[i] = threadIdx.x + i;
    }
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
      for (int i = 0; i < kChains; ++i) {
        acc[i] = ffma2(a, b, acc[i]);
        if (MODE == 4) asm volatile("shl.b32 %0, %0, 1;" : "+r"(sh[i]));
      }
    }
    unsigned long long s = 0;
    unsigned t = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
      s ^= acc[i];
      t ^= sh[i];
    }
    if (s == 0x1234567ull && t == 77u) out[threadIdx.x] = 1.f;
  } else if (MODE == 5 || MODE == 6) {
    // the depthwise tap loop's operand pattern: 14 accumulators, each FFMA2 (5) / FFMA pair (6) reads a DIFFERENT input pair and
    // tap pair (no operand shared with its neighbours except the input of one column): is the 64-bit operand fetch the limit?
    float2 acc[14], in[13], w[7];
#pragma unroll
    for (int i = 0; i < 14; ++i) acc[i] = make_float2(seed + i, seed - i);
#pragma unroll
    for (int i = 0; i < 13; ++i) in[i] = make_float2(seed * (i + 1), seed * 0.5f * (i + 2));
#pragma unroll
    for (int i = 0; i < 7; ++i) w[i] = make_float2(seed * 0.25f * (i + 1), seed * 0.125f * (i + 3));
    for (int it = 0; it < kIters / 8; ++it) {
#pragma unroll
      for (int ix = 0; ix < 13; ++ix) {
#pragma unroll
        for (int pp = 0; pp < 7; ++pp) {
          const int dx = ix - pp;
          if (dx >= 0 && dx < 7) {
            if (MODE == 5) {
              unsigned long long ra = *reinterpret_cast<unsigned long long*>(&in[ix]), rb = *reinterpret_cast<unsigned long long*>(&w[dx]);
              unsigned long long rc = *reinterpret_cast<unsigned long long*>(&acc[pp]);
              rc = ffma2(ra, rb, rc);
              acc[pp] = *reinterpret_cast<float2*>(&rc);
              ra = *reinterpret_cast<unsigned long long*>(&in[(ix + 1) % 13]);
              rc = *reinterpret_cast<unsigned long long*>(&acc[7 + pp]);
              rc = ffma2(ra, rb, rc);
              acc[7 + pp] = *reinterpret_cast<float2*>(&rc);
            } else {
              asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[pp].x) : "f"(in[ix].x), "f"(w[dx].x));
              asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[pp].y) : "f"(in[ix].y), "f"(w[dx].y));
              asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[7 + pp].x) : "f"(in[(ix + 1) % 13].x), "f"(w[dx].x));
              asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[7 + pp].y) : "f"(in[(ix + 1) % 13].y), "f"(w[dx].y));
            }
          }
        }
      }
    }
    float sacc = 0.f;
#pragma unroll
    for (int i = 0; i < 14; ++i) sacc += acc[i].x + acc[i].y;
    if (sacc == 123.456f) out[threadIdx.x] = sacc;
  } else {  // HFMA2 fp16 (2) / bf16 (3)
    unsigned acc[kChains];
    const unsigned a = 0x3c003c00u, b = 0x38003800u;
#pragma unroll
    for (int i = 0; i < kChains; ++i) acc[i] = 0x3c003c00u + i;
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
      for (int i = 0; i < kChains; ++i) {
        if (MODE == 2) asm volatile("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b));
        else asm volatile("fma.rn.bf16x2 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b));
      }
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s ^= acc[i];
    if (s == 0x1234567u) out[threadIdx.x] = 1.f;
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(const char* name, int fma_per_instr, int sms, float* out, long long* cyc, double instr_scale = 1.0) {
  const int blocks = sms;  // one 1024-thread CTA per SM: 8 warps per scheduler
  fma_kernel<MODE><<<blocks, 1024>>>(out, 1.0001f, cyc);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  fma_kernel<MODE><<<blocks, 1024>>>(out, 1.0001f, cyc);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h[1024];
  cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < blocks; ++i) mean += double(h[i]);
  mean /= blocks;
  const double instr_per_sm = double(kIters) * kChains * 32.0 * instr_scale;  // warp instructions per SM (32 warps)
  const double fma_per_clk_sm = instr_per_sm * 32.0 * fma_per_instr / mean;
  printf("%-28s %8.3f ms  %10.0f cycles/CTA  %6.1f FMA/clk/SM  issue interval per scheduler %.2f clk  (%.1f TFLOP/s at this clock x %d SMs)\n",
         name, ms, mean, fma_per_clk_sm, mean / (double(kIters) * kChains * 8.0 * instr_scale),
         fma_per_clk_sm * 2.0 * sms * (mean / (ms * 1e-3)) * 1e-12, sms);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  float* out;
  long long* cyc;
  cudaMalloc(&out, 4096);
  cudaMalloc(&cyc, sizeof(long long) * 1024);
  printf("device %s, %d SMs\n", p.name, p.multiProcessorCount);
  run<0>("FFMA (3-register)", 1, p.multiProcessorCount, out, cyc);
  run<1>("FFMA2 (fma.rn.f32x2)", 2, p.multiProcessorCount, out, cyc);
  run<2>("HFMA2 (fp16x2)", 2, p.multiProcessorCount, out, cyc);
  run<3>("HFMA2.BF16 (bf16x2)", 2, p.multiProcessorCount, out, cyc);
  // tap-loop pattern: (kIters / 8) iterations x 98 instruction pairs; normalise to the kIters x kChains instruction count above
  run<5>("FFMA2, tap-loop operands", 2, p.multiProcessorCount, out, cyc, (kIters / 8) * 98.0 / (double(kIters) * kChains));
  run<6>("FFMA, tap-loop operands", 1, p.multiProcessorCount, out, cyc, (kIters / 8) * 196.0 / (double(kIters) * kChains));
  return 0;
}
