// MUFU.EX2 issue-rate microbenchmark for the forward softmax: how fast can ONE warp (per SM sub-partition) stream the
// exp2 block of a 64-score half row, and how does it change with a second warp in the same sub-partition?
//   mode 0: 64 independent MUFU.EX2 per iteration (pipe ceiling)
//   mode 1: the kernel's block: FFMA2 scale/shift, 2 x EX2, FADD2 row sum, F2FP pack (compiler's schedule)
//   mode 2: same arithmetic, all 64 EX2 first, sums / packs afterwards (software-pipelined in source)
//   mode 3: mode 1 plus the FMNMX3 row-maximum chain of the NEXT 128 scores interleaved
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/mufu_bench tools/mufu_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint32_t bf2(float lo, float hi) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ float fmax3(float a, float b, float c) { float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(float* out, long long* cyc, int iters, float sc, float nm) {
  float s[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) s[i] = -0.01f * (threadIdx.x + i);
  float t[128];
  if (MODE == 3) {
#pragma unroll
    for (int i = 0; i < 128; ++i) t[i] = 0.001f * (threadIdx.x * 3 + i);
  }
  uint64_t sa = pk(0.f, 0.f), sb = sa;
  uint32_t acc = 0;
  float mxacc = 0.f;
  const uint64_t sc2 = pk(sc, sc), nm2 = pk(nm, nm);
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 64; ++i) s[i] = ex2(s[i]);
    } else if (MODE == 1 || MODE == 3) {
      float mx[4];
      if (MODE == 3) { mx[0] = t[0]; mx[1] = t[1]; mx[2] = t[2]; mx[3] = t[3]; }
#pragma unroll
      for (int x = 0; x < 32; ++x) {
        float a0, a1;
        upk(fma2(pk(s[2 * x], s[2 * x + 1]), sc2, nm2), a0, a1);
        const float p0 = ex2(a0), p1 = ex2(a1);
        if (x & 1) sb = add2(sb, pk(p0, p1)); else sa = add2(sa, pk(p0, p1));
        acc ^= bf2(p0, p1);
        s[2 * x] = p0 * -1.f; s[2 * x + 1] = p1 * -1.f;     // keep the next iteration's inputs negative
        if (MODE == 3) {
          mx[x & 3] = fmax3(mx[x & 3], t[4 + 2 * x], t[5 + 2 * x]);
          mx[(x + 1) & 3] = fmax3(mx[(x + 1) & 3], t[64 + 2 * x], t[65 + 2 * x]);
        }
      }
      if (MODE == 3) { mxacc += fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])); t[it & 127] += 1.f; }
    } else {
      float p[64];
#pragma unroll
      for (int x = 0; x < 32; ++x) {
        float a0, a1;
        upk(fma2(pk(s[2 * x], s[2 * x + 1]), sc2, nm2), a0, a1);
        p[2 * x] = a0; p[2 * x + 1] = a1;
      }
#pragma unroll
      for (int i = 0; i < 64; ++i) p[i] = ex2(p[i]);
#pragma unroll
      for (int x = 0; x < 32; ++x) {
        if (x & 1) sb = add2(sb, pk(p[2 * x], p[2 * x + 1])); else sa = add2(sa, pk(p[2 * x], p[2 * x + 1]));
        acc ^= bf2(p[2 * x], p[2 * x + 1]);
        s[2 * x] = p[2 * x] * -1.f; s[2 * x + 1] = p[2 * x + 1] * -1.f;
      }
    }
  }
  long long t1 = clock64();
  float a, b; upk(add2(sa, sb), a, b);
  float r = a + b + mxacc;
#pragma unroll
  for (int i = 0; i < 64; ++i) r += s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r + (float)acc;
  if (threadIdx.x % 32 == 0) cyc[blockIdx.x * 8 + threadIdx.x / 32] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps_per_smsp) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 148 * 8 * 8);
  const int iters = 2000, threads = 128 * warps_per_smsp;
  k<MODE><<<148, threads>>>(out, cyc, iters, 0.147f, -0.5f);
  k<MODE><<<148, threads>>>(out, cyc, iters, 0.147f, -0.5f);
  cudaDeviceSynchronize();
  long long h[148 * 8];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double s = 0; int n = 0;
  for (int b = 0; b < 148; ++b) for (int w = 0; w < threads / 32; ++w) { s += h[b * 8 + w]; ++n; }
  printf("%-44s warps/SMSP=%d  cycles per 64-exp block per warp: %.1f  (%.2f cycles per MUFU warp-instr per SMSP)\n", name,
         warps_per_smsp, s / n / iters, s / n / iters / 64.0 / warps_per_smsp);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w = 1; w <= 2; ++w) {
    run<0>("0: 64 independent EX2", w);
    run<1>("1: softmax block, compiler schedule", w);
    run<2>("2: softmax block, EX2 burst then consumers", w);
    run<3>("3: softmax block + FMNMX3 chain of 128", w);
  }
  return 0;
}
