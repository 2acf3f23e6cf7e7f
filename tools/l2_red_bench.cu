// Chip-wide throughput of fp32 vector reductions into an L2-resident buffer (the dQ accumulation pattern of the
// backward kernel: every warp instruction adds 512 contiguous bytes), next to plain 16-byte stores and 16-byte loads
// of the same pattern.  usage: l2_red_bench <mode 0 red.v4.f32 | 1 st.v4 | 2 ld.v4 | 3 red.f32 scalar | 4 even CTAs ld.v4, odd CTAs red.v4 | 5 every thread alternates ld.v4 / red.v4> [footprint MB]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__global__ void __launch_bounds__(128) k(float* buf, size_t n_chunks, int iters, int mode, float* sink) {
  // chunk = 48 KB (one dQ tile: 24 column quads x 128 rows x 16 B); a CTA walks chunks like the kernel walks query tiles
  const int r = threadIdx.x;
  float acc = 0.f;
  if (mode == 4) mode = (blockIdx.x & 1) ? 0 : 2;
  for (int it = 0; it < iters; ++it) {
    const size_t chunk = ((size_t)blockIdx.x * 7 + (size_t)it * 131) % n_chunks;
    float* dst = buf + chunk * 12288 + r * 4;
#pragma unroll
    for (int c = 0; c < 24; ++c) {
      float* p = dst + c * 512;
      if (mode == 0) asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p), "f"(1.0f) : "memory");
      else if (mode == 1) asm volatile("st.global.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p), "f"(1.0f) : "memory");
      else if (mode == 5) {
        if (c & 1) asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p), "f"(1.0f) : "memory");
        else { float4 v; asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p)); acc += v.x + v.y + v.z + v.w; }
      }
      else if (mode == 2) { float4 v; asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p)); acc += v.x + v.y + v.z + v.w; }
      else { for (int e = 0; e < 4; ++e) asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(p + e), "f"(1.0f) : "memory"); }
    }
  }
  if (acc == 12345.f) *sink = acc;
}

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  const size_t mb = argc > 2 ? atoi(argv[2]) : 24;
  const size_t n_chunks = mb * 1024 * 1024 / 49152;
  float *buf, *sink;
  CK(cudaMalloc(&buf, n_chunks * 49152)); CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(buf, 0, n_chunks * 49152));
  const int grid = 148 * 4, iters = 400;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k<<<grid, 128>>>(buf, n_chunks, iters, mode, sink);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep == 2) printf("mode %d footprint %zu MB: %.3f ms, %.2f TB/s of payload\n", mode, mb, ms, (double)grid * iters * 49152 / (ms * 1e-3) / 1e12);
  }
  return 0;
}
