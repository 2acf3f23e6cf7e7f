// Decode-sized linear layers of the Phi-3 decoder layer around the attention op: y = epi(pro(x) W^T) for a
// handful of tokens (B <= 8: one token per sequence of a decode step), so the whole cost is streaming the bf16 weights
// from HBM once.  Replaces, for decode steps, Phi3RMSNorm.forward + nn.Linear (qkv_proj / o_proj / gate_up_proj /
// down_proj) + the SiLU-gate of Phi3MLP + the residual adds of Phi3DecoderLayer.forward
// (transformers/models/phi3/modeling_phi3.py:49-64, 295-335) -- ~15 launches per layer in the eager / cuBLAS path.
//   prologue  optional RMSNorm of x with weight gamma, with HF's rounding points (normalise in fp32, round to bf16,
//             multiply by the weight, round to bf16); every CTA recomputes it for its own copy of x (B x K bf16 in shared
//             memory, L2-resident source)
//   product   one CTA = 16 warps = 1 or 4 tiles of 16 output features (x2 weight rows for the gate/up pair of the SwiGLU
//             mode) x 16 or 4 K slices; a warp streams its 16 x K/slices weights with 16-byte no-allocate loads through
//             a register ring (64 registers = 8 KB per warp in flight) straight into mma.sync m16n8k16 A fragments
//             (M = weight rows, N = 8 = tokens, fp32 accumulation).  The K positions of a fragment are permuted so that
//             one thread's 8 consecutive weights / activations are one 16-byte load -- a dot product does not care.
//             (tcgen05 needs M = 128 x N >= 16 tiles fed through shared memory; at 8 tokens the legacy warp-level MMA with
//             register operands is the shorter path to the same HBM bound.)
//   epilogue  0: store; 1: + residual; 2: SwiGLU, out = silu(gate) * up with gate = rows [0,N), up = rows [N,2N) of W
// HBM-bound: algorithmic bytes = N*K*2 (x2 in mode 2) per call.
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>
#include "api_common.cuh"

namespace aki {

constexpr int SK_ROWS = 16;

__device__ __forceinline__ uint4 ldg_stream_v4(const void* p) {
  uint4 r;
  asm("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void mma_bf16_16x8x16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                                 uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// 16 warps = TILES row tiles (16 weight rows each; x2 in the SwiGLU mode) x (16 / TILES) K slices.
template <int TILES, int MODE>
__global__ void __launch_bounds__(512)
skinny_linear_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_stride, const __nv_bfloat16* __restrict__ w,
                     const __nv_bfloat16* __restrict__ gamma, float eps, const __nv_bfloat16* __restrict__ residual,
                     int64_t res_stride, __nv_bfloat16* __restrict__ y, int64_t y_stride, int B, int N, int K) {
  constexpr int WARPS = 16, KSPLIT = WARPS / TILES;
  extern __shared__ __align__(16) uint8_t sk_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, c = lane & 3;
  const int row_bytes = K * 2 + 64;            // +64: the two rows of a quarter-warp's 16-byte reads hit disjoint banks
  __shared__ float inv_rms[8];
  __shared__ float red[WARPS][8];
  // ---- this warp: row tile `rt` of the CTA, K slice `ks`.  Its weights go through a ring of R 32-wide K chunks held in
  // registers: half a ring is reloaded for K + 32 R as soon as the tensor core has consumed it, so 4 - 8 KB per warp stay
  // in flight, and the reloads of one weight row are 64 R / 2 contiguous bytes issued back to back (single 64-byte
  // requests per row, spread in time, measured 35 % slower: DRAM page locality).  The first R chunks are requested before anything else -- weights do not depend on the previous kernel,
  // so with programmatic dependent launch they stream in while the producer of x is still draining.
  const int rt = warp / KSPLIT, ks = warp % KSPLIT;
  const int n0 = (blockIdx.x * TILES + rt) * SK_ROWS;
  const int k_per_warp = K / KSPLIT;                 // multiple of 64
  const int kb = ks * k_per_warp;
  const bool live = n0 < N;
  const __nv_bfloat16* w_lo = w + (size_t)((live ? n0 : 0) + g) * K + kb + 8 * c;
  const __nv_bfloat16* w_hi = w_lo + (size_t)8 * K;
  constexpr int NB = (MODE == 2) ? 4 : 2;            // uint4 per chunk: lo, hi rows (+ the up rows)
  constexpr int R = (MODE == 2) ? 4 : 8;             // 64 registers of weights per thread either way
  uint4 ring[R][NB];
  auto load = [&](uint4 (&d)[NB], int k0) {
    d[0] = ldg_stream_v4(w_lo + k0);
    d[1] = ldg_stream_v4(w_hi + k0);
    if (MODE == 2) {
      d[2] = ldg_stream_v4(w_lo + (size_t)N * K + k0);
      d[3] = ldg_stream_v4(w_hi + (size_t)N * K + k0);
    }
  };
  if (live) {
#pragma unroll
    for (int u = 0; u < R; ++u)
      if (32 * u < k_per_warp) load(ring[u], 32 * u);
  }
  pdl_wait();          // x / residual come from the previous kernel
  pdl_trigger();
  // ---- stage x (B x K) into shared memory in ONE pass over global memory (every CTA reads all of x: keep it to one
  // pass and to few CTAs); the RMSNorm is then applied in place
  {
    float ss[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) ss[b] = 0.f;
    for (int k8 = tid; k8 < K / 8; k8 += WARPS * 32) {
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        uint4 u = make_uint4(0, 0, 0, 0);
        if (b < B) u = *reinterpret_cast<const uint4*>(x + (size_t)b * x_stride + k8 * 8);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(h[e]); ss[b] += f.x * f.x + f.y * f.y; }
        *reinterpret_cast<uint4*>(sk_smem + (size_t)b * row_bytes + k8 * 16) = u;
      }
    }
    if (gamma) {
#pragma unroll
      for (int b = 0; b < 8; ++b) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss[b] += __shfl_xor_sync(0xffffffffu, ss[b], o);
        if (lane == 0) red[warp][b] = ss[b];
      }
      __syncthreads();
      if (tid < 8) {
        float t = 0.f;
        for (int wv = 0; wv < WARPS; ++wv) t += red[wv][tid];
        inv_rms[tid] = rsqrtf(t / (float)K + eps);
      }
      __syncthreads();
      for (int k8 = tid; k8 < K / 8; k8 += WARPS * 32) {          // the elements this thread wrote itself
        const uint4 gm = *reinterpret_cast<const uint4*>(gamma + k8 * 8);
        const __nv_bfloat162* gw = reinterpret_cast<const __nv_bfloat162*>(&gm);
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          uint4* p = reinterpret_cast<uint4*>(sk_smem + (size_t)b * row_bytes + k8 * 16);
          uint4 u = *p;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
          const float r = inv_rms[b];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __bfloat1622float2(h[e]);
            // Phi3RMSNorm: weight * (x * rsqrt(mean(x^2) + eps)).to(bf16)   -- two roundings
            h[e] = __hmul2(gw[e], __floats2bfloat162_rn(f.x * r, f.y * r));
          }
          *p = u;
        }
      }
    }
  }
  __syncthreads();
  const uint8_t* xs = sk_smem + (size_t)g * row_bytes + (size_t)(kb + 8 * c) * 2;
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};
  if (live) {
    constexpr int H = R / 2;                         // reload half a ring at a time: 64 H contiguous bytes per weight row
    for (int k0 = 0; k0 < k_per_warp; k0 += 32 * R) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
#pragma unroll
        for (int u = half * H; u < half * H + H; ++u) {
          const int kk = k0 + 32 * u;
          if (kk < k_per_warp) {
            const uint4 xv = *reinterpret_cast<const uint4*>(xs + (size_t)kk * 2);
            const uint4 (&d)[NB] = ring[u];
            mma_bf16_16x8x16(acc, d[0].x, d[1].x, d[0].y, d[1].y, xv.x, xv.y);
            mma_bf16_16x8x16(acc, d[0].z, d[1].z, d[0].w, d[1].w, xv.z, xv.w);
            if (MODE == 2) {
              mma_bf16_16x8x16(acc2, d[2].x, d[3].x, d[2].y, d[3].y, xv.x, xv.y);
              mma_bf16_16x8x16(acc2, d[2].z, d[3].z, d[2].w, d[3].w, xv.z, xv.w);
            }
          }
        }
#pragma unroll
        for (int u = half * H; u < half * H + H; ++u) {
          const int kk = k0 + 32 * u + 32 * R;
          if (kk < k_per_warp) load(ring[u], kk);
        }
      }
    }
  }
  // ---- sum the K slices' partial 16 x 8 tiles, epilogue
  __syncthreads();                                  // x in shared memory is dead: reuse it for the partials
  float* part = reinterpret_cast<float*>(sk_smem);  // [WARPS][2][16][8]
  {
    float* p = part + (size_t)warp * 256;
    p[g * 8 + 2 * c] = acc[0]; p[g * 8 + 2 * c + 1] = acc[1];
    p[(g + 8) * 8 + 2 * c] = acc[2]; p[(g + 8) * 8 + 2 * c + 1] = acc[3];
    if (MODE == 2) {
      p[128 + g * 8 + 2 * c] = acc2[0]; p[128 + g * 8 + 2 * c + 1] = acc2[1];
      p[128 + (g + 8) * 8 + 2 * c] = acc2[2]; p[128 + (g + 8) * 8 + 2 * c + 1] = acc2[3];
    }
  }
  __syncthreads();
  for (int idx = tid; idx < 128 * TILES; idx += WARPS * 32) {
    const int tile = idx >> 7, e = idx & 127, r = e >> 3, b = e & 7;       // output feature n + r of token b
    const int n = (blockIdx.x * TILES + tile) * SK_ROWS;
    float s = 0.f, s2 = 0.f;
#pragma unroll
    for (int kk = 0; kk < KSPLIT; ++kk) {
      s += part[(size_t)(tile * KSPLIT + kk) * 256 + e];
      if (MODE == 2) s2 += part[(size_t)(tile * KSPLIT + kk) * 256 + 128 + e];
    }
    if (b < B && n < N) {
      float out = s;
      if (MODE == 2) {
        // Phi3MLP: down_proj(up * silu(gate)) with gate / up rounded to bf16 by gate_up_proj first
        const float gq = __bfloat162float(__float2bfloat16(s)), uq = __bfloat162float(__float2bfloat16(s2));
        const float act = __bfloat162float(__float2bfloat16(gq / (1.f + __expf(-gq))));
        out = uq * act;
      } else if (MODE == 1) {
        out = __bfloat162float(__float2bfloat16(s)) + __bfloat162float(residual[(size_t)b * res_stride + n + r]);
      }
      y[(size_t)b * y_stride + n + r] = __float2bfloat16(out);
    }
  }
}

template <int TILES, int MODE>
static int launch_skinny(const void* x, int64_t xs, const void* w, const void* gamma, float eps, const void* res, int64_t rs,
                         void* y, int64_t ys, int B, int N, int K, cudaStream_t st) {
  const size_t smem = (size_t)8 * (K * 2 + 64);
  auto kern = skinny_linear_kernel<TILES, MODE>;
  if (smem > 48 * 1024) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_last_cuda_error(cudaGetErrorString(cudaGetLastError()));
      return AKI_ERR_CUDA;
    }
  }
  const int tiles = N / SK_ROWS;
  // programmatic dependent launch (api_common.cuh): the CTAs start, and request their first weights, while the previous
  // kernel of the stream drains; the kernel waits before it touches x / residual / y
  launch_pdl(kern, dim3((tiles + TILES - 1) / TILES), dim3(512), smem, st, static_cast<const __nv_bfloat16*>(x), xs,
             static_cast<const __nv_bfloat16*>(w), static_cast<const __nv_bfloat16*>(gamma), eps,
             static_cast<const __nv_bfloat16*>(res), rs, static_cast<__nv_bfloat16*>(y), ys, B, N, K);
  return check_launch();
}

}  // namespace aki

using namespace aki;

extern "C" int aki_mma_skinny_linear(const void* x, int64_t x_stride, const void* w, const void* rms_weight, float rms_eps,
                                     const void* residual, int64_t residual_stride, void* y, int64_t y_stride, int B, int N,
                                     int K, int mode, aki_stream_t stream) {
  AKI_REQUIRE(x && w && y, AKI_ERR_NULL);
  AKI_REQUIRE(B >= 1 && B <= 8 && N > 0 && K > 0 && mode >= 0 && mode <= 2, AKI_ERR_BAD_SHAPE);
  AKI_REQUIRE(N % SK_ROWS == 0 && K % (16 * 64) == 0, AKI_ERR_UNSUPPORTED);   // K slices of a multiple of 64 elements
  AKI_REQUIRE((mode == 1) == (residual != nullptr), AKI_ERR_NULL);
  AKI_REQUIRE(x_stride % 8 == 0 && (size_t)8 * (K * 2 + 64) <= 200 * 1024, AKI_ERR_UNSUPPORTED);
  AKI_REQUIRE(aligned16(x) && aligned16(w) && (!rms_weight || aligned16(rms_weight)), AKI_ERR_MISALIGNED);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // one wave of CTAs: the fewest row tiles per CTA (1, 2, 4 or 16; the 16 warps split K 16 / 4 / ... / 1 ways) that keeps
  // the grid within the SM count -- every CTA stages all of x, and a second, partial wave would cost a whole pass
  static int sm_count[64] = {};                       // benign race: every writer stores the same value
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { set_last_cuda_error("cudaGetDevice failed"); return AKI_ERR_CUDA; }
  if (sm_count[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    sm_count[dev] = n;
  }
  const int sms = sm_count[dev];
  const int tiles = N / SK_ROWS;
  const int per = tiles <= sms ? 1 : tiles <= 2 * sms ? 2 : tiles <= 4 * sms ? 4 : 16;
#define AKI_SK(T, M) launch_skinny<T, M>(x, x_stride, w, rms_weight, rms_eps, residual, residual_stride, y, y_stride, B, N, K, st)
#define AKI_SK_MODE(M) (per == 1 ? AKI_SK(1, M) : per == 2 ? AKI_SK(2, M) : per == 4 ? AKI_SK(4, M) : AKI_SK(16, M))
  if (mode == 0) return AKI_SK_MODE(0);
  if (mode == 1) return AKI_SK_MODE(1);
  return AKI_SK_MODE(2);
#undef AKI_SK_MODE
#undef AKI_SK
}
