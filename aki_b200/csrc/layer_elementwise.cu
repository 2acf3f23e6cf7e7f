// The element-wise work of a Phi-3 decoder layer around the attention op at PREFILL size: one pass over HBM
// per step instead of the 8 + 3 + 1 ATen kernels HF's eager modules launch for it
//   add_rmsnorm   h = residual + x (bf16 add, rounded like the eager `residual + hidden_states`), y = Phi3RMSNorm(h)
//                 = weight * (h_fp32 * rsqrt(mean(h_fp32^2) + eps)).to(bf16)          (modeling_phi3.py:49-64, 317-335)
//   swiglu        y = up * silu(gate) for gate_up = [gate | up] per token                  (Phi3MLP.forward, :295-306)
// A torch-profiler run of the AKI-4B prefill (B = 8, T = 655) put 42 % of the kernel time into those ATen element-wise
// kernels (tools/prefill_profile.py); the GEMMs between them stay cuBLAS.
// HBM-bound: add_rmsnorm reads x (+ residual) and writes y (+ h): 2-4 x M*K*2 bytes; swiglu reads 2 and writes 1 x M*N*2.
// One warp per token row, the row lives in registers between the two passes (K <= 4096), 16-byte accesses.
#include <cuda_bf16.h>
#include <math.h>
#include "api_common.cuh"

namespace aki {

constexpr int NORM_WARPS = 8;
constexpr int NORM_MAX_CHUNKS = 16;      // 16-byte chunks per lane: K <= 32 * 8 * 16 = 4096

__global__ void __launch_bounds__(NORM_WARPS * 32)
add_rmsnorm_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_stride, const __nv_bfloat16* residual,
                   int64_t res_stride, const __nv_bfloat16* __restrict__ weight, float eps, __nv_bfloat16* h_out,
                   int64_t h_stride, __nv_bfloat16* __restrict__ y, int64_t y_stride, int M, int K) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * NORM_WARPS + (threadIdx.x >> 5);
  pdl_wait();
  pdl_trigger();
  if (row >= M) return;
  const int n_chunks = K >> 8;             // K / (32 lanes * 8 elements)
  uint4 v[NORM_MAX_CHUNKS];
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < NORM_MAX_CHUNKS; ++c) {
    if (c < n_chunks) {
      const int k = (c * 32 + lane) * 8;
      uint4 a = *reinterpret_cast<const uint4*>(x + (size_t)row * x_stride + k);
      if (residual) {
        const uint4 r = *reinterpret_cast<const uint4*>(residual + (size_t)row * res_stride + k);
        __nv_bfloat162* a2 = reinterpret_cast<__nv_bfloat162*>(&a);
        const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
        for (int e = 0; e < 4; ++e) {                      // residual + hidden_states in bf16: fp32 add, one rounding
          const float2 fa = __bfloat1622float2(a2[e]), fr = __bfloat1622float2(r2[e]);
          a2[e] = __floats2bfloat162_rn(fr.x + fa.x, fr.y + fa.y);
        }
        if (h_out) *reinterpret_cast<uint4*>(h_out + (size_t)row * h_stride + k) = a;
      }
      v[c] = a;
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(h2[e]); ss += f.x * f.x + f.y * f.y; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float r = rsqrtf(ss / (float)K + eps);
#pragma unroll
  for (int c = 0; c < NORM_MAX_CHUNKS; ++c) {
    if (c < n_chunks) {
      const int k = (c * 32 + lane) * 8;
      const uint4 g = *reinterpret_cast<const uint4*>(weight + k);
      const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&g);
      uint4 out = v[c];
      __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&out);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(o2[e]);
        o2[e] = __hmul2(g2[e], __floats2bfloat162_rn(f.x * r, f.y * r));   // two roundings, as Phi3RMSNorm
      }
      *reinterpret_cast<uint4*>(y + (size_t)row * y_stride + k) = out;
    }
  }
}

// one thread per 8 output features of one token
__global__ void __launch_bounds__(256)
swiglu_kernel(const __nv_bfloat16* __restrict__ gate_up, int64_t gu_stride, __nv_bfloat16* __restrict__ y, int64_t y_stride,
              int M, int N) {
  pdl_wait();
  pdl_trigger();
  const int per_row = N >> 3;
  const long long total = (long long)M * per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(i / per_row), k = (int)(i % per_row) * 8;
    const uint4 g = *reinterpret_cast<const uint4*>(gate_up + (size_t)row * gu_stride + k);
    const uint4 u = *reinterpret_cast<const uint4*>(gate_up + (size_t)row * gu_stride + N + k);
    const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&g);
    const __nv_bfloat162* u2 = reinterpret_cast<const __nv_bfloat162*>(&u);
    uint4 out;
    __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&out);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 fg = __bfloat1622float2(g2[e]);
      // silu on a bf16 tensor: fp32 inside, rounded to bf16; then a bf16 multiply with up
      const __nv_bfloat162 act = __floats2bfloat162_rn(fg.x / (1.f + expf(-fg.x)), fg.y / (1.f + expf(-fg.y)));
      o2[e] = __hmul2(u2[e], act);
    }
    *reinterpret_cast<uint4*>(y + (size_t)row * y_stride + k) = out;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Next-token cross-entropy over the LM head's logits: what `AKI.forward` gets back from
// Phi3ForCausalLM(labels=...) (codes/open_flamingo/src/aki.py:125-130; HF: logits.float(), shift by one, CrossEntropyLoss
// with ignore_index -100) without the fp32 copy of the (B,T,32064) logits and the ~10 ATen kernels around it.
//   forward : row (b,t), t < T-1, target = labels[b,t+1]: loss = logsumexp(logits[b,t,:]) - logits[b,t,target] in fp32
//             from the bf16 logits (0 for ignored targets); the lse is kept for the backward
//   backward: dlogits[b,t,:] = (exp(logits - lse) - onehot(target)) * scale, rounded to bf16 (what autograd hands the
//             bf16 logits through the .float() cast); rows without a target get zeros.  scale = dloss / n_valid lives in
//             device memory (no host read).
// One CTA per row, one pass over the row per kernel (online softmax), 16-byte loads: HBM-bound, V*2 bytes read per row
// forward, V*2 read + V*2 written backward.
constexpr int CE_THREADS = 256;

__device__ __forceinline__ void ce_combine(float& m, float& s, float m2, float s2) {
  const float mn = fmaxf(m, m2);
  s = (m == -INFINITY ? 0.f : s * expf(m - mn)) + (m2 == -INFINITY ? 0.f : s2 * expf(m2 - mn));
  m = mn;
}

__global__ void __launch_bounds__(CE_THREADS)
cross_entropy_fwd_kernel(const __nv_bfloat16* __restrict__ logits, int64_t stride_b, int64_t stride_t,
                         const int64_t* __restrict__ labels, int64_t lab_stride_b, int T, int V, long long ignore_index,
                         float* __restrict__ row_loss, float* __restrict__ row_lse) {
  const int t = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const long long target = (t + 1 < T) ? labels[(size_t)b * lab_stride_b + t + 1] : ignore_index;
  const size_t out = (size_t)b * T + t;
  if (target == ignore_index || target < 0 || target >= V) {          // uniform per CTA
    if (tid == 0) { row_loss[out] = 0.f; row_lse[out] = 0.f; }
    return;
  }
  const __nv_bfloat16* row = logits + (size_t)b * stride_b + (size_t)t * stride_t;
  float m = -INFINITY, s = 0.f;
  for (int c = tid; c < V / 8; c += CE_THREADS) {
    const uint4 u = *reinterpret_cast<const uint4*>(row + c * 8);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
    float f[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) { const float2 v = __bfloat1622float2(h[e]); f[2 * e] = v.x; f[2 * e + 1] = v.y; }
    float cm = f[0];
#pragma unroll
    for (int e = 1; e < 8; ++e) cm = fmaxf(cm, f[e]);
    const float mn = fmaxf(m, cm);
    float cs = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) cs += expf(f[e] - mn);
    s = (m == -INFINITY ? 0.f : s * expf(m - mn)) + cs;
    m = mn;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    ce_combine(m, s, m2, s2);
  }
  __shared__ float sm[CE_THREADS / 32], ss[CE_THREADS / 32];
  if ((tid & 31) == 0) { sm[tid >> 5] = m; ss[tid >> 5] = s; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < CE_THREADS / 32; ++w) ce_combine(m, s, sm[w], ss[w]);
    const float lse = m + logf(s);
    row_lse[out] = lse;
    row_loss[out] = lse - __bfloat162float(row[target]);
  }
}

__global__ void __launch_bounds__(CE_THREADS)
cross_entropy_bwd_kernel(const __nv_bfloat16* __restrict__ logits, int64_t stride_b, int64_t stride_t,
                         const int64_t* __restrict__ labels, int64_t lab_stride_b, int T, int V, long long ignore_index,
                         const float* __restrict__ row_lse, const float* __restrict__ scale_dev,
                         __nv_bfloat16* __restrict__ dlogits, int64_t d_stride_b, int64_t d_stride_t) {
  const int t = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const long long target = (t + 1 < T) ? labels[(size_t)b * lab_stride_b + t + 1] : ignore_index;
  const bool live = !(target == ignore_index || target < 0 || target >= V);
  const __nv_bfloat16* row = logits + (size_t)b * stride_b + (size_t)t * stride_t;
  __nv_bfloat16* drow = dlogits + (size_t)b * d_stride_b + (size_t)t * d_stride_t;
  const float lse = live ? row_lse[(size_t)b * T + t] : 0.f;
  const float scale = live ? *scale_dev : 0.f;
  for (int c = tid; c < V / 8; c += CE_THREADS) {
    uint4 u = make_uint4(0, 0, 0, 0);
    if (live) {
      u = *reinterpret_cast<const uint4*>(row + c * 8);
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 v = __bfloat1622float2(h[e]);
        float p0 = expf(v.x - lse), p1 = expf(v.y - lse);
        if (c * 8 + 2 * e == target) p0 -= 1.f;
        if (c * 8 + 2 * e + 1 == target) p1 -= 1.f;
        h[e] = __floats2bfloat162_rn(p0 * scale, p1 * scale);
      }
    }
    *reinterpret_cast<uint4*>(drow + c * 8) = u;
  }
}

}  // namespace aki

using namespace aki;

extern "C" int aki_mma_add_rmsnorm(const void* x, int64_t x_stride, const void* residual, int64_t residual_stride,
                                   const void* weight, float eps, void* h_out, int64_t h_stride, void* y, int64_t y_stride,
                                   int M, int K, aki_stream_t stream) {
  AKI_REQUIRE(x && weight && y, AKI_ERR_NULL);
  AKI_REQUIRE(M > 0 && K > 0, AKI_ERR_BAD_SHAPE);
  AKI_REQUIRE(K % 256 == 0 && K <= 256 * NORM_MAX_CHUNKS, AKI_ERR_UNSUPPORTED);
  AKI_REQUIRE(!h_out || residual, AKI_ERR_BAD_SHAPE);       // h_out is the sum: meaningless without a residual
  AKI_REQUIRE(x_stride % 8 == 0 && y_stride % 8 == 0 && residual_stride % 8 == 0 && h_stride % 8 == 0, AKI_ERR_UNSUPPORTED);
  AKI_REQUIRE(aligned16(x) && aligned16(weight) && aligned16(y) && (!residual || aligned16(residual)) &&
                  (!h_out || aligned16(h_out)),
              AKI_ERR_MISALIGNED);
  launch_pdl(add_rmsnorm_kernel, dim3((M + NORM_WARPS - 1) / NORM_WARPS), dim3(NORM_WARPS * 32), 0,
             static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(x), x_stride,
             static_cast<const __nv_bfloat16*>(residual), residual_stride, static_cast<const __nv_bfloat16*>(weight), eps,
             static_cast<__nv_bfloat16*>(h_out), h_stride, static_cast<__nv_bfloat16*>(y), y_stride, M, K);
  return check_launch();
}

extern "C" int aki_mma_swiglu(const void* gate_up, int64_t gate_up_stride, void* y, int64_t y_stride, int M, int N,
                              aki_stream_t stream) {
  AKI_REQUIRE(gate_up && y, AKI_ERR_NULL);
  AKI_REQUIRE(M > 0 && N > 0, AKI_ERR_BAD_SHAPE);
  AKI_REQUIRE(N % 8 == 0 && gate_up_stride % 8 == 0 && y_stride % 8 == 0, AKI_ERR_UNSUPPORTED);
  AKI_REQUIRE(aligned16(gate_up) && aligned16(y), AKI_ERR_MISALIGNED);
  const long long total = (long long)M * (N / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  launch_pdl(swiglu_kernel, dim3((unsigned)blocks), dim3(256), 0, static_cast<cudaStream_t>(stream),
             static_cast<const __nv_bfloat16*>(gate_up), gate_up_stride, static_cast<__nv_bfloat16*>(y), y_stride, M, N);
  return check_launch();
}

extern "C" int aki_mma_cross_entropy_fwd(const void* logits, int64_t stride_b, int64_t stride_t, const int64_t* labels,
                                         int64_t labels_stride_b, int B, int T, int V, long long ignore_index,
                                         float* row_loss, float* row_lse, aki_stream_t stream) {
  AKI_REQUIRE(logits && labels && row_loss && row_lse, AKI_ERR_NULL);
  AKI_REQUIRE(B > 0 && T > 0 && V > 0 && B <= 65535, AKI_ERR_BAD_SHAPE);
  AKI_REQUIRE(V % 8 == 0 && stride_b % 8 == 0 && stride_t % 8 == 0, AKI_ERR_UNSUPPORTED);
  AKI_REQUIRE(aligned16(logits), AKI_ERR_MISALIGNED);
  cross_entropy_fwd_kernel<<<dim3(T, B), CE_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(logits), stride_b, stride_t, labels, labels_stride_b, T, V, ignore_index, row_loss,
      row_lse);
  return check_launch();
}

extern "C" int aki_mma_cross_entropy_bwd(const void* logits, int64_t stride_b, int64_t stride_t, const int64_t* labels,
                                         int64_t labels_stride_b, int B, int T, int V, long long ignore_index,
                                         const float* row_lse, const float* scale_dev, void* dlogits, int64_t d_stride_b,
                                         int64_t d_stride_t, aki_stream_t stream) {
  AKI_REQUIRE(logits && labels && row_lse && scale_dev && dlogits, AKI_ERR_NULL);
  AKI_REQUIRE(B > 0 && T > 0 && V > 0 && B <= 65535, AKI_ERR_BAD_SHAPE);
  AKI_REQUIRE(V % 8 == 0 && stride_b % 8 == 0 && stride_t % 8 == 0 && d_stride_b % 8 == 0 && d_stride_t % 8 == 0,
              AKI_ERR_UNSUPPORTED);
  AKI_REQUIRE(aligned16(logits) && aligned16(dlogits), AKI_ERR_MISALIGNED);
  cross_entropy_bwd_kernel<<<dim3(T, B), CE_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(logits), stride_b, stride_t, labels, labels_stride_b, T, V, ignore_index, row_lse,
      scale_dev, static_cast<__nv_bfloat16*>(dlogits), d_stride_b, d_stride_t);
  return check_launch();
}

// ------------------------------------------------------------------------------------------------------------------
// TRAINING layout of the same element-wise work (SURVEY 8 f-1 for the SFT step; reference precision amp_bf16,
// configs/sft.yaml:55): the residual stream and the norm weights are fp32, the GEMM operands bf16.
//   add_rmsnorm_amp_fwd  h' = h + float(a) (a = the bf16 output of o_proj / down_proj; skipped when NULL),
//                        r = rsqrt(mean(h'^2) + eps), x = bf16(w * (h' * r))      = residual add + Phi3RMSNorm + the
//                        autocast cast of the next Linear's input, one pass instead of ~9 ATen kernels
//   rmsnorm_amp_bwd      g = float(dx) * w;  dh = dh_out + r * g - h' * (r^3 / K) * sum_k(g_k h'_k);
//                        dw_partial[warp] += float(dx) * h' * r over the rows this warp walks (summed by the caller:
//                        deterministic, no atomics).  dh is the gradient of both h and (after a cast) a.
//   swiglu_bwd           d_up = d_out * silu(gate).bf16;  d_act = (d_out * up).bf16;  d_gate = d_act * silu'(gate)
// One warp per token row; the backward reads its row twice (second read from L1/L2).
namespace aki {

constexpr int AMP_MAX_CHUNKS = 12;       // 8-element chunks per lane: K <= 32 * 8 * 12 = 3072

__device__ __forceinline__ void ld8_f32(const float* p, float* f) {
  *reinterpret_cast<float4*>(f) = *reinterpret_cast<const float4*>(p);
  *reinterpret_cast<float4*>(f + 4) = *reinterpret_cast<const float4*>(p + 4);
}
__device__ __forceinline__ void st8_f32(float* p, const float* f) {
  *reinterpret_cast<float4*>(p) = *reinterpret_cast<const float4*>(f);
  *reinterpret_cast<float4*>(p + 4) = *reinterpret_cast<const float4*>(f + 4);
}
__device__ __forceinline__ void ld8_bf16(const __nv_bfloat16* p, float* f) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) { const float2 v = __bfloat1622float2(h[e]); f[2 * e] = v.x; f[2 * e + 1] = v.y; }
}
__device__ __forceinline__ void st8_bf16(__nv_bfloat16* p, const float* f) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

__global__ void __launch_bounds__(NORM_WARPS * 32)
add_rmsnorm_amp_fwd_kernel(const float* h_in, const __nv_bfloat16* __restrict__ a, const float* __restrict__ w, float eps,
                           float* h_out, __nv_bfloat16* __restrict__ x, float* __restrict__ r_out, int M, int K) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * NORM_WARPS + (threadIdx.x >> 5);
  if (row >= M) return;
  const int n_chunks = K >> 8;
  float v[AMP_MAX_CHUNKS][8];
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < AMP_MAX_CHUNKS; ++c) {
    if (c < n_chunks) {
      const size_t k = (size_t)row * K + (c * 32 + lane) * 8;
      ld8_f32(h_in + k, v[c]);
      if (a) {
        float fa[8];
        ld8_bf16(a + k, fa);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[c][e] += fa[e];
        st8_f32(h_out + k, v[c]);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) ss += v[c][e] * v[c][e];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float r = rsqrtf(ss / (float)K + eps);
  if (lane == 0) r_out[row] = r;
#pragma unroll
  for (int c = 0; c < AMP_MAX_CHUNKS; ++c) {
    if (c < n_chunks) {
      const int kk = (c * 32 + lane) * 8;
      float wv[8], out[8];
      ld8_f32(w + kk, wv);
#pragma unroll
      for (int e = 0; e < 8; ++e) out[e] = wv[e] * (v[c][e] * r);
      st8_bf16(x + (size_t)row * K + kk, out);
    }
  }
}

__global__ void __launch_bounds__(NORM_WARPS * 32)
rmsnorm_amp_bwd_kernel(const __nv_bfloat16* __restrict__ dx, const float* dh_out, const float* __restrict__ h,
                       const float* __restrict__ r_in, const float* __restrict__ w, float* dh,
                       float* __restrict__ dw_partial, int M, int K) {
  const int lane = threadIdx.x & 31;
  const int gwarp = blockIdx.x * NORM_WARPS + (threadIdx.x >> 5), n_warps = gridDim.x * NORM_WARPS;
  const int n_chunks = K >> 8;
  float dw[AMP_MAX_CHUNKS][8];
#pragma unroll
  for (int c = 0; c < AMP_MAX_CHUNKS; ++c)
#pragma unroll
    for (int e = 0; e < 8; ++e) dw[c][e] = 0.f;
  for (int row = gwarp; row < M; row += n_warps) {
    const float r = r_in[row];
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < AMP_MAX_CHUNKS; ++c) {
      if (c < n_chunks) {
        const int kk = (c * 32 + lane) * 8;
        float g[8], hv[8], wv[8];
        ld8_bf16(dx + (size_t)row * K + kk, g);
        ld8_f32(h + (size_t)row * K + kk, hv);
        ld8_f32(w + kk, wv);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          dw[c][e] += g[e] * hv[e] * r;
          dot += g[e] * wv[e] * hv[e];
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    const float coef = r * r * r * dot / (float)K;
#pragma unroll
    for (int c = 0; c < AMP_MAX_CHUNKS; ++c) {
      if (c < n_chunks) {
        const int kk = (c * 32 + lane) * 8;
        const size_t k = (size_t)row * K + kk;
        float g[8], hv[8], wv[8], out[8];
        ld8_bf16(dx + k, g);
        ld8_f32(h + k, hv);
        ld8_f32(w + kk, wv);
        if (dh_out) ld8_f32(dh_out + k, out);
#pragma unroll
        for (int e = 0; e < 8; ++e) out[e] = (dh_out ? out[e] : 0.f) + r * g[e] * wv[e] - hv[e] * coef;
        st8_f32(dh + k, out);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < AMP_MAX_CHUNKS; ++c)
    if (c < n_chunks) st8_f32(dw_partial + (size_t)gwarp * K + (c * 32 + lane) * 8, dw[c]);
}

__global__ void __launch_bounds__(256)
swiglu_bwd_kernel(const __nv_bfloat16* __restrict__ d_out, const __nv_bfloat16* __restrict__ gate_up,
                  __nv_bfloat16* __restrict__ d_gate_up, int M, int N) {
  const int per_row = N >> 3;
  const long long total = (long long)M * per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(i / per_row), k = (int)(i % per_row) * 8;
    float d[8], g[8], u[8], dg[8], du[8];
    ld8_bf16(d_out + (size_t)row * N + k, d);
    ld8_bf16(gate_up + (size_t)row * 2 * N + k, g);
    ld8_bf16(gate_up + (size_t)row * 2 * N + N + k, u);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float s = 1.f / (1.f + expf(-g[e]));
      const float act = __bfloat162float(__float2bfloat16(g[e] * s));          // silu(gate) as the forward rounded it
      du[e] = d[e] * act;                                                      // d_up = d_out * act (bf16 product)
      const float d_act = __bfloat162float(__float2bfloat16(d[e] * u[e]));     // d_act = d_out * up, rounded to bf16
      dg[e] = d_act * (s * (1.f + g[e] * (1.f - s)));                          // silu_backward in fp32
    }
    st8_bf16(d_gate_up + (size_t)row * 2 * N + k, dg);
    st8_bf16(d_gate_up + (size_t)row * 2 * N + N + k, du);
  }
}

}  // namespace aki

extern "C" int aki_mma_add_rmsnorm_amp_fwd(const float* h_in, const void* a, const float* weight, float eps, float* h_out,
                                           void* x, float* r_out, int M, int K, aki_stream_t stream) {
  AKI_REQUIRE(h_in && weight && x && r_out, AKI_ERR_NULL);
  AKI_REQUIRE((a != nullptr) == (h_out != nullptr), AKI_ERR_NULL);
  AKI_REQUIRE(M > 0 && K > 0, AKI_ERR_BAD_SHAPE);
  AKI_REQUIRE(K % 256 == 0 && K <= 256 * aki::AMP_MAX_CHUNKS, AKI_ERR_UNSUPPORTED);
  AKI_REQUIRE(aki::aligned16(h_in) && aki::aligned16(weight) && aki::aligned16(x) && (!a || aki::aligned16(a)) &&
                  (!h_out || aki::aligned16(h_out)),
              AKI_ERR_MISALIGNED);
  aki::add_rmsnorm_amp_fwd_kernel<<<(M + aki::NORM_WARPS - 1) / aki::NORM_WARPS, aki::NORM_WARPS * 32, 0,
                                    static_cast<cudaStream_t>(stream)>>>(
      h_in, static_cast<const __nv_bfloat16*>(a), weight, eps, h_out, static_cast<__nv_bfloat16*>(x), r_out, M, K);
  return aki::check_launch();
}

extern "C" int aki_mma_rmsnorm_amp_bwd_partials(int M) {
  if (M <= 0) return 0;
  const int ctas = (M + aki::NORM_WARPS - 1) / aki::NORM_WARPS;
  return (ctas < 296 ? ctas : 296) * aki::NORM_WARPS;
}

extern "C" int aki_mma_rmsnorm_amp_bwd(const void* dx, const float* dh_out, const float* h, const float* r,
                                       const float* weight, float* dh, float* dw_partial, int M, int K,
                                       aki_stream_t stream) {
  AKI_REQUIRE(dx && h && r && weight && dh && dw_partial, AKI_ERR_NULL);
  AKI_REQUIRE(M > 0 && K > 0, AKI_ERR_BAD_SHAPE);
  AKI_REQUIRE(K % 256 == 0 && K <= 256 * aki::AMP_MAX_CHUNKS, AKI_ERR_UNSUPPORTED);
  AKI_REQUIRE(aki::aligned16(dx) && aki::aligned16(h) && aki::aligned16(weight) && aki::aligned16(dh) &&
                  aki::aligned16(dw_partial) && (!dh_out || aki::aligned16(dh_out)),
              AKI_ERR_MISALIGNED);
  const int grid = aki_mma_rmsnorm_amp_bwd_partials(M) / aki::NORM_WARPS;
  aki::rmsnorm_amp_bwd_kernel<<<grid, aki::NORM_WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dx), dh_out, h, r, weight, dh, dw_partial, M, K);
  return aki::check_launch();
}

extern "C" int aki_mma_swiglu_bwd(const void* d_out, const void* gate_up, void* d_gate_up, int M, int N,
                                  aki_stream_t stream) {
  AKI_REQUIRE(d_out && gate_up && d_gate_up, AKI_ERR_NULL);
  AKI_REQUIRE(M > 0 && N > 0, AKI_ERR_BAD_SHAPE);
  AKI_REQUIRE(N % 8 == 0, AKI_ERR_UNSUPPORTED);
  AKI_REQUIRE(aki::aligned16(d_out) && aki::aligned16(gate_up) && aki::aligned16(d_gate_up), AKI_ERR_MISALIGNED);
  const long long total = (long long)M * (N / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  aki::swiglu_bwd_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(d_out), static_cast<const __nv_bfloat16*>(gate_up),
      static_cast<__nv_bfloat16*>(d_gate_up), M, N);
  return aki::check_launch();
}
