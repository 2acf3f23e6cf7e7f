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
