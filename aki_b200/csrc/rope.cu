// Phi-3 longrope tables and the fused RoPE + KV-cache write (HBM-bound elementwise kernels).
//   rope_table    : Phi3RotaryEmbedding.forward  (installed equivalent models/phi3/modeling_phi3.py:118-131)
//   rope_kv_write : apply_rotary_pos_emb on K (+ optionally Q) and DynamicCache.update
//                   (models/phi3/modeling_phi3.py:178-205, :248-251), written straight into (B,H,t_cap,D) caches.
#include <cuda_bf16.h>
#include "api_common.cuh"

namespace aki {

__global__ void __launch_bounds__(256)
rope_table_kernel(const int64_t* __restrict__ position_ids, const float* __restrict__ inv_freq, float attention_factor,
                  long long n, int half_dim, float* __restrict__ cos_out, float* __restrict__ sin_out) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= n) return;
  const long long bt = idx / half_dim;
  const int k = (int)(idx - bt * half_dim);
  const float f = (float)position_ids[bt] * inv_freq[k];   // fp32 product, as the reference's fp32 matmul
  float s, c;
  sincosf(f, &s, &c);
  cos_out[idx] = c * attention_factor;
  sin_out[idx] = s * attention_factor;
}

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}
// x' = x*cos + rotate_half(x)*sin with pairs (d, d+D/2):  lo' = lo*c - hi*s ; hi' = hi*c + lo*s
__device__ __forceinline__ void rotate8(const uint4& lo_in, const uint4& hi_in, const float* c, const float* s,
                                        uint4& lo_out, uint4& hi_out) {
  float lo[8], hi[8], lo2[8], hi2[8];
  unpack8(lo_in, lo);
  unpack8(hi_in, hi);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    lo2[i] = lo[i] * c[i] - hi[i] * s[i];
    hi2[i] = hi[i] * c[i] + lo[i] * s[i];
  }
  lo_out = pack8(lo2);
  hi_out = pack8(hi2);
}

// one thread per (b, t, h, c) with c in [0, 6): chunk c and c+6 of the 12 x 16-byte chunks of a 96-wide head row
__global__ void __launch_bounds__(192)
rope_kv_write_kernel(const __nv_bfloat16* __restrict__ qkv, int64_t qkv_stride_b, int64_t qkv_stride_t,
                     const float* __restrict__ cos_t, const float* __restrict__ sin_t, int64_t rope_stride_b, int T,
                     int H, __nv_bfloat16* __restrict__ k_cache, __nv_bfloat16* __restrict__ v_cache,
                     int64_t cache_stride_b, int64_t cache_stride_h, int past_len_host,
                     const int32_t* __restrict__ past_len_dev, int t_cap, __nv_bfloat16* __restrict__ q_rot) {
  constexpr int D = 96, HALF = 48;
  const int t = blockIdx.x, b = blockIdx.y;
  pdl_wait();          // qkv comes from the previous kernel of a decode step
  pdl_trigger();
  const int past_len = past_len_dev ? past_len_dev[b] : past_len_host;   // device-resident length: CUDA-graph decode
  const float* cr = cos_t + (size_t)b * rope_stride_b + (size_t)t * HALF;
  const float* sr = sin_t + (size_t)b * rope_stride_b + (size_t)t * HALF;
  const __nv_bfloat16* row = qkv + (size_t)b * qkv_stride_b + (size_t)t * qkv_stride_t;
  for (int idx = threadIdx.x; idx < H * 6; idx += 192) {
    const int h = idx / 6, c = idx - h * 6;
    float cs[8], sn[8];
    *reinterpret_cast<float4*>(cs) = *reinterpret_cast<const float4*>(cr + c * 8);
    *reinterpret_cast<float4*>(cs + 4) = *reinterpret_cast<const float4*>(cr + c * 8 + 4);
    *reinterpret_cast<float4*>(sn) = *reinterpret_cast<const float4*>(sr + c * 8);
    *reinterpret_cast<float4*>(sn + 4) = *reinterpret_cast<const float4*>(sr + c * 8 + 4);
    const size_t dst = (size_t)b * cache_stride_b + (size_t)h * cache_stride_h + (size_t)(past_len + t) * D;
    const bool in_cache = (past_len + t) < t_cap;     // a row past the capacity is dropped, never written next door
    if (in_cache) {
      const __nv_bfloat16* kp = row + (size_t)H * D + h * D + c * 8;
      uint4 lo = *reinterpret_cast<const uint4*>(kp), hi = *reinterpret_cast<const uint4*>(kp + HALF), lo2, hi2;
      rotate8(lo, hi, cs, sn, lo2, hi2);
      *reinterpret_cast<uint4*>(k_cache + dst + c * 8) = lo2;
      *reinterpret_cast<uint4*>(k_cache + dst + HALF + c * 8) = hi2;
    }
    if (v_cache && in_cache) {
      const __nv_bfloat16* vp = row + (size_t)2 * H * D + h * D + c * 8;
      *reinterpret_cast<uint4*>(v_cache + dst + c * 8) = *reinterpret_cast<const uint4*>(vp);
      *reinterpret_cast<uint4*>(v_cache + dst + HALF + c * 8) = *reinterpret_cast<const uint4*>(vp + HALF);
    }
    if (q_rot) {
      const __nv_bfloat16* qp = row + h * D + c * 8;
      uint4 lo = *reinterpret_cast<const uint4*>(qp), hi = *reinterpret_cast<const uint4*>(qp + HALF), lo2, hi2;
      rotate8(lo, hi, cs, sn, lo2, hi2);
      const size_t qd = (((size_t)b * H + h) * T + t) * D;
      *reinterpret_cast<uint4*>(q_rot + qd + c * 8) = lo2;
      *reinterpret_cast<uint4*>(q_rot + qd + HALF + c * 8) = hi2;
    }
  }
}

}  // namespace aki

using namespace aki;

extern "C" int aki_mma_rope_table(const int64_t* position_ids, const float* inv_freq, float attention_factor, int B,
                                  int T, int half_dim, float* cos_out, float* sin_out, aki_stream_t stream) {
  AKI_REQUIRE(position_ids && inv_freq && cos_out && sin_out, AKI_ERR_NULL);
  AKI_REQUIRE(B > 0 && T > 0 && half_dim > 0, AKI_ERR_BAD_SHAPE);
  const long long n = (long long)B * T * half_dim;
  rope_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      position_ids, inv_freq, attention_factor, n, half_dim, cos_out, sin_out);
  return check_launch();
}

static int rope_kv_write_impl(const void* qkv, int64_t qkv_stride_b, int64_t qkv_stride_t, const float* cos,
                              const float* sin, int64_t rope_stride_b, int B, int T, int H, int D, void* k_cache,
                              void* v_cache, int64_t cache_stride_b, int64_t cache_stride_h, int past_len,
                              const int32_t* past_len_dev, int t_cap, void* q_rot, aki_stream_t stream) {
  AKI_REQUIRE(qkv && cos && sin && k_cache, AKI_ERR_NULL);
  AKI_REQUIRE(B > 0 && T > 0 && H > 0 && past_len >= 0 && B <= 65535 && t_cap > 0, AKI_ERR_BAD_SHAPE);
  AKI_REQUIRE(past_len_dev || past_len + T <= t_cap, AKI_ERR_BAD_SHAPE);
  AKI_REQUIRE(D == AKI_MMA_HEAD_DIM, AKI_ERR_UNSUPPORTED);
  AKI_REQUIRE(qkv_stride_b % 8 == 0 && qkv_stride_t % 8 == 0 && cache_stride_b % 8 == 0 && cache_stride_h % 8 == 0,
              AKI_ERR_UNSUPPORTED);
  AKI_REQUIRE(aligned16(qkv) && aligned16(k_cache) && aligned16(cos) && aligned16(sin) &&
                  (!v_cache || aligned16(v_cache)) && (!q_rot || aligned16(q_rot)),
              AKI_ERR_MISALIGNED);
  launch_pdl(rope_kv_write_kernel, dim3(T, B), dim3(192), 0, static_cast<cudaStream_t>(stream),
             static_cast<const __nv_bfloat16*>(qkv), qkv_stride_b, qkv_stride_t, cos, sin, rope_stride_b, T, H,
             static_cast<__nv_bfloat16*>(k_cache), static_cast<__nv_bfloat16*>(v_cache), cache_stride_b, cache_stride_h,
             past_len, past_len_dev, t_cap, static_cast<__nv_bfloat16*>(q_rot));
  return check_launch();
}

extern "C" int aki_mma_rope_kv_write(const void* qkv, int64_t qkv_stride_b, int64_t qkv_stride_t, const float* cos,
                                     const float* sin, int64_t rope_stride_b, int B, int T, int H, int D,
                                     void* k_cache, void* v_cache, int64_t cache_stride_b, int64_t cache_stride_h,
                                     int past_len, int t_cap, void* q_rot, aki_stream_t stream) {
  return rope_kv_write_impl(qkv, qkv_stride_b, qkv_stride_t, cos, sin, rope_stride_b, B, T, H, D, k_cache, v_cache,
                            cache_stride_b, cache_stride_h, past_len, nullptr, t_cap, q_rot, stream);
}

extern "C" int aki_mma_rope_kv_write_dev(const void* qkv, int64_t qkv_stride_b, int64_t qkv_stride_t, const float* cos,
                                         const float* sin, int64_t rope_stride_b, int B, int T, int H, int D,
                                         void* k_cache, void* v_cache, int64_t cache_stride_b, int64_t cache_stride_h,
                                         const int32_t* past_len_dev, int t_cap, void* q_rot, aki_stream_t stream) {
  AKI_REQUIRE(past_len_dev, AKI_ERR_NULL);
  return rope_kv_write_impl(qkv, qkv_stride_b, qkv_stride_t, cos, sin, rope_stride_b, B, T, H, D, k_cache, v_cache,
                            cache_stride_b, cache_stride_h, 0, past_len_dev, t_cap, q_rot, stream);
}
