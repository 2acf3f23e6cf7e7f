// Helpers shared by the backward kernels: workspace carving, the preprocess kernel (delta = rowsum(O*dO),
// rotated copy of Q) and the dQ finalize kernel (fp32 accumulator -> inverse RoPE -> bf16).
#pragma once
#include <cuda_bf16.h>
#include "api_common.cuh"
#include "mask_pred.cuh"

namespace aki {

struct BwdWorkspace {
  __nv_bfloat16* q_rot;   // (B,H,T,D) bf16, post-RoPE queries
  float* delta;           // (B,H,T)
  float* dq_accum;        // (B,H,T,D) fp32
  __nv_bfloat16* row_stats;  // (B,H,t_pad,8) bf16: [-LSE/scale hi,mid,lo, 0, -delta hi,mid,lo, 0]: the B operand of the
                             // statistics k-step of S^T (selected by ones [1,1,1,0,0,0,0,0]) and of dP^T ([0,0,0,0,1,1,1,0])
  int t_pad;              // T rounded up to the 128-row tile: the statistics tiles are never out of bounds
  size_t bytes;
};

inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

inline BwdWorkspace carve_bwd_workspace(void* base, int B, int H, int T, int D) {
  BwdWorkspace w;
  const size_t n = (size_t)B * H * T;
  size_t off = 0;
  char* p = static_cast<char*>(base);
  w.q_rot = reinterpret_cast<__nv_bfloat16*>(p + off); off += align256(n * D * 2);
  w.delta = reinterpret_cast<float*>(p + off);         off += align256(n * 4);
  w.dq_accum = reinterpret_cast<float*>(p + off);      off += align256(n * D * 4);
  w.t_pad = (T + 127) / 128 * 128;
  const size_t na = (size_t)B * H * w.t_pad;
  w.row_stats = reinterpret_cast<__nv_bfloat16*>(p + off);  off += align256(na * 16);
  w.bytes = off;
  return w;
}

struct TensorView {  // (B,T,H,D) bf16 view
  __nv_bfloat16* ptr;
  int64_t sb, st, sh;
  __device__ __forceinline__ __nv_bfloat16* row(int b, int t, int h) const {
    return ptr + (size_t)b * sb + (size_t)t * st + (size_t)h * sh;
  }
};
inline TensorView view_of(const AkiMmaTensor4& t) {
  return TensorView{static_cast<__nv_bfloat16*>(t.ptr), t.stride_b, t.stride_t, t.stride_h};
}

inline int check_tensor(const AkiMmaTensor4& t) {
  if (!t.ptr) return AKI_ERR_NULL;
  if (!aligned16(t.ptr)) return AKI_ERR_MISALIGNED;
  if (t.stride_b % 8 || t.stride_t % 8 || t.stride_h % 8) return AKI_ERR_UNSUPPORTED;
  return AKI_OK;
}
int check_attn_params(const AkiMmaAttnParams& p);

int make_tile_map(CUtensorMap* m, const AkiMmaTensor4& t, int B, int H, int T, int box_rows);
int make_dq_accum_map(CUtensorMap* m, float* dq_accum, int B, int H, int T);
int make_row_stats_map(CUtensorMap* m, void* base, int B, int H, int t_pad);
int launch_bwd_preprocess(const AkiMmaAttnBwdParams& p, const BwdWorkspace& w, cudaStream_t st);
int launch_dq_finalize(const AkiMmaAttnBwdParams& p, const BwdWorkspace& w, float scale, cudaStream_t st);

}  // namespace aki
