// The MMA visibility predicate shared by every attention kernel:
//   allowed(i,j) = i<len & j<len & ( (j<=i & valid[j]) | (row_lo[i]<=j<row_hi[i] & mutual_ok[j]) )
// == the reference's (B,1,T,T) 0/1 mask (codes/open_flamingo/src/vlm.py:410-443 + utils.py:99-108).
#pragma once
#include <stdint.h>
#include "../../include/aki_mma.h"

namespace aki {

struct MaskMeta {
  const int32_t* seq_len;
  const int32_t* row_lo;
  const int32_t* row_hi;
  const uint32_t* vbits;
  const uint32_t* mbits;
  const int32_t* q_tile_kv_end;
  const uint32_t* kv_tile_q_mask;
  int meta_pitch, bits_pitch;
};

inline MaskMeta mask_meta_from(const AkiMmaAttnParams& p) {
  MaskMeta m;
  m.seq_len = p.seq_len; m.row_lo = p.row_lo; m.row_hi = p.row_hi;
  m.vbits = p.kv_valid_bits; m.mbits = p.kv_mutual_bits;
  m.q_tile_kv_end = p.q_tile_kv_end; m.kv_tile_q_mask = p.kv_tile_q_mask;
  m.meta_pitch = p.meta_pitch; m.bits_pitch = p.bits_pitch;
  return m;
}

__device__ __forceinline__ int meta_len(const MaskMeta& m, int b, int T) {
  return m.seq_len ? min(m.seq_len[b], T) : T;
}
__device__ __forceinline__ bool mma_allowed(const MaskMeta& m, int b, int i, int j, int len) {
  if (i >= len || j >= len) return false;
  bool vb = true, mb = true;
  if (m.vbits) vb = (m.vbits[(size_t)b * m.bits_pitch + (j >> 5)] >> (j & 31)) & 1u;
  if (m.mbits) mb = (m.mbits[(size_t)b * m.bits_pitch + (j >> 5)] >> (j & 31)) & 1u;
  int lo = 0, hi = 0;
  if (m.row_lo) { lo = m.row_lo[(size_t)b * m.meta_pitch + i]; hi = m.row_hi[(size_t)b * m.meta_pitch + i]; }
  return (j <= i && vb) || (j >= lo && j < hi && mb);
}
// right-most key (exclusive) row i can see
__device__ __forceinline__ int mma_row_end(const MaskMeta& m, int b, int i, int len) {
  int e = i + 1;
  if (m.row_lo) {
    const int lo = m.row_lo[(size_t)b * m.meta_pitch + i], hi = m.row_hi[(size_t)b * m.meta_pitch + i];
    if (hi > lo) e = max(e, hi);
  }
  return min(e, len);
}

}  // namespace aki
