// Modality-mutual attention forward for sm_100a: QK^T, online softmax and PV on tcgen05 tensor cores with
// TMEM accumulators, operands staged by TMA, mbarrier pipelines, warp-specialised roles.
//
// Replaces the eager core of Phi3Attention.forward (softmax_fp32(QK^T/sqrt(96) + mask) V; installed
// equivalent transformers/models/phi3/modeling_phi3.py:153-175) fed by the reference's materialised
// (B,1,T,T) mask (codes/open_flamingo/src/vlm.py:410-443).  No mask is read from HBM: the predicate
//   allowed(i,j) = (j<=i & valid[j]) | (row_lo[i]<=j<row_hi[i] & mutual_ok[j])
// is evaluated in registers on the few tiles that are not fully visible, and tiles beyond
// q_tile_kv_end[b][qt] are never visited.  RoPE is applied to Q in shared memory right after the TMA load.
//
// CTA = 2 query tiles x 128 rows of one (batch, head); 20 warps:
//   warp 0        TMA producer (Q once, then K_j / V_j through two 3-deep rings)
//   warp 1        MMA issuer (one elected lane): S_t = Q_t K_j^T (SS), O_t += P_t V_j (TS, P read from TMEM)
//   warp 2        TMEM allocator;   warp 3 idle
//   warps 4-11    softmax of tile 0: thread <-> row r <-> TMEM lane r; warps 4-7 own key columns [0,64) of the
//                 128-key tile, warps 8-11 columns [64,128) (row max / row sum exchanged through shared memory)
//   warps 12-19   softmax of tile 1, same split
// Four softmax warps per SM sub-partition (instead of two) keep the MUFU (exp2) pipe -- the real bound of this
// head_dim (128 exp vs 768 tensor cycles per row tile) -- busy while other warps wait on TMEM / barriers.
// TMEM columns: S0 [0,128) S1 [128,256) O0 [256,352) O1 [352,448); P_t (bf16 pairs) aliases S_t[0,64).
// Shared memory: Q 2x24 KB, K ring 3x24 KB, V ring 3x24 KB; every tile is 3 SWIZZLE_64B atoms [128][64 B]
// (head_dim 96 = 3 x 32), the layout both the TMA boxes and the UMMA descriptors use (validated by
// tools/umma_probe.cu).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "attn_aux.cuh"
#include "sm100_ptx.cuh"

namespace aki {

namespace fwd {
constexpr int BM = 128, BN = 128, HD = 96;
constexpr int ATOM_BYTES = 128 * 64;
constexpr int TILE_BYTES = 3 * ATOM_BYTES;  // 24576
constexpr int STAGES = 3;
constexpr int THREADS = 640;
constexpr int SMEM_Q = 0;
constexpr int SMEM_K = SMEM_Q + 2 * TILE_BYTES;
constexpr int SMEM_V = SMEM_K + STAGES * TILE_BYTES;
constexpr int SMEM_X = SMEM_V + STAGES * TILE_BYTES;      // exchange: [tile 2][parity 2][half 2][128] floats
constexpr int SMEM_TOTAL = SMEM_X + 2 * 2 * 2 * 128 * 4;  // 196608 + 4096
constexpr int SMEM_ALLOC = SMEM_TOTAL + 1024;             // slack for 1024-byte alignment
constexpr uint32_t TM_S0 = 0, TM_S1 = 128, TM_O0 = 256, TM_O1 = 352;
constexpr int REGS_CTRL = 40, REGS_SOFTMAX = 104;  // setmaxnreg draws from the CTA pool: 128*40 + 512*104 = 58368 <= 640*96 = 61440
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 units: P may grow to 2^8 before O is rescaled
}  // namespace fwd

struct FwdKernelParams {
  TensorView q, o;
  float* lse;
  const float* rope_cos;
  const float* rope_sin;
  int64_t rope_stride_b;
  MaskMeta mm;
  int B, H, T, n_qt, n_qp;
  float scale_log2, scale;
  unsigned long long* trace;  // debug: per-iteration clock64 stamps of one CTA (AKI_MMA_FWD_TRACE=<cta>)
  int trace_cta;
};

__device__ __forceinline__ uint32_t low_mask(int n) {  // n low bits set, n clamped to [0,32]
  return n <= 0 ? 0u : (n >= 32 ? 0xffffffffu : ((1u << n) - 1u));
}

#define TR(slot, j, k) do { if (tracing && (j) < 128) P.trace[((slot) * 128 + (j)) * 8 + (k)] = clock64(); } while (0)

template <bool ROPE>
__global__ void __launch_bounds__(fwd::THREADS, 1)
attn_fwd_sm100_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                      const __grid_constant__ CUtensorMap map_v, const FwdKernelParams P) {
  using namespace fwd;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  __shared__ __align__(8) uint64_t bars[2 + 2 + 4 * STAGES + 6];
  __shared__ uint32_t tmem_base_s;
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  // barrier indices
  constexpr int Q_FULL = 0, Q_READY = 2, K_FULL = 4, K_EMPTY = K_FULL + STAGES, V_FULL = K_EMPTY + STAGES,
                V_EMPTY = V_FULL + STAGES, S_FULL = V_EMPTY + STAGES, P_FULL = S_FULL + 2, O_FULL = P_FULL + 2;

  const int tid = threadIdx.x, warp = tid >> 5;
  // work decomposition: consecutive CTAs share (b,h) so K/V stay in L2; heaviest query tiles first
  const int bh = blockIdx.x / P.n_qp;
  const int qp = P.n_qp - 1 - (blockIdx.x % P.n_qp);
  const int b = bh / P.H, h = bh % P.H;
  const int n_kt = (P.T + BN - 1) / BN;
  int n_kv0, n_kv1;
  {
    int n[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int qt = 2 * qp + t;
      if (qt >= P.n_qt) n[t] = 0;
      else if (P.mm.q_tile_kv_end) n[t] = min(P.mm.q_tile_kv_end[(size_t)b * P.n_qt + qt], n_kt);
      else n[t] = min(qt + 1, n_kt);
    }
    n_kv0 = n[0]; n_kv1 = n[1];
  }
  const int n_max = max(n_kv0, n_kv1);

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(Q_FULL + i), 1); mbar_init(BAR(Q_READY + i), 256); }
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(BAR(K_FULL + i), 1); mbar_init(BAR(K_EMPTY + i), 1);
      mbar_init(BAR(V_FULL + i), 1); mbar_init(BAR(V_EMPTY + i), 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(S_FULL + i), 1); mbar_init(BAR(P_FULL + i), 256); mbar_init(BAR(O_FULL + i), 1); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(smem_u32(&tmem_base_s));
  if (warp == 0 && elect_one()) { tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    setmaxnreg_dec<REGS_CTRL>();
    if (elect_one()) {
      for (int t = 0; t < 2; ++t) {
        if ((t ? n_kv1 : n_kv0) == 0) continue;
        mbar_arrive_expect_tx(BAR(Q_FULL + t), TILE_BYTES);
        for (int a = 0; a < 3; ++a)
          tma_load_4d(smem_base + SMEM_Q + t * TILE_BYTES + a * ATOM_BYTES, &map_q, BAR(Q_FULL + t), a * 32,
                      (2 * qp + t) * BM, h, b);
      }
      for (int j = 0; j < n_max; ++j) {
        const int s = j % STAGES;
        const uint32_t ph = (j / STAGES) & 1;
        mbar_wait(BAR(K_EMPTY + s), ph ^ 1);
        mbar_arrive_expect_tx(BAR(K_FULL + s), TILE_BYTES);
        for (int a = 0; a < 3; ++a)
          tma_load_4d(smem_base + SMEM_K + s * TILE_BYTES + a * ATOM_BYTES, &map_k, BAR(K_FULL + s), a * 32, j * BN, h, b);
        mbar_wait(BAR(V_EMPTY + s), ph ^ 1);
        mbar_arrive_expect_tx(BAR(V_FULL + s), TILE_BYTES);
        for (int a = 0; a < 3; ++a)
          tma_load_4d(smem_base + SMEM_V + s * TILE_BYTES + a * ATOM_BYTES, &map_v, BAR(V_FULL + s), a * 32, j * BN, h, b);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    setmaxnreg_dec<REGS_CTRL>();
    if (elect_one()) {
      const bool tracing = P.trace && (int)blockIdx.x == P.trace_cta;
      constexpr uint32_t IDESC_QK = umma_idesc_bf16(BM, BN, 0, 0);
      constexpr uint32_t IDESC_PV = umma_idesc_bf16(BM, HD, 0, 1);
      // descriptors differ only in the 14-bit start-address field (units of 16 B)
      const uint64_t DESC_KMAJ = umma_smem_desc(0, 16, 512, UMMA_SW64);
      const uint64_t DESC_MNMAJ = umma_smem_desc(0, ATOM_BYTES, 512, UMMA_SW64);
      const uint32_t q_lo = (smem_base + SMEM_Q) >> 4, k_lo = (smem_base + SMEM_K) >> 4, v_lo = (smem_base + SMEM_V) >> 4;
      auto issue_qk = [&](int t, int j) {
        const uint32_t qa = q_lo + t * (TILE_BYTES >> 4), ka = k_lo + (j % STAGES) * (TILE_BYTES >> 4);
        const uint32_t d = tmem + (t ? TM_S1 : TM_S0);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          const uint32_t off = ((k >> 1) * ATOM_BYTES + (k & 1) * 32) >> 4;
          umma_ss(d, DESC_KMAJ | (uint64_t)(qa + off), DESC_KMAJ | (uint64_t)(ka + off), IDESC_QK, k > 0);
        }
      };
      auto issue_pv = [&](int t, int j) {
        const uint32_t va = v_lo + (j % STAGES) * (TILE_BYTES >> 4);
        const uint32_t d = tmem + (t ? TM_O1 : TM_O0), a = tmem + (t ? TM_S1 : TM_S0);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ts(d, a + 8 * k, DESC_MNMAJ | (uint64_t)(va + k * 64), IDESC_PV, (j > 0 || k > 0));
      };
      if (n_kv0 > 0) mbar_wait(BAR((ROPE ? Q_READY : Q_FULL) + 0), 0);
      if (n_kv1 > 0) mbar_wait(BAR((ROPE ? Q_READY : Q_FULL) + 1), 0);
      if (n_max > 0) {
        mbar_wait(BAR(K_FULL + 0), 0);
        tc_fence_after();
        if (n_kv0 > 0) { issue_qk(0, 0); umma_commit(BAR(S_FULL + 0)); }
        if (n_kv1 > 0) { issue_qk(1, 0); umma_commit(BAR(S_FULL + 1)); }
        umma_commit(BAR(K_EMPTY + 0));
      }
      for (int j = 0; j < n_max; ++j) {
        const int sv = j % STAGES, jn = j + 1, sk = jn % STAGES;
        // operand-ready waits first: they are long complete in steady state and must not sit between the
        // softmax's P_FULL arrival and the MMA issue (the critical chain of each tile)
        mbar_wait(BAR(V_FULL + sv), (j / STAGES) & 1);
        if (jn < n_max) mbar_wait(BAR(K_FULL + sk), (jn / STAGES) & 1);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int nk = t ? n_kv1 : n_kv0, nk_other = t ? n_kv0 : n_kv1;
          if (j >= nk) continue;
          TR(2 + t, j, 0);
          mbar_wait(BAR(P_FULL + t), j & 1);
          tc_fence_after();
          TR(2 + t, j, 1);
          issue_pv(t, j);
          umma_commit(BAR(O_FULL + t));
          // V_j is released by its last user: tile 1 if it uses j, else tile 0
          if (t == 1 || j >= nk_other) umma_commit(BAR(V_EMPTY + sv));
          TR(2 + t, j, 2);
          if (jn < nk) {
            issue_qk(t, jn);
            umma_commit(BAR(S_FULL + t));
            if (t == 1 || jn >= nk_other) umma_commit(BAR(K_EMPTY + sk));
          }
          TR(2 + t, j, 3);
        }
      }
    }
  } else if (warp < 4) {
    setmaxnreg_dec<REGS_CTRL>();
  } else {
    // ------------------------------------------------------------------ softmax / correction / epilogue
    setmaxnreg_inc<REGS_SOFTMAX>();
    const int st_ = tid - 128;                    // 0..511
    const int t = st_ >> 8;                       // query tile of the pair
    const int hf = (st_ >> 7) & 1;                // half of the key columns (and of the output columns)
    const int r = st_ & 127;                      // row within the tile == TMEM lane
    const int qt = 2 * qp + t;
    const int i = qt * BM + r;                    // query index in mask coordinates
    const int len = meta_len(P.mm, b, P.T);
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tm_s = tmem + (t ? TM_S1 : TM_S0) + lane_base;
    const uint32_t tm_o = tmem + (t ? TM_O1 : TM_O0) + lane_base;
    const int nk = t ? n_kv1 : n_kv0;
    const int bar_id = 1 + t;                     // named barrier of the 256 threads of this tile
    float* const xch = reinterpret_cast<float*>(smem_gen + SMEM_X) + t * 512;   // [parity][half][128]
    const bool tracing = P.trace && (int)blockIdx.x == P.trace_cta && r == 0 && hf == 0;

    if (ROPE && nk > 0) {
      mbar_wait(BAR(Q_FULL + t), 0);
      if (i < P.T) {
        const uint32_t qa = smem_base + SMEM_Q + t * TILE_BYTES;
        const float* cr = P.rope_cos + (size_t)b * P.rope_stride_b + (size_t)i * 48;
        const float* sr = P.rope_sin + (size_t)b * P.rope_stride_b + (size_t)i * 48;
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          const int c = 3 * hf + cc;              // 16-byte chunk c pairs with chunk c+6 (d <-> d+48)
          const uint32_t a_lo = qa + (c >> 2) * ATOM_BYTES + sw64_offset(r, c & 3);
          const uint32_t a_hi = qa + ((c + 6) >> 2) * ATOM_BYTES + sw64_offset(r, (c + 6) & 3);
          uint4 lo, hi;
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w) : "r"(a_lo));
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w) : "r"(a_hi));
          float cs[8], sn[8];
          *reinterpret_cast<float4*>(cs) = __ldg(reinterpret_cast<const float4*>(cr + c * 8));
          *reinterpret_cast<float4*>(cs + 4) = __ldg(reinterpret_cast<const float4*>(cr + c * 8 + 4));
          *reinterpret_cast<float4*>(sn) = __ldg(reinterpret_cast<const float4*>(sr + c * 8));
          *reinterpret_cast<float4*>(sn + 4) = __ldg(reinterpret_cast<const float4*>(sr + c * 8 + 4));
          const __nv_bfloat162* l2 = reinterpret_cast<const __nv_bfloat162*>(&lo);
          const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&hi);
          uint32_t lo_w[4], hi_w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 lf = __bfloat1622float2(l2[e]), hf2 = __bfloat1622float2(h2[e]);
            lo_w[e] = pack_bf16x2(lf.x * cs[2 * e] - hf2.x * sn[2 * e], lf.y * cs[2 * e + 1] - hf2.y * sn[2 * e + 1]);
            hi_w[e] = pack_bf16x2(hf2.x * cs[2 * e] + lf.x * sn[2 * e], hf2.y * cs[2 * e + 1] + lf.y * sn[2 * e + 1]);
          }
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a_lo), "r"(lo_w[0]), "r"(lo_w[1]), "r"(lo_w[2]), "r"(lo_w[3]) : "memory");
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a_hi), "r"(hi_w[0]), "r"(hi_w[1]), "r"(hi_w[2]), "r"(hi_w[3]) : "memory");
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(BAR(Q_READY + t));
    }

    int row_lo = 0, row_hi = 0;
    const bool row_live = (i < len);
    if (row_live && P.mm.row_lo) {
      row_lo = P.mm.row_lo[(size_t)b * P.mm.meta_pitch + i];
      row_hi = P.mm.row_hi[(size_t)b * P.mm.meta_pitch + i];
    }
    float m_used = -INFINITY;  // running max (raw score units) the accumulators are expressed against
    float l = 0.f;             // partial row sum over this thread's key columns

    for (int j = 0; j < nk; ++j) {
      TR(t, j, 0);
      mbar_wait(BAR(S_FULL + t), j & 1);
      tc_fence_after();
      TR(t, j, 1);
      float s[64];
      tmem_ld_x32(tm_s + 64 * hf, reinterpret_cast<uint32_t*>(s));
      tmem_ld_x32(tm_s + 64 * hf + 32, reinterpret_cast<uint32_t*>(s) + 32);

      // ---- tile classification (warp-uniform): fully visible tiles skip the predicate
      const int j0 = j * BN + 64 * hf;              // first key column of this thread
      uint32_t vw[2], mw[2];
      bool full = (j < qt) && (j * BN + BN <= len);
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        const int jw = j0 + 32 * w;
        const uint32_t in_len = low_mask(len - jw);
        vw[w] = P.mm.vbits ? (__ldg(P.mm.vbits + (size_t)b * P.mm.bits_pitch + (jw >> 5)) & in_len) : in_len;
        mw[w] = P.mm.mbits ? (__ldg(P.mm.mbits + (size_t)b * P.mm.bits_pitch + (jw >> 5)) & in_len) : in_len;
        full = full && (vw[w] == 0xffffffffu);
      }
      tmem_wait_ld();
      TR(t, j, 2);
      if (!full) {
        const int d = row_live ? (i - j0) : -1;             // causal: column c visible iff c <= d
        const int a = row_lo - j0, e = row_hi - j0;         // mutual: a <= c < e
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          const uint32_t causal = low_mask(d + 1 - 32 * w) & vw[w];
          const uint32_t mutual = row_live ? (low_mask(e - 32 * w) & ~low_mask(a - 32 * w) & mw[w]) : 0u;
          const uint32_t ok = causal | mutual;
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (!((ok >> c) & 1u)) s[32 * w + c] = -INFINITY;
        }
      }
      float mx0 = s[0], mx1 = s[1], mx2 = s[2], mx3 = s[3];
#pragma unroll
      for (int c = 4; c < 64; c += 4) {
        mx0 = fmaxf(mx0, s[c]); mx1 = fmaxf(mx1, s[c + 1]); mx2 = fmaxf(mx2, s[c + 2]); mx3 = fmaxf(mx3, s[c + 3]);
      }
      const float mx_half = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      // ---- exchange the half-row maxima with the partner warpgroup
      float* xp = xch + (j & 1) * 256;
      xp[hf * 128 + r] = mx_half;
      named_bar_sync(bar_id, 256);
      const float m_new = fmaxf(m_used, fmaxf(mx_half, xp[(hf ^ 1) * 128 + r]));
      TR(t, j, 3);
      // ---- lazy rescale of O (correction merged into the softmax warps; rare after the first tiles).
      // Both halves take the same decision (same m_used / m_new); each rescales its 48 output columns.
      if (j == 0) {
        m_used = m_new;
      } else {
        const bool need = (m_new - m_used) * P.scale_log2 > RESCALE_THRESHOLD || (m_used == -INFINITY && m_new > -INFINITY);
        if (__any_sync(0xffffffffu, need)) {
          const float alpha = (m_used == -INFINITY) ? 0.f : ex2_approx((m_used - m_new) * P.scale_log2);
          m_used = m_new;
          l *= alpha;
          mbar_wait(BAR(O_FULL + t), (j - 1) & 1);
          tc_fence_after();
          uint32_t o[48];
          tmem_ld_x32(tm_o + 48 * hf, o);
          tmem_ld_x16(tm_o + 48 * hf + 32, o + 32);
          tmem_wait_ld();
#pragma unroll
          for (int x = 0; x < 48; ++x) o[x] = __float_as_uint(__uint_as_float(o[x]) * alpha);
          tmem_st_x32(tm_o + 48 * hf, o);
          tmem_st_x16(tm_o + 48 * hf + 32, o + 32);
          tmem_wait_st();
        }
      }
      // ---- P = exp2((S - m) * scale*log2e), row sum, bf16 pack, store to TMEM (aliases S)
      const float neg_m = (m_used == -INFINITY) ? 0.f : -m_used * P.scale_log2;
      float sum0 = 0.f, sum1 = 0.f;
      uint32_t pk[32];
#pragma unroll
      for (int x = 0; x < 32; ++x) {
        const float p0 = ex2_approx(fmaf(s[2 * x], P.scale_log2, neg_m));
        const float p1 = ex2_approx(fmaf(s[2 * x + 1], P.scale_log2, neg_m));
        sum0 += p0; sum1 += p1;
        pk[x] = pack_bf16x2(p0, p1);
      }
      // every thread of the tile has read its S (barrier above) -> the aliased columns may be overwritten
      tmem_st_x32(tm_s + 32 * hf, pk);
      l += sum0 + sum1;
      TR(t, j, 4);
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(BAR(P_FULL + t));
      TR(t, j, 5);
    }

    // ---- epilogue: O / l -> bf16 -> global; LSE.  Each half stores 48 of the 96 output columns.
    if (qt < P.n_qt) {
      float inv_l = 0.f, l_tot = 0.f;
      if (nk > 0) {
        float* xp = xch + (nk & 1) * 256;          // parity not used by the last main-loop iteration
        xp[hf * 128 + r] = l;
        named_bar_sync(bar_id, 256);
        l_tot = l + xp[(hf ^ 1) * 128 + r];
        mbar_wait(BAR(O_FULL + t), (nk - 1) & 1);
        tc_fence_after();
        inv_l = (row_live && l_tot > 0.f) ? 1.f / l_tot : 0.f;  // batch-padding rows: zeros
      }
      // tcgen05.ld is warp-collective (.sync.aligned): load unconditionally, predicate only the global stores
      const bool store_row = (i < P.T);
      __nv_bfloat16* orow = P.o.row(b, store_row ? i : 0, h) + 48 * hf;
      uint32_t o[48];
      if (nk > 0) {
        tmem_ld_x32(tm_o + 48 * hf, o);
        tmem_ld_x16(tm_o + 48 * hf + 32, o + 32);
        tmem_wait_ld();
      }
#pragma unroll
      for (int x = 0; x < 6; ++x) {
        uint4 u;
        uint32_t* w = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e)
          w[e] = (inv_l > 0.f) ? pack_bf16x2(__uint_as_float(o[8 * x + 2 * e]) * inv_l,
                                               __uint_as_float(o[8 * x + 2 * e + 1]) * inv_l)
                               : 0u;
        if (store_row) *reinterpret_cast<uint4*>(orow + 8 * x) = u;
      }
      if (P.lse && store_row && hf == 0)
        P.lse[((size_t)b * P.H + h) * P.T + i] = (inv_l > 0.f) ? (m_used * P.scale + __logf(l_tot)) : INFINITY;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// (D=96, T, H, B) bf16 view -> tensor map with [32 x 128 x 1 x 1] SWIZZLE_64B boxes
int make_tile_map(CUtensorMap* m, const AkiMmaTensor4& t, int B, int H, int T, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_last_cuda_error("cuTensorMapEncodeTiled entry point not found"); return AKI_ERR_CUDA; }
  cuuint64_t dims[4] = {96, (cuuint64_t)T, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)t.stride_t * 2, (cuuint64_t)t.stride_h * 2, (cuuint64_t)t.stride_b * 2};
  cuuint32_t box[4] = {32, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, t.ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_last_cuda_error("cuTensorMapEncodeTiled failed"); return AKI_ERR_CUDA; }
  return AKI_OK;
}

// dQ accumulator (B,H,T,96) fp32 contiguous -> [32 x 128 x 1 x 1] fp32 SWIZZLE_128B boxes for TMA reductions
int make_dq_accum_map(CUtensorMap* m, float* dq_accum, int B, int H, int T) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_last_cuda_error("cuTensorMapEncodeTiled entry point not found"); return AKI_ERR_CUDA; }
  cuuint64_t dims[4] = {96, (cuuint64_t)T, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {96 * 4, (cuuint64_t)T * 96 * 4, (cuuint64_t)H * T * 96 * 4};
  cuuint32_t box[4] = {32, 128, 1, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dq_accum, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_last_cuda_error("cuTensorMapEncodeTiled(dq) failed"); return AKI_ERR_CUDA; }
  return AKI_OK;
}

}  // namespace aki

using namespace aki;

extern "C" int aki_mma_attn_fwd(const AkiMmaAttnParams* p, aki_stream_t stream) {
  AKI_REQUIRE(p, AKI_ERR_NULL);
  int rc = check_attn_params(*p);
  if (rc) return rc;
  CUtensorMap mq, mk, mv;
  if ((rc = make_tile_map(&mq, p->q, p->B, p->H, p->T, fwd::BM))) return rc;
  if ((rc = make_tile_map(&mk, p->k, p->B, p->H, p->T, fwd::BN))) return rc;
  if ((rc = make_tile_map(&mv, p->v, p->B, p->H, p->T, fwd::BN))) return rc;
  FwdKernelParams kp;
  kp.q = view_of(p->q); kp.o = view_of(p->o);
  kp.lse = p->lse; kp.rope_cos = p->rope_cos; kp.rope_sin = p->rope_sin; kp.rope_stride_b = p->rope_stride_b;
  kp.mm = mask_meta_from(*p);
  kp.B = p->B; kp.H = p->H; kp.T = p->T;
  kp.n_qt = (p->T + fwd::BM - 1) / fwd::BM;
  kp.n_qp = (kp.n_qt + 1) / 2;
  kp.scale = p->scale;
  kp.scale_log2 = p->scale * 1.4426950408889634f;
  kp.trace = nullptr; kp.trace_cta = -1;
  // Debug only (tools/fwd_trace.py): AKI_MMA_FWD_TRACE=<cta> dumps clock64 stamps of one CTA and SYNCHRONISES.
  const char* trace_env = getenv("AKI_MMA_FWD_TRACE");
  const size_t trace_bytes = 4 * 128 * 8 * sizeof(unsigned long long);
  if (trace_env) {
    kp.trace_cta = atoi(trace_env);
    cudaMalloc(&kp.trace, trace_bytes);
    cudaMemset(kp.trace, 0, trace_bytes);
  }
  const long long grid = (long long)kp.n_qp * p->H * p->B;
  AKI_REQUIRE(grid > 0 && grid < (1ll << 31), AKI_ERR_BAD_SHAPE);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(attn_fwd_sm100_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd::SMEM_ALLOC) != cudaSuccess ||
        cudaFuncSetAttribute(attn_fwd_sm100_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd::SMEM_ALLOC) != cudaSuccess) {
      set_last_cuda_error(cudaGetErrorString(cudaGetLastError()));
      return AKI_ERR_CUDA;
    }
    attr_done = true;
  }
  if (p->rope_cos)
    attn_fwd_sm100_kernel<true><<<(unsigned)grid, fwd::THREADS, fwd::SMEM_ALLOC, st>>>(mq, mk, mv, kp);
  else
    attn_fwd_sm100_kernel<false><<<(unsigned)grid, fwd::THREADS, fwd::SMEM_ALLOC, st>>>(mq, mk, mv, kp);
  if (trace_env) {
    cudaDeviceSynchronize();
    static unsigned long long host[4 * 128 * 8];
    cudaMemcpy(host, kp.trace, trace_bytes, cudaMemcpyDeviceToHost);
    cudaFree(kp.trace);
    unsigned long long t0 = ~0ull;
    for (size_t i = 0; i < 4 * 128 * 8; ++i) if (host[i] && host[i] < t0) t0 = host[i];
    const char* names[4] = {"softmax0", "softmax1", "mma_t0", "mma_t1"};
    for (int slot = 0; slot < 4; ++slot)
      for (int j = 0; j < 128; ++j) {
        if (!host[(slot * 128 + j) * 8]) continue;
        fprintf(stderr, "TRACE %s j=%d:", names[slot], j);
        for (int k = 0; k < 6; ++k) fprintf(stderr, " %llu", host[(slot * 128 + j) * 8 + k] ? host[(slot * 128 + j) * 8 + k] - t0 : 0ull);
        fprintf(stderr, "\n");
      }
  }
  return check_launch();
}
