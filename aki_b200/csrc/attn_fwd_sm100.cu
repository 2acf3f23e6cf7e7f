// Modality-mutual attention forward for sm_100a: QK^T, online softmax and PV on tcgen05 tensor cores with
// TMEM accumulators, operands staged by TMA, mbarrier pipelines, warp-specialised roles.
//
// Replaces the eager core of Phi3Attention.forward (softmax_fp32(QK^T/sqrt(96) + mask) V; installed
// equivalent transformers/models/phi3/modeling_phi3.py:153-175) fed by the reference's materialised
// (B,1,T,T) mask (codes/open_flamingo/src/vlm.py:410-443).  No mask is read from HBM: the predicate
//   allowed(i,j) = (j<=i & valid[j]) | (row_lo[i]<=j<row_hi[i] & mutual_ok[j])
// is evaluated in registers on the few tiles that are not fully visible, and key tiles beyond
// q_tile_kv_end[b][qt] are never visited.  RoPE is applied to Q in shared memory right after the TMA load.
//
// CTA = 2 query tiles x 128 rows of one (batch, head); a PASS covers 128 keys = two 64-key softmax tiles; 20 warps:
//   warp 0      TMA producer (Q once; K as 128-key tiles through a 2-deep ring, V as 64-key tiles through a 4-deep ring)
//   warp 1 / 3  QK^T issuers of query tile 0 / 1: [S_t(j) | S_t(j+1)] = Q_t [K_j ; K_j+1]^T, ONE SS group of N=128 per pass
//               (an M=128,K=16 MMA costs 64-76 cycles whether N is 64 or 128: tools/mma_mix_bench.cu)
//   warp 2      TMEM allocator, then PV issuer of both tiles: O_t += P_t V_j (TS, P read from TMEM); polls both P_FULL
//   (warp 3 first finds the first key tile that holds padding)
//   warps 4-19  softmax: warp = (tile t, column half c, lane group g); thread <-> row 32g+lane <-> TMEM lane,
//               32 of the 64 key columns of each softmax tile.  FOUR softmax warps per SM sub-partition: ncu on the
//               2-warp layout showed the exp2 pipe 47% busy because one warp's fixed latencies (mbarrier probes, TMEM
//               round trips, max chain) exceed its own exp2 time.  The two column halves of a row agree on the pass
//               maximum through shared memory and a 64-thread named barrier.
// Per pass a softmax warp: waits for both score halves, loads its 64 scores (2 x tcgen05.ld.x32), masks the tiles that
// are not fully visible (j >= n_full), reduces the maximum (FMNMX3), exponentiates tile j OPTIMISTICALLY against the
// running maximum of the earlier passes while the maxima are exchanged (redo + rescale only when the new maximum exceeds
// it by 2^8), publishes P(j) (PV(j) enters the tensor pipe), hands both S buffers back (QK^T of the next pass), then
// exponentiates tile j+1 under PV(j) and publishes P(j+1).  Element-wise math is packed (FFMA2 / FADD2).
// Neither the exp2 (MUFU, 49 %) nor the tensor pipe (36 %) is saturated: a pass is bound by this serial chain with the two
// query tiles' exp2 phases coinciding (DESIGN.md 4.1 lists what was tried against that).
// TMEM columns: S0 [0,128) S1 [128,256) (two 64-column halves each) | O0 [256,352) O1 [352,448) | P0 [448,480) P1 [480,512).
// Shared memory: Q 2x24 KB; K ring 2x24 KB; V ring 4x12 KB.  Every tile is 3 SWIZZLE_64B atoms [rows][64 B]
// (head_dim 96 = 3 x 32), the layout both the TMA boxes and the UMMA descriptors use (tools/umma_probe.cu).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "attn_aux.cuh"
#include "sm100_ptx.cuh"

namespace aki {

namespace fwd {
constexpr int BM = 128, BN = 64, HD = 96;
constexpr int Q_ATOM = 128 * 64, Q_TILE = 3 * Q_ATOM;     // 24576
constexpr int KV_ATOM = BN * 64, KV_TILE = 3 * KV_ATOM;   // 12288
constexpr int STAGES = 4;                                 // V ring: 64-key tiles
constexpr int K_PAIR_ATOM = 2 * BN * 64, K_PAIR_TILE = 3 * K_PAIR_ATOM;   // K ring: 128-key tiles (one per pass), 24576
constexpr int K_STAGES = 2;
constexpr int THREADS = 640;
constexpr int SMEM_Q = 0;
constexpr int SMEM_K = SMEM_Q + 2 * Q_TILE;
constexpr int SMEM_V = SMEM_K + K_STAGES * K_PAIR_TILE;
constexpr int SMEM_TOTAL = SMEM_V + STAGES * KV_TILE;     // 147456
constexpr int SMEM_ALLOC = SMEM_TOTAL + 1024;             // slack for 1024-byte alignment
constexpr uint32_t TM_S = 0, TM_O = 256, TM_P = 448;      // S: tile t buffer u at 128 t + 64 u; O: 96 t; P: 32 t
constexpr int REGS_CTRL = 56, REGS_SOFTMAX = 104;         // CTA pool (640 x 96 = 61440): 128*56 + 512*104 = 60416
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
  unsigned long long* trace;  // debug (AKI_MMA_FWD_TRACE=<cta>, tools/fwd_trace.py): clock64 stamps of one CTA
  int trace_cta;
};

#ifdef AKI_FWD_TRACE
#define TR(slot, j, k) do { if (tracing && (j) < 64) P.trace[((slot) * 64 + (j)) * 8 + (k)] = clock64(); } while (0)
#else
#define TR(slot, j, k) do { } while (0)
#endif

__device__ __forceinline__ uint32_t low_mask(int n) {  // n low bits set, n clamped to [0,32]
  return n <= 0 ? 0u : (n >= 32 ? 0xffffffffu : ((1u << n) - 1u));
}

// D[tmem] (+)= A[smem] * B[smem] with descriptors given as (low word, shared high word)
__device__ __forceinline__ void umma_ss_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_ts_lh(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t desc_hi,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Rare path of the online softmax: the running max grew by more than the threshold, this thread's 48 columns of O_t
// are rescaled by alpha.  Kept out of line so that the per-tile loop stays compact.
__device__ __noinline__ void rescale_o48(uint32_t tm_o, float alpha) {
  uint32_t o[32];
  tmem_ld_x32(tm_o, o);
  tmem_wait_ld();
#pragma unroll
  for (int x = 0; x < 32; ++x) o[x] = __float_as_uint(__uint_as_float(o[x]) * alpha);
  tmem_st_x32(tm_o, o);
  tmem_ld_x16(tm_o + 32, o);
  tmem_wait_ld();
#pragma unroll
  for (int x = 0; x < 16; ++x) o[x] = __float_as_uint(__uint_as_float(o[x]) * alpha);
  tmem_st_x16(tm_o + 32, o);
  tmem_wait_st();
}

template <bool ROPE>
__global__ void __launch_bounds__(fwd::THREADS, 1)
attn_fwd_sm100_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                      const __grid_constant__ CUtensorMap map_v, const FwdKernelParams P) {
  using namespace fwd;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // barrier indices
  constexpr int Q_FULL = 0, Q_READY = 2, K_FULL = 4, K_EMPTY = K_FULL + K_STAGES, V_FULL = K_EMPTY + K_STAGES,
                V_EMPTY = V_FULL + STAGES, S_FULL = V_EMPTY + STAGES /* [t][buf] */, S_FREE = S_FULL + 4,
                P_FULL = S_FREE + 4, O_FULL = P_FULL + 2, N_BARS = O_FULL + 2;
  __shared__ __align__(8) uint64_t bars[N_BARS];
  __shared__ uint32_t tmem_base_s;
  __shared__ int first_bad_s;
  __shared__ float xch[2][2][2][128];   // [key-tile parity][query tile][column half][row]: tile maxima, then row sums
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };

  const int tid = threadIdx.x, warp = tid >> 5;
  // work decomposition: consecutive CTAs share (b,h) so K/V stay in L2; heaviest query tiles first
  const int bh = blockIdx.x / P.n_qp;
  const int qp = P.n_qp - 1 - (blockIdx.x % P.n_qp);
  const int b = bh / P.H, h = bh % P.H;
  const int n_kt = (P.T + BN - 1) / BN;            // 64-key tiles
  int n_kv0, n_kv1;
  {
    int n[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int qt = 2 * qp + t;
      if (qt >= P.n_qt) n[t] = 0;
      else if (P.mm.q_tile_kv_end) n[t] = min(2 * P.mm.q_tile_kv_end[(size_t)b * P.n_qt + qt], n_kt);  // 128 -> 64 units
      else n[t] = min(2 * (qt + 1), n_kt);
    }
    n_kv0 = n[0]; n_kv1 = n[1];
  }
  const int n_max = max(n_kv0, n_kv1);

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(Q_FULL + i), 1); mbar_init(BAR(Q_READY + i), 256); }
    for (int i = 0; i < K_STAGES; ++i) { mbar_init(BAR(K_FULL + i), 1); mbar_init(BAR(K_EMPTY + i), 2); }
    for (int i = 0; i < STAGES; ++i) { mbar_init(BAR(V_FULL + i), 1); mbar_init(BAR(V_EMPTY + i), 2); }   // one release per query tile
    for (int i = 0; i < 4; ++i) { mbar_init(BAR(S_FULL + i), 1); mbar_init(BAR(S_FREE + i), 256); }
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(P_FULL + i), 256); mbar_init(BAR(O_FULL + i), 1); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(smem_u32(&tmem_base_s));
  if (warp == 0 && elect_one()) { tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v); }
  if (warp == 3) {
    // first 64-key tile (among those this CTA visits) that is not entirely inside the sequence and causally valid
    const int lane = tid & 31;
    const int len = meta_len(P.mm, b, P.T);
    int first = n_max;
    for (int base = 0; base < n_max; base += 32) {
      const int jt = base + lane;
      bool bad = false;
      if (jt < n_max) {
        bad = (jt * BN + BN > len);
        if (!bad && P.mm.vbits) {
          const uint32_t* w = P.mm.vbits + (size_t)b * P.mm.bits_pitch + 2 * jt;
          bad = (__ldg(w) != 0xffffffffu) || (__ldg(w + 1) != 0xffffffffu);
        }
      }
      const uint32_t m = __ballot_sync(0xffffffffu, bad);
      if (m) { first = base + __ffs(m) - 1; break; }
    }
    if (lane == 0) first_bad_s = first;
  }
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
        mbar_arrive_expect_tx(BAR(Q_FULL + t), Q_TILE);
        for (int a = 0; a < 3; ++a)
          tma_load_4d(smem_base + SMEM_Q + t * Q_TILE + a * Q_ATOM, &map_q, BAR(Q_FULL + t), a * 32, (2 * qp + t) * BM, h, b);
      }
      auto load_k = [&](int j) {                 // the 128 keys of pass j/2 (rows beyond T are zero-filled)
        const int pp = j >> 1, s = pp % K_STAGES;
        mbar_wait(BAR(K_EMPTY + s), ((pp / K_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(BAR(K_FULL + s), K_PAIR_TILE);
        for (int a = 0; a < 3; ++a)
          tma_load_4d(smem_base + SMEM_K + s * K_PAIR_TILE + a * K_PAIR_ATOM, &map_k, BAR(K_FULL + s), a * 32, j * BN, h, b);
      };
      auto load_v = [&](int j) {
        const int s = j % STAGES;
        mbar_wait(BAR(V_EMPTY + s), ((j / STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(BAR(V_FULL + s), KV_TILE);
        for (int a = 0; a < 3; ++a)
          tma_load_4d(smem_base + SMEM_V + s * KV_TILE + a * KV_ATOM, &map_v, BAR(V_FULL + s), a * 32, j * BN, h, b);
      };
      // consumption order of the MMA warps (two key tiles per pass): K01 | K23 V0 V1 | K45 V2 V3 | ...
      if (n_max > 0) load_k(0);
      for (int j = 0; j < n_max; j += 2) {
        if (j + 2 < n_max) load_k(j + 2);
        load_v(j);
        if (j + 1 < n_max) load_v(j + 1);
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ------------------------------------------------------------------ QK^T issuers: warp 1 -> query tile 0, warp 3 -> tile 1
    // An issuing thread streams M=128,K=16 MMAs at ~64-80 cycles each whatever N is (tools/mma_mix_bench.cu), so the
    // MMA work is spread over three warps: one QK^T issuer per query tile and one PV issuer (warp 2).  The whole warp
    // runs the loop so that addresses / descriptors stay in uniform registers; one elected lane issues.  Every K / V
    // stage is released by one arrival per query tile: a commit behind the MMA that read it, or a plain arrive when
    // that tile does not visit the key tile.
    setmaxnreg_dec<REGS_CTRL>();
    const int t = (warp == 3) ? 1 : 0;
    const int nk = t ? n_kv1 : n_kv0;
    const bool leader = elect_one();
#ifdef AKI_FWD_TRACE
    const bool tracing = P.trace && (int)blockIdx.x == P.trace_cta && leader;
#endif
    // One N=128 MMA group per pass: S_t(j) | S_t(j+1) = Q_t [K_j ; K_j+1]^T lands in the tile's 128 score columns.  An
    // M=128,K=16 MMA costs ~64-76 cycles whether N is 64 or 128, so two N=64 groups took twice the tensor-pipe time.
    constexpr uint32_t IDESC_QK128 = umma_idesc_bf16(BM, 2 * BN, 0, 0), IDESC_QK64 = umma_idesc_bf16(BM, BN, 0, 0);
    const uint64_t DESC_KMAJ = umma_smem_desc(0, 16, 512, UMMA_SW64);
    const uint32_t HI = (uint32_t)(DESC_KMAJ >> 32), KMAJ_LO = (uint32_t)DESC_KMAJ;
    const uint32_t qa = KMAJ_LO + ((smem_base + SMEM_Q + t * Q_TILE) >> 4), k_lo = KMAJ_LO + ((smem_base + SMEM_K) >> 4);
    if (nk > 0) mbar_wait(BAR((ROPE ? Q_READY : Q_FULL) + t), 0);
    // Pass p handles key tiles j = 2p, 2p+1.  The softmax frees both S buffers at the same moment; the chain
    // S_FREE -> S(j+2), S(j+3) is the critical path of a pass.  Every K stage is released by one arrival per query
    // tile: a commit behind the MMAs that read it, or a plain arrive when this tile does not visit these keys.
    for (int j = 0; j < n_max; j += 2) {
      const int pp = j >> 1, s = pp % K_STAGES;
      // waited for even when this tile skips the keys: it keeps the tile from running a whole ring ahead and
      // arriving twice in one K_EMPTY phase
      mbar_wait(BAR(K_FULL + s), (pp / K_STAGES) & 1);
      if (j < nk) {
        const bool two = (j + 1 < nk);
        if (j >= 2) {                           // the softmax holds S_t(j-2), S_t(j-1) in registers
          mbar_wait(BAR(S_FREE + 2 * t + 0), ((j - 2) >> 1) & 1);
          if (two) mbar_wait(BAR(S_FREE + 2 * t + 1), ((j - 1) >> 1) & 1);
        }
        tc_fence_after();
        TR(4 + t, j, 0);
        const uint32_t ka = k_lo + s * (K_PAIR_TILE >> 4);
        const uint32_t d = tmem + TM_S + 128 * t;
        if (leader) {
#pragma unroll
          for (int k = 0; k < 6; ++k)
            umma_ss_lh(d, qa + (((k >> 1) * Q_ATOM + (k & 1) * 32) >> 4), ka + (((k >> 1) * K_PAIR_ATOM + (k & 1) * 32) >> 4), HI,
                       two ? IDESC_QK128 : IDESC_QK64, k > 0);
          umma_commit(BAR(S_FULL + 2 * t + 0));
          if (two) umma_commit(BAR(S_FULL + 2 * t + 1));
          umma_commit(BAR(K_EMPTY + s));
        }
        TR(4 + t, j, 1);
      } else if (leader) {
        mbar_arrive(BAR(K_EMPTY + s));
      }
      __syncwarp();
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ PV issuer of both query tiles
    // O_t += P_t V_j (TS, P read from TMEM).  Polls the two tiles' P_FULL barriers and serves whichever is ready, so
    // neither tile waits behind the other; tiles that do not visit key tile j just release the V stage.
    setmaxnreg_dec<REGS_CTRL>();
    const bool leader = elect_one();
#ifdef AKI_FWD_TRACE
    const bool tracing = P.trace && (int)blockIdx.x == P.trace_cta && leader;
#endif
    constexpr uint32_t IDESC_PV = umma_idesc_bf16(BM, HD, 0, 1);
    const uint64_t DESC_V = umma_smem_desc(0, KV_ATOM, 512, UMMA_SW64);     // MN-major: LBO = atom stride
    const uint32_t HI = (uint32_t)(DESC_V >> 32), v_lo = (uint32_t)DESC_V + ((smem_base + SMEM_V) >> 4);
    int jt[2] = {0, 0};
    while (jt[0] < n_max || jt[1] < n_max) {
      bool progressed = false;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int j = jt[t];
        if (j >= n_max) continue;
        const int nk = t ? n_kv1 : n_kv0;
        const int s = j % STAGES;
        if (!mbar_test(BAR(V_FULL + s), (j / STAGES) & 1)) continue;
        if (j < nk) {
          if (!mbar_test(BAR(P_FULL + t), j & 1)) continue;
          tc_fence_after();
          TR(4 + t, j, 3);
          const uint32_t va = v_lo + s * (KV_TILE >> 4);
          const uint32_t d_o = tmem + TM_O + 96 * t, a_p = tmem + TM_P + 32 * t;
          if (leader) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ts_lh(d_o, a_p + 8 * k, va + k * 64, HI, IDESC_PV, (j > 0 || k > 0));
            umma_commit(BAR(O_FULL + t));
            umma_commit(BAR(V_EMPTY + s));
          }
          TR(4 + t, j, 4);
        } else if (leader) {
          mbar_arrive(BAR(V_EMPTY + s));
        }
        __syncwarp();
        jt[t] = j + 1;
        progressed = true;
      }
#ifndef AKI_PV_SLEEP
#define AKI_PV_SLEEP 32         // ns between polls of the PV issuer when neither tile is ready (A/B: 0 = busy poll)
#endif
      if (!progressed && AKI_PV_SLEEP > 0) __nanosleep(AKI_PV_SLEEP);
    }
  } else {
    // ------------------------------------------------------------------ softmax / correction / epilogue
    setmaxnreg_inc<REGS_SOFTMAX>();
    const int sw = warp - 4;
    const int t = sw >> 3, ch = (sw >> 2) & 1, g = sw & 3;   // query tile, column half, lane group (== warp % 4)
    const int r = 32 * g + (tid & 31);            // row within the tile == TMEM lane
    const int qt = 2 * qp + t;
    const int i = qt * BM + r;                    // query index in mask coordinates
    const int len = meta_len(P.mm, b, P.T);
    const uint32_t lane_base = (uint32_t)(g * 32) << 16;
    const uint32_t tm_s = tmem + TM_S + 128 * t + 32 * ch + lane_base;
    const uint32_t tm_o = tmem + TM_O + 96 * t + 48 * ch + lane_base;
    const uint32_t tm_p = tmem + TM_P + 32 * t + 16 * ch + lane_base;
    const int pair_bar = 1 + 4 * t + g;           // named barrier of the two warps that share these rows
    const int nk = t ? n_kv1 : n_kv0;
    // key tiles j < n_full lie entirely below the diagonal of this query tile and hold no padding
    const int n_full = min(first_bad_s, 2 * qt);

    if (ROPE && nk > 0) {
      mbar_wait(BAR(Q_FULL + t), 0);
      if (i < P.T) {
        const uint32_t qa = smem_base + SMEM_Q + t * Q_TILE;
        const float* cr = P.rope_cos + (size_t)b * P.rope_stride_b + (size_t)i * 48;
        const float* sr = P.rope_sin + (size_t)b * P.rope_stride_b + (size_t)i * 48;
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {          // 16-byte chunk c pairs with chunk c+6 (d <-> d+48)
          const int c = 3 * ch + cc;
          const uint32_t a_lo = qa + (c >> 2) * Q_ATOM + sw64_offset(r, c & 3);
          const uint32_t a_hi = qa + ((c + 6) >> 2) * Q_ATOM + sw64_offset(r, (c + 6) & 3);
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
    float l = 0.f;             // row sum over this warp's column halves
    int o_waited = 0;          // number of O_FULL phases this thread has already observed

#ifdef AKI_FWD_TRACE
    const bool tracing = P.trace && (int)blockIdx.x == P.trace_cta && g == 0 && (tid & 31) == 0;
    const int slot = 2 * t + ch;
#endif
    // All four softmax warps of an SM sub-partition share its 4 MUFU lanes.  Tile 1 starts when tile 0 has finished
    // its first batch of exponentials, so that from then on one tile's exp2 phase covers the other tile's barrier
    // probes / TMEM round trips (clock64 trace: started together, the tiles stay in phase and the exp2 phase of
    // all four warps takes 1240 cycles while the pipe idles for the other 1100 of each key tile).
#ifndef AKI_PV_FIRST
#define AKI_PV_FIRST 1          // 1: publish P(j) before handing the S buffers back, so that PV(j) queues ahead of the
                                // next pass's QK^T in the tensor pipe (same-box A/B: 1.047 -> 1.038 ms causal); 0: the reverse
#endif
#ifndef AKI_ONE_EXP_PHASE
#define AKI_ONE_EXP_PHASE 0     // 1: both key tiles of a pass are exponentiated in ONE phase before P(j) is published (A/B)
#endif
#ifndef AKI_STAGGER_POINT
#define AKI_STAGGER_POINT 1     // where tile 0 releases tile 1: 0 never staggered, 1 after its first exponentials (shipped),
#endif                          // 2 after its first TMEM load, 3 at the end of its first pass  (A/B builds only)
    const bool stagger = (AKI_STAGGER_POINT != 0) && (n_kv0 > 0 && n_kv1 > 0);
    if (stagger && t == 1) named_bar_sync(9, 512);
    // Two key tiles per pass: the fixed per-tile latencies (mbarrier probes, TMEM round trips, the max exchange of the
    // two column halves) cost ~1000 cycles against ~500 of exponentials, so they are paid once per PAIR of key tiles:
    // both S buffers are fetched together, one exchange covers both, P(j) is published and PV(j) runs while the
    // exponentials of tile j+1 are computed, then P(j+1) follows.  Exponentials overwrite the scores in place (bf16
    // pairs compacted into the low registers), so the 64 scores of a pass are the only large register array.
    // (A lock that made the two tiles' exp2 phases alternate on the 4 MUFU lanes of a sub-partition was tried: the
    // extra barrier + CAS round trips cost more than the idle MUFU time they removed, 1.38 -> 1.59 ms.)
    bool s_ready0 = false, s_ready1 = false;        // early probes of S_FULL for the next pass
    for (int j = 0; j < nk; j += 2) {
      TR(slot, j, 0);
      const bool two = (j + 1 < nk);                // warp-uniform
      const int xp = (j >> 1) & 1;                  // exchange-buffer parity of this pass
      const int c0 = j * BN + 32 * ch;              // first key column of this thread's half of tile j
      const bool partial0 = (j >= n_full), partial1 = two && (j + 1 >= n_full);
      uint32_t vw0 = 0xffffffffu, mw0 = 0xffffffffu, vw1 = 0xffffffffu, mw1 = 0xffffffffu;
      if (partial0) {                               // one 32-bit word of each bit-vector covers a half tile
        if (P.mm.vbits) vw0 = __ldg(P.mm.vbits + (size_t)b * P.mm.bits_pitch + (c0 >> 5));
        if (P.mm.mbits) mw0 = __ldg(P.mm.mbits + (size_t)b * P.mm.bits_pitch + (c0 >> 5));
      }
      if (partial1) {
        if (P.mm.vbits) vw1 = __ldg(P.mm.vbits + (size_t)b * P.mm.bits_pitch + ((c0 + BN) >> 5));
        if (P.mm.mbits) mw1 = __ldg(P.mm.mbits + (size_t)b * P.mm.bits_pitch + ((c0 + BN) >> 5));
      }
      // probe "PV(j-1) has consumed P(j-1)" now, use the answer just before the P store
      const bool o_known = (j == 0) || (o_waited >= j) || mbar_test(BAR(O_FULL + t), (j - 1) & 1);

      float s[64];                                   // [0,32): this half of S_t(j); [32,64): this half of S_t(j+1)
      auto mask32 = [&](int off, int col0, uint32_t vw, uint32_t mw) {
        const int d = row_live ? (i - col0) : -1;             // causal: column c visible iff c <= d
        const int a = row_lo - col0, e = row_hi - col0;       // mutual: a <= c < e
        const uint32_t in_len = low_mask(len - col0);
        const uint32_t causal = low_mask(d + 1) & vw & in_len;
        const uint32_t mutual = row_live ? (low_mask(e) & ~low_mask(a) & mw & in_len) : 0u;
        const uint32_t ok = causal | mutual;
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (!((ok >> c) & 1u)) s[off + c] = -INFINITY;
      };
      auto fetch0 = [&]() {                          // (re)load this half of S_t(j) and mask it
        tmem_ld_x32(tm_s + 64 * (j & 1), reinterpret_cast<uint32_t*>(s));
        tmem_wait_ld();
        if (partial0) mask32(0, c0, vw0, mw0);
      };
      if (!s_ready0) mbar_wait(BAR(S_FULL + 2 * t + (j & 1)), (j >> 1) & 1);
      if (two && !s_ready1) mbar_wait(BAR(S_FULL + 2 * t + ((j + 1) & 1)), ((j + 1) >> 1) & 1);
      tc_fence_after();
      TR(slot, j, 1);
      tmem_ld_x32(tm_s + 64 * (j & 1), reinterpret_cast<uint32_t*>(s));
      if (two) tmem_ld_x32(tm_s + 64 * ((j + 1) & 1), reinterpret_cast<uint32_t*>(s) + 32);
      tmem_wait_ld();
      TR(slot, j, 2);
      if (AKI_STAGGER_POINT == 2 && stagger && t == 0 && j == 0) asm volatile("bar.arrive 9, 512;" ::: "memory");
      if (partial0) mask32(0, c0, vw0, mw0);
      if (partial1) mask32(32, c0 + BN, vw1, mw1);
      // ---- maximum of this thread's scores of the pass; the partner half's arrives through shared memory
      // three-input maxima (FMNMX3): two scores per instruction, four independent chains
      float mx[4] = {fmax3(s[0], s[1], s[2]), fmax3(s[3], s[4], s[5]), fmax3(s[6], s[7], s[8]), fmax3(s[9], s[10], s[11])};
#pragma unroll
      for (int c = 12; c < 28; c += 8) {
        mx[0] = fmax3(mx[0], s[c], s[c + 1]); mx[1] = fmax3(mx[1], s[c + 2], s[c + 3]);
        mx[2] = fmax3(mx[2], s[c + 4], s[c + 5]); mx[3] = fmax3(mx[3], s[c + 6], s[c + 7]);
      }
      mx[0] = fmax3(mx[0], s[28], s[29]); mx[1] = fmax3(mx[1], s[30], s[31]);
      if (two) {
#pragma unroll
        for (int c = 32; c < 64; c += 8) {
          mx[0] = fmax3(mx[0], s[c], s[c + 1]); mx[1] = fmax3(mx[1], s[c + 2], s[c + 3]);
          mx[2] = fmax3(mx[2], s[c + 4], s[c + 5]); mx[3] = fmax3(mx[3], s[c + 6], s[c + 7]);
        }
      }
      const float m_half = fmaxf(fmax3(mx[0], mx[1], mx[2]), mx[3]);
      xch[xp][t][ch][r] = m_half;
      // exp2((S - m_used) * scale*log2e) of 32 scores starting at `off`, packed IN PLACE into s[off .. off+16)
      auto exps = [&](int off) {
        const float neg_m = (m_used == -INFINITY) ? 0.f : -m_used * P.scale_log2;
        // packed FFMA2 / FADD2: one scale-and-shift and one row-sum instruction per PAIR of scores
        const uint64_t sc2 = f32x2_pack(P.scale_log2, P.scale_log2), nm2 = f32x2_pack(neg_m, neg_m);
        uint64_t sum_a = f32x2_pack(0.f, 0.f), sum_b = sum_a;
#pragma unroll
        for (int x = 0; x < 16; ++x) {
          float a0, a1;
          f32x2_unpack(f32x2_fma(f32x2_pack(s[off + 2 * x], s[off + 2 * x + 1]), sc2, nm2), a0, a1);
          const float p0 = ex2_approx(a0), p1 = ex2_approx(a1);
          if (x & 1) sum_b = f32x2_add(sum_b, f32x2_pack(p0, p1));
          else sum_a = f32x2_add(sum_a, f32x2_pack(p0, p1));
          s[off + x] = __uint_as_float(pack_bf16x2(p0, p1));
        }
        float t0, t1;
        f32x2_unpack(f32x2_add(sum_a, sum_b), t0, t1);
        return t0 + t1;
      };
      auto fetch1 = [&]() {                          // (re)load this half of S_t(j+1) and mask it
        tmem_ld_x32(tm_s + 64 * ((j + 1) & 1), reinterpret_cast<uint32_t*>(s) + 32);
        tmem_wait_ld();
        if (partial1) mask32(32, c0 + BN, vw1, mw1);
      };
      float sum_j;
      if (j == 0) {
        named_bar_sync(pair_bar, 64);
        m_used = fmaxf(m_half, xch[xp][t][ch ^ 1][r]);
        sum_j = exps(0);
        if (AKI_ONE_EXP_PHASE && two) sum_j += exps(32);
      } else {
        // Optimistic: exponentiate tile j against the running max of the EARLIER passes while the maxima are being
        // exchanged; if the pass maximum exceeds the reference by more than the threshold (rare after the first
        // tiles) O and l are rescaled and tile j is redone -- published P values never exceed 2^threshold.
        sum_j = exps(0);
        if (AKI_ONE_EXP_PHASE && two) sum_j += exps(32);
        named_bar_sync(pair_bar, 64);
        const float m_new = fmaxf(m_used, fmaxf(m_half, xch[xp][t][ch ^ 1][r]));
        TR(slot, j, 3);
        const bool need = (m_new - m_used) * P.scale_log2 > RESCALE_THRESHOLD || (m_used == -INFINITY && m_new > -INFINITY);
        if (__any_sync(0xffffffffu, need)) {
          const float alpha = (m_used == -INFINITY) ? 0.f : ex2_approx((m_used - m_new) * P.scale_log2);
          m_used = m_new;
          l *= alpha;
          if (o_waited < j) { mbar_wait(BAR(O_FULL + t), (j - 1) & 1); o_waited = j; }   // PV(j-1) has landed
          tc_fence_after();
          rescale_o48(tm_o, alpha);
          fetch0();
          sum_j = exps(0);
          if (AKI_ONE_EXP_PHASE && two) { fetch1(); sum_j += exps(32); }
        }
      }
      l += sum_j;
      if (AKI_STAGGER_POINT == 1 && stagger && t == 0 && j == 0) asm volatile("bar.arrive 9, 512;" ::: "memory");   // release tile 1
      // ---- both S buffers go back to the MMA warp: QK^T(j+2), QK^T(j+3) run during the rest of this pass
      if (!AKI_PV_FIRST) {
        tc_fence_before();
        mbar_arrive(BAR(S_FREE + 2 * t + (j & 1)));
        if (two) mbar_arrive(BAR(S_FREE + 2 * t + ((j + 1) & 1)));
      }
      TR(slot, j, 4);
      // ---- publish P(j): its own TMEM columns, single-buffered -- PV(j-1) has consumed P(j-1)
      if (o_waited < j) {
        if (!o_known) mbar_wait(BAR(O_FULL + t), (j - 1) & 1);
        o_waited = j;
      }
      TR(slot, j, 5);
      tmem_st_x16(tm_p, reinterpret_cast<const uint32_t*>(s));
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(BAR(P_FULL + t));
      if (AKI_PV_FIRST) {            // PV(j) enters the tensor-pipe queue ahead of QK^T(j+2), QK^T(j+3)
        mbar_arrive(BAR(S_FREE + 2 * t + (j & 1)));
        if (two) mbar_arrive(BAR(S_FREE + 2 * t + ((j + 1) & 1)));
      }
      if (two) {
        // ---- tile j+1 while PV(j) runs, then P(j+1) into the same columns
        if (!AKI_ONE_EXP_PHASE) l += exps(32);
        mbar_wait(BAR(O_FULL + t), j & 1);
        o_waited = j + 1;
        tc_fence_after();
        tmem_st_x16(tm_p, reinterpret_cast<const uint32_t*>(s) + 32);
      }
      // probe S_FULL of the next pass while the P store drains
      s_ready0 = (j + 2 < nk) && mbar_test(BAR(S_FULL + 2 * t + (j & 1)), ((j + 2) >> 1) & 1);
      s_ready1 = (j + 3 < nk) && mbar_test(BAR(S_FULL + 2 * t + ((j + 1) & 1)), ((j + 3) >> 1) & 1);
      if (two) {
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(BAR(P_FULL + t));
      }
      TR(slot, j, 6);
      if (AKI_STAGGER_POINT == 3 && stagger && t == 0 && j == 0) asm volatile("bar.arrive 9, 512;" ::: "memory");
    }

    // ---- epilogue: O / l -> bf16 -> global; LSE
    if (qt < P.n_qt) {
      float inv_l = 0.f;
      if (nk > 0) {
        // total row sum = sum of the two halves (the exchange buffer the last pass did not use)
        const int xe = (((nk - 1) >> 1) + 1) & 1;
        xch[xe][t][ch][r] = l;
        named_bar_sync(pair_bar, 64);
        l += xch[xe][t][ch ^ 1][r];
        mbar_wait(BAR(O_FULL + t), (nk - 1) & 1);
        tc_fence_after();
        inv_l = (row_live && l > 0.f) ? 1.f / l : 0.f;  // batch-padding rows: zeros
      }
      // tcgen05.ld is warp-collective (.sync.aligned): load unconditionally, predicate only the global stores
      const bool store_row = (i < P.T);
      __nv_bfloat16* orow = P.o.row(b, store_row ? i : 0, h) + 48 * ch;
      uint32_t o[48];
      if (nk > 0) {
        tmem_ld_x32(tm_o, o);
        tmem_ld_x16(tm_o + 32, o + 32);
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
      if (P.lse && store_row && ch == 0)
        P.lse[((size_t)b * P.H + h) * P.T + i] = (inv_l > 0.f) ? (m_used * P.scale + __logf(l)) : INFINITY;
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

// row statistics (B,H,t_pad,8) bf16 contiguous -> [8 x 128 x 1 x 1] boxes, no swizzle: one box is the K-major
// SWIZZLE_NONE operand [16 row groups][8 rows][16 B] of the backward's statistics k-step
int make_row_stats_map(CUtensorMap* m, void* base, int B, int H, int t_pad) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_last_cuda_error("cuTensorMapEncodeTiled entry point not found"); return AKI_ERR_CUDA; }
  cuuint64_t dims[4] = {8, (cuuint64_t)t_pad, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {16, (cuuint64_t)t_pad * 16, (cuuint64_t)H * t_pad * 16};
  cuuint32_t box[4] = {8, 128, 1, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_last_cuda_error("cuTensorMapEncodeTiled(row stats) failed"); return AKI_ERR_CUDA; }
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
  if ((rc = make_tile_map(&mk, p->k, p->B, p->H, p->T, 2 * fwd::BN))) return rc;   // K: 128-key tiles (one per pass)
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
#ifdef AKI_FWD_TRACE
  // Debug build only (make TRACE=1; tools/fwd_trace.py): dumps clock64 stamps of one CTA and SYNCHRONISES.
  const char* trace_env = getenv("AKI_MMA_FWD_TRACE");
  const size_t trace_bytes = 6 * 64 * 8 * sizeof(unsigned long long);
  if (trace_env) {
    kp.trace_cta = atoi(trace_env);
    cudaMalloc(&kp.trace, trace_bytes);
    cudaMemset(kp.trace, 0, trace_bytes);
  }
#endif
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
  timing_hook_begin(st);
  if (p->rope_cos)
    attn_fwd_sm100_kernel<true><<<(unsigned)grid, fwd::THREADS, fwd::SMEM_ALLOC, st>>>(mq, mk, mv, kp);
  else
    attn_fwd_sm100_kernel<false><<<(unsigned)grid, fwd::THREADS, fwd::SMEM_ALLOC, st>>>(mq, mk, mv, kp);
  timing_hook_end(st);
#ifdef AKI_FWD_TRACE
  if (trace_env) {
    cudaDeviceSynchronize();
    static unsigned long long host[6 * 64 * 8];
    cudaMemcpy(host, kp.trace, trace_bytes, cudaMemcpyDeviceToHost);
    cudaFree(kp.trace);
    unsigned long long t0 = ~0ull;
    for (size_t i = 0; i < 6 * 64 * 8; ++i) if (host[i] && host[i] < t0) t0 = host[i];
    const char* names[6] = {"sm_t0c0", "sm_t0c1", "sm_t1c0", "sm_t1c1", "mma_t0", "mma_t1"};
    for (int slot = 0; slot < 6; ++slot)
      for (int j = 0; j < 64; ++j) {
        if (!host[(slot * 64 + j) * 8]) continue;
        fprintf(stderr, "TRACE %s j=%d:", names[slot], j);
        for (int k = 0; k < 7; ++k) fprintf(stderr, " %llu", host[(slot * 64 + j) * 8 + k] ? host[(slot * 64 + j) * 8 + k] - t0 : 0ull);
        fprintf(stderr, "\n");
      }
  }
#endif
  return check_launch();
}
