// EXPERIMENT (not compiled into the library): backward with 64-query sub-tiles, S^T double-buffered, dQ in its own TMEM
// columns and P / dS stages in separate warpgroups.  Correct (same parity as the shipped kernel) but slower on B200
// (3.9-4.3 ms vs 3.4 ms at T=8192 B=2): N=64 SS MMAs are shared-memory-bandwidth bound (6 KB per 32-cycle MMA), and with
// 227 KB of smem the dS^T buffer / Q ring / dQ staging cannot all be double-buffered, so the stages stay coupled.
// Trace: profiles/r01_bwd_trace_v3_subtile.txt.  Kept as the starting point for the next round.
// Modality-mutual attention backward for sm_100a (tcgen05 / TMEM / TMA).
//
// Gradient of softmax_fp32(QK^T*scale + mask) V (the eager core of Phi3Attention.forward; installed equivalent
// transformers/models/phi3/modeling_phi3.py:153-175; the reference obtains it from autograd over five passes
// on a (B,32,T,T) tensor).  One CTA owns one 128-key tile of one (batch, head) and walks exactly the query tiles
// that can see it (kv_tile_q_mask: with MMA that set is the image-row tiles before the diagonal plus everything
// from the diagonal on).  Everything is transposed so that keys sit on TMEM lanes, and a 128-query tile is
// processed as two 64-query sub-tiles g = 2*tile + u so that every stage can be double-buffered / decoupled
// inside the 512 TMEM columns:
//     S^T_g  = K Q_g^T        (SS, N=64)      P^T_g  = exp2(S^T*c - LSE)            WG1 -> TMEM bf16 (own columns)
//     dP^T_g = V dO_g^T       (SS, N=64)      dS^T_g = P^T o (dP^T - delta)         WG2 -> smem bf16
//     dV  += P^T_g dO_g       (TS, K=64)
//     dK  += dS^T_g Q_g       (SS, K=64; A K-major = dS^T, B MN-major = Q_g)         (x scale in the epilogue)
//     dQ_tile = dS K          (SS, M=128 queries, K=128 keys; A MN-major = the dS^T buffer, B MN-major = K)
// S^T is double-buffered, dQ has its own columns (it used to alias dP^T, which serialised dS -> dQ -> drain ->
// next dP^T and left the tensor pipe idle 70 % of the time), P^T has its own columns, and the two exp / dS stages
// run in different warpgroups on different sub-tiles.  dK / dV accumulate in TMEM over the whole loop; dQ is
// drained by a third warpgroup: TMEM -> registers -> fp32 SWIZZLE_128B staging -> TMA tensor reduce-add into a
// fp32 accumulator in HBM; aki_mma_attn_bwd's finalize kernel applies scale, the inverse RoPE and the bf16 cast.
//
// 16 warps: 0 TMA producer | 1 MMA issuer | 2 TMEM allocator | 3 builds the query-tile list | 4-7 WG1 (P^T) |
// 8-11 WG2 (dS^T) | 12-15 dQ drain.  Threads of WG1/WG2 <-> key row r <-> TMEM lane r.
// TMEM columns: S^T 2x64 [0,128) | dP^T [128,192) | P^T [192,224) | dQ [224,320) | dV [320,416) | dK [416,512).
// Shared memory: K, V 24 KB each (resident), Q ring 2x24 KB, dO ring 2x24 KB, dS^T 32 KB, dQ staging 2x16 KB,
// per-tile row statistics, query-tile list.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "attn_aux.cuh"
#include "sm100_ptx.cuh"

namespace aki {

namespace bwd {
constexpr int BN = 128, BM = 128, HD = 96;
constexpr int ATOM_BYTES = 128 * 64;          // bf16 operand atom [128 rows][64 B], SWIZZLE_64B
constexpr int TILE_BYTES = 3 * ATOM_BYTES;
constexpr int SUB_BYTES = 64 * 64;            // byte offset of the second 64-row sub-tile inside an atom
constexpr int DQ_ATOM_BYTES = 128 * 128;      // fp32 staging atom [128 rows][32 floats], SWIZZLE_128B
constexpr int Q_STAGES = 2, DO_STAGES = 2;
constexpr int THREADS = 512;
constexpr int MAX_TILES = 1024;               // T <= 131072 (Phi-3.5's max_position_embeddings)
constexpr int SMEM_K = 0;
constexpr int SMEM_V = SMEM_K + TILE_BYTES;
constexpr int SMEM_Q = SMEM_V + TILE_BYTES;
constexpr int SMEM_DO = SMEM_Q + Q_STAGES * TILE_BYTES;
constexpr int SMEM_DS = SMEM_DO + DO_STAGES * TILE_BYTES;   // 4 atoms [128 keys][64 B] = 128 keys x 128 queries
constexpr int DQ_BUFS = 2;                                  // dQ staging atoms in flight (TMA reduce latency ~1000 cycles)
constexpr int SMEM_DQ = SMEM_DS + 4 * ATOM_BYTES;
constexpr int SMEM_STATS = SMEM_DQ + DQ_BUFS * DQ_ATOM_BYTES;         // 2 stages x {lse2, delta, lo, hi} x 128 x 4 B
constexpr int SMEM_QLIST = SMEM_STATS + 2 * 4 * 128 * 4;    // uint16[MAX_TILES]
constexpr int SMEM_TOTAL = SMEM_QLIST + MAX_TILES * 2;
constexpr int SMEM_ALLOC = SMEM_TOTAL + 1024;
static_assert(SMEM_ALLOC <= 232448, "shared memory budget");
constexpr uint32_t TM_S = 0, TM_DP = 128, TM_P = 192, TM_DQ = 224, TM_DV = 320, TM_DK = 416;
constexpr int REGS_CTRL = 48, REGS_WG = 152, REGS_DRAIN = 112;   // 128*48 + 256*152 + 128*112 = 59392 <= 512*128
}  // namespace bwd

struct BwdKernelParams {
  TensorView d_k, d_v;
  const float* lse;
  const float* delta;
  const float* rope_cos;
  const float* rope_sin;
  int64_t rope_stride_b;
  MaskMeta mm;
  int B, H, T, n_t, n_words;
  float scale_log2, scale;
  unsigned long long* trace;   // debug (AKI_MMA_BWD_TRACE=<cta>): clock64 stamps of one CTA
  int trace_cta;
};

#define TRB(slot, g, k) do { if (tracing && (g) < 128) P.trace[((slot) * 128 + (g)) * 8 + (k)] = clock64(); } while (0)

__global__ void __launch_bounds__(bwd::THREADS, 1)
attn_bwd_sm100_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                      const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_do,
                      const __grid_constant__ CUtensorMap map_dq, const BwdKernelParams P) {
  using namespace bwd;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));   // generic pointer to the aligned base
  constexpr int KV_FULL = 0, Q_FULL = 1, Q_EMPTY = Q_FULL + Q_STAGES, DO_FULL = Q_EMPTY + Q_STAGES,
                DO_EMPTY = DO_FULL + DO_STAGES, S_FULL = DO_EMPTY + DO_STAGES /*2*/, S_FREE = S_FULL + 2 /*2*/,
                P_READY = S_FREE + 2, P_FREE = P_READY + 1, DP_FULL = P_FREE + 1, DP_FREE = DP_FULL + 1,
                DS_READY = DP_FREE + 1 /*2*/, DS_FREE = DS_READY + 2, DQ_FULL = DS_FREE + 1, DQ_DRAINED = DQ_FULL + 1,
                N_BARS = DQ_DRAINED + 1;
  __shared__ __align__(8) uint64_t bars[N_BARS];
  __shared__ uint32_t tmem_base_s;
  __shared__ int n_q_s;
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };

  const int tid = threadIdx.x, warp = tid >> 5;
  const int bh = blockIdx.x / P.n_t, kt = blockIdx.x % P.n_t;   // key tiles ascending: heaviest first
  const int b = bh / P.H, h = bh % P.H;
  const int len = meta_len(P.mm, b, P.T);
  const int j0 = kt * BN;
  uint16_t* const qlist = reinterpret_cast<uint16_t*>(smem_gen + SMEM_QLIST);

  if (tid == 0) {
    mbar_init(BAR(KV_FULL), 1);
    for (int i = 0; i < Q_STAGES; ++i) { mbar_init(BAR(Q_FULL + i), 1); mbar_init(BAR(Q_EMPTY + i), 1); }
    for (int i = 0; i < DO_STAGES; ++i) { mbar_init(BAR(DO_FULL + i), 1); mbar_init(BAR(DO_EMPTY + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(S_FULL + i), 1); mbar_init(BAR(S_FREE + i), 128); }
    mbar_init(BAR(P_READY), 128);
    mbar_init(BAR(P_FREE), 129);      // 128 WG2 threads (P^T read) + the commit of the dV MMA that consumed it
    mbar_init(BAR(DP_FULL), 1); mbar_init(BAR(DP_FREE), 128);
    mbar_init(BAR(DS_READY + 0), 128); mbar_init(BAR(DS_READY + 1), 128); mbar_init(BAR(DS_FREE), 1);
    mbar_init(BAR(DQ_FULL), 1); mbar_init(BAR(DQ_DRAINED), 128);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(smem_u32(&tmem_base_s));
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v); tma_prefetch_desc(&map_do);
    tma_prefetch_desc(&map_dq);
  }
  if (warp == 3) {
    // list of query tiles to visit, ascending
    const int n_live = (len + BM - 1) / BM;
    const int lane = tid & 31;
    int n = 0;
    if (j0 < len) {
      if (P.mm.kv_tile_q_mask) {
        const uint32_t* mrow = P.mm.kv_tile_q_mask + ((size_t)b * P.n_t + kt) * P.n_words;
        for (int w0 = 0; w0 < P.n_words; w0 += 32) {
          const int w = w0 + lane;
          uint32_t word = (w < P.n_words) ? mrow[w] : 0u;
          if (w * 32 >= n_live) word = 0u;                                        // keep only live tiles
          else if (w * 32 + 32 > n_live) word &= (1u << (n_live - w * 32)) - 1u;
          const int cnt = __popc(word);
          int incl = cnt;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
          }
          int pos = n + incl - cnt;
          while (word) {
            const int bit = __ffs(word) - 1;
            word &= word - 1;
            qlist[pos++] = (uint16_t)(w * 32 + bit);
          }
          n += __shfl_sync(0xffffffffu, incl, 31);
        }
      } else {
        for (int qt = kt + lane; qt < n_live; qt += 32) qlist[qt - kt] = (uint16_t)qt;
        n = max(0, n_live - kt);
      }
    }
    if (lane == 0) n_q_s = n;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const int n_q = n_q_s;       // query tiles
  const int G = 2 * n_q;       // 64-query sub-tiles

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    setmaxnreg_dec<REGS_CTRL>();
    if (elect_one() && n_q > 0) {
      mbar_arrive_expect_tx(BAR(KV_FULL), 2 * TILE_BYTES);
      for (int a = 0; a < 3; ++a) {
        tma_load_4d(smem_base + SMEM_K + a * ATOM_BYTES, &map_k, BAR(KV_FULL), a * 32, j0, h, b);
        tma_load_4d(smem_base + SMEM_V + a * ATOM_BYTES, &map_v, BAR(KV_FULL), a * 32, j0, h, b);
      }
      for (int it = 0; it < n_q; ++it) {
        const int i0 = (int)qlist[it] * BM;
        const int sq = it % Q_STAGES, sd = it % DO_STAGES;
        mbar_wait(BAR(Q_EMPTY + sq), ((it / Q_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(BAR(Q_FULL + sq), TILE_BYTES);
        for (int a = 0; a < 3; ++a)
          tma_load_4d(smem_base + SMEM_Q + sq * TILE_BYTES + a * ATOM_BYTES, &map_q, BAR(Q_FULL + sq), a * 32, i0, h, b);
        mbar_wait(BAR(DO_EMPTY + sd), ((it / DO_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(BAR(DO_FULL + sd), TILE_BYTES);
        for (int a = 0; a < 3; ++a)
          tma_load_4d(smem_base + SMEM_DO + sd * TILE_BYTES + a * ATOM_BYTES, &map_do, BAR(DO_FULL + sd), a * 32, i0, h, b);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp runs the loop (addresses / descriptors stay in uniform registers); one lane issues.
    setmaxnreg_dec<REGS_CTRL>();
    const bool leader = elect_one();
    const bool tracing = P.trace && (int)blockIdx.x == P.trace_cta && leader;
    if (n_q > 0) {
      constexpr uint32_t IDESC_N64 = umma_idesc_bf16(128, 64, 0, 0);           // S^T, dP^T
      constexpr uint32_t IDESC_N96_BMN = umma_idesc_bf16(128, 96, 0, 1);      // dV (TS), dK (SS)
      constexpr uint32_t IDESC_N96_AMN_BMN = umma_idesc_bf16(128, 96, 1, 1);  // dQ
      const uint64_t DESC_KMAJ = umma_smem_desc(0, 16, 512, UMMA_SW64);
      const uint64_t DESC_MNMAJ = umma_smem_desc(0, ATOM_BYTES, 512, UMMA_SW64);
      const uint32_t k_lo = (smem_base + SMEM_K) >> 4, v_lo = (smem_base + SMEM_V) >> 4,
                     q_lo = (smem_base + SMEM_Q) >> 4, do_lo = (smem_base + SMEM_DO) >> 4,
                     ds_lo = (smem_base + SMEM_DS) >> 4;
      auto kmaj = [&](uint32_t lo, int k) { return DESC_KMAJ | (uint64_t)(lo + (((k >> 1) * ATOM_BYTES + (k & 1) * 32) >> 4)); };
      auto mnmaj = [&](uint32_t lo, int k) { return DESC_MNMAJ | (uint64_t)(lo + k * 64); };
      auto q_of = [&](int g) { return q_lo + ((g >> 1) % Q_STAGES) * (TILE_BYTES >> 4) + (g & 1) * (SUB_BYTES >> 4); };
      auto do_of = [&](int g) { return do_lo + ((g >> 1) % DO_STAGES) * (TILE_BYTES >> 4) + (g & 1) * (SUB_BYTES >> 4); };
      auto issue_s = [&](int g) {          // S^T_g = K Q_g^T  -> S buffer g&1
        if ((g & 1) == 0) mbar_wait(BAR(Q_FULL + (g >> 1) % Q_STAGES), ((g >> 1) / Q_STAGES) & 1);
        if (g >= 2) mbar_wait(BAR(S_FREE + (g & 1)), ((g >> 1) - 1) & 1);     // WG1 has read S^T_{g-2}
        tc_fence_after();
        if (leader) {
          const uint32_t qa = q_of(g);
#pragma unroll
          for (int k = 0; k < 6; ++k) umma_ss(tmem + TM_S + 64 * (g & 1), kmaj(k_lo, k), kmaj(qa, k), IDESC_N64, k > 0);
          umma_commit(BAR(S_FULL + (g & 1)));
        }
        __syncwarp();
      };
      auto issue_dp = [&](int g) {         // dP^T_g = V dO_g^T
        if ((g & 1) == 0) mbar_wait(BAR(DO_FULL + (g >> 1) % DO_STAGES), ((g >> 1) / DO_STAGES) & 1);
        if (g >= 1) mbar_wait(BAR(DP_FREE), (g - 1) & 1);                     // WG2 has read dP^T_{g-1}
        tc_fence_after();
        if (leader) {
          const uint32_t da = do_of(g);
#pragma unroll
          for (int k = 0; k < 6; ++k) umma_ss(tmem + TM_DP, kmaj(v_lo, k), kmaj(da, k), IDESC_N64, k > 0);
          umma_commit(BAR(DP_FULL));
        }
        __syncwarp();
      };
      auto issue_dv = [&](int g) {         // dV += P^T_g dO_g
        mbar_wait(BAR(P_READY), g & 1);
        tc_fence_after();
        if (leader) {
          const uint32_t da = do_of(g);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ts(tmem + TM_DV, tmem + TM_P + 8 * k, mnmaj(da, k), IDESC_N96_BMN, (g > 0 || k > 0));
          umma_commit(BAR(P_FREE));                                              // one of the 129 arrivals
          if (g & 1) umma_commit(BAR(DO_EMPTY + (g >> 1) % DO_STAGES));          // dO tile fully consumed
        }
        __syncwarp();
      };
      auto issue_dk_dq = [&](int g) {      // dK += dS^T_g Q_g ; after the second sub-tile: dQ_tile = dS K
        // one barrier per sub-tile parity: WG2 may run a whole sub-tile ahead of this consumer
        mbar_wait(BAR(DS_READY + (g & 1)), (g >> 1) & 1);
        const int tile = g >> 1;
        if ((g & 1) && tile >= 1) mbar_wait(BAR(DQ_DRAINED), (tile - 1) & 1);   // dQ columns free again
        tc_fence_after();
        if (leader) {
          const uint32_t qa = q_of(g);
          const uint32_t dsa = ds_lo + (g & 1) * (2 * ATOM_BYTES >> 4);          // atoms 2u, 2u+1 hold queries 64u..64u+63
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tmem + TM_DK, kmaj(dsa, k), mnmaj(qa, k), IDESC_N96_BMN, (g > 0 || k > 0));
          if (g & 1) {
#pragma unroll
            for (int k = 0; k < 8; ++k) umma_ss(tmem + TM_DQ, mnmaj(ds_lo, k), mnmaj(k_lo, k), IDESC_N96_AMN_BMN, k > 0);
            umma_commit(BAR(DQ_FULL));
            umma_commit(BAR(DS_FREE));
            umma_commit(BAR(Q_EMPTY + tile % Q_STAGES));
          }
        }
        __syncwarp();
      };
      mbar_wait(BAR(KV_FULL), 0);
      issue_s(0);
      issue_dp(0);
      // steady state per sub-tile g: dV_g | S_{g+1} | dP_{g+1} | dK_{g-1} (+dQ) -- the dS consumer runs one
      // sub-tile behind so that WG2 has a whole step of slack
      for (int g = 0; g < G; ++g) {
        TRB(0, g, 0);
        if (g + 1 < G) issue_s(g + 1);
        TRB(0, g, 1);
        issue_dv(g);
        TRB(0, g, 2);
        if (g + 1 < G) issue_dp(g + 1);
        TRB(0, g, 3);
        if (g >= 1) issue_dk_dq(g - 1);
        TRB(0, g, 4);
      }
      issue_dk_dq(G - 1);
    }
  } else if (warp < 4) {
    setmaxnreg_dec<REGS_CTRL>();
  } else if (warp < 8) {
    // ------------------------------------------------------------------ WG1: P^T = exp2(S^T * c - LSE)
    setmaxnreg_inc<REGS_WG>();
    const int r = tid - 128;                  // key row within the tile == TMEM lane
    const int j = j0 + r;                     // key index
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    float* const stats_gen = reinterpret_cast<float*>(smem_gen + SMEM_STATS);   // [stage][lse2 | delta | lo | hi][128]
    const size_t bhT = ((size_t)b * P.H + h) * P.T;
    bool k_valid = (j < len), k_mutual = (j < len);
    if (j < len && P.mm.vbits) k_valid = (P.mm.vbits[(size_t)b * P.mm.bits_pitch + (j >> 5)] >> (j & 31)) & 1u;
    if (j < len && P.mm.mbits) k_mutual = (P.mm.mbits[(size_t)b * P.mm.bits_pitch + (j >> 5)] >> (j & 31)) & 1u;
    const bool warp_keys_valid = __all_sync(0xffffffffu, k_valid);
    const bool tracing = P.trace && (int)blockIdx.x == P.trace_cta && r == 0;

    float pre_lse = INFINITY; int pre_lo = 0, pre_hi = 0;
    auto prefetch = [&](int it) {
      const int i = (int)qlist[it] * BM + r;
      pre_lse = (i < len) ? __ldg(P.lse + bhT + i) * 1.4426950408889634f : INFINITY;
      pre_lo = pre_hi = 0;
      if (i < len && P.mm.row_lo) {
        pre_lo = __ldg(P.mm.row_lo + (size_t)b * P.mm.meta_pitch + i);
        pre_hi = __ldg(P.mm.row_hi + (size_t)b * P.mm.meta_pitch + i);
      }
    };
    if (n_q > 0) prefetch(0);
    for (int g = 0; g < G; ++g) {
      const int it = g >> 1, u = g & 1;
      float* st = stats_gen + (it & 1) * 512;
      if (u == 0) {            // publish this tile's row statistics (WG1 owns lse2 / lo / hi), fetch the next tile's
        st[r] = pre_lse;
        reinterpret_cast<int*>(st)[256 + r] = pre_lo;
        reinterpret_cast<int*>(st)[384 + r] = pre_hi;
        if (it + 1 < n_q) prefetch(it + 1);
        named_bar_sync(1, 128);
      }
      const int qt = (int)qlist[it], i0 = qt * BM + 64 * u;
      TRB(1, g, 0);
      mbar_wait(BAR(S_FULL + u), (g >> 1) & 1);
      TRB(1, g, 1);
      tc_fence_after();
      uint32_t sraw[64];
      tmem_ld_x32(tmem + TM_S + lane_base + 64 * u, sraw);
      tmem_ld_x32(tmem + TM_S + lane_base + 64 * u + 32, sraw + 32);
      const bool full = (qt > kt) && (qt * BM + BM <= len) && warp_keys_valid;
      const float4* lse4 = reinterpret_cast<const float4*>(st + 64 * u);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(BAR(S_FREE + u));            // S^T_g is in registers: its buffer may take S^T_{g+2}
      TRB(1, g, 2);
      uint32_t pk[32];
      if (full) {
#pragma unroll
        for (int c4 = 0; c4 < 16; ++c4) {
          const float4 v = lse4[c4];
          const float p0 = ex2_approx(fmaf(__uint_as_float(sraw[4 * c4 + 0]), P.scale_log2, -v.x));
          const float p1 = ex2_approx(fmaf(__uint_as_float(sraw[4 * c4 + 1]), P.scale_log2, -v.y));
          const float p2 = ex2_approx(fmaf(__uint_as_float(sraw[4 * c4 + 2]), P.scale_log2, -v.z));
          const float p3 = ex2_approx(fmaf(__uint_as_float(sraw[4 * c4 + 3]), P.scale_log2, -v.w));
          pk[2 * c4] = pack_bf16x2(p0, p1);
          pk[2 * c4 + 1] = pack_bf16x2(p2, p3);
        }
      } else {
        const int4* lo4 = reinterpret_cast<const int4*>(st + 256 + 64 * u);
        const int4* hi4 = reinterpret_cast<const int4*>(st + 384 + 64 * u);
#pragma unroll
        for (int c4 = 0; c4 < 16; ++c4) {
          const int4 a4 = lo4[c4], e4 = hi4[c4];
          const float4 l4 = lse4[c4];
          const float lse2[4] = {l4.x, l4.y, l4.z, l4.w};
          const int lo[4] = {a4.x, a4.y, a4.z, a4.w}, hi[4] = {e4.x, e4.y, e4.z, e4.w};
          float pv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = i0 + 4 * c4 + e;
            const bool ok = (i < len) && ((j <= i && k_valid) || (j >= lo[e] && j < hi[e] && k_mutual));
            const float val = ex2_approx(fmaf(__uint_as_float(sraw[4 * c4 + e]), P.scale_log2, -lse2[e]));
            pv[e] = ok ? val : 0.f;
          }
          pk[2 * c4] = pack_bf16x2(pv[0], pv[1]);
          pk[2 * c4 + 1] = pack_bf16x2(pv[2], pv[3]);
        }
      }
      TRB(1, g, 3);
      if (g >= 1) mbar_wait(BAR(P_FREE), (g - 1) & 1);   // dV_{g-1} and WG2 are done with P^T_{g-1}
      TRB(1, g, 4);
      tc_fence_after();
      tmem_st_x32(tmem + TM_P + lane_base, pk);
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(BAR(P_READY));
      TRB(1, g, 5);
    }
  } else if (warp < 12) {
    // ------------------------------------------------------------------ WG2: dS^T = P^T o (dP^T - delta) -> smem
    setmaxnreg_inc<REGS_WG>();
    const int r = tid - 256;
    const int j = j0 + r;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    float* const stats_gen = reinterpret_cast<float*>(smem_gen + SMEM_STATS);
    const size_t bhT = ((size_t)b * P.H + h) * P.T;
    const bool tracing = P.trace && (int)blockIdx.x == P.trace_cta && r == 0;
    float pre_delta = 0.f;
    auto prefetch = [&](int it) {
      const int i = (int)qlist[it] * BM + r;
      pre_delta = (i < len) ? __ldg(P.delta + bhT + i) : 0.f;
    };
    if (n_q > 0) prefetch(0);
    for (int g = 0; g < G; ++g) {
      const int it = g >> 1, u = g & 1;
      float* st = stats_gen + (it & 1) * 512 + 128;      // delta
      if (u == 0) {
        st[r] = pre_delta;
        if (it + 1 < n_q) prefetch(it + 1);
        named_bar_sync(2, 128);
      }
      TRB(2, g, 0);
      mbar_wait(BAR(P_READY), g & 1);
      tc_fence_after();
      TRB(2, g, 1);
      uint32_t pk[32];
      tmem_ld_x32(tmem + TM_P + lane_base, pk);
      mbar_wait(BAR(DP_FULL), g & 1);
      TRB(2, g, 2);
      tc_fence_after();
      uint32_t draw[64];
      tmem_ld_x32(tmem + TM_DP + lane_base, draw);
      tmem_ld_x32(tmem + TM_DP + lane_base + 32, draw + 32);
      const float4* dl4 = reinterpret_cast<const float4*>(st + 64 * u);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(BAR(P_FREE));
      mbar_arrive(BAR(DP_FREE));
      TRB(2, g, 3);
      uint32_t w[32];
#pragma unroll
      for (int c4 = 0; c4 < 16; ++c4) {
        const float4 d = dl4[c4];
        const float p0 = __uint_as_float(pk[2 * c4] << 16), p1 = __uint_as_float(pk[2 * c4] & 0xffff0000u);
        const float p2 = __uint_as_float(pk[2 * c4 + 1] << 16), p3 = __uint_as_float(pk[2 * c4 + 1] & 0xffff0000u);
        w[2 * c4] = pack_bf16x2(p0 * (__uint_as_float(draw[4 * c4]) - d.x), p1 * (__uint_as_float(draw[4 * c4 + 1]) - d.y));
        w[2 * c4 + 1] = pack_bf16x2(p2 * (__uint_as_float(draw[4 * c4 + 2]) - d.z), p3 * (__uint_as_float(draw[4 * c4 + 3]) - d.w));
      }
      // the dS^T buffer is read by dK (per sub-tile) and dQ (per tile): wait for the previous tile's MMAs
      TRB(2, g, 4);
      if (it >= 1) mbar_wait(BAR(DS_FREE), (it - 1) & 1);
      TRB(2, g, 5);
      const uint32_t ds_base = smem_base + SMEM_DS;
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {   // 8 query columns -> one 16-byte chunk of the dS^T row
        const int col = 64 * u + 8 * c8;
        const uint32_t addr = ds_base + (col >> 5) * ATOM_BYTES + sw64_offset(r, (col & 31) >> 3);
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(w[4 * c8]), "r"(w[4 * c8 + 1]),
                     "r"(w[4 * c8 + 2]), "r"(w[4 * c8 + 3]) : "memory");
      }
      fence_proxy_async_smem();
      mbar_arrive(BAR(DS_READY + u));
      TRB(2, g, 6);
    }
    if (n_q > 0) {       // the last DQ_FULL commit covers every dK / dV MMA
      mbar_wait(BAR(DQ_FULL), (n_q - 1) & 1);
      tc_fence_after();
    }
    // ---- epilogue (WG2): dV -> bf16 -> global
    {
      const bool store_row = (j < P.T);
      __nv_bfloat16* dvrow = P.d_v.row(b, store_row ? j : 0, h);
      if (n_q > 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          uint32_t acc[32];
          tmem_ld_x32(tmem + TM_DV + lane_base + 32 * c, acc);
          tmem_wait_ld();
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            uint4 v4;
            v4.x = pack_bf16x2(__uint_as_float(acc[8 * x]), __uint_as_float(acc[8 * x + 1]));
            v4.y = pack_bf16x2(__uint_as_float(acc[8 * x + 2]), __uint_as_float(acc[8 * x + 3]));
            v4.z = pack_bf16x2(__uint_as_float(acc[8 * x + 4]), __uint_as_float(acc[8 * x + 5]));
            v4.w = pack_bf16x2(__uint_as_float(acc[8 * x + 6]), __uint_as_float(acc[8 * x + 7]));
            if (store_row) *reinterpret_cast<uint4*>(dvrow + 32 * c + 8 * x) = v4;
          }
        }
      } else if (store_row) {
        const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int x = 0; x < 12; ++x) *reinterpret_cast<uint4*>(dvrow + 8 * x) = z;
      }
    }
  } else {
    // ------------------------------------------------------------------ WG3: dQ drain, then the dK epilogue
    setmaxnreg_dec<REGS_DRAIN>();
    const int r = tid - 384;                  // QUERY row r of the tile while draining; KEY row r in the epilogue
    const bool tracing = P.trace && (int)blockIdx.x == P.trace_cta && r == 0;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    for (int it = 0; it < n_q; ++it) {
      const int i0 = (int)qlist[it] * BM;
      TRB(3, it, 0);
      mbar_wait(BAR(DQ_FULL), it & 1);
      tc_fence_after();
      TRB(3, it, 1);
      uint32_t dq[96];
      tmem_ld_x32(tmem + TM_DQ + lane_base, dq);
      tmem_ld_x32(tmem + TM_DQ + lane_base + 32, dq + 32);
      tmem_ld_x32(tmem + TM_DQ + lane_base + 64, dq + 64);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(BAR(DQ_DRAINED));
      TRB(3, it, 2);
      // three [128][32 x fp32] SWIZZLE_128B atoms, each through its own staging buffer (DQ_BUFS reductions in flight)
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const uint32_t abase = smem_base + SMEM_DQ + ((it * 3 + a) % DQ_BUFS) * DQ_ATOM_BYTES;
        if (r == 0) tma_store_wait_read<DQ_BUFS - 1>();   // the reduction that last used this buffer has read it
        named_bar_sync(3, 128);
#pragma unroll
        for (int x = 0; x < 8; ++x) {
          const uint32_t addr = abase + r * 128 + ((x ^ (r & 7)) << 4);
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(dq[32 * a + 4 * x]),
                       "r"(dq[32 * a + 4 * x + 1]), "r"(dq[32 * a + 4 * x + 2]), "r"(dq[32 * a + 4 * x + 3]) : "memory");
        }
        fence_proxy_async_smem();
        named_bar_sync(4, 128);
        if (r == 0) {
          tma_reduce_add_4d(&map_dq, abase, 32 * a, i0, h, b);
          tma_store_commit();
        }
      }
      TRB(3, it, 3);
    }
    if (r == 0) tma_store_wait<0>();   // all dQ reductions have landed before the CTA retires its smem
    // ---- epilogue (WG3): dK (x scale, inverse RoPE) -> bf16 -> global.  The last DQ_FULL covers every dK MMA.
    {
      const int j = j0 + r;
      const bool store_row = (j < P.T);
      const int js = store_row ? j : 0;
      __nv_bfloat16* dkrow = P.d_k.row(b, js, h);
      if (n_q > 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {     // 16 columns d and their RoPE partners d+48
          uint32_t lo[16], hi[16];
          tmem_ld_x16(tmem + TM_DK + lane_base + 16 * c, lo);
          tmem_ld_x16(tmem + TM_DK + lane_base + 48 + 16 * c, hi);
          tmem_wait_ld();
          float flo[16], fhi[16];
#pragma unroll
          for (int x = 0; x < 16; ++x) {
            float a = __uint_as_float(lo[x]) * P.scale, e = __uint_as_float(hi[x]) * P.scale;
            if (P.rope_cos) {   // g = R^T g'
              const float cs = __ldg(P.rope_cos + (size_t)b * P.rope_stride_b + (size_t)js * 48 + 16 * c + x);
              const float sn = __ldg(P.rope_sin + (size_t)b * P.rope_stride_b + (size_t)js * 48 + 16 * c + x);
              const float a2 = a * cs + e * sn, e2 = e * cs - a * sn;
              a = a2; e = e2;
            }
            flo[x] = a; fhi[x] = e;
          }
#pragma unroll
          for (int x = 0; x < 2; ++x) {
            uint4 u4, w4;
            u4.x = pack_bf16x2(flo[8 * x], flo[8 * x + 1]); u4.y = pack_bf16x2(flo[8 * x + 2], flo[8 * x + 3]);
            u4.z = pack_bf16x2(flo[8 * x + 4], flo[8 * x + 5]); u4.w = pack_bf16x2(flo[8 * x + 6], flo[8 * x + 7]);
            w4.x = pack_bf16x2(fhi[8 * x], fhi[8 * x + 1]); w4.y = pack_bf16x2(fhi[8 * x + 2], fhi[8 * x + 3]);
            w4.z = pack_bf16x2(fhi[8 * x + 4], fhi[8 * x + 5]); w4.w = pack_bf16x2(fhi[8 * x + 6], fhi[8 * x + 7]);
            if (store_row) {
              *reinterpret_cast<uint4*>(dkrow + 16 * c + 8 * x) = u4;
              *reinterpret_cast<uint4*>(dkrow + 48 + 16 * c + 8 * x) = w4;
            }
          }
        }
      } else if (store_row) {
        const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int x = 0; x < 12; ++x) *reinterpret_cast<uint4*>(dkrow + 8 * x) = z;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem);
}

}  // namespace aki

using namespace aki;

extern "C" int aki_mma_attn_bwd(const AkiMmaAttnBwdParams* p, aki_stream_t stream) {
  AKI_REQUIRE(p, AKI_ERR_NULL);
  const AkiMmaAttnParams& f = p->fwd;
  int rc = check_attn_params(f);
  if (rc) return rc;
  if ((rc = check_tensor(p->d_o)) || (rc = check_tensor(p->d_q)) || (rc = check_tensor(p->d_k)) ||
      (rc = check_tensor(p->d_v)))
    return rc;
  AKI_REQUIRE(f.lse && p->workspace, AKI_ERR_NULL);
  AKI_REQUIRE(p->workspace_bytes >= aki_mma_attn_bwd_workspace_bytes(f.B, f.H, f.T, f.D), AKI_ERR_BAD_SHAPE);
  AKI_REQUIRE((reinterpret_cast<uintptr_t>(p->workspace) & 255u) == 0, AKI_ERR_MISALIGNED);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BwdWorkspace w = carve_bwd_workspace(p->workspace, f.B, f.H, f.T, f.D);
  if ((rc = launch_bwd_preprocess(*p, w, st))) return rc;
  if (cudaMemsetAsync(w.dq_accum, 0, (size_t)f.B * f.H * f.T * f.D * sizeof(float), st) != cudaSuccess) {
    set_last_cuda_error(cudaGetErrorString(cudaGetLastError()));
    return AKI_ERR_CUDA;
  }
  AkiMmaTensor4 qrot{w.q_rot, (int64_t)f.H * f.T * f.D, (int64_t)f.D, (int64_t)f.T * f.D};
  CUtensorMap mq, mk, mv, mdo, mdq;
  if ((rc = make_tile_map(&mq, qrot, f.B, f.H, f.T, bwd::BM))) return rc;
  if ((rc = make_tile_map(&mk, f.k, f.B, f.H, f.T, bwd::BN))) return rc;
  if ((rc = make_tile_map(&mv, f.v, f.B, f.H, f.T, bwd::BN))) return rc;
  if ((rc = make_tile_map(&mdo, p->d_o, f.B, f.H, f.T, bwd::BM))) return rc;
  if ((rc = make_dq_accum_map(&mdq, w.dq_accum, f.B, f.H, f.T))) return rc;
  BwdKernelParams kp;
  kp.d_k = view_of(p->d_k); kp.d_v = view_of(p->d_v);
  kp.lse = f.lse; kp.delta = w.delta;
  kp.rope_cos = f.rope_cos; kp.rope_sin = f.rope_sin; kp.rope_stride_b = f.rope_stride_b;
  kp.mm = mask_meta_from(f);
  kp.B = f.B; kp.H = f.H; kp.T = f.T;
  kp.n_t = (f.T + bwd::BN - 1) / bwd::BN;
  kp.n_words = (kp.n_t + 31) / 32;
  AKI_REQUIRE(kp.n_t <= bwd::MAX_TILES, AKI_ERR_UNSUPPORTED);
  kp.scale = f.scale;
  kp.scale_log2 = f.scale * 1.4426950408889634f;
  kp.trace = nullptr; kp.trace_cta = -1;
  const char* trace_env = getenv("AKI_MMA_BWD_TRACE");   // debug only; synchronises
  const size_t trace_bytes = 4 * 128 * 8 * sizeof(unsigned long long);
  if (trace_env) { kp.trace_cta = atoi(trace_env); cudaMalloc(&kp.trace, trace_bytes); cudaMemset(kp.trace, 0, trace_bytes); }
  const long long grid = (long long)kp.n_t * f.H * f.B;
  AKI_REQUIRE(grid > 0 && grid < (1ll << 31), AKI_ERR_BAD_SHAPE);
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(attn_bwd_sm100_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd::SMEM_ALLOC) !=
        cudaSuccess) {
      set_last_cuda_error(cudaGetErrorString(cudaGetLastError()));
      return AKI_ERR_CUDA;
    }
    attr_done = true;
  }
  attn_bwd_sm100_kernel<<<(unsigned)grid, bwd::THREADS, bwd::SMEM_ALLOC, st>>>(mq, mk, mv, mdo, mdq, kp);
  if (trace_env) {
    cudaDeviceSynchronize();
    static unsigned long long host[4 * 128 * 8];
    cudaMemcpy(host, kp.trace, trace_bytes, cudaMemcpyDeviceToHost);
    cudaFree(kp.trace);
    unsigned long long t0 = ~0ull;
    for (size_t i = 0; i < 4 * 128 * 8; ++i) if (host[i] && host[i] < t0) t0 = host[i];
    const char* names[4] = {"mma", "wg1_P", "wg2_dS", "drain"};
    for (int slot = 0; slot < 4; ++slot)
      for (int g = 0; g < 128; ++g) {
        if (!host[(slot * 128 + g) * 8]) continue;
        fprintf(stderr, "TRACE %s g=%d:", names[slot], g);
        for (int k = 0; k < 7; ++k) fprintf(stderr, " %llu", host[(slot * 128 + g) * 8 + k] ? host[(slot * 128 + g) * 8 + k] - t0 : 0ull);
        fprintf(stderr, "\n");
      }
  }
  if ((rc = check_launch())) return rc;
  return launch_dq_finalize(*p, w, f.scale, st);   // dS^T is kept unscaled inside the kernel
}
