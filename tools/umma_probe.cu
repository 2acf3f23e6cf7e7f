// Bring-up probe for the five tcgen05.mma operand forms the attention kernels use (run on the B200 box):
//   mode 0: dump a TMA SWIZZLE_64B tile and check the swizzle formula
//   mode 1: SS  A K-major  (TMA tile)        x B K-major  (TMA tile)   -> S  = Q K^T      (128x128, K=96)
//   mode 2: TS  A in TMEM  (bf16 pairs)      x B MN-major (TMA tile)   -> O  = P V        (128x96,  K=128)
//   mode 3: SS  A MN-major (thread-written)  x B MN-major (TMA tile)   -> dQ = dS K       (128x96,  K=128)
//   mode 4: SS  A K-major  (thread-written)  x B MN-major (TMA tile)   -> dK = dS^T Q     (128x96,  K=128)
//   mode 5: mode 1 plus a 7th k-step in the SWIZZLE_NONE K-major form: A = thread-written "ones" rows [1,1,1,0..],
//           B = TMA-loaded [128][8] bf16 row statistics (box 8x128, no swizzle), second 16-byte K chunk of both
//           operands = one shared zero region reached through LBO  -> S'[m][n] = S[m][n] + sum_k aug[n][k]
//           (this is how the backward folds -LSE / -delta into the S^T and dP^T MMAs).
//           mode 6: same with SBO = 0 for the ones operand (all row groups alias one core matrix);
//           mode 7: same with LBO = 0 for the statistics operand (second K chunk aliases the first).
// Descriptor strides (LBO/SBO) and the per-k-step start-address advance are arguments so that one GPU
// call can sweep candidates:   umma_probe <mode> <a_lbo> <a_sbo> <a_kadv> <b_lbo> <b_sbo> <b_kadv>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../aki_b200/csrc/sm100_ptx.cuh"

using namespace aki;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

struct ProbeArgs {
  int mode;
  uint32_t a_lbo, a_sbo, a_kadv, b_lbo, b_sbo, b_kadv;
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
             const __grid_constant__ CUtensorMap mapAug, const __nv_bfloat16* __restrict__ X,   // [128][128] bf16 (modes 2,3,4)
             float* __restrict__ D, uint8_t* __restrict__ dump, ProbeArgs pa) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // 32 KB region
  uint8_t* sB = smem + 32768;         // 24 KB
  uint8_t* sAug = smem + 57344;       // 2 KB  [16 groups][8 rows][16 B]
  uint8_t* sOnes = smem + 59392;      // 2 KB
  uint8_t* sZero = smem + 61440;      // 2 KB
  const bool aug = pa.mode >= 5;
  const int mode_in = pa.mode;
  if (aug) pa.mode = 1;
  __shared__ __align__(8) uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(smem_u32(&bar_load), 1);
    mbar_init(smem_u32(&bar_mma), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_base_s));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (aug) {
    for (int r = tid; r < 128; r += 128) {
      const uint4 one = make_uint4(0x3f803f80u, 0x00003f80u, 0u, 0u);   // bf16 [1,1,1,0,0,0,0,0]
      *reinterpret_cast<uint4*>(sOnes + (r >> 3) * 128 + (r & 7) * 16) = one;
      *reinterpret_cast<uint4*>(sZero + r * 16) = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
  }
  if (tid == 0) {
    uint32_t bytes = 24576 * ((pa.mode <= 1) ? 2 : 1) + (aug ? 2048 : 0);
    mbar_arrive_expect_tx(smem_u32(&bar_load), bytes);
    if (aug) tma_load_4d(smem_u32(sAug), &mapAug, smem_u32(&bar_load), 0, 0, 0, 0);
    if (pa.mode <= 1)
      for (int a = 0; a < 3; ++a) tma_load_4d(smem_u32(sA + a * 8192), &mapA, smem_u32(&bar_load), a * 32, 0, 0, 0);
    for (int a = 0; a < 3; ++a) tma_load_4d(smem_u32(sB + a * 8192), &mapB, smem_u32(&bar_load), a * 32, 0, 0, 0);
  }
  // thread-written A operand: X[r][c] -> 4 atoms [128 rows][64 B], SW64
  if (pa.mode == 3 || pa.mode == 4) {
    for (int idx = tid; idx < 128 * 16; idx += 128) {
      int r = idx >> 4, ch = idx & 15;          // 16-byte chunk ch of row r (8 elements)
      uint4 v = *reinterpret_cast<const uint4*>(X + r * 128 + ch * 8);
      int atom = ch >> 2, c = ch & 3;
      *reinterpret_cast<uint4*>(sA + atom * 8192 + sw64_offset(r, c)) = v;
    }
    fence_proxy_async_smem();
  }
  if (pa.mode == 2) {
    // P row r -> TMEM lane r, 64 columns of packed bf16 pairs at column offset 256
    const uint32_t* xr = reinterpret_cast<const uint32_t*>(X + (warp * 32 + lane) * 128);
    uint32_t regs[32];
    for (int h = 0; h < 2; ++h) {
      for (int i = 0; i < 32; ++i) regs[i] = xr[h * 32 + i];
      tmem_st_x32(tmem + ((warp * 32u) << 16) + 256 + h * 32, regs);
    }
    tmem_wait_st();
    tc_fence_before();
  }
  __syncthreads();
  mbar_wait(smem_u32(&bar_load), 0);
  tc_fence_after();

  if (pa.mode == 0) {
    for (int i = tid; i < 24576 / 16; i += 128)
      reinterpret_cast<uint4*>(dump)[i] = reinterpret_cast<const uint4*>(sB)[i];
  }

  if (warp == 0 && pa.mode >= 1) {
    if (elect_one()) {
      if (pa.mode == 1) {
        const uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
        for (int k = 0; k < 6; ++k) {
          uint32_t aoff = (k >> 1) * 8192 + (k & 1) * pa.a_kadv;
          uint32_t boff = (k >> 1) * 8192 + (k & 1) * pa.b_kadv;
          umma_ss(tmem, umma_smem_desc(smem_u32(sA) + aoff, pa.a_lbo, pa.a_sbo, UMMA_SW64),
                  umma_smem_desc(smem_u32(sB) + boff, pa.b_lbo, pa.b_sbo, UMMA_SW64), idesc, k > 0);
        }
        if (aug) {
          const uint32_t ones_sbo = (mode_in == 6) ? 0u : 128u;
          const uint32_t aug_lbo = (mode_in == 7) ? 0u : (uint32_t)(sZero - sAug);
          umma_ss(tmem, umma_smem_desc(smem_u32(sOnes), (uint32_t)(sZero - sOnes), ones_sbo, UMMA_SW_NONE),
                  umma_smem_desc(smem_u32(sAug), aug_lbo, 128, UMMA_SW_NONE), idesc, 1);
        }
      } else if (pa.mode == 2) {
        const uint32_t idesc = umma_idesc_bf16(128, 96, 0, 1);
        for (int k = 0; k < 8; ++k)
          umma_ts(tmem, tmem + 256 + k * 8, umma_smem_desc(smem_u32(sB) + k * pa.b_kadv, pa.b_lbo, pa.b_sbo, UMMA_SW64),
                  idesc, k > 0);
      } else if (pa.mode == 3) {
        const uint32_t idesc = umma_idesc_bf16(128, 96, 1, 1);
        for (int k = 0; k < 8; ++k)
          umma_ss(tmem, umma_smem_desc(smem_u32(sA) + k * pa.a_kadv, pa.a_lbo, pa.a_sbo, UMMA_SW64),
                  umma_smem_desc(smem_u32(sB) + k * pa.b_kadv, pa.b_lbo, pa.b_sbo, UMMA_SW64), idesc, k > 0);
      } else if (pa.mode == 4) {
        const uint32_t idesc = umma_idesc_bf16(128, 96, 0, 1);
        for (int k = 0; k < 8; ++k) {
          uint32_t aoff = (k >> 1) * 8192 + (k & 1) * pa.a_kadv;
          umma_ss(tmem, umma_smem_desc(smem_u32(sA) + aoff, pa.a_lbo, pa.a_sbo, UMMA_SW64),
                  umma_smem_desc(smem_u32(sB) + k * pa.b_kadv, pa.b_lbo, pa.b_sbo, UMMA_SW64), idesc, k > 0);
        }
      }
      umma_commit(smem_u32(&bar_mma));
    }
    __syncwarp();
  }
  if (pa.mode >= 1) {
    mbar_wait(smem_u32(&bar_mma), 0);
    tc_fence_after();
    const int ncol = (pa.mode == 1) ? 128 : 96;
    uint32_t regs[32];
    for (int c0 = 0; c0 < ncol; c0 += 32) {
      tmem_ld_x32(tmem + ((warp * 32u) << 16) + c0, regs);
      tmem_wait_ld();
      for (int i = 0; i < 32; ++i) D[(warp * 32 + lane) * ncol + c0 + i] = __uint_as_float(regs[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main(int argc, char** argv) {
  ProbeArgs pa{};
  pa.mode = argc > 1 ? atoi(argv[1]) : 1;
  pa.a_lbo = argc > 2 ? atoi(argv[2]) : 16;
  pa.a_sbo = argc > 3 ? atoi(argv[3]) : 512;
  pa.a_kadv = argc > 4 ? atoi(argv[4]) : 32;
  pa.b_lbo = argc > 5 ? atoi(argv[5]) : 16;
  pa.b_sbo = argc > 6 ? atoi(argv[6]) : 512;
  pa.b_kadv = argc > 7 ? atoi(argv[7]) : 32;

  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  EncodeFn encode = reinterpret_cast<EncodeFn>(fn);

  std::vector<float> A(128 * 96), Bm(128 * 96), X(128 * 128);
  srand(1234);
  auto rnd = []() { return (rand() % 2001 - 1000) / 1000.0f; };
  for (auto& x : A) x = bf(rnd());
  for (auto& x : Bm) x = bf(rnd());
  for (auto& x : X) x = bf(rnd());
  std::vector<__nv_bfloat16> hA(A.size()), hB(Bm.size()), hX(X.size());
  for (size_t i = 0; i < A.size(); ++i) hA[i] = __float2bfloat16(A[i]);
  for (size_t i = 0; i < Bm.size(); ++i) hB[i] = __float2bfloat16(Bm[i]);
  for (size_t i = 0; i < X.size(); ++i) hX[i] = __float2bfloat16(X[i]);
  __nv_bfloat16 *dA, *dB, *dX;
  float* dD;
  uint8_t* ddump;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dX, hX.size() * 2));
  CK(cudaMalloc(&dD, 128 * 128 * 4)); CK(cudaMalloc(&ddump, 24576));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dX, hX.data(), hX.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, 128 * 128 * 4));

  auto make_map = [&](void* ptr) {
    CUtensorMap m;
    cuuint64_t dims[4] = {96, 128, 1, 1};
    cuuint64_t strides[3] = {192, 192 * 128, 192 * 128};
    cuuint32_t box[4] = {32, 128, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(3); }
    return m;
  };
  CUtensorMap mA = make_map(dA), mB = make_map(dB);
  std::vector<float> AUG(128 * 8, 0.f);
  std::vector<__nv_bfloat16> hAug(128 * 8);
  for (int n = 0; n < 128; ++n) {
    AUG[n * 8 + 0] = bf(100.f * rnd()); AUG[n * 8 + 1] = bf(rnd()); AUG[n * 8 + 2] = bf(0.01f * rnd());
    for (int k = 0; k < 8; ++k) hAug[n * 8 + k] = __float2bfloat16(AUG[n * 8 + k]);
  }
  __nv_bfloat16* dAug;
  CK(cudaMalloc(&dAug, hAug.size() * 2));
  CK(cudaMemcpy(dAug, hAug.data(), hAug.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap mAug;
  {
    cuuint64_t dims[4] = {8, 128, 1, 1};
    cuuint64_t strides[3] = {16, 16 * 128, 16 * 128};
    cuuint32_t box[4] = {8, 128, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = encode(&mAug, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dAug, dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode(aug) failed %d\n", (int)r); exit(3); }
  }
  const int smem_bytes = 32768 + 24576 + 6144 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  probe_kernel<<<1, 128, smem_bytes>>>(mA, mB, mAug, dX, dD, ddump, pa);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());

  if (pa.mode == 0) {
    std::vector<uint8_t> dump(24576);
    CK(cudaMemcpy(dump.data(), ddump, 24576, cudaMemcpyDeviceToHost));
    const __nv_bfloat16* d16 = reinterpret_cast<const __nv_bfloat16*>(dump.data());
    int bad = 0;
    for (int r = 0; r < 128; ++r)
      for (int c = 0; c < 96; ++c) {
        int atom = c / 32, ch = (c % 32) / 8, e = c % 8;
        uint32_t off = atom * 8192 + sw64_offset(r, ch) + e * 2;
        if (__bfloat162float(d16[off / 2]) != Bm[r * 96 + c]) ++bad;
      }
    printf("mode 0: TMA SW64 layout mismatches = %d / %d\n", bad, 128 * 96);
    return 0;
  }
  const int ncol = (pa.mode == 1 || pa.mode >= 5) ? 128 : 96;
  std::vector<float> D(128 * ncol), R(128 * ncol, 0.f);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < ncol; ++n) {
      double acc = 0;
      if (pa.mode == 1 || pa.mode >= 5) for (int k = 0; k < 96; ++k) acc += (double)A[m * 96 + k] * Bm[n * 96 + k];
      if (pa.mode >= 5) acc += (double)AUG[n * 8] + AUG[n * 8 + 1] + AUG[n * 8 + 2];
      if (pa.mode == 2) for (int k = 0; k < 128; ++k) acc += (double)X[m * 128 + k] * Bm[k * 96 + n];
      if (pa.mode == 3) for (int k = 0; k < 128; ++k) acc += (double)X[k * 128 + m] * Bm[k * 96 + n];
      if (pa.mode == 4) for (int k = 0; k < 128; ++k) acc += (double)X[m * 128 + k] * Bm[k * 96 + n];
      R[m * ncol + n] = (float)acc;
    }
  double maxerr = 0, maxref = 0;
  for (size_t i = 0; i < D.size(); ++i) { maxerr = fmax(maxerr, fabs(D[i] - R[i])); maxref = fmax(maxref, fabs(R[i])); }
  printf("mode %d a(lbo=%u sbo=%u kadv=%u) b(lbo=%u sbo=%u kadv=%u): max_err=%.5f max_ref=%.3f  %s\n", pa.mode, pa.a_lbo,
         pa.a_sbo, pa.a_kadv, pa.b_lbo, pa.b_sbo, pa.b_kadv, maxerr, maxref, maxerr < 1e-2 ? "PASS" : "FAIL");
  return 0;
}
