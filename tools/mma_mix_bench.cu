// Tensor-side floor of the attention kernels' tcgen05.mma instruction mixes, measured in isolation on one SM:
// one CTA, one issuing thread, operands = zero-filled shared memory / TMEM, no softmax, no TMA, no barriers in
// the loop.  Prints cycles per "iteration" of each mix, i.e. the time the tensor pipe (incl. its shared-memory /
// TMEM operand fetch) needs for the MMAs of one tile step when nothing else competes.
//   mode 0  backward step:  S^T 7xSS(128x128) | dP^T 7xSS(128x128) | dV 8xTS(128x96) | dQ 8xSS(A,B MN-major, 128x96)
//                           | dK 8xTS(128x96)                                        (nominal 2048 cycles)
//   mode 1  same with dK as SS (A K-major from smem)
//   mode 2  forward step:   2 x [QK^T 6xSS(128x64) + PV 4xTS(128x96)]                 (nominal 768 cycles)
//   mode 10 SS128x64 with ONE constant descriptor pair (no per-instruction descriptor arithmetic), mode 11 the same
//           for SS128x128, mode 12 SS128x64 issued by TWO warps at once (thread 0 and thread 32, own accumulators)
//   mode 13 backward step split over two issuing warps: {dV, S^T} on thread 0, {dQ, dK, dP^T} on thread 32
//   mode 14 forward step split per query tile: {QK^T0, PV0} on thread 0, {QK^T1, PV1} on thread 32
//   mode 3..9 single op types: 3 SS128x128 K-major/K-major | 4 TS128x96 B MN-major | 5 SS128x96 A MN/B MN
//                              | 6 SS128x96 A K-major/B MN | 7 SS128x64 | 8 TS128x96 with B K-major | 9 SS128x256
//   mode 15 dQ / dK k-steps interleaved (SS dK), 16 dV / S^T interleaved, 17 whole backward step interleaved that way from
//           one thread, 18 the same from two threads ({dV,S^T} | {dQ,dK,dP^T}), 19 dQ / dK interleaved with dK as TS,
//           21 / 22 the backward step from four / three issuing threads, 20 the shipped order (dV S^T dQ dK dP^T, SS dK) from one thread
// usage: mma_mix_bench <mode> [iters]
#include <cstdio>
#include <cstdlib>
#include "../aki_b200/csrc/sm100_ptx.cuh"

using namespace aki;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__global__ void __launch_bounds__(128, 1) mix_kernel(int mode, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bar, bar2, bars4[4];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 200 * 1024 / 16; i += 128)
    *reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)) + i * 16) = make_uint4(0, 0, 0, 0);
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar2), 1); for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars4[i]), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_base_s));
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    constexpr int ATOM = 8192;
    const uint32_t sA = base, sB = base + 32768, sC = base + 65536, sD = base + 98304;
    const uint64_t KMAJ = umma_smem_desc(0, 16, 512, UMMA_SW64), MNMAJ = umma_smem_desc(0, ATOM, 512, UMMA_SW64);
    auto km = [&](uint32_t b, int k) { return KMAJ | (uint64_t)(((b + (k >> 1) * ATOM + (k & 1) * 32) >> 4) & 0x3FFFu); };
    auto mn = [&](uint32_t b, int k) { return MNMAJ | (uint64_t)(((b + k * 1024) >> 4) & 0x3FFFu); };
    const uint32_t I128 = umma_idesc_bf16(128, 128, 0, 0), I96B = umma_idesc_bf16(128, 96, 0, 1),
                   I96AB = umma_idesc_bf16(128, 96, 1, 1), I64 = umma_idesc_bf16(128, 64, 0, 0),
                   I96 = umma_idesc_bf16(128, 96, 0, 0), I256 = umma_idesc_bf16(128, 256, 0, 0);
    auto ss128 = [&](uint32_t d, int n) { for (int k = 0; k < n; ++k) umma_ss(tmem + d, km(sA, k % 6), km(sB, k % 6), I128, k > 0); };
    auto ts96 = [&](uint32_t d, uint32_t a, int n) { for (int k = 0; k < n; ++k) umma_ts(tmem + d, tmem + a + 8 * k, mn(sC, k), I96B, 1); };
    auto ss96_amn = [&](uint32_t d, int n) { for (int k = 0; k < n; ++k) umma_ss(tmem + d, mn(sD, k), mn(sA, k), I96AB, k > 0); };
    auto ss96_akm = [&](uint32_t d, int n) { for (int k = 0; k < n; ++k) umma_ss(tmem + d, km(sD, k), mn(sB, k), I96B, 1); };
    auto ss64 = [&](uint32_t d, int n) { for (int k = 0; k < n; ++k) umma_ss(tmem + d, km(sA, k), km(sB, k), I64, k > 0); };
    auto ts96_bk = [&](uint32_t d, uint32_t a, int n) { for (int k = 0; k < n; ++k) umma_ts(tmem + d, tmem + a + 8 * k, km(sC, k % 6), I96, 1); };
    auto ss256 = [&](uint32_t d, int n) { for (int k = 0; k < n; ++k) umma_ss(tmem + d, km(sA, k % 6), km(sB, k % 6), I256, k > 0); };
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      switch (mode) {
        case 0: ss128(0, 7); ss128(128, 7); ts96(256, 0, 8); ss96_amn(128, 8); ts96(352, 448, 8); break;
        case 1: ss128(0, 7); ss128(128, 7); ts96(256, 0, 8); ss96_amn(128, 8); ss96_akm(352, 8); break;
        case 2: ss64(0, 6); ts96(256, 448, 4); ss64(128, 6); ts96(352, 480, 4); break;
        case 3: ss128(0, 8); break;
        case 4: ts96(256, 0, 8); break;
        case 5: ss96_amn(128, 8); break;
        case 6: ss96_akm(352, 8); break;
        case 7: ss64(0, 6); break;
        case 8: ts96_bk(256, 0, 8); break;
        case 9: ss256(0, 8); break;
        case 10: { const uint64_t a = km(sA, 0), b = km(sB, 0);
#pragma unroll
                   for (int k = 0; k < 6; ++k) umma_ss(tmem, a, b, I64, 1); } break;
        case 11: { const uint64_t a = km(sA, 0), b = km(sB, 0);
#pragma unroll
                   for (int k = 0; k < 8; ++k) umma_ss(tmem, a, b, I128, 1); } break;
        case 12: ss64(0, 6); break;
        case 13: ts96(256, 0, 8); ss128(0, 7); break;
        case 14: ss64(0, 6); ts96(256, 448, 4); break;
        // 15-19: consecutive MMAs alternate between two accumulators (is the per-MMA floor a dependency latency?)
        case 15: for (int k = 0; k < 8; ++k) { umma_ss(tmem + 128, mn(sD, k), mn(sA, k), I96AB, k > 0);
                                               umma_ss(tmem + 352, km(sD, k), mn(sB, k), I96B, 1); } break;
        case 16: for (int k = 0; k < 8; ++k) { umma_ts(tmem + 256, tmem + 448 + 8 * k, mn(sC, k), I96B, 1);
                                               if (k < 7) umma_ss(tmem, km(sA, k % 6), km(sB, k % 6), I128, k > 0); } break;
        case 17: for (int k = 0; k < 8; ++k) { umma_ts(tmem + 256, tmem + 448 + 8 * k, mn(sC, k), I96B, 1);
                                               if (k < 7) umma_ss(tmem, km(sA, k % 6), km(sB, k % 6), I128, k > 0); }
                 for (int k = 0; k < 8; ++k) { umma_ss(tmem + 128, mn(sD, k), mn(sA, k), I96AB, k > 0);
                                               umma_ss(tmem + 352, km(sD, k), mn(sB, k), I96B, 1); }
                 ss128(128, 7); break;
        case 18: for (int k = 0; k < 8; ++k) { umma_ts(tmem + 256, tmem + 448 + 8 * k, mn(sC, k), I96B, 1);
                                               if (k < 7) umma_ss(tmem, km(sA, k % 6), km(sB, k % 6), I128, k > 0); } break;
        case 19: for (int k = 0; k < 8; ++k) { umma_ss(tmem + 128, mn(sD, k), mn(sA, k), I96AB, k > 0);
                                               umma_ts(tmem + 352, tmem + 480 + 4 * (k & 7), mn(sB, k), I96B, 1); } break;
        case 21: ts96(256, 448, 8); break;                       // four issuing threads: dV | S^T | dQ, dP^T | dK
        case 22: ts96(256, 448, 8); ss128(0, 7); break;          // three: dV, S^T | dQ, dP^T | dK
        case 20: ts96(256, 448, 8); ss128(0, 7); ss96_amn(128, 8); ss96_akm(352, 8); ss128(128, 7); break;   // shipped order, one thread
      }
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    out[0] = t1 - t0;
  }
  if (tid == 32 && mode == 12) {   // second issuing thread (another warp): same mix into other TMEM columns
    const uint64_t KMAJ = umma_smem_desc(0, 16, 512, UMMA_SW64);
    const uint32_t sC = base + 65536, sD = base + 98304, I64 = umma_idesc_bf16(128, 64, 0, 0);
    auto km = [&](uint32_t b, int k) { return KMAJ | (uint64_t)(((b + (k >> 1) * 8192 + (k & 1) * 32) >> 4) & 0x3FFFu); };
    for (int it = 0; it < iters; ++it)
      for (int k = 0; k < 6; ++k) umma_ss(tmem + 256, km(sC, k), km(sD, k), I64, k > 0);
    umma_commit(smem_u32(&bar2));
    mbar_wait(smem_u32(&bar2), 0);
  }
  if ((mode == 21 || mode == 22) && tid > 0 && tid < 128 && (tid & 31) == 0) {
    constexpr int ATOM = 8192;
    const uint32_t sA = base, sB = base + 32768, sD = base + 98304;
    const uint64_t KMAJ = umma_smem_desc(0, 16, 512, UMMA_SW64), MNMAJ = umma_smem_desc(0, ATOM, 512, UMMA_SW64);
    auto km = [&](uint32_t b, int k) { return KMAJ | (uint64_t)(((b + (k >> 1) * ATOM + (k & 1) * 32) >> 4) & 0x3FFFu); };
    auto mn = [&](uint32_t b, int k) { return MNMAJ | (uint64_t)(((b + k * 1024) >> 4) & 0x3FFFu); };
    const uint32_t I128 = umma_idesc_bf16(128, 128, 0, 0), I96B = umma_idesc_bf16(128, 96, 0, 1),
                   I96AB = umma_idesc_bf16(128, 96, 1, 1);
    const int w = tid >> 5;
    for (int it = 0; it < iters; ++it) {
      if (w == 1 && mode == 21) for (int k = 0; k < 7; ++k) umma_ss(tmem, km(sA, k % 6), km(sB, k % 6), I128, k > 0);
      if (w == 2) { for (int k = 0; k < 8; ++k) umma_ss(tmem + 128, mn(sD, k), mn(sA, k), I96AB, k > 0);
                    for (int k = 0; k < 7; ++k) umma_ss(tmem + 128, km(sA, k % 6), km(sB, k % 6), I128, k > 0); }
      if (w == 3) for (int k = 0; k < 8; ++k) umma_ss(tmem + 352, km(sD, k), mn(sB, k), I96B, 1);
    }
    umma_commit(smem_u32(&bars4[w]));
    mbar_wait(smem_u32(&bars4[w]), 0);
  }
  if (tid == 32 && mode == 18) {   // stream B of mode 18: dQ / dK interleaved, then dP
    constexpr int ATOM = 8192;
    const uint32_t sA = base, sB = base + 32768, sD = base + 98304;
    const uint64_t KMAJ = umma_smem_desc(0, 16, 512, UMMA_SW64), MNMAJ = umma_smem_desc(0, ATOM, 512, UMMA_SW64);
    auto km = [&](uint32_t b, int k) { return KMAJ | (uint64_t)(((b + (k >> 1) * ATOM + (k & 1) * 32) >> 4) & 0x3FFFu); };
    auto mn = [&](uint32_t b, int k) { return MNMAJ | (uint64_t)(((b + k * 1024) >> 4) & 0x3FFFu); };
    const uint32_t I128 = umma_idesc_bf16(128, 128, 0, 0), I96B = umma_idesc_bf16(128, 96, 0, 1),
                   I96AB = umma_idesc_bf16(128, 96, 1, 1);
    for (int it = 0; it < iters; ++it) {
      for (int k = 0; k < 8; ++k) { umma_ss(tmem + 128, mn(sD, k), mn(sA, k), I96AB, k > 0);
                                    umma_ss(tmem + 352, km(sD, k), mn(sB, k), I96B, 1); }
      for (int k = 0; k < 7; ++k) umma_ss(tmem + 128, km(sA, k % 6), km(sB, k % 6), I128, k > 0);
    }
    umma_commit(smem_u32(&bar2));
    mbar_wait(smem_u32(&bar2), 0);
  }
  if (tid == 32 && (mode == 13 || mode == 14)) {
    constexpr int ATOM = 8192;
    const uint32_t sA = base, sB = base + 32768, sC = base + 65536, sD = base + 98304;
    const uint64_t KMAJ = umma_smem_desc(0, 16, 512, UMMA_SW64), MNMAJ = umma_smem_desc(0, ATOM, 512, UMMA_SW64);
    auto km = [&](uint32_t b, int k) { return KMAJ | (uint64_t)(((b + (k >> 1) * ATOM + (k & 1) * 32) >> 4) & 0x3FFFu); };
    auto mn = [&](uint32_t b, int k) { return MNMAJ | (uint64_t)(((b + k * 1024) >> 4) & 0x3FFFu); };
    const uint32_t I128 = umma_idesc_bf16(128, 128, 0, 0), I96B = umma_idesc_bf16(128, 96, 0, 1),
                   I96AB = umma_idesc_bf16(128, 96, 1, 1), I64 = umma_idesc_bf16(128, 64, 0, 0);
    for (int it = 0; it < iters; ++it) {
      if (mode == 13) {
        for (int k = 0; k < 8; ++k) umma_ss(tmem + 128, mn(sD, k), mn(sA, k), I96AB, k > 0);          // dQ
        for (int k = 0; k < 8; ++k) umma_ts(tmem + 352, tmem + 448 + 8 * k, mn(sC, k), I96B, 1);       // dK
        for (int k = 0; k < 7; ++k) umma_ss(tmem + 128, km(sA, k % 6), km(sB, k % 6), I128, k > 0);    // dP
      } else {
        for (int k = 0; k < 6; ++k) umma_ss(tmem + 128, km(sA, k), km(sB, k), I64, k > 0);
        for (int k = 0; k < 4; ++k) umma_ts(tmem + 352, tmem + 480 + 8 * k, mn(sC, k), I96B, 1);
      }
    }
    umma_commit(smem_u32(&bar2));
    mbar_wait(smem_u32(&bar2), 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0, iters = argc > 2 ? atoi(argv[2]) : 200;
  long long* d;
  CK(cudaMalloc(&d, 8));
  const int smem = 201 * 1024;
  CK(cudaFuncSetAttribute(mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int rep = 0; rep < 2; ++rep) {
    mix_kernel<<<1, 128, smem>>>(mode, iters, d);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
  }
  long long h;
  CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
  const double nominal[] = {2048, 2048, 768, 512, 384, 384, 384, 192, 384, 1024, 192, 512, 384, 2048, 768,
                            768, 832, 2048, 2048, 768, 2048, 2048, 2048};
  printf("mode %d: %.1f cycles per iteration (nominal tensor math %.0f)\n", mode, (double)h / iters, nominal[mode]);
  return 0;
}
