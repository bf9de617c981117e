// shift_probe.cu -- feasibility probe for round 2's tap-shifted A operands ("Next" item ii).
//
// Question: can a SWIZZLE_128B K-major UMMA operand start at an arbitrary ROW of a tile that TMA wrote (start address = tile +
// s * 128 bytes, s not a multiple of 8), and what must the descriptor's base_offset field (bits 49-51) hold?  If yes, the 9 taps of
// a 3x3 convolution are 9 descriptors into ONE halo tile in shared memory instead of 9 TMA loads.
//
// One CTA: TMA loads A[256 rows][64 fp16] and B[64][64] (small integers: every product and sum is exact), then for each shift s and
// base_offset convention issues D = A[s : s+128] . B^T (M128 x N64 x K64, four K=16 steps) and copies D out of TMEM.  The host compares
// each D with the exact reference of EVERY shift and prints which rows the hardware really read.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../straps-3dhumanshapepose_b200/csrc shift_probe.cu -o shift_probe
#include "common.cuh"
#include <cuda_fp16.h>
#include <vector>
#include <cstdlib>

namespace straps {
void set_error(const char*, ...) {}
std::atomic<unsigned long long> g_launches{0};
}
using namespace straps;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int A_ROWS = 256, NB = 64, KD = 64, MROWS = 128;
constexpr int NSHIFT = 10, NMODE = 3;
__constant__ int c_shift[NSHIFT] = {0, 1, 2, 3, 5, 7, 8, 9, 66, 67};

__global__ void __launch_bounds__(192, 1) shift_probe(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                                      float* __restrict__ out /*[NSHIFT][NMODE][128][64]*/) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = smem;                           // 256 rows x 128 B = 32 KB
  unsigned char* sB = smem + A_ROWS * 128;            // 64 rows x 128 B = 8 KB
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar_load, 1); mbar_init(&bar_mma, 1); fence_mbar_init(); }
  if (warp == 1) tmem_alloc<64>(&tslot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tslot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar_load, A_ROWS * 128 + NB * 128);
    tma_load_2d(sA, &map_a, &bar_load, 0, 0);
    tma_load_2d(sB, &map_b, &bar_load, 0, 0);
  }
  mbar_wait(&bar_load, 0);
  uint32_t phase = 0;
  for (int si = 0; si < NSHIFT; ++si)
    for (int mode = 0; mode < NMODE; ++mode) {
      if (threadIdx.x == 0) {
        const uint32_t start = smem_u32(sA) + (uint32_t)c_shift[si] * 128u;
        const uint32_t row_phase = (start >> 7) & 7u;
        const uint32_t bo = mode == 0 ? 0u : mode == 1 ? row_phase : ((8u - row_phase) & 7u);
        const uint64_t a = umma_desc_sw128(start) | (static_cast<uint64_t>(bo) << 49);
        const uint64_t b = umma_desc_sw128(smem_u32(sB));
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < KD / 16; ++k) umma_f16(tbase, a + 2 * k, b + 2 * k, umma_idesc_f16(MROWS, NB), k != 0);
        umma_commit(&bar_mma);
      }
      mbar_wait(&bar_mma, phase);
      phase ^= 1;
      tc_fence_after();
      if (warp >= 2) {
        const int quad = warp & 3;
        float* o = out + (((size_t)si * NMODE + mode) * MROWS + quad * 32 + lane) * NB;
        for (int c0 = 0; c0 < NB; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tbase + ((uint32_t)(quad * 32) << 16) + c0, v);
          tmem_ld_wait();
          for (int i = 0; i < 32; ++i) o[c0 + i] = __uint_as_float(v[i]);
        }
      }
      tc_fence_before();
      __syncthreads();
    }
  if (warp == 1) { tc_fence_after(); tmem_dealloc<64>(tbase); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  std::vector<__half> hA((size_t)A_ROWS * KD), hB((size_t)NB * KD);
  std::vector<float> fA(hA.size()), fB(hB.size());
  srand(7);
  for (size_t i = 0; i < hA.size(); ++i) { fA[i] = (float)(rand() % 9 - 4); hA[i] = __float2half(fA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { fB[i] = (float)(rand() % 9 - 4); hB[i] = __float2half(fB[i]); }
  __half *dA, *dB;
  float* dOut;
  const size_t out_elems = (size_t)NSHIFT * NMODE * MROWS * NB;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dOut, out_elems * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dOut, 0xff, out_elems * 4));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
  CUtensorMap ma, mb;
  cuuint64_t da[2] = {KD, A_ROWS}, db[2] = {KD, NB};
  cuuint64_t str[1] = {KD * 2};
  cuuint32_t ba[2] = {KD, A_ROWS}, bb[2] = {KD, NB}, es[2] = {1, 1};
  if (enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dA, da, str, ba, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
      enc(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dB, db, str, bb, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    printf("tensor map encode failed\n");
    return 1;
  }
  const int smem = A_ROWS * 128 + NB * 128 + 2048;
  CK(cudaFuncSetAttribute(shift_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  shift_probe<<<1, 192, smem>>>(ma, mb, dOut);
  CK(cudaDeviceSynchronize());
  std::vector<float> hOut(out_elems);
  CK(cudaMemcpy(hOut.data(), dOut, out_elems * 4, cudaMemcpyDeviceToHost));
  // exact references for every possible shift
  const int MAXS = A_ROWS - MROWS;
  std::vector<float> ref((size_t)(MAXS + 1) * MROWS * NB);
  for (int s = 0; s <= MAXS; ++s)
    for (int m = 0; m < MROWS; ++m)
      for (int n = 0; n < NB; ++n) {
        float acc = 0.f;
        for (int k = 0; k < KD; ++k) acc += fA[(size_t)(s + m) * KD + k] * fB[(size_t)n * KD + k];
        ref[((size_t)s * MROWS + m) * NB + n] = acc;
      }
  const int shifts[NSHIFT] = {0, 1, 2, 3, 5, 7, 8, 9, 66, 67};
  const char* modes[NMODE] = {"base_offset=0", "base_offset=(addr>>7)&7", "base_offset=(8-(addr>>7)&7)&7"};
  int ok_mode[NMODE] = {0, 0, 0};
  for (int si = 0; si < NSHIFT; ++si)
    for (int mode = 0; mode < NMODE; ++mode) {
      const float* d = &hOut[((size_t)si * NMODE + mode) * MROWS * NB];
      int rows_ok = 0, best_s = -1, best_rows = -1;
      for (int s = 0; s <= MAXS; ++s) {
        int good = 0;
        for (int m = 0; m < MROWS; ++m) {
          bool same = true;
          for (int n = 0; n < NB && same; ++n) same = d[m * NB + n] == ref[((size_t)s * MROWS + m) * NB + n];
          good += same;
        }
        if (s == shifts[si]) rows_ok = good;
        if (good > best_rows) { best_rows = good; best_s = s; }
      }
      ok_mode[mode] += rows_ok == MROWS;
      printf("shift %3d  %-32s rows exact at the requested shift: %3d / 128   (best matching shift %3d: %3d rows)\n", shifts[si],
             modes[mode], rows_ok, best_s, best_rows);
    }
  for (int mode = 0; mode < NMODE; ++mode) printf("SUMMARY %-32s exact for %d of %d shifts\n", modes[mode], ok_mode[mode], NSHIFT);
  return 0;
}
