// tc_probe.cu -- micro-benchmarks that size the conv_tc_kernel design (run under gpurun; results in profiles/).
//   1. MMA-only: back-to-back tcgen05.mma (kind::f16, M=128) from static shared-memory operands, one thread issuing,
//      for N = 64/128/256 and 1 or 2 alternating accumulators -> cycles per MMA.
//   2. TMA-only: a producer lane streams [rows x 128 B] boxes (SWIZZLE_128B, 2-D map over an L2-resident buffer)
//      through an S-stage ring, the consumer releases a stage as soon as it lands -> bytes/clk/SM versus
//      stages, boxes per stage and rows per box.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../straps-3dhumanshapepose_b200/csrc tc_probe.cu -o tc_probe
#include "common.cuh"
#include <vector>
#include <cstdlib>

namespace straps {
void set_error(const char*, ...) {}
std::atomic<unsigned long long> g_launches{0};
}
using namespace straps;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

// ---------------------------------------------------------------- MMA only
template <int N, int NACC>
__global__ void __launch_bounds__(128, 1) mma_probe(int iters, long long* out_cycles) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  // operands: zeros are fine for timing
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&tslot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tslot;
  if (threadIdx.x == 0) {
    const uint64_t a = umma_desc_sw128(smem_u32(smem)), b = umma_desc_sw128(smem_u32(smem + 16384));
    const uint32_t idesc = umma_idesc_f16(128, N);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_f16(tbase + ((i * 4 + k) % NACC) * N, a + 2 * k, b + 2 * k, idesc, 1);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out_cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tbase); }
}

// ---------------------------------------------------------------- TMA only
__global__ void __launch_bounds__(64, 1) tma_probe(const __grid_constant__ CUtensorMap map, int stages, int boxes, int box_bytes,
                                                   int iters, int total_rows, int rows, long long* out_cycles) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[16], empty[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    fence_mbar_init();
  }
  __syncthreads();
  const int stage_bytes = boxes * box_bytes;
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    uint32_t row = (blockIdx.x * 7919u * rows) % (uint32_t)total_rows;
    for (int it = 0; it < iters; ++it) {
      const int st = it % stages;
      mbar_wait(&empty[st], ((it / stages) & 1) ^ 1);
      mbar_arrive_expect_tx(&full[st], stage_bytes);
      for (int b = 0; b < boxes; ++b) {
        tma_load_2d(smem + st * stage_bytes + b * box_bytes, &map, &full[st], 0, (int)row);
        row += rows;
        if (row + rows > (uint32_t)total_rows) row = 0;
      }
    }
  } else if (threadIdx.x == 32) {
    for (int it = 0; it < iters; ++it) {
      const int st = it % stages;
      mbar_wait(&full[st], (it / stages) & 1);
      mbar_arrive(&empty[st]);
    }
    out_cycles[blockIdx.x] = clock64() - t0;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int N, int NACC>
static void run_mma(int sms) {
  long long* d;
  CK(cudaMalloc(&d, sms * sizeof(long long)));
  const int smem = 16384 + 32768 + 2048, iters = 2000;
  CK(cudaFuncSetAttribute(mma_probe<N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  mma_probe<N, NACC><<<sms, 128, smem>>>(iters, d);
  CK(cudaDeviceSynchronize());
  std::vector<long long> h(sms);
  CK(cudaMemcpy(h.data(), d, sms * sizeof(long long), cudaMemcpyDeviceToHost));
  double avg = 0;
  for (long long c : h) avg += c;
  avg /= sms;
  printf("MMA  N=%3d accumulators=%d : %.1f cycles per tcgen05.mma (M128 x N x K16), %d CTAs\n", N, NACC, avg / (iters * 4.0), sms);
  cudaFree(d);
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  run_mma<64, 1>(sms); run_mma<128, 1>(sms); run_mma<256, 1>(sms); run_mma<128, 2>(sms); run_mma<64, 4>(sms);
  run_mma<128, 1>(1);

  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
  const int total_rows = 512 * 1024;   // 64 MB of 128-byte rows: stays in the 126 MB L2
  void* buf;
  CK(cudaMalloc(&buf, (size_t)total_rows * 128));
  CK(cudaMemset(buf, 0, (size_t)total_rows * 128));
  long long* d;
  CK(cudaMalloc(&d, sms * sizeof(long long)));
  CK(cudaFuncSetAttribute(tma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int cfgs[][3] = {  // stages, boxes per stage, rows per box
      {2, 4, 128}, {3, 4, 128}, {4, 4, 128}, {6, 4, 128}, {3, 2, 256}, {6, 2, 128}, {6, 1, 256}, {12, 1, 128},
      {3, 6, 128}, {2, 6, 128}, {4, 3, 128}, {4, 6, 64}, {8, 3, 64}};
  for (auto& c : cfgs) {
    const int stages = c[0], boxes = c[1], rows = c[2];
    CUtensorMap map;
    cuuint64_t dims[2] = {64, (cuuint64_t)total_rows};
    cuuint64_t str[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)rows};
    cuuint32_t es[2] = {1, 1};
    CUresult rc = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { printf("encode failed %d\n", (int)rc); return 1; }
    const int iters = 400, box_bytes = rows * 128;
    for (int pass = 0; pass < 2; ++pass) {   // pass 0 warms L2
      tma_probe<<<sms, 64, stages * boxes * box_bytes + 2048>>>(map, stages, boxes, box_bytes, iters, total_rows, rows, d);
      CK(cudaDeviceSynchronize());
    }
    std::vector<long long> h(sms);
    CK(cudaMemcpy(h.data(), d, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    double avg = 0;
    for (long long x : h) avg += x;
    avg /= sms;
    printf("TMA  stages=%2d boxes/stage=%d rows/box=%3d (stage %3d KB): %.1f B/clk/SM, %.0f cycles per stage\n", stages, boxes, rows,
           boxes * box_bytes / 1024, (double)iters * boxes * box_bytes / avg, avg / iters);
  }
  return 0;
}
