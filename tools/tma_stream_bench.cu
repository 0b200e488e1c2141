// Raw TMA streaming microbenchmark (not part of the product): how much HBM bandwidth can ONE producer lane per CTA pull
// with [rows][128 B] swizzle-128B boxes into an mbarrier ring when the consumer releases stages immediately?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I mhla_b200/csrc -o tools/tma_stream_bench tools/tma_stream_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include "ptx.cuh"

using namespace mhla;

struct P {
  CUtensorMap tm;
  long long nboxes;   // total boxes in the tensor
  int boxes_per_stage, nstages, box_rows, hold_cycles;
};

__global__ void __launch_bounds__(128) stream_kernel(const __grid_constant__ P p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[16], empty[16];
  const int box_bytes = p.box_rows * 128;
  const int stage_bytes = box_bytes * p.boxes_per_stage;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.nstages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    fence_barrier_init();
  }
  __syncthreads();
  const long long nst = p.nboxes / p.boxes_per_stage;   // stage-sized work units, strided over CTAs
  if (threadIdx.x == 0) {
    int st = 0; uint32_t ph = 0;
    for (long long u = blockIdx.x; u < nst; u += gridDim.x) {
      mbar_wait(&empty[st], ph ^ 1);
      mbar_arrive_expect_tx(&full[st], stage_bytes);
      for (int b = 0; b < p.boxes_per_stage; ++b)
        tma_load_2d(smem + st * stage_bytes + b * box_bytes, &p.tm, &full[st], 0,
                    (int)((u * p.boxes_per_stage + b) * p.box_rows), kEvictFirst);
      if (++st == p.nstages) { st = 0; ph ^= 1; }
    }
  } else if (threadIdx.x == 32) {
    int st = 0; uint32_t ph = 0;
    for (long long u = blockIdx.x; u < nst; u += gridDim.x) {
      mbar_wait(&full[st], ph);
      if (p.hold_cycles) { const long long t0 = clock64(); while (clock64() - t0 < p.hold_cycles) {} }
      mbar_arrive(&empty[st]);
      if (++st == p.nstages) { st = 0; ph ^= 1; }
    }
  }
}

int main() {
  const size_t bytes = 512ull << 20;   // 512 MiB > L2
  void* buf;
  cudaMalloc(&buf, bytes);
  cudaMemset(buf, 1, bytes);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  auto enc = reinterpret_cast<CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                           const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                           CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill)>(fn);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  struct Cfg { int box_rows, boxes_per_stage, nstages, ctas_per_sm, hold; };
  const Cfg cfgs[] = {
      {128, 2, 6, 1, 0}, {128, 2, 3, 1, 0}, {128, 2, 2, 1, 0}, {128, 1, 12, 1, 0}, {128, 4, 3, 1, 0},
      {256, 1, 6, 1, 0}, {64, 4, 6, 1, 0}, {128, 2, 3, 2, 0}, {128, 1, 6, 2, 0}, {128, 1, 4, 3, 0},
      {128, 2, 6, 1, 1500}, {128, 2, 6, 1, 3000},
  };
  for (const Cfg& c : cfgs) {
    P p;
    const uint64_t rows = bytes / 128;
    cuuint64_t dims[2] = {64, rows};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)c.box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&p.tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    p.nboxes = rows / c.box_rows;
    p.boxes_per_stage = c.boxes_per_stage; p.nstages = c.nstages; p.box_rows = c.box_rows; p.hold_cycles = c.hold;
    const int smem = c.box_rows * 128 * c.boxes_per_stage * c.nstages;
    const int grid = sms * c.ctas_per_sm;
    float best = 1e9f;
    for (int it = 0; it < 4; ++it) {
      cudaEventRecord(e0);
      stream_kernel<<<grid, 128, smem>>>(p);
      cudaEventRecord(e1);
      cudaError_t err = cudaEventSynchronize(e1);
      if (err != cudaSuccess) { printf("error %s\n", cudaGetErrorString(err)); return 1; }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    printf("box %3d rows x128B, %d boxes/stage, %2d stages (%3d KB smem), %d CTA/SM, hold %4d cyc: %7.1f us  %7.1f GB/s\n",
           c.box_rows, c.boxes_per_stage, c.nstages, smem / 1024, c.ctas_per_sm, c.hold, best * 1e3, bytes / (best * 1e-3) / 1e9);
  }
  return 0;
}
