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
  CUtensorMap tmo;    // output tensor (TMA stores)
  int store_every;    // consumer stores the stage's first box back after every `store_every`-th stage (0: never)
  long long wrap;     // stage units wrap modulo this (small value: the stream is served from L2)
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
      const long long uu = p.wrap ? (u % p.wrap) : u;
      for (int b = 0; b < p.boxes_per_stage; ++b)
        tma_load_2d(smem + st * stage_bytes + b * box_bytes, &p.tm, &full[st], 0,
                    (int)((uu * p.boxes_per_stage + b) * p.box_rows), p.wrap ? kEvictLast : kEvictFirst);
      if (++st == p.nstages) { st = 0; ph ^= 1; }
    }
  } else if (threadIdx.x == 32) {
    int st = 0; uint32_t ph = 0; long long n = 0;
    for (long long u = blockIdx.x; u < nst; u += gridDim.x) {
      mbar_wait(&full[st], ph);
      if (p.hold_cycles) { const long long t0 = clock64(); while (clock64() - t0 < p.hold_cycles) {} }
      if (p.store_every && (n++ % p.store_every) == 0) {
        fence_proxy_async_smem();
        for (int b = 0; b < p.boxes_per_stage; ++b)
          tma_store_2d(&p.tmo, smem + st * stage_bytes + b * box_bytes, 0, (int)((u * p.boxes_per_stage + b) * p.box_rows));
        tma_store_commit();
        tma_store_wait_read<0>();
      }
      mbar_arrive(&empty[st]);
      if (++st == p.nstages) { st = 0; ph ^= 1; }
    }
    tma_store_wait_all<0>();
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
  void* obuf;
  cudaMalloc(&obuf, bytes);
  struct Cfg { int box_rows, boxes_per_stage, nstages, ctas_per_sm, hold, grid, store_every, l2_mb; };
  const Cfg cfgs[] = {
      {128, 2, 6, 1, 0, 148, 0, 0},  {128, 2, 5, 1, 0, 148, 0, 0},  {128, 2, 5, 1, 0, 130, 0, 0}, {128, 2, 5, 1, 0, 112, 0, 0},
      {128, 2, 5, 1, 0, 74, 0, 0},   {128, 2, 5, 1, 0, 37, 0, 0},   {128, 2, 5, 1, 0, 8, 0, 0},   {128, 2, 5, 1, 0, 1, 0, 0},
      {128, 2, 5, 1, 1500, 148, 0, 0}, {128, 2, 5, 1, 3000, 148, 0, 0}, {128, 2, 3, 1, 1500, 148, 0, 0},
      {128, 2, 5, 1, 0, 148, 4, 0},  {128, 2, 5, 1, 0, 148, 3, 0},  {128, 2, 5, 1, 0, 148, 2, 0},  {128, 2, 5, 1, 0, 148, 1, 0},
      {128, 2, 5, 1, 0, 112, 3, 0},
      {128, 2, 5, 1, 0, 148, 0, 32}, {128, 2, 5, 1, 0, 112, 0, 32}, {128, 2, 5, 1, 0, 18, 0, 32},  {128, 2, 5, 1, 0, 1, 0, 32},
      {128, 2, 2, 1, 0, 148, 0, 32}, {128, 2, 5, 1, 0, 148, 3, 32},
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
    enc(&p.tmo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, obuf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    p.store_every = c.store_every;
    p.wrap = c.l2_mb ? ((long long)c.l2_mb << 20) / (c.box_rows * 128 * c.boxes_per_stage) : 0;
    const int smem = c.box_rows * 128 * c.boxes_per_stage * c.nstages;
    const int grid = c.grid * c.ctas_per_sm;
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
    const double wr = c.store_every ? (double)bytes / c.store_every : 0.0;
    printf("box %3d x128B, %d boxes/stage, %2d stages, grid %3d, hold %4d cyc, store 1/%d, src %s: %7.1f us  read %7.1f GB/s  read+write %7.1f GB/s  (%.1f B/clk/SM @1.9GHz)\n",
           c.box_rows, c.boxes_per_stage, c.nstages, grid, c.hold, c.store_every, c.l2_mb ? "L2 " : "HBM", best * 1e3,
           bytes / (best * 1e-3) / 1e9, (bytes + wr) / (best * 1e-3) / 1e9, (bytes + wr) / (best * 1e-3) / grid / 1.9e9);
  }
  return 0;
}
