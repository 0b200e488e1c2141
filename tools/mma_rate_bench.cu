// tcgen05.mma issue-rate microbenchmark for the operand layouts the kernels use (not part of the product).
#include <cstdio>
#include "ptx.cuh"
using namespace mhla;

struct Cfg { int a_major, b_major, M, N, with_ones, reps; const char* name; };

__global__ void rate_kernel(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 65536, o0 = a0 + 140 * 1024;
    const uint32_t idesc = make_idesc(1, c.a_major, c.b_major, c.M, c.N);
    const uint32_t idesc1 = make_idesc(1, c.a_major, 1, c.M, 16);
    const uint64_t dones = make_smem_desc(o0, 256, 128, kSwizzleNone);
    const long long t0 = clock64();
    for (int r = 0; r < c.reps; ++r) {
      for (int ks = 0; ks < 8; ++ks) {
        const uint64_t da = c.a_major ? make_smem_desc(a0 + ks * 2048, 16384, 1024, kSwizzle128)
                                      : make_smem_desc(a0 + (ks >> 2) * 16384 + (ks & 3) * 32, 0, 1024, kSwizzle128);
        const uint64_t db = c.b_major ? make_smem_desc(b0 + ks * 2048, 16384, 1024, kSwizzle128)
                                      : make_smem_desc(b0 + (ks >> 2) * 16384 + (ks & 3) * 32, 0, 1024, kSwizzle128);
        mma_f16_ss(tmem, da, db, idesc, 1);
        if (c.with_ones) mma_f16_ss(tmem + 256, da, dones, idesc1, 1);
      }
    }
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  const Cfg cfgs[] = {
      {1, 1, 64, 64, 0, 64, "P1 D=64  : A MN, B MN, M=64  N=64"},
      {1, 1, 64, 64, 1, 64, "P1 D=64 + ones MMA (N=16)"},
      {1, 1, 128, 128, 0, 64, "P1 D=128 : A MN, B MN, M=128 N=128"},
      {1, 1, 128, 128, 1, 64, "P1 D=128 + ones"},
      {0, 1, 128, 64, 0, 64, "P3 D=64  : A K,  B MN, M=128 N=64"},
      {0, 1, 128, 128, 0, 64, "P3 D=128 : A K,  B MN, M=128 N=128"},
      {0, 1, 128, 256, 0, 64, "P2       : A K,  B MN, M=128 N=256"},
      {0, 0, 128, 256, 0, 64, "ref      : A K,  B K,  M=128 N=256"},
      {0, 0, 128, 64, 0, 64, "ref      : A K,  B K,  M=128 N=64"},
      {0, 0, 64, 64, 0, 64, "ref      : A K,  B K,  M=64  N=64"},
  };
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  for (const Cfg& c : cfgs) {
    long long h = 0;
    for (int it = 0; it < 2; ++it) {
      rate_kernel<<<1, 128, 160 * 1024>>>(c, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: error %s\n", c.name, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    }
    const int n = c.reps * 8;
    printf("%-44s : %8lld cycles for %d k-steps -> %.1f cycles / k-step (K=16)\n", c.name, h, n, (double)h / n);
  }
  return 0;
}
