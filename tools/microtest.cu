// Standalone bring-up test for UMMA descriptor / layout assumptions (not part of the product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I mhla_b200/csrc -o gpurun_out/microtest tools/microtest.cu
// One CTA; operands are written to shared memory by hand in the swizzle-128B layouts the kernels assume, one
// tcgen05.mma chain is issued, the accumulator is read back and compared with a CPU product.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ptx.cuh"

using namespace mhla;

struct Cfg {
  int kind;      // 0 = bf16, 2 = tf32
  int a_major;   // 0 = K-major, 1 = MN-major
  int b_major;
  int M, N, K;   // K = total contraction length (multiple of the per-instruction K)
  int a_lbo, b_lbo;  // bytes (as used by the kernels), -1: default
};

// Element (mn, k) of an operand stored as swizzle-128B tiles.
//  K-major : tiles of [rows = MN][128 B of K]; tile t covers k in [t*epr, (t+1)*epr); tile pitch = MNext*128
//  MN-major: tiles of [rows = K][128 B of MN]; tile t covers mn in [t*epr, ...); tile pitch = Kext*128
__host__ __device__ inline size_t sw128_offset(int row, int byte_in_row) {
  const int chunk = byte_in_row >> 4, within = byte_in_row & 15;
  return (size_t)row * 128 + (size_t)((chunk ^ (row & 7)) << 4) + within;
}

__global__ void mma_test(Cfg c, const float* A, const float* Bm, float* Dout) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int es = c.kind == 2 ? 4 : 2;
  const int epr = 128 / es;  // elements per 128-byte row
  uint8_t* sa = smem;
  uint8_t* sb = smem + 65536;
  const int tid = threadIdx.x;
  // ---- fill A (M x K) and B (N x K) ----
  for (int idx = tid; idx < c.M * c.K; idx += blockDim.x) {
    const int m = idx / c.K, k = idx % c.K;
    const float val = A[idx];
    size_t off;
    if (c.a_major == 0) { const int t = k / epr; off = (size_t)t * c.M * 128 + sw128_offset(m, (k % epr) * es); }
    else { const int t = m / epr; off = (size_t)t * c.K * 128 + sw128_offset(k, (m % epr) * es); }
    if (es == 4) *reinterpret_cast<float*>(sa + off) = val;
    else *reinterpret_cast<__nv_bfloat16*>(sa + off) = __float2bfloat16(val);
  }
  for (int idx = tid; idx < c.N * c.K; idx += blockDim.x) {
    const int n = idx / c.K, k = idx % c.K;
    const float val = Bm[idx];
    size_t off;
    if (c.b_major == 0) { const int t = k / epr; off = (size_t)t * c.N * 128 + sw128_offset(n, (k % epr) * es); }
    else { const int t = n / epr; off = (size_t)t * c.K * 128 + sw128_offset(k, (n % epr) * es); }
    if (es == 4) *reinterpret_cast<float*>(sb + off) = val;
    else *reinterpret_cast<__nv_bfloat16*>(sb + off) = __float2bfloat16(val);
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (tid < 32) { tmem_alloc(&tmem_slot, 256); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const int kinstr = 32 / es;  // K per instruction
    const uint32_t idesc = make_idesc(c.kind == 2 ? 2 : 1, c.a_major, c.b_major, c.M, c.N);
    const uint32_t a0 = smem_u32(sa), b0 = smem_u32(sb);
    for (int ks = 0; ks < c.K / kinstr; ++ks) {
      uint64_t da, db;
      const int k0 = ks * kinstr;
      if (c.a_major == 0) {
        const int t = k0 / epr;
        da = make_smem_desc(a0 + t * c.M * 128 + (k0 % epr) * es, c.a_lbo < 0 ? 0 : c.a_lbo, 1024, kSwizzle128);
      } else {
        da = make_smem_desc(a0 + k0 * 128, c.a_lbo < 0 ? c.K * 128 : c.a_lbo, 1024, kSwizzle128);
      }
      if (c.b_major == 0) {
        const int t = k0 / epr;
        db = make_smem_desc(b0 + t * c.N * 128 + (k0 % epr) * es, c.b_lbo < 0 ? 0 : c.b_lbo, 1024, kSwizzle128);
      } else {
        db = make_smem_desc(b0 + k0 * 128, c.b_lbo < 0 ? c.K * 128 : c.b_lbo, 1024, kSwizzle128);
      }
      if (c.kind == 2) mma_tf32_ss(tmem, da, db, idesc, ks != 0);
      else mma_f16_ss(tmem, da, db, idesc, ks != 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int warp = tid >> 5, lane = tid & 31;
  uint32_t v[32];
  for (int c0 = 0; c0 < c.N; c0 += 32) {
    tmem_ld_x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    int row = -1;
    if (c.M == 128) row = tid;
    else if (lane < 16) row = warp * 16 + lane;
    if (row >= 0)
      for (int e = 0; e < 32 && c0 + e < c.N; ++e) Dout[row * c.N + c0 + e] = __uint_as_float(v[e]);
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 256);
}

static float tf32_round(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xFFFFE000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}
static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  const Cfg cfgs[] = {
      {0, 0, 0, 128, 64, 64, -1, -1},     // bf16 K/K (sanity)
      {0, 0, 1, 128, 64, 64, -1, -1},     // bf16 A K-major, B MN-major  (P3, D=64)
      {0, 0, 1, 128, 128, 128, -1, -1},   // P3, D=128 (B two MN atoms, A two K tiles)
      {0, 1, 1, 64, 64, 128, -1, -1},     // P1, D=64
      {2, 0, 0, 128, 128, 32, -1, -1},    // tf32 K/K
      {2, 0, 0, 128, 128, 32, 16, 16},    // tf32 K/K, LBO = 16 bytes
      {2, 0, 1, 128, 128, 32, -1, -1},    // tf32 A K-major, B MN-major  (P2)
      {2, 0, 1, 128, 128, 32, 16, -1},    // P2 with A LBO = 16 bytes
      {2, 1, 1, 128, 128, 32, -1, -1},    // tf32 MN/MN
      {2, 0, 1, 128, 32, 32, -1, -1},     // tf32 B MN-major single atom
  };
  float *dA, *dB, *dD;
  cudaMalloc(&dA, 128 * 256 * 4);
  cudaMalloc(&dB, 256 * 256 * 4);
  cudaMalloc(&dD, 128 * 256 * 4);
  cudaFuncSetAttribute(mma_test, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
  for (const Cfg& c : cfgs) {
    std::vector<float> A(c.M * c.K), B(c.N * c.K), D(c.M * c.N), R(c.M * c.N);
    srand(1);
    for (auto& x : A) x = (rand() % 2001 - 1000) / 1000.f;
    for (auto& x : B) x = (rand() % 2001 - 1000) / 1000.f;
    for (int m = 0; m < c.M; ++m)
      for (int n = 0; n < c.N; ++n) {
        double s = 0;
        for (int k = 0; k < c.K; ++k) {
          const float a = c.kind == 2 ? tf32_round(A[m * c.K + k]) : bf16_round(A[m * c.K + k]);
          const float b = c.kind == 2 ? tf32_round(B[n * c.K + k]) : bf16_round(B[n * c.K + k]);
          s += (double)a * b;
        }
        R[m * c.N + n] = (float)s;
      }
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xFF, 128 * 256 * 4);
    mma_test<<<1, 128, 140 * 1024>>>(c, dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cfg kind=%d a=%d b=%d: CUDA error %s\n", c.kind, c.a_major, c.b_major, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double err = 0, ref = 0, sumabs = 0;
    for (size_t i = 0; i < D.size(); ++i) { err += (D[i] - R[i]) * (double)(D[i] - R[i]); ref += R[i] * (double)R[i]; sumabs += fabs(D[i]); }
    printf("kind=%d a_major=%d b_major=%d M=%d N=%d K=%d lbo=(%d,%d): rel_err=%.3e  mean|D|=%.4f  D[0..3]=%.4f %.4f %.4f %.4f  R[0..3]=%.4f %.4f %.4f %.4f\n",
           c.kind, c.a_major, c.b_major, c.M, c.N, c.K, c.a_lbo, c.b_lbo, sqrt(err / ref), sumabs / D.size(), D[0], D[1], D[2],
           D[3], R[0], R[1], R[2], R[3]);
  }
  return 0;
}
