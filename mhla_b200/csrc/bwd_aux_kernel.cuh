// Elementwise companions of the native backward (mhla_b200/autograd.py).  The gradient CONTRACTIONS of the MHLA operator
// run as launches of the forward kernel with permuted operands; what is left around them when the normaliser
//   den_i[t] = sum_j W_ij n_loc[j, t] + eps,  n_loc[j, t] = q_{j,t} . ksum_j        (mhla_dit/mhla/mhla.py:265-268)
// is on are two streaming passes over token rows, one before and one after those launches:
//   bwd_prep :  dO~[r, :] = dO[r, :] / den[r]                 (the upstream gradient of the un-normalised numerator, 16 bit)
//               dden[r]   = -(dO[r, :] . O[r, :]) / den[r]    (num = O * den, so d(num/den)/d den = -dO.O / den)
//   bwd_post :  dq[r, :] = dQn[r, :] + dnl[r] * ksum[blk(r), :]      (blk(r) = r / w: the row's block)
//               dk[r, :] = dKn[r, :] + dksum[blk(r), :]              (every token of a block receives the same dksum)
// Rows are contiguous [rows, D] 16-bit; each thread owns 8 channels (16-byte accesses), D / 8 threads per row.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace mhla {

struct BwdPrepParams {
  const void* dout; const void* out;   // [rows, D] 16-bit
  const float* den;                    // [rows]
  void* dnum;                          // [rows, D] 16-bit
  float* dden;                         // [rows]
  long long rows;
  int D, fp16;
};

struct BwdPostParams {
  const void* dqn; const void* dkn;    // [rows, D] 16-bit or NULL (roped numerator: the un-roped q, k only see the normaliser)
  const float* dnl;                    // [rows]
  const float* ksum; const float* dksum;   // [rows / w, D] fp32
  void* dq; void* dk;                  // [rows, D] 16-bit
  long long rows;
  int w, D, fp16;
};

__device__ __forceinline__ void aux_load8(const void* base, long long idx, int fp16, float (&f)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(base) + idx));
  const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t;
    if (fp16) t = __half22float2(*reinterpret_cast<const __half2*>(&w4[i]));
    else t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[i]));
    f[2 * i] = t.x; f[2 * i + 1] = t.y;
  }
}

__device__ __forceinline__ void aux_store8(void* base, long long idx, int fp16, const float (&f)[8]) {
  uint32_t w4[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (fp16) { __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]); w4[i] = *reinterpret_cast<uint32_t*>(&h); }
    else { __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]); w4[i] = *reinterpret_cast<uint32_t*>(&h); }
  }
  *reinterpret_cast<uint4*>(static_cast<uint16_t*>(base) + idx) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
}

// TPR = D / 8 threads per row (8 or 16: a power of two, so a row never straddles a warp and the dot product is a
// butterfly over TPR lanes).  Grid-stride over groups of blockDim.x / TPR rows; every lane of a warp runs the same number
// of iterations (the shuffles need the full warp), out-of-range rows are masked.
template <int TPR>
__global__ void __launch_bounds__(256) bwd_prep_kernel(const BwdPrepParams p) {
  const int rpc = blockDim.x / TPR;
  const int sub = threadIdx.x % TPR, rl = threadIdx.x / TPR;
  const long long stride = (long long)gridDim.x * rpc;
  for (long long base = (long long)blockIdx.x * rpc; base < p.rows; base += stride) {
    const long long row = base + rl;
    const bool ok = row < p.rows;
    float a[8], b[8];
    float s = 0.f, r = 0.f;
    if (ok) {
      const long long idx = row * p.D + sub * 8;
      aux_load8(p.dout, idx, p.fp16, a);
      aux_load8(p.out, idx, p.fp16, b);
      r = 1.0f / __ldg(p.den + row);
#pragma unroll
      for (int i = 0; i < 8; ++i) s = fmaf(a[i], b[i], s);
    }
#pragma unroll
    for (int o = TPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (ok) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] *= r;
      aux_store8(p.dnum, row * p.D + sub * 8, p.fp16, a);
      if (sub == 0) p.dden[row] = -s * r;
    }
  }
}

__global__ void __launch_bounds__(256) bwd_post_kernel(const BwdPostParams p) {
  const int tpr = p.D / 8;
  const long long total = p.rows * tpr;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long row = t / tpr;
    const int c0 = (int)(t % tpr) * 8;
    const long long blk = row / p.w;
    const long long idx = row * p.D + c0;
    float a[8], b[8];
    if (p.dqn) { aux_load8(p.dqn, idx, p.fp16, a); aux_load8(p.dkn, idx, p.fp16, b); }
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) { a[i] = 0.f; b[i] = 0.f; }
    }
    const float dn = __ldg(p.dnl + row);
    const float4* ks = reinterpret_cast<const float4*>(p.ksum + blk * p.D + c0);
    const float4* dk = reinterpret_cast<const float4*>(p.dksum + blk * p.D + c0);
    const float4 k0 = __ldg(ks), k1 = __ldg(ks + 1), d0 = __ldg(dk), d1 = __ldg(dk + 1);
    const float kk[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
    const float dd[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = fmaf(dn, kk[i], a[i]); b[i] += dd[i]; }
    aux_store8(p.dq, idx, p.fp16, a);
    aux_store8(p.dk, idx, p.fp16, b);
  }
}

// Per-block (weighted) column sums over the token axis: out[blk, :] = sum_t wgt[blk, t] * x[blk, t, :]  (wgt NULL: 1).
//   ksum_j = sum_t k_{j,t}            (mhla.py:265)            dksum_j = sum_t dnl[j, t] q_{j,t}   (its gradient w.r.t. ksum)
// One CTA per block at a time (grid-stride), 256 threads = (256 / (D/8)) token rows x (D/8) channel groups of 8.
struct BlockSumParams {
  const void* x;        // [blocks, w, D] 16-bit
  const float* wgt;     // [blocks, w] or NULL
  float* out;           // [blocks, D]
  long long blocks;
  int w, D, fp16;
};

__global__ void __launch_bounds__(256) block_wsum_kernel(const BlockSumParams p) {
  __shared__ float red[32][129];
  const int tpr = p.D / 8, rpc = 256 / tpr;
  const int sub = threadIdx.x % tpr, rl = threadIdx.x / tpr;
  for (long long blk = blockIdx.x; blk < p.blocks; blk += gridDim.x) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    const long long base = blk * p.w;
    for (int t = rl; t < p.w; t += rpc) {
      float a[8];
      aux_load8(p.x, (base + t) * p.D + sub * 8, p.fp16, a);
      const float wv = p.wgt ? __ldg(p.wgt + base + t) : 1.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(wv, a[i], acc[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) red[rl][sub * 8 + i] = acc[i];
    __syncthreads();
    if ((int)threadIdx.x < p.D) {
      float s = 0.f;
      for (int r = 0; r < rpc; ++r) s += red[r][threadIdx.x];
      p.out[blk * p.D + threadIdx.x] = s;
    }
    __syncthreads();
  }
}

}  // namespace mhla
