// Fused pre-processing of the Wan MHLA layer: what the reference does with ~10 elementwise
// launches, fp32 / complex128 intermediates and a 5-tensor concatenation
//   mhla_videogen/diffusion/model/wan/mhla_utils.py:267-276 (WanRMSNorm over the FULL channel dim, relu + eps),
//   :127-156 (3-axis RoPE on interleaved pairs, complex128), :303-316 (.float(), head split, cat)
// as ONE pass: per token row  y = relu(x * rsqrt(mean_C(x^2) + eps_n) * w) + eps ;  y_rope = rotate pairs (2i, 2i+1) of
// every head by the token's angle ; both written once, in the 16-bit I/O type, token-major [B, N, C] - exactly the
// layout the blockmix kernel's 3-D block view consumes (no rearrange, no fp32 copies, no host sync).
// One CTA per token, C / 8 threads, 8 consecutive channels (4 rotation pairs) per thread: 16-byte loads and stores.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace mhla {

struct WanPrepParams {
  const void* xq; const void* xk;          // [B*N, C] rows (row pitch ld_in elements), in_dtype
  void* q_rope; void* k_rope;              // [B*N, C] 16-bit outputs (roped); k_rope/q_rope may alias nothing else
  void* q_plain; void* k_plain;            // optional un-roped outputs (normaliser operands); NULL: not written
  const float* wq; const float* wk;        // RMSNorm weights [C] or NULL (no norm: qk_norm = False)
  const float* cos_t; const float* sin_t;  // [N, D/2] fp32 or NULL (no rope)
  long long ld_in;
  int rows, N, C, D;                       // rows = B*N
  int in_dtype;                            // 0 bf16, 1 fp16, 2 fp32
  int out_fp16;
  float eps_norm, eps;
};

template <int IN>   // 0 bf16, 1 fp16, 2 fp32
__device__ __forceinline__ void load8(const void* base, long long idx, float (&f)[8]) {
  if constexpr (IN == 2) {
    const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(base) + idx);
    const float4 a = __ldg(p), b = __ldg(p + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(base) + idx));
    const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t;
      if constexpr (IN == 0) t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[i]));
      else t = __half22float2(*reinterpret_cast<const __half2*>(&w4[i]));
      f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
  }
}

__device__ __forceinline__ void store8(void* base, long long idx, const float (&f)[8], int fp16) {
  uint32_t w4[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (fp16) { __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]); w4[i] = *reinterpret_cast<uint32_t*>(&h); }
    else { __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]); w4[i] = *reinterpret_cast<uint32_t*>(&h); }
  }
  *reinterpret_cast<uint4*>(static_cast<uint16_t*>(base) + idx) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
}

template <int IN>
__global__ void wan_prep_kernel(const WanPrepParams p) {
  // Persistent over token rows (grid = a few CTAs per SM): the loads of the NEXT row are issued before the current row is
  // reduced, normalised and stored, so every thread keeps two rows of reads in flight (one CTA per token left the
  // memory system at 2.6 TB/s: each row's load -> block reduction -> store chain was fully exposed).
  __shared__ float red[2][2][32];
  const int tid = threadIdx.x;
  const int c0 = tid * 8;
  const bool act = c0 < p.C;
  const int nw = (blockDim.x + 31) >> 5;
  float wqv[8], wkv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { wqv[i] = 1.f; wkv[i] = 1.f; }
  if (act && p.wq) load8<2>(p.wq, c0, wqv);
  if (act && p.wk) load8<2>(p.wk, c0, wkv);
  float qn[8], kn[8];
  int row = blockIdx.x;
  if (act && row < p.rows) {
    load8<IN>(p.xq, (long long)row * p.ld_in + c0, qn);
    load8<IN>(p.xk, (long long)row * p.ld_in + c0, kn);
  }
  for (int it = 0; row < p.rows; row += gridDim.x, ++it) {
    float q[8], k[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { q[i] = qn[i]; k[i] = kn[i]; }
    const int nrow = row + gridDim.x;
    if (act && nrow < p.rows) {
      load8<IN>(p.xq, (long long)nrow * p.ld_in + c0, qn);
      load8<IN>(p.xk, (long long)nrow * p.ld_in + c0, kn);
    }
    float sq = 0.f, sk = 0.f;
    if (act) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { sq = fmaf(q[i], q[i], sq); sk = fmaf(k[i], k[i], sk); }
    }
    // block-wide sums of squares (the norm runs over the FULL channel dim, across heads: wan/model.py:181-196)
    for (int o = 16; o > 0; o >>= 1) { sq += __shfl_xor_sync(0xffffffffu, sq, o); sk += __shfl_xor_sync(0xffffffffu, sk, o); }
    float (*rd)[32] = red[it & 1];                 // alternate buffers: one barrier per row
    if ((tid & 31) == 0) { rd[0][tid >> 5] = sq; rd[1][tid >> 5] = sk; }
    __syncthreads();
    if (!act) continue;
    sq = 0.f; sk = 0.f;
    for (int i = 0; i < nw; ++i) { sq += rd[0][i]; sk += rd[1][i]; }
    const float rq = p.wq ? rsqrtf(sq / (float)p.C + p.eps_norm) : 1.f;
    const float rk = p.wk ? rsqrtf(sk / (float)p.C + p.eps_norm) : 1.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      q[i] = fmaxf(q[i] * rq * wqv[i], 0.f) + p.eps;
      k[i] = fmaxf(k[i] * rk * wkv[i], 0.f) + p.eps;
    }
    const long long o = (long long)row * p.C + c0;
    if (p.q_plain) { store8(p.q_plain, o, q, p.out_fp16); store8(p.k_plain, o, k, p.out_fp16); }
    if (p.cos_t) {
      // interleaved-pair rotation (view_as_complex, mhla_utils.py:144-151): pair i of a head takes angle [token, i]
      const int tok = row % p.N, d0 = c0 % p.D;                  // 8 channels never straddle a head (D % 8 == 0)
      const float4 cs = __ldg(reinterpret_cast<const float4*>(p.cos_t + (long long)tok * (p.D / 2) + d0 / 2));
      const float4 sn = __ldg(reinterpret_cast<const float4*>(p.sin_t + (long long)tok * (p.D / 2) + d0 / 2));
      const float c4[4] = {cs.x, cs.y, cs.z, cs.w}, s4[4] = {sn.x, sn.y, sn.z, sn.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float a = q[2 * i], b = q[2 * i + 1], a2 = k[2 * i], b2 = k[2 * i + 1];
        q[2 * i] = a * c4[i] - b * s4[i]; q[2 * i + 1] = a * s4[i] + b * c4[i];
        k[2 * i] = a2 * c4[i] - b2 * s4[i]; k[2 * i + 1] = a2 * s4[i] + b2 * c4[i];
      }
    }
    store8(p.q_rope, o, q, p.out_fp16);
    store8(p.k_rope, o, k, p.out_fp16);
  }
}

}  // namespace mhla
