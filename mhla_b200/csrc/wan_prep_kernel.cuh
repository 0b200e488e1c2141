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

#include "ptx.cuh"

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
  int stages, stage_bytes;                 // shared-memory ring of staged rows: [xq row | xk row | cos row | sin row] per stage
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

// 8 consecutive channels of a staged row (shared memory) as floats
template <int IN>
__device__ __forceinline__ void smem_get8(const uint8_t* rowp, int c0, float (&f)[8]) {
  if constexpr (IN == 2) {
    const float4 a = *reinterpret_cast<const float4*>(rowp + c0 * 4), b = *reinterpret_cast<const float4*>(rowp + c0 * 4 + 16);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else {
    const uint4 u = *reinterpret_cast<const uint4*>(rowp + c0 * 2);
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

// global -> shared bulk copy (TMA, 1-D), completion counted in bytes on an mbarrier; 16-byte aligned, size % 16 == 0
__device__ __forceinline__ void prep_bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int IN>
__global__ void wan_prep_kernel(const WanPrepParams p) {
  // Persistent over token rows.  The rows are STAGED: one thread issues bulk copies (TMA) of the next `stages` rows of
  // xq, xk and of the tokens' cos / sin rows into a shared-memory ring, completion on one mbarrier per stage; the CTA
  // reads a landed row, reduces, normalises, rotates and stores it, and the reduction's own __syncthreads is the point
  // after which the stage is refilled.  Bytes in flight per SM = CTAs x (stages - 1) x 6.5 KB (Wan) - not bounded by
  // registers.  History (Wan layer, B = 1): one CTA per token 2.6 TB/s; one / two rows of register look-ahead 130 / 103 us
  // (a load takes ~3 us under the kernel's own traffic, 64 B per thread in flight were not enough).
  extern __shared__ __align__(128) uint8_t prep_smem[];
  __shared__ float red[2][2][32];
  const int tid = threadIdx.x;
  const int c0 = tid * 8;
  const bool act = c0 < p.C;
  const int nw = (blockDim.x + 31) >> 5;
  const int G = (int)gridDim.x;
  const int S = p.stages;
  const uint32_t rowb = (uint32_t)p.C * (IN == 2 ? 4u : 2u);           // bytes of one input row
  const uint32_t angb = p.cos_t ? (uint32_t)p.D * 2u : 0u;             // bytes of one cos (or sin) row: D/2 floats
  uint64_t* bars = reinterpret_cast<uint64_t*>(prep_smem + (size_t)S * p.stage_bytes);
  const int d0 = c0 % p.D;                                             // 8 channels never straddle a head (D % 8 == 0)
  if (tid == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](int s, long long r) {   // thread 0
    uint8_t* st = prep_smem + (size_t)s * p.stage_bytes;
    mbar_arrive_expect_tx(&bars[s], 2u * rowb + 2u * angb);
    prep_bulk_load(st, static_cast<const uint8_t*>(p.xq) + (size_t)r * p.ld_in * (IN == 2 ? 4 : 2), rowb, &bars[s]);
    prep_bulk_load(st + rowb, static_cast<const uint8_t*>(p.xk) + (size_t)r * p.ld_in * (IN == 2 ? 4 : 2), rowb, &bars[s]);
    if (angb) {
      const long long tok = r % p.N;
      prep_bulk_load(st + 2 * rowb, p.cos_t + tok * (p.D / 2), angb, &bars[s]);
      prep_bulk_load(st + 2 * rowb + angb, p.sin_t + tok * (p.D / 2), angb, &bars[s]);
    }
  };
  int row = blockIdx.x;
  if (tid == 0)
    for (int s = 0; s < S; ++s)
      if ((long long)row + (long long)s * G < p.rows) issue(s, (long long)row + (long long)s * G);
  for (int it = 0; row < p.rows; row += G, ++it) {
    const int s = it % S;
    const uint8_t* st = prep_smem + (size_t)s * p.stage_bytes;
    mbar_wait(&bars[s], (uint32_t)(it / S) & 1u);
    float q[8], k[8];
    float4 cs = make_float4(1.f, 1.f, 1.f, 1.f), sn = make_float4(0.f, 0.f, 0.f, 0.f);
    float sq = 0.f, sk = 0.f;
    if (act) {
      smem_get8<IN>(st, c0, q);
      smem_get8<IN>(st + rowb, c0, k);
      if (angb) {
        cs = *reinterpret_cast<const float4*>(st + 2 * rowb + (d0 / 2) * 4);
        sn = *reinterpret_cast<const float4*>(st + 2 * rowb + angb + (d0 / 2) * 4);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) { sq = fmaf(q[i], q[i], sq); sk = fmaf(k[i], k[i], sk); }
      // (the angle loads must have RETURNED before the barrier below frees the stage: make their results live here)
      asm volatile("" ::"f"(cs.x), "f"(cs.w), "f"(sn.x), "f"(sn.w));
    }
    // block-wide sums of squares (the norm runs over the FULL channel dim, across heads: wan/model.py:181-196)
    for (int o = 16; o > 0; o >>= 1) { sq += __shfl_xor_sync(0xffffffffu, sq, o); sk += __shfl_xor_sync(0xffffffffu, sk, o); }
    float (*rd)[32] = red[it & 1];                 // alternate buffers: one barrier per row
    if ((tid & 31) == 0) { rd[0][tid >> 5] = sq; rd[1][tid >> 5] = sk; }
    __syncthreads();                               // ... which also says: every thread has read stage s
    if (tid == 0) {
      const long long nr = (long long)row + (long long)S * G;
      if (nr < p.rows) {
        fence_proxy_async_smem();                  // generic-proxy reads of the stage before the async-proxy refill
        issue(s, nr);
      }
    }
    if (!act) continue;
    sq = 0.f; sk = 0.f;
    for (int i = 0; i < nw; ++i) { sq += rd[0][i]; sk += rd[1][i]; }
    const float rq = p.wq ? rsqrtf(sq / (float)p.C + p.eps_norm) : 1.f;
    const float rk = p.wk ? rsqrtf(sk / (float)p.C + p.eps_norm) : 1.f;
    {
      float wv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) wv[i] = 1.f;
      if (p.wq) load8<2>(p.wq, c0, wv);
#pragma unroll
      for (int i = 0; i < 8; ++i) q[i] = fmaxf(q[i] * rq * wv[i], 0.f) + p.eps;
#pragma unroll
      for (int i = 0; i < 8; ++i) wv[i] = 1.f;
      if (p.wk) load8<2>(p.wk, c0, wv);
#pragma unroll
      for (int i = 0; i < 8; ++i) k[i] = fmaxf(k[i] * rk * wv[i], 0.f) + p.eps;
    }
    const long long o = (long long)row * p.C + c0;
    if (p.q_plain) { store8(p.q_plain, o, q, p.out_fp16); store8(p.k_plain, o, k, p.out_fp16); }
    if (angb) {
      // interleaved-pair rotation (view_as_complex, mhla_utils.py:144-151): pair i of a head takes angle [token, i]
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
