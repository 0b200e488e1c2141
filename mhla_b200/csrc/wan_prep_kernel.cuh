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
  // after which the stage is refilled.
  // ncu on the first ring version (profiles/r02c_wanprep_ncu_details_ring.txt): 43 % DRAM throughput but IPC 2.8 - the
  // kernel is ISSUE-bound, ~390 instructions per thread and row for 8 channels.  Hence: 16 channels (two 16-byte chunks,
  // blockDim apart) per thread so that the per-row fixed cost (mbarrier wait, shuffles, block reduction, rsqrt) is paid by
  // half as many threads; ring stage / phase, row pointers and the token index advance by additions (no 64-bit
  // multiplies, no integer divisions in the loop); relu(x * r * w) = relu(x * w) * r as FMUL + FMNMX + FFMA.
  // History (Wan layer, B = 1): one CTA per token 2.6 TB/s; one / two rows of register look-ahead 130 / 103 us; ring 98 us.
  extern __shared__ __align__(128) uint8_t prep_smem[];
  __shared__ float2 red[2][32];
  const int tid = threadIdx.x;
  const int nchunk = p.C >> 3;
  const int cA = tid * 8, cB = (tid + (int)blockDim.x) * 8;            // first channels of this thread's two chunks
  const bool actA = tid < nchunk, actB = tid + (int)blockDim.x < nchunk;
  const int nw = (blockDim.x + 31) >> 5;
  const int G = (int)gridDim.x;
  const int S = p.stages;
  constexpr int kEsz = IN == 2 ? 4 : 2;
  const uint32_t rowb = (uint32_t)p.C * kEsz;                          // bytes of one input row
  const uint32_t angb = p.cos_t ? (uint32_t)p.D * 2u : 0u;             // bytes of one cos (or sin) row: D/2 floats
  uint64_t* bars = reinterpret_cast<uint64_t*>(prep_smem + (size_t)S * p.stage_bytes);
  const int aA = (cA % p.D) * 2, aB = (cB % p.D) * 2;                  // byte offset of the chunk's 4 angles in a cos row
  if (tid == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  // producer state (thread 0): next row to stage, its token, its source pointers - all advanced by additions
  const size_t in_step = (size_t)G * p.ld_in * kEsz;
  long long irow = blockIdx.x;
  int itok = (int)(irow % p.N);
  const int gmod = G % p.N;
  const uint8_t* iq = static_cast<const uint8_t*>(p.xq) + (size_t)irow * p.ld_in * kEsz;
  const uint8_t* ik = static_cast<const uint8_t*>(p.xk) + (size_t)irow * p.ld_in * kEsz;
  auto issue = [&](int s) {   // thread 0: stage row `irow` into stage s (caller checked irow < rows), then advance
    uint8_t* st = prep_smem + (size_t)s * p.stage_bytes;
    mbar_arrive_expect_tx(&bars[s], 2u * rowb + 2u * angb);
    prep_bulk_load(st, iq, rowb, &bars[s]);
    prep_bulk_load(st + rowb, ik, rowb, &bars[s]);
    if (angb) {
      prep_bulk_load(st + 2 * rowb, p.cos_t + (size_t)itok * (p.D / 2), angb, &bars[s]);
      prep_bulk_load(st + 2 * rowb + angb, p.sin_t + (size_t)itok * (p.D / 2), angb, &bars[s]);
    }
    irow += G; iq += in_step; ik += in_step;
    itok += gmod; if (itok >= p.N) itok -= p.N;
  };
  if (tid == 0)
    for (int s = 0; s < S && irow < p.rows; ++s) issue(s);
  // consumer state: output element offset of this row's first channel, ring stage and phase
  size_t orow = (size_t)blockIdx.x * p.C;
  const size_t out_step = (size_t)G * p.C;
  int s = 0;
  uint32_t ph = 0;
  uint32_t flip = 0;
  // one 8-channel chunk: normalise + relu + eps, optional plain store, rotation, roped store
  auto finish = [&](float (&q)[8], float (&k)[8], const float4& cs, const float4& sn, int c0, float rq, float rk, size_t o) {
    float wv[8];
    if (p.wq) {
      load8<2>(p.wq, c0, wv);
#pragma unroll
      for (int i = 0; i < 8; ++i) q[i] = fmaf(fmaxf(q[i] * wv[i], 0.f), rq, p.eps);   // relu(x r w) = relu(x w) r, r > 0
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) q[i] = fmaxf(q[i], 0.f) + p.eps;
    }
    if (p.wk) {
      load8<2>(p.wk, c0, wv);
#pragma unroll
      for (int i = 0; i < 8; ++i) k[i] = fmaf(fmaxf(k[i] * wv[i], 0.f), rk, p.eps);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) k[i] = fmaxf(k[i], 0.f) + p.eps;
    }
    if (p.q_plain) { store8(p.q_plain, (long long)o, q, p.out_fp16); store8(p.k_plain, (long long)o, k, p.out_fp16); }
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
    store8(p.q_rope, (long long)o, q, p.out_fp16);
    store8(p.k_rope, (long long)o, k, p.out_fp16);
  };
  for (long long row = blockIdx.x; row < p.rows; row += G) {
    const uint8_t* st = prep_smem + (size_t)s * p.stage_bytes;
    mbar_wait(&bars[s], ph);
    float qa[8], ka[8], qb[8], kb[8];
    float4 csa = make_float4(1.f, 1.f, 1.f, 1.f), sna = make_float4(0.f, 0.f, 0.f, 0.f), csb = csa, snb = sna;
    float sq = 0.f, sk = 0.f;
    if (actA) {
      smem_get8<IN>(st, cA, qa);
      smem_get8<IN>(st + rowb, cA, ka);
      if (angb) {
        csa = *reinterpret_cast<const float4*>(st + 2 * rowb + aA);
        sna = *reinterpret_cast<const float4*>(st + 2 * rowb + angb + aA);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) { sq = fmaf(qa[i], qa[i], sq); sk = fmaf(ka[i], ka[i], sk); }
    }
    if (actB) {
      smem_get8<IN>(st, cB, qb);
      smem_get8<IN>(st + rowb, cB, kb);
      if (angb) {
        csb = *reinterpret_cast<const float4*>(st + 2 * rowb + aB);
        snb = *reinterpret_cast<const float4*>(st + 2 * rowb + angb + aB);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) { sq = fmaf(qb[i], qb[i], sq); sk = fmaf(kb[i], kb[i], sk); }
    }
    // (the angle loads must have RETURNED before the barrier below frees the stage: make their results live here)
    asm volatile("" ::"f"(csa.x), "f"(sna.x), "f"(csb.x), "f"(snb.x));
    // block-wide sums of squares (the norm runs over the FULL channel dim, across heads: wan/model.py:181-196)
    for (int o = 16; o > 0; o >>= 1) { sq += __shfl_xor_sync(0xffffffffu, sq, o); sk += __shfl_xor_sync(0xffffffffu, sk, o); }
    float2* rd = red[flip];                        // alternate buffers: one barrier per row
    if ((tid & 31) == 0) rd[tid >> 5] = make_float2(sq, sk);
    __syncthreads();                               // ... which also says: every thread has read stage s
    if (tid == 0 && irow < p.rows) {
      fence_proxy_async_smem();                    // generic-proxy reads of the stage before the async-proxy refill
      issue(s);
    }
    sq = 0.f; sk = 0.f;
#pragma unroll 4
    for (int i = 0; i < nw; ++i) { const float2 t = rd[i]; sq += t.x; sk += t.y; }
    const float inv_c = 1.0f / (float)p.C;
    const float rq = p.wq ? rsqrtf(fmaf(sq, inv_c, p.eps_norm)) : 1.f;
    const float rk = p.wk ? rsqrtf(fmaf(sk, inv_c, p.eps_norm)) : 1.f;
    if (actA) finish(qa, ka, csa, sna, cA, rq, rk, orow + cA);
    if (actB) finish(qb, kb, csb, snb, cB, rq, rk, orow + cB);
    orow += out_step;
    flip ^= 1u;
    if (++s == S) { s = 0; ph ^= 1u; }
  }
}

}  // namespace mhla
