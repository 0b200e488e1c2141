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

// One row's 8 channels of one tensor as loaded (16 bytes of 16-bit data or 32 bytes of fp32): the conversion to float
// happens when the row is consumed, so a prefetched row costs registers but never a scoreboard stall at issue time.
template <int IN>
struct Raw8 {
  uint4 a, b;   // b: fp32 inputs only
  __device__ __forceinline__ void load(const void* base, long long idx) {
    if constexpr (IN == 2) {
      const uint4* p = reinterpret_cast<const uint4*>(static_cast<const float*>(base) + idx);
      a = __ldg(p); b = __ldg(p + 1);
    } else {
      a = __ldg(reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(base) + idx));
    }
  }
  __device__ __forceinline__ void get(float (&f)[8]) const {
    if constexpr (IN == 2) {
      f[0] = __uint_as_float(a.x); f[1] = __uint_as_float(a.y); f[2] = __uint_as_float(a.z); f[3] = __uint_as_float(a.w);
      f[4] = __uint_as_float(b.x); f[5] = __uint_as_float(b.y); f[6] = __uint_as_float(b.z); f[7] = __uint_as_float(b.w);
    } else {
      const uint32_t w4[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 t;
        if constexpr (IN == 0) t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[i]));
        else t = __half22float2(*reinterpret_cast<const __half2*>(&w4[i]));
        f[2 * i] = t.x; f[2 * i + 1] = t.y;
      }
    }
  }
};

template <int IN>
__global__ void wan_prep_kernel(const WanPrepParams p) {
  // Persistent over token rows (grid = a few CTAs per SM).  Every thread keeps the reads of its next TWO rows in flight
  // as raw 16-byte registers (two register sets, loop unrolled by two - no register-to-register rotation that would wait
  // for the load) while the current row is reduced, normalised, rotated and stored: 64 bytes per thread x 1536 threads
  // per SM in flight.  History: one CTA per token left the memory system at 2.6 TB/s (load -> block reduction -> store
  // chain fully exposed), one row of look-ahead whose conversion sat right behind the load at 2.9 TB/s.
  __shared__ float red[2][2][32];
  const int tid = threadIdx.x;
  const int c0 = tid * 8;
  const bool act = c0 < p.C;
  const int nw = (blockDim.x + 31) >> 5;
  const int G = (int)gridDim.x;
  // (the RMSNorm weights are re-read per row - 64 bytes per thread from L1 - instead of living in 16 registers: the
  //  register budget decides how many CTAs, i.e. how many rows in flight, an SM holds)
  Raw8<IN> qa, ka, qb, kb;   // set a: rows blockIdx.x + 2nG, set b: rows blockIdx.x + (2n+1)G
  qa.a = qa.b = ka.a = ka.b = qb.a = qb.b = kb.a = kb.b = make_uint4(0u, 0u, 0u, 0u);
  // the token's rotation angles travel with the row's set: the tables are streamed from HBM once per launch (16 MB at the
  // Wan size), and a load issued where the rotation needs it exposed a DRAM round trip per row (122 us for the layer)
  float4 csa = make_float4(1.f, 1.f, 1.f, 1.f), sna = make_float4(0.f, 0.f, 0.f, 0.f), csb = csa, snb = sna;
  const int d0 = c0 % p.D;                                       // 8 channels never straddle a head (D % 8 == 0)
  auto load_angles = [&](long long r, float4& cs, float4& sn) {
    if (p.cos_t == nullptr) return;
    const long long tok = r % p.N;
    cs = __ldg(reinterpret_cast<const float4*>(p.cos_t + tok * (p.D / 2) + d0 / 2));
    sn = __ldg(reinterpret_cast<const float4*>(p.sin_t + tok * (p.D / 2) + d0 / 2));
  };
  int row = blockIdx.x;
  if (act && row < p.rows) {
    qa.load(p.xq, (long long)row * p.ld_in + c0);
    ka.load(p.xk, (long long)row * p.ld_in + c0);
    load_angles(row, csa, sna);
  }
  if (act && row + G < p.rows) {
    qb.load(p.xq, (long long)(row + G) * p.ld_in + c0);
    kb.load(p.xk, (long long)(row + G) * p.ld_in + c0);
    load_angles(row + G, csb, snb);
  }
  // one row: consume the register set, refill it with the row two steps ahead, then reduce / normalise / rotate / store
  auto body = [&](const int r, Raw8<IN>& qr, Raw8<IN>& kr, float4& csr, float4& snr, const int it) {
    float q[8], k[8];
    qr.get(q); kr.get(k);
    const float4 cs = csr, sn = snr;
    const long long nrow = (long long)r + 2ll * G;
    if (act && nrow < p.rows) {
      qr.load(p.xq, nrow * p.ld_in + c0);
      kr.load(p.xk, nrow * p.ld_in + c0);
      load_angles(nrow, csr, snr);
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
    if (!act) return;
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
    const long long o = (long long)r * p.C + c0;
    if (p.q_plain) { store8(p.q_plain, o, q, p.out_fp16); store8(p.k_plain, o, k, p.out_fp16); }
    if (p.cos_t) {
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
  };
  for (int it = 0; row < p.rows; it += 2) {
    body(row, qa, ka, csa, sna, it);
    row += G;
    if (row >= p.rows) break;
    body(row, qb, kb, csb, snb, it + 1);
    row += G;
  }
}

}  // namespace mhla
