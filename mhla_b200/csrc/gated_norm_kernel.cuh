// Post-op of the NLP layer: FusedRMSNormGated(o, g) of fla/modules/fused_norm_gate.py:77-99
// as called by mhla_nlp/fla/layers/mhla.py:350-356 - per (token, head) row of the operator's output
//     y = o * rsqrt(mean_V(o^2) + eps) * weight * g * sigmoid(g)
// in ONE streaming pass (16-byte loads of o and g, fp32 math, one rounding).  It is a separate launch, not part of the
// causal kernel's readout epilogue: that kernel splits a V = 256 row over two items (two accumulator halves, possibly on
// different SMs), so the row's sum of squares is not available where the row is written.  g == NULL gives the plain
// RMSNorm (`g_norm`, layers/mhla.py:358-360).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "bwd_aux_kernel.cuh"   // aux_load8 / aux_store8

namespace mhla {

struct GatedNormParams {
  const void* x; const void* g; void* out;   // [rows, D] 16-bit; row pitches ld_x / ld_g elements, out contiguous
  const float* weight;                       // [D] fp32 or NULL
  long long rows, ld_x, ld_g;
  int D, fp16;
  float eps;
};

// TPR = D / 8 threads per row (a power of two <= 32)
template <int TPR>
__global__ void __launch_bounds__(256) gated_norm_kernel(const GatedNormParams p) {
  const int rpc = blockDim.x / TPR;
  const int sub = threadIdx.x % TPR, rl = threadIdx.x / TPR;
  float wv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) wv[i] = p.weight ? __ldg(p.weight + sub * 8 + i) : 1.f;
  const long long stride = (long long)gridDim.x * rpc;
  for (long long base = (long long)blockIdx.x * rpc; base < p.rows; base += stride) {
    const long long row = base + rl;
    const bool ok = row < p.rows;
    float a[8], b[8];
    float s = 0.f;
    if (ok) {
      aux_load8(p.x, row * p.ld_x + sub * 8, p.fp16, a);
      if (p.g) aux_load8(p.g, row * p.ld_g + sub * 8, p.fp16, b);
#pragma unroll
      for (int i = 0; i < 8; ++i) s = fmaf(a[i], a[i], s);
    }
#pragma unroll
    for (int o = TPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (ok) {
      const float r = rsqrtf(s / (float)p.D + p.eps);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float y = a[i] * r * wv[i];
        if (p.g) y *= __fdividef(b[i], 1.0f + __expf(-b[i]));
        a[i] = y;
      }
      aux_store8(p.out, row * p.D + sub * 8, p.fp16, a);
    }
  }
}

// Post-ops of the Wan / DiT layers as ONE streaming pass behind the operator (mhla_utils.py:360-366, wan/model.py:1001-1003,
// mhla.py:268-273):   out = x * silu(g) + add     (g == NULL: no gate, add == NULL: no additive term), fp32 math, one
// rounding.  Rows of C elements (C % 8 == 0) with their own pitches; 16 bytes per thread and tensor, grid-stride.
// (The same ops also exist INSIDE the readout epilogue - blockmix_kernel<D, G3D, true> - but an epilogue thread owns one
// token row, and its dependent global loads queue behind the TMA stream of a saturated memory system: measured 283 us
// fused vs 124 us + this pass on the Wan layer, profiles/r02c_notes.md.)
struct GateAddParams {
  const void* x; const void* g; const void* add; void* out;
  long long rows, ld_x, ld_g, ld_a, ld_o;   // row pitches in elements
  int C, fp16;
};

__global__ void __launch_bounds__(256) gate_add_kernel(const GateAddParams p) {
  const int vec = p.C / 8;
  const long long total = p.rows * vec;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const long long row = idx / vec;
    const int c = (int)(idx - row * vec) * 8;
    float a[8], b[8];
    aux_load8(p.x, row * p.ld_x + c, p.fp16, a);
    if (p.g) {
      aux_load8(p.g, row * p.ld_g + c, p.fp16, b);
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] *= __fdividef(b[i], 1.0f + __expf(-b[i]));
    }
    if (p.add) {
      aux_load8(p.add, row * p.ld_a + c, p.fp16, b);
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] += b[i];
    }
    aux_store8(p.out, row * p.ld_o + c, p.fp16, a);
  }
}

// LePE of the Wan layers (mhla_videogen/diffusion/model/wan/model.py: `self.lepe = nn.Conv3d(dim, dim, 3, padding=1,
// groups=dim)` applied to v, mhla_utils.py:289-296): a depthwise 3x3x3 convolution over the (F, H, W) token grid, computed
// directly on the TOKEN-major [B, F*H*W, C] tensor the v projection produces and written token-major - what the operator's
// post-op consumes.  The reference path rearranges to NCDHW, runs cuDNN's depthwise Conv3d (10 ms at the Wan size on a B200)
// and rearranges back (a transposing copy).  Here: one thread per (token, 8 channels), 27 taps with zero padding, each tap
// a 16-byte load of the neighbour token's channels (coalesced across the warp, L1/L2 hits for 26 of 27) and 8 FMAs with
// the tap's weights from a [27][C] fp32 table; fp32 accumulation, bias, one rounding.
struct DwConv3dParams {
  const void* x; void* out;           // [B, F*H*W, C] 16-bit; x row pitch ld_x elements, out contiguous
  const float* wt;                    // [27][C] fp32: wt[(kf*3 + kh)*3 + kw][c] = conv.weight[c, 0, kf, kh, kw]
  const float* bias;                  // [C] fp32 or NULL
  long long ld_x;
  int B, F, H, W, C, fp16;
};

// 16 bytes of 16-bit channels -> 8 floats
__device__ __forceinline__ void dw_unpack8(const uint4& u, int fp16, float (&f)[8]) {
  const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t;
    if (fp16) t = __half22float2(*reinterpret_cast<const __half2*>(&w4[i]));
    else t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[i]));
    f[2 * i] = t.x; f[2 * i + 1] = t.y;
  }
}

// One thread = 8 channels of kDwTile consecutive tokens along W: the 3 x 3 neighbour rows are walked once, each row's
// kDwTile + 2 columns are loaded once (16 bytes each) and every loaded column feeds up to three outputs; the three taps'
// weights of a row are loaded once for the whole tile.  (One output per thread: 408 us at the Wan size - 27 x-loads, 54
// weight loads and 216 conversions per output; tile of 4: 13.5 / 13.5 / 108.)
constexpr int kDwTile = 4;
__global__ void __launch_bounds__(256) dwconv3d_kernel(const DwConv3dParams p) {
  const int vec = p.C / 8;
  const int wt_n = (p.W + kDwTile - 1) / kDwTile;
  const long long total = (long long)p.B * p.F * p.H * wt_n * vec;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    long long t = idx / vec;
    const int c = (int)(idx - t * vec) * 8;
    const int w0 = (int)(t % wt_n) * kDwTile; t /= wt_n;
    const int h = (int)(t % p.H); t /= p.H;
    const int f = (int)(t % p.F);
    const long long b = t / p.F;
    float acc[kDwTile][8];
    {
      float bv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) bv[i] = 0.f;
      if (p.bias) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + c)), b1 = __ldg(reinterpret_cast<const float4*>(p.bias + c + 4));
        bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w; bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
      }
#pragma unroll
      for (int o = 0; o < kDwTile; ++o)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[o][i] = bv[i];
    }
#pragma unroll
    for (int kf = 0; kf < 3; ++kf) {
      const int ff = f + kf - 1;
      if (ff < 0 || ff >= p.F) continue;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int hh = h + kh - 1;
        if (hh < 0 || hh >= p.H) continue;
        const long long rowtok = ((b * p.F + ff) * p.H + hh) * (long long)p.W;     // token index of (b, ff, hh, 0)
        // the row's three taps
        float wv[3][8];
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float* wrow = p.wt + (long long)((kf * 3 + kh) * 3 + kw) * p.C + c;
          const float4 a0 = __ldg(reinterpret_cast<const float4*>(wrow)), a1 = __ldg(reinterpret_cast<const float4*>(wrow + 4));
          wv[kw][0] = a0.x; wv[kw][1] = a0.y; wv[kw][2] = a0.z; wv[kw][3] = a0.w;
          wv[kw][4] = a1.x; wv[kw][5] = a1.y; wv[kw][6] = a1.z; wv[kw][7] = a1.w;
        }
        // columns w0 - 1 .. w0 + kDwTile: issue all loads first, then convert and accumulate
        uint4 raw[kDwTile + 2];
#pragma unroll
        for (int col = 0; col < kDwTile + 2; ++col) {
          const int ww = w0 + col - 1;
          raw[col] = make_uint4(0u, 0u, 0u, 0u);                                      // zero padding (16-bit zeros)
          if (ww >= 0 && ww < p.W)
            raw[col] = __ldg(reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(p.x) + (rowtok + ww) * p.ld_x + c));
        }
#pragma unroll
        for (int col = 0; col < kDwTile + 2; ++col) {
          float xv[8];
          dw_unpack8(raw[col], p.fp16, xv);
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const int o = col - kw;                                                     // output w0 + o reads column o + kw
            if (o >= 0 && o < kDwTile) {
#pragma unroll
              for (int i = 0; i < 8; ++i) acc[o][i] = fmaf(xv[i], wv[kw][i], acc[o][i]);
            }
          }
        }
      }
    }
    const long long otok = ((b * p.F + f) * p.H + h) * (long long)p.W + w0;
#pragma unroll
    for (int o = 0; o < kDwTile; ++o)
      if (w0 + o < p.W) aux_store8(p.out, (otok + o) * p.C + c, p.fp16, acc[o]);
  }
}

}  // namespace mhla
