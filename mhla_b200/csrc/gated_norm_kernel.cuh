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

}  // namespace mhla
