// Block-mixed MHLA forward for SHORT sequences (DiT / ViT: N = M*w <= 256 tokens per (b,h) unit, D = 64) as one
// persistent sm_100a kernel with NO workspace and NO cross-CTA dependencies.
//
// Replaces mhla_dit/mhla/mhla.py:262-268 (and its twin mhla_image_classification/.../attention/mhla.py:275-282) at the
// shapes those models run (DiT-S/2 256x256: M = 16 blocks of w = 16 tokens; ViT: M = 4 / 16 blocks of 49 / 16 tokens).
// For so few tokens the three-phase kernel (blockmix_kernel.cuh) is all latency: one P1 and one P3 item per 16-token
// block and a P1 -> P2 -> P3 dependency chain through L2 (166 us for batch 64, round 1).  Here the whole (b,h) unit
// lives on chip and the mixing becomes a MASK on the token-token scores - the same algebra the reference's formulas
// expand to:
//     O[t,:]  = ( sum_s  P[t,s] W[blk(t), blk(s)] V[s,:] ) / den[t],          P = Q K^T   (256 x 256, fp32 in TMEM)
//     n_loc[t] = sum_{s in blk(t)} P[t,s]                 (= q_t . ksum_blk(t), mhla.py:265-266)
//     den[t]  = sum_j W[blk(t), j] n_loc[j*w + t % w] + eps                   (the reference's quirk, kept)
// Roles (384 threads, 1 CTA / SM, grid = min(#units, #SMs), units strided over the CTAs):
//   warp 0  TMA producer: one 3-D box load per tensor brings the unit's [N][64] tile (rows in (block, token) order)
//           into a 2-stage ring (Q | K | V, 32 KB each), so unit g+1 streams in while unit g computes
//   warp 1  tcgen05 issuer: P = Q K^T (two M=128 row tiles x N<=256 key columns, SS MMA), then O = A V with the masked,
//           16-bit A read straight from TMEM (TS MMA): the scores never touch shared memory
//   warps 4-11  two epilogue warpgroups, one per row tile, one token row per thread: P (fp32) -> weight by W in
//           registers (exact fp32 weights - no hi/lo split needed here) -> n_loc -> round to 16 bit -> tcgen05.st back
//           over the columns just read; later O -> 1/den -> 16 bit -> the unit's Q buffer (free by then) -> one TMA store
// HBM traffic is exactly Q, K, V in and O out.
#pragma once
#include <cuda.h>
#include "ptx.cuh"

namespace mhla {

constexpr int kSnThreads = 384;
constexpr int kSnRows = 256;                       // token rows per unit (padded)
constexpr int kSnTile = kSnRows * 128;             // one [256][64 x 16-bit] swizzle-128B tile = 32 KB
constexpr int kSnStage = 3 * kSnTile;              // Q | K | V
constexpr int kSnMaxM = 64;
constexpr int kSnSmemW = 2 * kSnStage;             // weight table: [M][NP] per-column weights (<= 32 KB), or W [M][M] (generic path)
constexpr int kSnTable = 8192;                     // floats
constexpr int kSnSmemNl = kSnSmemW + kSnTable * 4; // n_loc [256] fp32
constexpr int kSnSmemBars = kSnSmemNl + kSnRows * 4;
constexpr int kSnSmemTotal = kSnSmemBars + 256;
constexpr int kSnSmemAlloc = kSnSmemTotal + 1024;

struct alignas(64) SmallNParams {
  CUtensorMap tmQ, tmK, tmV, tmO;   // rank-5 (d, w, M, H, B), box (64, w, M, 1, 1): one box = one (b,h) unit
  const float* mix;                 // [M][mix_ld] fp32
  long long mix_ld;
  int G, H, M, w, N;                // G = B*H units
  int normalize, is_fp16;
  float eps;
  unsigned long long* prof;         // optional [gridDim][16] cycle counters of warpgroup 0 (debug, tools/prof_smalln.py)
  // fused "+ lepe" of the readout (mhla_dit/mhla/mhla.py:271-273; ABI v4 out_add): 16-bit tensor laid out like out with its
  // own element strides, NULL = off.  Every epilogue thread adds its token row's 128 bytes before the rounding.
  const uint16_t* post_add;
  long long pa_sb, pa_sh, pa_sm, pa_sw;
};

__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]  (A: M lanes x K 16-bit elements, two per 32-bit column)
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(kSnThreads, 1) smalln_kernel(const __grid_constant__ SmallNParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* Wsm = reinterpret_cast<float*>(smem + kSnSmemW);
  float* nl_s = reinterpret_cast<float*>(smem + kSnSmemNl);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kSnSmemBars);   // [2] stage loaded
  uint64_t* empty = full + 2;                                         // [2] stage free (MMA commit + output store read)
  uint64_t* pfull = empty + 2;                                        // [2] scores of row tile ready (MMA -> epilogue)
  uint64_t* aready = pfull + 2;                                       // [2] masked A written to TMEM (epilogue -> MMA)
  uint64_t* ofull = aready + 2;                                       // [2] O accumulator ready
  uint64_t* tfree = ofull + 2;                                        // [1] TMEM drained by both warpgroups
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfree + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = p.N;
  const int ntile = N > 128 ? 2 : 1;                 // row tiles of 128 tokens
  const int NP = (N + 15) / 16 * 16;                 // key columns of the score tile (MMA N, multiple of 16)
  const uint32_t fmt16 = p.is_fp16 ? 0u : 1u;
  const uint32_t unit_bytes = (uint32_t)N * 128u;    // one tensor's tile of one unit

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 1); mbar_init(&empty[i], 2); mbar_init(&pfull[i], 1); mbar_init(&aready[i], 1);
      mbar_init(&ofull[i], 1);
    }
    mbar_init(tfree, 2);
    fence_barrier_init();
    const CUtensorMap* maps = &p.tmQ;
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(maps + i);
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  // rows N .. 255 of the K and V tiles are never written by TMA: zero both stages once (0 x garbage could be NaN)
  if (N < kSnRows) {
    uint4* z = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < 2 * kSnStage / 16; i += kSnThreads) z[i] = make_uint4(0, 0, 0, 0);
  }
  grid_dependency_wait();   // PDL: the mixing matrix may have been written by the previous kernel in the stream
  grid_launch_dependents();
  // Per-column weights Wsm[bt][col] = W[bt][blk(col)] (0 beyond the N real key columns) when the table fits: the epilogue
  // then masks a score with ONE multiply and vectorised broadcast loads.  Otherwise (M * NP > 8192, e.g. 64 blocks of 4
  // tokens) Wsm holds W [M][M] and the epilogue tracks the block boundaries itself.
  const bool use_table = p.M * NP <= kSnTable;
  const int tld = use_table ? NP : p.M;              // row pitch of Wsm
  if (use_table) {
    for (int i = threadIdx.x; i < p.M * NP; i += kSnThreads) {
      const int bt = i / NP, col = i - bt * NP;
      Wsm[i] = col < N ? __ldg(p.mix + (long long)bt * p.mix_ld + col / p.w) : 0.f;
    }
  } else {
    for (int i = threadIdx.x; i < p.M * p.M; i += kSnThreads) Wsm[i] = __ldg(p.mix + (long long)(i / p.M) * p.mix_ld + i % p.M);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================================================ TMA producer
    if (elect_one()) {
      uint32_t it = 0;
      for (int g = blockIdx.x; g < p.G; g += gridDim.x, ++it) {
        const int s = it & 1;
        mbar_wait(&empty[s], ((it >> 1) & 1) ^ 1);
        uint8_t* st = smem + s * kSnStage;
        const int b = g / p.H, h = g % p.H;
        mbar_arrive_expect_tx(&full[s], 3 * unit_bytes);
        tma_load_5d(st, &p.tmQ, &full[s], 0, 0, 0, h, b, kEvictFirst);
        tma_load_5d(st + kSnTile, &p.tmK, &full[s], 0, 0, 0, h, b, kEvictFirst);
        tma_load_5d(st + 2 * kSnTile, &p.tmV, &full[s], 0, 0, 0, h, b, kEvictFirst);
      }
    }
  } else if (warp == 1) {
    // ============================================================ tcgen05 issuer
    if (elect_one()) {
      const uint32_t idesc_qk = make_idesc(fmt16, 0, 0, 128, (uint32_t)NP);   // A = Q (K-major), B = K (K-major)
      const uint32_t idesc_pv = make_idesc(fmt16, 0, 1, 128, 64);             // A = scores (TMEM), B = V (MN-major)
      uint32_t it = 0;
      for (int g = blockIdx.x; g < p.G; g += gridDim.x, ++it) {
        const int s = it & 1;
        const uint32_t ph = it & 1;
        const uint32_t st = smem_u32(smem + s * kSnStage);
        mbar_wait(&full[s], (it >> 1) & 1);
        mbar_wait(tfree, ph ^ 1);                    // the previous unit's accumulators have been read
        tc_fence_after();
        for (int t = 0; t < ntile; ++t) {
          const uint32_t acc = tmem_base + t * 256;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {           // 64 channels = 4 k-steps of 32 bytes
            const uint64_t da = make_smem_desc(st + t * 16384 + ks * 32, 0, 1024, kSwizzle128);
            const uint64_t db = make_smem_desc(st + kSnTile + ks * 32, 0, 1024, kSwizzle128);
            mma_f16_ss(acc, da, db, idesc_qk, ks != 0);
          }
          mma_commit(&pfull[t]);
        }
        for (int t = 0; t < ntile; ++t) {
          const uint32_t acc = tmem_base + t * 256;
          mbar_wait(&aready[t], ph);
          tc_fence_after();
          for (int ks = 0; ks < NP / 16; ++ks) {     // 16 key tokens per k-step: 8 TMEM columns of A, 2048 B of V
            const uint64_t db = make_smem_desc(st + 2 * kSnTile + ks * 2048, 16384, 1024, kSwizzle128);
            mma_f16_ts(acc + 128, acc + ks * 8, db, idesc_pv, ks != 0);
          }
          mma_commit(&ofull[t]);
        }
        mma_commit(&empty[s]);                       // K and V of this stage are free (Q: see the output store)
      }
    }
  } else if (warp >= 4) {
    // ============================================================ epilogue warpgroups: warpgroup t owns row tile t
    const int q4 = warp & 3;
    const int t = (warp - 4) >> 2;
    const int et = threadIdx.x - 128 - t * 128;
    const int row = t * 128 + q4 * 32 + lane;        // token row = TMEM lane (128 t + ...)
    const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
    const bool active = t < ntile;
    const int w = p.w, M = p.M;
    const int blk_t = row / w, pos_t = row - blk_t * w;
    const bool row_ok = row < N;
    const float* Wrow = Wsm + (row_ok ? blk_t : 0) * tld;
    const int wstep = use_table ? w : 1;             // W[bt][j] = Wrow[j * wstep]
    const int lo = blk_t * w, hi = lo + w;           // key columns of the row's own block (n_loc)
    uint32_t it = 0;
    const bool prof_on = p.prof != nullptr && threadIdx.x == 128;
    long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tp = prof_on ? clock64() : 0;
    auto lap = [&](int k) { if (prof_on) { const long long n_ = clock64(); pc[k] += n_ - tp; tp = n_; } };
    for (int g = blockIdx.x; g < p.G; g += gridDim.x, ++it) {
      const int s = it & 1;
      const uint32_t ph = it & 1;
      uint8_t* st = smem + s * kSnStage;
      const uint32_t acc = tmem_base + t * 256 + lane_sel;
      float nl = 0.f;
      if (active) {
        mbar_wait(&pfull[t], ph);
        lap(0);   // wait for the scores (load + Q K^T)
        tc_fence_after();
        // scores -> masked 16-bit A, in place (the 16-bit columns [16c, 16c+16) lie behind the fp32 columns already read)
        const int nchunk = (NP + 31) / 32;
        if (use_table) {
          // one 32-column chunk: weights by vectorised broadcast loads, n_loc from the row's own block, 16-bit pack, store
          auto process = [&](const uint32_t (&v)[32], int c) {
            const float4* w4 = reinterpret_cast<const float4*>(Wrow + c * 32);
            float4 ww[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) ww[e] = w4[e];
            if (c * 32 < hi && c * 32 + 32 > lo) {
#pragma unroll
              for (int e = 0; e < 32; ++e) {
                const int col = c * 32 + e;
                if (col >= lo && col < hi) nl += __uint_as_float(v[e]);
              }
            }
            uint32_t pk[16];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float a0 = __uint_as_float(v[4 * e]) * ww[e].x, a1 = __uint_as_float(v[4 * e + 1]) * ww[e].y;
              const float a2 = __uint_as_float(v[4 * e + 2]) * ww[e].z, a3 = __uint_as_float(v[4 * e + 3]) * ww[e].w;
              if (p.is_fp16) {
                __half2 h0 = __floats2half2_rn(a0, a1), h1 = __floats2half2_rn(a2, a3);
                pk[2 * e] = *reinterpret_cast<uint32_t*>(&h0); pk[2 * e + 1] = *reinterpret_cast<uint32_t*>(&h1);
              } else {
                pk[2 * e] = pack_bf16x2(a0, a1); pk[2 * e + 1] = pack_bf16x2(a2, a3);
              }
            }
            tmem_st_x16(acc + c * 16, pk);
          };
          // (double-buffering the TMEM reads across chunks was measured: no gain, 32 more registers - profiles/r02_notes.md)
          for (int c = 0; c < nchunk; ++c) {
            uint32_t v[32];
            tmem_ld_x32(acc + c * 32, v);
            tmem_ld_wait();
            process(v, c);
          }
        } else {
          int j = 0, nb = w;                            // current key block and its end column
          float wv = row_ok ? Wrow[0] : 0.f;
          for (int c = 0; c < nchunk; ++c) {
            uint32_t v[32];
            tmem_ld_x32(acc + c * 32, v);
            tmem_ld_wait();
            uint32_t pk[16];
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
              float a2[2];
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const int col = c * 32 + e + u;
                if (col == nb) { ++j; nb += w; wv = (row_ok && j < M) ? Wrow[j] : 0.f; }
                const float pv = __uint_as_float(v[e + u]);
                if (j == blk_t) nl += pv;
                a2[u] = pv * wv;
              }
              if (p.is_fp16) { __half2 hh = __floats2half2_rn(a2[0], a2[1]); pk[e >> 1] = *reinterpret_cast<uint32_t*>(&hh); }
              else pk[e >> 1] = pack_bf16x2(a2[0], a2[1]);
            }
            tmem_st_x16(acc + c * 16, pk);
          }
        }
        tmem_st_wait();
        if (row_ok) nl_s[row] = nl;
        tc_fence_before();
        named_bar_sync(1 + t, 128);
        if (et == 0) mbar_arrive(&aready[t]);
        lap(1);   // masking epilogue
      }
      // ---- normaliser: mix equal in-block positions across blocks (needs the n_loc of BOTH row tiles)
      named_bar_sync(3, 256);
      float rden = 1.f;
      if (p.normalize && row_ok) {
        float den = p.eps;
        for (int jj = 0; jj < M; ++jj) den = fmaf(Wrow[jj * wstep], nl_s[jj * w + pos_t], den);
        rden = 1.0f / den;
      }
      lap(2);     // barrier with the other row tile + normaliser
      // ---- O -> 16 bit -> this unit's Q buffer (the scores MMAs have completed: pfull) -> one TMA store
      if (active) {
        mbar_wait(&ofull[t], ph);
        lap(3);   // wait for O = A V
        tc_fence_after();
        uint32_t pk[32];
        if (p.post_add != nullptr) {
          // fused additive term: 32 columns at a time, the row's 64-byte pieces of the add tensor read directly
          const uint16_t* arow = p.post_add + (long long)(g / p.H) * p.pa_sb + (long long)(g % p.H) * p.pa_sh +
                                 (long long)blk_t * p.pa_sm + (long long)pos_t * p.pa_sw;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t v[32], aw[16];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 a4 = make_uint4(0u, 0u, 0u, 0u);
              if (row_ok) a4 = __ldg(reinterpret_cast<const uint4*>(arow + hh * 32) + i);
              aw[4 * i] = a4.x; aw[4 * i + 1] = a4.y; aw[4 * i + 2] = a4.z; aw[4 * i + 3] = a4.w;
            }
            tmem_ld_x32(acc + 128 + hh * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              float2 a2;
              if (p.is_fp16) a2 = __half22float2(*reinterpret_cast<const __half2*>(&aw[e]));
              else a2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&aw[e]));
              const float x0 = fmaf(__uint_as_float(v[2 * e]), rden, a2.x), x1 = fmaf(__uint_as_float(v[2 * e + 1]), rden, a2.y);
              if (p.is_fp16) { __half2 h0 = __floats2half2_rn(x0, x1); pk[hh * 16 + e] = *reinterpret_cast<uint32_t*>(&h0); }
              else pk[hh * 16 + e] = pack_bf16x2(x0, x1);
            }
          }
        } else {
          uint32_t v[32], v2[32];
          tmem_ld_x32(acc + 128, v);
          tmem_ld_x32(acc + 160, v2);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float a = __uint_as_float(v[2 * e]) * rden, bq = __uint_as_float(v[2 * e + 1]) * rden;
            const float c2 = __uint_as_float(v2[2 * e]) * rden, d2 = __uint_as_float(v2[2 * e + 1]) * rden;
            if (p.is_fp16) {
              __half2 h0 = __floats2half2_rn(a, bq), h1 = __floats2half2_rn(c2, d2);
              pk[e] = *reinterpret_cast<uint32_t*>(&h0); pk[16 + e] = *reinterpret_cast<uint32_t*>(&h1);
            } else {
              pk[e] = pack_bf16x2(a, bq); pk[16 + e] = pack_bf16x2(c2, d2);
            }
          }
        }
        uint4* dst = reinterpret_cast<uint4*>(st + row * 128);
#pragma unroll
        for (int c = 0; c < 8; ++c)
          dst[c ^ (row & 7)] = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        fence_proxy_async_smem();
        tc_fence_before();
      }
      lap(4);     // readout epilogue
      named_bar_sync(4, 256);                        // both tiles staged, n_loc consumed, accumulators read
      if (threadIdx.x == 128) {
        mbar_arrive(tfree);
        tma_store_5d_hint(&p.tmO, st, 0, 0, 0, g % p.H, g / p.H, kEvictFirst);
        tma_store_commit();
        tma_store_wait_read<0>();                    // the Q buffer may be refilled
        mbar_arrive(&empty[s]);
      }
      if (threadIdx.x == 256) mbar_arrive(tfree);
      lap(5);     // store hand-off
    }
    if (threadIdx.x == 128) tma_store_wait_all<0>();
    if (prof_on) {
      for (int k = 0; k < 6; ++k) p.prof[(size_t)blockIdx.x * 16 + k] = (unsigned long long)pc[k];
      p.prof[(size_t)blockIdx.x * 16 + 6] = it;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace mhla
