// Causal chunked MHLA forward (variant C) -- replaces naive_chunk_simple_mhla_fixed,
// mhla_nlp/fla/ops/mhla/naive.py:10-83:
//     o_i = K^-1/2 ( q_i . sum_{j<i} mm[i,j] S_j  +  mm[i,i] . tril(q_i k_i^T) v_i ),   S_j = k_j^T v_j,  chunk = 64.
// Same persistent, warp-specialised structure as blockmix_kernel.cuh (TMA ring -> tcgen05 issuer -> epilogue
// warpgroup, per-group dependency counters) with three item kinds:
//   P1 (g, j)       S_j = K_j^T V_j                                   -> 16-bit workspace
//   P2 (g, it, ic)  S~ = (scale * strict_lower(mm)) . S   (hi+lo split mixing matrix)   -> 16-bit workspace
//   P3 (g, i, vh)   P = q_i k_i^T (M=64 MMA) ; O = q_i S~_i[:, vh] ; epilogue masks P (s <= t) * scale * mm[i,i],
//                   rounds it to 16 bit into shared memory ; O += P_masked v_i[:, vh] ; store.
// V is processed in halves of <= 128 columns (vh) so that O (<=128 cols) and P (64 cols) share one TMEM buffer.
#pragma once
#include <string>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../include/mhla_b200.h"
#include "blockmix_kernel.cuh"

namespace mhla {

constexpr int kChunk = 64;
constexpr int kCausalThreads = 256;   // warps 0-3: producer / MMA / TMEM alloc / idle, 4-7: epilogue
constexpr int kCTile = kChunk * 128;  // one [64 tokens][64 channels] 16-bit swizzle-128B tile = 8 KB
constexpr int kPCol = 192;            // P accumulator columns [192,256) inside an accumulator buffer

struct alignas(64) CausalParams {
  CUtensorMap tmQ, tmK, tmV, tmO;  // rank-5 (d, 64, n, H, B) views of [B, T, H, d]
  CUtensorMap tmSst;               // S store : (DV, DK, G*n)   box (64, DK, 1)
  CUtensorMap tmSld;               // S load  : (DK*DV, n, G)   box (64, 64, 1)
  CUtensorMap tmW;                 // mix     : (Mp, n, 2)      box (64, 128, 1)
  CUtensorMap tmStst;              // S~ store: (DK*DV, n, G)   box (64, 128, 1)
  CUtensorMap tmStld;              // S~ load : (DV, DK, G*n)   box (64, DK, 1)
  const float* mm;                 // original fp32 mixing matrix (diagonal is read by the epilogue)
  long long mm_ld;
  uint32_t* counters;              // [2*G]
  int G, H, n;                     // as scheduled: with packing G = groups / pack, n = pack * n0
  int pack, n0;                    // `pack` consecutive (b,h) groups share one 128-row mixing tile (block-diagonal mm)
  int n2_rows, n2_cols, kslabs;
  int is_fp16, mode, lag2, lag3;
  float scale;
};

struct CausalSched {
  int G, n1, n2, n3, lag2, lag3, mode, nsteps, stride, s;
  long long off;
  __device__ void init(const CausalParams& p, int nvh) {
    G = p.G; n1 = p.n; n2 = p.n2_rows * p.n2_cols; n3 = p.n * nvh;
    lag2 = p.lag2; lag3 = p.lag3; mode = p.mode;
    nsteps = G + (lag2 > lag3 ? lag2 : lag3);
    stride = gridDim.x; s = 0; off = blockIdx.x;
  }
  __device__ bool next(Item& it) {
    if (mode != 0) {
      const int nn = mode == 1 ? n1 : (mode == 2 ? n2 : n3);
      if (off >= (long long)G * nn) return false;
      it.type = mode; it.g = (int)(off / nn); it.t = (int)(off % nn);
      off += stride;
      return true;
    }
    while (s < nsteps) {
      const int c1 = (s < G) ? n1 : 0;
      const int g3 = s - lag3, g2 = s - lag2;
      const int c3 = (g3 >= 0 && g3 < G) ? n3 : 0;
      const int c2 = (g2 >= 0 && g2 < G) ? n2 : 0;
      const int tot = c1 + c3 + c2;
      if (off >= tot) { off -= tot; ++s; continue; }
      if (off < c1) { it.type = 1; it.g = s; it.t = (int)off; }
      else if (off < c1 + c3) { it.type = 3; it.g = g3; it.t = (int)off - c1; }
      else { it.type = 2; it.g = g2; it.t = (int)off - c1 - c3; }
      off += stride;
      return true;
    }
    return false;
  }
};

template <int DK, int DV>
__global__ void __launch_bounds__(kCausalThreads, 1) causal_kernel(const __grid_constant__ CausalParams p) {
  constexpr int DVH = DV > 128 ? 128 : DV;  // V columns per P3 item
  constexpr int NVH = DV / DVH;
  constexpr bool kKVOneStage = (DK + DV) <= 256;  // K and V tiles of one chunk fit one 32 KB stage
  constexpr int kP1Stages = kKVOneStage ? 1 : 2;
  constexpr int kP3Stages = 3;                    // [q | k], [v half | P tile], [S~ half]

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem + kSmemRing;
  uint8_t* staging = smem + kSmemStaging;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kSmemBars);
  uint64_t* empty = full + kNumStages;
  uint64_t* tfull = empty + kNumStages;
  uint64_t* tempty = tfull + 2;
  uint64_t* pfull = tempty + 2;    // P accumulator ready (MMA -> epilogue)
  uint64_t* pready = pfull + 2;    // masked P tile in shared memory (epilogue -> MMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pready + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t fmt16 = p.is_fp16 ? 0u : 1u;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kNumStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); mbar_init(&pfull[i], 1); mbar_init(&pready[i], 1);
    }
    fence_barrier_init();
    const CUtensorMap* maps = &p.tmQ;
    for (int i = 0; i < 9; ++i) tma_prefetch_desc(maps + i);
  }
  if (warp == 2) { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  CausalSched sched; sched.init(p, NVH);
  Item it;

  if (warp == 0) {
    // ============================================================ TMA producer
    if (elect_one()) {
      Ring r;
      while (sched.next(it)) {
        // (packing: scheduled chunk c of scheduled group g is chunk c % n0 of real group g * pack + c / n0)
        const int cs = it.type == 3 ? it.t / NVH : it.t;
        const int gr = it.g * p.pack + cs / p.n0;
        const int b = gr / p.H, h = gr % p.H;
        if (it.type == 1) {
          const int j = it.t % p.n0;
          mbar_wait(&empty[r.stage], r.phase ^ 1);
          uint8_t* st = ring + r.stage * kStageBytes;
          mbar_arrive_expect_tx(&full[r.stage], (kKVOneStage ? (DK + DV) / 64 : DK / 64) * kCTile);
          for (int c = 0; c < DK / 64; ++c) tma_load_5d(st + c * kCTile, &p.tmK, &full[r.stage], c * 64, 0, j, h, b, kEvictNormal);
          if constexpr (!kKVOneStage) {
            r.advance();
            mbar_wait(&empty[r.stage], r.phase ^ 1);
            st = ring + r.stage * kStageBytes;
            mbar_arrive_expect_tx(&full[r.stage], (DV / 64) * kCTile);
            for (int c = 0; c < DV / 64; ++c) tma_load_5d(st + c * kCTile, &p.tmV, &full[r.stage], c * 64, 0, j, h, b, kEvictNormal);
          } else {
            for (int c = 0; c < DV / 64; ++c)
              tma_load_5d(st + (DK / 64 + c) * kCTile, &p.tmV, &full[r.stage], c * 64, 0, j, h, b, kEvictNormal);
          }
          r.advance();
        } else if (it.type == 2) {
          if (p.mode == 0) { spin_until(&p.counters[it.g], (uint32_t)p.n); fence_proxy_async_all(); }
          const int ti = it.t / p.n2_cols, tc = it.t % p.n2_cols;
          for (int slab = 0; slab < p.kslabs; ++slab) {
            mbar_wait(&empty[r.stage], r.phase ^ 1);
            uint8_t* st = ring + r.stage * kStageBytes;
            mbar_arrive_expect_tx(&full[r.stage], 32768);
            tma_load_3d(st, &p.tmW, &full[r.stage], slab * 64, ti * 128, 0, kEvictLast);
            tma_load_3d(st + 16384, &p.tmW, &full[r.stage], slab * 64, ti * 128, 1, kEvictLast);
            r.advance();
            mbar_wait(&empty[r.stage], r.phase ^ 1);
            st = ring + r.stage * kStageBytes;
            mbar_arrive_expect_tx(&full[r.stage], 32768);
            for (int n4 = 0; n4 < 4; ++n4)
              tma_load_3d(st + n4 * 8192, &p.tmSld, &full[r.stage], tc * 256 + n4 * 64, slab * 64, it.g, kEvictNormal);
            r.advance();
          }
        } else {
          if (p.mode == 0) {
            spin_until(&p.counters[p.G + it.g], (uint32_t)(p.n2_rows * p.n2_cols));
            fence_proxy_async_all();
          }
          const int is = it.t / NVH, vh = it.t % NVH;   // is: scheduled chunk (workspace row), i: chunk in its real group
          const int i = is % p.n0;
          // stage A: q tiles | k tiles
          mbar_wait(&empty[r.stage], r.phase ^ 1);
          uint8_t* st = ring + r.stage * kStageBytes;
          mbar_arrive_expect_tx(&full[r.stage], 2 * (DK / 64) * kCTile);
          for (int c = 0; c < DK / 64; ++c) {
            tma_load_5d(st + c * kCTile, &p.tmQ, &full[r.stage], c * 64, 0, i, h, b, kEvictFirst);
            tma_load_5d(st + (DK / 64 + c) * kCTile, &p.tmK, &full[r.stage], c * 64, 0, i, h, b, kEvictFirst);
          }
          r.advance();
          // stage B: v half (the masked P tile is written at +16 KB by the epilogue warps)
          mbar_wait(&empty[r.stage], r.phase ^ 1);
          st = ring + r.stage * kStageBytes;
          mbar_arrive_expect_tx(&full[r.stage], (DVH / 64) * kCTile);
          for (int c = 0; c < DVH / 64; ++c)
            tma_load_5d(st + c * kCTile, &p.tmV, &full[r.stage], vh * DVH + c * 64, 0, i, h, b, kEvictFirst);
          r.advance();
          // stage C: S~_i[:, vh half] as DVH/64 tiles of [DK rows][64 cols]
          mbar_wait(&empty[r.stage], r.phase ^ 1);
          st = ring + r.stage * kStageBytes;
          mbar_arrive_expect_tx(&full[r.stage], (DVH / 64) * DK * 128);
          for (int c = 0; c < DVH / 64; ++c)
            tma_load_3d(st + c * DK * 128, &p.tmStld, &full[r.stage], vh * DVH + c * 64, 0, it.g * p.n + is, kEvictFirst);
          r.advance();
        }
      }
    }
  } else if (warp == 1) {
    // ============================================================ tcgen05 issuer
    if (elect_one()) {
      Ring r;
      uint32_t nitem = 0;
      uint32_t pcnt[2] = {0, 0};   // P3 items seen per accumulator buffer (phase of pfull / pready)
      const uint32_t ring_addr = smem_u32(ring);
      const uint32_t idesc_p1 = make_idesc(fmt16, 1, 1, DK, DV);
      const uint32_t idesc_p2 = make_idesc(fmt16, 0, 1, 128, 256);
      const uint32_t idesc_qk = make_idesc(fmt16, 0, 0, 64, 64);
      const uint32_t idesc_qs = make_idesc(fmt16, 0, 1, 64, DVH);
      while (sched.next(it)) {
        const uint32_t ab = nitem & 1, aphase = (nitem >> 1) & 1;
        const uint32_t acc = tmem_base + ab * kAccCols;
        mbar_wait(&tempty[ab], aphase ^ 1);
        tc_fence_after();
        if (it.type == 1) {
          mbar_wait(&full[r.stage], r.phase);
          const int s0 = r.stage;
          const uint32_t a_addr = ring_addr + r.stage * kStageBytes;
          uint32_t b_addr = a_addr + (DK / 64) * kCTile;
          if constexpr (!kKVOneStage) {
            r.advance();
            mbar_wait(&full[r.stage], r.phase);
            b_addr = ring_addr + r.stage * kStageBytes;
          }
          tc_fence_after();
          for (int ks = 0; ks < kChunk / 16; ++ks) {
            const uint64_t da = make_smem_desc(a_addr + ks * 2048, kCTile, 1024, kSwizzle128);   // MN-major
            const uint64_t db = make_smem_desc(b_addr + ks * 2048, kCTile, 1024, kSwizzle128);
            mma_f16_ss(acc, da, db, idesc_p1, ks != 0);
          }
          mma_commit(&empty[s0]);
          if constexpr (!kKVOneStage) mma_commit(&empty[r.stage]);
          r.advance();
          mma_commit(&tfull[ab]);
        } else if (it.type == 2) {
          for (int slab = 0; slab < p.kslabs; ++slab) {
            mbar_wait(&full[r.stage], r.phase);
            const uint32_t a_addr = ring_addr + r.stage * kStageBytes;
            const int sa = r.stage;
            r.advance();
            mbar_wait(&full[r.stage], r.phase);
            tc_fence_after();
            const uint32_t b_addr = ring_addr + r.stage * kStageBytes;
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t dhi = make_smem_desc(a_addr + ks * 32, 0, 1024, kSwizzle128);
              const uint64_t dlo = make_smem_desc(a_addr + 16384 + ks * 32, 0, 1024, kSwizzle128);
              const uint64_t db = make_smem_desc(b_addr + ks * 2048, 8192, 1024, kSwizzle128);
              mma_f16_ss(acc, dhi, db, idesc_p2, (slab | ks) != 0);
              mma_f16_ss(acc, dlo, db, idesc_p2, 1u);
            }
            mma_commit(&empty[sa]);
            mma_commit(&empty[r.stage]);
            r.advance();
          }
          mma_commit(&tfull[ab]);
        } else {
          const Ring rA = r, rB = r.at(1), rC = r.at(2);
          mbar_wait(&full[rA.stage], rA.phase);
          tc_fence_after();
          const uint32_t q_addr = ring_addr + rA.stage * kStageBytes;
          const uint32_t k_addr = q_addr + (DK / 64) * kCTile;
          // (1) P = q k^T : both K-major, M = 64 tokens t, N = 64 tokens s
          for (int ks = 0; ks < DK / 16; ++ks) {
            const uint32_t off = (ks >> 2) * kCTile + (ks & 3) * 32;
            const uint64_t da = make_smem_desc(q_addr + off, 0, 1024, kSwizzle128);
            const uint64_t db = make_smem_desc(k_addr + off, 0, 1024, kSwizzle128);
            mma_f16_ss(acc + kPCol, da, db, idesc_qk, ks != 0);
          }
          mma_commit(&pfull[ab]);
          // (2) O = q S~_i[:, vh]
          mbar_wait(&full[rC.stage], rC.phase);
          tc_fence_after();
          const uint32_t s_addr = ring_addr + rC.stage * kStageBytes;
          for (int ks = 0; ks < DK / 16; ++ks) {
            const uint32_t off = (ks >> 2) * kCTile + (ks & 3) * 32;
            const uint64_t da = make_smem_desc(q_addr + off, 0, 1024, kSwizzle128);
            const uint64_t db = make_smem_desc(s_addr + ks * 2048, DK * 128, 1024, kSwizzle128);   // MN-major
            mma_f16_ss(acc, da, db, idesc_qs, ks != 0);
          }
          // (3) O += P_masked v_i[:, vh]   (P tile written by the epilogue warps into stage B + 16 KB)
          mbar_wait(&full[rB.stage], rB.phase);
          mbar_wait(&pready[ab], pcnt[ab] & 1);
          ++pcnt[ab];
          tc_fence_after();
          const uint32_t v_addr = ring_addr + rB.stage * kStageBytes;
          const uint32_t p_addr = v_addr + 16384;
          for (int ks = 0; ks < kChunk / 16; ++ks) {
            const uint64_t da = make_smem_desc(p_addr + ks * 32, 0, 1024, kSwizzle128);           // K-major [64 t][64 s]
            const uint64_t db = make_smem_desc(v_addr + ks * 2048, kCTile, 1024, kSwizzle128);    // MN-major
            mma_f16_ss(acc, da, db, idesc_qs, 1u);
          }
          mma_commit(&empty[rA.stage]);
          mma_commit(&empty[rB.stage]);
          mma_commit(&empty[rC.stage]);
          mma_commit(&tfull[ab]);
          r.advance(kP3Stages);
        }
        ++nitem;
      }
    }
  } else if (warp >= 4) {
    // ============================================================ epilogue warpgroup
    const int q4 = warp & 3;
    const int et = threadIdx.x - 128;
    const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
    Ring r;
    uint32_t nitem = 0, nstore = 0;
    uint32_t pcnt[2] = {0, 0};
    uint32_t v[32];

    auto stage_row = [&](uint8_t* buf, int row, const uint32_t* w32) {
      uint4* dst = reinterpret_cast<uint4*>(buf + row * 128);
#pragma unroll
      for (int c = 0; c < 8; ++c)
        dst[c ^ (row & 7)] = make_uint4(w32[4 * c], w32[4 * c + 1], w32[4 * c + 2], w32[4 * c + 3]);
    };
    auto staging_acquire = [&]() -> uint8_t* {
      if (et == 0) tma_store_wait_read<1>();
      named_bar_sync(1, kEpiThreads);
      return staging + (nstore & 1) * kStagingBytes;
    };
    auto staging_publish = [&]() {
      fence_proxy_async_smem();
      named_bar_sync(2, kEpiThreads);
      ++nstore;
    };
    auto pack2 = [&](float a, float bq) -> uint32_t {
      if (p.is_fp16) { __half2 hv = __floats2half2_rn(a, bq); return *reinterpret_cast<uint32_t*>(&hv); }
      return pack_bf16x2(a, bq);
    };
    auto load_pack64 = [&](uint32_t taddr, uint32_t* pk) {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        tmem_ld_x32(taddr + hh * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) pk[hh * 16 + e] = pack2(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]));
      }
    };
    // M = 64 accumulators: row r sits in TMEM lane 32*(r/16) + r%16
    const bool row64_ok = lane < 16;
    const int row64 = q4 * 16 + (lane & 15);

    while (sched.next(it)) {
      const uint32_t ab = nitem & 1, aphase = (nitem >> 1) & 1;
      const uint32_t acc = tmem_base + ab * kAccCols + lane_sel;
      if (it.type == 1) {
        mbar_wait(&tfull[ab], aphase);
        tc_fence_after();
        const bool row_ok = (DK == 128) || row64_ok;
        const int row = (DK == 128) ? et : row64;
        for (int c = 0; c < DV / 64; ++c) {
          uint32_t pk[32];
          load_pack64(acc + c * 64, pk);
          uint8_t* buf = staging_acquire();
          if (row_ok) stage_row(buf, row, pk);
          staging_publish();
          if (et == 0) {
            tma_store_3d(&p.tmSst, buf, c * 64, 0, it.g * p.n + it.t);
            tma_store_commit();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[ab]);
        r.advance(kP1Stages);
        if (p.mode == 0 && et == 0) {
          tma_store_wait_all<0>();
          fence_proxy_async_all();
          __threadfence();
          red_release_gpu_add(&p.counters[it.g], 1u);
        }
      } else if (it.type == 2) {
        const int ti = it.t / p.n2_cols, tc = it.t % p.n2_cols;
        mbar_wait(&tfull[ab], aphase);
        tc_fence_after();
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[32];
          load_pack64(acc + c * 64, pk);
          uint8_t* buf = staging_acquire();
          stage_row(buf, et, pk);
          staging_publish();
          if (et == 0) {
            tma_store_3d(&p.tmStst, buf, tc * 256 + c * 64, ti * 128, it.g);
            tma_store_commit();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[ab]);
        r.advance(2 * p.kslabs);
        if (p.mode == 0 && et == 0) {
          tma_store_wait_all<0>();
          fence_proxy_async_all();
          __threadfence();
          red_release_gpu_add(&p.counters[p.G + it.g], 1u);
        }
      } else {
        const int is = it.t / NVH, vh = it.t % NVH;
        const int i = is % p.n0;                          // chunk inside its real group
        const int gr = it.g * p.pack + is / p.n0;         // real (b,h) group
        const Ring rB = r.at(1);
        const float dscale = p.scale * __ldg(p.mm + (long long)i * p.mm_ld + i);
        // ---- masked P: TMEM -> registers -> 16-bit K-major swizzled tile in stage B (+16 KB)
        mbar_wait(&pfull[ab], pcnt[ab] & 1);
        ++pcnt[ab];
        tc_fence_after();
        uint32_t pk[32];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          tmem_ld_x32(acc + kPCol + hh * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int s0 = hh * 32 + 2 * e;
            const float a = (s0 <= row64) ? __uint_as_float(v[2 * e]) * dscale : 0.f;
            const float bq = (s0 + 1 <= row64) ? __uint_as_float(v[2 * e + 1]) * dscale : 0.f;
            pk[hh * 16 + e] = pack2(a, bq);
          }
        }
        // stage B must have landed before we write next to the v tiles?  No: the P tile lives at +16 KB, outside the
        // bytes TMA writes, and the stage was handed to this item by the producer (its previous user released it).
        uint8_t* ptile = ring + rB.stage * kStageBytes + 16384;
        if (row64_ok) stage_row(ptile, row64, pk);
        fence_proxy_async_smem();
        tc_fence_before();
        named_bar_sync(3, kEpiThreads);
        if (et == 0) mbar_arrive(&pready[ab]);
        // ---- O
        mbar_wait(&tfull[ab], aphase);
        tc_fence_after();
        for (int c = 0; c < DVH / 64; ++c) {
          load_pack64(acc + c * 64, pk);
          uint8_t* buf = staging_acquire();
          if (row64_ok) stage_row(buf, row64, pk);
          staging_publish();
          if (et == 0) {
            const int b = gr / p.H, h = gr % p.H;
            tma_store_5d(&p.tmO, buf, vh * DVH + c * 64, 0, i, h, b);
            tma_store_commit();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[ab]);
        r.advance(kP3Stages);
      }
      ++nitem;
    }
    if (et == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}

// ---------------------------------------------------------------------------------------------------- host side
struct CausalPlan {
  int G, n, Mp, n2_rows, n2_cols, kslabs;
  int pack, Gs, ns;   // scheduled groups / chunks per scheduled group (packing of consecutive (b,h) groups)
  size_t off_S, off_St, off_W, off_cnt, total;
};

inline int plan_causal(const mhla_causal_desc* d, CausalPlan* pl) {
  if (!d) return MHLA_ERR_INVALID_ARGUMENT;
  if (d->dtype != MHLA_BF16 && d->dtype != MHLA_FP16) return MHLA_ERR_INVALID_ARGUMENT;
  if (d->B < 1 || d->H < 1 || d->T < 1) return MHLA_ERR_UNSUPPORTED_SHAPE;
  if (d->chunk != kChunk) return MHLA_ERR_UNSUPPORTED_SHAPE;
  if (d->K != 64 && d->K != 128) return MHLA_ERR_UNSUPPORTED_SHAPE;
  if (d->V != 64 && d->V != 128 && d->V != 256) return MHLA_ERR_UNSUPPORTED_SHAPE;
  if (d->T % kChunk != 0) return MHLA_ERR_UNSUPPORTED_SHAPE;  // the host shim zero-pads (naive.py:46-51)
  pl->G = d->B * d->H;
  pl->n = d->T / kChunk;
  if (pl->n > d->L) return MHLA_ERR_INVALID_ARGUMENT;
  // few chunks per sequence: schedule `pack` consecutive (b,h) groups as one group with a block-diagonal mixing matrix
  pl->pack = 1;
  for (int pk = 128 / pl->n; pk >= 2; --pk)
    if (pl->G % pk == 0) { pl->pack = pk; break; }
  pl->Gs = pl->G / pl->pack;
  pl->ns = pl->n * pl->pack;
  pl->Mp = (pl->ns + 7) / 8 * 8;
  pl->n2_rows = (pl->ns + 127) / 128;
  pl->n2_cols = d->K * d->V / 256;
  pl->kslabs = (pl->ns + 63) / 64;
  const size_t Gn = (size_t)pl->G * pl->n, KV = (size_t)d->K * d->V;
  auto up = [](size_t x) { return (x + 1023) / 1024 * 1024; };
  size_t off = 0;
  pl->off_S = off;   off = up(off + Gn * KV * 2);
  pl->off_St = off;  off = up(off + Gn * KV * 2);
  pl->off_W = off;   off = up(off + (size_t)2 * pl->ns * pl->Mp * 2);
  pl->off_cnt = off; off = up(off + (size_t)2 * pl->G * 4);
  pl->total = off;
  return MHLA_OK;
}

inline size_t causal_workspace_bytes(const mhla_causal_desc* d) {
  CausalPlan pl;
  return plan_causal(d, &pl) == MHLA_OK ? pl.total : 0;
}

using CausalEncodeFn = bool (*)(CUtensorMap*, CUtensorMapDataType, int, void*, const uint64_t*, const uint64_t*,
                                const uint32_t*);

template <int DK, int DV>
inline int causal_launch(CausalParams& P, const CausalPlan& pl, int unfused, int num_sms, cudaStream_t stream,
                         int* launches) {
  auto kern = causal_kernel<DK, DV>;
  // the dynamic shared-memory opt-in is a per-device function attribute: track it per device ordinal
  static bool attr[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return MHLA_ERR_CUDA;
  if (!attr[dev]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemAlloc) != cudaSuccess)
      return MHLA_ERR_CUDA;
    attr[dev] = true;
  }
  constexpr int NVH = DV > 128 ? DV / 128 : 1;
  const long long n1 = pl.ns, n2 = (long long)pl.n2_rows * pl.n2_cols, n3 = (long long)pl.ns * NVH;
  if (unfused) {
    for (int mode = 1; mode <= (unfused >= 2 ? unfused - 1 : 3); ++mode) {   // unfused = 2 / 3: stop after phase 1 / 2 (timing)
      P.mode = mode;
      const long long items = (long long)pl.Gs * (mode == 1 ? n1 : (mode == 2 ? n2 : n3));
      kern<<<(int)(items < num_sms ? items : num_sms), kCausalThreads, kSmemAlloc, stream>>>(P);
      ++*launches;
    }
  } else {
    P.mode = 0;
    const long long items = (long long)pl.Gs * (n1 + n2 + n3);
    kern<<<(int)(items < num_sms ? items : num_sms), kCausalThreads, kSmemAlloc, stream>>>(P);
    ++*launches;
  }
  return MHLA_OK;
}

}  // namespace mhla
