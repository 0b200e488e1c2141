// Non-causal block-mixed MHLA forward (variants A/B) as ONE persistent, warp-specialised sm_100a kernel.
//
// Replaces the inline PyTorch operator of the reference:
//   mhla_dit/mhla/mhla.py:262-268 and mhla_videogen/diffusion/model/wan/mhla_utils.py:328-341.
//
// Work is cut into three kinds of items that flow through one TMA->smem ring, one tcgen05 issuer and one
// epilogue warpgroup (see DESIGN.md "Kernel"):
//   P1 (g, j)        S_j = K_j^T V_j  (+ ksum_j via an all-ones B operand, n_loc[j,t] = q_{j,t}.ksum_j)
//   P2 (g, it, ic)   [S~ | den] rows it*128.., cols ic*256.. = mix . [S | n_loc]    (GEMM over blocks; S is kept
//                    in 16-bit, mix and n_loc are split hi+lo so the product carries ~16 mantissa bits.  A TF32
//                    formulation is not possible: tcgen05 kind::tf32 returns zeros for an MN-major operand with
//                    the plain 128B swizzle - see profiles/r01_microtest_umma_layouts.log)
//   P3 (g, i)        O_i = (Q_i S~_i) / den_i
// Items of different (b,h) groups g are interleaved in a fixed global order (P1(s), P3(s-lag3), P2(s-lag2))
// so that the S / S~ / den workspace and the second read of Q are served from L2; cross-CTA dependencies
// are per-group arrival counters in global memory (release/acquire), all CTAs are co-resident.
#pragma once
#include <cuda.h>
#include "ptx.cuh"

namespace mhla {

constexpr int kStageBytes = 32768;
constexpr int kNumStages = 5;            // default ring depth (fused kernel, P2/P3 launches, causal kernel)
constexpr int kMaxStages = 6;            // P1-only launches with D = 64 trade staging for a sixth stage
constexpr int kStagingBytes = 16384;  // one staging slot: [128 rows][128 B] swizzle-128B tile
constexpr int kStagingPerWg = 2 * kStagingBytes;   // each epilogue warpgroup owns two slots = one hand-off of <= 2 chunks
constexpr int kThreads = 384;         // warp 0: TMA producer, 1: MMA issuer, 2: dependency poller (+TMEM alloc),
                                      // 3: store/signal, 4-7: epilogue warpgroup 0 (even items), 8-11: warpgroup 1 (odd items)
constexpr int kEpiThreads = 128;
constexpr int kTmemCols = 512;
constexpr int kAccCols = 256;         // two accumulator buffers of 256 columns
constexpr int kKsumCol = 128;         // ksum accumulator columns [128,144) inside a P1 buffer

constexpr int kSmemRing = 0;
constexpr int kSmemStaging = kStageBytes * kNumStages;
constexpr int kSmemOnes = kSmemStaging + 2 * kStagingPerWg;
constexpr int kSmemKsum = kSmemOnes + 512;
constexpr int kSmemBars = kSmemKsum + 1024;   // ksum: 128 floats per epilogue warpgroup
constexpr int kSmemTotal = kSmemBars + 256;
constexpr int kSmemAlloc = kSmemTotal + 1024;  // slack for manual 1024-byte alignment

struct alignas(64) BlockmixParams {
  CUtensorMap tmK, tmV, tmKn, tmQn, tmQr;   // rank-5 (d, w, M, H, B) views; Kn/Qn: un-roped (normaliser)
  CUtensorMap tmSst;                        // S store   : (Dv, Dk, G*M)      16-bit, box (64, Dk, 1)
  CUtensorMap tmSld;                        // S load    : (ncols, M, G)      16-bit, box (64, 64, 1)
  CUtensorMap tmW;                          // mix hi/lo : (Mp, M, 2)         16-bit, box (64, 128, 1)
  CUtensorMap tmStst;                       // S~ store  : (D*D, M, G)        16-bit, box (64, 128, 1)
  CUtensorMap tmDen;                        // den store : (2*wpad, M, G)     fp32, box (32, 128, 1)  (hi | lo parts)
  CUtensorMap tmStld;                       // S~ load   : (Dv, Dk, G*M)      bf16/fp16, box (64, Dk, 1)
  CUtensorMap tmO;                          // out       : rank-5 like q
  uint16_t* ws_S;                           // [G*M][ncols] 16-bit: S_j (Dk*Dv) | n_loc_j hi (wpad) | n_loc_j lo (wpad)
  const float* den;                         // [G*M][2*wpad]: mix . n_loc_hi | mix . n_loc_lo
  uint32_t* counters;                       // [2*G]: finished P1 items, finished P2 items per group
  int G, H, M, w, TW, nsub;
  int ncols, wpad;
  int n2_rows, n2_cols, n2_scols;           // P2 tile grid; first n2_scols column tiles are S columns
  int kslabs;                               // ceil(M / 64)
  int normalize, ropenorm, is_fp16;
  int mode;                                 // 0: fused; 1/2/3: only that phase (unfused debugging path)
  int lag2, lag3;
  float eps;
  unsigned long long* prof;                 // optional [gridDim][16] cycle counters (debug, tools/prof_roles.py)
  int ring_stages, slot_bytes;              // smem carve-up of this launch (see kernel prologue)
  int dep_mode;                             // tuning: 0 = 32-lane dependency warp, 1 = single polling lane
  int sig_mode;                             // tuning: 0 = deferred completion signals, 1 = drain after every item
};

// Event trace of CTA 0 (debug): trace[role][item][slot] = clock64, laid out behind the per-CTA counters.
__device__ __forceinline__ void trace_ev(const BlockmixParams& p, int role, uint32_t item, int slot) {
  if (p.prof != nullptr && blockIdx.x == 0 && item < 256)
    p.prof[(size_t)gridDim.x * 16 + ((size_t)role * 256 + item) * 4 + slot] = (unsigned long long)clock64();
}

// wait on an mbarrier and, when profiling, add the waited cycles to `acc`
__device__ __forceinline__ void mbar_wait_prof(uint64_t* bar, uint32_t parity, bool on, long long& acc) {
  if (!on) { mbar_wait(bar, parity); return; }
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  acc += clock64() - t0;
}

struct Item {
  int type, g, t;
};

struct Sched {
  int G, n1, n2, n3, lag2, lag3, mode, nsteps, stride;
  int s;
  long long off;
  __device__ void init(const BlockmixParams& p) {
    G = p.G; n1 = p.M; n2 = p.n2_rows * p.n2_cols; n3 = p.M;
    lag2 = p.lag2; lag3 = p.lag3; mode = p.mode;
    nsteps = G + (lag2 > lag3 ? lag2 : lag3);
    stride = gridDim.x; s = 0; off = blockIdx.x;
  }
  __device__ bool next(Item& it) {
    if (mode != 0) {
      const int n = mode == 1 ? n1 : (mode == 2 ? n2 : n3);
      if (off >= (long long)G * n) return false;
      it.type = mode; it.g = (int)(off / n); it.t = (int)(off % n);
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

struct Ring {
  int stage = 0;
  uint32_t phase = 0;
  int depth = kNumStages;
  __device__ Ring() {}
  __device__ explicit Ring(int d) : depth(d) {}
  __device__ __forceinline__ void advance(int n = 1) {
    stage += n;
    while (stage >= depth) { stage -= depth; phase ^= 1; }
  }
  __device__ __forceinline__ Ring at(int k) const { Ring r = *this; r.advance(k); return r; }
};

__device__ __forceinline__ void spin_until(const uint32_t* cnt, uint32_t target) {
  uint32_t spins = 0;
  while (ld_acquire_gpu(cnt) < target) {
    __nanosleep(64);
    if (++spins > (1u << 24)) { printf("mhla: dependency wait timed out (block %d)\n", blockIdx.x); __trap(); }
  }
}

// Number of ring stages an item occupies (identical in every role).
template <int D>
__device__ __forceinline__ int p1_stages(const BlockmixParams& p) {
  if constexpr (D == 64) return p.nsub * (1 + p.ropenorm) + p.normalize;
  else return p.nsub * (2 + p.ropenorm) + p.normalize * p.nsub;
}
template <int D>
__device__ __forceinline__ int p3_stages(const BlockmixParams& p) {
  if constexpr (D == 64) return p.nsub;
  else return 1 + p.nsub;
}

template <int D>
__global__ void __launch_bounds__(kThreads, 1) blockmix_kernel(const __grid_constant__ BlockmixParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem + kSmemRing;
  // smem carve-up: [ring: nst x 32 KB][staging: 2 warpgroups x 2 slots][ones][ksum][barriers]; ring + staging = 224 KB
  const int nst = p.ring_stages;
  const int slot_bytes = p.slot_bytes;          // one staging slot (16 KB; 8 KB in P1-only launches with D = 64)
  uint8_t* staging = smem + nst * kStageBytes;
  uint16_t* ones = reinterpret_cast<uint16_t*>(smem + kSmemOnes);
  float* ksum_s = reinterpret_cast<float*>(smem + kSmemKsum);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kSmemBars);
  uint64_t* empty = full + kMaxStages;
  uint64_t* tfull = empty + kMaxStages;
  uint64_t* tempty = tfull + 2;
  uint64_t* sfull = tempty + 2;    // staging buffer written by the epilogue warps   (epilogue -> store warp)
  uint64_t* sfree = sfull + 2;     // staging buffer read out by TMA                 (store warp -> epilogue)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sfree + 2);
  uint32_t* dep_count = tmem_slot + 1;   // dependencies confirmed by the dependency warp (warp 2), read by the producer

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile_bytes = p.TW * 128;  // one [TW rows][64 x 16-bit] swizzle-128B tile
  const uint32_t fmt16 = p.is_fp16 ? 0u : 1u;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); mbar_init(&sfull[i], 4); mbar_init(&sfree[i], 1);
    }
    fence_barrier_init();
    *dep_count = 0;
    const CUtensorMap* maps = &p.tmK;
    for (int i = 0; i < 12; ++i) tma_prefetch_desc(maps + i);
  }
  if (threadIdx.x < 256) ones[threadIdx.x] = p.is_fp16 ? 0x3C00 : 0x3F80;  // 1.0 in fp16 / bf16
  if (warp == 2) { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
  fence_proxy_async_smem();  // the all-ones tile is read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dependency_wait();   // PDL: everything above overlapped the tail of the previous kernel in the stream
  grid_launch_dependents(); // ... and the next kernel may start its own prologue as soon as SMs free up

  Sched sched; sched.init(p);
  Item it;

  if (warp == 0) {
    // ============================================================ TMA producer (one lane)
    if (elect_one()) {
      Ring r(nst);
      uint32_t ndep = 0;   // dependency-bearing items seen so far
      const bool prof_on = p.prof != nullptr;
      long long w_empty = 0, w_dep = 0;
      const long long t_begin = clock64();
      unsigned long long gt_begin;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_begin));
      uint32_t pitem = 0;
      auto wait_dependency = [&]() {
        // the dependency warp has already polled the group counter (global, ~1 us) - here it is a smem read
        uint32_t spins = 0;
        const long long t0 = clock64();
        while (ld_acquire_cta_shared(dep_count) <= ndep) {
          if (++spins > MHLA_SPIN_LIMIT) { printf("mhla: dep wait timed out (block %d)\n", blockIdx.x); __trap(); }
        }
        ++ndep;
        fence_proxy_async_all();
        w_dep += clock64() - t0;
        trace_ev(p, 0, pitem, 2);
      };
      while (sched.next(it)) {
        const int b = it.g / p.H, h = it.g % p.H;
        trace_ev(p, 0, pitem, 0);
        if (p.prof != nullptr && blockIdx.x == 0 && pitem < 256)
          p.prof[(size_t)gridDim.x * 16 + ((size_t)0 * 256 + pitem) * 4 + 3] = (unsigned long long)it.type;
        if (it.type == 1) {
          const int j = it.t;
          for (int sub = 0; sub < p.nsub; ++sub) {
            const int t0 = sub * p.TW;
            if constexpr (D == 64) {
              mbar_wait_prof(&empty[r.stage], r.phase ^ 1, prof_on, w_empty);
              uint8_t* st = ring + r.stage * kStageBytes;
              mbar_arrive_expect_tx(&full[r.stage], 2 * tile_bytes);
              tma_load_5d(st, &p.tmK, &full[r.stage], 0, t0, j, h, b, kEvictFirst);
              tma_load_5d(st + 16384, &p.tmV, &full[r.stage], 0, t0, j, h, b, kEvictFirst);
              r.advance();
            } else {
              for (int kv = 0; kv < 2; ++kv) {
                mbar_wait_prof(&empty[r.stage], r.phase ^ 1, prof_on, w_empty);
                uint8_t* st = ring + r.stage * kStageBytes;
                mbar_arrive_expect_tx(&full[r.stage], 2 * tile_bytes);
                const CUtensorMap* tm = kv ? &p.tmV : &p.tmK;
                tma_load_5d(st, tm, &full[r.stage], 0, t0, j, h, b, kEvictFirst);
                tma_load_5d(st + 16384, tm, &full[r.stage], 64, t0, j, h, b, kEvictFirst);
                r.advance();
              }
            }
            if (p.ropenorm) {
              mbar_wait_prof(&empty[r.stage], r.phase ^ 1, prof_on, w_empty);
              uint8_t* st = ring + r.stage * kStageBytes;
              mbar_arrive_expect_tx(&full[r.stage], (D / 64) * tile_bytes);
              tma_load_5d(st, &p.tmKn, &full[r.stage], 0, t0, j, h, b, kEvictFirst);
              if constexpr (D == 128) tma_load_5d(st + 16384, &p.tmKn, &full[r.stage], 64, t0, j, h, b, kEvictFirst);
              r.advance();
            }
          }
          if (p.normalize) {
            if constexpr (D == 64) {
              mbar_wait_prof(&empty[r.stage], r.phase ^ 1, prof_on, w_empty);
              uint8_t* st = ring + r.stage * kStageBytes;
              mbar_arrive_expect_tx(&full[r.stage], p.nsub * tile_bytes);
              for (int sub = 0; sub < p.nsub; ++sub)
                tma_load_5d(st + sub * tile_bytes, &p.tmQn, &full[r.stage], 0, sub * p.TW, j, h, b, kEvictLast);
              r.advance();
            } else {
              for (int sub = 0; sub < p.nsub; ++sub) {
                mbar_wait_prof(&empty[r.stage], r.phase ^ 1, prof_on, w_empty);
                uint8_t* st = ring + r.stage * kStageBytes;
                mbar_arrive_expect_tx(&full[r.stage], 2 * tile_bytes);
                tma_load_5d(st, &p.tmQn, &full[r.stage], 0, sub * p.TW, j, h, b, kEvictLast);
                tma_load_5d(st + 16384, &p.tmQn, &full[r.stage], 64, sub * p.TW, j, h, b, kEvictLast);
                r.advance();
              }
            }
          }
        } else if (it.type == 2) {
          if (p.mode == 0) wait_dependency();
          const int ti = it.t / p.n2_cols, tc = it.t % p.n2_cols;
          for (int slab = 0; slab < p.kslabs; ++slab) {
            // stage X: mix hi | mix lo, [128 i][64 j] each; stage Y: [64 j][256 cols] as 4 tiles of 64 columns
            mbar_wait_prof(&empty[r.stage], r.phase ^ 1, prof_on, w_empty);
            uint8_t* st = ring + r.stage * kStageBytes;
            mbar_arrive_expect_tx(&full[r.stage], 32768);
            tma_load_3d(st, &p.tmW, &full[r.stage], slab * 64, ti * 128, 0, kEvictLast);
            tma_load_3d(st + 16384, &p.tmW, &full[r.stage], slab * 64, ti * 128, 1, kEvictLast);
            r.advance();
            mbar_wait_prof(&empty[r.stage], r.phase ^ 1, prof_on, w_empty);
            st = ring + r.stage * kStageBytes;
            mbar_arrive_expect_tx(&full[r.stage], 32768);
            for (int n4 = 0; n4 < 4; ++n4)
              tma_load_3d(st + n4 * 8192, &p.tmSld, &full[r.stage], tc * 256 + n4 * 64, slab * 64, it.g, kEvictNormal);
            r.advance();
          }
        } else {
          if (p.mode == 0) wait_dependency();
          const int i = it.t;
          const CUtensorMap* tq = &p.tmQr;
          if constexpr (D == 64) {
            for (int sub = 0; sub < p.nsub; ++sub) {
              mbar_wait_prof(&empty[r.stage], r.phase ^ 1, prof_on, w_empty);
              uint8_t* st = ring + r.stage * kStageBytes;
              mbar_arrive_expect_tx(&full[r.stage], tile_bytes + (sub == 0 ? 8192 : 0));
              tma_load_5d(st, tq, &full[r.stage], 0, sub * p.TW, i, h, b, kEvictFirst);
              if (sub == 0) tma_load_3d(st + 16384, &p.tmStld, &full[r.stage], 0, 0, it.g * p.M + i, kEvictFirst);
              r.advance();
            }
          } else {
            mbar_wait_prof(&empty[r.stage], r.phase ^ 1, prof_on, w_empty);
            uint8_t* st = ring + r.stage * kStageBytes;
            mbar_arrive_expect_tx(&full[r.stage], 32768);
            tma_load_3d(st, &p.tmStld, &full[r.stage], 0, 0, it.g * p.M + i, kEvictFirst);
            tma_load_3d(st + 16384, &p.tmStld, &full[r.stage], 64, 0, it.g * p.M + i, kEvictFirst);
            r.advance();
            for (int sub = 0; sub < p.nsub; ++sub) {
              mbar_wait_prof(&empty[r.stage], r.phase ^ 1, prof_on, w_empty);
              st = ring + r.stage * kStageBytes;
              mbar_arrive_expect_tx(&full[r.stage], 2 * tile_bytes);
              tma_load_5d(st, tq, &full[r.stage], 0, sub * p.TW, i, h, b, kEvictFirst);
              tma_load_5d(st + 16384, tq, &full[r.stage], 64, sub * p.TW, i, h, b, kEvictFirst);
              r.advance();
            }
          }
        }
        trace_ev(p, 0, pitem, 1);
        ++pitem;
      }
      if (prof_on) {
        unsigned long long* pr = p.prof + (size_t)blockIdx.x * 16;
        pr[0] = (unsigned long long)w_empty; pr[1] = (unsigned long long)w_dep;
        pr[2] = (unsigned long long)(clock64() - t_begin);
        unsigned long long gt_end;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_end));
        pr[14] = gt_end - gt_begin;   // nanoseconds: pr[2] / pr[14] = SM clock in GHz
      }
    }
  } else if (warp == 1) {
    // ============================================================ tcgen05 issuer (one lane)
    if (elect_one()) {
      Ring r(nst);
      uint32_t nitem = 0;
      const bool prof_on = p.prof != nullptr;
      long long w_full = 0, w_tempty = 0;
      const uint32_t ring_addr = smem_u32(ring);
      const uint32_t ones_addr = smem_u32(ones);
      // all-ones B operand: no swizzle, MN-major, 2x2 core matrices of 128 B (LBO: K direction, SBO: N direction)
      const uint64_t desc_ones = make_smem_desc(ones_addr, 256, 128, kSwizzleNone);
      const uint32_t idesc_p1 = make_idesc(fmt16, 1, 1, D, D);
      const uint32_t idesc_p1_ones = make_idesc(fmt16, 1, 1, D, 16);
      const uint32_t idesc_p2 = make_idesc(fmt16, 0, 1, 128, 256);
      const uint32_t idesc_p3 = make_idesc(fmt16, 0, 1, 128, D);
      // This lane's instruction stream is latency-bound (one thread, dependent integer ops), so the issue loops carry
      // as few instructions as possible: descriptors are built once per stage and advanced by adding to the 14-bit
      // start-address field ((bytes >> 4); no carry can leave the field for addresses < 256 KB).
      const uint64_t tmpl_mn16k = make_smem_desc(0, 16384, 1024, kSwizzle128);   // MN-major, 64-channel tiles 16 KB apart
      const uint64_t tmpl_mn8k = make_smem_desc(0, 8192, 1024, kSwizzle128);     // MN-major, tiles 8 KB apart (P2 B operand)
      const uint64_t tmpl_k = make_smem_desc(0, 0, 1024, kSwizzle128);           // K-major
      auto dsc = [](uint64_t tmpl, uint32_t saddr) -> uint64_t { return tmpl | (uint64_t)((saddr & 0x3FFFF) >> 4); };
      while (sched.next(it)) {
        const uint32_t ab = nitem & 1, aphase = (nitem >> 1) & 1;
        const uint32_t acc = tmem_base + ab * kAccCols;
        trace_ev(p, 1, nitem, 0);
        mbar_wait_prof(&tempty[ab], aphase ^ 1, prof_on, w_tempty);
        tc_fence_after();
        trace_ev(p, 1, nitem, 1);
        if (it.type == 1) {
          const int ksteps = p.TW / 16;
          for (int sub = 0; sub < p.nsub; ++sub) {
            uint32_t a_addr, b_addr;
            Ring r0 = r;
            if constexpr (D == 64) {
              mbar_wait_prof(&full[r.stage], r.phase, prof_on, w_full);
              a_addr = ring_addr + r.stage * kStageBytes;
              b_addr = a_addr + 16384;
              r.advance();
            } else {
              mbar_wait_prof(&full[r.stage], r.phase, prof_on, w_full);
              a_addr = ring_addr + r.stage * kStageBytes;
              r.advance();
              mbar_wait_prof(&full[r.stage], r.phase, prof_on, w_full);
              b_addr = ring_addr + r.stage * kStageBytes;
              r.advance();
            }
            tc_fence_after();
            // MN-major, 128B swizzle: 8 token rows per atom (SBO = 1024 B), 64 channels per atom (LBO = 16 KB);
            // one k-step = 16 tokens = 2048 B = 128 descriptor units
            const uint64_t da0 = dsc(tmpl_mn16k, a_addr), db0 = dsc(tmpl_mn16k, b_addr);
            const uint32_t first = sub != 0;
            // ksum accumulates K^T . 1 with the un-roped K: the same tile as A (variant A) or its own stage (variant B with
            // the normaliser).  The narrow N=16 MMAs are always interleaved with the main ones - a back-to-back run of
            // them produced wrong sums on B200 (profiles/r01_bringup_notes.md).
            if (p.normalize && !p.ropenorm) {
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                if (ks < ksteps) {
                  mma_f16_ss(acc, da0 + ks * 128, db0 + ks * 128, idesc_p1, ks ? 1u : first);
                  mma_f16_ss(acc + kKsumCol, da0 + ks * 128, desc_ones, idesc_p1_ones, ks ? 1u : first);
                }
            } else {
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                if (ks < ksteps) mma_f16_ss(acc, da0 + ks * 128, db0 + ks * 128, idesc_p1, ks ? 1u : first);
            }
            mma_commit(&empty[r0.stage]);
            if constexpr (D == 128) mma_commit(&empty[r0.at(1).stage]);
            if (p.ropenorm) {
              // ksum of the un-roped K (variant B with the normaliser) from its own stage.  NOTE: issued as a plain
              // (not unrolled, descriptor rebuilt per step) loop on purpose - a full-speed back-to-back run of these
              // narrow N=16 accumulating MMAs gave wrong sums on B200 (profiles/r01_bringup_notes.md).
              mbar_wait_prof(&full[r.stage], r.phase, prof_on, w_full);
              tc_fence_after();
              const uint32_t n_addr = ring_addr + r.stage * kStageBytes;
#pragma unroll 1
              for (int ks = 0; ks < ksteps; ++ks) {
                const uint64_t da = make_smem_desc(n_addr + ks * 2048, 16384, 1024, kSwizzle128);
                mma_f16_ss(acc + kKsumCol, da, desc_ones, idesc_p1_ones, (sub | ks) != 0);
              }
              mma_commit(&empty[r.stage]);
              r.advance();
            }
          }
          mma_commit(&tfull[ab]);
          if (p.normalize) r.advance(D == 64 ? 1 : p.nsub);  // Q stages are consumed by the epilogue warps
        } else if (it.type == 2) {
          for (int slab = 0; slab < p.kslabs; ++slab) {
            mbar_wait_prof(&full[r.stage], r.phase, prof_on, w_full);
            const uint32_t a_addr = ring_addr + r.stage * kStageBytes;
            const int sa = r.stage;
            r.advance();
            mbar_wait_prof(&full[r.stage], r.phase, prof_on, w_full);
            tc_fence_after();
            const uint64_t dhi0 = dsc(tmpl_k, a_addr), dlo0 = dsc(tmpl_k, a_addr + 16384);   // K-major: 32 B per k-step
            const uint64_t db0 = dsc(tmpl_mn8k, ring_addr + r.stage * kStageBytes);           // MN-major: 2048 B per k-step
            const uint32_t first = slab != 0;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              mma_f16_ss(acc, dhi0 + ks * 2, db0 + ks * 128, idesc_p2, ks ? 1u : first);
              mma_f16_ss(acc, dlo0 + ks * 2, db0 + ks * 128, idesc_p2, 1u);
            }
            mma_commit(&empty[sa]);
            mma_commit(&empty[r.stage]);
            r.advance();
          }
          mma_commit(&tfull[ab]);
        } else {
          Ring r0 = r;
          uint32_t b_addr;
          if constexpr (D == 128) {
            mbar_wait_prof(&full[r.stage], r.phase, prof_on, w_full);
            b_addr = ring_addr + r.stage * kStageBytes;
            r.advance();
          }
          for (int sub = 0; sub < p.nsub; ++sub) {
            mbar_wait_prof(&full[r.stage], r.phase, prof_on, w_full);
            tc_fence_after();
            const uint32_t a_addr = ring_addr + r.stage * kStageBytes;
            if constexpr (D == 64) { if (sub == 0) b_addr = a_addr + 16384; }
            const uint64_t da0 = dsc(tmpl_k, a_addr), db0 = dsc(tmpl_mn16k, b_addr);
#pragma unroll
            for (int ks = 0; ks < D / 16; ++ks)   // K-major A: 4 k-steps (32 B each) per 64-channel tile, tiles 16 KB apart
              mma_f16_ss(acc + sub * 128, da0 + (ks >> 2) * 1024 + (ks & 3) * 2, db0 + ks * 128, idesc_p3, ks != 0);
            r.advance();
          }
          const int ns = p3_stages<D>(p);
          for (int k = 0; k < ns; ++k) mma_commit(&empty[r0.at(k).stage]);
          mma_commit(&tfull[ab]);
        }
        trace_ev(p, 1, nitem, 2);
        ++nitem;
      }
      if (prof_on) {
        unsigned long long* pr = p.prof + (size_t)blockIdx.x * 16;
        pr[3] = (unsigned long long)w_full; pr[4] = (unsigned long long)w_tempty;
      }
    }
  } else if (warp == 2) {
    // ============================================================ dependency warp (32 lanes)
    // Polls the per-group counters in global memory for the producer.  An L2 round trip under load costs microseconds,
    // so the 32 lanes watch the next 32 dependency-bearing items concurrently (lane L owns dependencies L, L+32, ...)
    // and the confirmed prefix is published through a monotonic counter in shared memory.
    if (p.mode == 0 && p.dep_mode == 1) {
      if (lane == 0) {
        uint32_t n = 0;
        while (sched.next(it)) {
          if (it.type == 1) continue;
          if (it.type == 2) spin_until(&p.counters[it.g], (uint32_t)p.M);
          else spin_until(&p.counters[p.G + it.g], (uint32_t)(p.n2_rows * p.n2_cols));
          st_release_cta_shared(dep_count, ++n);
        }
      }
    } else if (p.mode == 0) {
      uint32_t n = 0;            // dependencies confirmed so far (warp-uniform)
      uint32_t my_idx = lane;    // the dependency this lane is watching
      uint32_t seen = 0;         // dependency-bearing items this lane's private schedule walk has passed
      const uint32_t* my_ptr = nullptr;
      uint32_t my_target = 0;
      bool have = false;
      auto advance_to = [&](uint32_t idx) {
        have = false;
        while (sched.next(it)) {
          if (it.type == 1) continue;
          if (seen++ == idx) {
            my_ptr = (it.type == 2) ? &p.counters[it.g] : &p.counters[p.G + it.g];
            my_target = (it.type == 2) ? (uint32_t)p.M : (uint32_t)(p.n2_rows * p.n2_cols);
            have = true;
            break;
          }
        }
      };
      advance_to(my_idx);
      uint32_t idle = 0;
      while (true) {
        const bool ok = have ? (ld_acquire_gpu(my_ptr) >= my_target) : true;   // past-the-end lanes never block
        const uint32_t mask = __ballot_sync(0xffffffffu, ok);
        const uint32_t any_have = __ballot_sync(0xffffffffu, have);
        const uint32_t rot = n & 31;
        const uint32_t rmask = rot ? ((mask >> rot) | (mask << (32 - rot))) : mask;
        const uint32_t cnt = (rmask == 0xffffffffu) ? 32u : (uint32_t)(__ffs(~rmask) - 1);
        if (cnt > 0) {
          if (lane == 0) st_release_cta_shared(dep_count, n + cnt);
          if (((lane - rot) & 31) < cnt && have) { my_idx += 32; advance_to(my_idx); }
          n += cnt;
          idle = 0;
        } else {
          __nanosleep(128);
          if (++idle > (1u << 22)) { if (lane == 0) printf("mhla: dependency wait timed out (block %d)\n", blockIdx.x); __trap(); }
        }
        if (any_have == 0) break;
      }
    }
  } else if (warp >= 4) {
    // ============================================================ epilogue warpgroup (TMEM -> regs -> smem -> TMA)
    const int q4 = warp & 3;                 // TMEM sub-partition (lanes 32*q4 .. 32*q4+31)
    const int wg = (warp - 4) >> 2;          // epilogue warpgroup: handles the items whose accumulator buffer is `wg`
    const int et = threadIdx.x - 128 - wg * kEpiThreads;   // 0..127 within the warpgroup
    ksum_s += wg * 128;
    const uint32_t bar_base = 1 + wg * 4;    // named barrier ids of this warpgroup
    const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
    Ring r(nst);
    uint32_t nitem = 0;
    uint32_t nstore = 0;                     // staging buffer toggles per TMA-store chunk
    uint32_t v[32];
    const bool prof_on = p.prof != nullptr && et == 0 && wg == 0;
    long long w_tfull = 0, w_sfree = 0, w_q = 0, t_p1 = 0, t_p2 = 0, t_p3 = 0;
    uint32_t ndep_e = 0;                     // dependency-bearing (P2/P3) items seen so far

    // write one [rows][128 B] chunk row into the swizzle-128B staging tile
    auto stage_row = [&](uint8_t* buf, int row, const uint32_t* w32) {
      uint4* dst = reinterpret_cast<uint4*>(buf + row * 128);
#pragma unroll
      for (int c = 0; c < 8; ++c)
        dst[c ^ (row & 7)] = make_uint4(w32[4 * c], w32[4 * c + 1], w32[4 * c + 2], w32[4 * c + 3]);
    };
    // Staging hand-off with the store warp.  Warpgroup `wg` owns two 16 KB slots; the chunks of an item are written
    // alternately into them and handed over two at a time (one mbarrier round trip per pair - the round trip, not
    // the copy, is what costs ~1500 cycles).  cj = chunk index within the item, nch = chunks of the item.
    int cj = 0, nch = 0;
    auto staging_acquire = [&]() -> uint8_t* {
      if ((cj & 1) == 0) mbar_wait_prof(&sfree[wg], (nstore & 1) ^ 1, prof_on, w_sfree);   // first use passes immediately
      return staging + wg * 2 * slot_bytes + (cj & 1) * slot_bytes;
    };
    auto staging_publish = [&]() {
      fence_proxy_async_smem();                 // my rows -> visible to the async proxy (TMA)
      ++cj;
      if ((cj & 1) == 0 || cj == nch) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&sfull[wg]);   // one arrival per epilogue warp
        ++nstore;
      }
    };
    // load 64 fp32 accumulator columns, scale, round to the 16-bit I/O type: 32 packed words = one 128-byte row
    auto load_pack64 = [&](uint32_t taddr, float scale, uint32_t* pk) {
      uint32_t v2[32];
      tmem_ld_x32(taddr, v);
      tmem_ld_x32(taddr + 32, v2);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const float a = __uint_as_float(v[2 * e]) * scale, bq = __uint_as_float(v[2 * e + 1]) * scale;
        const float c2 = __uint_as_float(v2[2 * e]) * scale, d2 = __uint_as_float(v2[2 * e + 1]) * scale;
        if (p.is_fp16) {
          __half2 h0 = __floats2half2_rn(a, bq), h1 = __floats2half2_rn(c2, d2);
          pk[e] = *reinterpret_cast<uint32_t*>(&h0); pk[16 + e] = *reinterpret_cast<uint32_t*>(&h1);
        } else {
          pk[e] = pack_bf16x2(a, bq); pk[16 + e] = pack_bf16x2(c2, d2);
        }
      }
    };
    // x = hi + lo with hi, lo in the 16-bit I/O type
    auto split16 = [&](float x, uint16_t& hi, uint16_t& lo) {
      if (p.is_fp16) {
        const __half h = __float2half_rn(x);
        const __half l = __float2half_rn(x - __half2float(h));
        hi = __half_as_ushort(h); lo = __half_as_ushort(l);
      } else {
        const __nv_bfloat16 h = __float2bfloat16_rn(x);
        const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
        hi = __bfloat16_as_ushort(h); lo = __bfloat16_as_ushort(l);
      }
    };
    // q_t . ksum over one 64-channel swizzled tile row
    auto dot_row64 = [&](const uint8_t* tile, int rrow, const float* ks, float acc_n) -> float {
      const uint4* rowp = reinterpret_cast<const uint4*>(tile + rrow * 128);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 u = rowp[c ^ (rrow & 7)];
        const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 f;
          if (p.is_fp16) f = __half22float2(*reinterpret_cast<const __half2*>(&uw[e]));
          else f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&uw[e]));
          acc_n = fmaf(f.x, ks[c * 8 + e * 2], acc_n);
          acc_n = fmaf(f.y, ks[c * 8 + e * 2 + 1], acc_n);
        }
      }
      return acc_n;
    };

    while (sched.next(it)) {
      const uint32_t ab = nitem & 1, aphase = (nitem >> 1) & 1;
      const uint32_t acc = tmem_base + ab * kAccCols + lane_sel;
      if ((int)ab != wg) {   // the other warpgroup's item: only keep the ring / dependency bookkeeping in step
        r.advance(it.type == 1 ? p1_stages<D>(p) : (it.type == 2 ? 2 * p.kslabs : p3_stages<D>(p)));
        if (it.type != 1) ++ndep_e;
        ++nitem;
        continue;
      }
      const long long t_item = prof_on ? clock64() : 0;
      if (et == 0) trace_ev(p, 2, nitem, 0);
      cj = 0;
      if (it.type == 1) nch = (p.normalize ? 1 : 0) + D / 64;
      else if (it.type == 3) nch = p.nsub * (D / 64);
      else {
        const int tcx = it.t % p.n2_cols;
        nch = 4;
        if (tcx >= p.n2_scols) { const int rem = (2 * p.wpad - (tcx - p.n2_scols) * 256 + 31) / 32; nch = rem < 8 ? rem : 8; }
      }
      if (it.type == 1) {
        mbar_wait_prof(&tfull[ab], aphase, prof_on, w_tfull);
        if (et == 0) trace_ev(p, 2, nitem, 1);
        tc_fence_after();
        // rows of S live in TMEM lanes: D == 128 -> lane = row; D == 64 (M=64 MMA) -> row r in lane 32*(r/16)+r%16
        const bool row_ok = (D == 128) || (lane < 16);
        const int row = (D == 128) ? et : (q4 * 16 + (lane & 15));
        const int kvs = (D == 64 ? 1 : 2) + p.ropenorm;
        r.advance(p.nsub * kvs);
        if (p.normalize) {
          // ---- n_loc first: it frees the Q stage(s) of the ring early
          uint32_t ks;
          tmem_ld_x1(acc + kKsumCol, ks);
          tmem_ld_wait();
          if (row_ok) ksum_s[row] = __uint_as_float(ks);
          named_bar_sync(bar_base + 2, kEpiThreads);  // ksum_s complete
          uint16_t* nbuf = reinterpret_cast<uint16_t*>(staging_acquire());   // [hi: wpad][lo: wpad] -> one bulk store
          if constexpr (D == 64) {
            mbar_wait_prof(&full[r.stage], r.phase, prof_on, w_q);
            const uint8_t* qs = ring + r.stage * kStageBytes;
            for (int t = et; t < p.wpad; t += kEpiThreads) {
              uint16_t hi, lo;
              split16(dot_row64(qs, t, ksum_s, 0.f), hi, lo);
              nbuf[t] = hi;
              nbuf[p.wpad + t] = lo;
            }
            named_bar_sync(bar_base, kEpiThreads);
            if (et == 0) mbar_arrive(&empty[r.stage]);
            r.advance();
          } else {
            for (int sub = 0; sub < p.nsub; ++sub) {
              mbar_wait_prof(&full[r.stage], r.phase, prof_on, w_q);
              const uint8_t* qs = ring + r.stage * kStageBytes;
              if (et < p.TW) {
                const int t = sub * p.TW + et;
                float a = dot_row64(qs, et, ksum_s, 0.f);
                a = dot_row64(qs + 16384, et, ksum_s + 64, a);
                uint16_t hi, lo;
                split16(a, hi, lo);
                nbuf[t] = hi;
                nbuf[p.wpad + t] = lo;
              }
              named_bar_sync(bar_base, kEpiThreads);
              if (et == 0) mbar_arrive(&empty[r.stage]);
              r.advance();
            }
          }
          staging_publish();
        }
        for (int c = 0; c < D / 64; ++c) {
          uint32_t pk[32];
          load_pack64(acc + c * 64, 1.0f, pk);
          uint8_t* buf = staging_acquire();
          if (row_ok) stage_row(buf, row, pk);
          staging_publish();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[ab]);
      } else if (it.type == 2) {
        const int tc = it.t % p.n2_cols;
        mbar_wait_prof(&tfull[ab], aphase, prof_on, w_tfull);
        if (et == 0) trace_ev(p, 2, nitem, 1);
        tc_fence_after();
        if (tc < p.n2_scols) {
          for (int c = 0; c < 4; ++c) {
            uint32_t pk[32];
            load_pack64(acc + c * 64, 1.0f, pk);
            uint8_t* buf = staging_acquire();
            stage_row(buf, et, pk);
            staging_publish();
          }
        } else {
          for (int q = 0; q < 8; ++q) {
            if ((tc - p.n2_scols) * 256 + q * 32 >= 2 * p.wpad) break;   // uniform: nothing left in this tile
            tmem_ld_x32(acc + q * 32, v);
            tmem_ld_wait();
            uint8_t* buf = staging_acquire();
            stage_row(buf, et, v);
            staging_publish();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[ab]);
        r.advance(2 * p.kslabs);
      } else {
        const int i = it.t;
        float dsum[2] = {1.f, 1.f};
        if (p.normalize) {
          // den = mix.n_loc_hi + mix.n_loc_lo + eps was written by other CTAs.  Once the dependency warp has confirmed
          // this item's group (monotonic smem counter - NOT the ring barriers, whose phase may already have wrapped
          // because P3 stages are recycled by the MMA warp alone), den is visible: fetch it through L2 now and let the
          // latency overlap the wait for the accumulator.
          if (p.mode == 0) {
            if (lane == 0) {     // one poller per warp: 128 threads spinning on one smem word would starve the banks
              uint32_t spins = 0;
              while (ld_acquire_cta_shared(dep_count) <= ndep_e) {
                __nanosleep(32);
                if (++spins > MHLA_SPIN_LIMIT) { printf("mhla: epilogue dep wait timed out (block %d)\n", blockIdx.x); __trap(); }
              }
            }
            __syncwarp();
          }
          const float* dg = p.den + (size_t)(it.g * p.M + i) * (2 * p.wpad);
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            const int t = sub * p.TW + et;
            if (sub < p.nsub && et < p.TW && t < p.w) dsum[sub] = __ldcg(dg + t) + __ldcg(dg + p.wpad + t) + p.eps;
          }
        }
        mbar_wait_prof(&tfull[ab], aphase, prof_on, w_tfull);
        if (et == 0) trace_ev(p, 2, nitem, 1);
        tc_fence_after();
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          if (sub >= p.nsub) break;
          const float rden = p.normalize ? 1.0f / dsum[sub] : 1.0f;
          for (int c = 0; c < D / 64; ++c) {
            uint32_t pk[32];
            load_pack64(acc + sub * 128 + c * 64, rden, pk);
            uint8_t* buf = staging_acquire();
            stage_row(buf, et, pk);
            staging_publish();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[ab]);
        r.advance(p3_stages<D>(p));
      }
      if (it.type != 1) ++ndep_e;
      if (et == 0) trace_ev(p, 2, nitem, 2);
      if (prof_on) {
        const long long dt = clock64() - t_item;
        if (it.type == 1) t_p1 += dt; else if (it.type == 2) t_p2 += dt; else t_p3 += dt;
      }
      ++nitem;
    }
    if (prof_on) {
      unsigned long long* pr = p.prof + (size_t)blockIdx.x * 16;
      pr[5] = (unsigned long long)w_tfull; pr[6] = (unsigned long long)w_sfree; pr[7] = (unsigned long long)w_q;
      pr[8] = (unsigned long long)t_p1; pr[9] = (unsigned long long)t_p2; pr[10] = (unsigned long long)t_p3;
      pr[11] = nitem;
    }
  } else if (warp == 3) {
    // ============================================================ store / signal warp (one lane)
    // Issues every TMA store (so the bulk async-groups belong to this thread), recycles the staging buffers and
    // publishes the per-group dependency counters once an item's stores have fully completed - none of this sits on
    // the epilogue warps' critical path.
    if (elect_one()) {
      uint32_t k = 0;              // bulk groups committed so far
      const bool prof_on = p.prof != nullptr;
      long long w_sfull = 0, w_done = 0;
      // Completion signals are deferred instead of draining the store queue after every item: bulk groups retire in
      // order, so once kSigLag younger groups have been committed, `wait_group kSigLag` (normally already satisfied)
      // proves the item's bytes are in global memory.  Whenever the warp would idle it flushes everything pending, so
      // a signal never waits on work that (transitively) depends on it.
      constexpr int kSigLag = 1, kMaxPending = 8;
      uint32_t* pend_ptr[kMaxPending] = {};
      uint32_t pend_seq[kMaxPending] = {};
      int pend_head = 0, pend_n = 0;
      auto fire = [&](uint32_t upto_seq) {     // publish every pending signal whose last group index is <= upto_seq
        bool fenced = false;
        while (pend_n > 0 && pend_seq[pend_head] <= upto_seq) {
          if (!fenced) { fence_proxy_async_all(); __threadfence(); fenced = true; }
          red_release_gpu_add(pend_ptr[pend_head], 1u);
          pend_head = (pend_head + 1) % kMaxPending;
          --pend_n;
        }
      };
      uint32_t kw[2] = {0, 0};   // chunks taken from each warpgroup's staging buffer
      int cur = 0;               // warpgroup (= item parity) of the item being stored
      int unfreed = -1;          // warpgroup whose previous hand-off has been issued but not yet handed back
      auto drain_all = [&]() {
        const long long t0 = prof_on ? clock64() : 0;
        tma_store_wait_all<0>();
        if (unfreed >= 0) { mbar_arrive(&sfree[unfreed]); unfreed = -1; }
        fire(k);
        if (prof_on) w_done += clock64() - t0;
      };
      int cj = 0, nch = 0;       // chunk index within the item / chunks of the item (same rule as the epilogue)
      auto take = [&]() -> uint8_t* {
        if ((cj & 1) == 0) {
          if (pend_n > 0 && !mbar_try_wait(&sfull[cur], kw[cur] & 1)) drain_all();   // idle: flush the signals
          mbar_wait_prof(&sfull[cur], kw[cur] & 1, prof_on, w_sfull);
          ++kw[cur];
        }
        return staging + cur * 2 * slot_bytes + (cj & 1) * slot_bytes;
      };
      // Buffers are handed back lazily: the TMA store engine drains a 32 KB pair in ~2000 cycles, so blocking on every
      // read-out would make this lane the bottleneck.  After committing hand-off k we only wait until hand-off k-1 has
      // been read (`wait_group.read 1`) - unless the same warpgroup produces the next hand-off too (an item with more
      // than two chunks), in which case its slots must come back before it can continue.
      auto issued = [&]() {
        ++cj;
        if ((cj & 1) != 0 && cj != nch) return;   // second chunk of the pair still to come
        tma_store_commit();
        ++k;
        if (cj != nch) {                       // more hand-offs of this item follow from the same warpgroup
          tma_store_wait_read<0>();
          if (unfreed >= 0) { mbar_arrive(&sfree[unfreed]); unfreed = -1; }
          mbar_arrive(&sfree[cur]);
        } else {
          tma_store_wait_read<1>();            // everything but the newest group has been read out of smem
          if (unfreed >= 0) mbar_arrive(&sfree[unfreed]);
          unfreed = cur;
        }
        if (pend_n > 0 && k >= pend_seq[pend_head] + kSigLag) {
          tma_store_wait_all<kSigLag>();       // groups 1..k-kSigLag are complete
          fire(k - kSigLag);
        }
      };
      uint32_t sitem = 0;
      while (sched.next(it)) {
        cur = (int)(sitem & 1);
        cj = 0;
        if (it.type == 1) nch = (p.normalize ? 1 : 0) + D / 64;
        else if (it.type == 3) nch = p.nsub * (D / 64);
        else {
          const int tcx = it.t % p.n2_cols;
          nch = 4;
          if (tcx >= p.n2_scols) { const int rem = (2 * p.wpad - (tcx - p.n2_scols) * 256 + 31) / 32; nch = rem < 8 ? rem : 8; }
        }
        trace_ev(p, 3, sitem, 0);
        if (it.type == 1) {
          const int row = it.g * p.M + it.t;
          if (p.normalize) {
            uint8_t* buf = take();
            bulk_store_1d(p.ws_S + (size_t)row * p.ncols + D * D, buf, (uint32_t)(4 * p.wpad));
            issued();
          }
          for (int c = 0; c < D / 64; ++c) {
            uint8_t* buf = take();
            tma_store_3d(&p.tmSst, buf, c * 64, 0, row);
            issued();
          }
        } else if (it.type == 2) {
          const int ti = it.t / p.n2_cols, tc = it.t % p.n2_cols;
          if (tc < p.n2_scols) {
            for (int c = 0; c < 4; ++c) {
              uint8_t* buf = take();
              tma_store_3d(&p.tmStst, buf, tc * 256 + c * 64, ti * 128, it.g);
              issued();
            }
          } else {
            for (int q = 0; q < 8; ++q) {
              if ((tc - p.n2_scols) * 256 + q * 32 >= 2 * p.wpad) break;
              uint8_t* buf = take();
              tma_store_3d(&p.tmDen, buf, (tc - p.n2_scols) * 256 + q * 32, ti * 128, it.g);
              issued();
            }
          }
        } else {
          const int b = it.g / p.H, h = it.g % p.H;
          for (int sub = 0; sub < p.nsub; ++sub)
            for (int c = 0; c < D / 64; ++c) {
              uint8_t* buf = take();
              tma_store_5d(&p.tmO, buf, c * 64, sub * p.TW, it.t, h, b);
              issued();
            }
        }
        trace_ev(p, 3, sitem, 1);
        if (p.mode == 0 && it.type != 3 && p.sig_mode == 1) {
          tma_store_wait_all<0>();
          fence_proxy_async_all();
          __threadfence();
          red_release_gpu_add(&p.counters[(it.type == 1 ? 0 : p.G) + it.g], 1u);
        } else if (p.mode == 0 && it.type != 3) {
          if (pend_n == kMaxPending) drain_all();
          const int slot = (pend_head + pend_n) % kMaxPending;
          pend_ptr[slot] = &p.counters[(it.type == 1 ? 0 : p.G) + it.g];
          pend_seq[slot] = k;                  // the item's last group
          ++pend_n;
        }
        trace_ev(p, 3, sitem, 2);
        ++sitem;
      }
      drain_all();
      tma_store_wait_all<0>();
      if (prof_on) {
        unsigned long long* pr = p.prof + (size_t)blockIdx.x * 16;
        pr[12] = (unsigned long long)w_sfull; pr[13] = (unsigned long long)w_done;
      }
    }
  }

  // ---------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}

// Tiny prologue: split the fp32 mixing matrix into hi + lo 16-bit planes [2][M][Mp] (optionally keeping only the
// strictly-lower triangle and folding a scale, for the causal variant) and zero the dependency counters.
__global__ void prep_mix_kernel(const float* __restrict__ mix, long long ld, uint16_t* __restrict__ out, int M, int Mp,
                                int strict_lower, float scale, int is_fp16, uint32_t* counters, int ncounters) {
  grid_launch_dependents();   // the main kernel may start its prologue now; it waits (griddepcontrol.wait) for our results
  const int n = M * Mp;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
    const int i = idx / Mp, j = idx % Mp;
    float v = 0.f;
    if (j < M && (!strict_lower || j < i)) v = mix[(long long)i * ld + j] * scale;
    uint16_t hi, lo;
    if (is_fp16) {
      const __half h = __float2half_rn(v);
      hi = __half_as_ushort(h); lo = __half_as_ushort(__float2half_rn(v - __half2float(h)));
    } else {
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      hi = __bfloat16_as_ushort(h); lo = __bfloat16_as_ushort(__float2bfloat16_rn(v - __bfloat162float(h)));
    }
    out[idx] = hi;
    out[n + idx] = lo;
  }
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < ncounters; idx += gridDim.x * blockDim.x)
    counters[idx] = 0u;
}

}  // namespace mhla
