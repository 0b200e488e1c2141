// Non-causal block-mixed MHLA forward (variants A/B) as ONE persistent, warp-specialised sm_100a kernel.
//
// Replaces the inline PyTorch operator of the reference:
//   mhla_dit/mhla/mhla.py:262-268 and mhla_videogen/diffusion/model/wan/mhla_utils.py:328-341.
//
// Work is cut into three kinds of items that flow through one TMA->smem ring, one tcgen05 issuer and two
// epilogue warpgroups:
//   P1 (g, j)        S_j = K_j^T V_j  (+ ksum_j via an all-ones B operand, n_loc[j,t] = q_{j,t}.ksum_j)
//   P2 (g, it, ic)   [S~ | den] rows it*128.., cols ic*256.. = mix . [S | n_loc]    (GEMM over blocks; S is kept
//                    in 16-bit, mix and n_loc are split hi+lo so the product carries ~16 mantissa bits.  A TF32
//                    formulation is not possible: tcgen05 kind::tf32 returns zeros for an MN-major operand with
//                    the plain 128B swizzle - see profiles/r01_microtest_umma_layouts.log)
//   P3 (g, i)        O_i = (Q_i S~_i) / den_i
// Items are handed out through global tickets and the ORDER in which a CTA runs them is decided at run time by a
// scheduler warp: a dependent item (P2 needs all P1 of its group, P3 all P2) is only enqueued once its group's arrival
// counter in global memory (release/acquire) says it is ready, so no role ever blocks on a dependency.  Default
// policy: ready P2 items first, then P1, the readout (P3) last; `policy 0` puts ready P3 items before new P1 items
// within a window of groups (keeps Q in L2, but stalls on the P1 -> P2 -> P3 latency chain).  The scheduler feeds the
// other roles through a small FIFO in shared memory; all CTAs are co-resident (grid <= #SMs, 1 CTA/SM).
#pragma once
#include <cuda.h>
#include <type_traits>
#include "ptx.cuh"

// Early hand-back of the accumulator buffer to the tcgen05 issuer (measured on the headline shape, profiles/r02c_ab_early.log:
// 142.6 -> 140.5 us): a P1 epilogue reads its whole accumulator (64 S columns + ksum, D = 64) into registers FIRST and
// releases the buffer before the n_loc dot products and the staging of S (MHLA_EARLY_P1); a readout epilogue releases it
// right after its last TMEM read, before the last tile is staged and stored (MHLA_EARLY_P3 = 2; = 1 holds both 64-column
// halves in registers and releases before any staging - slower, register pressure; = 0: release at the end of the item).
#ifndef MHLA_EARLY_P1
#define MHLA_EARLY_P1 1
#endif
#ifndef MHLA_EARLY_P3
#define MHLA_EARLY_P3 2
#endif
// 1: the first summary item of every CTA is fixed (no ticket round trip before the first loads); 2: only for D = 128
// (measured: Wan layer 119.3 -> 117.7 us, headline 141.0 -> 141.6 us, profiles/r02c_ab_2.log)
#ifndef MHLA_STATIC_FIRST
#define MHLA_STATIC_FIRST 2
#endif
#ifndef MHLA_EPI_POLL1
#define MHLA_EPI_POLL1 0
#endif
// D = 64 instantiations: FOUR accumulator buffers of 128 TMEM columns instead of two of 256 (a P1 item needs 64 + 16
// columns, a readout item 2 x 64), so the tcgen05 issuer may run up to four items ahead of the epilogues; mixing items
// then cover 128 instead of 256 columns of [S | n_loc] (the host's tile grid is refined inside the kernel).
#ifndef MHLA_ACC4
#define MHLA_ACC4 0
#endif

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
constexpr int kSmemFifo = kSmemBars + 256;    // scheduler -> roles item FIFO
constexpr int kFifoDepth = 64;
constexpr int kSmemTotal = kSmemFifo + kFifoDepth * 4;
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
  uint16_t* ws_St;                          // [G*M][Dk*Dv] 16-bit: S~_i
  const float* mix;                         // self_prep: the caller's fp32 mixing matrix [M][mix_ld]
  long long mix_ld;
  uint16_t* w_planes;                       // [2][M][Mp] hi | lo planes of the mixing matrix (I/O type)
  int Mp, self_prep;                        // self_prep: no prologue kernel - the CTAs split the matrix and the last one to
                                            // finish re-zeroes the control block (persistent, library-owned workspace)
  const float* rms_w;                       // optional fused output RMSNorm weight [D] (NULL: off)
  float rms_eps;
  const float* wscale;                      // [1]: power of two the mixing matrix was divided by (prep_mix_scaled_kernel)
  const float* den;                         // [G*M][2*wpad]: mix . n_loc_hi | mix . n_loc_lo
  uint32_t* counters;                       // [2*G]: finished P1 items, finished P2 items per group
  int G, H, M, w, TW, nsub;                 // G, M: as scheduled - with packing, G = groups / pack and M = pack * M0
  int pack, M0;                             // small M: `pack` consecutive (b,h) groups share one 128-row mixing tile (block-diagonal W)
  int ncols, wpad;
  int n2_rows, n2_cols, n2_scols;           // P2 tile grid; first n2_scols column tiles are S columns
  int kslabs;                               // ceil(M / 64)
  int normalize, ropenorm, is_fp16;
  int mode;                                 // 0: fused; 1/2/3: only that phase (unfused debugging path)
  int window;                               // fused mode: P1 may run this many groups ahead of the CTA's next P3 item
  int run_ahead;                            // fused mode: items the scheduler may enqueue ahead of the producer
  int np2;                                  // fused mode: CTAs [0, np2) run only P2 items (0: every CTA owns P2 items too)
  float eps;
  unsigned long long* prof;                 // optional [gridDim][16] cycle counters (debug, tools/prof_roles.py)
  int mix_hi_only;                          // bf16: S columns of the block mixing take the 8-bit hi plane only (MHLA_FLAG_FAST_MIX)
  int slots_per_wg;                         // staging slots per epilogue warpgroup (2, or 1 to buy another ring stage)
  int ring_stages, slot_bytes;              // smem carve-up of this launch (see kernel prologue)
  int q_hint;                               // 1: evict-first on the normaliser's Q loads when the readout comes much later
  int o_hint;                               // 1: evict-first L2 hint on the output stores
  int policy;                               // mode 0: 0 = ready P3 items before new P1 items (window), 1 = P3 items last
  int reverse3;                             // mode 3: walk the groups backwards
  int pf_dist;                              // L2 prefetch distance of the producer, in own streaming items (0: off)
  int trace_cta;                            // debug: CTA whose event trace is recorded
  int cnt_stride;                           // words between two dependency counters (padded to separate L2 sectors)
  // ---- 3-D block view (Wan, blockmix_kernel<D, true>): q,k,v,out are TOKEN-major [B, (F H W), heads, D]; block j =
  // (fbi, hbi, wbi) of a (fb, hb, wb) layout covers the (p1, p2, p3) sub-grid of tokens, in-block order (p1 p2 p3) as in
  // mhla_utils.py:317-326.  Rank-5 maps (d, W, H, B*F, heads) with box (64, p3, p2, a, 1) fetch a sub-tile of `a` frames
  // in ONE TMA box - the reference's rearrange copies (:317-326) and their inverse (:345-354) cost nothing.
  // tm3[t][0]: box of g3_aper frames, tm3[t][1]: box of the last sub-tile when p1 is not a multiple of g3_aper;
  // t = 0..5: K(numerator), V, K(normaliser), Q(normaliser), Q(numerator), out.
  CUtensorMap tm3[6][2];
  const void* g3_zero;                      // >= 2 KB of zeros (workspace): fills a tile up to a multiple of 16 token rows
  int g3_F, g3_hb, g3_wb, g3_p1, g3_p2, g3_p3, g3_aper, g3_tail;
  int g3_rows[2], g3_kpad[2];               // token rows of sub-tile s, and rounded up to the MMA k-step (16)
  // ---- fused post-ops of the readout epilogue (blockmix_kernel<D, G3D, true>):
  //   out = (o [/ den] [* rms * rms_w]) * silu(gate) + add       (mhla_utils.py:360-366: g_norm, SiLU gate, + lepe)
  // gate / add are 16-bit tensors laid out like `out` with their own element strides (3-D view: token stride in *_sw,
  // *_sm unused); NULL = that op is off.  Each epilogue thread owns one token row and reads its 128-byte pieces directly.
  const uint16_t* post_gate; const uint16_t* post_add;
  long long pg_sb, pg_sh, pg_sm, pg_sw, pa_sb, pa_sh, pa_sm, pa_sw;
  int g3_H, g3_W;                           // token grid (3-D view): rows of a frame, tokens of a row
};

// 1-D bulk copy global -> shared, completion counted on an mbarrier (bytes multiple of 16, 16-byte aligned addresses)
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// One [rows][64 channels] tile of tensor `which` (0 K, 1 V, 2 Kn, 3 Qn, 4 Qr) for sub-tile `sub` of block j of (b, h).
template <bool G3D>
__device__ __forceinline__ void load_tile(const BlockmixParams& p, uint8_t* dst, int which, uint64_t* bar, int c0, int sub,
                                          int j, int h, int b, uint64_t hint) {
  if constexpr (!G3D) {
    tma_load_5d(dst, &p.tmK + which, bar, c0, sub * p.TW, j, h, b, hint);
  } else {
    const int wbi = j % p.g3_wb, jj = j / p.g3_wb, hbi = jj % p.g3_hb, fbi = jj / p.g3_hb;
    const CUtensorMap* tm = &p.tm3[which][(sub == p.nsub - 1) ? p.g3_tail : 0];
    tma_load_5d(dst, tm, bar, c0, wbi * p.g3_p3, hbi * p.g3_p2, b * p.g3_F + fbi * p.g3_p1 + sub * p.g3_aper, h, hint);
    const int rows = p.g3_rows[sub], kp = p.g3_kpad[sub];
    if (kp > rows) bulk_load_1d(dst + rows * 128, p.g3_zero, (uint32_t)(kp - rows) * 128u, bar);   // zero token rows
  }
}
template <bool G3D>
__device__ __forceinline__ int tile_bytes_of(const BlockmixParams& p, int sub) {
  if constexpr (!G3D) return p.TW * 128;
  else return p.g3_kpad[sub] * 128;
}

// Event trace of CTA 0 (debug): trace[role][item][slot] = clock64, laid out behind the per-CTA counters.
__device__ __forceinline__ void trace_ev(const BlockmixParams& p, int role, uint32_t item, int slot) {
  if (p.prof != nullptr && (int)blockIdx.x == p.trace_cta && item < 256)
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

// Item FIFO entry: type [0,2) | g [2,16) | t [16,32); type 0 terminates the stream.
__device__ __forceinline__ uint32_t fifo_encode(int type, int g, int t) {
  return (uint32_t)type | ((uint32_t)g << 2) | ((uint32_t)t << 16);
}

// Consumer side of the item FIFO (one per role; every role sees the same item sequence).
struct ItemStream {
  const uint32_t* fifo;
  const uint32_t* published;
  uint32_t idx = 0;
  __device__ ItemStream(const uint32_t* f, const uint32_t* pub) : fifo(f), published(pub) {}
  __device__ __forceinline__ bool ready() const { return ld_acquire_cta_shared(published) > idx; }
  // single-lane roles
  // (the product build keeps the exact round-1 polling loops: a few more integer instructions in them cost the headline
  //  kernel 3 %, see ptx.cuh; -DMHLA_DIAG swaps in the time-bounded, record-writing variants)
  __device__ __forceinline__ bool next(Item& it) {
#ifdef MHLA_DIAG
    SpinGuard guard;
    while (!ready()) {
      __nanosleep(20);
      if (guard.expired()) report_stall(2, 0, idx);
    }
#else
    uint32_t spins = 0;
    while (!ready()) {
      __nanosleep(20);   // keep the poll off the shared-memory pipe the epilogue warps of this SM sub-partition use
      if (++spins > MHLA_SPIN_LIMIT) { printf("mhla: item stream stalled (block %d)\n", blockIdx.x); __trap(); }
    }
#endif
    return take(it);
  }
  // whole-warp roles: one lane polls, the warp then reads the entry together
  __device__ __forceinline__ bool next_warp(Item& it, int lane) {
    if (lane == 0) {
#ifdef MHLA_DIAG
      SpinGuard guard;
      while (!ready()) {
        __nanosleep(32);
        if (guard.expired()) report_stall(2, 8, idx);
      }
#else
      uint32_t spins = 0;
      while (!ready()) {
        __nanosleep(32);
        if (++spins > MHLA_SPIN_LIMIT) { printf("mhla: item stream stalled (block %d)\n", blockIdx.x); __trap(); }
      }
#endif
    }
    __syncwarp();
    return take(it);
  }
  __device__ __forceinline__ bool take(Item& it) {
    const uint32_t e = *reinterpret_cast<const volatile uint32_t*>(fifo + (idx & (kFifoDepth - 1)));
    ++idx;
    it.type = (int)(e & 3u); it.g = (int)((e >> 2) & 0x3FFFu); it.t = (int)(e >> 16);
    return it.type != 0;
  }
};

struct Ring {
  int stage = 0;
  uint32_t phase = 0;
  int depth = kNumStages;
  int base = 0;     // first physical stage of this ring (stages below it hold resident data)
  __device__ Ring() {}
  __device__ explicit Ring(int d, int b = 0) : depth(d - b), base(b) {}
  __device__ __forceinline__ int idx() const { return base + stage; }
  __device__ __forceinline__ void advance(int n = 1) {
    stage += n;
    while (stage >= depth) { stage -= depth; phase ^= 1; }
  }
  __device__ __forceinline__ Ring at(int k) const { Ring r = *this; r.advance(k); return r; }
};

__device__ __forceinline__ void spin_until(const uint32_t* cnt, uint32_t target) {
#ifdef MHLA_DIAG
  SpinGuard guard;
  while (ld_acquire_gpu(cnt) < target) {
    __nanosleep(64);
    if (guard.expired()) report_stall(6, target, ld_acquire_gpu(cnt));
  }
#else
  uint32_t spins = 0;
  while (ld_acquire_gpu(cnt) < target) {
    __nanosleep(64);
    if (++spins > (1u << 24)) { printf("mhla: dependency wait timed out (block %d)\n", blockIdx.x); __trap(); }
  }
#endif
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

template <int D, bool G3D = false, bool POST = false>
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
  uint64_t* tempty = tfull + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 4);
  uint32_t* q_published = tmem_slot + 1;   // items the scheduler warp (warp 2) has enqueued
  uint32_t* q_started = tmem_slot + 2;     // items the producer has picked up (throttles the scheduler's run-ahead)
  uint32_t* wg_done = tmem_slot + 3;       // [2]: workspace-producing items finished by each epilogue warpgroup
  float* wscale_s = reinterpret_cast<float*>(tmem_slot + 5);   // self_prep: power of two the mixing matrix was divided by
  uint32_t* last_flag = tmem_slot + 6;     // self_prep: this CTA is the last one to finish
  uint32_t* fifo = reinterpret_cast<uint32_t*>(smem + kSmemFifo);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile_bytes = p.TW * 128;  // one [TW rows][64 x 16-bit] swizzle-128B tile
  const uint32_t fmt16 = p.is_fp16 ? 0u : 1u;
  // Dedicated block-mixing CTA of the fused kernel: the mixing matrix (hi | lo, one ring stage per 64-block slab) stays
  // resident in ring stages [0, kslabs) and only the S / n_loc tiles stream through the remaining stages.
  // in-kernel dependencies through the per-group counters: 0 = all phases, 4 = P1 + P2 only, 5 = P3 only, started by PDL
  // while the mode-4 grid is still draining (no griddepcontrol.wait: the counters carry the dependency)
  const bool dynamic = p.mode == 0 || p.mode == 4 || p.mode == 5;
  const bool wres = dynamic && (int)blockIdx.x < p.np2 && p.n2_rows == 1 && p.kslabs <= 2;
  // accumulator buffers (see MHLA_ACC4): number, width, column of ksum inside a P1 buffer, column stride between the two
  // sub-tiles of a readout item, width of a mixing item and the mixing tile grid that follows from it
  constexpr bool kAcc4 = (D == 64) && (MHLA_ACC4 != 0);
  constexpr uint32_t kNB = kAcc4 ? 4 : 2, kACols = kAcc4 ? 128 : kAccCols, kKsC = kAcc4 ? 64 : kKsumCol, kSubC = kAcc4 ? 64 : 128;
  constexpr int kPN = kAcc4 ? 128 : 256;
  const int n2_cols = p.n2_cols * (256 / kPN), n2_scols = p.n2_scols * (256 / kPN);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&tfull[i], 2); mbar_init(&tempty[i], 4);   // tfull: tcgen05.commit + the issuer's own arrive (see below)
    }
    fence_barrier_init();
    *q_published = 0;
    *q_started = 0;
    wg_done[0] = 0; wg_done[1] = 0;
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
  if (p.mode != 5) grid_dependency_wait();   // PDL: everything above overlapped the tail of the previous kernel in the stream
  grid_launch_dependents(); // ... and the next kernel may start its own prologue as soon as SMs free up
  // debug timeline: [mode][cta] = {globaltimer at start of work, at end}, behind the per-CTA counters and the CTA trace
  unsigned long long* const tl = p.prof == nullptr ? nullptr
      : p.prof + (size_t)gridDim.x * 16 + 4 * 256 * 4 + ((size_t)p.mode * 148 + blockIdx.x) * 4;
  if (tl != nullptr && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); tl[0] = t; }

  ItemStream sched(fifo, q_published);
  Item it;

  if (warp == 0) {
    // ============================================================ TMA producer (one lane)
    if (elect_one()) {
      Ring r(nst, wres ? p.kslabs : 0);
      const bool prof_on = p.prof != nullptr;
      long long w_empty = 0, w_dep = 0;
      const long long t_begin = clock64();
      unsigned long long gt_begin;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_begin));
      uint32_t pitem = 0;
      // A dependent item is only enqueued after the scheduler lane has acquired its group's counter; the FIFO hand-off
      // (release/acquire in shared memory) extends that to this lane, the proxy fence to the TMA loads issued below.
      auto wait_dependency = [&]() { fence_proxy_async_all(); };
      // Q tiles of the normaliser: worth keeping in L2 only if the readout follows within a few groups (policy 0)
      const uint64_t q_hint = (p.mode == 0 && p.policy == 0) ? kEvictLast : (p.q_hint ? kEvictFirst : kEvictNormal);
      // The streaming items a CTA owns are a fixed arithmetic sequence of linear block indices, so the producer can pull
      // the tiles of the item `pf_dist` places further down its own list into L2 while it fills shared memory for the
      // current one: the shared-memory fill then sees L2 latency instead of HBM latency, and the bytes in flight towards
      // HBM are no longer bounded by the ring.
      const long long n1tot_p = (long long)p.G * p.M;
      const long long nct13_p = (dynamic && p.np2 > 0) ? (long long)gridDim.x - p.np2 : (long long)gridDim.x;
      auto prefetch_block = [&](long long idx, bool kv, bool q_norm, bool q_read) {
        if (p.pf_dist <= 0) return;
        idx += (long long)p.pf_dist * nct13_p;
        if (idx >= n1tot_p) return;
        const int g_ = (int)(idx / p.M0), j_ = (int)(idx % p.M0);   // linear block index = real group * M0 + block
        const int b_ = g_ / p.H, h_ = g_ % p.H;
        for (int sub = 0; sub < p.nsub; ++sub)
          for (int c0 = 0; c0 < D; c0 += 64) {
            if (kv) {
              tma_prefetch_5d(&p.tmK, c0, sub * p.TW, j_, h_, b_);
              tma_prefetch_5d(&p.tmV, c0, sub * p.TW, j_, h_, b_);
              if (p.ropenorm) tma_prefetch_5d(&p.tmKn, c0, sub * p.TW, j_, h_, b_);
            }
            if (q_norm) tma_prefetch_5d(&p.tmQn, c0, sub * p.TW, j_, h_, b_);
            if (q_read) tma_prefetch_5d(&p.tmQr, c0, sub * p.TW, j_, h_, b_);
          }
      };
      if (wres) {
        if (p.self_prep) {   // the planes are being written by the CTAs of this very launch: wait until all have announced
          spin_until(p.counters + (size_t)2 * p.G * p.cnt_stride + 64, gridDim.x);
          fence_proxy_async_all();
        }
        for (int slab = 0; slab < p.kslabs; ++slab) {
          uint8_t* st = ring + slab * kStageBytes;
          mbar_arrive_expect_tx(&full[slab], 32768);
          tma_load_3d(st, &p.tmW, &full[slab], slab * 64, 0, 0, kEvictLast);
          tma_load_3d(st + 16384, &p.tmW, &full[slab], slab * 64, 0, 1, kEvictLast);
        }
      }
      while (true) {
        const long long tq0 = prof_on ? clock64() : 0;
        const bool more = sched.next(it);
        if (prof_on) w_dep += clock64() - tq0;      // time spent waiting for the scheduler (nothing ready)
        if (!more) break;
        st_release_cta_shared(q_started, sched.idx);
        // (packing: item (g, t) is block t % M0 of real group g * pack + t / M0)
        const int gr = it.g * p.pack + it.t / p.M0;
        const int b = gr / p.H, h = gr % p.H;
        trace_ev(p, 0, pitem, 0);
        if (p.prof != nullptr && (int)blockIdx.x == p.trace_cta && pitem < 256)
          p.prof[(size_t)gridDim.x * 16 + ((size_t)0 * 256 + pitem) * 4 + 3] = (unsigned long long)it.type;
        if (it.type == 1) {
          const int j = it.t % p.M0;
          for (int sub = 0; sub < p.nsub; ++sub) {
            const int t0 = sub * p.TW;
            const int tb = tile_bytes_of<G3D>(p, sub);
            (void)t0;
            if constexpr (D == 64) {
              mbar_wait_prof(&empty[r.idx()], r.phase ^ 1, prof_on, w_empty);
              uint8_t* st = ring + r.idx() * kStageBytes;
              mbar_arrive_expect_tx(&full[r.idx()], 2 * tb);
              load_tile<G3D>(p, st, 0, &full[r.idx()], 0, sub, j, h, b, kEvictFirst);
              load_tile<G3D>(p, st + 16384, 1, &full[r.idx()], 0, sub, j, h, b, kEvictFirst);
              r.advance();
            } else {
              for (int kv = 0; kv < 2; ++kv) {
                mbar_wait_prof(&empty[r.idx()], r.phase ^ 1, prof_on, w_empty);
                uint8_t* st = ring + r.idx() * kStageBytes;
                mbar_arrive_expect_tx(&full[r.idx()], 2 * tb);
                load_tile<G3D>(p, st, kv, &full[r.idx()], 0, sub, j, h, b, kEvictFirst);
                load_tile<G3D>(p, st + 16384, kv, &full[r.idx()], 64, sub, j, h, b, kEvictFirst);
                r.advance();
              }
            }
            if (p.ropenorm) {
              mbar_wait_prof(&empty[r.idx()], r.phase ^ 1, prof_on, w_empty);
              uint8_t* st = ring + r.idx() * kStageBytes;
              mbar_arrive_expect_tx(&full[r.idx()], (D / 64) * tb);
              load_tile<G3D>(p, st, 2, &full[r.idx()], 0, sub, j, h, b, kEvictFirst);
              if constexpr (D == 128) load_tile<G3D>(p, st + 16384, 2, &full[r.idx()], 64, sub, j, h, b, kEvictFirst);
              r.advance();
            }
          }
          if (p.normalize) {
            if constexpr (D == 64) {
              mbar_wait_prof(&empty[r.idx()], r.phase ^ 1, prof_on, w_empty);
              uint8_t* st = ring + r.idx() * kStageBytes;
              int qbytes = 0;
              for (int sub = 0; sub < p.nsub; ++sub) qbytes += tile_bytes_of<G3D>(p, sub);
              mbar_arrive_expect_tx(&full[r.idx()], qbytes);
              for (int sub = 0; sub < p.nsub; ++sub)
                load_tile<G3D>(p, st + sub * tile_bytes, 3, &full[r.idx()], 0, sub, j, h, b, q_hint);
              r.advance();
            } else {
              for (int sub = 0; sub < p.nsub; ++sub) {
                mbar_wait_prof(&empty[r.idx()], r.phase ^ 1, prof_on, w_empty);
                uint8_t* st = ring + r.idx() * kStageBytes;
                mbar_arrive_expect_tx(&full[r.idx()], 2 * tile_bytes_of<G3D>(p, sub));
                load_tile<G3D>(p, st, 3, &full[r.idx()], 0, sub, j, h, b, q_hint);
                load_tile<G3D>(p, st + 16384, 3, &full[r.idx()], 64, sub, j, h, b, q_hint);
                r.advance();
              }
            }
          }
          // (after this item's own loads: the TMA unit serves its queue in order)
          prefetch_block((long long)it.g * p.M + it.t, true, p.normalize != 0, false);
        } else if (it.type == 2) {
          if (dynamic) wait_dependency();
          const int ti = it.t / n2_cols, tc = it.t % n2_cols;
          for (int slab = 0; slab < p.kslabs; ++slab) {
            // stage X: mix hi | mix lo, [128 i][64 j] each; stage Y: [64 j][256 cols] as 4 tiles of 64 columns
            uint8_t* st;
            if (!wres) {
              mbar_wait_prof(&empty[r.idx()], r.phase ^ 1, prof_on, w_empty);
              st = ring + r.idx() * kStageBytes;
              const bool need_lo = !(p.is_fp16 || p.mix_hi_only) || tc >= n2_scols;
              mbar_arrive_expect_tx(&full[r.idx()], need_lo ? 32768 : 16384);
              tma_load_3d(st, &p.tmW, &full[r.idx()], slab * 64, ti * 128, 0, kEvictLast);
              if (need_lo) tma_load_3d(st + 16384, &p.tmW, &full[r.idx()], slab * 64, ti * 128, 1, kEvictLast);
              r.advance();
            }
            mbar_wait_prof(&empty[r.idx()], r.phase ^ 1, prof_on, w_empty);
            st = ring + r.idx() * kStageBytes;
            mbar_arrive_expect_tx(&full[r.idx()], (kPN / 64) * 8192);
            for (int n4 = 0; n4 < kPN / 64; ++n4)
              tma_load_3d(st + n4 * 8192, &p.tmSld, &full[r.idx()], tc * kPN + n4 * 64, slab * 64, it.g, kEvictNormal);
            r.advance();
          }
        } else {
          if (dynamic) wait_dependency();
          const int i = it.t % p.M0;                 // block inside its real group (tensor coordinate)
          const int irow = it.g * p.M + it.t;        // row of the block in the S~ workspace
          if constexpr (D == 64) {
            for (int sub = 0; sub < p.nsub; ++sub) {
              mbar_wait_prof(&empty[r.idx()], r.phase ^ 1, prof_on, w_empty);
              uint8_t* st = ring + r.idx() * kStageBytes;
              mbar_arrive_expect_tx(&full[r.idx()], tile_bytes_of<G3D>(p, sub) + (sub == 0 ? 8192 : 0));
              load_tile<G3D>(p, st, 4, &full[r.idx()], 0, sub, i, h, b, kEvictFirst);
              if (sub == 0) tma_load_3d(st + 16384, &p.tmStld, &full[r.idx()], 0, 0, irow, kEvictFirst);
              r.advance();
            }
          } else {
            mbar_wait_prof(&empty[r.idx()], r.phase ^ 1, prof_on, w_empty);
            uint8_t* st = ring + r.idx() * kStageBytes;
            mbar_arrive_expect_tx(&full[r.idx()], 32768);
            tma_load_3d(st, &p.tmStld, &full[r.idx()], 0, 0, irow, kEvictFirst);
            tma_load_3d(st + 16384, &p.tmStld, &full[r.idx()], 64, 0, irow, kEvictFirst);
            r.advance();
            for (int sub = 0; sub < p.nsub; ++sub) {
              mbar_wait_prof(&empty[r.idx()], r.phase ^ 1, prof_on, w_empty);
              st = ring + r.idx() * kStageBytes;
              mbar_arrive_expect_tx(&full[r.idx()], 2 * tile_bytes_of<G3D>(p, sub));
              load_tile<G3D>(p, st, 4, &full[r.idx()], 0, sub, i, h, b, kEvictFirst);
              load_tile<G3D>(p, st + 16384, 4, &full[r.idx()], 64, sub, i, h, b, kEvictFirst);
              r.advance();
            }
          }
          // the readout's Q tile has not been read before unless the normaliser pulled the same tensor through L2
          if (!p.normalize || p.ropenorm || p.mode != 0) prefetch_block((long long)it.g * p.M + it.t, false, false, true);
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
      Ring r(nst, wres ? p.kslabs : 0);
      uint32_t nitem = 0;
      const bool prof_on = p.prof != nullptr;
      long long w_full = 0, w_tempty = 0;
      const uint32_t ring_addr = smem_u32(ring);
      const uint32_t ones_addr = smem_u32(ones);
      // all-ones B operand: no swizzle, MN-major, 2x2 core matrices of 128 B (LBO: K direction, SBO: N direction)
      const uint64_t desc_ones = make_smem_desc(ones_addr, 256, 128, kSwizzleNone);
      const uint32_t idesc_p1 = make_idesc(fmt16, 1, 1, D, D);
      const uint32_t idesc_p1_ones = make_idesc(fmt16, 1, 1, D, 16);
      const uint32_t idesc_p2 = make_idesc(fmt16, 0, 1, 128, kPN);
      const uint32_t idesc_p3 = make_idesc(fmt16, 0, 1, 128, D);
      // This lane's instruction stream is latency-bound (one thread, dependent integer ops), so the issue loops carry
      // as few instructions as possible: descriptors are built once per stage and advanced by adding to the 14-bit
      // start-address field ((bytes >> 4); no carry can leave the field for addresses < 256 KB).
      const uint64_t tmpl_mn16k = make_smem_desc(0, 16384, 1024, kSwizzle128);   // MN-major, 64-channel tiles 16 KB apart
      const uint64_t tmpl_mn8k = make_smem_desc(0, 8192, 1024, kSwizzle128);     // MN-major, tiles 8 KB apart (P2 B operand)
      const uint64_t tmpl_k = make_smem_desc(0, 0, 1024, kSwizzle128);           // K-major
      auto dsc = [](uint64_t tmpl, uint32_t saddr) -> uint64_t { return tmpl | (uint64_t)((saddr & 0x3FFFF) >> 4); };
      while (sched.next(it)) {
        const uint32_t ab = nitem & (kNB - 1), aphase = (nitem / kNB) & 1;
        const uint32_t acc = tmem_base + ab * kACols;
        trace_ev(p, 1, nitem, 0);
        mbar_wait_prof(&tempty[ab], aphase ^ 1, prof_on, w_tempty);
        tc_fence_after();
        trace_ev(p, 1, nitem, 1);
        if (it.type == 1) {
          for (int sub = 0; sub < p.nsub; ++sub) {
            const int ksteps = G3D ? p.g3_kpad[sub] / 16 : p.TW / 16;
            uint32_t a_addr, b_addr;
            Ring r0 = r;
            if constexpr (D == 64) {
              mbar_wait_prof(&full[r.idx()], r.phase, prof_on, w_full);
              a_addr = ring_addr + r.idx() * kStageBytes;
              b_addr = a_addr + 16384;
              r.advance();
            } else {
              mbar_wait_prof(&full[r.idx()], r.phase, prof_on, w_full);
              a_addr = ring_addr + r.idx() * kStageBytes;
              r.advance();
              mbar_wait_prof(&full[r.idx()], r.phase, prof_on, w_full);
              b_addr = ring_addr + r.idx() * kStageBytes;
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
                  mma_f16_ss(acc + kKsC, da0 + ks * 128, desc_ones, idesc_p1_ones, ks ? 1u : first);
                }
            } else {
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                if (ks < ksteps) mma_f16_ss(acc, da0 + ks * 128, db0 + ks * 128, idesc_p1, ks ? 1u : first);
            }
            mma_commit(&empty[r0.idx()]);
            if constexpr (D == 128) mma_commit(&empty[r0.at(1).idx()]);
            if (p.ropenorm) {
              // ksum of the un-roped K (variant B with the normaliser) from its own stage.  NOTE: issued as a plain
              // (not unrolled, descriptor rebuilt per step) loop on purpose - a full-speed back-to-back run of these
              // narrow N=16 accumulating MMAs gave wrong sums on B200 (profiles/r01_bringup_notes.md).
              mbar_wait_prof(&full[r.idx()], r.phase, prof_on, w_full);
              tc_fence_after();
              const uint32_t n_addr = ring_addr + r.idx() * kStageBytes;
#pragma unroll 1
              for (int ks = 0; ks < ksteps; ++ks) {
                const uint64_t da = make_smem_desc(n_addr + ks * 2048, 16384, 1024, kSwizzle128);
                mma_f16_ss(acc + kKsC, da, desc_ones, idesc_p1_ones, (sub | ks) != 0);
              }
              mma_commit(&empty[r.idx()]);
              r.advance();
            }
          }
          mma_commit(&tfull[ab]);
          if (p.normalize) {
            // The Q stages are consumed by the epilogue warps, but THIS lane waits for them: every fill of every ring
            // slot must be observed by one role, in order.  mbarrier.try_wait.parity cannot tell "phase k complete"
            // from "phase k-2 complete": if this lane skipped the Q fills it could later wait on a slot whose previous
            // (Q) fill is still in flight, pass at once and feed a half-filled stage to the tensor core - the rare
            // launch failure of round 1 (profiles/r02_stall_root_cause.md).  The arrive below forwards "Q has landed"
            // to the epilogue through tfull (count 2), which also carries the visibility of the TMA-written tile.
            const int nq = D == 64 ? 1 : p.nsub;
            for (int k = 0; k < nq; ++k) { mbar_wait_prof(&full[r.idx()], r.phase, prof_on, w_full); r.advance(); }
          }
          mbar_arrive(&tfull[ab]);
        } else if (it.type == 2) {
          for (int slab = 0; slab < p.kslabs; ++slab) {
            uint32_t a_addr;
            int sa = -1;
            if (wres) {
              if (nitem == 0) mbar_wait_prof(&full[slab], 0, prof_on, w_full);   // resident mixing matrix: loaded once
              a_addr = ring_addr + slab * kStageBytes;
            } else {
              mbar_wait_prof(&full[r.idx()], r.phase, prof_on, w_full);
              a_addr = ring_addr + r.idx() * kStageBytes;
              sa = r.idx();
              r.advance();
            }
            mbar_wait_prof(&full[r.idx()], r.phase, prof_on, w_full);
            tc_fence_after();
            const uint64_t dhi0 = dsc(tmpl_k, a_addr), dlo0 = dsc(tmpl_k, a_addr + 16384);   // K-major: 32 B per k-step
            const uint64_t db0 = dsc(tmpl_mn8k, ring_addr + r.idx() * kStageBytes);           // MN-major: 2048 B per k-step
            const uint32_t first = slab != 0;
            // fp16 I/O: the S columns take the "hi" plane only (11 significant bits on a power-of-two normalised matrix -
            // finer than the 16-bit S it multiplies); the normaliser columns, and everything in bf16 (8-bit planes;
            // tcgen05 kind::f16 does not take an f16 A with a bf16 B), add the "lo" plane.
            if (!(p.is_fp16 || p.mix_hi_only) || (it.t % n2_cols) >= n2_scols) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                mma_f16_ss(acc, dhi0 + ks * 2, db0 + ks * 128, idesc_p2, ks ? 1u : first);
                mma_f16_ss(acc, dlo0 + ks * 2, db0 + ks * 128, idesc_p2, 1u);
              }
            } else {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) mma_f16_ss(acc, dhi0 + ks * 2, db0 + ks * 128, idesc_p2, ks ? 1u : first);
            }
            if (sa >= 0) mma_commit(&empty[sa]);
            mma_commit(&empty[r.idx()]);
            r.advance();
          }
          mma_commit(&tfull[ab]);
          mbar_arrive(&tfull[ab]);
        } else {
          Ring r0 = r;
          uint32_t b_addr;
          if constexpr (D == 128) {
            mbar_wait_prof(&full[r.idx()], r.phase, prof_on, w_full);
            b_addr = ring_addr + r.idx() * kStageBytes;
            r.advance();
          }
          for (int sub = 0; sub < p.nsub; ++sub) {
            mbar_wait_prof(&full[r.idx()], r.phase, prof_on, w_full);
            tc_fence_after();
            const uint32_t a_addr = ring_addr + r.idx() * kStageBytes;
            if constexpr (D == 64) { if (sub == 0) b_addr = a_addr + 16384; }
            const uint64_t da0 = dsc(tmpl_k, a_addr), db0 = dsc(tmpl_mn16k, b_addr);
#pragma unroll
            for (int ks = 0; ks < D / 16; ++ks)   // K-major A: 4 k-steps (32 B each) per 64-channel tile, tiles 16 KB apart
              mma_f16_ss(acc + sub * kSubC, da0 + (ks >> 2) * 1024 + (ks & 3) * 2, db0 + ks * 128, idesc_p3, ks != 0);
            r.advance();
          }
          const int ns = p3_stages<D>(p);
          for (int k = 0; k < ns; ++k) mma_commit(&empty[r0.at(k).idx()]);
          mma_commit(&tfull[ab]);
          mbar_arrive(&tfull[ab]);
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
    // ============================================================ scheduler (one lane)
    // Decides the order in which this CTA runs the items it owns and feeds the other roles through the FIFO.
    // Items of each kind are handed out through a global ticket counter (first come, first served: an SM that runs
    // faster simply takes more of them, so all CTAs finish together); the scheduler keeps at most one claimed item per
    // kind and decides which of them to run next:
    //   1. the claimed P2 item, if every P1 item of its group has signalled     (unblocks the whole group's readout)
    //   2. the claimed P3 item, if every P2 item of its group has signalled
    //   3. the claimed P1 item, unless it is `window` or more groups ahead of the claimed P3 item (bounds the L2 footprint)
    // An item enters the FIFO only when it is ready, so no other role ever waits on a dependency and the in-order
    // pipeline behind the FIFO cannot deadlock: tickets are claimed in group order, P1 items never depend on anything,
    // and a claimed-but-not-ready item never blocks the other kinds except through the window - whose P1 items all
    // belong to later groups than the P3 item that is being waited for.
    if (elect_one()) {
      const long long n1tot = (long long)p.G * p.M;
      const int n2per = p.n2_rows * n2_cols;
      const long long n2tot = (long long)p.G * n2per;
      const bool dyn = dynamic;
      // The block mixing of a group sits on the critical path between its summaries and its readout; with np2 > 0 a few
      // CTAs do nothing else, so a P2 item never queues behind streaming items in an in-order pipeline.
      const int np2 = dyn ? p.np2 : 0;
      bool has1 = true, has2 = true, has3 = true;
      const bool dedicated = np2 > 0 && (int)blockIdx.x < np2;
      if (np2 > 0) {
        if (dedicated) has1 = false; else has2 = false;   // (a dedicated CTA joins the readout once the mixing is done)
      }
      if (p.mode == 1) { has2 = false; has3 = false; }
      if (p.mode == 2) { has1 = false; has3 = false; }
      if (p.mode == 3 || p.mode == 5) { has1 = false; has2 = false; }
      if (p.mode == 4) has3 = false;
      const bool dep2 = p.mode == 0 || p.mode == 4;   // P2 items wait for their group's P1 counter
      const bool dep3 = p.mode == 0 || p.mode == 5;   // P3 items wait for their group's P2 counter
      const bool rev3 = (p.mode == 3 || p.mode == 5) && p.reverse3;   // stand-alone readout: last groups first
      unsigned long long* const tickets = reinterpret_cast<unsigned long long*>(p.counters + (size_t)2 * p.G * p.cnt_stride);
      // (a short queue also keeps the tickets balanced: a CTA never hoards items it will only reach much later)
      const uint32_t run_ahead = (uint32_t)p.run_ahead;
      uint32_t pub = 0;
      int ready1_g = -1, ready2_g = -1;   // groups already known to have all P1 / all P2 items done
      bool w_ok = !p.self_prep;           // the planes of the mixing matrix are complete (self_prep: every CTA has announced)
      auto emit = [&](int type, int g, int t) {
#ifdef MHLA_DIAG
        SpinGuard guard;
        while (pub - ld_acquire_cta_shared(q_started) >= run_ahead) {
          __nanosleep(32);
          if (guard.expired()) report_stall(3, pub, ld_acquire_cta_shared(q_started));
        }
#else
        uint32_t spins = 0;
        while (pub - ld_acquire_cta_shared(q_started) >= run_ahead) {
          __nanosleep(32);
          if (++spins > MHLA_SPIN_LIMIT) { printf("mhla: scheduler stalled (block %d)\n", blockIdx.x); __trap(); }
        }
#endif
        // (the slowest role is never more than a handful of items behind the producer: the ring, the two accumulators
        //  and the staging slots bound the distance well below kFifoDepth - run_ahead)
        *reinterpret_cast<volatile uint32_t*>(fifo + (pub & (kFifoDepth - 1))) = fifo_encode(type, g, t);
        st_release_cta_shared(q_published, ++pub);
      };
#ifdef MHLA_DIAG
      SpinGuard idle;
#else
      uint32_t idle = 0;
#endif
      uint32_t wg_load[2] = {0, 0};   // epilogue work handed to each warpgroup so far (arbitrary units)
      long long cur1 = -1, cur2 = -1, cur3 = -1;   // claimed, not yet enqueued
      // The first summary item of a CTA is fixed (linear block index = CTA index) and enqueued before anything else: the
      // first TMA loads of a launch do not wait for a ticket atomic and a readiness poll (two L2 round trips).  The
      // ticket counter then hands out the items from `base1` on.
      long long base1 = 0;
      if ((MHLA_STATIC_FIRST == 1 || (MHLA_STATIC_FIRST == 2 && D == 128)) && has1 && np2 == 0) {
        base1 = n1tot < (long long)gridDim.x ? n1tot : (long long)gridDim.x;
        if ((long long)blockIdx.x < base1) {
          wg_load[pub & 1] += p.normalize ? 7 : 3;
          emit(1, (int)(blockIdx.x / p.M), (int)(blockIdx.x % p.M));
        }
      }
      while (true) {
        {  // claim what is missing; the atomics are independent and overlap
          const bool n1 = has1 && cur1 < 0, n2 = has2 && cur2 < 0, n3 = has3 && cur3 < 0 && !(dedicated && (has2 || cur2 >= 0));
          unsigned long long a1 = 0, a2 = 0, a3 = 0;
          if (n1) a1 = atomicAdd(tickets + 0, 1ull);
          if (n2) a2 = atomicAdd(tickets + 8, 1ull);
          if (n3) a3 = atomicAdd(tickets + 16, 1ull);
          if (n1) {
            a1 += (unsigned long long)base1;
            if ((long long)a1 < n1tot) cur1 = (long long)a1;
            else {
              has1 = false;
              if (tl != nullptr) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); tl[2] = t; }
            }
          }
          if (n2) {
            if ((long long)a2 < n2tot) cur2 = (long long)a2;
            else {
              has2 = false;
              if (tl != nullptr) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); tl[3] = t; }
            }
          }
          if (n3) { if ((long long)a3 < n1tot) cur3 = (long long)a3; else has3 = false; }
        }
        if (cur1 < 0 && cur2 < 0 && cur3 < 0 && !(dedicated && has3)) break;   // every kind is exhausted
        if (cur1 < 0 && cur2 < 0 && cur3 < 0) continue;                         // (dedicated CTA: now claim readout items)
        const int g2 = cur2 >= 0 ? (int)(cur2 / n2per) : -1;
        int g3 = cur3 >= 0 ? (int)(cur3 / p.M) : -1;
        if (g3 >= 0 && rev3) g3 = p.G - 1 - g3;
        // fused kernel with the readout last: walk the groups backwards (the Q tiles the normaliser pulled in last are
        // still in L2), except that the final `tail` groups - whose block mixing is still in flight when the readout
        // begins - come at the very end
        if (g3 >= 0 && p.mode == 0 && p.policy == 1 && p.reverse3) {
          const int tail = p.G > 2 * p.reverse3 ? p.reverse3 : 0;
          if (g3 < p.G - tail) g3 = p.G - tail - 1 - g3;
        }
        {  // both polls are in flight together: one L2 round trip per decision
          const bool need1 = dep2 && g2 >= 0 && g2 != ready1_g, need2 = dep3 && g3 >= 0 && g3 != ready2_g;
          uint32_t c1 = 0, c2 = 0;
          if (need1) c1 = ld_acquire_gpu(p.counters + (size_t)g2 * p.cnt_stride);
          if (need2) c2 = ld_acquire_gpu(p.counters + (size_t)(p.G + g3) * p.cnt_stride);
          if (need1 && c1 >= (uint32_t)p.M) ready1_g = g2;
          if (!w_ok && g2 >= 0) w_ok = ld_acquire_gpu(p.counters + (size_t)2 * p.G * p.cnt_stride + 64) >= gridDim.x;
          if (need2 && c2 >= (uint32_t)n2per) ready2_g = g3;
        }
        const bool can2 = g2 >= 0 && (!dep2 || ready1_g == g2) && w_ok;
        const bool can3 = g3 >= 0 && (!dep3 || ready2_g == g3);
        const bool can1 = cur1 >= 0 && (p.mode != 0 || p.policy >= 1 || g3 < 0 || (int)(cur1 / p.M) < g3 + p.window);
        // Items alternate between the two epilogue warpgroups (FIFO index parity) and a P1 epilogue (normaliser) costs
        // about twice a P3 epilogue: when both kinds are available, give the heavier one to the less loaded warpgroup
        // instead of letting a strict P1/P3 alternation pile every P1 item onto the same warpgroup.
        int pick = 0;
        if (can1 && p.policy == 2) pick = 1;          // policy 2: phase by phase (all summaries, then mixing, then readout)
        else if (can2) pick = 2;
        else if (can1 && p.policy == 1) pick = 1;
        else if (can3 && can1) pick = (wg_load[pub & 1] <= wg_load[(pub & 1) ^ 1]) ? 1 : 3;
        else if (can3) pick = 3;
        else if (can1) pick = 1;
        if (pick == 2) {
          wg_load[pub & 1] += 8; emit(2, g2, (int)(cur2 % n2per)); cur2 = -1; idle = {};
        } else if (pick == 3) {
          wg_load[pub & 1] += 4; emit(3, g3, (int)(cur3 % p.M)); cur3 = -1; idle = {};
        } else if (pick == 1) {
          wg_load[pub & 1] += p.normalize ? 7 : 3; emit(1, (int)(cur1 / p.M), (int)(cur1 % p.M)); cur1 = -1; idle = {};
        } else {
          __nanosleep(100);
#ifdef MHLA_DIAG
          if (idle.expired()) report_stall(4, (uint32_t)(g2 & 0xFFFF) | ((uint32_t)(g3 & 0xFFFF) << 16), 0);
#else
          if (++idle > (1u << 23)) { printf("mhla: dependency wait timed out (block %d)\n", blockIdx.x); __trap(); }
#endif
        }
      }
      emit(0, 0, 0);
    }
  } else if (warp >= 4) {
    // ============================================================ epilogue warpgroups (TMEM -> regs -> smem -> global)
    // Workspace results (S, n_loc, S~, den) leave through a swizzled staging slot and coalesced 16-byte global stores
    // of the warpgroup itself - no TMA hand-off on the path that other CTAs are waiting for; completion is published
    // through wg_done[wg] and turned into the group's arrival counter by the signal warp.  Only the readout tiles (O)
    // go out by TMA (issued by thread 0 of the warpgroup, two slots in flight).
    const int q4 = warp & 3;                 // TMEM sub-partition (lanes 32*q4 .. 32*q4+31)
    const int wg = (warp - 4) >> 2;          // epilogue warpgroup: handles the items whose accumulator buffer is `wg`
    const int et = threadIdx.x - 128 - wg * kEpiThreads;   // 0..127 within the warpgroup
    ksum_s += wg * 128;
    const uint32_t bar_base = 1 + wg * 4;    // named barrier ids of this warpgroup
    const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
    const int nslots = p.slots_per_wg;
    uint8_t* const my_staging = staging + wg * nslots * slot_bytes;
    Ring r(nst, wres ? p.kslabs : 0);
    uint32_t nitem = 0;
    uint32_t nchunk = 0;                     // staging chunks written so far (slot = nchunk & 1)
    uint32_t ndone = 0;                      // workspace-producing items finished (published through wg_done[wg])
    int last_tma_slot = -1;                  // slot read by the most recent TMA store of this warpgroup (thread 0's view)
    uint32_t v[32];
    const bool prof_on = p.prof != nullptr && et == 0 && wg == 0;
    long long w_tfull = 0, w_sfree = 0, w_q = 0, t_p1 = 0, t_p2 = 0, t_p3 = 0, t_ld = 0, t_out = 0, t_pack = 0, t_stage = 0;

    // write one [rows][128 B] chunk row into the swizzle-128B staging tile
    auto stage_row = [&](uint8_t* buf, int row, const uint32_t* w32) {
      const long long t0 = prof_on ? clock64() : 0;
      uint4* dst = reinterpret_cast<uint4*>(buf + row * 128);
#pragma unroll
      for (int c = 0; c < 8; ++c)
        dst[c ^ (row & 7)] = make_uint4(w32[4 * c], w32[4 * c + 1], w32[4 * c + 2], w32[4 * c + 3]);
      if (prof_on) t_stage += clock64() - t0;
    };
    // next staging slot; a TMA store that is still reading it (issued two chunks ago) must have finished
    auto slot_acquire = [&]() -> uint8_t* {
      const int s_ = nslots == 1 ? 0 : (int)(nchunk & 1);
      const long long t0 = prof_on ? clock64() : 0;
      if (et == 0 && last_tma_slot >= 0) {
        if (last_tma_slot == s_) tma_store_wait_read<0>(); else tma_store_wait_read<1>();
      }
      named_bar_sync(bar_base + 1, kEpiThreads);
      if (prof_on) w_sfree += clock64() - t0;
      return my_staging + s_ * slot_bytes;
    };
    // rows of the slot are complete: copy [nrows][128 B] to global memory, row r to gbase + r * row_stride (bytes)
    auto chunk_copy = [&](const uint8_t* buf, uint8_t* gbase, size_t row_stride, int nrows, int nvalid) {
      named_bar_sync(bar_base, kEpiThreads);
      const long long t0 = prof_on ? clock64() : 0;
      // thread -> (row = k * 16 + et / 8, 16-byte piece et % 8): a warp instruction covers 4 rows x 128 contiguous bytes
      const int c = et & 7, r0_ = et >> 3;
      const uint8_t* sp = buf + r0_ * 128 + ((c ^ (r0_ & 7)) << 4);   // (k * 16 + r0_) & 7 == r0_ & 7
      uint8_t* gp = gbase + (size_t)r0_ * row_stride + c * 16;
      auto run = [&](auto KN) {
        constexpr int kn = decltype(KN)::value;
        uint4 val[kn];
#pragma unroll
        for (int k = 0; k < kn; ++k) val[k] = *reinterpret_cast<const uint4*>(sp + k * 16 * 128);
#pragma unroll
        for (int k = 0; k < kn; ++k)
          if (k * 16 + r0_ < nvalid) __stcg(reinterpret_cast<uint4*>(gp + (size_t)k * 16 * row_stride), val[k]);
      };
      if (nrows == 64) run(std::integral_constant<int, 4>{}); else run(std::integral_constant<int, 8>{});
      if (prof_on) t_out += clock64() - t0;
      ++nchunk;
    };
    // rows of the slot are complete: thread 0 sends it with TMA
    auto chunk_tma_begin = [&]() { fence_proxy_async_smem(); named_bar_sync(bar_base, kEpiThreads); };
    auto chunk_tma_end = [&]() {
      if (et == 0) { tma_store_commit(); last_tma_slot = nslots == 1 ? 0 : (int)(nchunk & 1); }
      ++nchunk;
    };
    // every global store of this item has been issued by all 128 threads: publish
    auto item_done = [&]() {
      named_bar_sync(bar_base + 3, kEpiThreads);
      ++ndone;
      if (et == 0) st_release_cta_shared(&wg_done[wg], ndone);
    };
    // load 64 fp32 accumulator columns, scale, round to the 16-bit I/O type: 32 packed words = one 128-byte row
    auto load_pack64 = [&](uint32_t taddr, float scale, uint32_t* pk) {
      uint32_t v2[32];
      const long long t0 = prof_on ? clock64() : 0;
      tmem_ld_x32(taddr, v);
      tmem_ld_x32(taddr + 32, v2);
      tmem_ld_wait();
      if (prof_on) t_ld += clock64() - t0;
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
      if (prof_on) t_pack += clock64() - t0;
    };
    // same with a per-column weight (fused output RMSNorm)
    auto load_pack64_w = [&](uint32_t taddr, float scale, const float* wcol, uint32_t* pk) {
      uint32_t v2[32];
      tmem_ld_x32(taddr, v);
      tmem_ld_x32(taddr + 32, v2);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const float a = __uint_as_float(v[2 * e]) * scale * __ldg(wcol + 2 * e);
        const float bq = __uint_as_float(v[2 * e + 1]) * scale * __ldg(wcol + 2 * e + 1);
        const float c2 = __uint_as_float(v2[2 * e]) * scale * __ldg(wcol + 32 + 2 * e);
        const float d2 = __uint_as_float(v2[2 * e + 1]) * scale * __ldg(wcol + 32 + 2 * e + 1);
        if (p.is_fp16) {
          __half2 h0 = __floats2half2_rn(a, bq), h1 = __floats2half2_rn(c2, d2);
          pk[e] = *reinterpret_cast<uint32_t*>(&h0); pk[16 + e] = *reinterpret_cast<uint32_t*>(&h1);
        } else {
          pk[e] = pack_bf16x2(a, bq); pk[16 + e] = pack_bf16x2(c2, d2);
        }
      }
    };
    // same with the fused post-ops (POST instantiations only): optional per-column weight, SiLU gate and additive term.
    // An epilogue thread owns one token ROW (its TMEM lane), but a warp instruction in which every lane reads 16 bytes of
    // its own row touches 32 cache lines: measured, the L1 tag stage made the fused epilogue slower than separate
    // elementwise passes (275 vs 262 us on the Wan layer).  So the warp fetches the gate / add pieces of ITS 32 rows
    // coalesced - 8 lanes cover one row's [gate 64 B | add 64 B] of a 32-column half, 4 rows per instruction - bounces
    // them through the (still free) rows of the staging slot the output tile is about to be written to, and every lane
    // reads its own row back from shared memory.  32 columns at a time (register budget).
    //   gbase / abase: gate / add tensors at this item's (b, h[, block]); ts_mine: this thread's token row (-1: none)
    auto load_pack64_post = [&](uint32_t taddr, float scale, const float* wcol, const uint16_t* gbase, long long g_sw,
                                const uint16_t* abase, long long a_sw, int ts_mine, int coff, uint8_t* buf, uint32_t* pk) {
      const int pc = lane & 7, rsel = lane >> 3;
      const uint16_t* const tb = pc < 4 ? gbase : abase;
      const long long tsw = pc < 4 ? g_sw : a_sw;
      const int row0 = et & ~31;   // first row of this warp inside the tile
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint4 tmp[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int ts = __shfl_sync(0xffffffffu, ts_mine, i * 4 + rsel);
          tmp[i] = make_uint4(0u, 0u, 0u, 0u);
          if (tb != nullptr && ts >= 0) tmp[i] = __ldg(reinterpret_cast<const uint4*>(tb + (long long)ts * tsw + coff + hh * 32 + (pc & 3) * 8));
        }
        __syncwarp();   // the previous half's own-row reads are done
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = row0 + i * 4 + rsel;
          *reinterpret_cast<uint4*>(buf + rr * 128 + ((pc ^ (rr & 7)) << 4)) = tmp[i];
        }
        __syncwarp();
        uint32_t gw[16], aw[16];
        const uint4* rowp = reinterpret_cast<const uint4*>(buf + et * 128);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 g4 = rowp[i ^ (et & 7)], a4 = rowp[(4 + i) ^ (et & 7)];
          gw[4 * i] = g4.x; gw[4 * i + 1] = g4.y; gw[4 * i + 2] = g4.z; gw[4 * i + 3] = g4.w;
          aw[4 * i] = a4.x; aw[4 * i + 1] = a4.y; aw[4 * i + 2] = a4.z; aw[4 * i + 3] = a4.w;
        }
        tmem_ld_x32(taddr + hh * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          float x0 = __uint_as_float(v[2 * e]) * scale, x1 = __uint_as_float(v[2 * e + 1]) * scale;
          if (wcol != nullptr) { x0 *= __ldg(wcol + hh * 32 + 2 * e); x1 *= __ldg(wcol + hh * 32 + 2 * e + 1); }
          float2 g2, a2;
          if (p.is_fp16) {
            g2 = __half22float2(*reinterpret_cast<const __half2*>(&gw[e]));
            a2 = __half22float2(*reinterpret_cast<const __half2*>(&aw[e]));
          } else {
            g2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gw[e]));
            a2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&aw[e]));
          }
          if (gbase != nullptr) {   // SiLU(g) = g / (1 + exp(-g))
            x0 *= __fdividef(g2.x, 1.0f + __expf(-g2.x));
            x1 *= __fdividef(g2.y, 1.0f + __expf(-g2.y));
          }
          x0 += a2.x; x1 += a2.y;   // (zeros when the additive term is off)
          if (p.is_fp16) { __half2 h0 = __floats2half2_rn(x0, x1); pk[hh * 16 + e] = *reinterpret_cast<uint32_t*>(&h0); }
          else pk[hh * 16 + e] = pack_bf16x2(x0, x1);
        }
      }
      __syncwarp();   // every lane has read its row: the rows may now be overwritten with the output tile
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

    // wait for the accumulator of the current item.  MHLA_EPI_POLL1: only the first warp of the warpgroup polls the
    // mbarrier, the other three sleep in a named barrier (their polling loops would share issue slots with the other
    // warpgroup's epilogue)
    auto wait_tfull = [&](uint32_t ab_, uint32_t aphase_) {
#if MHLA_EPI_POLL1
      if (q4 == 0) mbar_wait_prof(&tfull[ab_], aphase_, prof_on, w_tfull);
      named_bar_sync(bar_base + 2, kEpiThreads);
#else
      mbar_wait_prof(&tfull[ab_], aphase_, prof_on, w_tfull);
#endif
    };
    if (p.self_prep) {
      // No prologue kernel: the 256 epilogue threads of every CTA split their share of the fp32 mixing matrix into the
      // hi | lo planes the P2 items load by TMA (rows blockIdx.x, blockIdx.x + gridDim.x, ...) and announce it; the
      // scheduler holds P2 items back until every CTA has done so.  The first summaries are still in flight meanwhile.
      const int t256 = (int)threadIdx.x - 128;
      float down = 1.0f;
      if (p.is_fp16) {   // fp16 planes: normalise by a power of two (every CTA reduces the whole matrix - it is tiny)
        float amax = 0.f;
        const int nn = p.M0 * p.M0;
        for (int idx = t256; idx < nn; idx += 256) amax = fmaxf(amax, fabsf(__ldg(p.mix + (long long)(idx / p.M0) * p.mix_ld + idx % p.M0)));
        for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        float* red = reinterpret_cast<float*>(smem + kSmemKsum);   // scratch: the ksum area is not in use yet
        if (lane == 0) red[warp - 4] = amax;
        named_bar_sync(9, 256);
        amax = 0.f;
        for (int i = 0; i < 8; ++i) amax = fmaxf(amax, red[i]);
        named_bar_sync(9, 256);
        int e = 0;
        if (amax > 0.f && amax < 3.0e38f) { (void)frexpf(amax, &e); }
        e = e < -100 ? -100 : (e > 100 ? 100 : e);
        down = ldexpf(1.0f, -e);
        if (t256 == 0) *wscale_s = ldexpf(1.0f, e);
      } else if (t256 == 0) {
        *wscale_s = 1.0f;
      }
      const int nplane = p.M * p.Mp;
      for (int row = blockIdx.x; row < p.M; row += gridDim.x)
        for (int j = t256; j < p.Mp; j += 256) {
          // packed groups: the scheduled matrix is block-diagonal, `pack` copies of the caller's M0 x M0 matrix
          const float val = (j < p.M && row / p.M0 == j / p.M0)
                                ? __ldg(p.mix + (long long)(row % p.M0) * p.mix_ld + j % p.M0) * down : 0.f;
          uint16_t hi, lo;
          split16(val, hi, lo);
          p.w_planes[row * p.Mp + j] = hi;
          p.w_planes[nplane + row * p.Mp + j] = lo;
        }
      named_bar_sync(9, 256);
      if (t256 == 0) red_release_gpu_add(p.counters + (size_t)2 * p.G * p.cnt_stride + 64, 1u);
    }
    while (sched.next_warp(it, lane)) {
      const uint32_t ab = nitem & (kNB - 1), aphase = (nitem / kNB) & 1;
      const uint32_t acc = tmem_base + ab * kACols + lane_sel;
      if ((int)(ab & 1) != wg) {   // the other warpgroup's item: only keep the ring bookkeeping in step
        r.advance(it.type == 1 ? p1_stages<D>(p) : (it.type == 2 ? (wres ? 1 : 2) * p.kslabs : p3_stages<D>(p)));
        ++nitem;
        continue;
      }
      const long long t_item = prof_on ? clock64() : 0;
      if (et == 0) trace_ev(p, 2, nitem, 0);
      if (it.type == 1) {
        wait_tfull(ab, aphase);
        if (et == 0) trace_ev(p, 2, nitem, 1);
        tc_fence_after();
        // rows of S live in TMEM lanes: D == 128 -> lane = row; D == 64 (M=64 MMA) -> row r in lane 32*(r/16)+r%16
        const bool row_ok = (D == 128) || (lane < 16);
        const int row = (D == 128) ? et : (q4 * 16 + (lane & 15));
        const int kvs = (D == 64 ? 1 : 2) + p.ropenorm;
        const size_t blk = (size_t)it.g * p.M + it.t;
        uint16_t* const srow = p.ws_S + blk * p.ncols;
        r.advance(p.nsub * kvs);
#if MHLA_EARLY_P1
        // D = 64: the whole accumulator (64 S columns + ksum) fits the register budget, so it is read FIRST and handed
        // back to the issuer before the n_loc dot products and the staging of S - the MMAs of the item after next no
        // longer wait for this epilogue
        uint32_t pk_early[32];
        uint32_t ks_early = 0;
        if constexpr (D == 64) {
          if (p.normalize) tmem_ld_x1(acc + kKsC, ks_early);
          load_pack64(acc, 1.0f, pk_early);   // (its tcgen05.wait::ld covers the ksum column as well)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[ab]);
        }
#endif
        if (p.normalize) {
          // ---- n_loc first: it frees the Q stage(s) of the ring early
          uint32_t ks;
#if MHLA_EARLY_P1
          if constexpr (D == 64) ks = ks_early;
          else
#endif
          {
            tmem_ld_x1(acc + kKsC, ks);
            tmem_ld_wait();
          }
          if (row_ok) ksum_s[row] = __uint_as_float(ks);
          named_bar_sync(bar_base + 2, kEpiThreads);  // ksum_s complete
          uint16_t* const nbuf = srow + D * D;        // [hi: wpad][lo: wpad]
          // (the Q tiles have landed: the issuer waited for them before its arrive on tfull)
          if constexpr (D == 64) {
            const uint8_t* qs = ring + r.idx() * kStageBytes;
            for (int t = et; t < p.wpad; t += kEpiThreads) {
              uint16_t hi, lo;
              split16(dot_row64(qs, t, ksum_s, 0.f), hi, lo);
              nbuf[t] = hi;
              nbuf[p.wpad + t] = lo;
            }
            named_bar_sync(bar_base + 2, kEpiThreads);
            if (et == 0) mbar_arrive(&empty[r.idx()]);
            r.advance();
          } else {
            for (int sub = 0; sub < p.nsub; ++sub) {
              const uint8_t* qs = ring + r.idx() * kStageBytes;
              if (et < p.TW) {
                const int t = sub * p.TW + et;
                float a = dot_row64(qs, et, ksum_s, 0.f);
                a = dot_row64(qs + 16384, et, ksum_s + 64, a);
                uint16_t hi, lo;
                split16(a, hi, lo);
                nbuf[t] = hi;
                nbuf[p.wpad + t] = lo;
              }
              named_bar_sync(bar_base + 2, kEpiThreads);
              if (et == 0) mbar_arrive(&empty[r.idx()]);
              r.advance();
            }
          }
        }
#if MHLA_EARLY_P1
        if constexpr (D == 64) {
          uint8_t* buf = slot_acquire();
          if (row_ok) stage_row(buf, row, pk_early);
          chunk_copy(buf, reinterpret_cast<uint8_t*>(srow), (size_t)D * 2, D, D);
        } else
#endif
        {
          for (int c = 0; c < D / 64; ++c) {
            uint32_t pk[32];
            load_pack64(acc + c * 64, 1.0f, pk);
            uint8_t* buf = slot_acquire();
            if (row_ok) stage_row(buf, row, pk);
            chunk_copy(buf, reinterpret_cast<uint8_t*>(srow) + c * 128, (size_t)D * 2, D, D);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[ab]);
        }
        item_done();
      } else if (it.type == 2) {
        const int ti = it.t / n2_cols, tc = it.t % n2_cols;
        const int nvalid = p.M - ti * 128;       // rows of this tile inside the matrix (TMA used to clip them)
        const size_t row0 = (size_t)it.g * p.M + (size_t)ti * 128;
        wait_tfull(ab, aphase);
        if (et == 0) trace_ev(p, 2, nitem, 1);
        tc_fence_after();
        const float wsc = p.self_prep ? *wscale_s : __ldg(p.wscale);   // undo the power-of-two normalisation (exact)
        if (tc < n2_scols) {
          for (int c = 0; c < kPN / 64; ++c) {
            uint32_t pk[32];
            load_pack64(acc + c * 64, wsc, pk);
#ifdef MHLA_EARLY_P2
            if (c == kPN / 64 - 1) {   // last read of the accumulator: hand it back before the tile is staged and copied out
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tempty[ab]);
            }
#endif
            uint8_t* buf = slot_acquire();
            stage_row(buf, et, pk);
            chunk_copy(buf, reinterpret_cast<uint8_t*>(p.ws_St + row0 * (size_t)(D * D) + tc * kPN + c * 64),
                       (size_t)D * D * 2, 128, nvalid);
          }
        } else {
          for (int q = 0; q < kPN / 32; ++q) {
            const int col0 = (tc - n2_scols) * kPN + q * 32;
            if (col0 >= 2 * p.wpad) break;   // uniform: nothing left in this tile
            tmem_ld_x32(acc + q * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * wsc);
            uint8_t* buf = slot_acquire();
            stage_row(buf, et, v);
            chunk_copy(buf, reinterpret_cast<uint8_t*>(const_cast<float*>(p.den) + row0 * (size_t)(2 * p.wpad) + col0),
                       (size_t)2 * p.wpad * 4, 128, nvalid);
          }
        }
#ifdef MHLA_EARLY_P2
        if (tc >= n2_scols)
#endif
        {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[ab]);
        }
        r.advance((wres ? 1 : 2) * p.kslabs);
        item_done();
      } else {
        const int i = it.t;                               // block row within the scheduled (packed) group
        const int gr = it.g * p.pack + it.t / p.M0;       // real (b,h) group and block inside it: tensor coordinates
        const int b = gr / p.H, h = gr % p.H, ib = it.t % p.M0;
        float dsum[2] = {1.f, 1.f};
        if (p.normalize) {
          // den = mix.n_loc_hi + mix.n_loc_lo + eps was written by other CTAs; the item was only enqueued after its
          // group's P2 counter had been acquired, so it is visible: fetch it through L2 now and let the latency overlap
          // the wait for the accumulator.
          const float* dg = p.den + (size_t)(it.g * p.M + i) * (2 * p.wpad);
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            const int t = sub * p.TW + et;
            const bool valid = G3D ? (sub < p.nsub && et < p.g3_rows[sub]) : (sub < p.nsub && et < p.TW && t < p.w);
            if (valid) dsum[sub] = __ldcg(dg + t) + __ldcg(dg + p.wpad + t) + p.eps;
          }
        }
        // fused post-ops: this thread's token rows inside the gate / add tensors (-1: the row lies beyond the block and is
        // never stored).  Formed BEFORE the wait for the accumulator, and the rows' 128-byte pieces are pulled into L2 now.
        int ts_s[2] = {-1, -1};
        const uint16_t* gbase = nullptr;
        const uint16_t* abase = nullptr;
        if constexpr (POST) {
          if (p.post_gate != nullptr) gbase = p.post_gate + b * p.pg_sb + h * p.pg_sh + (G3D ? 0 : ib * p.pg_sm);
          if (p.post_add != nullptr) abase = p.post_add + b * p.pa_sb + h * p.pa_sh + (G3D ? 0 : ib * p.pa_sm);
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            if (sub < p.nsub) {
              if constexpr (!G3D) {   // block-major: token inside the block
                const int t = sub * p.TW + et;
                if (et < p.TW && t < p.w) ts_s[sub] = t;
              } else {                // 3-D view: token inside the sample
                if (et < p.g3_rows[sub]) {
                  const int pp = p.g3_p2 * p.g3_p3, a = et / pp, rem = et - a * pp, y = rem / p.g3_p3, x = rem - y * p.g3_p3;
                  const int wbi = ib % p.g3_wb, jj = ib / p.g3_wb, hbi = jj % p.g3_hb, fbi = jj / p.g3_hb;
                  ts_s[sub] = ((fbi * p.g3_p1 + sub * p.g3_aper + a) * p.g3_H + hbi * p.g3_p2 + y) * p.g3_W + wbi * p.g3_p3 + x;
                }
              }
              if (ts_s[sub] >= 0) {
#pragma unroll
                for (int c = 0; c < D / 64; ++c) {
                  if (gbase != nullptr) prefetch_l2(gbase + (long long)ts_s[sub] * p.pg_sw + c * 64);
                  if (abase != nullptr) prefetch_l2(abase + (long long)ts_s[sub] * p.pa_sw + c * 64);
                }
              }
            }
          }
        }
        wait_tfull(ab, aphase);
        if (et == 0) trace_ev(p, 2, nitem, 1);
        tc_fence_after();
#if MHLA_EARLY_P3 == 1
        // D = 64, no fused post-ops: both 64-column halves of the readout accumulator are scaled and rounded into
        // registers first and the buffer goes back to the issuer BEFORE the two tiles are staged and stored
        bool early_done = false;
        if constexpr (D == 64 && !POST && !G3D) {
          if (p.rms_w == nullptr) {
            uint32_t pk0[32], pk1[32];
            auto load_pack64_lr = [&](uint32_t taddr, float scale, uint32_t* pk) {   // 32 columns at a time: fewer live registers
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                tmem_ld_x32(taddr + hh * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                  const float a = __uint_as_float(v[2 * e]) * scale, bq = __uint_as_float(v[2 * e + 1]) * scale;
                  if (p.is_fp16) { __half2 h0 = __floats2half2_rn(a, bq); pk[hh * 16 + e] = *reinterpret_cast<uint32_t*>(&h0); }
                  else pk[hh * 16 + e] = pack_bf16x2(a, bq);
                }
              }
            };
            load_pack64_lr(acc, p.normalize ? 1.0f / dsum[0] : 1.0f, pk0);
            if (p.nsub > 1) load_pack64_lr(acc + kSubC, p.normalize ? 1.0f / dsum[1] : 1.0f, pk1);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[ab]);
            const uint64_t oh = p.o_hint ? kEvictFirst : kEvictNormal;
            auto send = [&](int sub, const uint32_t* pk) {
              uint8_t* buf = slot_acquire();
              stage_row(buf, et, pk);
              chunk_tma_begin();
              if (et == 0) tma_store_5d_hint(&p.tmO, buf, 0, sub * p.TW, ib, h, b, oh);
              chunk_tma_end();
            };
            send(0, pk0);
            if (p.nsub > 1) send(1, pk1);
            early_done = true;
          }
        }
        if (!early_done) {
#endif
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          if (sub >= p.nsub) break;
          float rden = p.normalize ? 1.0f / dsum[sub] : 1.0f;
          if (p.rms_w != nullptr) {
            // fused per-(token, head) RMS normalisation: first pass over the row's D accumulator columns for the sum of
            // squares (TMEM reads are cheap), the scale then rides along with the normaliser in the second pass
            float ss = 0.f;
            for (int c = 0; c < D / 64; ++c) {
              uint32_t v2[32];
              tmem_ld_x32(acc + sub * kSubC + c * 64, v);
              tmem_ld_x32(acc + sub * kSubC + c * 64 + 32, v2);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 32; ++e) {
                const float a = __uint_as_float(v[e]) * rden, b2 = __uint_as_float(v2[e]) * rden;
                ss = fmaf(a, a, ss); ss = fmaf(b2, b2, ss);
              }
            }
            rden *= rsqrtf(ss * (1.0f / D) + p.rms_eps);
          }
          for (int c = 0; c < D / 64; ++c) {
            uint32_t pk[32];
            uint8_t* buf = nullptr;
            if constexpr (POST) {
              buf = slot_acquire();   // (the slot doubles as the bounce buffer of the gate / add pieces)
              load_pack64_post(acc + sub * kSubC + c * 64, rden, p.rms_w != nullptr ? p.rms_w + c * 64 : nullptr, gbase, p.pg_sw,
                               abase, p.pa_sw, ts_s[sub], c * 64, buf, pk);
            } else if (p.rms_w != nullptr) load_pack64_w(acc + sub * kSubC + c * 64, rden, p.rms_w + c * 64, pk);
            else load_pack64(acc + sub * kSubC + c * 64, rden, pk);
#if MHLA_EARLY_P3 == 2
            // last read of this accumulator buffer: hand it back to the issuer before the tile is staged and stored
            if (sub == p.nsub - 1 && c == D / 64 - 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tempty[ab]);
            }
#endif
            if constexpr (!POST) buf = slot_acquire();
            stage_row(buf, et, pk);
            chunk_tma_begin();
            // the output is never read again: mark its lines evict-first so that they leave L2 before the Q tiles the
            // readout of later groups still needs
            if (et == 0) {
              const uint64_t oh = p.o_hint ? kEvictFirst : kEvictNormal;
              if constexpr (!G3D) {
                tma_store_5d_hint(&p.tmO, buf, c * 64, sub * p.TW, ib, h, b, oh);
              } else {   // inverse of the block gather (mhla_utils.py:345-354): one box per sub-tile
                const int wbi = ib % p.g3_wb, jj = ib / p.g3_wb, hbi = jj % p.g3_hb, fbi = jj / p.g3_hb;
                tma_store_5d_hint(&p.tm3[5][(sub == p.nsub - 1) ? p.g3_tail : 0], buf, c * 64, wbi * p.g3_p3, hbi * p.g3_p2,
                                  b * p.g3_F + fbi * p.g3_p1 + sub * p.g3_aper, h, oh);
              }
            }
            chunk_tma_end();
          }
        }
#if MHLA_EARLY_P3 != 2
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[ab]);
#endif
#if MHLA_EARLY_P3 == 1
        }
#endif
        r.advance(p3_stages<D>(p));
      }
      if (et == 0) trace_ev(p, 2, nitem, 2);
      if (prof_on) {
        const long long dt = clock64() - t_item;
        if (it.type == 1) t_p1 += dt; else if (it.type == 2) t_p2 += dt; else t_p3 += dt;
      }
      ++nitem;
    }
    if (et == 0) tma_store_wait_all<0>();
    if (prof_on) {
      unsigned long long* pr = p.prof + (size_t)blockIdx.x * 16;
      pr[5] = (unsigned long long)w_tfull; pr[6] = (unsigned long long)w_sfree; pr[7] = (unsigned long long)w_q;
      pr[8] = (unsigned long long)t_p1; pr[9] = (unsigned long long)t_p2; pr[10] = (unsigned long long)t_p3;
      pr[11] = nitem;
      pr[12] = (unsigned long long)t_ld; pr[13] = (unsigned long long)t_out;
      pr[7] = (unsigned long long)t_pack; pr[15] = (unsigned long long)t_stage;
    }
  } else if (warp == 3) {
    // ============================================================ signal warp (one lane)
    // Turns "warpgroup finished item n" (wg_done, shared memory) into the group's arrival counter in global memory.
    // The release (gpu scope) orders every global store of the warpgroup's 128 threads before the increment: they
    // happen before the warpgroup's named barrier, thread 0's st.release.cta and this lane's ld.acquire.cta.  The
    // ~1 us the release takes under load is spent here, not in the epilogue.
    if (dynamic && elect_one()) {
      uint32_t seen[2] = {0, 0};
      uint32_t n = 0;
      while (sched.next(it)) {
        const int wgi = (int)(n & 1);
        trace_ev(p, 3, n, 0);
        if (it.type != 3) {
          ++seen[wgi];
#ifdef MHLA_DIAG
          SpinGuard guard;
          while (ld_acquire_cta_shared(&wg_done[wgi]) < seen[wgi]) {
            __nanosleep(32);
            if (guard.expired()) report_stall(5, (uint32_t)wgi, seen[wgi]);
          }
#else
          uint32_t spins = 0;
          while (ld_acquire_cta_shared(&wg_done[wgi]) < seen[wgi]) {
            __nanosleep(32);
            if (++spins > MHLA_SPIN_LIMIT) { printf("mhla: signal wait timed out (block %d)\n", blockIdx.x); __trap(); }
          }
#endif
          trace_ev(p, 3, n, 1);
          red_release_gpu_add(p.counters + (size_t)((it.type == 1 ? 0 : p.G) + it.g) * p.cnt_stride, 1u);
        }
        trace_ev(p, 3, n, 2);
        ++n;
      }
    }
  }

  // ---------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (tl != nullptr && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); tl[1] = t; }
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
  if (p.self_prep) {
    // The last CTA to get here re-zeroes the control block (counters, tickets, flags) for the next call: every other
    // CTA has made its last access to it before incrementing the exit counter.
    uint32_t* const tail = p.counters + (size_t)2 * p.G * p.cnt_stride;
    if (threadIdx.x == 0) {
      __threadfence();
      *last_flag = (atomicAdd(tail + 80, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (*last_flag) {
      __threadfence();
      const int nwords = 2 * p.G * p.cnt_stride + 128;
      for (int idx = threadIdx.x; idx < nwords; idx += blockDim.x) p.counters[idx] = 0u;
    }
  }
}

// Tiny prologue: split the fp32 mixing matrix into hi + lo 16-bit planes [2][M][Mp] (optionally keeping only the
// strictly-lower triangle and folding a scale, for the causal variant) and zero the dependency counters.
__global__ void prep_mix_kernel(const float* __restrict__ mix, long long ld, uint16_t* __restrict__ out, int M, int Mp,
                                int M0, int strict_lower, float scale, int is_fp16, uint32_t* counters, int ncounters) {
  grid_launch_dependents();   // the main kernel may start its prologue now; it waits (griddepcontrol.wait) for our results
  // M = pack * M0: block-diagonal, `pack` copies of the caller's M0 x M0 matrix (packing of consecutive (b,h) groups)
  const int n = M * Mp;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
    const int i = idx / Mp, j = idx % Mp;
    float v = 0.f;
    if (j < M && i / M0 == j / M0 && (!strict_lower || j % M0 < i % M0)) v = mix[(long long)(i % M0) * ld + j % M0] * scale;
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


// Prologue of the block-mixed path: normalise the fp32 mixing matrix by a power of two (largest magnitude in [0.5, 1)),
// split it into hi + lo planes [2][M][Mp] of the I/O type (fp16: 22 significant bits together, bf16: 16), publish the power of two for the
// epilogue and zero the dependency counters / item tickets.  Every block reduces the whole matrix (M*M floats,
// L2-resident after the first touch) so that no inter-block synchronisation is needed.
__global__ void prep_mix_scaled_kernel(const float* __restrict__ mix, long long ld, uint16_t* __restrict__ out, int M,
                                       int Mp, int M0, int is_fp16, float* __restrict__ wscale, uint32_t* counters,
                                       int ncounters) {
  // M = pack * M0: the scheduled matrix is block-diagonal with `pack` copies of the caller's M0 x M0 matrix
  grid_launch_dependents();
  // bf16 planes carry the fp32 exponent range: only fp16 needs the normalisation (and pays for the reduction)
  int e = 0;
  if (is_fp16) {
    __shared__ float red[32];
    float amax = 0.f;
    const int nn = M0 * M0;
#pragma unroll 4
    for (int idx = threadIdx.x; idx < nn; idx += blockDim.x)
      amax = fmaxf(amax, fabsf(__ldg(mix + (long long)(idx / M0) * ld + idx % M0)));
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = amax;
    __syncthreads();
    amax = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) amax = fmaxf(amax, red[i]);
    if (amax > 0.f && amax < 3.0e38f) { (void)frexpf(amax, &e); }
    e = e < -100 ? -100 : (e > 100 ? 100 : e);
  }
  const float down = ldexpf(1.0f, -e);
  if (blockIdx.x == 0 && threadIdx.x == 0) *wscale = ldexpf(1.0f, e);
  const int n = M * Mp;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
    const int i = idx / Mp, j = idx % Mp;
    float v = 0.f;
    if (j < M && i / M0 == j / M0) v = mix[(long long)(i % M0) * ld + j % M0] * down;
    if (is_fp16) {
      const __half h = __float2half_rn(v);
      out[idx] = __half_as_ushort(h);
      out[n + idx] = __half_as_ushort(__float2half_rn(v - __half2float(h)));
    } else {
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      out[idx] = __bfloat16_as_ushort(h);
      out[n + idx] = __bfloat16_as_ushort(__float2bfloat16_rn(v - __bfloat162float(h)));
    }
  }
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < ncounters; idx += gridDim.x * blockDim.x)
    counters[idx] = 0u;
}

}  // namespace mhla
