// C-ABI host side of libmhla_b200.so: argument validation, workspace carving, TMA tensor-map encoding
// (driver entry point resolved at run time - no link-time dependency on libcuda) and kernel launches.
// See include/mhla_b200.h for the contract.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/mhla_b200.h"
#include "blockmix_kernel.cuh"
#include "causal_kernel.cuh"
#include "smalln_kernel.cuh"
#include "wan_prep_kernel.cuh"
#include "bwd_aux_kernel.cuh"
#include "gated_norm_kernel.cuh"

namespace {

thread_local std::string g_last_cuda_error;
thread_local int g_last_launches = 0;

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode_fn() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
  });
  return fn;
}

bool cuda_ok(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return true;
  g_last_cuda_error = std::string(what) + ": " + cudaGetErrorString(e);
  return false;
}

struct MapSpec {
  CUtensorMapDataType dt;
  int rank;
  void* base;
  uint64_t dims[5];
  uint64_t strides[4];  // bytes, dims 1..rank-1
  uint32_t box[5];
};

bool encode_map(CUtensorMap* out, const MapSpec& s) {
  EncodeFn fn = get_encode_fn();
  if (!fn) { g_last_cuda_error = "cuTensorMapEncodeTiled entry point not found"; return false; }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  cuuint64_t dims[5], strides[4];
  cuuint32_t box[5];
  for (int i = 0; i < s.rank; ++i) { dims[i] = s.dims[i]; box[i] = s.box[i]; }
  for (int i = 0; i + 1 < s.rank; ++i) strides[i] = s.strides[i];
  CUresult r = fn(out, s.dt, (cuuint32_t)s.rank, s.base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu %llu %llu box %u %u %u",
             (int)r, s.rank, (unsigned long long)s.dims[0], (unsigned long long)s.dims[1],
             (unsigned long long)s.dims[2], (unsigned long long)s.dims[3], (unsigned long long)s.dims[4], s.box[0],
             s.box[1], s.box[2]);
    g_last_cuda_error = buf;
    return false;
  }
  return true;
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
constexpr int kCntStride = 16;   // dependency counters sit 64 bytes apart (separate L2 sectors)

// Tuning / debugging knobs of the blockmix kernel, read from the environment ONCE (first call).
struct Knobs {
  int run_ahead = 2, mix_hi_only = -1, o_hint = 1, q_hint = 1, trace_cta = 0, slots = 2, reverse3 = 0;
  bool no_pack = false, no_self_prep = false;
};
const Knobs& knobs() {
  static Knobs k;
  static std::once_flag once;
  std::call_once(once, [] {
    auto geti = [](const char* name, int dflt) { const char* e = std::getenv(name); return e ? std::atoi(e) : dflt; };
    k.run_ahead = geti("MHLA_RUNAHEAD", k.run_ahead);
    if (k.run_ahead < 1) k.run_ahead = 1;
    if (k.run_ahead > 16) k.run_ahead = 16;
    k.mix_hi_only = geti("MHLA_MIX_HI_ONLY", -1);   // -1: automatic (see mhla_fwd_blockmix)
    k.o_hint = geti("MHLA_OHINT", 1);
    k.q_hint = geti("MHLA_QHINT", 1);
    k.trace_cta = geti("MHLA_TRACE_CTA", 0);
    k.slots = geti("MHLA_SLOTS", 2) == 1 ? 1 : 2;
    k.reverse3 = geti("MHLA_REVERSE3", 0);   // n > 0: fused readout walks the groups backwards, the last n groups at the very end
    if (k.reverse3 < 0) k.reverse3 = 0;
    k.no_pack = std::getenv("MHLA_NO_PACK") != nullptr;
    k.no_self_prep = std::getenv("MHLA_NO_SELF_PREP") != nullptr;
  });
  return k;
}

// ------------------------------------------------------------------------------------------------ blockmix
struct BlockmixPlan {
  int G, TW, nsub, wpad, ncols, Mp, n2_rows, n2_cols, n2_scols, kslabs, normalize, ropenorm;
  int pack, Gs, Ms;   // small M: `pack` consecutive groups are scheduled as one group of Ms = pack * M blocks (Gs = G / pack)
  int g3d, p1, p2, p3, aper, tail, rows[2], kpad[2];   // 3-D block view (desc->grid / layout)
  size_t off_S, off_St, off_den, off_W, off_cnt, off_zero, total;
};

int plan_blockmix(const mhla_blockmix_desc* d, BlockmixPlan* pl) {
  if (!d) return MHLA_ERR_INVALID_ARGUMENT;
  if (d->dtype != MHLA_BF16 && d->dtype != MHLA_FP16) return MHLA_ERR_INVALID_ARGUMENT;
  if (d->B < 1 || d->H < 1 || d->M < 1 || d->w < 1) return MHLA_ERR_UNSUPPORTED_SHAPE;
  if (d->D != 64 && d->D != 128) return MHLA_ERR_UNSUPPORTED_SHAPE;
  if (d->w > 256) return MHLA_ERR_UNSUPPORTED_SHAPE;
  if ((long long)d->B * d->H > 16383 * 8 || d->M > 65535) return MHLA_ERR_UNSUPPORTED_SHAPE;   // item FIFO field widths (checked again below)
  const int D = d->D;
  pl->G = d->B * d->H;
  pl->TW = d->w >= 128 ? 128 : (d->w + 15) / 16 * 16;
  pl->nsub = (d->w + pl->TW - 1) / pl->TW;
  pl->g3d = (d->grid[0] | d->grid[1] | d->grid[2] | d->layout[0] | d->layout[1] | d->layout[2]) != 0;
  pl->tail = 0;
  pl->rows[0] = pl->rows[1] = pl->kpad[0] = pl->kpad[1] = 0;
  if (pl->g3d) {
    const int F = d->grid[0], Hh = d->grid[1], Ww = d->grid[2], fb = d->layout[0], hb = d->layout[1], wb = d->layout[2];
    if (F < 1 || Hh < 1 || Ww < 1 || fb < 1 || hb < 1 || wb < 1) return MHLA_ERR_INVALID_ARGUMENT;
    if (F % fb || Hh % hb || Ww % wb) return MHLA_ERR_UNSUPPORTED_SHAPE;
    pl->p1 = F / fb; pl->p2 = Hh / hb; pl->p3 = Ww / wb;
    if (d->M != fb * hb * wb || d->w != pl->p1 * pl->p2 * pl->p3) return MHLA_ERR_INVALID_ARGUMENT;
    if (pl->p2 * pl->p3 > 128 || pl->p2 > 256 || pl->p3 > 256) return MHLA_ERR_UNSUPPORTED_SHAPE;
    pl->aper = 128 / (pl->p2 * pl->p3);
    if (pl->aper > pl->p1) pl->aper = pl->p1;
    pl->TW = 128;
    pl->nsub = (pl->p1 + pl->aper - 1) / pl->aper;
    if (pl->nsub > 2) return MHLA_ERR_UNSUPPORTED_SHAPE;
    for (int sidx = 0; sidx < pl->nsub; ++sidx) {
      const int a = (sidx == pl->nsub - 1) ? pl->p1 - sidx * pl->aper : pl->aper;
      pl->rows[sidx] = a * pl->p2 * pl->p3;
      pl->kpad[sidx] = (pl->rows[sidx] + 15) / 16 * 16;
    }
    pl->tail = (pl->p1 % pl->aper) != 0 ? 1 : 0;
  }
  pl->normalize = (d->flags & MHLA_FLAG_NORMALIZE) ? 1 : 0;
  pl->ropenorm = (pl->normalize && d->k_rope.ptr != nullptr) ? 1 : 0;
  pl->wpad = pl->normalize ? (pl->nsub * pl->TW) : 0;
  pl->ncols = D * D + 2 * pl->wpad;          // 16-bit elements: S_j | n_loc hi | n_loc lo
  // Small block counts waste the 128-row mixing tile: schedule `pack` consecutive (b,h) groups as ONE group whose
  // mixing matrix is block-diagonal (pack copies of the caller's matrix).  Their rows are consecutive in the workspace,
  // so nothing else changes; pack must divide the number of groups.
  pl->pack = 1;
  if (!knobs().no_pack && !pl->g3d)
    for (int pk = 128 / d->M; pk >= 2; --pk)
      if (pl->G % pk == 0) { pl->pack = pk; break; }
  pl->Gs = pl->G / pl->pack;
  if (pl->Gs > 16383) return MHLA_ERR_UNSUPPORTED_SHAPE;   // item FIFO: 14-bit group field
  pl->Ms = d->M * pl->pack;
  pl->Mp = (pl->Ms + 7) / 8 * 8;
  pl->n2_rows = (pl->Ms + 127) / 128;
  pl->n2_scols = D * D / 256;
  pl->n2_cols = pl->n2_scols + (2 * pl->wpad + 255) / 256;
  pl->kslabs = (pl->Ms + 63) / 64;
  size_t off = 0;
  const size_t GM = (size_t)pl->G * d->M;
  pl->off_S = off;   off = align_up(off + GM * pl->ncols * 2, 1024);
  pl->off_St = off;  off = align_up(off + GM * D * D * 2, 1024);
  pl->off_den = off; off = align_up(off + GM * (pl->wpad ? 2 * pl->wpad : 32) * 4, 1024);
  pl->off_W = off;   off = align_up(off + (size_t)2 * pl->Ms * pl->Mp * 2, 1024);
  pl->off_cnt = off; off = align_up(off + ((size_t)2 * pl->G * kCntStride + 128) * 4, 1024);   // + item tickets, flags
  pl->off_zero = off; off += 4096;   // zeros (3-D block view: pad rows of a tile); never written by the kernels
  pl->total = off;
  return MHLA_OK;
}

bool t5_ok(const mhla_tensor5& t) {
  if ((reinterpret_cast<uintptr_t>(t.ptr) & 15) != 0) return false;
  return t.stride_b % 8 == 0 && t.stride_h % 8 == 0 && t.stride_m % 8 == 0 && t.stride_w % 8 == 0;
}

MapSpec spec_t5(const mhla_tensor5& t, const mhla_blockmix_desc* d, int TW) {
  MapSpec s{};
  s.dt = d->dtype == MHLA_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  s.rank = 5;
  s.base = const_cast<void*>(t.ptr);
  const uint64_t dims[5] = {(uint64_t)d->D, (uint64_t)d->w, (uint64_t)d->M, (uint64_t)d->H, (uint64_t)d->B};
  int64_t str[4] = {t.stride_w, t.stride_m, t.stride_h, t.stride_b};
  for (int i = 0; i < 5; ++i) s.dims[i] = dims[i];
  uint64_t prev = (uint64_t)d->D * 2;  // a valid (16-byte multiple) stand-in for size-1 dimensions
  for (int i = 0; i < 4; ++i) {
    uint64_t bytes = (uint64_t)str[i] * 2;
    if (dims[i + 1] == 1 || bytes == 0) bytes = prev;
    s.strides[i] = bytes;
    prev = bytes * dims[i + 1];
  }
  s.box[0] = 64; s.box[1] = (uint32_t)TW; s.box[2] = s.box[3] = s.box[4] = 1;
  return s;
}

struct CacheEntry {
  mhla_blockmix_desc key;
  mhla::BlockmixParams params;
};
std::mutex g_cache_mu;
std::vector<CacheEntry> g_cache;
struct CausalCacheEntry {
  mhla_causal_desc key;
  mhla::CausalParams params;
};
std::vector<CausalCacheEntry> g_causal_cache;

int build_blockmix_params(const mhla_blockmix_desc* d, const BlockmixPlan& pl, mhla::BlockmixParams* P) {
  const int D = d->D;
  uint8_t* ws = static_cast<uint8_t*>(d->workspace);
  uint16_t* S = reinterpret_cast<uint16_t*>(ws + pl.off_S);
  void* St = ws + pl.off_St;
  float* den = reinterpret_cast<float*>(ws + pl.off_den);
  uint16_t* Wp = reinterpret_cast<uint16_t*>(ws + pl.off_W);
  const uint64_t GM = (uint64_t)pl.G * d->M;
  const CUtensorMapDataType dt16 =
      d->dtype == MHLA_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const bool rope = d->k_rope.ptr != nullptr;

  if (!pl.g3d) {
    if (!encode_map(&P->tmK, spec_t5(rope ? d->k_rope : d->k, d, pl.TW))) return MHLA_ERR_CUDA;
    if (!encode_map(&P->tmV, spec_t5(d->v, d, pl.TW))) return MHLA_ERR_CUDA;
    if (!encode_map(&P->tmKn, spec_t5(d->k, d, pl.TW))) return MHLA_ERR_CUDA;
    if (!encode_map(&P->tmQn, spec_t5(d->q, d, pl.TW))) return MHLA_ERR_CUDA;
    if (!encode_map(&P->tmQr, spec_t5(d->q_rope.ptr ? d->q_rope : d->q, d, pl.TW))) return MHLA_ERR_CUDA;
    if (!encode_map(&P->tmO, spec_t5(d->out, d, pl.TW))) return MHLA_ERR_CUDA;
  }
  {
    MapSpec s{dt16, 3, S, {(uint64_t)D, (uint64_t)D, GM}, {(uint64_t)D * 2, (uint64_t)pl.ncols * 2},
              {64, (uint32_t)D, 1}};
    if (!encode_map(&P->tmSst, s)) return MHLA_ERR_CUDA;
  }
  {
    MapSpec s{dt16, 3, S, {(uint64_t)pl.ncols, (uint64_t)pl.Ms, (uint64_t)pl.Gs},
              {(uint64_t)pl.ncols * 2, (uint64_t)pl.Ms * pl.ncols * 2}, {64, 64, 1}};
    if (!encode_map(&P->tmSld, s)) return MHLA_ERR_CUDA;
  }
  {
    MapSpec s{dt16, 3, Wp, {(uint64_t)pl.Mp, (uint64_t)pl.Ms, 2}, {(uint64_t)pl.Mp * 2, (uint64_t)pl.Ms * pl.Mp * 2},
              {64, 128, 1}};
    if (!encode_map(&P->tmW, s)) return MHLA_ERR_CUDA;
  }
  {
    MapSpec s{dt16, 3, St, {(uint64_t)D * D, (uint64_t)pl.Ms, (uint64_t)pl.Gs},
              {(uint64_t)D * D * 2, (uint64_t)pl.Ms * D * D * 2}, {64, 128, 1}};
    if (!encode_map(&P->tmStst, s)) return MHLA_ERR_CUDA;
  }
  {
    const uint64_t wp = pl.wpad ? 2 * pl.wpad : 32;
    MapSpec s{CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, den, {wp, (uint64_t)pl.Ms, (uint64_t)pl.Gs},
              {wp * 4, (uint64_t)pl.Ms * wp * 4}, {32, 128, 1}};
    if (!encode_map(&P->tmDen, s)) return MHLA_ERR_CUDA;
  }
  {
    MapSpec s{dt16, 3, St, {(uint64_t)D, (uint64_t)D, GM}, {(uint64_t)D * 2, (uint64_t)D * D * 2},
              {64, (uint32_t)D, 1}};
    if (!encode_map(&P->tmStld, s)) return MHLA_ERR_CUDA;
  }
  if (pl.g3d) {
    // token-major [B, F*H*W, heads, D] viewed as (d, W, H, B*F, heads); one box = (64 channels, p3, p2, a frames, 1 head)
    const mhla_tensor5* ts[6] = {rope ? &d->k_rope : &d->k, &d->v, &d->k, &d->q, d->q_rope.ptr ? &d->q_rope : &d->q, &d->out};
    const uint64_t F = d->grid[0], Hh = d->grid[1], Ww = d->grid[2];
    for (int t = 0; t < 6; ++t) {
      const uint64_t tok = (uint64_t)ts[t]->stride_w * 2;
      for (int v = 0; v < 2; ++v) {
        const uint32_t a = v == 0 ? (uint32_t)pl.aper : (uint32_t)(pl.p1 - (pl.nsub - 1) * pl.aper);
        MapSpec s{dt16, 5, const_cast<void*>(ts[t]->ptr), {(uint64_t)D, Ww, Hh, (uint64_t)d->B * F, (uint64_t)d->H},
                  {tok, tok * Ww, tok * Ww * Hh, (uint64_t)ts[t]->stride_h * 2}, {64, (uint32_t)pl.p3, (uint32_t)pl.p2, a, 1}};
        if (d->H == 1) s.strides[3] = tok * Ww * Hh * (uint64_t)d->B * F;   // a valid stand-in for the size-1 dimension
        if (!encode_map(&P->tm3[t][v], s)) return MHLA_ERR_CUDA;
      }
    }
    P->g3_zero = ws + pl.off_zero;
    P->g3_F = d->grid[0]; P->g3_hb = d->layout[1]; P->g3_wb = d->layout[2];
    P->g3_p1 = pl.p1; P->g3_p2 = pl.p2; P->g3_p3 = pl.p3; P->g3_aper = pl.aper; P->g3_tail = pl.tail;
    for (int i = 0; i < 2; ++i) { P->g3_rows[i] = pl.rows[i]; P->g3_kpad[i] = pl.kpad[i]; }
  }
  P->ws_S = S;
  P->ws_St = reinterpret_cast<uint16_t*>(St);
  P->den = den;
  P->counters = reinterpret_cast<uint32_t*>(ws + pl.off_cnt);
  P->wscale = reinterpret_cast<const float*>(P->counters + (size_t)2 * pl.Gs * kCntStride + 48);
  P->rms_w = d->out_rms_weight; P->rms_eps = d->out_rms_eps;
  P->post_gate = static_cast<const uint16_t*>(d->out_gate.ptr);
  P->post_add = static_cast<const uint16_t*>(d->out_add.ptr);
  P->pg_sb = d->out_gate.stride_b; P->pg_sh = d->out_gate.stride_h; P->pg_sm = d->out_gate.stride_m; P->pg_sw = d->out_gate.stride_w;
  P->pa_sb = d->out_add.stride_b; P->pa_sh = d->out_add.stride_h; P->pa_sm = d->out_add.stride_m; P->pa_sw = d->out_add.stride_w;
  P->g3_H = d->grid[1]; P->g3_W = d->grid[2];
  P->mix = d->mix; P->mix_ld = d->mix_ld; P->w_planes = Wp; P->Mp = pl.Mp; P->self_prep = 0;
  P->G = pl.Gs; P->H = d->H; P->M = pl.Ms; P->pack = pl.pack; P->M0 = d->M; P->w = d->w; P->TW = pl.TW; P->nsub = pl.nsub;
  P->ncols = pl.ncols; P->wpad = pl.wpad;
  P->n2_rows = pl.n2_rows; P->n2_cols = pl.n2_cols; P->n2_scols = pl.n2_scols; P->kslabs = pl.kslabs;
  P->normalize = pl.normalize; P->ropenorm = pl.ropenorm; P->is_fp16 = d->dtype == MHLA_FP16;
  P->mode = 0; P->run_ahead = 2; P->cnt_stride = kCntStride;
  P->eps = d->eps;
  (void)Wp;
  return MHLA_OK;
}

unsigned long long* g_prof_buffer = nullptr;   // debug: per-CTA role counters (mhla_debug_set_profile_buffer)

// Per-device state (SM count, dynamic shared-memory opt-in per kernel instantiation), guarded by g_cache_mu.
constexpr int kMaxDevices = 64;
struct DeviceState {
  int sms = 0;          // 0: not queried yet; < 0: not an sm_100 device
  bool attr64 = false, attr128 = false, attr64g = false, attr128g = false, attr_smalln = false;
  bool attr_post[4] = {false, false, false, false};   // POST instantiations: [D == 128][3-D view]
};
DeviceState g_dev[kMaxDevices];

// resolves the current device; returns MHLA_OK and its state, or an error
int device_state(DeviceState** out) {
  int dev = 0;
  if (!cuda_ok(cudaGetDevice(&dev), "cudaGetDevice")) return MHLA_ERR_CUDA;
  if (dev < 0 || dev >= kMaxDevices) return MHLA_ERR_NO_DEVICE;
  std::lock_guard<std::mutex> lk(g_cache_mu);
  DeviceState& st = g_dev[dev];
  if (st.sms == 0) {
    int major = 0, sms = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    st.sms = (major == 10 && sms > 0) ? sms : -1;
  }
  if (st.sms < 0) return MHLA_ERR_NO_DEVICE;
  *out = &st;
  return MHLA_OK;
}

struct SmallNCacheEntry {
  mhla_blockmix_desc key;
  mhla::SmallNParams params;
};
std::vector<SmallNCacheEntry> g_smalln_cache;

// Short sequences (DiT / ViT): the whole (b,h) unit fits one CTA - see smalln_kernel.cuh.
bool smalln_eligible(const mhla_blockmix_desc* d) {
  if (d->D != 64 || d->M > mhla::kSnMaxM || (long long)d->M * d->w > mhla::kSnRows) return false;
  if (d->q_rope.ptr || d->k_rope.ptr || d->out_rms_weight || d->out_gate.ptr) return false;
  if (d->grid[0] | d->grid[1] | d->grid[2] | d->layout[0] | d->layout[1] | d->layout[2]) return false;
  if (d->flags & (MHLA_FLAG_NO_SMALLN | MHLA_FLAG_UNFUSED | MHLA_FLAG_TWO_LAUNCH | MHLA_FLAG_FUSED | MHLA_FLAG_STOP_AFTER_P1 |
                  MHLA_FLAG_STOP_AFTER_P2 | MHLA_FLAG_ONLY_P3 | MHLA_FLAG_ONLY_P2))
    return false;
  return true;
}

int launch_smalln(const mhla_blockmix_desc* d, DeviceState* dst, cudaStream_t stream) {
  mhla::SmallNParams P;
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    bool hit = false;
    for (auto& e : g_smalln_cache)
      if (std::memcmp(&e.key, d, sizeof(*d)) == 0) { P = e.params; hit = true; break; }
    if (!hit) {
      std::memset(&P, 0, sizeof P);
      const mhla_tensor5* ts[4] = {&d->q, &d->k, &d->v, &d->out};
      CUtensorMap* maps[4] = {&P.tmQ, &P.tmK, &P.tmV, &P.tmO};
      for (int i = 0; i < 4; ++i) {
        MapSpec s = spec_t5(*ts[i], d, d->w);
        s.box[2] = (uint32_t)d->M;               // one box = all blocks of one (b,h) unit, rows in (block, token) order
        if (!encode_map(maps[i], s)) return MHLA_ERR_CUDA;
      }
      P.mix = d->mix; P.mix_ld = d->mix_ld;
      P.G = d->B * d->H; P.H = d->H; P.M = d->M; P.w = d->w; P.N = d->M * d->w;
      P.normalize = (d->flags & MHLA_FLAG_NORMALIZE) ? 1 : 0;
      P.is_fp16 = d->dtype == MHLA_FP16;
      P.eps = d->eps;
      P.post_add = static_cast<const uint16_t*>(d->out_add.ptr);
      P.pa_sb = d->out_add.stride_b; P.pa_sh = d->out_add.stride_h; P.pa_sm = d->out_add.stride_m; P.pa_sw = d->out_add.stride_w;
      if (g_smalln_cache.size() >= 32) g_smalln_cache.erase(g_smalln_cache.begin());
      SmallNCacheEntry e;
      std::memcpy(&e.key, d, sizeof(*d));
      e.params = P;
      g_smalln_cache.push_back(e);
    }
    if (!dst->attr_smalln) {
      if (!cuda_ok(cudaFuncSetAttribute(mhla::smalln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mhla::kSnSmemAlloc),
                   "cudaFuncSetAttribute(smalln)"))
        return MHLA_ERR_CUDA;
      dst->attr_smalln = true;
    }
  }
  P.prof = g_prof_buffer;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(P.G < dst->sms ? P.G : dst->sms);
  cfg.blockDim = dim3(mhla::kSnThreads);
  cfg.dynamicSmemBytes = mhla::kSnSmemAlloc;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  if (!cuda_ok(cudaLaunchKernelEx(&cfg, mhla::smalln_kernel, P), "cudaLaunchKernelEx(smalln)")) return MHLA_ERR_CUDA;
  if (!cuda_ok(cudaGetLastError(), "kernel launch")) return MHLA_ERR_CUDA;
  g_last_launches = 1;
  return MHLA_OK;
}

}  // namespace

extern "C" {

int mhla_abi_version(void) { return MHLA_B200_ABI_VERSION; }

const char* mhla_strerror(int status) {
  switch (status) {
    case MHLA_OK: return "ok";
    case MHLA_ERR_INVALID_ARGUMENT: return "invalid argument";
    case MHLA_ERR_UNSUPPORTED_SHAPE: return "shape outside the supported envelope";
    case MHLA_ERR_ALIGNMENT: return "pointer or stride alignment";
    case MHLA_ERR_WORKSPACE: return "workspace missing or too small";
    case MHLA_ERR_CUDA: return "CUDA error";
    case MHLA_ERR_NO_DEVICE: return "current device is not an sm_100 GPU";
    default: return "unknown status";
  }
}

const char* mhla_last_cuda_error(void) { return g_last_cuda_error.c_str(); }
int mhla_last_launch_count(void) { return g_last_launches; }

size_t mhla_blockmix_workspace_bytes(const mhla_blockmix_desc* desc) {
  BlockmixPlan pl;
  if (plan_blockmix(desc, &pl) != MHLA_OK) return 0;
  return pl.total;
}

/* 1 when mhla_fwd_blockmix(desc) will touch the workspace, 0 when the shape takes the short-sequence kernel (whole unit
 * on chip, csrc/smalln_kernel.cuh) and workspace may be NULL. */
int mhla_blockmix_needs_workspace(const mhla_blockmix_desc* desc) {
  BlockmixPlan pl;
  if (plan_blockmix(desc, &pl) != MHLA_OK) return 1;
  return smalln_eligible(desc) ? 0 : 1;
}

int mhla_blockmix_workspace_layout(const mhla_blockmix_desc* desc, size_t out[8]) {
  BlockmixPlan pl;
  int rc = plan_blockmix(desc, &pl);
  if (rc != MHLA_OK) return rc;
  out[0] = pl.off_S; out[1] = pl.off_St; out[2] = pl.off_den; out[3] = pl.off_W; out[4] = pl.off_cnt;
  out[5] = (size_t)pl.ncols; out[6] = (size_t)pl.wpad; out[7] = (size_t)pl.Mp;
  return MHLA_OK;
}

int mhla_fwd_blockmix(const mhla_blockmix_desc* d, void* stream_) {
  BlockmixPlan pl;
  int rc = plan_blockmix(d, &pl);
  if (rc != MHLA_OK) return rc;
  if (!d->q.ptr || !d->k.ptr || !d->v.ptr || !d->out.ptr || !d->mix) return MHLA_ERR_INVALID_ARGUMENT;
  if ((d->q_rope.ptr == nullptr) != (d->k_rope.ptr == nullptr)) return MHLA_ERR_INVALID_ARGUMENT;
  const bool small = smalln_eligible(d);      // short sequences need no workspace
  if (!small && (!d->workspace || d->workspace_bytes < pl.total)) return MHLA_ERR_WORKSPACE;
  if (!small && (reinterpret_cast<uintptr_t>(d->workspace) & 1023) != 0) return MHLA_ERR_ALIGNMENT;
  if (!t5_ok(d->q) || !t5_ok(d->k) || !t5_ok(d->v) || !t5_ok(d->out)) return MHLA_ERR_ALIGNMENT;
  if (d->q_rope.ptr && (!t5_ok(d->q_rope) || !t5_ok(d->k_rope))) return MHLA_ERR_ALIGNMENT;
  if ((d->out_gate.ptr && !t5_ok(d->out_gate)) || (d->out_add.ptr && !t5_ok(d->out_add))) return MHLA_ERR_ALIGNMENT;
  if (d->mix_ld < d->M) return MHLA_ERR_INVALID_ARGUMENT;
  if (pl.g3d) {   // token-major tensors: the batch must follow the token axis directly ((b, f) is ONE tensor-map dimension)
    const long long ntok = (long long)d->grid[0] * d->grid[1] * d->grid[2];
    const mhla_tensor5* ts[6] = {&d->q, &d->k, &d->v, &d->out, &d->q_rope, &d->k_rope};
    for (const mhla_tensor5* t : ts)
      if (t->ptr && d->B > 1 && t->stride_b != ntok * t->stride_w) return MHLA_ERR_ALIGNMENT;
  }
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DeviceState* dst = nullptr;
  rc = device_state(&dst);
  if (rc != MHLA_OK) return rc;
  const int num_sms = dst->sms;
  if (small) return launch_smalln(d, dst, stream);

  mhla::BlockmixParams P;
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    bool hit = false;
    for (auto& e : g_cache)
      if (std::memcmp(&e.key, d, sizeof(*d)) == 0) { P = e.params; hit = true; break; }
    if (!hit) {
      std::memset(&P, 0, sizeof P);
      rc = build_blockmix_params(d, pl, &P);
      if (rc != MHLA_OK) return rc;
      if (g_cache.size() >= 32) g_cache.erase(g_cache.begin());
      CacheEntry e;
      std::memcpy(&e.key, d, sizeof(*d));
      e.params = P;
      g_cache.push_back(e);
    }
  }

  const Knobs& kn = knobs();
  P.prof = g_prof_buffer;
  P.run_ahead = kn.run_ahead;
  P.trace_cta = kn.trace_cta;
  // bf16, D = 128 (Wan): the S columns of the block mixing take the 8-bit hi plane of the mixing matrix only - what the
  // reference's 1x1 conv computes under the bf16 autocast its sampler and trainer run in (its weight is rounded to bf16;
  // SURVEY.md 8a row B3).  Halves the matrix bytes and MMAs of the 1536 mixing items of a Wan layer: 132 -> 120 us,
  // RMS error vs the fp32 oracle 2.9e-3 -> 3.3e-3 (budget 5e-3).  The normaliser columns always use hi + lo.
  P.mix_hi_only = kn.mix_hi_only >= 0 ? kn.mix_hi_only : ((d->D == 128 && d->dtype == MHLA_BF16) ? 1 : 0);
  P.o_hint = kn.o_hint;
  P.q_hint = kn.q_hint;
  P.window = 8; P.np2 = 0; P.policy = 1; P.pf_dist = 0; P.reverse3 = kn.reverse3;   // (round-1 tuning options, fixed at their best values)
  // fused gate / additive term: separate instantiations, so the plain kernels' code is untouched by them
  const bool post = d->out_gate.ptr != nullptr || d->out_add.ptr != nullptr;
  auto kern = pl.g3d ? (d->D == 64 ? mhla::blockmix_kernel<64, true> : mhla::blockmix_kernel<128, true>)
                     : (d->D == 64 ? mhla::blockmix_kernel<64, false> : mhla::blockmix_kernel<128, false>);
  if (post)
    kern = pl.g3d ? (d->D == 64 ? mhla::blockmix_kernel<64, true, true> : mhla::blockmix_kernel<128, true, true>)
                  : (d->D == 64 ? mhla::blockmix_kernel<64, false, true> : mhla::blockmix_kernel<128, false, true>);
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    bool& attr = post ? dst->attr_post[(d->D == 128 ? 2 : 0) + (pl.g3d ? 1 : 0)]
                      : (pl.g3d ? (d->D == 64 ? dst->attr64g : dst->attr128g) : (d->D == 64 ? dst->attr64 : dst->attr128));
    if (!attr) {
      if (!cuda_ok(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, mhla::kSmemAlloc),
                   "cudaFuncSetAttribute"))
        return MHLA_ERR_CUDA;
      // The fused kernel's CTAs wait for each other (self-prep announcement, per-group counters), so the whole grid has
      // to be co-resident: check once per device that a CTA of this instantiation fits an SM at all (grid <= #SMs, 1 CTA
      // per SM).  Other work sharing the GPU only delays the waits - they are bounded in the seconds range, a whole
      // launch takes < 1 ms.
      int occ = 0;
      if (!cuda_ok(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, mhla::kThreads, mhla::kSmemAlloc),
                   "cudaOccupancyMaxActiveBlocksPerMultiprocessor"))
        return MHLA_ERR_CUDA;
      if (occ < 1) { g_last_cuda_error = "blockmix kernel does not fit one SM on this device"; return MHLA_ERR_NO_DEVICE; }
      attr = true;
    }
  }

  uint8_t* ws = static_cast<uint8_t*>(d->workspace);
  int launches = 0;
  const bool dbg_phase = (d->flags & (MHLA_FLAG_STOP_AFTER_P1 | MHLA_FLAG_STOP_AFTER_P2 | MHLA_FLAG_ONLY_P3 | MHLA_FLAG_ONLY_P2)) != 0;
  const bool single = (d->flags & MHLA_FLAG_FUSED) || !(dbg_phase || (d->flags & (MHLA_FLAG_UNFUSED | MHLA_FLAG_TWO_LAUNCH)));
  P.self_prep = (single && (d->flags & MHLA_FLAG_WS_PERSISTENT) && !kn.no_self_prep) ? 1 : 0;
  if (!P.self_prep && pl.g3d &&
      !cuda_ok(cudaMemsetAsync(static_cast<uint8_t*>(d->workspace) + pl.off_zero, 0, 4096, stream), "cudaMemsetAsync(zero rows)"))
    return MHLA_ERR_CUDA;
  if (!P.self_prep) {
    mhla::prep_mix_scaled_kernel<<<16, 1024, 0, stream>>>(d->mix, (long long)d->mix_ld,
                                                       reinterpret_cast<uint16_t*>(ws + pl.off_W), pl.Ms, pl.Mp, d->M,
                                                       d->dtype == MHLA_FP16, const_cast<float*>(P.wscale), P.counters,
                                                       2 * pl.Gs * kCntStride + 48);
    // (words 48.. of the tail: wscale and the self_prep flags - the prologue kernel leaves them alone)
    ++launches;
  }
  const long long n1 = pl.Ms, n2 = (long long)pl.n2_rows * pl.n2_cols, n3 = pl.Ms;
  auto launch_pdl = [&](int grid) -> bool {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(mhla::kThreads);
    cfg.dynamicSmemBytes = mhla::kSmemAlloc;
    cfg.stream = stream;
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // prologue overlaps the previous kernel's tail
    attrs[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 1;
    return cuda_ok(cudaLaunchKernelEx(&cfg, kern, P), "cudaLaunchKernelEx");
  };
  if (!single) {
    // MHLA_FLAG_UNFUSED / the debugging flags: plain phase-by-phase launches of the same kernel, chained with PDL
    int last = (d->flags & MHLA_FLAG_STOP_AFTER_P1) ? 1 : ((d->flags & MHLA_FLAG_STOP_AFTER_P2) ? 2 : 3);
    int first = (d->flags & MHLA_FLAG_ONLY_P3) ? 3 : 1;
    if (d->flags & MHLA_FLAG_ONLY_P2) first = last = 2;
    for (int mode = first; mode <= last; ++mode) {
      P.mode = mode;
      // P1 with D = 64 only stages 8 KB (S) per item: give the ring a sixth stage instead (a multiple of the 3 stages
      // per item keeps the long-lived Q stage out of the K/V recycling path)
      const bool small_staging = (mode == 1 && d->D == 64);
      P.slot_bytes = small_staging ? 8192 : 16384;
      P.ring_stages = small_staging ? 6 : 5;
      P.slots_per_wg = 2;
      const long long items = (long long)pl.Gs * (mode == 1 ? n1 : (mode == 2 ? n2 : n3));
      const int grid = (int)(items < num_sms ? items : num_sms);
      if (!launch_pdl(grid)) return MHLA_ERR_CUDA;
      ++launches;
    }
  } else {
    P.mode = 0;
    P.slot_bytes = 16384;
    P.slots_per_wg = kn.slots;
    P.ring_stages = P.slots_per_wg == 1 ? 6 : 5;   // one staging slot per warpgroup buys a sixth ring stage
    const long long items = (long long)pl.Gs * (n1 + n2 + n3);
    // 1 CTA per SM (224 KB of shared memory): the grid never exceeds what is co-resident on an otherwise idle device
    const int grid = (int)(items < num_sms ? items : num_sms);
    if (!launch_pdl(grid)) return MHLA_ERR_CUDA;
    ++launches;
  }
  if (!cuda_ok(cudaGetLastError(), "kernel launch")) return MHLA_ERR_CUDA;
  g_last_launches = launches;
  return MHLA_OK;
}

/* Debug hook (not part of the stable ABI): device buffer of [#SMs][16] uint64 that receives per-CTA wait-cycle counters
 * of the blockmix kernel's warp roles; NULL switches the instrumentation off. */
void mhla_debug_set_profile_buffer(void* dev_ptr) { g_prof_buffer = static_cast<unsigned long long*>(dev_ptr); }

/* Debug hook (not part of the stable ABI): host-mapped (pinned, zero-copy) buffer of 1 + 148*64 uint64 that receives a
 * record from every wait that hits its time bound before the kernel traps; NULL switches it off.  Applies to the
 * current device. */
int mhla_debug_set_diag_buffer(void* host_mapped_ptr) {
  unsigned long long* p = static_cast<unsigned long long*>(host_mapped_ptr);
  if (!cuda_ok(cudaMemcpyToSymbol(mhla::g_mhla_diag, &p, sizeof p), "cudaMemcpyToSymbol(g_mhla_diag)")) return MHLA_ERR_CUDA;
  return MHLA_OK;
}

int mhla_blockmix_workspace_init(const mhla_blockmix_desc* d, void* stream_) {
  BlockmixPlan pl;
  int rc = plan_blockmix(d, &pl);
  if (rc != MHLA_OK) return rc;
  if (!d->workspace || d->workspace_bytes < pl.total) return MHLA_ERR_WORKSPACE;
  uint8_t* ws = static_cast<uint8_t*>(d->workspace);
  if (!cuda_ok(cudaMemsetAsync(ws + pl.off_cnt, 0, pl.total - pl.off_cnt, static_cast<cudaStream_t>(stream_)),
               "cudaMemsetAsync"))
    return MHLA_ERR_CUDA;
  return MHLA_OK;
}

namespace { __global__ void stall_selftest_kernel() { mhla::report_stall_diag(99, 1, 2); } }
/* Debug hook: launches one thread that reports a stall (code 99) and traps - checks the diagnostics path end to end. */
int mhla_debug_trigger_stall(void* stream_) {
  stall_selftest_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream_)>>>();
  return cuda_ok(cudaGetLastError(), "stall_selftest_kernel") ? MHLA_OK : MHLA_ERR_CUDA;
}

int mhla_wan_prep(const mhla_wan_prep_desc* d, void* stream_) {
  if (!d || !d->xq || !d->xk || !d->q_rope || !d->k_rope) return MHLA_ERR_INVALID_ARGUMENT;
  if ((d->q_plain == nullptr) != (d->k_plain == nullptr) || (d->cos_table == nullptr) != (d->sin_table == nullptr))
    return MHLA_ERR_INVALID_ARGUMENT;
  if (d->rows < 1 || d->N < 1 || d->C < 8 || d->C % 8 || d->D < 8 || d->D % 8 || d->C % d->D || d->C > 8192)
    return MHLA_ERR_UNSUPPORTED_SHAPE;
  if (d->in_dtype < 0 || d->in_dtype > 2 || (d->out_dtype != MHLA_BF16 && d->out_dtype != MHLA_FP16)) return MHLA_ERR_INVALID_ARGUMENT;
  const uintptr_t al = reinterpret_cast<uintptr_t>(d->xq) | reinterpret_cast<uintptr_t>(d->xk) |
                       reinterpret_cast<uintptr_t>(d->q_rope) | reinterpret_cast<uintptr_t>(d->k_rope) |
                       reinterpret_cast<uintptr_t>(d->q_plain) | reinterpret_cast<uintptr_t>(d->k_plain) |
                       reinterpret_cast<uintptr_t>(d->cos_table) | reinterpret_cast<uintptr_t>(d->sin_table) |
                       reinterpret_cast<uintptr_t>(d->wq) | reinterpret_cast<uintptr_t>(d->wk);
  if ((al & 15) != 0 || d->ld_in % 8 != 0 || (d->D / 2) % 4 != 0) return MHLA_ERR_ALIGNMENT;
  mhla::WanPrepParams P{};
  P.xq = d->xq; P.xk = d->xk; P.q_rope = d->q_rope; P.k_rope = d->k_rope; P.q_plain = d->q_plain; P.k_plain = d->k_plain;
  P.wq = d->wq; P.wk = d->wk; P.cos_t = d->cos_table; P.sin_t = d->sin_table; P.ld_in = d->ld_in;
  P.rows = d->rows; P.N = d->N; P.C = d->C; P.D = d->D; P.in_dtype = d->in_dtype; P.out_fp16 = d->out_dtype == MHLA_FP16;
  P.eps_norm = d->eps_norm; P.eps = d->eps;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int threads = (((d->C / 8) + 1) / 2 + 31) / 32 * 32;   // two 8-channel chunks per thread
  DeviceState* dst = nullptr;
  int rc = device_state(&dst);
  if (rc != MHLA_OK) return rc;
  // shared-memory ring of staged rows (see the kernel): as many stages as fit 47 KB (no opt-in needed), at most 8, at
  // least 2 - very wide fp32 rows take the opt-in; persistent grid = exactly the CTAs that are resident at once
  auto kern = d->in_dtype == 0 ? mhla::wan_prep_kernel<0> : (d->in_dtype == 1 ? mhla::wan_prep_kernel<1> : mhla::wan_prep_kernel<2>);
  const size_t rowb = (size_t)d->C * (d->in_dtype == 2 ? 4 : 2), angb = d->cos_table ? (size_t)d->D * 2 : 0;
  const size_t stage = align_up(2 * rowb + 2 * angb, 128);
  int stages = (int)((47 * 1024) / stage);
  if (stages > 4) stages = 4;     // (the kernel is issue-bound, not latency-bound: more CTAs per SM beat deeper rings)
  if (stages < 2) stages = 2;
  const size_t smem = stages * stage + 8 * (size_t)stages;
  if (smem > 48 * 1024 &&
      !cuda_ok(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(wan_prep)"))
    return MHLA_ERR_CUDA;
  P.stages = stages; P.stage_bytes = (int)stage;
  int per_sm = 0;
  if (!cuda_ok(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem), "cudaOccupancyMaxActiveBlocksPerMultiprocessor"))
    return MHLA_ERR_CUDA;
  if (per_sm < 1) per_sm = 1;
  const long long want = (long long)dst->sms * per_sm;
  const int grid = (int)(d->rows < want ? d->rows : want);
  kern<<<grid, threads, smem, stream>>>(P);
  if (!cuda_ok(cudaGetLastError(), "wan_prep_kernel")) return MHLA_ERR_CUDA;
  g_last_launches = 1;
  return MHLA_OK;
}

int mhla_gated_rmsnorm(const mhla_gated_rmsnorm_desc* d, void* stream_) {
  if (!d || !d->x || !d->out) return MHLA_ERR_INVALID_ARGUMENT;
  if (d->rows < 1 || (d->D != 64 && d->D != 128 && d->D != 256)) return MHLA_ERR_UNSUPPORTED_SHAPE;
  if (d->dtype != MHLA_BF16 && d->dtype != MHLA_FP16) return MHLA_ERR_INVALID_ARGUMENT;
  const uintptr_t al = reinterpret_cast<uintptr_t>(d->x) | reinterpret_cast<uintptr_t>(d->g) | reinterpret_cast<uintptr_t>(d->out);
  if ((al & 15) != 0 || d->ld_x % 8 != 0 || (d->g && d->ld_g % 8 != 0) || d->ld_x < d->D || (d->g && d->ld_g < d->D))
    return MHLA_ERR_ALIGNMENT;
  DeviceState* dst = nullptr;
  int rc = device_state(&dst);
  if (rc != MHLA_OK) return rc;
  mhla::GatedNormParams P{d->x, d->g, d->out, d->weight, (long long)d->rows, (long long)d->ld_x, (long long)d->ld_g, d->D,
                          d->dtype == MHLA_FP16, d->eps};
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int tpr = d->D / 8, rpc = 256 / tpr;
  const long long want = (d->rows + rpc - 1) / rpc, cap = (long long)dst->sms * 8;
  const int grid = (int)(want < cap ? want : cap);
  if (tpr == 8) mhla::gated_norm_kernel<8><<<grid, 256, 0, stream>>>(P);
  else if (tpr == 16) mhla::gated_norm_kernel<16><<<grid, 256, 0, stream>>>(P);
  else mhla::gated_norm_kernel<32><<<grid, 256, 0, stream>>>(P);
  if (!cuda_ok(cudaGetLastError(), "gated_norm_kernel")) return MHLA_ERR_CUDA;
  g_last_launches = 1;
  return MHLA_OK;
}

int mhla_gate_add(const mhla_gate_add_desc* d, void* stream_) {
  if (!d || !d->x || !d->out) return MHLA_ERR_INVALID_ARGUMENT;
  if (d->rows < 1 || d->C < 8 || d->C % 8) return MHLA_ERR_UNSUPPORTED_SHAPE;
  if (d->dtype != MHLA_BF16 && d->dtype != MHLA_FP16) return MHLA_ERR_INVALID_ARGUMENT;
  const uintptr_t al = reinterpret_cast<uintptr_t>(d->x) | reinterpret_cast<uintptr_t>(d->g) |
                       reinterpret_cast<uintptr_t>(d->add) | reinterpret_cast<uintptr_t>(d->out);
  if ((al & 15) != 0 || d->ld_x % 8 || d->ld_out % 8 || (d->g && d->ld_g % 8) || (d->add && d->ld_add % 8) ||
      d->ld_x < d->C || d->ld_out < d->C || (d->g && d->ld_g < d->C) || (d->add && d->ld_add < d->C))
    return MHLA_ERR_ALIGNMENT;
  DeviceState* dst = nullptr;
  int rc = device_state(&dst);
  if (rc != MHLA_OK) return rc;
  mhla::GateAddParams P{d->x, d->g, d->add, d->out, (long long)d->rows, (long long)d->ld_x, (long long)d->ld_g,
                        (long long)d->ld_add, (long long)d->ld_out, d->C, d->dtype == MHLA_FP16};
  const long long want = (d->rows * (d->C / 8) + 255) / 256, cap = (long long)dst->sms * 8;
  const int grid = (int)(want < cap ? want : cap);
  mhla::gate_add_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(P);
  if (!cuda_ok(cudaGetLastError(), "gate_add_kernel")) return MHLA_ERR_CUDA;
  g_last_launches = 1;
  return MHLA_OK;
}

int mhla_dwconv3d(const mhla_dwconv3d_desc* d, void* stream_) {
  if (!d || !d->x || !d->wt || !d->out) return MHLA_ERR_INVALID_ARGUMENT;
  if (d->B < 1 || d->F < 1 || d->H < 1 || d->W < 1 || d->C < 8 || d->C % 8) return MHLA_ERR_UNSUPPORTED_SHAPE;
  if (d->dtype != MHLA_BF16 && d->dtype != MHLA_FP16) return MHLA_ERR_INVALID_ARGUMENT;
  const uintptr_t al = reinterpret_cast<uintptr_t>(d->x) | reinterpret_cast<uintptr_t>(d->wt) |
                       reinterpret_cast<uintptr_t>(d->bias) | reinterpret_cast<uintptr_t>(d->out);
  if ((al & 15) != 0 || d->ld_x % 8 || d->ld_x < d->C) return MHLA_ERR_ALIGNMENT;
  DeviceState* dst = nullptr;
  int rc = device_state(&dst);
  if (rc != MHLA_OK) return rc;
  mhla::DwConv3dParams P{d->x, d->out, d->wt, d->bias, (long long)d->ld_x, d->B, d->F, d->H, d->W, d->C, d->dtype == MHLA_FP16};
  const long long total = (long long)d->B * d->F * d->H * ((d->W + mhla::kDwTile - 1) / mhla::kDwTile) * (d->C / 8);
  const long long want = (total + 255) / 256, cap = (long long)dst->sms * 16;
  const int grid = (int)(want < cap ? want : cap);
  mhla::dwconv3d_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(P);
  if (!cuda_ok(cudaGetLastError(), "dwconv3d_kernel")) return MHLA_ERR_CUDA;
  g_last_launches = 1;
  return MHLA_OK;
}

int mhla_bwd_prep(const mhla_bwd_prep_desc* d, void* stream_) {
  if (!d || !d->dout || !d->out || !d->den || !d->dnum || !d->dden) return MHLA_ERR_INVALID_ARGUMENT;
  if (d->rows < 1 || (d->D != 64 && d->D != 128)) return MHLA_ERR_UNSUPPORTED_SHAPE;
  if (d->dtype != MHLA_BF16 && d->dtype != MHLA_FP16) return MHLA_ERR_INVALID_ARGUMENT;
  const uintptr_t al = reinterpret_cast<uintptr_t>(d->dout) | reinterpret_cast<uintptr_t>(d->out) |
                       reinterpret_cast<uintptr_t>(d->dnum);
  if ((al & 15) != 0) return MHLA_ERR_ALIGNMENT;
  DeviceState* dst = nullptr;
  int rc = device_state(&dst);
  if (rc != MHLA_OK) return rc;
  mhla::BwdPrepParams P{d->dout, d->out, d->den, d->dnum, d->dden, (long long)d->rows, d->D, d->dtype == MHLA_FP16};
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int tpr = d->D / 8, rpc = 256 / tpr;
  const long long want = (d->rows + rpc - 1) / rpc, cap = (long long)dst->sms * 8;
  const int grid = (int)(want < cap ? want : cap);
  if (tpr == 8) mhla::bwd_prep_kernel<8><<<grid, 256, 0, stream>>>(P);
  else mhla::bwd_prep_kernel<16><<<grid, 256, 0, stream>>>(P);
  if (!cuda_ok(cudaGetLastError(), "bwd_prep_kernel")) return MHLA_ERR_CUDA;
  g_last_launches = 1;
  return MHLA_OK;
}

int mhla_block_wsum(const mhla_block_wsum_desc* d, void* stream_) {
  if (!d || !d->x || !d->out) return MHLA_ERR_INVALID_ARGUMENT;
  if (d->blocks < 1 || d->w < 1 || (d->D != 64 && d->D != 128)) return MHLA_ERR_UNSUPPORTED_SHAPE;
  if (d->dtype != MHLA_BF16 && d->dtype != MHLA_FP16) return MHLA_ERR_INVALID_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(d->x) & 15) != 0) return MHLA_ERR_ALIGNMENT;
  DeviceState* dst = nullptr;
  int rc = device_state(&dst);
  if (rc != MHLA_OK) return rc;
  mhla::BlockSumParams P{d->x, d->wgt, d->out, (long long)d->blocks, d->w, d->D, d->dtype == MHLA_FP16};
  const long long cap = (long long)dst->sms * 8;
  const int grid = (int)(d->blocks < cap ? d->blocks : cap);
  mhla::block_wsum_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(P);
  if (!cuda_ok(cudaGetLastError(), "block_wsum_kernel")) return MHLA_ERR_CUDA;
  g_last_launches = 1;
  return MHLA_OK;
}

int mhla_bwd_post(const mhla_bwd_post_desc* d, void* stream_) {
  if (!d || !d->dnl || !d->ksum || !d->dksum || !d->dq || !d->dk) return MHLA_ERR_INVALID_ARGUMENT;
  if ((d->dqn == nullptr) != (d->dkn == nullptr)) return MHLA_ERR_INVALID_ARGUMENT;
  if (d->rows < 1 || d->w < 1 || d->rows % d->w != 0 || d->D < 8 || d->D % 8 != 0) return MHLA_ERR_UNSUPPORTED_SHAPE;
  if (d->dtype != MHLA_BF16 && d->dtype != MHLA_FP16) return MHLA_ERR_INVALID_ARGUMENT;
  const uintptr_t al = reinterpret_cast<uintptr_t>(d->dqn) | reinterpret_cast<uintptr_t>(d->dkn) |
                       reinterpret_cast<uintptr_t>(d->dq) | reinterpret_cast<uintptr_t>(d->dk) |
                       reinterpret_cast<uintptr_t>(d->ksum) | reinterpret_cast<uintptr_t>(d->dksum);
  if ((al & 15) != 0) return MHLA_ERR_ALIGNMENT;
  DeviceState* dst = nullptr;
  int rc = device_state(&dst);
  if (rc != MHLA_OK) return rc;
  mhla::BwdPostParams P{d->dqn, d->dkn, d->dnl, d->ksum, d->dksum, d->dq, d->dk, (long long)d->rows, d->w, d->D,
                        d->dtype == MHLA_FP16};
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const long long want = ((long long)d->rows * (d->D / 8) + 255) / 256, cap = (long long)dst->sms * 8;
  const int grid = (int)(want < cap ? want : cap);
  mhla::bwd_post_kernel<<<grid, 256, 0, stream>>>(P);
  if (!cuda_ok(cudaGetLastError(), "bwd_post_kernel")) return MHLA_ERR_CUDA;
  g_last_launches = 1;
  return MHLA_OK;
}

size_t mhla_causal_workspace_bytes(const mhla_causal_desc* desc) { return mhla::causal_workspace_bytes(desc); }

int mhla_fwd_causal(const mhla_causal_desc* d, void* stream_) {
  mhla::CausalPlan pl;
  int rc = mhla::plan_causal(d, &pl);
  if (rc != MHLA_OK) return rc;
  if (!d->q.ptr || !d->k.ptr || !d->v.ptr || !d->out.ptr || !d->mm) return MHLA_ERR_INVALID_ARGUMENT;
  if (!d->workspace || d->workspace_bytes < pl.total) return MHLA_ERR_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(d->workspace) & 1023) != 0) return MHLA_ERR_ALIGNMENT;
  const mhla_tensor4* ts[4] = {&d->q, &d->k, &d->v, &d->out};
  for (const mhla_tensor4* t : ts) {
    if ((reinterpret_cast<uintptr_t>(t->ptr) & 15) != 0) return MHLA_ERR_ALIGNMENT;
    if (t->stride_b % 8 || t->stride_t % 8 || t->stride_h % 8) return MHLA_ERR_ALIGNMENT;
  }
  if (d->mm_ld < pl.n) return MHLA_ERR_INVALID_ARGUMENT;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DeviceState* dst = nullptr;
  rc = device_state(&dst);
  if (rc != MHLA_OK) return rc;
  const int num_sms = dst->sms;

  mhla::CausalParams P;
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    bool hit = false;
    for (auto& e : g_causal_cache)
      if (std::memcmp(&e.key, d, sizeof(*d)) == 0) { P = e.params; hit = true; break; }
    if (!hit) {
      std::memset(&P, 0, sizeof P);
      const CUtensorMapDataType dt16 =
          d->dtype == MHLA_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
      uint8_t* ws = static_cast<uint8_t*>(d->workspace);
      void* S = ws + pl.off_S;
      void* St = ws + pl.off_St;
      void* Wp = ws + pl.off_W;
      const uint64_t Gn = (uint64_t)pl.G * pl.n, KV = (uint64_t)d->K * d->V;   // = Gs * ns: rows of consecutive groups
      auto t4 = [&](const mhla_tensor4& t, int dim) {
        // [B, T, H, dim] viewed as (dim, 64, n, H, B)
        MapSpec s{};
        s.dt = dt16; s.rank = 5; s.base = const_cast<void*>(t.ptr);
        const uint64_t dims[5] = {(uint64_t)dim, 64, (uint64_t)pl.n, (uint64_t)d->H, (uint64_t)d->B};
        const int64_t str[4] = {t.stride_t, t.stride_t * 64, t.stride_h, t.stride_b};
        uint64_t prev = (uint64_t)dim * 2;
        for (int i = 0; i < 5; ++i) s.dims[i] = dims[i];
        for (int i = 0; i < 4; ++i) {
          uint64_t bytes = (uint64_t)str[i] * 2;
          if (dims[i + 1] == 1 || bytes == 0) bytes = prev;
          s.strides[i] = bytes;
          prev = bytes * dims[i + 1];
        }
        s.box[0] = 64; s.box[1] = 64; s.box[2] = s.box[3] = s.box[4] = 1;
        return s;
      };
      if (!encode_map(&P.tmQ, t4(d->q, d->K))) return MHLA_ERR_CUDA;
      if (!encode_map(&P.tmK, t4(d->k, d->K))) return MHLA_ERR_CUDA;
      if (!encode_map(&P.tmV, t4(d->v, d->V))) return MHLA_ERR_CUDA;
      if (!encode_map(&P.tmO, t4(d->out, d->V))) return MHLA_ERR_CUDA;
      {
        MapSpec s{dt16, 3, S, {(uint64_t)d->V, (uint64_t)d->K, Gn}, {(uint64_t)d->V * 2, KV * 2},
                  {64, (uint32_t)d->K, 1}};
        if (!encode_map(&P.tmSst, s)) return MHLA_ERR_CUDA;
        s.base = St;
        if (!encode_map(&P.tmStld, s)) return MHLA_ERR_CUDA;
      }
      {
        MapSpec s{dt16, 3, S, {KV, (uint64_t)pl.ns, (uint64_t)pl.Gs}, {KV * 2, (uint64_t)pl.ns * KV * 2}, {64, 64, 1}};
        if (!encode_map(&P.tmSld, s)) return MHLA_ERR_CUDA;
        s.base = St; s.box[1] = 128;
        if (!encode_map(&P.tmStst, s)) return MHLA_ERR_CUDA;
      }
      {
        MapSpec s{dt16, 3, Wp, {(uint64_t)pl.Mp, (uint64_t)pl.ns, 2}, {(uint64_t)pl.Mp * 2, (uint64_t)pl.ns * pl.Mp * 2},
                  {64, 128, 1}};
        if (!encode_map(&P.tmW, s)) return MHLA_ERR_CUDA;
      }
      P.mm = d->mm; P.mm_ld = d->mm_ld;
      P.counters = reinterpret_cast<uint32_t*>(ws + pl.off_cnt);
      P.G = pl.Gs; P.H = d->H; P.n = pl.ns; P.pack = pl.pack; P.n0 = pl.n;
      P.n2_rows = pl.n2_rows; P.n2_cols = pl.n2_cols; P.kslabs = pl.kslabs;
      P.is_fp16 = d->dtype == MHLA_FP16; P.mode = 0; P.lag2 = 1; P.lag3 = 3;
      P.scale = d->scale;
      if (g_causal_cache.size() >= 32) g_causal_cache.erase(g_causal_cache.begin());
      CausalCacheEntry e;
      std::memcpy(&e.key, d, sizeof(*d));
      e.params = P;
      g_causal_cache.push_back(e);
    }
  }
  uint8_t* ws = static_cast<uint8_t*>(d->workspace);
  int launches = 0;
  mhla::prep_mix_kernel<<<8, 256, 0, stream>>>(d->mm, (long long)d->mm_ld, reinterpret_cast<uint16_t*>(ws + pl.off_W),
                                               pl.ns, pl.Mp, pl.n, 1, d->scale, d->dtype == MHLA_FP16, P.counters, 2 * pl.G);
  ++launches;
  // Measured on B200 (tools/causal_modes.py): with >= 512 chunks in flight the three phase-by-phase launches beat the
  // statically scheduled single kernel (97 vs 105 us at B=8,H=4,T=2048,K=128,V=256; 101 vs 146 us at H=16,K=V=64), below
  // that the single kernel wins (68 vs 80 us at B=2).  MHLA_FLAG_UNFUSED / MHLA_FLAG_FUSED force either.
  int unfused = (long long)pl.G * pl.n >= 512 ? 1 : 0;
  if (d->flags & MHLA_FLAG_UNFUSED) unfused = 1;
  if (d->flags & MHLA_FLAG_FUSED) unfused = 0;
  if (d->flags & MHLA_FLAG_STOP_AFTER_P1) unfused = 2;   // debugging / phase timing: summaries only
  if (d->flags & MHLA_FLAG_STOP_AFTER_P2) unfused = 3;   // ... summaries + mixing
  rc = MHLA_ERR_UNSUPPORTED_SHAPE;
#define MHLA_CAUSAL_CASE(KK, VV) \
  if (d->K == KK && d->V == VV) rc = mhla::causal_launch<KK, VV>(P, pl, unfused, num_sms, stream, &launches);
  MHLA_CAUSAL_CASE(64, 64)
  MHLA_CAUSAL_CASE(64, 128)
  MHLA_CAUSAL_CASE(128, 128)
  MHLA_CAUSAL_CASE(128, 256)
  MHLA_CAUSAL_CASE(64, 256)
  MHLA_CAUSAL_CASE(128, 64)
#undef MHLA_CAUSAL_CASE
  if (rc != MHLA_OK) return rc;
  if (!cuda_ok(cudaGetLastError(), "kernel launch")) return MHLA_ERR_CUDA;
  g_last_launches = launches;
  return MHLA_OK;
}

}  // extern "C"
