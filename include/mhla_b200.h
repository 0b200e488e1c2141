/*
 * mhla_b200 -- C ABI of the B200-native MHLA forward operator (sm_100a).
 *
 * The reference (DAGroup-PKU/MHLA) is pure Python and has no FFI / plugin interface for this path
 * (SURVEY.md 8b): the operator is inline PyTorch in
 *   - mhla_dit/mhla/mhla.py:262-268                              (variant A, block-mixed + normaliser)
 *   - mhla_image_classification/models/modules/attention/mhla.py:275-282   (twin of A)
 *   - mhla_videogen/diffusion/model/wan/mhla_utils.py:328-341    (variant B, roped numerator)
 *   - mhla_nlp/fla/ops/mhla/naive.py:10-83, :88-142              (variant C, causal chunked / recurrent)
 * This header therefore DEFINES the boundary a binding would use; each entry point cites the reference
 * lines it replaces.  INTEGRATION.md shows the ctypes stub a maintainer adds on the reference side.
 *
 * Conventions: plain pointers and sizes only (no torch types); all pointers are DEVICE pointers unless
 * stated; the caller owns inputs, outputs and workspace; calls only enqueue work on `stream` (a
 * cudaStream_t passed as void*) -- they never allocate device memory and never synchronise the device.
 * Return value: 0 on success, a negative mhla_status otherwise (see mhla_strerror).  Thread-safe and
 * re-entrant across streams (the only process-wide state is a mutex-guarded tensor-map cache).
 */
#ifndef MHLA_B200_H_
#define MHLA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MHLA_B200_ABI_VERSION 4

typedef enum mhla_status {
  MHLA_OK = 0,
  MHLA_ERR_INVALID_ARGUMENT = -1, /* NULL pointer, bad dtype / flag */
  MHLA_ERR_UNSUPPORTED_SHAPE = -2, /* shape outside the kernel's envelope (see each entry point) */
  MHLA_ERR_ALIGNMENT = -3,        /* pointer not 16-byte aligned or stride not a multiple of 8 elements */
  MHLA_ERR_WORKSPACE = -4,        /* workspace NULL or smaller than mhla_*_workspace_bytes() */
  MHLA_ERR_CUDA = -5,             /* a CUDA runtime/driver call failed (see mhla_last_cuda_error) */
  MHLA_ERR_NO_DEVICE = -6         /* current device is not sm_100 */
} mhla_status;

typedef enum mhla_dtype { MHLA_BF16 = 0, MHLA_FP16 = 1 } mhla_dtype;

enum {
  MHLA_FLAG_NORMALIZE = 1 << 0, /* divide by the (quirky) block-mixed normaliser, mhla.py:265-268 */
  MHLA_FLAG_FUSED = 1 << 7,     /* explicit request for the default path: ONE persistent kernel, items scheduled at run
                                 * time, cross-CTA dependencies through per-group counters */
  MHLA_FLAG_UNFUSED = 1 << 8,   /* three back-to-back launches (summaries, mixing, readout) chained with PDL; used by the
                                 * tests as an independent cross-check of the fused kernel */
  MHLA_FLAG_TWO_LAUNCH = 1 << 13, /* summaries + mixing in one dynamically scheduled kernel, then the readout as a second
                                 * launch that starts on the counters while the first one drains */
  MHLA_FLAG_WS_PERSISTENT = 1 << 14, /* the workspace is owned by this library's caller exclusively for this descriptor
                                 * shape: it was zeroed once (mhla_blockmix_workspace_init) and has only been used by
                                 * calls carrying this flag since.  The fused kernel then needs no prologue launch (its
                                 * CTAs split the mixing matrix themselves and the last one to finish re-zeroes the
                                 * control block): ONE launch per call.  Ignored by the multi-launch paths. */
  MHLA_FLAG_NO_SMALLN = 1 << 15, /* do not take the short-sequence kernel (M*w <= 256 tokens per unit, D = 64: the whole
                                 * unit is processed on chip by one CTA, no workspace); used by the tests as a cross-check */
  MHLA_FLAG_STOP_AFTER_P1 = 1 << 9,  /* debugging (with UNFUSED): stop after the block summaries */
  MHLA_FLAG_STOP_AFTER_P2 = 1 << 10, /* debugging (with UNFUSED): stop after the block mixing */
  MHLA_FLAG_ONLY_P3 = 1 << 11,       /* debugging (with UNFUSED): run only the readout on a caller-filled workspace */
  MHLA_FLAG_ONLY_P2 = 1 << 12        /* debugging (with UNFUSED): run only the block mixing */
};

/*
 * A strided view of a [B, H, M, w, D] tensor (batch, head, block, token-in-block, channel).
 * Strides are in ELEMENTS; the channel stride is 1.  This covers the reference's block-major
 * "(b h) n w d" layout (mhla.py:232-236) as well as token-major [B, N, H, D] tensors whose blocks
 * are contiguous token ranges.
 */
typedef struct mhla_tensor5 {
  const void* ptr;
  int64_t stride_b, stride_h, stride_m, stride_w;
} mhla_tensor5;

/*
 * Non-causal block-mixed forward (variants A and B):
 *   S_j   = Kn_j^T V_j                              (mhla.py:262 / mhla_utils.py:331)
 *   S~_i  = sum_j mix[i, j] S_j                     (mhla.py:263 / mhla_utils.py:332; the 1x1 Conv2d)
 *   den_i[t] = sum_j mix[i, j] (q_{j,t} . sum_s k_{j,s}) + eps     (mhla.py:265-266, the reference's quirk)
 *   out_i = (Qn_i S~_i) / den_i                     (mhla.py:268 / mhla_utils.py:339-341)
 * where Qn/Kn are q_rope/k_rope when given (variant B) and q/k otherwise, while the normaliser always
 * uses the un-roped q/k (mhla_utils.py:334-338).
 * Envelope: D in {64, 128} (Dk == Dv); 1 <= w <= 256; M >= 1; bf16 or fp16 I/O, fp32 accumulation,
 * block mixing as a hi+lo 16-bit GEMM (~16 mantissa bits); mix is fp32 [M, M] row-major with leading dimension mix_ld (elements).
 */
typedef struct mhla_blockmix_desc {
  int32_t B, H, M, w, D;
  int32_t dtype;         /* mhla_dtype */
  uint32_t flags;        /* MHLA_FLAG_* */
  float eps;
  mhla_tensor5 q, k, v;  /* q,k: un-roped (used by the normaliser; also the numerator if *_rope are NULL) */
  mhla_tensor5 q_rope, k_rope; /* ptr == NULL when absent */
  mhla_tensor5 out;      /* written; same dtype as the inputs */
  const float* mix;      /* [M, M] fp32, y_i = sum_j mix[i*mix_ld + j] x_j */
  int64_t mix_ld;
  void* workspace;       /* >= mhla_blockmix_workspace_bytes(desc) bytes, 1024-byte aligned */
  size_t workspace_bytes;
  /* Optional fused epilogue: per-(token, head) RMS normalisation of the output row,
   * out = o * rsqrt(mean_d(o^2) + out_rms_eps) * out_rms_weight[d]  - the per-head `g_norm` of MHLA_Video_Uni
   * (mhla_videogen/diffusion/model/wan/mhla_utils.py:360-362, WanRMSNorm wan/model.py:181-196).  NULL: off. */
  const float* out_rms_weight; /* [D] fp32 device pointer or NULL */
  float out_rms_eps;
  /* 3-D block view (ABI v3; SURVEY.md 8f rank 1, mhla_videogen/diffusion/model/wan/mhla_utils.py:317-326 and :345-354).
   * grid = (F, H, W) token grid, layout = (fb, hb, wb) blocks per axis; all zero = block-major tensors as above.
   * When set, q, k, v, q_rope, k_rope and out are TOKEN-major [B, F*H*W, heads, D] tensors: stride_w is the TOKEN
   * stride, stride_h the head stride, stride_m is ignored and stride_b must equal F*H*W*stride_w.  Block
   * j = (fbi*hb + hbi)*wb + wbi covers tokens (fbi*p1 + a, hbi*p2 + b, wbi*p3 + c), in-block order (a, b, c), with
   * (p1, p2, p3) = (F/fb, H/hb, W/wb); M must equal fb*hb*wb and w = p1*p2*p3.  The kernel gathers every block with one
   * TMA box per sub-tile and scatters the output the same way, so the reference's five rearrange copies and their
   * inverse disappear.  Envelope: p2*p3 <= 128 and ceil(p1 / floor(128 / (p2*p3))) <= 2 sub-tiles. */
  int32_t grid[3];
  int32_t layout[3];
  /* Further fused post-ops of the readout epilogue (ABI v4), applied in fp32 before the single
   * rounding, after the normaliser and the optional per-head RMS normalisation:
   *     out = o * silu(out_gate) + out_add
   * - the SiLU output gate and the "+ lepe" term of the Wan classes (mhla_videogen/diffusion/model/wan/mhla_utils.py
   * :363-365, wan/model.py:995-1001) and the "+ lepe" of MHLA4DiT (mhla_dit/mhla/mhla.py:271-273).  Both are 16-bit
   * tensors of the I/O dtype shaped like `out`, with their own strides (block-major or, with the 3-D block view,
   * token-major: e.g. plain views of the gate projection g(x) and of the depthwise-conv output); ptr == NULL: off. */
  mhla_tensor5 out_gate;
  mhla_tensor5 out_add;
} mhla_blockmix_desc;

size_t mhla_blockmix_workspace_bytes(const mhla_blockmix_desc* desc);
/* 1 when mhla_fwd_blockmix(desc) uses the workspace; 0 for shapes that take the short-sequence kernel (M*w <= 256 tokens
 * per (b,h) unit, D = 64, M <= 64, no rope / fused output norm / gate): workspace may then be NULL.  Pointers in desc are
 * ignored except that q_rope / k_rope / out_rms_weight / out_gate must be NULL or non-NULL as in the later call. */
int mhla_blockmix_needs_workspace(const mhla_blockmix_desc* desc);
/* White-box view of the workspace for tests: out[0..7] = byte offsets of S, S~, den, padded mix, counters,
 * then ncols (floats per S row: D*D summaries followed by wpad n_loc entries), wpad, padded mix pitch. */
int mhla_blockmix_workspace_layout(const mhla_blockmix_desc* desc, size_t out[8]);
int mhla_fwd_blockmix(const mhla_blockmix_desc* desc, void* stream);
/* Zero the control block of a workspace (enqueued on `stream`) before its first use with MHLA_FLAG_WS_PERSISTENT. */
int mhla_blockmix_workspace_init(const mhla_blockmix_desc* desc, void* stream);

/*
 * Causal chunked forward (variant C), replaces naive_chunk_simple_mhla_fixed
 * (mhla_nlp/fla/ops/mhla/naive.py:10-83) and, for T <= chunk, naive_recurrent_mhla (:88-142):
 *   o = K^-1/2 ((Q K^T) * Mask) V,  Mask[t, s] = mm[t / chunk, s / chunk] * 1[s <= t]
 * q,k: [B, T, H, K], v,o: [B, T, H, V] with element strides (stride_b, stride_t, stride_h, 1);
 * T is zero-padded to a multiple of `chunk` (naive.py:46-51); mm is fp32 [L, L] row-major, L >= ceil(T/chunk).
 * Envelope: chunk == 64; K in {64, 128}; V in {64, 128, 256}.
 */
typedef struct mhla_tensor4 {
  const void* ptr;
  int64_t stride_b, stride_t, stride_h;
} mhla_tensor4;

typedef struct mhla_causal_desc {
  int32_t B, T, H, K, V;
  int32_t chunk;
  int32_t dtype;
  uint32_t flags;
  float scale;           /* K^-1/2 in the reference (naive.py:42) */
  mhla_tensor4 q, k, v, out;
  const float* mm;       /* [L, L] fp32 */
  int64_t mm_ld;
  int32_t L;
  void* workspace;
  size_t workspace_bytes;
} mhla_causal_desc;

size_t mhla_causal_workspace_bytes(const mhla_causal_desc* desc);

/*
 * Fused pre-processing of the Wan MHLA layer (mhla_videogen/diffusion/model/wan/mhla_utils.py:267-276, :127-156,
 * :303-316): per token row  y = relu(x * rsqrt(mean_C(x^2) + eps_norm) * w) + eps  (WanRMSNorm over the FULL channel
 * dim, wan/model.py:181-196), then the 3-axis RoPE as a rotation of the interleaved pairs (2i, 2i+1) of every head by
 * the angle table [N, D/2] (cos, sin of the reference's complex128 freqs gathered for the (F, H, W) token grid).
 * Writes the roped q, k - and the un-roped ones when q_plain / k_plain are given (normaliser operands) - token-major
 * [B*N, C] in the 16-bit dtype: the layout the 3-D block view of mhla_fwd_blockmix consumes.  One launch.
 */
typedef struct mhla_wan_prep_desc {
  int32_t rows, N, C, D;        /* rows = B*N token rows, N tokens per sample, C = heads*D channels, D head dim */
  int32_t in_dtype;             /* 0 bf16, 1 fp16, 2 fp32 (the q / k projection outputs) */
  int32_t out_dtype;            /* mhla_dtype */
  const void* xq; const void* xk;   /* [rows, C], row pitch ld_in elements */
  int64_t ld_in;
  void* q_rope; void* k_rope;   /* [rows, C] 16-bit, contiguous */
  void* q_plain; void* k_plain; /* optional un-roped outputs, NULL to skip */
  const float* wq; const float* wk; /* RMSNorm weights [C] fp32, NULL = no normalisation */
  const float* cos_table; const float* sin_table; /* [N, D/2] fp32, NULL = no RoPE */
  float eps_norm, eps;
} mhla_wan_prep_desc;

int mhla_wan_prep(const mhla_wan_prep_desc* desc, void* stream);
int mhla_fwd_causal(const mhla_causal_desc* desc, void* stream);

/*
 * Backward (trainers mhla_dit/train.py:298-310, mhla_videogen/train_wan.py:717).  The gradient
 * CONTRACTIONS of variants A / B are calls of mhla_fwd_blockmix itself with permuted operands,
 *   dQ = blockmix(q = dO~, k = V,   v = K,   mix)          dV = blockmix(q = K, k = Q, v = dO~, mix^T)
 *   dK = blockmix(q = V,   k = dO~, v = Q,   mix^T)        (no normaliser flag; dO~ = dO / den, or dO itself)
 * and d mix contracts the block summaries those launches leave in their workspaces (S of the dK launch against S of the
 * dQ launch).  Variant C likewise through mhla_fwd_causal on time-reversed tensors (mhla_b200/autograd.py).  With the
 * normaliser (mhla.py:265-268) two streaming passes over contiguous [rows, D] token rows remain:
 *   mhla_bwd_prep:  dnum[r,:] = dout[r,:] / den[r]   and   dden[r] = -(dout[r,:] . out[r,:]) / den[r]
 *   mhla_bwd_post:  dq[r,:] = dqn[r,:] + dnl[r] * ksum[r / w, :]   and   dk[r,:] = dkn[r,:] + dksum[r / w, :]
 * (den: the forward's normaliser; dnl = mix^T dden; ksum = sum_t k[.,t,:]; dksum = sum_t dnl[.,t] q[.,t,:]).
 * dqn / dkn may both be NULL (roped numerator: the un-roped q, k only feed the normaliser).  D in {64, 128} for prep.
 */
typedef struct mhla_bwd_prep_desc {
  int64_t rows; int32_t D; int32_t dtype;
  const void* dout; const void* out;  /* [rows, D] 16-bit */
  const float* den;                   /* [rows] */
  void* dnum;                         /* [rows, D] 16-bit, written */
  float* dden;                        /* [rows], written */
} mhla_bwd_prep_desc;
typedef struct mhla_bwd_post_desc {
  int64_t rows; int32_t w, D, dtype;
  const void* dqn; const void* dkn;   /* [rows, D] 16-bit or NULL */
  const float* dnl;                   /* [rows] */
  const float* ksum; const float* dksum; /* [rows / w, D] fp32 */
  void* dq; void* dk;                 /* [rows, D] 16-bit, written (may alias dqn / dkn) */
} mhla_bwd_post_desc;
/* out[b, :] = sum_t wgt[b, t] * x[b, t, :] over the w token rows of every block (wgt NULL: plain sums): ksum_j = sum_t
 * k_{j,t} (mhla.py:265) and its gradient partner dksum_j = sum_t dnl[j,t] q_{j,t}.  x: [blocks, w, D] 16-bit contiguous,
 * wgt: [blocks, w] fp32 or NULL, out: [blocks, D] fp32; D in {64, 128}. */
typedef struct mhla_block_wsum_desc {
  int64_t blocks; int32_t w, D, dtype;
  const void* x; const float* wgt; float* out;
} mhla_block_wsum_desc;
int mhla_block_wsum(const mhla_block_wsum_desc* desc, void* stream);
int mhla_bwd_prep(const mhla_bwd_prep_desc* desc, void* stream);
int mhla_bwd_post(const mhla_bwd_post_desc* desc, void* stream);

/*
 * Post-op of the NLP layer: FusedRMSNormGated(o, g) (mhla_nlp/fla/modules/fused_norm_gate.py:77-99,
 * called at mhla_nlp/fla/layers/mhla.py:350-356) per (token, head) row of the causal operator's output, one pass:
 *   out[r, :] = x[r, :] * rsqrt(mean(x[r, :]^2) + eps) * weight * g[r, :] * sigmoid(g[r, :])
 * x, g: [rows, D] 16-bit with row pitches ld_x / ld_g (elements); out: [rows, D] contiguous; weight fp32 [D] or NULL;
 * g == NULL: plain RMSNorm (layers/mhla.py:358-360).  D in {64, 128, 256}.
 */
typedef struct mhla_gated_rmsnorm_desc {
  int64_t rows; int32_t D; int32_t dtype;
  const void* x; int64_t ld_x;
  const void* g; int64_t ld_g;
  const float* weight;
  float eps;
  void* out;
} mhla_gated_rmsnorm_desc;
int mhla_gated_rmsnorm(const mhla_gated_rmsnorm_desc* desc, void* stream);

/*
 * Post-ops of the Wan / DiT layers behind the block-mixed operator (SiLU gate and "+ lepe",
 * mhla_videogen/diffusion/model/wan/mhla_utils.py:360-366, wan/model.py:1001-1003, mhla_dit/mhla/mhla.py:268-273), one
 * streaming pass:   out[r, :] = x[r, :] * silu(g[r, :]) + add[r, :]     (g == NULL: no gate; add == NULL: nothing added)
 * x, g, add, out: [rows, C] 16-bit, contiguous along C (C % 8 == 0), row pitches ld_* in elements; out may alias x.
 */
typedef struct mhla_gate_add_desc {
  int64_t rows; int32_t C; int32_t dtype;
  const void* x; int64_t ld_x;
  const void* g; int64_t ld_g;
  const void* add; int64_t ld_add;
  void* out; int64_t ld_out;
} mhla_gate_add_desc;
int mhla_gate_add(const mhla_gate_add_desc* desc, void* stream);

/*
 * LePE of the Wan layers: the depthwise 3x3x3 convolution `self.lepe = nn.Conv3d(dim, dim, 3, padding=1, groups=dim)`
 * applied to v (mhla_videogen/diffusion/model/wan/mhla_utils.py:289-296, wan/model.py `lepe`), computed on the token-major
 * [B, F*H*W, C] tensor without the NCDHW rearrangements:
 *   out[b, (f,h,w), c] = bias[c] + sum_{kf,kh,kw in 0..2} wt[(kf*3+kh)*3+kw][c] * x[b, (f+kf-1, h+kh-1, w+kw-1), c]   (zero padding)
 * x: 16-bit, row pitch ld_x elements; wt: fp32 [27][C] (the Conv3d weight [C,1,3,3,3] transposed); bias fp32 [C] or NULL;
 * out: [B, F*H*W, C] contiguous, dtype of x.  C % 8 == 0.
 */
typedef struct mhla_dwconv3d_desc {
  int32_t B, F, H, W, C, dtype;
  const void* x; int64_t ld_x;
  const float* wt;
  const float* bias;
  void* out;
} mhla_dwconv3d_desc;
int mhla_dwconv3d(const mhla_dwconv3d_desc* desc, void* stream);

/* Misc. */
int mhla_abi_version(void);
const char* mhla_strerror(int status);
const char* mhla_last_cuda_error(void);   /* thread-local text of the last CUDA failure, "" if none */
/* Number of kernel launches (memsets excluded) the last successful mhla_fwd_* call on this thread made. */
int mhla_last_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MHLA_B200_H_ */
