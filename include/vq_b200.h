/*
 * vq_b200.h — C ABI of libvq_b200.so: the B200-native (sm_100a) implementation of the visual-encoding
 * hot path of Sycamorers/vggt-qwen3 (frozen VGGT Aggregator -> trainable Perceiver projector -> tokens
 * spliced into Qwen3 inputs_embeds).
 *
 * The reference is pure Python/PyTorch and has NO foreign-function interface of its own; the entry
 * points below are what a binding for this path would bind.  Each one cites the reference code it
 * replaces (paths relative to the reference repo; [UPSTREAM] = the un-vendored `vggt` package that
 * src/models/vggt_qwen3_vlm.py:72 imports, restated in oracle/aggregator_ref.py).
 *
 * Conventions
 *  - Every function returns 0 on success or a negative VQ_ERR_* code; no exceptions cross the ABI.
 *    vq_last_error() returns a thread-local human readable message for the last failure.
 *  - All pointers are CUDA device pointers owned by the caller (PyTorch); the library never allocates
 *    or frees caller memory.  Rows are 16-byte aligned.  Work is enqueued on `stream` and the call
 *    returns without synchronising.  One process per GPU; functions are re-entrant; the only global
 *    state is a bounded (LRU, 2048 entries) cache of immutable TMA descriptors keyed by (pointer, shape, box).
 *  - "bf16" buffers hold IEEE bfloat16; "f32" buffers hold float.
 */
#ifndef VQ_B200_H_
#define VQ_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* vq_stream_t; /* cudaStream_t */

enum {
  VQ_OK = 0,
  VQ_ERR_INVALID = -1,   /* bad shape / alignment / argument */
  VQ_ERR_CUDA = -2,      /* a CUDA runtime or driver call failed */
  VQ_ERR_UNSUPPORTED = -3
};

enum { VQ_BF16 = 0, VQ_F32 = 1 };

/* Train-mode dropout (nn.Dropout / nn.MultiheadAttention(dropout=p) of src/models/projector_perceiver.py:33,37,42,46,49).
 * Masks are counter-based (Philox4x32-10, csrc/philox.cuh): element (row, col) of a site's [rows, ld] tensor is dropped
 * when its 16-bit lane of philox(seed, offset, site, row * ld/8 + col/8) is below round(p * 65536); kept elements are
 * scaled by 1/(1-p).  Forward and backward kernels regenerate the same mask from this descriptor; nothing is stored.
 * NULL or p == 0 means no dropout. */
typedef struct vq_dropout_desc {
  float p;
  uint32_t site;    /* distinguishes the dropout sites of one forward (layer * 4 + {0 attn-prob, 1 post-attn, 2 mlp, 3 post-mlp}) */
  uint64_t seed;    /* generator seed */
  uint64_t offset;  /* generator offset of this forward call */
} vq_dropout_desc;

/* GEMM epilogues */
enum {
  VQ_EPI_LINEAR = 0,   /* C = [add +] [gamma *] (alpha*acc [+ bias])                                  */
  VQ_EPI_GELU = 1,     /* C = gelu_erf(alpha*acc + bias); aux (optional, bf16) receives the pre-act   */
  VQ_EPI_GELU_BWD = 2, /* C = [add +] acc * gelu'(aux)   (aux = saved pre-activation, bf16)           */
  VQ_EPI_QKV_ROPE = 3, /* C = qkv + bias, per-head QK LayerNorm(64) + 2-D RoPE on q,k; q *= q_scale  */
  VQ_EPI_PATCH = 4     /* x[frame*P + patch_start + idx] = acc + bias + pos_embed[1 + idx]  (f32 out) */
};

const char* vq_last_error(void);
int vq_version(void);
/* Number of SMs of the current device (148 on B200); used to size persistent grids. */
int vq_num_sms(void);
/* Launch accounting: total kernels launched by this library so far (monotonic). */
long long vq_launch_count(void);
/* Optional CUDA-event profiling: enable(1) starts a fresh recording in which every launch is bracketed by two
 * events on its stream; dump() synchronises on them and writes one text line per (kind, tag, m, n, k):
 * "kind tag m n k launches total_ms total_flops total_bytes" (kind 0 = GEMM, 1 = attention, 2 = row kernel).
 * Returns bytes written, 0 if `cap` is too small. */
int vq_profile_enable(int on);
long long vq_profile_dump(char* buf, long long cap);

/* ------------------------------------------------------------------------------------------------
 * vq_gemm — C[M,N] = epilogue(A[M,K] * B[N,K]^T), bf16 inputs, fp32 accumulation in TMEM.
 * TMA-fed, warp-specialised, persistent tcgen05 kernel (csrc/gemm.cu).
 * Replaces every nn.Linear / F.linear / bmm call site on the path:
 *   [UPSTREAM] vggt/layers/attention.py qkv/proj, vggt/layers/mlp.py fc1/fc2, patch_embed.proj (as im2col GEMM);
 *   src/models/projector_perceiver.py:33 (nn.MultiheadAttention in/out projections and its two bmm),
 *   :34-39 (MLP), :60 (in_proj), :67 (out_proj), and their autograd dgrad/wgrad counterparts.
 *
 * Operand layout: X_major = 0 means the contraction dimension K is contiguous (A is [M,K] row-major,
 * B is [N,K] row-major = nn.Linear.weight); X_major = 1 means M (resp. N) is contiguous (A stored as
 * [K,M] row-major, B stored as [K,N] row-major).  ld* are in elements between consecutive outer rows.
 * Batching: Z0*Z1 independent problems; problem (z0,z1) shifts each 2-D tensor by
 * (outer += z0*x_r0 + z1*x_r1, inner += z0*x_c0 + z1*x_c1).  With Z0*Z1 > 1, K must be a multiple of 64.
 * ---------------------------------------------------------------------------------------------- */
typedef struct vq_gemm_desc {
  int32_t M, N, K;
  int32_t a_major, b_major;
  const void* A;
  int64_t lda;
  const void* B;
  int64_t ldb;
  void* C;
  int64_t ldc;
  int32_t c_dtype; /* VQ_BF16 or VQ_F32 */
  int32_t Z0, Z1;  /* 0 is treated as 1 */
  int64_t a_r0, a_c0, a_r1, a_c1;
  int64_t b_r0, b_c0, b_r1, b_c1;
  int64_t c_r0, c_c0, c_r1, c_c1;
  /* extents (outer rows, inner cols) of the 2-D tensors A and B live in; 0 = derive from M/N/K + batch offsets */
  int64_t a_rows, a_cols, b_rows, b_cols;

  int32_t epilogue; /* VQ_EPI_* */
  float alpha;      /* 0 is treated as 1 */
  const float* bias;  /* [N] or NULL */
  const float* gamma; /* [N] or NULL (LayerScale) */
  const void* add;    /* [M,N] in C's dtype, may alias C, or NULL; uses C's batch offsets */
  int64_t ldadd;
  void* aux; /* bf16 [M,N]; see VQ_EPI_GELU / VQ_EPI_GELU_BWD */
  int64_t ldaux;

  /* VQ_EPI_QKV_ROPE ([UPSTREAM] vggt/layers/attention.py q_norm/k_norm + vggt/layers/rope.py) and VQ_EPI_PATCH */
  int32_t tokens_per_frame; /* P = patch_start + gh*gw */
  int32_t patch_start;      /* 5 */
  int32_t grid_w;           /* gw: patches per image row */
  int32_t apply_qknorm_rope; /* 0 for the DINOv2 stem blocks (bias and q_scale only) */
  const float *qn_w, *qn_b, *kn_w, *kn_b; /* [64] each */
  float qk_eps;
  float q_scale;                     /* folded softmax scale: log2(e)/sqrt(64) */
  const float *rope_cos, *rope_sin;  /* [max_pos+1][16], position 0 = identity */
  int32_t rope_positions;            /* rows of the rope tables (max_pos + 1); 0 = unknown (tables read from global) */
  int32_t n_patches;                 /* VQ_EPI_PATCH: gh*gw */
  const float* pos_embed;            /* VQ_EPI_PATCH: [(1+n_patches), N] f32 */

  /* Dropout on the epilogue value BEFORE gamma / add (unbatched LINEAR, GELU: after the activation; GELU_BWD: the
   * incoming gradient is masked with the forward's mask); element index = (row m, column n) with ld = N. */
  vq_dropout_desc dropout;

  /* LayerNorm folded into the GEMMs around it (the aggregator's pre-norm blocks: no standalone LayerNorm launch).
   * Consumer side (ln_stats != NULL; LINEAR / GELU / QKV_ROPE, unbatched, A K-major): A holds the RAW rows x (bf16),
   * B the gamma-scaled weight W' = W diag(gamma_ln); with ln_stats[m] = (sum_k x, sum_k x^2) of row m (f32, over the
   * K = ln_dim features), ln_colsum[n] = sum_k W'[n,k] and bias = W beta_ln + b the epilogue computes
   *     LN(x) W^T + b = rstd_m (x W'^T - mean_m ln_colsum) + bias.
   * Producer side (LINEAR, f32 output): stats_out[m] += (sum_n C[m,n], sum_n C[m,n]^2) over the columns this launch
   * writes (atomics; the buffer must be zero when the launch starts), c_copy_bf16 receives a bf16 copy of C (same ldc),
   * and stats_zero (2*M floats) is cleared after the launch's stream dependency has resolved -- the buffer the NEXT
   * producer accumulates into, so two buffers ping-pong through a chain of blocks without memset launches. */
  const float* ln_stats;
  const float* ln_colsum;
  float ln_eps;
  float* stats_out;
  float* stats_zero;
  void* c_copy_bf16;
} vq_gemm_desc;

int vq_gemm(const vq_gemm_desc* d, vq_stream_t stream);
/* Tuning / A-B switch (process-wide; negative = keep): maximum number of K slices a tail tile of the persistent
 * schedule is split into (0 or 1 = never split = default; env VQ_GEMM_TAIL_SPLIT).  Measured not to pay on B200. */
int vq_gemm_tune(int32_t tail_split_max);

/* ------------------------------------------------------------------------------------------------
 * vq_attention_fwd — non-causal softmax(QK^T)V for head_dim 64 over a packed qkv buffer.
 * Flash-style tcgen05 kernel (csrc/attention.cu): S and O accumulate in TMEM, P is fed back to the
 * tensor core from TMEM, K/V tiles stream through a TMA pipeline.
 * Replaces [UPSTREAM] F.scaled_dot_product_attention in vggt/layers/attention.py for the DINOv2
 * stem, the frame-attention blocks (sequences = frames) and the global-attention blocks
 * (sequences = scenes, S*P tokens).
 *   qkv  : bf16 [n_seq*seq_len, 3*heads*64]  (q | k | v column blocks; q already scaled by log2(e)/8)
 *   out  : bf16 [n_seq*seq_len, heads*64]
 * ---------------------------------------------------------------------------------------------- */
int vq_attention_fwd(const void* qkv, void* out, int32_t n_seq, int32_t seq_len, int32_t heads,
                     vq_stream_t stream);
/* Kernel-variant overrides for tests and tuning (process-wide; negative = keep).
 *   q_tiles_per_cta : 0 = choose by sequence length (default: variant 4 up to 16384 keys, variant 3 beyond);
 *                     1 / 2 = split-KV kernel with 128- / 256-row CTAs (four 64-key streams, in-place P);
 *                     3 = wide kernel, 256-row CTAs, one per SM;  4 = wide kernel, 128-row CTAs, two per SM
 *   emu_pairs       : 0..6 of every 8 exponential pairs evaluated on the FMA pipe instead of the MUFU (default 4;
 *                     the split-KV kernels are instantiated for 0..4) */
int vq_attention_tune(int32_t q_tiles_per_cta, int32_t emu_pairs);
/* The same op for callers that can GUARANTEE |q . k| <= score_bound <= 64 for every (query, key) pair of the buffer (q
 * already carries the log2(e)/8 factor) -- the aggregator's frame / global blocks, whose q and k rows are per-head
 * LayerNorm-ed so that the bound follows from the frozen q_norm / k_norm weights.  P = 2^s needs no row maximum, no
 * range check and no rescale then, and every query row is served by two threads (csrc/attention.cu, "bounded"
 * variant).  score_bound <= 0 or > 64 falls back to vq_attention_fwd.  Scores outside the promised bound give
 * undefined probabilities (overflow), so the bound must be a proof, not an estimate. */
int vq_attention_fwd_bounded(const void* qkv, void* out, int32_t n_seq, int32_t seq_len, int32_t heads,
                             float score_bound, vq_stream_t stream);
/* enable: 0 routes vq_attention_fwd_bounded to the checked kernel (A/B runs); emu_pairs as in vq_attention_tune (default 2);
 * negative = keep */
int vq_attention_tune_bounded(int32_t enable, int32_t emu_pairs);

/* ------------------------------------------------------------------------------------------------
 * HBM-bound row kernels (csrc/rowops.cu)
 * ---------------------------------------------------------------------------------------------- */

/* K1+K2: ImageNet normalise ([UPSTREAM] aggregator.forward `(images - _resnet_mean) / _resnet_std`) fused with
 * the im2col of the 14x14/14 patch-embed conv ([UPSTREAM] PatchEmbed.proj): images f32 [frames,3,H,W] in [0,1]
 * -> bf16 [frames*gh*gw, Kp], column = c*patch^2 + dy*patch + dx, zero padded to Kp (multiple of 8). */
int vq_im2col_patches(const float* images, void* out_bf16, int32_t frames, int32_t H, int32_t W, int32_t patch,
                      int32_t Kp, vq_stream_t stream);
/* The same from uint8 views [frames, H, W, 3] (what the image decoder / the H2D copy delivers): ToTensor's x / 255
 * (src/dataio/collate_multiview.py:17) is applied inside, bit-identical to the f32 path, and the f32 image is never
 * materialised. */
int vq_im2col_patches_u8(const uint8_t* views_hwc, void* out_bf16, int32_t frames, int32_t H, int32_t W, int32_t patch,
                         int32_t Kp, vq_stream_t stream);

/* K3: DINOv2 token assembly ([UPSTREAM] DinoVisionTransformer.prepare_tokens_with_masks): row 0 of every frame =
 * cls_token + pos_embed[0]; rows 1..n_special-1 = register_tokens (no positional embedding). x is f32 [frames*P, D]. */
int vq_init_special_tokens(float* x, const float* cls_token, const float* pos_embed0, const float* register_tokens,
                           int32_t frames, int32_t P, int32_t D, int32_t n_special, vq_stream_t stream);

/* K4: LayerNorm over the last dim (nn.LayerNorm: [UPSTREAM] Block.norm1/norm2; src/models/projector_perceiver.py:40-41).
 * in/out dtype VQ_F32 or VQ_BF16; optional per-row mean / rstd outputs (saved for backward). D % 128 == 0. */
int vq_layernorm(const void* in, int32_t in_dtype, void* out, int32_t out_dtype, const float* gamma,
                 const float* beta, int64_t rows, int32_t D, float eps, float* mean_out, float* rstd_out,
                 vq_stream_t stream);

/* Residual add fused with the LayerNorm after it ([UPSTREAM] Block.forward: x = x + ls(branch(norm(x)))): x (f32 [rows, D],
 * in place) += y (bf16 [rows, D]: the LayerScale-d branch output written by the proj / fc2 GEMM), then
 * xn = LayerNorm(x) (bf16) for the GEMM that follows; xn == NULL: the add only (last block of a chain). */
int vq_residual_layernorm(float* x, const void* y_bf16, void* xn_bf16, const float* gamma, const float* beta, int64_t rows,
                          int32_t D, float eps, vq_stream_t stream);

/* Entry of a LayerNorm-folded block chain: stats[m] = (sum, sum of squares) of row m of x (f32 [rows, D]), xb = bf16(x),
 * stats_zero (optional, 2*rows floats) cleared: what the proj / fc2 epilogues maintain from then on. */
int vq_row_stats_cast(const float* x, void* xb_bf16, float* stats, float* stats_zero, int64_t rows, int32_t D,
                      vq_stream_t stream);

/* K13: DINOv2 final norm on the patch rows (x_norm_patchtokens) and swap of the special rows for the aggregator's
 * camera_token [1,2,1,D] / register_token [1,2,R,D] ([UPSTREAM] Aggregator.forward + slice_expand_and_flatten:
 * index 0 for the first frame of each scene, 1 for the rest).  In place on the f32 stream [rows = B*S*P, D]. */
int vq_stem_finalize(float* x, const float* gamma, const float* beta, const float* camera_token,
                     const float* register_token, int64_t rows, int32_t D, int32_t P, int32_t S, int32_t n_special,
                     float eps, vq_stream_t stream);

/* K14+K15: out[b, t, col_off:col_off+D] = bf16(x[b*tokens_per_scene + t, :]), t < n_take.  Writes one half of
 * `torch.cat([frame_out, global_out], -1)` restricted to the rows the caller consumes
 * (src/models/vggt_qwen3_vlm.py:148-156 keeps `[:, :num_vis_tokens]` of the flattened last entry). */
int vq_gather_tokens(const float* x, void* out_bf16, int32_t B, int64_t tokens_per_scene, int32_t n_take, int32_t D,
                     int32_t out_ld, int32_t col_off, vq_stream_t stream);

/* K22: training splice (src/models/vggt_qwen3_vlm.py:190-195): for every <image> position pos of sample b,
 * inputs_embeds[b, pos:pos+n_feat] = features[b]; later positions overwrite earlier ones.  No host sync.
 * *overflow_count (optional, device int) is incremented for every span that does not fit (the reference raises). */
int vq_splice_overwrite(void* inputs_embeds, int32_t dtype, const int64_t* input_ids, const void* features, int32_t B,
                        int32_t L, int32_t hidden, int32_t n_feat, int64_t image_id, int32_t* overflow_count,
                        vq_stream_t stream);

/* a11 + f1: geometry tokens and the fused training splice (csrc/geom.cu).
 * vq_geom_head_forward: encode_geom + geom_head (src/models/vggt_qwen3_vlm.py:51-56,164-177):
 *   pooled[b] = mean over views of cat(parts[0..3]) (R 9 | t 3 | K 9 | depth_hist 16; parts[k] f32 [B, V, widths[k]],
 *   NULL = zeros like the reference's `geom_token.get(key, zeros)`), pre = pooled W1^T + b1, act = silu(pre),
 *   y = act W2^T + b2 -> f32 [B, H].  pooled / pre / act are kept for the backward.  Weights f32 [H, in], [H, H].
 * vq_geom_head_backward: dW1, db1, dW2, db2 (overwritten, or += when accumulate) from dy [B, H]; scratch: B*H floats. */
int vq_geom_head_forward(const float* const* parts, const int32_t* widths, const float* w1, const float* b1,
                         const float* w2, const float* b2, float* pooled, float* pre, float* act, float* y, int32_t B,
                         int32_t V, int32_t in_dim, int32_t H, vq_stream_t stream);
int vq_geom_head_backward(const float* dy, const float* pooled, const float* pre, const float* act, const float* w2,
                          float* dw1, float* db1, float* dw2, float* db2, float* scratch, int32_t B, int32_t in_dim,
                          int32_t H, int32_t accumulate, vq_stream_t stream);
/* vq_splice_features: vq_splice_overwrite with the span cat([geom.expand(n_geom), vis])[b] never materialised
 * (src/models/vggt_qwen3_vlm.py:187-195): feature row t < n_geom is geom[b] (f32 [B, hidden], cast on the fly), row
 * t >= n_geom is vis[b, t - n_geom] (embedding dtype, [B, n_vis, hidden]).
 * vq_splice_features_backward: autograd backward of that assignment in one launch: d_embeds = grad with the written
 * rows zeroed (NULL to skip), d_geom f32 [B, hidden] / d_vis f32 [B, n_vis, hidden] = gradient rows gathered back to
 * the feature they were written from (either may be NULL). */
int vq_splice_features(void* inputs_embeds, int32_t dtype, const int64_t* input_ids, const float* geom, int32_t n_geom,
                       const void* vis, int32_t n_vis, int32_t B, int32_t L, int32_t hidden, int64_t image_id,
                       int32_t* overflow_count, vq_stream_t stream);
int vq_splice_features_backward(const void* grad, int32_t dtype, const int64_t* input_ids, void* d_embeds, float* d_geom,
                                float* d_vis, int32_t n_geom, int32_t n_vis, int32_t B, int32_t L, int32_t hidden,
                                int64_t image_id, vq_stream_t stream);

/* f1: inference splice (src/inference/qa_inference.py:119-145, src/inference/arkit_inference.py:182-214).
 * vq_find_first_token: *first (device int64, pre-set to n by the caller) = index of the first `token` in the
 * flattened [B*L] ids, i.e. `(input_ids == id).nonzero()[0]`; left at n when absent.
 * vq_splice_insert: with pos = *first % L, out_embeds[B, L-1+n_vis, hidden] = cat(emb[:, :pos], vis, emb[:, pos+1:])
 * and out_mask likewise with ones over the inserted span (mask elements of 8, 4 or 1 bytes).  Writes nothing when
 * *first == B*L (the reference returns its inputs unchanged).  No host synchronisation. */
int vq_find_first_token(const int64_t* input_ids, int64_t n, int64_t token, int64_t* first, vq_stream_t stream);
int vq_splice_insert(const void* inputs_embeds, int32_t dtype, const void* attention_mask, int32_t mask_elt_bytes,
                     const void* vis_tokens, const int64_t* first_pos, void* out_embeds, void* out_mask, int32_t B,
                     int32_t L, int32_t hidden, int32_t n_vis, vq_stream_t stream);

/* f2: image preprocessing from uint8 pixels (src/dataio/collate_multiview.py:12-19: Resize(image_size, BICUBIC) on a
 * PIL image + CenterCrop(image_size) + ToTensor), bit-exact with Pillow's antialiased bicubic (Resample.c: 22-bit
 * fixed-point taps, uint8 rounding after each pass) and torchvision's integer geometry; csrc/preprocess.cu.
 *   views_hwc : uint8 [frames, H, W, 3] (device), equally sized RGB views
 *   out_chw   : f32   [frames, 3, image_size, image_size] in [0, 1] (device) — the layout vq_aggregator_forward takes
 *   workspace : vq_preprocess_workspace_bytes(...) bytes (device, 16-byte aligned): taps + the horizontal-pass image */
size_t vq_preprocess_workspace_bytes(int32_t frames, int32_t H, int32_t W, int32_t image_size);
/* Host-only (no GPU): out6 = {resized H, resized W, crop top, crop left, horizontal taps, vertical taps}; and the
 * {first input index, tap count, taps[ksize]} rows of output indices [first, first + count) — returns ksize (> 0). */
int vq_preprocess_geometry(int32_t H, int32_t W, int32_t image_size, int32_t* out6);
int vq_preprocess_taps(int32_t in_size, int32_t out_size, int32_t first, int32_t count, int32_t* table,
                       int64_t capacity);
int vq_preprocess_views(const uint8_t* views_hwc, float* out_chw, int32_t frames, int32_t H, int32_t W,
                        int32_t image_size, void* workspace, size_t workspace_bytes, vq_stream_t stream);

/* f3: fused AdamW step for the projector parameter group (src/train/train_sft.py:138-163: torch.optim.AdamW, defaults
 * betas (0.9, 0.999), eps 1e-8, decoupled weight decay).  One pass over HBM: reads p, m, v (f32) and the gradient
 * (f32 or bf16, multiplied by grad_scale), writes p, m, v and, when param_bf16 is non-null, the bf16 compute copy of
 * the updated weights that the forward GEMMs read.  `step` counts from 1 (bias correction).  n % 4 == 0. */
int vq_adamw_step(float* param, float* exp_avg, float* exp_avg_sq, const void* grad, int32_t grad_dtype,
                  void* param_bf16, int64_t n, int32_t step, float lr, float beta1, float beta2, float eps,
                  float weight_decay, float grad_scale, vq_stream_t stream);
/* f3, sharded (ZeRO-1 style; the reference's equivalent is DeepSpeed's reduce-scatter / all-gather,
 * configs/deepspeed_zero3.json:7-14): the reduce step of a reduce-scatter fused into the optimizer.  The gradient of
 * the n elements this rank owns is the SUM of n_srcs (1..16) sources of the same dtype -- this rank's shard plus the
 * peers' shards pulled over NVLink -- accumulated in fp32 and never written back; AdamW as above on the f32 master
 * shard; weight_out (optional, bf16 or f32) receives the updated compute copy the forward kernels read and the peers
 * all-gather.  n may be any positive count (scalar tail handled). */
int vq_adamw_reduce_step(float* param, float* exp_avg, float* exp_avg_sq, const void* const* grad_srcs, int32_t n_srcs,
                         int32_t grad_dtype, void* weight_out, int32_t weight_dtype, int64_t n, int32_t step, float lr,
                         float beta1, float beta2, float eps, float weight_decay, float grad_scale, vq_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * vq_aggregator_forward — the whole frozen VGGT Aggregator forward (csrc/aggregator.cu): replaces
 * `self.vision_model.aggregator(images)` at src/models/vggt_qwen3_vlm.py:144 ([UPSTREAM] vggt/models/aggregator.py).
 * All weight pointers are device pointers; the structs themselves live in host memory.
 * Matrix weights are bf16 in nn.Linear layout [out, in]; vectors are f32.
 * ---------------------------------------------------------------------------------------------- */
typedef struct vq_block_weights {
  const float *norm1_w, *norm1_b;
  const void* qkv_w; /* bf16 [3D, D] */
  const float* qkv_b;
  const float *q_norm_w, *q_norm_b, *k_norm_w, *k_norm_b; /* [64]; all NULL for the DINOv2 stem blocks */
  const void* proj_w; /* bf16 [D, D] */
  const float* proj_b;
  const float* ls1; /* LayerScale gamma [D] */
  const float *norm2_w, *norm2_b;
  const void* fc1_w; /* bf16 [hidden, D] */
  const float* fc1_b;
  const void* fc2_w; /* bf16 [D, hidden] */
  const float* fc2_b;
  const float* ls2;
  /* LayerNorm folding (0 = off: qkv_w / fc1_w are the plain weights and vq_layernorm runs before them).  When set,
   * qkv_w = bf16(W_qkv diag(norm1_w)), qkv_b = W_qkv norm1_b + b_qkv, qkv_colsum[n] = sum_k qkv_w[n,k] (of the ROUNDED
   * bf16 weight, f32), and likewise fc1_w / fc1_b / fc1_colsum with norm2. */
  int32_t ln_folded;
  const float *qkv_colsum, *fc1_colsum;
  /* upper bound of |q . k| * log2(e)/8 over all token pairs, proven from q_norm / k_norm (host side:
   * (8 max|w_q| + |b_q|_2)(8 max|w_k| + |b_k|_2) * log2(e)/8, rotation by RoPE preserves norms); 0 = unknown */
  float qk_score_bound;
} vq_block_weights;

typedef struct vq_aggregator_weights {
  const void* patch_w; /* bf16 [D, Kp]: conv weight flattened (c, dy, dx), zero padded to Kp */
  const float* patch_b;
  const float* cls_token;       /* [D] */
  const float* pos_embed;       /* [1 + gh*gw, D], already interpolated to the patch grid */
  const float* register_tokens; /* DINOv2 registers [R, D] */
  const vq_block_weights* stem_blocks;
  int32_t n_stem;
  const float *stem_norm_w, *stem_norm_b;
  const float* camera_token;   /* [2, 1, D] */
  const float* register_token; /* [2, R, D] */
  const vq_block_weights* frame_blocks;
  const vq_block_weights* global_blocks;
  int32_t depth;
  const float *rope_cos, *rope_sin; /* [rope_max_pos + 1, 16] */
  int32_t rope_max_pos;
} vq_aggregator_weights;

typedef struct vq_aggregator_cfg {
  int32_t B, S;        /* scenes, frames per scene */
  int32_t H, W, patch; /* 518, 518, 14 */
  int32_t D, heads, hidden; /* 1024, 16 (head_dim must be 64), 4096 */
  int32_t n_special;   /* patch_start_idx = 1 + registers = 5 */
  int32_t Kp;          /* padded im2col width (592) */
  float stem_eps, block_eps, qk_eps; /* 1e-6, 1e-5, 1e-5 */
  /* 1 (recommended): the residual adds are deferred into the LayerNorm pass that follows them (proj / fc2 write a bf16
   * branch output, one row kernel does x += y and the next LayerNorm); 0: f32 read-modify-write in the GEMM epilogues
   * + standalone LayerNorm launches.  Ignored when the weights are LayerNorm-folded. */
  int32_t deferred_residual;
  int32_t images_u8;   /* 0: images are f32 [B,S,3,H,W] in [0,1];  1: uint8 views [B,S,H,W,3] (x / 255 applied inside) */
} vq_aggregator_cfg;

size_t vq_aggregator_workspace_bytes(const vq_aggregator_cfg* cfg);

/* images: f32 [B,S,3,H,W] in [0,1], or uint8 [B,S,H,W,3] when cfg->images_u8.
 * out_tokens: bf16 [B, n_take, 2D] = cat(frame_out, global_out)[-1] flattened over views, first n_take tokens
 *             (n_take = num_vis_tokens for encode_images; S*P for the full last list entry).
 * out_all: NULL, or `depth` device pointers, each bf16 [B, S*P, 2D], receiving every iteration's concat
 *          (the full list the upstream call returns; the reference only reads [-1]). */
int vq_aggregator_forward(const vq_aggregator_cfg* cfg, const vq_aggregator_weights* w, const void* images,
                          void* out_tokens, int32_t n_take, void* const* out_all, void* workspace,
                          size_t workspace_bytes, vq_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Perceiver projector row kernels (csrc/perceiver_ops.cu)
 * ---------------------------------------------------------------------------------------------- */

/* softmax over the first T columns of each f32 score row -> bf16 probabilities [rows, Tp] (columns T..Tp-1 = 0).
 * The attention core of nn.MultiheadAttention (src/models/projector_perceiver.py:33,45). T, Tp <= 1024.
 * With dropout (attention-probability dropout of nn.MultiheadAttention in train mode) probs_dropped_bf16 receives
 * mask * P / (1-p) (mask index: row, col with ld = Tp) and probs_bf16 the undropped P the backward needs. */
int vq_softmax_fwd(const float* scores, void* probs_bf16, int64_t rows, int32_t T, int32_t ld_scores, int32_t Tp,
                   const vq_dropout_desc* dropout, void* probs_dropped_bf16, vq_stream_t stream);
/* dS = scale * P .* (dP - rowsum(dP .* P)) -> bf16 [rows, Tp]; with dropout, dP = mask * dprobs / (1-p) first. */
int vq_softmax_bwd(const void* probs_bf16, const float* dprobs, void* dscores_bf16, int64_t rows, int32_t T,
                   int32_t ld_dprobs, int32_t Tp, float scale, const vq_dropout_desc* dropout, vq_stream_t stream);
/* One-tile attention core of nn.MultiheadAttention(latents, context, context) (src/models/projector_perceiver.py:33,45)
 * and its autograd backward (csrc/perceiver_attn.cu): one CTA per (sample, head), scores / dP never leave TMEM.
 * Replaces S = q k^T | vq_softmax_fwd | P v (and dv | dP | vq_softmax_bwd | dq | dk backward) when num_latents is 64 or
 * 128 and T <= 128, T % 8 == 0, head_dim % 64 == 0 (`vq_perceiver_attention_supported` says so; VQ_ERR_INVALID
 * otherwise).  All bf16: q [B*Lq, E]; kv [B*T, 2E] = k | v; probs / probs_dropped [B*heads*Lq, Tp], Tp = T rounded up to
 * 64 (columns T..Tp-1 = 0); att, datt, dq [B*Lq, E]; dkv [B*T, 2E] = dk | dv.  The scale is 1/sqrt(E/heads).  With dropout
 * (site stream as in vq_softmax_fwd: row = (b*heads + h)*Lq + i, ld = Tp) probs_dropped receives mask * P / (1-p); the
 * backward regenerates it from probs.  `vq_perceiver_attention_tune(0)` (or VQ_PERCEIVER_FUSED_ATTN=0) makes
 * `_supported` answer 0, so that vq_perceiver_forward / _backward take the multi-launch formulation (A/B, debugging). */
int vq_perceiver_attention_supported(int32_t Lq, int32_t T, int32_t E, int32_t heads);
int vq_perceiver_attention_tune(int32_t fused);
int vq_perceiver_attention_fwd(const void* q, const void* kv, void* probs, void* probs_dropped, void* att, int32_t B,
                               int32_t heads, int32_t Lq, int32_t T, int32_t E, const vq_dropout_desc* dropout,
                               vq_stream_t stream);
int vq_perceiver_attention_bwd(const void* q, const void* kv, const void* probs, const void* datt, void* dq, void* dkv,
                               int32_t B, int32_t heads, int32_t Lq, int32_t T, int32_t E,
                               const vq_dropout_desc* dropout, vq_stream_t stream);
/* LayerNorm backward (autograd of nn.LayerNorm at src/models/projector_perceiver.py:47,49): dz (bf16) from dy, the
 * saved LN input z and its row statistics; dgamma/dbeta (f32 [D]) are ACCUMULATED (+=).
 * With dropout, dz_dropped_bf16 additionally receives mask * dz / (1-p): the gradient of the branch that went through
 * `self.dropout(...)` before the residual add (mask index: row, col with ld = D). */
int vq_layernorm_bwd(const void* dy_bf16, const void* z_bf16, const float* mean, const float* rstd,
                     const float* gamma, void* dz_bf16, float* dgamma, float* dbeta, int64_t rows, int32_t D,
                     const vq_dropout_desc* dropout, void* dz_dropped_bf16, vq_stream_t stream);
/* Test hook: keep[r * cols + c] = 1 if element (r, c) of a [rows, ld] site survives the mask, else 0. */
int vq_dropout_mask(uint8_t* keep, int64_t rows, int32_t cols, int32_t ld, const vq_dropout_desc* dropout,
                    vq_stream_t stream);
/* out[n] += sum_m x[m, n]  (bias gradients) */
int vq_colsum(const void* x, int32_t dtype, float* out, int64_t rows, int32_t N, int64_t ld, vq_stream_t stream);
/* f32 master -> bf16 compute copy */
int vq_cast_f32_bf16(const float* in, void* out_bf16, int64_t n, vq_stream_t stream);
/* out[b, :] = src[:]  — `self.latents.unsqueeze(0).expand(B, -1, -1)` (src/models/projector_perceiver.py:79) */
int vq_broadcast_rows(const void* src_bf16, void* out_bf16, int64_t n, int32_t B, vq_stream_t stream);
/* out[i] (+)= sum_b in[b, i] — gradient of the expanded latents */
int vq_reduce_batch(const void* in_bf16, float* out, int64_t n, int32_t B, int32_t accumulate, vq_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * vq_perceiver_forward / vq_perceiver_backward (csrc/perceiver.cu): PerceiverProjector.forward
 * (src/models/projector_perceiver.py:70-82, layer :44-50) and its autograd backward.
 * bf16 compute copies of the f32 master weights are passed in (DeepSpeed-bf16 policy of the reference,
 * configs/deepspeed_zero3.json:4-6); gradients are produced in f32 directly from the TMEM accumulators.
 * Train mode (cfg.dropout_p > 0): attention-probability dropout (:33), dropout after the attention output projection
 * (:46), inside the MLP after GELU (:37) and after the MLP (:49), with counter-based masks (vq_dropout_desc).
 * ---------------------------------------------------------------------------------------------- */
typedef struct vq_perceiver_cfg {
  int32_t B, T;       /* samples, context tokens per sample (T % 8 == 0, T <= 1024) */
  int32_t in_dim;     /* 2048 */
  int32_t E, heads;   /* latent_dim 4096, num_heads 8  (E/heads % 64 == 0) */
  int32_t Lq;         /* num_latents 128 */
  int32_t layers;     /* 6 */
  int32_t ffn;        /* 16384 */
  int32_t out_dim;    /* text hidden size, 2560 */
  float ln_eps;       /* 1e-5 */
  int32_t out_dtype;  /* VQ_BF16 or VQ_F32 */
  /* train mode: dropout probability of the four sites of every layer (0 = eval), generator seed / offset of this
   * forward call; the backward must be given the same three values */
  float dropout_p;
  uint64_t dropout_seed, dropout_offset;
} vq_perceiver_cfg;

typedef struct vq_perceiver_layer_weights {
  const void* in_proj_w; /* bf16 [3E, E]  (nn.MultiheadAttention.in_proj_weight: q | k | v row blocks) */
  const float* in_proj_b; /* [3E] */
  const void* out_proj_w; /* bf16 [E, E] */
  const float* out_proj_b;
  const float *norm1_w, *norm1_b;
  const void* fc1_w; /* bf16 [ffn, E]  (mlp.0) */
  const float* fc1_b;
  const void* fc2_w; /* bf16 [E, ffn]  (mlp.3) */
  const float* fc2_b;
  const float *norm2_w, *norm2_b;
} vq_perceiver_layer_weights;

typedef struct vq_perceiver_weights {
  const void* latents;   /* bf16 [Lq, E] */
  const void* in_proj_w; /* bf16 [E, in_dim] */
  const float* in_proj_b;
  const vq_perceiver_layer_weights* layers;
  const void* out_proj_w; /* bf16 [out_dim, E] */
  const float* out_proj_b;
} vq_perceiver_weights;

/* Gradient destinations, same shapes as the parameters.  Every buffer is OVERWRITTEN unless accumulate != 0.
 * Matrix gradients (the *_w of the Linear layers, produced by the wgrad GEMMs) are f32 or -- weight_grad_bf16 = 1 --
 * bf16 written straight from the fp32 TMEM accumulators (the wire format of the data-parallel reduce: no f32 gradient
 * buffer, no cast pass); vectors (biases, LayerNorm affine) and the latents are always f32. */
typedef struct vq_perceiver_layer_grads {
  void* in_proj_w;
  float* in_proj_b;
  void* out_proj_w;
  float *out_proj_b, *norm1_w, *norm1_b;
  void* fc1_w;
  float* fc1_b;
  void* fc2_w;
  float *fc2_b, *norm2_w, *norm2_b;
} vq_perceiver_layer_grads;

typedef struct vq_perceiver_grads {
  float* latents;
  void* in_proj_w;
  float* in_proj_b;
  const vq_perceiver_layer_grads* layers;
  void* out_proj_w;
  float* out_proj_b;
  int32_t accumulate;
  int32_t weight_grad_bf16; /* 0: matrix gradients f32, 1: bf16 */
} vq_perceiver_grads;

/* Called on the host right after the work producing the gradients of `layer` has been ENQUEUED
 * (layer = cfg.layers for out_proj, cfg.layers-1 .. 0 for the Perceiver layers, -1 for in_proj + latents),
 * so the caller can record an event and start that bucket's NCCL all-reduce on another stream. */
typedef void (*vq_layer_done_fn)(int32_t layer, void* user);

size_t vq_perceiver_saved_bytes(const vq_perceiver_cfg* cfg);     /* activations kept for backward */
size_t vq_perceiver_workspace_bytes(const vq_perceiver_cfg* cfg); /* scratch for either direction */

/* tokens: bf16 [B*T, in_dim]; out: [B*Lq, out_dim] in cfg.out_dtype; saved: vq_perceiver_saved_bytes (1024-aligned). */
int vq_perceiver_forward(const vq_perceiver_cfg* cfg, const vq_perceiver_weights* w, const void* tokens_bf16,
                         void* out, void* saved, void* workspace, vq_stream_t stream);
/* grad_out: bf16 [B*Lq, out_dim]; grad_tokens: f32 [B*T, in_dim] or NULL. */
int vq_perceiver_backward(const vq_perceiver_cfg* cfg, const vq_perceiver_weights* w, const void* tokens_bf16,
                          const void* grad_out_bf16, const void* saved, const vq_perceiver_grads* grads,
                          float* grad_tokens, void* workspace, vq_layer_done_fn on_layer_done, void* user,
                          vq_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VQ_B200_H_ */
