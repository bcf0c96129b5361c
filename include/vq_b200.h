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
 *    state is a cache of immutable TMA descriptors keyed by (pointer, shape, box).
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
  int32_t n_patches;                 /* VQ_EPI_PATCH: gh*gw */
  const float* pos_embed;            /* VQ_EPI_PATCH: [(1+n_patches), N] f32 */
} vq_gemm_desc;

int vq_gemm(const vq_gemm_desc* d, vq_stream_t stream);

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

/* ------------------------------------------------------------------------------------------------
 * HBM-bound row kernels (csrc/rowops.cu)
 * ---------------------------------------------------------------------------------------------- */

/* K1+K2: ImageNet normalise ([UPSTREAM] aggregator.forward `(images - _resnet_mean) / _resnet_std`) fused with
 * the im2col of the 14x14/14 patch-embed conv ([UPSTREAM] PatchEmbed.proj): images f32 [frames,3,H,W] in [0,1]
 * -> bf16 [frames*gh*gw, Kp], column = c*patch^2 + dy*patch + dx, zero padded to Kp (multiple of 8). */
int vq_im2col_patches(const float* images, void* out_bf16, int32_t frames, int32_t H, int32_t W, int32_t patch,
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
} vq_aggregator_cfg;

size_t vq_aggregator_workspace_bytes(const vq_aggregator_cfg* cfg);

/* images: f32 [B,S,3,H,W] in [0,1].
 * out_tokens: bf16 [B, n_take, 2D] = cat(frame_out, global_out)[-1] flattened over views, first n_take tokens
 *             (n_take = num_vis_tokens for encode_images; S*P for the full last list entry).
 * out_all: NULL, or `depth` device pointers, each bf16 [B, S*P, 2D], receiving every iteration's concat
 *          (the full list the upstream call returns; the reference only reads [-1]). */
int vq_aggregator_forward(const vq_aggregator_cfg* cfg, const vq_aggregator_weights* w, const float* images,
                          void* out_tokens, int32_t n_take, void* const* out_all, void* workspace,
                          size_t workspace_bytes, vq_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VQ_B200_H_ */
