// vq_aggregator_forward: host-side schedule of the frozen VGGT Aggregator forward.
// One call enqueues the whole forward (DINOv2 stem + depth x (frame block, global block)) on the
// caller's stream: 7 kernels per transformer block, no host synchronisation, no allocation.
//
// Data layout in HBM (workspace, carved below):
//   x     f32  [M, D]       residual token stream, M = B*S*P rows ordered (scene, frame, token)
//   xn    bf16 [M, D]       LayerNorm output = GEMM A operand
//   qkv   bf16 [M, 3D]      q | k | v column blocks, q/k normalised + rotated, q pre-scaled
//   att   bf16 [M, D]       attention output
//   h     bf16 [M, hidden]  GELU(fc1) (the im2col patch matrix aliases this buffer before block 0)
//   y     bf16 [M, D]       LayerScale-d branch output of proj / fc2 (deferred-residual schedule)
// Frame attention views the rows as B*S sequences of P tokens, global attention as B sequences of
// S*P tokens; the buffer is never permuted.
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Workspace {
  float* x;
  void *xn, *qkv, *att, *h, *y;
  float* stats[2];   // LayerNorm-folded chain: (sum, sum of squares) per row, two buffers that ping-pong
  size_t total;
};

Workspace carve(const vq_aggregator_cfg* c, void* base) {
  const int gh = c->H / c->patch, gw = c->W / c->patch;
  const size_t P = c->n_special + (size_t)gh * gw;
  const size_t M = (size_t)c->B * c->S * P;
  const size_t D = c->D;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 1024);
    return o;
  };
  const size_t o_x = take(M * D * 4);
  const size_t o_xn = take(M * D * 2);
  const size_t o_qkv = take(M * 3 * D * 2);
  const size_t o_att = take(M * D * 2);
  const size_t h_bytes = M * (size_t)c->hidden * 2;
  const size_t patch_bytes = (size_t)c->B * c->S * gh * gw * c->Kp * 2;
  const size_t o_h = take(h_bytes > patch_bytes ? h_bytes : patch_bytes);
  const size_t o_s0 = take(M * 2 * 4), o_s1 = take(M * 2 * 4);
  const size_t o_y = take(M * D * 2);
  Workspace w;
  char* b = reinterpret_cast<char*>(base);
  w.x = reinterpret_cast<float*>(b + o_x);
  w.xn = b + o_xn;
  w.qkv = b + o_qkv;
  w.att = b + o_att;
  w.h = b + o_h;
  w.y = b + o_y;
  w.stats[0] = reinterpret_cast<float*>(b + o_s0);
  w.stats[1] = reinterpret_cast<float*>(b + o_s1);
  w.total = off;
  return w;
}

int check_cfg(const vq_aggregator_cfg* c) {
  VQ_CHECK_ARG(c != nullptr, "cfg is null");
  VQ_CHECK_ARG(c->B > 0 && c->S > 0, "B and S must be positive");
  VQ_CHECK_ARG(c->patch > 0 && c->H >= c->patch && c->W >= c->patch, "image geometry");
  VQ_CHECK_ARG(c->D > 0 && c->D % 128 == 0, "D must be a multiple of 128");
  VQ_CHECK_ARG(c->heads * 64 == c->D, "head_dim must be 64 (D == heads * 64)");
  VQ_CHECK_ARG(c->hidden > 0 && c->hidden % 64 == 0, "hidden must be a multiple of 64");
  VQ_CHECK_ARG(c->n_special >= 1, "n_special >= 1");
  VQ_CHECK_ARG(c->Kp >= 3 * c->patch * c->patch && c->Kp % 8 == 0, "Kp");
  return VQ_OK;
}

// One pre-norm transformer block over x (in place).  n_seq * seq_len == M.
int run_block(const vq_aggregator_cfg* c, const vq_block_weights& bw, const vq_aggregator_weights* w,
              const Workspace& ws, long long M, int P, int gw, int n_seq, int seq_len, float eps,
              vq_stream_t stream) {
  const int D = c->D;
  int rc;
  // LayerNorm folded into the GEMMs (bw.ln_folded): ws.xn holds the RAW bf16 rows of x and ws.stats[0] their
  // (sum, sum of squares); the proj epilogue rebuilds both for the MLP half into stats[1], the fc2 epilogue for the
  // next block into stats[0].  5 launches per block instead of 7.
  const bool fold = bw.ln_folded != 0;
  // LN1
  if (!fold) {
    rc = vq_layernorm(ws.x, VQ_F32, ws.xn, VQ_BF16, bw.norm1_w, bw.norm1_b, M, D, eps, nullptr, nullptr, stream);
    if (rc) return rc;
  }
  // qkv (+ QK-norm + RoPE for aggregator blocks), q pre-scaled by log2(e)/sqrt(64)
  {
    vq_gemm_desc g;
    memset(&g, 0, sizeof(g));
    g.M = (int)M; g.N = 3 * D; g.K = D;
    g.A = ws.xn; g.lda = D;
    g.B = bw.qkv_w; g.ldb = D;
    g.C = ws.qkv; g.ldc = 3 * D; g.c_dtype = VQ_BF16;
    g.epilogue = VQ_EPI_QKV_ROPE;
    g.bias = bw.qkv_b;
    g.tokens_per_frame = P; g.patch_start = c->n_special; g.grid_w = gw;
    g.apply_qknorm_rope = bw.q_norm_w != nullptr;
    g.qn_w = bw.q_norm_w; g.qn_b = bw.q_norm_b; g.kn_w = bw.k_norm_w; g.kn_b = bw.k_norm_b;
    g.qk_eps = c->qk_eps;
    g.q_scale = 1.4426950408889634f / 8.0f;
    g.rope_cos = w->rope_cos; g.rope_sin = w->rope_sin; g.rope_positions = w->rope_max_pos + 1;
    if (fold) { g.ln_stats = ws.stats[0]; g.ln_colsum = bw.qkv_colsum; g.ln_eps = eps; }
    rc = vq_gemm(&g, stream);
    if (rc) return rc;
  }
  rc = vq_attention_fwd_bounded(ws.qkv, ws.att, n_seq, seq_len, c->heads, bw.q_norm_w ? bw.qk_score_bound : 0.f, stream);
  if (rc) return rc;
  // x += ls1 * (att Wp^T + b)
  {
    vq_gemm_desc g;
    memset(&g, 0, sizeof(g));
    g.M = (int)M; g.N = D; g.K = D;
    g.A = ws.att; g.lda = D;
    g.B = bw.proj_w; g.ldb = D;
    g.C = ws.x; g.ldc = D; g.c_dtype = VQ_F32;
    g.epilogue = VQ_EPI_LINEAR;
    g.bias = bw.proj_b; g.gamma = bw.ls1;
    g.add = ws.x; g.ldadd = D;
    if (fold) { g.stats_out = ws.stats[1]; g.stats_zero = ws.stats[0]; g.c_copy_bf16 = ws.xn; }
    rc = vq_gemm(&g, stream);
    if (rc) return rc;
  }
  // LN2
  if (!fold) {
    rc = vq_layernorm(ws.x, VQ_F32, ws.xn, VQ_BF16, bw.norm2_w, bw.norm2_b, M, D, eps, nullptr, nullptr, stream);
    if (rc) return rc;
  }
  // h = gelu(xn W1^T + b1)
  {
    vq_gemm_desc g;
    memset(&g, 0, sizeof(g));
    g.M = (int)M; g.N = c->hidden; g.K = D;
    g.A = ws.xn; g.lda = D;
    g.B = bw.fc1_w; g.ldb = D;
    g.C = ws.h; g.ldc = c->hidden; g.c_dtype = VQ_BF16;
    g.epilogue = VQ_EPI_GELU;
    g.bias = bw.fc1_b;
    if (fold) { g.ln_stats = ws.stats[1]; g.ln_colsum = bw.fc1_colsum; g.ln_eps = eps; }
    rc = vq_gemm(&g, stream);
    if (rc) return rc;
  }
  // x += ls2 * (h W2^T + b2)
  {
    vq_gemm_desc g;
    memset(&g, 0, sizeof(g));
    g.M = (int)M; g.N = D; g.K = c->hidden;
    g.A = ws.h; g.lda = c->hidden;
    g.B = bw.fc2_w; g.ldb = c->hidden;
    g.C = ws.x; g.ldc = D; g.c_dtype = VQ_F32;
    g.epilogue = VQ_EPI_LINEAR;
    g.bias = bw.fc2_b; g.gamma = bw.ls2;
    g.add = ws.x; g.ldadd = D;
    if (fold) { g.stats_out = ws.stats[0]; g.stats_zero = ws.stats[1]; g.c_copy_bf16 = ws.xn; }
    rc = vq_gemm(&g, stream);
    if (rc) return rc;
  }
  return VQ_OK;
}

// The same block with the residual add DEFERRED into the LayerNorm pass that follows it (default schedule):
//   [LN1 -> xn]  qkv  attention  proj -> y = ls1 (att Wp^T + b)      (plain bf16 epilogue: no f32 read-modify-write)
//   x += y ; xn = LN2(x)                                              (one streaming row kernel)
//   fc1 -> h     fc2 -> y = ls2 (h W2^T + b2)
//   x += y ; xn = LN1_next(x)   (the NEXT block's first LayerNorm; add only at the end of a chain)
// `first` = the block opens a chain (xn is not prepared yet); next_w / next_b = norm1 of the next block or nullptr.
int run_block_deferred(const vq_aggregator_cfg* c, const vq_block_weights& bw, const vq_aggregator_weights* w,
                       const Workspace& ws, long long M, int P, int gw, int n_seq, int seq_len, float eps, bool first,
                       const float* next_w, const float* next_b, vq_stream_t stream) {
  const int D = c->D;
  int rc;
  if (first) {
    rc = vq_layernorm(ws.x, VQ_F32, ws.xn, VQ_BF16, bw.norm1_w, bw.norm1_b, M, D, eps, nullptr, nullptr, stream);
    if (rc) return rc;
  }
  {
    vq_gemm_desc g;
    memset(&g, 0, sizeof(g));
    g.M = (int)M; g.N = 3 * D; g.K = D;
    g.A = ws.xn; g.lda = D;
    g.B = bw.qkv_w; g.ldb = D;
    g.C = ws.qkv; g.ldc = 3 * D; g.c_dtype = VQ_BF16;
    g.epilogue = VQ_EPI_QKV_ROPE;
    g.bias = bw.qkv_b;
    g.tokens_per_frame = P; g.patch_start = c->n_special; g.grid_w = gw;
    g.apply_qknorm_rope = bw.q_norm_w != nullptr;
    g.qn_w = bw.q_norm_w; g.qn_b = bw.q_norm_b; g.kn_w = bw.k_norm_w; g.kn_b = bw.k_norm_b;
    g.qk_eps = c->qk_eps;
    g.q_scale = 1.4426950408889634f / 8.0f;
    g.rope_cos = w->rope_cos; g.rope_sin = w->rope_sin; g.rope_positions = w->rope_max_pos + 1;
    rc = vq_gemm(&g, stream);
    if (rc) return rc;
  }
  rc = vq_attention_fwd_bounded(ws.qkv, ws.att, n_seq, seq_len, c->heads, bw.q_norm_w ? bw.qk_score_bound : 0.f, stream);
  if (rc) return rc;
  {
    vq_gemm_desc g;
    memset(&g, 0, sizeof(g));
    g.M = (int)M; g.N = D; g.K = D;
    g.A = ws.att; g.lda = D;
    g.B = bw.proj_w; g.ldb = D;
    g.C = ws.y; g.ldc = D; g.c_dtype = VQ_BF16;
    g.epilogue = VQ_EPI_LINEAR;
    g.bias = bw.proj_b; g.gamma = bw.ls1;
    rc = vq_gemm(&g, stream);
    if (rc) return rc;
  }
  rc = vq_residual_layernorm(ws.x, ws.y, ws.xn, bw.norm2_w, bw.norm2_b, M, D, eps, stream);
  if (rc) return rc;
  {
    vq_gemm_desc g;
    memset(&g, 0, sizeof(g));
    g.M = (int)M; g.N = c->hidden; g.K = D;
    g.A = ws.xn; g.lda = D;
    g.B = bw.fc1_w; g.ldb = D;
    g.C = ws.h; g.ldc = c->hidden; g.c_dtype = VQ_BF16;
    g.epilogue = VQ_EPI_GELU;
    g.bias = bw.fc1_b;
    rc = vq_gemm(&g, stream);
    if (rc) return rc;
  }
  {
    vq_gemm_desc g;
    memset(&g, 0, sizeof(g));
    g.M = (int)M; g.N = D; g.K = c->hidden;
    g.A = ws.h; g.lda = c->hidden;
    g.B = bw.fc2_w; g.ldb = c->hidden;
    g.C = ws.y; g.ldc = D; g.c_dtype = VQ_BF16;
    g.epilogue = VQ_EPI_LINEAR;
    g.bias = bw.fc2_b; g.gamma = bw.ls2;
    rc = vq_gemm(&g, stream);
    if (rc) return rc;
  }
  return vq_residual_layernorm(ws.x, ws.y, next_w ? ws.xn : nullptr, next_w, next_b, M, D, eps, stream);
}

}  // namespace

extern "C" size_t vq_aggregator_workspace_bytes(const vq_aggregator_cfg* cfg) {
  if (check_cfg(cfg) != VQ_OK) return 0;
  return carve(cfg, nullptr).total;
}

extern "C" int vq_aggregator_forward(const vq_aggregator_cfg* c, const vq_aggregator_weights* w,
                                     const void* images, void* out_tokens, int32_t n_take,
                                     void* const* out_all, void* workspace, size_t workspace_bytes,
                                     vq_stream_t stream) {
  int rc = check_cfg(c);
  if (rc) return rc;
  VQ_CHECK_ARG(w && images && workspace, "null pointer");
  VQ_CHECK_ARG(w->n_stem >= 0 && w->depth >= 1, "depth");
  VQ_CHECK_ARG(out_tokens != nullptr || out_all != nullptr, "no output requested");
  const int gh = c->H / c->patch, gw = c->W / c->patch;
  const int n_patches = gh * gw;
  const int P = c->n_special + n_patches;
  const int F = c->B * c->S;
  const long long M = (long long)F * P;
  const int D = c->D;
  VQ_CHECK_ARG(M < (1LL << 31) - 1024, "too many tokens");
  VQ_CHECK_ARG(w->rope_max_pos >= (gh > gw ? gh : gw), "rope table too short");
  VQ_CHECK_ARG(!out_tokens || (n_take > 0 && n_take <= (long long)c->S * P), "n_take out of range");
  VQ_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) % 1024) == 0, "workspace must be 1024-byte aligned");
  const Workspace ws = carve(c, workspace);
  VQ_CHECK_ARG(workspace_bytes >= ws.total, "workspace too small");

  // ---- DINOv2 stem: normalise + im2col -> patch-embed GEMM (+bias +pos_embed) -> special rows
  rc = c->images_u8 ? vq_im2col_patches_u8(reinterpret_cast<const uint8_t*>(images), ws.h, F, c->H, c->W, c->patch, c->Kp, stream)
                    : vq_im2col_patches(reinterpret_cast<const float*>(images), ws.h, F, c->H, c->W, c->patch, c->Kp, stream);
  if (rc) return rc;
  {
    vq_gemm_desc g;
    memset(&g, 0, sizeof(g));
    g.M = F * n_patches; g.N = D; g.K = c->Kp;
    g.A = ws.h; g.lda = c->Kp;
    g.B = w->patch_w; g.ldb = c->Kp;
    g.C = ws.x; g.ldc = D; g.c_dtype = VQ_F32;
    g.epilogue = VQ_EPI_PATCH;
    g.bias = w->patch_b;
    g.tokens_per_frame = P; g.patch_start = c->n_special; g.n_patches = n_patches;
    g.pos_embed = w->pos_embed;
    rc = vq_gemm(&g, stream);
    if (rc) return rc;
  }
  rc = vq_init_special_tokens(ws.x, w->cls_token, w->pos_embed, w->register_tokens, F, P, D, c->n_special, stream);
  if (rc) return rc;
  // every block of a chain must agree on the folding (the statistics are handed from block to block)
  bool fold_stem = w->n_stem > 0, fold_agg = true;
  for (int i = 0; i < w->n_stem; ++i) fold_stem = fold_stem && w->stem_blocks[i].ln_folded;
  for (int i = 0; i < w->depth; ++i) fold_agg = fold_agg && w->frame_blocks[i].ln_folded && w->global_blocks[i].ln_folded;
  for (int i = 0; i < w->n_stem; ++i)
    VQ_CHECK_ARG((w->stem_blocks[i].ln_folded != 0) == fold_stem, "stem blocks disagree on ln_folded");
  for (int i = 0; i < w->depth; ++i)
    VQ_CHECK_ARG((w->frame_blocks[i].ln_folded != 0) == fold_agg && (w->global_blocks[i].ln_folded != 0) == fold_agg,
                 "aggregator blocks disagree on ln_folded");
  if (fold_stem) {
    rc = vq_row_stats_cast(ws.x, ws.xn, ws.stats[0], ws.stats[1], M, D, stream);
    if (rc) return rc;
  }
  const bool deferred = c->deferred_residual != 0 && !fold_stem && !fold_agg;
  for (int i = 0; i < w->n_stem; ++i) {
    if (deferred) {
      const bool last = i == w->n_stem - 1;
      rc = run_block_deferred(c, w->stem_blocks[i], w, ws, M, P, gw, F, P, c->stem_eps, i == 0,
                              last ? nullptr : w->stem_blocks[i + 1].norm1_w, last ? nullptr : w->stem_blocks[i + 1].norm1_b,
                              stream);
    } else {
      rc = run_block(c, w->stem_blocks[i], w, ws, M, P, gw, F, P, c->stem_eps, stream);
    }
    if (rc) return rc;
  }
  rc = vq_stem_finalize(ws.x, w->stem_norm_w, w->stem_norm_b, w->camera_token, w->register_token, M, D, P, c->S,
                        c->n_special, c->stem_eps, stream);
  if (rc) return rc;

  if (fold_agg) {
    rc = vq_row_stats_cast(ws.x, ws.xn, ws.stats[0], ws.stats[1], M, D, stream);
    if (rc) return rc;
  }

  // ---- alternating attention
  const long long SP = (long long)c->S * P;
  for (int i = 0; i < w->depth; ++i) {
    const bool last = i == w->depth - 1;
    if (deferred)
      rc = run_block_deferred(c, w->frame_blocks[i], w, ws, M, P, gw, F, P, c->block_eps, i == 0,
                              w->global_blocks[i].norm1_w, w->global_blocks[i].norm1_b, stream);
    else
      rc = run_block(c, w->frame_blocks[i], w, ws, M, P, gw, F, P, c->block_eps, stream);
    if (rc) return rc;
    if (out_all) {
      rc = vq_gather_tokens(ws.x, out_all[i], c->B, SP, (int)SP, D, 2 * D, 0, stream);
      if (rc) return rc;
    }
    if (last && out_tokens) {
      rc = vq_gather_tokens(ws.x, out_tokens, c->B, SP, n_take, D, 2 * D, 0, stream);
      if (rc) return rc;
    }
    if (deferred)
      rc = run_block_deferred(c, w->global_blocks[i], w, ws, M, P, gw, c->B, (int)SP, c->block_eps, false,
                              last ? nullptr : w->frame_blocks[i + 1].norm1_w, last ? nullptr : w->frame_blocks[i + 1].norm1_b,
                              stream);
    else
      rc = run_block(c, w->global_blocks[i], w, ws, M, P, gw, c->B, (int)SP, c->block_eps, stream);
    if (rc) return rc;
    if (out_all) {
      rc = vq_gather_tokens(ws.x, out_all[i], c->B, SP, (int)SP, D, 2 * D, D, stream);
      if (rc) return rc;
    }
    if (last && out_tokens) {
      rc = vq_gather_tokens(ws.x, out_tokens, c->B, SP, n_take, D, 2 * D, D, stream);
      if (rc) return rc;
    }
  }
  return VQ_OK;
}
