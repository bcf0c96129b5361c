// vq_perceiver_forward / vq_perceiver_backward: host-side schedule of the Perceiver projector
// (src/models/projector_perceiver.py:44-50, 70-82) and of its autograd backward, expressed as
// tcgen05 GEMMs (forward, dgrad with MN-major weights, wgrad with both operands MN-major, batched
// per-(sample, head) attention products) plus the row kernels of perceiver_ops.cu.
//
// Row spaces:  R = B*Lq latent rows,  C = B*T context rows,  A = B*heads*Lq attention rows (Tp = T padded to 64).
// Saved for backward (bf16 unless noted), per layer l:
//   lat[l] [R,E] layer input | q [R,E] | kv [C,2E] | P [A,Tp] | att [R,E] | z1 [R,E] + mean1,rstd1 f32 [R]
//   x1 [R,E] | pre [R,F] | g [R,F] | z2 [R,E] + mean2,rstd2 f32 [R];  plus ctx [C,E] and lat[layers] [R,E].
// Train mode (cfg.dropout_p > 0) additionally keeps Pd [A,Tp] = dropout(P); `g` then holds dropout(gelu(pre)).
// The four dropout sites of layer l use mask streams 4l + {0: attention probabilities, 1: after the attention output
// projection, 2: after GELU, 3: after the MLP}; the backward regenerates every mask from (seed, offset, site).
#include <math.h>
#include <string.h>

#include <cuda_bf16.h>

#include "common.cuh"

namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Dims {
  long long R, C, A;
  int E, F, T, Tp, Lq, H, hd, B;
};

Dims dims_of(const vq_perceiver_cfg* c) {
  Dims d;
  d.B = c->B; d.T = c->T; d.Tp = (c->T + 63) / 64 * 64;
  d.E = c->E; d.F = c->ffn; d.Lq = c->Lq; d.H = c->heads; d.hd = c->E / c->heads;
  d.R = (long long)c->B * c->Lq;
  d.C = (long long)c->B * c->T;
  d.A = (long long)c->B * c->heads * c->Lq;
  return d;
}

struct LayerSaved {
  char *q, *kv, *P, *Pd, *att, *z1, *x1, *pre, *g, *z2;
  float *mean1, *rstd1, *mean2, *rstd2;
};
struct Saved {
  char* ctx;
  char* lat[65];  // lat[l] = input of layer l; lat[layers] = final latents
  LayerSaved layer[64];
  size_t total;
};

Saved carve_saved(const vq_perceiver_cfg* c, void* base) {
  const Dims d = dims_of(c);
  Saved s;
  size_t off = 0;
  char* b = reinterpret_cast<char*>(base);
  auto take = [&](size_t bytes) {
    char* p = b + off;
    off = align_up(off + bytes, 1024);
    return p;
  };
  const size_t RE = (size_t)d.R * d.E * 2, RF = (size_t)d.R * d.F * 2;
  s.ctx = take((size_t)d.C * d.E * 2);
  for (int l = 0; l <= c->layers; ++l) s.lat[l] = take(RE);
  for (int l = 0; l < c->layers; ++l) {
    LayerSaved& L = s.layer[l];
    L.q = take(RE);
    L.kv = take((size_t)d.C * 2 * d.E * 2);
    L.P = take((size_t)d.A * d.Tp * 2);
    L.Pd = c->dropout_p > 0.f ? take((size_t)d.A * d.Tp * 2) : L.P;
    L.att = take(RE);
    L.z1 = take(RE);
    L.x1 = take(RE);
    L.pre = take(RF);
    L.g = take(RF);
    L.z2 = take(RE);
    L.mean1 = reinterpret_cast<float*>(take((size_t)d.R * 4));
    L.rstd1 = reinterpret_cast<float*>(take((size_t)d.R * 4));
    L.mean2 = reinterpret_cast<float*>(take((size_t)d.R * 4));
    L.rstd2 = reinterpret_cast<float*>(take((size_t)d.R * 4));
  }
  s.total = off;
  return s;
}

struct Work {
  float* S;       // f32 [A, Tp]  scores / dP
  char *dlat, *dz2, *dpre, *dx1, *dz1, *datt, *dS, *dq, *dkv, *dctx_bf16, *tmp_lat, *ddrop;
  float* dctx;    // f32 [C, E]
  size_t total;
};

Work carve_work(const vq_perceiver_cfg* c, void* base) {
  const Dims d = dims_of(c);
  Work w;
  size_t off = 0;
  char* b = reinterpret_cast<char*>(base);
  auto take = [&](size_t bytes) {
    char* p = b + off;
    off = align_up(off + bytes, 1024);
    return p;
  };
  const size_t RE = (size_t)d.R * d.E * 2;
  w.S = reinterpret_cast<float*>(take((size_t)d.A * d.Tp * 4));
  w.dlat = take(RE);
  w.dz2 = take(RE);
  w.dpre = take((size_t)d.R * d.F * 2);
  w.dx1 = take(RE);
  w.dz1 = take(RE);
  w.datt = take(RE);
  w.dS = take((size_t)d.A * d.Tp * 2);
  w.dq = take(RE);
  w.dkv = take((size_t)d.C * 2 * d.E * 2);
  w.dctx = reinterpret_cast<float*>(take((size_t)d.C * d.E * 4));
  w.dctx_bf16 = take((size_t)d.C * d.E * 2);
  w.tmp_lat = take(RE);
  w.ddrop = take(RE);
  w.total = off;
  return w;
}

int check_cfg(const vq_perceiver_cfg* c) {
  VQ_CHECK_ARG(c != nullptr, "cfg is null");
  VQ_CHECK_ARG(c->B > 0 && c->T > 0 && c->T % 8 == 0 && c->T <= 1024, "T must be a multiple of 8, <= 1024");
  VQ_CHECK_ARG(c->in_dim > 0 && c->in_dim % 8 == 0, "in_dim % 8");
  VQ_CHECK_ARG(c->E > 0 && c->heads > 0 && c->E % c->heads == 0 && (c->E / c->heads) % 64 == 0,
               "head_dim must be a multiple of 64");
  VQ_CHECK_ARG(c->E % 128 == 0, "latent_dim % 128");
  VQ_CHECK_ARG(c->Lq > 0 && c->Lq % 64 == 0, "num_latents must be a multiple of 64");
  VQ_CHECK_ARG(c->layers > 0 && c->layers <= 64, "1..64 layers");
  VQ_CHECK_ARG(c->ffn > 0 && c->ffn % 8 == 0 && c->out_dim > 0 && c->out_dim % 8 == 0, "ffn/out_dim % 8");
  VQ_CHECK_ARG(c->out_dtype == VQ_BF16 || c->out_dtype == VQ_F32, "out_dtype");
  VQ_CHECK_ARG(c->dropout_p >= 0.f && c->dropout_p < 1.f, "dropout_p must be in [0, 1)");
  return VQ_OK;
}

// Plain (unbatched) GEMM helper.  a_mn / b_mn select MN-major operands.
int gemm(vq_stream_t st, int M, int N, int K, const void* A, long long lda, int a_mn, const void* Bm, long long ldb,
         int b_mn, void* Cm, long long ldc, int c_dtype, const float* bias = nullptr, const void* add = nullptr,
         long long ldadd = 0, int epilogue = VQ_EPI_LINEAR, void* aux = nullptr, long long ldaux = 0,
         const vq_dropout_desc* dropout = nullptr) {
  vq_gemm_desc g;
  memset(&g, 0, sizeof(g));
  g.M = M; g.N = N; g.K = K;
  g.A = A; g.lda = lda; g.a_major = a_mn;
  g.B = Bm; g.ldb = ldb; g.b_major = b_mn;
  g.C = Cm; g.ldc = ldc; g.c_dtype = c_dtype;
  g.epilogue = epilogue;
  g.bias = bias;
  g.add = add; g.ldadd = ldadd;
  g.aux = aux; g.ldaux = ldaux;
  if (dropout) g.dropout = *dropout;
  return vq_gemm(&g, st);
}

// mask stream of dropout site `site` (0..3) of layer l; nullptr-equivalent (p = 0) in eval mode
vq_dropout_desc site_of(const vq_perceiver_cfg* c, int l, int site) {
  vq_dropout_desc d;
  d.p = c->dropout_p;
  d.site = (uint32_t)(4 * l + site);
  d.seed = c->dropout_seed;
  d.offset = c->dropout_offset;
  return d;
}

#define RUN(expr)        \
  do {                   \
    int _rc = (expr);    \
    if (_rc) return _rc; \
  } while (0)

}  // namespace

extern "C" size_t vq_perceiver_saved_bytes(const vq_perceiver_cfg* c) {
  if (check_cfg(c) != VQ_OK) return 0;
  return carve_saved(c, nullptr).total;
}
extern "C" size_t vq_perceiver_workspace_bytes(const vq_perceiver_cfg* c) {
  if (check_cfg(c) != VQ_OK) return 0;
  return carve_work(c, nullptr).total;
}

extern "C" int vq_perceiver_forward(const vq_perceiver_cfg* c, const vq_perceiver_weights* w, const void* tokens,
                                    void* out, void* saved, void* workspace, vq_stream_t st) {
  RUN(check_cfg(c));
  VQ_CHECK_ARG(w && tokens && out && saved && workspace, "null pointer");
  VQ_CHECK_ARG((reinterpret_cast<uintptr_t>(saved) % 1024) == 0 && (reinterpret_cast<uintptr_t>(workspace) % 1024) == 0,
               "saved/workspace must be 1024-byte aligned");
  const Dims d = dims_of(c);
  const Saved s = carve_saved(c, saved);
  const Work k = carve_work(c, workspace);
  const int E = d.E, F = d.F, R = (int)d.R, C = (int)d.C;
  const float scale = 1.0f / sqrtf((float)d.hd);
  const bool fused_attn = vq_perceiver_attention_supported(d.Lq, d.T, E, d.H) != 0;

  // context = in_proj(tokens); latents expanded over the batch
  RUN(gemm(st, C, E, c->in_dim, tokens, c->in_dim, 0, w->in_proj_w, c->in_dim, 0, s.ctx, E, VQ_BF16, w->in_proj_b));
  RUN(vq_broadcast_rows(w->latents, s.lat[0], (long long)d.Lq * E, d.B, st));

  for (int l = 0; l < c->layers; ++l) {
    const vq_perceiver_layer_weights& lw = w->layers[l];
    const LayerSaved& L = s.layer[l];
    const __nv_bfloat16* Wqkv = reinterpret_cast<const __nv_bfloat16*>(lw.in_proj_w);
    // q = lat Wq^T + bq ; kv = ctx [Wk;Wv]^T + [bk;bv]
    RUN(gemm(st, R, E, E, s.lat[l], E, 0, Wqkv, E, 0, L.q, E, VQ_BF16, lw.in_proj_b));
    RUN(gemm(st, C, 2 * E, E, s.ctx, E, 0, Wqkv + (size_t)E * E, E, 0, L.kv, 2 * E, VQ_BF16, lw.in_proj_b + E));
    const vq_dropout_desc dr0 = site_of(c, l, 0), dr1 = site_of(c, l, 1), dr2 = site_of(c, l, 2), dr3 = site_of(c, l, 3);
    if (fused_attn) {
      // one kernel per layer: S and P stay on chip (csrc/perceiver_attn.cu)
      RUN(vq_perceiver_attention_fwd(L.q, L.kv, L.P, L.Pd, L.att, d.B, d.H, d.Lq, d.T, E, &dr0, st));
    } else {
      // scores[b,h] = scale * q[b,:,h] k[b,:,h]^T  -> f32 [A, Tp]
      {
        vq_gemm_desc g;
        memset(&g, 0, sizeof(g));
        g.M = d.Lq; g.N = d.T; g.K = d.hd;
        g.Z0 = d.B; g.Z1 = d.H;
        g.A = L.q; g.lda = E; g.a_r0 = d.Lq; g.a_c1 = d.hd; g.a_rows = R; g.a_cols = E;
        g.B = L.kv; g.ldb = 2 * E; g.b_r0 = d.T; g.b_c1 = d.hd; g.b_rows = C; g.b_cols = E;
        g.C = k.S; g.ldc = d.Tp; g.c_dtype = VQ_F32; g.c_r0 = (long long)d.H * d.Lq; g.c_r1 = d.Lq;
        g.alpha = scale;
        RUN(vq_gemm(&g, st));
      }
      RUN(vq_softmax_fwd(k.S, L.P, d.A, d.T, d.Tp, d.Tp, &dr0, L.Pd, st));
      // att[b,:,h] = P[b,h] v[b,:,h]   (v is MN-major: keys along rows)
      {
        vq_gemm_desc g;
        memset(&g, 0, sizeof(g));
        g.M = d.Lq; g.N = d.hd; g.K = d.Tp;
        g.Z0 = d.B; g.Z1 = d.H;
        g.A = L.Pd; g.lda = d.Tp; g.a_r0 = (long long)d.H * d.Lq; g.a_r1 = d.Lq; g.a_rows = d.A; g.a_cols = d.Tp;
        g.B = reinterpret_cast<const __nv_bfloat16*>(L.kv) + E; g.ldb = 2 * E; g.b_major = 1;
        g.b_r0 = d.T; g.b_c1 = d.hd; g.b_rows = C; g.b_cols = E;
        g.C = L.att; g.ldc = E; g.c_dtype = VQ_BF16; g.c_r0 = d.Lq; g.c_c1 = d.hd;
        RUN(vq_gemm(&g, st));
      }
    }
    // z1 = lat + drop(att Wo^T + bo) ; x1 = LN1(z1)
    RUN(gemm(st, R, E, E, L.att, E, 0, lw.out_proj_w, E, 0, L.z1, E, VQ_BF16, lw.out_proj_b, s.lat[l], E, VQ_EPI_LINEAR,
             nullptr, 0, &dr1));
    RUN(vq_layernorm(L.z1, VQ_BF16, L.x1, VQ_BF16, lw.norm1_w, lw.norm1_b, R, E, c->ln_eps, L.mean1, L.rstd1, st));
    // g = drop(gelu(x1 W1^T + b1)) (pre-activation kept) ; z2 = x1 + drop(g W2^T + b2) ; lat' = LN2(z2)
    RUN(gemm(st, R, F, E, L.x1, E, 0, lw.fc1_w, E, 0, L.g, F, VQ_BF16, lw.fc1_b, nullptr, 0, VQ_EPI_GELU, L.pre, F,
             &dr2));
    RUN(gemm(st, R, E, F, L.g, F, 0, lw.fc2_w, F, 0, L.z2, E, VQ_BF16, lw.fc2_b, L.x1, E, VQ_EPI_LINEAR, nullptr, 0,
             &dr3));
    RUN(vq_layernorm(L.z2, VQ_BF16, s.lat[l + 1], VQ_BF16, lw.norm2_w, lw.norm2_b, R, E, c->ln_eps, L.mean2, L.rstd2,
                     st));
  }
  RUN(gemm(st, R, c->out_dim, E, s.lat[c->layers], E, 0, w->out_proj_w, E, 0, out, c->out_dim, c->out_dtype,
           w->out_proj_b));
  return VQ_OK;
}

extern "C" int vq_perceiver_backward(const vq_perceiver_cfg* c, const vq_perceiver_weights* w, const void* tokens,
                                     const void* grad_out, const void* saved, const vq_perceiver_grads* gr,
                                     float* grad_tokens, void* workspace, vq_layer_done_fn on_layer_done,
                                     void* user, vq_stream_t st) {
  RUN(check_cfg(c));
  VQ_CHECK_ARG(w && tokens && grad_out && saved && gr && workspace, "null pointer");
  VQ_CHECK_ARG((reinterpret_cast<uintptr_t>(saved) % 1024) == 0 && (reinterpret_cast<uintptr_t>(workspace) % 1024) == 0,
               "saved/workspace must be 1024-byte aligned");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(st);
  const Dims d = dims_of(c);
  const Saved s = carve_saved(c, const_cast<void*>(saved));
  const Work k = carve_work(c, workspace);
  const int E = d.E, F = d.F, R = (int)d.R, C = (int)d.C;
  const float scale = 1.0f / sqrtf((float)d.hd);
  const bool fused_attn = vq_perceiver_attention_supported(d.Lq, d.T, E, d.H) != 0;
  const int acc = gr->accumulate;

  // wgrad: dW[N_out, K_in] (+)= dY^T X, both operands MN-major over the row (contraction) dimension; f32, or bf16
  // straight from the fp32 accumulators when the caller reduces / steps on bf16 gradients
  const int wdt = gr->weight_grad_bf16 ? VQ_BF16 : VQ_F32;
  const size_t welt = gr->weight_grad_bf16 ? 2 : 4;
  auto wgrad = [&](void* dW, int n_out, int k_in, const void* dY, long long lddy, const void* X, long long ldx,
                   int rows) {
    return gemm(st, n_out, k_in, rows, dY, lddy, 1, X, ldx, 1, dW, k_in, wdt, nullptr, acc ? dW : nullptr, k_in);
  };
  // bias / LN-affine gradients are accumulated by the kernels: clear them first unless accumulating
  auto clear = [&](float* p, size_t n) -> int {
    if (!acc) VQ_CHECK_CUDA(cudaMemsetAsync(p, 0, n * sizeof(float), stream));
    return VQ_OK;
  };

  // ---- out_proj
  RUN(wgrad(gr->out_proj_w, c->out_dim, E, grad_out, c->out_dim, s.lat[c->layers], E, R));
  RUN(clear(gr->out_proj_b, c->out_dim));
  RUN(vq_colsum(grad_out, VQ_BF16, gr->out_proj_b, R, c->out_dim, c->out_dim, st));
  RUN(gemm(st, R, E, c->out_dim, grad_out, c->out_dim, 0, w->out_proj_w, E, 1, k.dlat, E, VQ_BF16));
  if (on_layer_done) on_layer_done(c->layers, user);
  VQ_CHECK_CUDA(cudaMemsetAsync(k.dctx, 0, (size_t)d.C * E * sizeof(float), stream));

  for (int l = c->layers - 1; l >= 0; --l) {
    const vq_perceiver_layer_weights& lw = w->layers[l];
    const vq_perceiver_layer_grads& lg = gr->layers[l];
    const LayerSaved& L = s.layer[l];
    const __nv_bfloat16* Wqkv = reinterpret_cast<const __nv_bfloat16*>(lw.in_proj_w);
    const bool drop = c->dropout_p > 0.f;
    const vq_dropout_desc dr0 = site_of(c, l, 0), dr1 = site_of(c, l, 1), dr2 = site_of(c, l, 2), dr3 = site_of(c, l, 3);
    // LN2 backward: dz2 (residual path) and dm = mask3 * dz2 / (1-p) (the branch through the post-MLP dropout)
    RUN(clear(lg.norm2_w, E)); RUN(clear(lg.norm2_b, E));
    RUN(vq_layernorm_bwd(k.dlat, L.z2, L.mean2, L.rstd2, lw.norm2_w, k.dz2, lg.norm2_w, lg.norm2_b, R, E, &dr3, k.ddrop,
                         st));
    const char* dm = drop ? k.ddrop : k.dz2;
    // MLP backward
    RUN(wgrad(lg.fc2_w, E, F, dm, E, L.g, F, R));
    RUN(clear(lg.fc2_b, E));
    RUN(vq_colsum(dm, VQ_BF16, lg.fc2_b, R, E, E, st));
    RUN(gemm(st, R, F, E, dm, E, 0, lw.fc2_w, F, 1, k.dpre, F, VQ_BF16, nullptr, nullptr, 0, VQ_EPI_GELU_BWD, L.pre, F,
             &dr2));
    RUN(wgrad(lg.fc1_w, F, E, k.dpre, F, L.x1, E, R));
    RUN(clear(lg.fc1_b, F));
    RUN(vq_colsum(k.dpre, VQ_BF16, lg.fc1_b, R, F, F, st));
    RUN(gemm(st, R, E, F, k.dpre, F, 0, lw.fc1_w, E, 1, k.dx1, E, VQ_BF16, nullptr, k.dz2, E));
    // LN1 backward: dz1 (residual path) and da = mask1 * dz1 / (1-p) (the branch through the post-attention dropout)
    RUN(clear(lg.norm1_w, E)); RUN(clear(lg.norm1_b, E));
    RUN(vq_layernorm_bwd(k.dx1, L.z1, L.mean1, L.rstd1, lw.norm1_w, k.dz1, lg.norm1_w, lg.norm1_b, R, E, &dr1, k.ddrop,
                         st));
    const char* da = drop ? k.ddrop : k.dz1;
    // attention out-projection
    RUN(wgrad(lg.out_proj_w, E, E, da, E, L.att, E, R));
    RUN(clear(lg.out_proj_b, E));
    RUN(vq_colsum(da, VQ_BF16, lg.out_proj_b, R, E, E, st));
    RUN(gemm(st, R, E, E, da, E, 0, lw.out_proj_w, E, 1, k.datt, E, VQ_BF16));
    // attention core, per (sample, head)
    __nv_bfloat16* dkv = reinterpret_cast<__nv_bfloat16*>(k.dkv);
    const __nv_bfloat16* kv = reinterpret_cast<const __nv_bfloat16*>(L.kv);
    if (fused_attn) {
      RUN(vq_perceiver_attention_bwd(L.q, L.kv, L.P, k.datt, k.dq, k.dkv, d.B, d.H, d.Lq, d.T, E, &dr0, st));
    } else {
      {  // dV = Pd^T datt  -> dkv[:, E + h*hd]   (Pd = dropout(P), = P in eval mode)
        vq_gemm_desc g;
        memset(&g, 0, sizeof(g));
        g.M = d.T; g.N = d.hd; g.K = d.Lq; g.Z0 = d.B; g.Z1 = d.H;
        g.A = L.Pd; g.lda = d.Tp; g.a_major = 1; g.a_r0 = (long long)d.H * d.Lq; g.a_r1 = d.Lq; g.a_rows = d.A; g.a_cols = d.Tp;
        g.B = k.datt; g.ldb = E; g.b_major = 1; g.b_r0 = d.Lq; g.b_c1 = d.hd; g.b_rows = R; g.b_cols = E;
        g.C = dkv + E; g.ldc = 2 * E; g.c_dtype = VQ_BF16; g.c_r0 = d.T; g.c_c1 = d.hd;
        RUN(vq_gemm(&g, st));
      }
      {  // dP = datt v^T -> f32 [A, Tp]
        vq_gemm_desc g;
        memset(&g, 0, sizeof(g));
        g.M = d.Lq; g.N = d.T; g.K = d.hd; g.Z0 = d.B; g.Z1 = d.H;
        g.A = k.datt; g.lda = E; g.a_r0 = d.Lq; g.a_c1 = d.hd; g.a_rows = R; g.a_cols = E;
        g.B = kv + E; g.ldb = 2 * E; g.b_r0 = d.T; g.b_c1 = d.hd; g.b_rows = C; g.b_cols = E;
        g.C = k.S; g.ldc = d.Tp; g.c_dtype = VQ_F32; g.c_r0 = (long long)d.H * d.Lq; g.c_r1 = d.Lq;
        RUN(vq_gemm(&g, st));
      }
      RUN(vq_softmax_bwd(L.P, k.S, k.dS, d.A, d.T, d.Tp, d.Tp, scale, &dr0, st));
      {  // dq = dS k -> [R, E]
        vq_gemm_desc g;
        memset(&g, 0, sizeof(g));
        g.M = d.Lq; g.N = d.hd; g.K = d.Tp; g.Z0 = d.B; g.Z1 = d.H;
        g.A = k.dS; g.lda = d.Tp; g.a_r0 = (long long)d.H * d.Lq; g.a_r1 = d.Lq; g.a_rows = d.A; g.a_cols = d.Tp;
        g.B = kv; g.ldb = 2 * E; g.b_major = 1; g.b_r0 = d.T; g.b_c1 = d.hd; g.b_rows = C; g.b_cols = E;
        g.C = k.dq; g.ldc = E; g.c_dtype = VQ_BF16; g.c_r0 = d.Lq; g.c_c1 = d.hd;
        RUN(vq_gemm(&g, st));
      }
      {  // dk = dS^T q -> dkv[:, h*hd]
        vq_gemm_desc g;
        memset(&g, 0, sizeof(g));
        g.M = d.T; g.N = d.hd; g.K = d.Lq; g.Z0 = d.B; g.Z1 = d.H;
        g.A = k.dS; g.lda = d.Tp; g.a_major = 1; g.a_r0 = (long long)d.H * d.Lq; g.a_r1 = d.Lq; g.a_rows = d.A; g.a_cols = d.Tp;
        g.B = L.q; g.ldb = E; g.b_major = 1; g.b_r0 = d.Lq; g.b_c1 = d.hd; g.b_rows = R; g.b_cols = E;
        g.C = dkv; g.ldc = 2 * E; g.c_dtype = VQ_BF16; g.c_r0 = d.T; g.c_c1 = d.hd;
        RUN(vq_gemm(&g, st));
      }
    }
    // projections: in_proj_weight.grad rows [0:E] <- q, rows [E:3E] <- k,v
    RUN(wgrad(lg.in_proj_w, E, E, k.dq, E, s.lat[l], E, R));
    RUN(wgrad(reinterpret_cast<char*>(lg.in_proj_w) + (size_t)E * E * welt, 2 * E, E, k.dkv, 2 * E, s.ctx, E, C));
    RUN(clear(lg.in_proj_b, 3 * (size_t)E));
    RUN(vq_colsum(k.dq, VQ_BF16, lg.in_proj_b, R, E, E, st));
    RUN(vq_colsum(k.dkv, VQ_BF16, lg.in_proj_b + E, C, 2 * E, 2 * E, st));
    // dlat_prev = dz1 + dq Wq ; dctx += dkv [Wk;Wv]
    RUN(gemm(st, R, E, E, k.dq, E, 0, Wqkv, E, 1, k.dlat, E, VQ_BF16, nullptr, k.dz1, E));
    RUN(gemm(st, C, E, 2 * E, k.dkv, 2 * E, 0, Wqkv + (size_t)E * E, E, 1, k.dctx, E, VQ_F32, nullptr, k.dctx, E));
    if (on_layer_done) on_layer_done(l, user);
  }

  // ---- latents (sum over the batch) and in_proj
  RUN(vq_reduce_batch(k.dlat, gr->latents, (long long)d.Lq * E, d.B, acc, st));
  RUN(vq_cast_f32_bf16(k.dctx, k.dctx_bf16, (long long)d.C * E, st));
  RUN(wgrad(gr->in_proj_w, E, c->in_dim, k.dctx_bf16, E, tokens, c->in_dim, C));
  RUN(clear(gr->in_proj_b, E));
  RUN(vq_colsum(k.dctx, VQ_F32, gr->in_proj_b, C, E, E, st));
  if (grad_tokens)
    RUN(gemm(st, C, c->in_dim, E, k.dctx_bf16, E, 0, w->in_proj_w, c->in_dim, 1, grad_tokens, c->in_dim, VQ_F32));
  if (on_layer_done) on_layer_done(-1, user);
  return VQ_OK;
}
