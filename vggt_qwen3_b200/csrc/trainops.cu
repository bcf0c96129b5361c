// HBM-bound kernels either side of the encode path (SURVEY.md §8f "next" rows):
//   f1  vq_splice_insert : the inference splice (src/inference/qa_inference.py:119-145), no host-side torch.cat
//   f3  vq_adamw_step    : fused AdamW for the projector parameter group (src/train/train_sft.py:138-163,
//                          torch.optim.AdamW defaults) that also refreshes the bf16 compute copy of the weights
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace vq {

// ---- first <image> position in (batch, position) order, as `(input_ids == id).nonzero()[0]` finds it
__global__ void __launch_bounds__(256)
find_first_token_kernel(const long long* __restrict__ ids, long long n, long long token, long long* __restrict__ first) {
  long long best = n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (ids[i] == token) { best = i; break; }   // a thread's indices ascend: its first hit is its smallest
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other < best ? other : best;
  }
  if ((threadIdx.x & 31) == 0 && best < n) atomicMin(reinterpret_cast<unsigned long long*>(first), (unsigned long long)best);
}

// out[b, r] = emb[b, r] (r < pos) | vis[b, r - pos] (pos <= r < pos + n) | emb[b, r - n + 1] (otherwise); one warp per
// output row.  mask_out likewise with ones over the inserted span.  pos = *first % L (the reference slices every
// sample at the position of the FIRST hit); if there is no hit (*first == B * L) nothing is written.
template <typename T, typename M>
__global__ void __launch_bounds__(256)
splice_insert_kernel(const T* __restrict__ emb, const M* __restrict__ mask, const T* __restrict__ vis,
                     T* __restrict__ out, M* __restrict__ mask_out, const long long* __restrict__ first, int B, int L,
                     int Hd, int n) {
  const long long f = *first;
  if (f >= (long long)B * L) return;
  const int pos = (int)(f % L);
  const int Lo = L - 1 + n;
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (long long)B * Lo) return;
  const int b = (int)(w / Lo), r = (int)(w - (long long)b * Lo);
  const T* src;
  M mval;
  if (r < pos) {
    src = emb + ((long long)b * L + r) * Hd;
    mval = mask[(long long)b * L + r];
  } else if (r < pos + n) {
    src = vis + ((long long)b * n + (r - pos)) * Hd;
    mval = (M)1;
  } else {
    src = emb + ((long long)b * L + (r - n + 1)) * Hd;
    mval = mask[(long long)b * L + (r - n + 1)];
  }
  T* dst = out + ((long long)b * Lo + r) * Hd;
  constexpr int kVec = 16 / sizeof(T);
  for (int c = lane * kVec; c < Hd; c += 32 * kVec)
    *reinterpret_cast<uint4*>(dst + c) = *reinterpret_cast<const uint4*>(src + c);
  if (lane == 0) mask_out[(long long)b * Lo + r] = mval;
}

// ---- AdamW, torch.optim.AdamW semantics (decoupled weight decay, bias correction, eps outside the sqrt of v_hat):
//   p *= 1 - lr * wd;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;
//   p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps),   bc1 = 1 - b1^t, bc2 = 1 - b2^t
// g = grad * grad_scale (1 / world size, 1 / accumulation steps, loss-scale inverse ...).  4 elements per thread with
// 128-bit accesses; 16 B/element read (p, m, v, g) + 12 B written (+ 2 B for the bf16 compute copy).
struct AdamArgs {
  float lr, beta1, beta2, eps, weight_decay, grad_scale, bc1, bc2_sqrt;
};
constexpr int kMaxGradSources = 16;
struct GradSources {
  const void* src[kMaxGradSources];   // same dtype, same length: the gradient is their sum (accumulated in fp32)
  int n;
};

__device__ __forceinline__ void adam_update(float& p, float& m, float& v, float g, const AdamArgs& a, float decay,
                                            float step) {
  g *= a.grad_scale;
  m = a.beta1 * m + (1.0f - a.beta1) * g;
  v = a.beta2 * v + (1.0f - a.beta2) * g * g;
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
  p = p * decay - step * (m / denom);
}

// GRAD_BF16: dtype of the gradient sources; W_BF16: dtype of the optional compute copy (bf16 for matrices, f32 for the
// vectors the kernels read in f32).  With more than one source this is the reduce step of a reduce-scatter fused into
// the optimizer: the shard of every rank's gradient (pulled over NVLink into local staging, or read from peer-mapped
// memory) is summed in fp32 and never written back.
template <bool GRAD_BF16, bool W_BF16>
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const GradSources gs,
             void* __restrict__ w_out, long long n, AdamArgs a) {
  const float decay = 1.0f - a.lr * a.weight_decay;
  const float step = a.lr / a.bc1;
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<const float4*>(p)[i];
    float4 mm = reinterpret_cast<const float4*>(m)[i];
    float4 vv = reinterpret_cast<const float4*>(v)[i];
    float g[4] = {0.f, 0.f, 0.f, 0.f};
    for (int s = 0; s < gs.n; ++s) {
      if constexpr (GRAD_BF16) {
        const uint2 u = reinterpret_cast<const uint2*>(gs.src[s])[i];
        g[0] += bf16_lo(u.x); g[1] += bf16_hi(u.x); g[2] += bf16_lo(u.y); g[3] += bf16_hi(u.y);
      } else {
        const float4 gg = reinterpret_cast<const float4*>(gs.src[s])[i];
        g[0] += gg.x; g[1] += gg.y; g[2] += gg.z; g[3] += gg.w;
      }
    }
    adam_update(pp.x, mm.x, vv.x, g[0], a, decay, step);
    adam_update(pp.y, mm.y, vv.y, g[1], a, decay, step);
    adam_update(pp.z, mm.z, vv.z, g[2], a, decay, step);
    adam_update(pp.w, mm.w, vv.w, g[3], a, decay, step);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (w_out) {
      if constexpr (W_BF16) {
        uint2 o;
        o.x = pack_bf16x2(pp.x, pp.y);
        o.y = pack_bf16x2(pp.z, pp.w);
        reinterpret_cast<uint2*>(w_out)[i] = o;
      } else {
        reinterpret_cast<float4*>(w_out)[i] = pp;
      }
    }
  }
  // scalar tail (n % 4 elements), one thread
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (long long i = n4 << 2; i < n; ++i) {
      float g = 0.f;
      for (int s = 0; s < gs.n; ++s)
        g += GRAD_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(gs.src[s])[i])
                       : reinterpret_cast<const float*>(gs.src[s])[i];
      adam_update(p[i], m[i], v[i], g, a, decay, step);
      if (w_out) {
        if constexpr (W_BF16) reinterpret_cast<__nv_bfloat16*>(w_out)[i] = __float2bfloat16(p[i]);
        else reinterpret_cast<float*>(w_out)[i] = p[i];
      }
    }
  }
}

}  // namespace vq

using namespace vq;

extern "C" int vq_find_first_token(const int64_t* input_ids, int64_t n, int64_t token, int64_t* first,
                                   vq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VQ_CHECK_ARG(input_ids && first && n > 0, "find_first_token arguments");
  {
    // *first must hold n ("not found") on entry; the kernel lowers it with atomicMin
    LaunchScope scope(kKindRow, "find_first_token", 1, (int)(n > INT32_MAX ? INT32_MAX : n), 0, 0, (double)n * 8, stream);
    const long long blocks = (n + 255) / 256;
    find_first_token_kernel<<<(unsigned)(blocks > 1184 ? 1184 : blocks), 256, 0, stream>>>(
        reinterpret_cast<const long long*>(input_ids), n, token, reinterpret_cast<long long*>(first));
  }
  VQ_CHECK_CUDA(cudaGetLastError());
  return VQ_OK;
}

extern "C" int vq_splice_insert(const void* inputs_embeds, int32_t dtype, const void* attention_mask,
                                int32_t mask_elt_bytes, const void* vis_tokens, const int64_t* first_pos,
                                void* out_embeds, void* out_mask, int32_t B, int32_t L, int32_t hidden, int32_t n_vis,
                                vq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VQ_CHECK_ARG(inputs_embeds && attention_mask && vis_tokens && first_pos && out_embeds && out_mask, "null pointer");
  VQ_CHECK_ARG(B > 0 && L > 0 && n_vis > 0, "geometry");
  VQ_CHECK_ARG(dtype == VQ_F32 || dtype == VQ_BF16, "dtype");
  VQ_CHECK_ARG(mask_elt_bytes == 8 || mask_elt_bytes == 4 || mask_elt_bytes == 1, "attention_mask must be int64/int32/bool");
  const int elt = dtype == VQ_F32 ? 4 : 2;
  VQ_CHECK_ARG((hidden * elt) % 16 == 0, "hidden size must give 16-byte rows");
  VQ_CHECK_ARG((reinterpret_cast<uintptr_t>(inputs_embeds) % 16) == 0 && (reinterpret_cast<uintptr_t>(vis_tokens) % 16) == 0 &&
                   (reinterpret_cast<uintptr_t>(out_embeds) % 16) == 0,
               "16-byte alignment");
  const long long rows = (long long)B * (L - 1 + n_vis);
  const unsigned grid = (unsigned)((rows + 7) / 8);
  LaunchScope scope(kKindRow, "splice_insert", B, L, n_vis, 0, 2.0 * rows * hidden * elt, stream);
#define VQ_INSERT(T, M)                                                                                          \
  splice_insert_kernel<T, M><<<grid, 256, 0, stream>>>(                                                          \
      reinterpret_cast<const T*>(inputs_embeds), reinterpret_cast<const M*>(attention_mask),                     \
      reinterpret_cast<const T*>(vis_tokens), reinterpret_cast<T*>(out_embeds), reinterpret_cast<M*>(out_mask), \
      reinterpret_cast<const long long*>(first_pos), B, L, hidden, n_vis)
  if (dtype == VQ_F32) {
    if (mask_elt_bytes == 8) VQ_INSERT(float, long long);
    else if (mask_elt_bytes == 4) VQ_INSERT(float, int);
    else VQ_INSERT(float, unsigned char);
  } else {
    if (mask_elt_bytes == 8) VQ_INSERT(__nv_bfloat16, long long);
    else if (mask_elt_bytes == 4) VQ_INSERT(__nv_bfloat16, int);
    else VQ_INSERT(__nv_bfloat16, unsigned char);
  }
#undef VQ_INSERT
  VQ_CHECK_CUDA(cudaGetLastError());
  return VQ_OK;
}

static int adamw_launch(float* param, float* exp_avg, float* exp_avg_sq, const void* const* grad_srcs, int n_srcs,
                        int grad_dtype, void* weight_out, int weight_dtype, long long n, int step, float lr, float beta1,
                        float beta2, float eps, float weight_decay, float grad_scale, cudaStream_t stream) {
  VQ_CHECK_ARG(param && exp_avg && exp_avg_sq && grad_srcs, "null pointer");
  VQ_CHECK_ARG(n > 0, "element count must be positive");
  VQ_CHECK_ARG(n_srcs >= 1 && n_srcs <= kMaxGradSources, "1..16 gradient sources");
  VQ_CHECK_ARG(step >= 1, "step counts from 1");
  VQ_CHECK_ARG(grad_dtype == VQ_F32 || grad_dtype == VQ_BF16, "grad dtype");
  VQ_CHECK_ARG(weight_dtype == VQ_F32 || weight_dtype == VQ_BF16, "weight copy dtype");
  VQ_CHECK_ARG(grad_scale != 0.f && grad_scale == grad_scale, "grad_scale must be a non-zero number");
  VQ_CHECK_ARG(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(exp_avg) |
                 reinterpret_cast<uintptr_t>(exp_avg_sq)) % 16) == 0,
               "param / state 16-byte alignment");
  GradSources gs;
  gs.n = n_srcs;
  for (int i = 0; i < n_srcs; ++i) {
    VQ_CHECK_ARG(grad_srcs[i] && (reinterpret_cast<uintptr_t>(grad_srcs[i]) % (grad_dtype == VQ_F32 ? 16 : 8)) == 0,
                 "gradient source alignment");
    gs.src[i] = grad_srcs[i];
  }
  VQ_CHECK_ARG(!weight_out || (reinterpret_cast<uintptr_t>(weight_out) % (weight_dtype == VQ_F32 ? 16 : 8)) == 0,
               "weight copy alignment");
  AdamArgs a;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
  a.grad_scale = grad_scale;
  a.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  a.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  const long long n4 = n / 4 > 0 ? n / 4 : 1;
  const int sms = num_sms() > 0 ? num_sms() : 148;
  long long blocks = (n4 + 255) / 256;
  static const int ctas_per_sm = [] {   // grid-stride; 8 resident CTAs of 256 threads per SM unless overridden (tuning)
    const char* e = getenv("VQ_ADAMW_CTAS_PER_SM");
    const int v = e ? atoi(e) : 8;
    return v < 1 ? 1 : v > 64 ? 64 : v;
  }();
  if (blocks > (long long)ctas_per_sm * sms) blocks = (long long)ctas_per_sm * sms;
  const double gbytes = grad_dtype == VQ_F32 ? 4.0 : 2.0, wbytes = !weight_out ? 0.0 : weight_dtype == VQ_F32 ? 4.0 : 2.0;
  LaunchScope scope(kKindRow, n_srcs > 1 ? "adamw_reduce_step" : "adamw_step", n_srcs, 0, 0, 0,
                    (double)n * (24.0 + gbytes * n_srcs + wbytes), stream);
#define VQ_ADAM(G, W) \
  adamw_kernel<G, W><<<(unsigned)blocks, 256, 0, stream>>>(param, exp_avg, exp_avg_sq, gs, weight_out, n, a)
  if (grad_dtype == VQ_BF16) {
    if (weight_dtype == VQ_BF16) VQ_ADAM(true, true); else VQ_ADAM(true, false);
  } else {
    if (weight_dtype == VQ_BF16) VQ_ADAM(false, true); else VQ_ADAM(false, false);
  }
#undef VQ_ADAM
  VQ_CHECK_CUDA(cudaGetLastError());
  return VQ_OK;
}

extern "C" int vq_adamw_step(float* param, float* exp_avg, float* exp_avg_sq, const void* grad, int32_t grad_dtype,
                             void* param_bf16, int64_t n, int32_t step, float lr, float beta1, float beta2, float eps,
                             float weight_decay, float grad_scale, vq_stream_t stream_) {
  const void* srcs[1] = {grad};
  return adamw_launch(param, exp_avg, exp_avg_sq, srcs, 1, grad_dtype, param_bf16, VQ_BF16, n, step, lr, beta1, beta2, eps,
                      weight_decay, grad_scale, reinterpret_cast<cudaStream_t>(stream_));
}

extern "C" int vq_adamw_reduce_step(float* param, float* exp_avg, float* exp_avg_sq, const void* const* grad_srcs,
                                    int32_t n_srcs, int32_t grad_dtype, void* weight_out, int32_t weight_dtype, int64_t n,
                                    int32_t step, float lr, float beta1, float beta2, float eps, float weight_decay,
                                    float grad_scale, vq_stream_t stream_) {
  return adamw_launch(param, exp_avg, exp_avg_sq, grad_srcs, n_srcs, grad_dtype, weight_out, weight_dtype, n, step, lr,
                      beta1, beta2, eps, weight_decay, grad_scale, reinterpret_cast<cudaStream_t>(stream_));
}
