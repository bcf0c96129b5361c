// HBM-bound kernels either side of the encode path (SURVEY.md §8f "next" rows):
//   f1  vq_splice_insert : the inference splice (src/inference/qa_inference.py:119-145), no host-side torch.cat
//   f3  vq_adamw_step    : fused AdamW for the projector parameter group (src/train/train_sft.py:138-163,
//                          torch.optim.AdamW defaults) that also refreshes the bf16 compute copy of the weights
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

template <bool GRAD_BF16>
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const void* __restrict__ grad,
             __nv_bfloat16* __restrict__ p_bf16, long long n4, AdamArgs a) {
  const float decay = 1.0f - a.lr * a.weight_decay;
  const float step = a.lr / a.bc1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<const float4*>(p)[i];
    float4 mm = reinterpret_cast<const float4*>(m)[i];
    float4 vv = reinterpret_cast<const float4*>(v)[i];
    float g[4];
    if constexpr (GRAD_BF16) {
      const uint2 u = reinterpret_cast<const uint2*>(grad)[i];
      g[0] = bf16_lo(u.x); g[1] = bf16_hi(u.x); g[2] = bf16_lo(u.y); g[3] = bf16_hi(u.y);
    } else {
      const float4 gg = reinterpret_cast<const float4*>(grad)[i];
      g[0] = gg.x; g[1] = gg.y; g[2] = gg.z; g[3] = gg.w;
    }
    float* pe = &pp.x;
    float* me = &mm.x;
    float* ve = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = g[k] * a.grad_scale;
      me[k] = a.beta1 * me[k] + (1.0f - a.beta1) * gk;
      ve[k] = a.beta2 * ve[k] + (1.0f - a.beta2) * gk * gk;
      const float denom = sqrtf(ve[k]) / a.bc2_sqrt + a.eps;
      pe[k] = pe[k] * decay - step * (me[k] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (p_bf16) {
      uint2 o;
      o.x = pack_bf16x2(pp.x, pp.y);
      o.y = pack_bf16x2(pp.z, pp.w);
      reinterpret_cast<uint2*>(p_bf16)[i] = o;
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

extern "C" int vq_adamw_step(float* param, float* exp_avg, float* exp_avg_sq, const void* grad, int32_t grad_dtype,
                             void* param_bf16, int64_t n, int32_t step, float lr, float beta1, float beta2, float eps,
                             float weight_decay, float grad_scale, vq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VQ_CHECK_ARG(param && exp_avg && exp_avg_sq && grad, "null pointer");
  VQ_CHECK_ARG(n > 0 && n % 4 == 0, "element count must be a positive multiple of 4");
  VQ_CHECK_ARG(step >= 1, "step counts from 1");
  VQ_CHECK_ARG(grad_dtype == VQ_F32 || grad_dtype == VQ_BF16, "grad dtype");
  VQ_CHECK_ARG(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(exp_avg) |
                 reinterpret_cast<uintptr_t>(exp_avg_sq)) % 16) == 0,
               "param / state 16-byte alignment");
  VQ_CHECK_ARG((reinterpret_cast<uintptr_t>(grad) % (grad_dtype == VQ_F32 ? 16 : 8)) == 0, "grad alignment");
  VQ_CHECK_ARG(!param_bf16 || (reinterpret_cast<uintptr_t>(param_bf16) % 8) == 0, "bf16 copy alignment");
  AdamArgs a;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
  a.grad_scale = grad_scale == 0.f ? 1.f : grad_scale;
  a.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  a.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  const long long n4 = n / 4;
  const int sms = num_sms() > 0 ? num_sms() : 148;
  long long blocks = (n4 + 255) / 256;
  if (blocks > 8LL * sms) blocks = 8LL * sms;   // grid-stride: 8 resident CTAs of 256 threads per SM
  const double bytes = (double)n * (12.0 + (grad_dtype == VQ_F32 ? 4.0 : 2.0) + 12.0 + (param_bf16 ? 2.0 : 0.0));
  LaunchScope scope(kKindRow, "adamw_step", 1, 0, 0, 0, bytes, stream);
  if (grad_dtype == VQ_BF16)
    adamw_kernel<true><<<(unsigned)blocks, 256, 0, stream>>>(param, exp_avg, exp_avg_sq, grad,
                                                             reinterpret_cast<__nv_bfloat16*>(param_bf16), n4, a);
  else
    adamw_kernel<false><<<(unsigned)blocks, 256, 0, stream>>>(param, exp_avg, exp_avg_sq, grad,
                                                              reinterpret_cast<__nv_bfloat16*>(param_bf16), n4, a);
  VQ_CHECK_CUDA(cudaGetLastError());
  return VQ_OK;
}
