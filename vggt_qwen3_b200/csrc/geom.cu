// Geometry tokens and the fused training splice:
//   vq_geom_head_forward / _backward : encode_geom + geom_head of the reference
//                                      (src/models/vggt_qwen3_vlm.py:51-56 Linear(37,H)-SiLU-Linear(H,H), :164-177
//                                      cat(R, t, K, depth_hist) -> mean over views -> head -> expand to geom_tokens)
//   vq_splice_features / _backward   : inputs_embeds[b, pos:pos+n] = cat([geom.expand(n_geom), vis])[b]
//                                      (src/models/vggt_qwen3_vlm.py:187-195) without materialising the expand / cat,
//                                      and its autograd backward (gather of the written rows) as one launch.
// All of it is HBM-bound byte moving plus a 6.6 M-parameter MLP on B rows (weight-read bound: 26 MB of f32 weights).
#include "common.cuh"
#include "ptx.cuh"

namespace vq {

constexpr int kGeomMaxIn = 64;    // pooled feature width (37 in the reference)
constexpr int kGeomBatch = 8;     // samples per pass of the H x H layer (accumulators in registers)

__device__ __forceinline__ float wsum_g(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct GeomParts {
  const float* ptr[4];   // [B, V, width[k]] each; nullptr = zeros
  int width[4];
};

// pooled[b, :] = mean over views of cat(parts); pre[b, j] = w1[j, :] . pooled[b] + b1[j]; act = silu(pre).
// grid = (ceil(H / 256), B): every CTA recomputes the 37 pooled features of its sample (a few hundred loads).
__global__ void __launch_bounds__(256)
geom_fc1_kernel(GeomParts parts, const float* __restrict__ w1, const float* __restrict__ b1, float* __restrict__ pooled,
                float* __restrict__ pre, float* __restrict__ act, int V, int in_dim, int H) {
  __shared__ float feat[kGeomMaxIn];
  const int b = blockIdx.y;
  if (threadIdx.x < in_dim) {
    int i = threadIdx.x, k = 0;
    while (i >= parts.width[k]) { i -= parts.width[k]; ++k; }
    float s = 0.f;
    if (parts.ptr[k]) {
      const float* src = parts.ptr[k] + (long long)b * V * parts.width[k] + i;
      for (int v = 0; v < V; ++v) s += src[(long long)v * parts.width[k]];
    }
    s /= (float)V;
    feat[threadIdx.x] = s;
    if (blockIdx.x == 0) pooled[b * in_dim + threadIdx.x] = s;
  }
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= H) return;
  float a = b1[j];
  for (int i = 0; i < in_dim; ++i) a += w1[(long long)j * in_dim + i] * feat[i];
  pre[(long long)b * H + j] = a;
  act[(long long)b * H + j] = a / (1.0f + expf(-a));
}

// y[b, j] = w2[j, :] . act[b, :] + b2[j]: one warp per output feature j, the row of w2 is read once (coalesced
// float4) for all samples of the pass.
__global__ void __launch_bounds__(256)
geom_fc2_kernel(const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ act,
                float* __restrict__ y, int B, int H) {
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= H) return;
  const float* wr = w2 + (long long)j * H;
  for (int b0 = 0; b0 < B; b0 += kGeomBatch) {
    float acc[kGeomBatch];
#pragma unroll
    for (int q = 0; q < kGeomBatch; ++q) acc[q] = 0.f;
    for (int k = lane * 4; k < H; k += 128) {
      const float4 w = *reinterpret_cast<const float4*>(wr + k);
#pragma unroll
      for (int q = 0; q < kGeomBatch; ++q) {
        if (b0 + q < B) {
          const float4 a = *reinterpret_cast<const float4*>(act + (long long)(b0 + q) * H + k);
          acc[q] += w.x * a.x + w.y * a.y + w.z * a.z + w.w * a.w;
        }
      }
    }
#pragma unroll
    for (int q = 0; q < kGeomBatch; ++q) {
      const float s = wsum_g(acc[q]);
      if (lane == 0 && b0 + q < B) y[(long long)(b0 + q) * H + j] = s + b2[j];
    }
  }
}

// ---- backward.  dw2[j, k] (+)= sum_b dy[b, j] act[b, k];  db2[j] (+)= sum_b dy[b, j]   (one thread per 4 k of a row j)
__global__ void __launch_bounds__(256)
geom_dw2_kernel(const float* __restrict__ dy, const float* __restrict__ act, float* __restrict__ dw2,
                float* __restrict__ db2, int B, int H, int accumulate) {
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int j = blockIdx.y;
  if (k >= H) return;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  float sb = 0.f;
  for (int b = 0; b < B; ++b) {
    const float d = dy[(long long)b * H + j];
    const float4 h = *reinterpret_cast<const float4*>(act + (long long)b * H + k);
    a.x += d * h.x; a.y += d * h.y; a.z += d * h.z; a.w += d * h.w;
    sb += d;
  }
  float4* dst = reinterpret_cast<float4*>(dw2 + (long long)j * H + k);
  if (accumulate) {
    const float4 o = *dst;
    a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
  }
  *dst = a;
  if (k == 0) db2[j] = accumulate ? db2[j] + sb : sb;
}

// dpre[b, k] = silu'(pre[b, k]) * sum_j dy[b, j] w2[j, k]: one thread per column k (coalesced over k), j walked in
// slabs by blockIdx.y and combined with one atomic per (b, k) and slab; dpre must be zeroed first.  The SiLU
// derivative is applied by geom_dw1_kernel (it needs the complete sum).
__global__ void __launch_bounds__(128)
geom_dact_kernel(const float* __restrict__ dy, const float* __restrict__ w2, float* __restrict__ dact, int B, int H,
                 int rows_per_slab) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= H) return;
  const int j0 = blockIdx.y * rows_per_slab, j1 = min(H, j0 + rows_per_slab);
  for (int b0 = 0; b0 < B; b0 += kGeomBatch) {
    float acc[kGeomBatch];
#pragma unroll
    for (int q = 0; q < kGeomBatch; ++q) acc[q] = 0.f;
    for (int j = j0; j < j1; ++j) {
      const float w = w2[(long long)j * H + k];
#pragma unroll
      for (int q = 0; q < kGeomBatch; ++q)
        if (b0 + q < B) acc[q] += __ldg(dy + (long long)(b0 + q) * H + j) * w;
    }
#pragma unroll
    for (int q = 0; q < kGeomBatch; ++q)
      if (b0 + q < B) atomicAdd(dact + (long long)(b0 + q) * H + k, acc[q]);
  }
}

// dw1[j, i] (+)= sum_b dpre[b, j] pooled[b, i];  db1[j] (+)= sum_b dpre[b, j];  dpre = dact * silu'(pre)
__global__ void __launch_bounds__(256)
geom_dw1_kernel(const float* __restrict__ dact, const float* __restrict__ pre, const float* __restrict__ pooled,
                float* __restrict__ dw1, float* __restrict__ db1, int B, int in_dim, int H, int accumulate) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= H) return;
  float sb = 0.f;
  float acc[kGeomMaxIn];
#pragma unroll
  for (int i = 0; i < kGeomMaxIn; ++i) acc[i] = 0.f;
  for (int b = 0; b < B; ++b) {
    const float x = pre[(long long)b * H + j];
    const float s = 1.0f / (1.0f + expf(-x));
    const float d = dact[(long long)b * H + j] * (s * (1.0f + x * (1.0f - s)));
    sb += d;
#pragma unroll
    for (int i = 0; i < kGeomMaxIn; ++i)
      if (i < in_dim) acc[i] += d * __ldg(pooled + b * in_dim + i);
  }
#pragma unroll
  for (int i = 0; i < kGeomMaxIn; ++i)
    if (i < in_dim) {
      float* dst = dw1 + (long long)j * in_dim + i;
      *dst = accumulate ? *dst + acc[i] : acc[i];
    }
  db1[j] = accumulate ? db1[j] + sb : sb;
}

// ------------------------------------------------------------------------------------------------
// Training splice with two feature sources.  Feature row t of sample b is geom[b, :] for t < n_geom (the expanded
// geometry token) and vis[b, t - n_geom, :] otherwise; for every row r of sample b the LAST <image> position
// pos <= r with r - pos < n wins (the reference writes the spans in ascending position order).
// One warp per (b, r).  BWD = the autograd backward of the same assignment: written rows route their gradient to the
// feature they came from (f32 atomics: several <image> positions may read the same feature) and are zeroed in
// d(inputs_embeds).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int wmax_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int winning_position(const long long* row_ids, int r, int n, long long image_id, int lane) {
  int best = -1;
  for (int p = r - lane; p > r - n && p >= 0; p -= 32)
    if (row_ids[p] == image_id) { best = p; break; }   // a lane's candidates descend: its first hit is its largest
  return wmax_i(best);
}

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

template <typename T>
__global__ void __launch_bounds__(256)
splice_features_kernel(T* __restrict__ embeds, const long long* __restrict__ ids, const float* __restrict__ geom,
                       const T* __restrict__ vis, int B, int L, int Hd, int n_geom, int n_vis, long long image_id,
                       int* __restrict__ overflow) {
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (long long)B * L) return;
  const int b = (int)(w / L), r = (int)(w - (long long)b * L);
  const int n = n_geom + n_vis;
  const long long* row_ids = ids + (long long)b * L;
  if (overflow && lane == 0 && row_ids[r] == image_id && r + n > L) atomicAdd(overflow, 1);
  const int best = winning_position(row_ids, r, n, image_id, lane);
  if (best < 0) return;
  const int t = r - best;
  T* dst = embeds + ((long long)b * L + r) * Hd;
  if (t < n_geom) {   // the expanded geometry token: geom_head output (f32) cast to the embedding dtype on the fly
    const float* src = geom + (long long)b * Hd;
    for (int c = lane; c < Hd; c += 32) dst[c] = from_f32<T>(src[c]);
    return;
  }
  const T* src = vis + ((long long)b * n_vis + (t - n_geom)) * Hd;
  constexpr int kVec = 16 / sizeof(T);
  for (int c = lane * kVec; c < Hd; c += 32 * kVec)
    *reinterpret_cast<uint4*>(dst + c) = *reinterpret_cast<const uint4*>(src + c);
}

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__global__ void __launch_bounds__(256)
splice_features_bwd_kernel(const T* __restrict__ grad, const long long* __restrict__ ids, T* __restrict__ d_embeds,
                           float* __restrict__ d_geom, float* __restrict__ d_vis, int B, int L, int Hd, int n_geom,
                           int n_vis, long long image_id) {
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (long long)B * L) return;
  const int b = (int)(w / L), r = (int)(w - (long long)b * L);
  const int n = n_geom + n_vis;
  const int best = winning_position(ids + (long long)b * L, r, n, image_id, lane);
  const T* g = grad + ((long long)b * L + r) * Hd;
  T* de = d_embeds ? d_embeds + ((long long)b * L + r) * Hd : nullptr;
  constexpr int kVec = 16 / sizeof(T);
  if (best < 0) {
    if (de)
      for (int c = lane * kVec; c < Hd; c += 32 * kVec)
        *reinterpret_cast<uint4*>(de + c) = *reinterpret_cast<const uint4*>(g + c);
    return;
  }
  const int t = r - best;
  float* dst = t < n_geom ? (d_geom ? d_geom + (long long)b * Hd : nullptr)
                          : (d_vis ? d_vis + ((long long)b * n_vis + (t - n_geom)) * Hd : nullptr);
  for (int c = lane * kVec; c < Hd; c += 32 * kVec) {
    if (dst) {
#pragma unroll
      for (int e = 0; e < kVec; ++e) atomicAdd(dst + c + e, to_f32<T>(g[c + e]));
    }
    if (de) *reinterpret_cast<uint4*>(de + c) = make_uint4(0u, 0u, 0u, 0u);
  }
}

}  // namespace vq

using namespace vq;

extern "C" int vq_geom_head_forward(const float* const* parts, const int32_t* widths, const float* w1, const float* b1,
                                    const float* w2, const float* b2, float* pooled, float* pre, float* act, float* y,
                                    int32_t B, int32_t V, int32_t in_dim, int32_t H, vq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VQ_CHECK_ARG(parts && widths && w1 && b1 && w2 && b2 && pooled && pre && act && y, "null pointer");
  VQ_CHECK_ARG(B > 0 && V > 0 && in_dim > 0 && in_dim <= kGeomMaxIn && H > 0 && H % 4 == 0, "geometry (in_dim <= 64, H % 4)");
  VQ_CHECK_ARG(B <= 65535, "batch too large");
  GeomParts gp;
  int total = 0;
  for (int k = 0; k < 4; ++k) {
    gp.ptr[k] = parts[k];
    gp.width[k] = widths[k];
    VQ_CHECK_ARG(widths[k] >= 0, "negative part width");
    total += widths[k];
  }
  VQ_CHECK_ARG(total == in_dim, "R | t | K | depth_hist widths must add up to the head's input width");
  VQ_CHECK_ARG(((reinterpret_cast<uintptr_t>(w2) | reinterpret_cast<uintptr_t>(act)) % 16) == 0, "16-byte alignment");
  {
    LaunchScope scope(kKindRow, "geom_fc1", B, H, in_dim, 0, (double)H * in_dim * 4 + 2.0 * B * H * 4, stream);
    geom_fc1_kernel<<<dim3((H + 255) / 256, B), 256, 0, stream>>>(gp, w1, b1, pooled, pre, act, V, in_dim, H);
  }
  VQ_CHECK_CUDA(cudaGetLastError());
  {
    LaunchScope scope(kKindRow, "geom_fc2", B, H, H, 0, (double)H * H * 4 + 2.0 * B * H * 4, stream);
    geom_fc2_kernel<<<(H + 7) / 8, 256, 0, stream>>>(w2, b2, act, y, B, H);
  }
  VQ_CHECK_CUDA(cudaGetLastError());
  return VQ_OK;
}

extern "C" int vq_geom_head_backward(const float* dy, const float* pooled, const float* pre, const float* act,
                                     const float* w2, float* dw1, float* db1, float* dw2, float* db2, float* scratch,
                                     int32_t B, int32_t in_dim, int32_t H, int32_t accumulate, vq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VQ_CHECK_ARG(dy && pooled && pre && act && w2 && dw1 && db1 && dw2 && db2 && scratch, "null pointer");
  VQ_CHECK_ARG(B > 0 && in_dim > 0 && in_dim <= kGeomMaxIn && H > 0 && H % 4 == 0 && H <= 65535, "geometry");
  {
    LaunchScope scope(kKindRow, "geom_dw2", B, H, H, 0, (double)H * H * (accumulate ? 8 : 4), stream);
    geom_dw2_kernel<<<dim3((H / 4 + 255) / 256, H), 256, 0, stream>>>(dy, act, dw2, db2, B, H, accumulate);
  }
  VQ_CHECK_CUDA(cudaGetLastError());
  VQ_CHECK_CUDA(cudaMemsetAsync(scratch, 0, (size_t)B * H * sizeof(float), stream));
  {
    const int rows_per_slab = 64;
    LaunchScope scope(kKindRow, "geom_dact", B, H, H, 0, (double)H * H * 4, stream);
    geom_dact_kernel<<<dim3((H + 127) / 128, (H + rows_per_slab - 1) / rows_per_slab), 128, 0, stream>>>(
        dy, w2, scratch, B, H, rows_per_slab);
  }
  VQ_CHECK_CUDA(cudaGetLastError());
  {
    LaunchScope scope(kKindRow, "geom_dw1", B, H, in_dim, 0, (double)H * in_dim * 4, stream);
    geom_dw1_kernel<<<(H + 255) / 256, 256, 0, stream>>>(scratch, pre, pooled, dw1, db1, B, in_dim, H, accumulate);
  }
  VQ_CHECK_CUDA(cudaGetLastError());
  return VQ_OK;
}

extern "C" int vq_splice_features(void* inputs_embeds, int32_t dtype, const int64_t* input_ids, const float* geom,
                                  int32_t n_geom, const void* vis, int32_t n_vis, int32_t B, int32_t L, int32_t hidden,
                                  int64_t image_id, int32_t* overflow_count, vq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VQ_CHECK_ARG(inputs_embeds && input_ids && vis, "null pointer");
  VQ_CHECK_ARG(B > 0 && L > 0 && n_vis > 0 && n_geom >= 0 && (n_geom == 0 || geom), "geometry");
  VQ_CHECK_ARG(dtype == VQ_F32 || dtype == VQ_BF16, "dtype");
  const int elt = dtype == VQ_F32 ? 4 : 2;
  VQ_CHECK_ARG((hidden * elt) % 16 == 0, "hidden size must give 16-byte rows");
  VQ_CHECK_ARG(((reinterpret_cast<uintptr_t>(inputs_embeds) | reinterpret_cast<uintptr_t>(vis) |
                 reinterpret_cast<uintptr_t>(geom)) % 16) == 0, "16-byte alignment");
  const long long total = (long long)B * L;
  const unsigned grid = (unsigned)((total + 7) / 8);
  LaunchScope scope(kKindRow, "splice_features", B, L, n_geom + n_vis, 0, 2.0 * B * (n_geom + n_vis) * hidden * elt, stream);
  if (dtype == VQ_F32)
    splice_features_kernel<float><<<grid, 256, 0, stream>>>(
        reinterpret_cast<float*>(inputs_embeds), reinterpret_cast<const long long*>(input_ids), geom,
        reinterpret_cast<const float*>(vis), B, L, hidden, n_geom, n_vis, image_id, overflow_count);
  else
    splice_features_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(
        reinterpret_cast<__nv_bfloat16*>(inputs_embeds), reinterpret_cast<const long long*>(input_ids), geom,
        reinterpret_cast<const __nv_bfloat16*>(vis), B, L, hidden, n_geom, n_vis, image_id, overflow_count);
  VQ_CHECK_CUDA(cudaGetLastError());
  return VQ_OK;
}

extern "C" int vq_splice_features_backward(const void* grad, int32_t dtype, const int64_t* input_ids, void* d_embeds,
                                           float* d_geom, float* d_vis, int32_t n_geom, int32_t n_vis, int32_t B,
                                           int32_t L, int32_t hidden, int64_t image_id, vq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VQ_CHECK_ARG(grad && input_ids, "null pointer");
  VQ_CHECK_ARG(B > 0 && L > 0 && n_vis > 0 && n_geom >= 0, "geometry");
  VQ_CHECK_ARG(dtype == VQ_F32 || dtype == VQ_BF16, "dtype");
  const int elt = dtype == VQ_F32 ? 4 : 2;
  VQ_CHECK_ARG((hidden * elt) % 16 == 0, "hidden size must give 16-byte rows");
  VQ_CHECK_ARG(((reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(d_embeds)) % 16) == 0, "16-byte alignment");
  if (d_geom) VQ_CHECK_CUDA(cudaMemsetAsync(d_geom, 0, (size_t)B * hidden * sizeof(float), stream));
  if (d_vis) VQ_CHECK_CUDA(cudaMemsetAsync(d_vis, 0, (size_t)B * n_vis * hidden * sizeof(float), stream));
  const long long total = (long long)B * L;
  const unsigned grid = (unsigned)((total + 7) / 8);
  LaunchScope scope(kKindRow, "splice_features_bwd", B, L, n_geom + n_vis, 0, 2.0 * total * hidden * elt, stream);
  if (dtype == VQ_F32)
    splice_features_bwd_kernel<float><<<grid, 256, 0, stream>>>(
        reinterpret_cast<const float*>(grad), reinterpret_cast<const long long*>(input_ids),
        reinterpret_cast<float*>(d_embeds), d_geom, d_vis, B, L, hidden, n_geom, n_vis, image_id);
  else
    splice_features_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(grad), reinterpret_cast<const long long*>(input_ids),
        reinterpret_cast<__nv_bfloat16*>(d_embeds), d_geom, d_vis, B, L, hidden, n_geom, n_vis, image_id);
  VQ_CHECK_CUDA(cudaGetLastError());
  return VQ_OK;
}
