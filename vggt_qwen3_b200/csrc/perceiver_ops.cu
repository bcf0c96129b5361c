// HBM-bound kernels of the Perceiver projector forward/backward: row softmax (fwd/bwd), LayerNorm backward,
// column sums (bias / LayerNorm-affine gradients), dtype casts, latent broadcast and batch reduction.
#include "common.cuh"
#include "ptx.cuh"

namespace vq {

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- softmax over the first T columns of each f32 row (scores already scaled); bf16 out, zero padded to Tp.
// One warp per row; T <= 32 * kMaxPerLane.
constexpr int kMaxPerLane = 32;  // T <= 1024

__global__ void __launch_bounds__(256)
softmax_fwd_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ p, long long rows, int T, int lds,
                   int Tp) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* src = s + r * lds;
  float v[kMaxPerLane];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int c = i * 32 + lane;
    v[i] = c < T ? src[c] : -INFINITY;
    mx = fmaxf(mx, v[i]);
  }
  mx = wmax(mx);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    v[i] = (i * 32 < T) ? __expf(v[i] - mx) : 0.f;
    sum += v[i];
  }
  const float inv = 1.0f / wsum(sum);
  __nv_bfloat16* dst = p + r * Tp;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int c = i * 32 + lane;
    if (c < Tp) dst[c] = __float2bfloat16(v[i] * inv);
  }
}

// dS = scale * P .* (dP - rowsum(dP .* P));  P bf16 [rows, Tp], dP f32 [rows, ld], dS bf16 [rows, Tp] (zero padded)
__global__ void __launch_bounds__(256)
softmax_bwd_kernel(const __nv_bfloat16* __restrict__ p, const float* __restrict__ dp,
                   __nv_bfloat16* __restrict__ ds, long long rows, int T, int lddp, int Tp, float scale) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  float pv[kMaxPerLane], dv[kMaxPerLane];
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int c = i * 32 + lane;
    pv[i] = c < T ? __bfloat162float(p[r * Tp + c]) : 0.f;
    dv[i] = c < T ? dp[r * lddp + c] : 0.f;
    dot += pv[i] * dv[i];
  }
  dot = wsum(dot);
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int c = i * 32 + lane;
    if (c < Tp) ds[r * Tp + c] = __float2bfloat16(scale * pv[i] * (dv[i] - dot));
  }
}

// ---- LayerNorm backward.  dz = rstd * (dy*g - mean(dy*g) - xhat * mean(dy*g*xhat)), xhat = (z - mean) * rstd.
// One warp per row for dz; per-CTA column partials of dgamma (sum dy*xhat) and dbeta (sum dy) are reduced in smem
// over the CTA's rows and added to the f32 outputs with one atomic per column per CTA.
template <int VEC>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ z,
                     const float* __restrict__ mean, const float* __restrict__ rstd,
                     const float* __restrict__ gamma, __nv_bfloat16* __restrict__ dz,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, long long rows, int D,
                     int rows_per_cta) {
  extern __shared__ float red[];  // [2][D]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float accg[VEC][4], accb[VEC][4];
#pragma unroll
  for (int i = 0; i < VEC; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) accg[i][j] = accb[i][j] = 0.f;

  const long long row0 = (long long)blockIdx.x * rows_per_cta;
  for (int rr = warp; rr < rows_per_cta; rr += nwarps) {
    const long long r = row0 + rr;
    if (r >= rows) break;
    const float mu = mean[r], rs = rstd[r];
    float dyv[VEC][4], xh[VEC][4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int col = (i * 32 + lane) * 4;
      const uint2 a = *reinterpret_cast<const uint2*>(dy + r * D + col);
      const uint2 b = *reinterpret_cast<const uint2*>(z + r * D + col);
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col));
      dyv[i][0] = bf16_lo(a.x); dyv[i][1] = bf16_hi(a.x); dyv[i][2] = bf16_lo(a.y); dyv[i][3] = bf16_hi(a.y);
      xh[i][0] = (bf16_lo(b.x) - mu) * rs; xh[i][1] = (bf16_hi(b.x) - mu) * rs;
      xh[i][2] = (bf16_lo(b.y) - mu) * rs; xh[i][3] = (bf16_hi(b.y) - mu) * rs;
      const float gg[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        accg[i][j] += dyv[i][j] * xh[i][j];
        accb[i][j] += dyv[i][j];
        dyv[i][j] *= gg[j];  // dy * gamma
        s1 += dyv[i][j];
        s2 += dyv[i][j] * xh[i][j];
      }
    }
    s1 = wsum(s1) / D;
    s2 = wsum(s2) / D;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int col = (i * 32 + lane) * 4;
      uint2 u;
      u.x = pack_bf16x2(rs * (dyv[i][0] - s1 - xh[i][0] * s2), rs * (dyv[i][1] - s1 - xh[i][1] * s2));
      u.y = pack_bf16x2(rs * (dyv[i][2] - s1 - xh[i][2] * s2), rs * (dyv[i][3] - s1 - xh[i][3] * s2));
      *reinterpret_cast<uint2*>(dz + r * D + col) = u;
    }
  }
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int col = (i * 32 + lane) * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      atomicAdd(&red[col + j], accg[i][j]);
      atomicAdd(&red[D + col + j], accb[i][j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    atomicAdd(dgamma + i, red[i]);
    atomicAdd(dbeta + i, red[D + i]);
  }
}

// ---- out[n] += sum_m x[m, n]   (x bf16 or f32 [rows, N], out f32).  grid = (N / 256 column slabs, row chunks)
template <bool F32>
__global__ void __launch_bounds__(256)
colsum_kernel(const void* __restrict__ x, float* __restrict__ out, long long rows, int N, long long ld,
              int rows_per_cta) {
  const int col = blockIdx.x * 256 + threadIdx.x;
  if (col >= N) return;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  const long long r1 = min(rows, r0 + rows_per_cta);
  float acc = 0.f;
  for (long long r = r0; r < r1; ++r) {
    if constexpr (F32) acc += reinterpret_cast<const float*>(x)[r * ld + col];
    else acc += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[r * ld + col]);
  }
  atomicAdd(out + col, acc);
}

// ---- f32 -> bf16 cast, 8 elements per thread
__global__ void __launch_bounds__(256)
cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n8, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n8) {
    const float4 a = reinterpret_cast<const float4*>(in)[2 * i];
    const float4 b = reinterpret_cast<const float4*>(in)[2 * i + 1];
    uint4 u;
    u.x = pack_bf16x2(a.x, a.y); u.y = pack_bf16x2(a.z, a.w);
    u.z = pack_bf16x2(b.x, b.y); u.w = pack_bf16x2(b.z, b.w);
    reinterpret_cast<uint4*>(out)[i] = u;
  }
  if (i == 0)
    for (long long j = n8 * 8; j < n; ++j) out[j] = __float2bfloat16(in[j]);
}

// ---- out[b, :] = src[:]  (latents.unsqueeze(0).expand(B)) ; 16 B per thread
__global__ void __launch_bounds__(256)
broadcast_rows_kernel(const uint4* __restrict__ src, uint4* __restrict__ out, long long n16, int B) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n16) return;
  const uint4 v = src[i];
  for (int b = 0; b < B; ++b) out[(long long)b * n16 + i] = v;
}

// ---- out[i] (+)= sum_b in[b, i]   (bf16 in, f32 out): gradient of the expanded latents
__global__ void __launch_bounds__(256)
reduce_batch_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, long long n, int B,
                    int accumulate) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (i >= n) return;
  float a0 = 0.f, a1 = 0.f;
  for (int b = 0; b < B; ++b) {
    const uint32_t v = *reinterpret_cast<const uint32_t*>(in + (long long)b * n + i);
    a0 += bf16_lo(v);
    a1 += bf16_hi(v);
  }
  if (accumulate) { a0 += out[i]; a1 += out[i + 1]; }
  out[i] = a0;
  out[i + 1] = a1;
}

}  // namespace vq

using namespace vq;

extern "C" int vq_softmax_fwd(const float* scores, void* probs_bf16, int64_t rows, int32_t T, int32_t ld_scores,
                              int32_t Tp, vq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VQ_CHECK_ARG(scores && probs_bf16, "null pointer");
  VQ_CHECK_ARG(rows > 0 && T > 0 && T <= 32 * kMaxPerLane && Tp >= T && Tp <= 32 * kMaxPerLane && ld_scores >= T,
               "softmax geometry (T <= 1024)");
  LaunchScope scope(kKindRow, "softmax_fwd", (int)rows, T, 0, 0, (double)rows * (T * 4.0 + Tp * 2.0), stream);
  softmax_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(
      scores, reinterpret_cast<__nv_bfloat16*>(probs_bf16), rows, T, ld_scores, Tp);
  VQ_CHECK_CUDA(cudaGetLastError());
  return VQ_OK;
}

extern "C" int vq_softmax_bwd(const void* probs_bf16, const float* dprobs, void* dscores_bf16, int64_t rows,
                              int32_t T, int32_t ld_dprobs, int32_t Tp, float scale, vq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VQ_CHECK_ARG(probs_bf16 && dprobs && dscores_bf16, "null pointer");
  VQ_CHECK_ARG(rows > 0 && T > 0 && Tp >= T && Tp <= 32 * kMaxPerLane && ld_dprobs >= T, "softmax geometry");
  LaunchScope scope(kKindRow, "softmax_bwd", (int)rows, T, 0, 0, (double)rows * (T * 6.0 + Tp * 2.0), stream);
  softmax_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(probs_bf16), dprobs, reinterpret_cast<__nv_bfloat16*>(dscores_bf16),
      rows, T, ld_dprobs, Tp, scale);
  VQ_CHECK_CUDA(cudaGetLastError());
  return VQ_OK;
}

extern "C" int vq_layernorm_bwd(const void* dy_bf16, const void* z_bf16, const float* mean, const float* rstd,
                                const float* gamma, void* dz_bf16, float* dgamma, float* dbeta, int64_t rows,
                                int32_t D, vq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VQ_CHECK_ARG(dy_bf16 && z_bf16 && mean && rstd && gamma && dz_bf16 && dgamma && dbeta, "null pointer");
  VQ_CHECK_ARG(rows > 0 && D % 128 == 0, "geometry");
  const int rows_per_cta = 16;
  const unsigned grid = (unsigned)((rows + rows_per_cta - 1) / rows_per_cta);
  const size_t smem = 2 * (size_t)D * sizeof(float);
  const __nv_bfloat16* dy = reinterpret_cast<const __nv_bfloat16*>(dy_bf16);
  const __nv_bfloat16* z = reinterpret_cast<const __nv_bfloat16*>(z_bf16);
  __nv_bfloat16* dz = reinterpret_cast<__nv_bfloat16*>(dz_bf16);
  LaunchScope scope(kKindRow, "layernorm_bwd", (int)rows, D, 0, 0, (double)rows * D * 6.0, stream);
#define VQ_LNB_CASE(V)                                                                                         \
  case V:                                                                                                      \
    layernorm_bwd_kernel<V><<<grid, 256, smem, stream>>>(dy, z, mean, rstd, gamma, dz, dgamma, dbeta, rows, D, \
                                                         rows_per_cta);                                        \
    break;
  switch (D / 128) {
    VQ_LNB_CASE(1) VQ_LNB_CASE(2) VQ_LNB_CASE(4) VQ_LNB_CASE(8) VQ_LNB_CASE(16) VQ_LNB_CASE(32)
    default:
      set_error("layernorm_bwd: unsupported width");
      return VQ_ERR_UNSUPPORTED;
  }
#undef VQ_LNB_CASE
  VQ_CHECK_CUDA(cudaGetLastError());
  return VQ_OK;
}

extern "C" int vq_colsum(const void* x, int32_t dtype, float* out, int64_t rows, int32_t N, int64_t ld,
                         vq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VQ_CHECK_ARG(x && out && rows > 0 && N > 0 && ld >= N, "colsum arguments");
  const int rows_per_cta = 64;
  dim3 grid((N + 255) / 256, (unsigned)((rows + rows_per_cta - 1) / rows_per_cta));
  LaunchScope scope(kKindRow, "colsum", (int)rows, N, 0, 0, (double)rows * N * (dtype == VQ_F32 ? 4.0 : 2.0), stream);
  if (dtype == VQ_F32) colsum_kernel<true><<<grid, 256, 0, stream>>>(x, out, rows, N, ld, rows_per_cta);
  else colsum_kernel<false><<<grid, 256, 0, stream>>>(x, out, rows, N, ld, rows_per_cta);
  VQ_CHECK_CUDA(cudaGetLastError());
  return VQ_OK;
}

extern "C" int vq_cast_f32_bf16(const float* in, void* out_bf16, int64_t n, vq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VQ_CHECK_ARG(in && out_bf16 && n > 0, "cast arguments");
  VQ_CHECK_ARG((reinterpret_cast<uintptr_t>(in) % 16) == 0 && (reinterpret_cast<uintptr_t>(out_bf16) % 16) == 0,
               "cast: 16-byte alignment");
  const long long n8 = n / 8;
  const long long threads = n8 > 0 ? n8 : 1;
  LaunchScope scope(kKindRow, "cast_f32_bf16", 0, 0, 0, 0, (double)n * 6.0, stream);
  cast_f32_bf16_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(
      in, reinterpret_cast<__nv_bfloat16*>(out_bf16), n8, n);
  VQ_CHECK_CUDA(cudaGetLastError());
  return VQ_OK;
}

extern "C" int vq_broadcast_rows(const void* src_bf16, void* out_bf16, int64_t n, int32_t B, vq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VQ_CHECK_ARG(src_bf16 && out_bf16 && n > 0 && n % 8 == 0 && B > 0, "broadcast arguments (n % 8 == 0)");
  const long long n16 = n / 8;
  LaunchScope scope(kKindRow, "broadcast_rows", B, 0, 0, 0, (double)n * 2.0 * (1 + B), stream);
  broadcast_rows_kernel<<<(unsigned)((n16 + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const uint4*>(src_bf16), reinterpret_cast<uint4*>(out_bf16), n16, B);
  VQ_CHECK_CUDA(cudaGetLastError());
  return VQ_OK;
}

extern "C" int vq_reduce_batch(const void* in_bf16, float* out, int64_t n, int32_t B, int32_t accumulate,
                               vq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VQ_CHECK_ARG(in_bf16 && out && n > 0 && n % 2 == 0 && B > 0, "reduce_batch arguments (n even)");
  const long long threads = n / 2;
  LaunchScope scope(kKindRow, "reduce_batch", B, 0, 0, 0, (double)n * (2.0 * B + 4.0), stream);
  reduce_batch_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(in_bf16), out, n, B, accumulate);
  VQ_CHECK_CUDA(cudaGetLastError());
  return VQ_OK;
}
