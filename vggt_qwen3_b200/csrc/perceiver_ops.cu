// HBM-bound kernels of the Perceiver projector forward/backward: row softmax (fwd/bwd), LayerNorm backward,
// column sums (bias / LayerNorm-affine gradients), dtype casts, latent broadcast and batch reduction.
#include "common.cuh"
#include "philox.cuh"
#include "ptx.cuh"

namespace vq {

// keep bit of element (r, c) of a [rows, ld] dropout site
__device__ __forceinline__ bool dropout_keep(const DropoutParams& d, long long r, int c, int ld) {
  const uint64_t group = (uint64_t)r * (uint64_t)(ld >> 3) + (uint64_t)(c >> 3);
  return (dropout_keep8(d, group) >> (c & 7)) & 1u;
}

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
                   int Tp, const DropoutParams drop, __nv_bfloat16* __restrict__ pd) {
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
    if (c < Tp) {
      const __nv_bfloat16 pv = __float2bfloat16(v[i] * inv);
      dst[c] = pv;
      // attention-probability dropout acts on the bf16 probability the P V product would otherwise read
      if (drop.thr)
        pd[r * Tp + c] = (c < T && dropout_keep(drop, r, c, Tp)) ? __float2bfloat16(__bfloat162float(pv) * drop.scale)
                                                                 : __float2bfloat16(0.f);
    }
  }
}

// dS = scale * P .* (dP - rowsum(dP .* P));  P bf16 [rows, Tp], dP f32 [rows, ld], dS bf16 [rows, Tp] (zero padded)
__global__ void __launch_bounds__(256)
softmax_bwd_kernel(const __nv_bfloat16* __restrict__ p, const float* __restrict__ dp,
                   __nv_bfloat16* __restrict__ ds, long long rows, int T, int lddp, int Tp, float scale,
                   const DropoutParams drop) {
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
    if (drop.thr && c < T) dv[i] = dropout_keep(drop, r, c, Tp) ? dv[i] * drop.scale : 0.f;
    dot += pv[i] * dv[i];
  }
  dot = wsum(dot);
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int c = i * 32 + lane;
    if (c < Tp) ds[r * Tp + c] = __float2bfloat16(scale * pv[i] * (dv[i] - dot));
  }
}

// ---- LayerNorm backward, two HBM-bound kernels.
// (1) dz = rstd * (dy*g - mean(dy*g) - xhat * mean(dy*g*xhat)), xhat = (z - mean) * rstd: one warp per row.
template <int VEC, bool DROP>   // DROP: the dropout-branch output is compiled in (keeps the eval kernel small: I-cache)
__global__ void __launch_bounds__(256)
layernorm_bwd_dz_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ z,
                        const float* __restrict__ mean, const float* __restrict__ rstd,
                        const float* __restrict__ gamma, __nv_bfloat16* __restrict__ dz, long long rows, int D,
                        const DropoutParams drop, __nv_bfloat16* __restrict__ dzd) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float mu = mean[r], rs = rstd[r];
  float dyg[VEC][4], xh[VEC][4];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int col = (i * 32 + lane) * 4;
    const uint2 a = *reinterpret_cast<const uint2*>(dy + r * D + col);
    const uint2 b = *reinterpret_cast<const uint2*>(z + r * D + col);
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col));
    dyg[i][0] = bf16_lo(a.x) * g.x; dyg[i][1] = bf16_hi(a.x) * g.y;
    dyg[i][2] = bf16_lo(a.y) * g.z; dyg[i][3] = bf16_hi(a.y) * g.w;
    xh[i][0] = (bf16_lo(b.x) - mu) * rs; xh[i][1] = (bf16_hi(b.x) - mu) * rs;
    xh[i][2] = (bf16_lo(b.y) - mu) * rs; xh[i][3] = (bf16_hi(b.y) - mu) * rs;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s1 += dyg[i][j];
      s2 += dyg[i][j] * xh[i][j];
    }
  }
  s1 = wsum(s1) / D;
  s2 = wsum(s2) / D;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int col = (i * 32 + lane) * 4;
    uint2 u;
    u.x = pack_bf16x2(rs * (dyg[i][0] - s1 - xh[i][0] * s2), rs * (dyg[i][1] - s1 - xh[i][1] * s2));
    u.y = pack_bf16x2(rs * (dyg[i][2] - s1 - xh[i][2] * s2), rs * (dyg[i][3] - s1 - xh[i][3] * s2));
    *reinterpret_cast<uint2*>(dz + r * D + col) = u;
    if (DROP && drop.thr) {
      // gradient of the dropped branch: the mask acts on the bf16 dz the residual path sees
      const uint32_t keep = dropout_keep8(drop, (uint64_t)r * (uint64_t)(D >> 3) + (uint64_t)(col >> 3)) >> (col & 4);
      uint2 w;
      w.x = pack_bf16x2((keep & 1u) ? bf16_lo(u.x) * drop.scale : 0.f, (keep & 2u) ? bf16_hi(u.x) * drop.scale : 0.f);
      w.y = pack_bf16x2((keep & 4u) ? bf16_lo(u.y) * drop.scale : 0.f, (keep & 8u) ? bf16_hi(u.y) * drop.scale : 0.f);
      *reinterpret_cast<uint2*>(dzd + r * D + col) = w;
    }
  }
}

// Wide rows (D = 1024 * G8, G8 in {1, 2, 4, 8}): one CTA of 128 threads per row instead of one warp -- 8 x more threads in
// flight for the B*Lq = 1024 rows of a c2 step, 128-bit loads, 8 * G8 elements per thread (one Philox call per 8-group).
template <int G8, bool DROP>
__global__ void __launch_bounds__(128)
layernorm_bwd_dz_row_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ z,
                            const float* __restrict__ mean, const float* __restrict__ rstd,
                            const float* __restrict__ gamma, __nv_bfloat16* __restrict__ dz, int D,
                            const DropoutParams drop, __nv_bfloat16* __restrict__ dzd) {
  __shared__ float red[2][4];
  const long long r = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float mu = mean[r], rs = rstd[r];
  float dyg[G8][8], xh[G8][8];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < G8; ++i) {
    const int col = (i * 128 + threadIdx.x) * 8;
    const uint4 a = *reinterpret_cast<const uint4*>(dy + r * D + col);
    const uint4 b = *reinterpret_cast<const uint4*>(z + r * D + col);
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + col));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + col + 4));
    const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      dyg[i][2 * j] = bf16_lo(aw[j]) * gv[2 * j];
      dyg[i][2 * j + 1] = bf16_hi(aw[j]) * gv[2 * j + 1];
      xh[i][2 * j] = (bf16_lo(bw[j]) - mu) * rs;
      xh[i][2 * j + 1] = (bf16_hi(bw[j]) - mu) * rs;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s1 += dyg[i][j];
      s2 += dyg[i][j] * xh[i][j];
    }
  }
  s1 = wsum(s1);
  s2 = wsum(s2);
  if (lane == 0) { red[0][warp] = s1; red[1][warp] = s2; }
  __syncthreads();
  s1 = (red[0][0] + red[0][1] + red[0][2] + red[0][3]) / D;
  s2 = (red[1][0] + red[1][1] + red[1][2] + red[1][3]) / D;
#pragma unroll
  for (int i = 0; i < G8; ++i) {
    const int col = (i * 128 + threadIdx.x) * 8;
    uint32_t u[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      u[j] = pack_bf16x2(rs * (dyg[i][2 * j] - s1 - xh[i][2 * j] * s2), rs * (dyg[i][2 * j + 1] - s1 - xh[i][2 * j + 1] * s2));
    *reinterpret_cast<uint4*>(dz + r * D + col) = make_uint4(u[0], u[1], u[2], u[3]);
    if (DROP && drop.thr) {
      // gradient of the dropped branch: the mask acts on the bf16 dz the residual path sees
      const uint32_t keep = dropout_keep8(drop, (uint64_t)r * (uint64_t)(D >> 3) + (uint64_t)(col >> 3));
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        w[j] = pack_bf16x2((keep >> (2 * j)) & 1u ? bf16_lo(u[j]) * drop.scale : 0.f,
                           (keep >> (2 * j + 1)) & 1u ? bf16_hi(u[j]) * drop.scale : 0.f);
      *reinterpret_cast<uint4*>(dzd + r * D + col) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

// test hook: materialise a site's mask
__global__ void __launch_bounds__(256)
dropout_mask_kernel(uint8_t* __restrict__ keep, long long rows, int cols, int ld, const DropoutParams drop) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const long long r = i / cols;
  const int c = (int)(i - r * cols);
  keep[i] = drop.thr ? (dropout_keep(drop, r, c, ld) ? 1 : 0) : 1;
}

// Column reductions over a chunk of rows.  CTA = 32 column groups (8 adjacent columns each: 128-bit loads, 512
// contiguous bytes per row across the groups) x 8 row lanes; the 8 row lanes of a column meet in shared memory, so a CTA
// issues ONE f32 atomic per column (the earlier one-thread-per-column-group version with 4-row chunks put 256 atomics
// on every output address).  grid = (column slabs of 32 * kPer, row chunks).
constexpr int kColGroups = 32, kRowLanes = 8;

template <int NACC, int kPer>
__device__ __forceinline__ void colreduce_finish(float (&acc)[NACC][kPer], float* __restrict__ const (&out)[NACC], int col0,
                                                 int N) {
  __shared__ float red[NACC][kRowLanes][kColGroups * kPer + 1];
  const int cg = threadIdx.x % kColGroups, rl = threadIdx.x / kColGroups;
#pragma unroll
  for (int a = 0; a < NACC; ++a)
#pragma unroll
    for (int j = 0; j < kPer; ++j) red[a][rl][cg * kPer + j] = acc[a][j];
  __syncthreads();
  for (int c = threadIdx.x; c < kColGroups * kPer; c += kColGroups * kRowLanes) {
    if (col0 + c >= N) continue;
#pragma unroll
    for (int a = 0; a < NACC; ++a) {
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < kRowLanes; ++r) t += red[a][r][c];
      atomicAdd(out[a] + col0 + c, t);
    }
  }
}

// (2) dgamma[n] += sum_m dy*xhat, dbeta[n] += sum_m dy
__global__ void __launch_bounds__(kColGroups * kRowLanes)
layernorm_bwd_affine_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ z,
                            const float* __restrict__ mean, const float* __restrict__ rstd,
                            float* __restrict__ dgamma, float* __restrict__ dbeta, long long rows, int D,
                            int rows_per_cta) {
  const int cg = threadIdx.x % kColGroups, rl = threadIdx.x / kColGroups;
  const int col0 = blockIdx.x * (kColGroups * 8);
  const int col = col0 + cg * 8;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  const long long r1 = min(rows, r0 + rows_per_cta);
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = 0.f;
  if (col < D) {
#pragma unroll 4
    for (long long r = r0 + rl; r < r1; r += kRowLanes) {
      const uint4 a = *reinterpret_cast<const uint4*>(dy + r * D + col);
      const uint4 b = *reinterpret_cast<const uint4*>(z + r * D + col);
      const float mu = __ldg(mean + r), rs = __ldg(rstd + r);
      const float d[8] = {bf16_lo(a.x), bf16_hi(a.x), bf16_lo(a.y), bf16_hi(a.y),
                          bf16_lo(a.z), bf16_hi(a.z), bf16_lo(a.w), bf16_hi(a.w)};
      const float x[8] = {bf16_lo(b.x), bf16_hi(b.x), bf16_lo(b.y), bf16_hi(b.y),
                          bf16_lo(b.z), bf16_hi(b.z), bf16_lo(b.w), bf16_hi(b.w)};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[0][j] += d[j] * (x[j] - mu) * rs;
        acc[1][j] += d[j];
      }
    }
  }
  float* const outs[2] = {dgamma, dbeta};
  colreduce_finish<2, 8>(acc, outs, col0, D);
}

// ---- out[n] += sum_m x[m, n]   (x bf16 or f32 [rows, N], out f32): 8 (bf16) / 4 (f32) columns per thread
template <bool F32>
__global__ void __launch_bounds__(kColGroups * kRowLanes)
colsum_kernel(const void* __restrict__ x, float* __restrict__ out, long long rows, int N, long long ld,
              int rows_per_cta) {
  constexpr int kPer = F32 ? 4 : 8;
  const int cg = threadIdx.x % kColGroups, rl = threadIdx.x / kColGroups;
  const int col0 = blockIdx.x * (kColGroups * kPer);
  const int col = col0 + cg * kPer;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  const long long r1 = min(rows, r0 + rows_per_cta);
  float acc[1][kPer];
#pragma unroll
  for (int j = 0; j < kPer; ++j) acc[0][j] = 0.f;
  if (col < N) {
#pragma unroll 4
    for (long long r = r0 + rl; r < r1; r += kRowLanes) {
      if constexpr (F32) {
        const float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + r * ld + col);
        acc[0][0] += v.x; acc[0][1] += v.y; acc[0][2] += v.z; acc[0][3] += v.w;
      } else {
        const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(x) + r * ld + col);
        acc[0][0] += bf16_lo(v.x); acc[0][1] += bf16_hi(v.x); acc[0][2] += bf16_lo(v.y); acc[0][3] += bf16_hi(v.y);
        acc[0][4] += bf16_lo(v.z); acc[0][5] += bf16_hi(v.z); acc[0][6] += bf16_lo(v.w); acc[0][7] += bf16_hi(v.w);
      }
    }
  }
  float* const outs[1] = {out};
  colreduce_finish<1, kPer>(acc, outs, col0, N);
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

static DropoutParams params_of(const vq_dropout_desc* d) {
  if (!d) return make_dropout_params(0.f, 0, 0, 0);
  return make_dropout_params(d->p, d->seed, d->offset, d->site);
}
#define VQ_CHECK_DROPOUT(d) VQ_CHECK_ARG(!(d) || ((d)->p >= 0.f && (d)->p < 1.f), "dropout p must be in [0, 1)")

extern "C" int vq_dropout_mask(uint8_t* keep, int64_t rows, int32_t cols, int32_t ld, const vq_dropout_desc* dropout,
                               vq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VQ_CHECK_ARG(keep && rows > 0 && cols > 0 && ld >= cols && ld % 8 == 0, "dropout_mask arguments (ld % 8 == 0)");
  VQ_CHECK_DROPOUT(dropout);
  const long long n = (long long)rows * cols;
  LaunchScope scope(kKindRow, "dropout_mask", (int)rows, cols, 0, 0, (double)n, stream);
  dropout_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(keep, rows, cols, ld, params_of(dropout));
  VQ_CHECK_CUDA(cudaGetLastError());
  return VQ_OK;
}

extern "C" int vq_softmax_fwd(const float* scores, void* probs_bf16, int64_t rows, int32_t T, int32_t ld_scores,
                              int32_t Tp, const vq_dropout_desc* dropout, void* probs_dropped_bf16,
                              vq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VQ_CHECK_ARG(scores && probs_bf16, "null pointer");
  VQ_CHECK_DROPOUT(dropout);
  const DropoutParams drop = params_of(dropout);
  VQ_CHECK_ARG(!drop.thr || (probs_dropped_bf16 && Tp % 8 == 0), "dropout needs probs_dropped and Tp % 8 == 0");
  VQ_CHECK_ARG(rows > 0 && T > 0 && T <= 32 * kMaxPerLane && Tp >= T && Tp <= 32 * kMaxPerLane && ld_scores >= T,
               "softmax geometry (T <= 1024)");
  LaunchScope scope(kKindRow, "softmax_fwd", (int)rows, T, 0, 0, (double)rows * (T * 4.0 + Tp * 2.0), stream);
  softmax_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(
      scores, reinterpret_cast<__nv_bfloat16*>(probs_bf16), rows, T, ld_scores, Tp, drop,
      reinterpret_cast<__nv_bfloat16*>(probs_dropped_bf16));
  VQ_CHECK_CUDA(cudaGetLastError());
  return VQ_OK;
}

extern "C" int vq_softmax_bwd(const void* probs_bf16, const float* dprobs, void* dscores_bf16, int64_t rows,
                              int32_t T, int32_t ld_dprobs, int32_t Tp, float scale, const vq_dropout_desc* dropout,
                              vq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VQ_CHECK_ARG(probs_bf16 && dprobs && dscores_bf16, "null pointer");
  VQ_CHECK_DROPOUT(dropout);
  const DropoutParams drop = params_of(dropout);
  VQ_CHECK_ARG(!drop.thr || Tp % 8 == 0, "dropout needs Tp % 8 == 0");
  VQ_CHECK_ARG(rows > 0 && T > 0 && Tp >= T && Tp <= 32 * kMaxPerLane && ld_dprobs >= T, "softmax geometry");
  LaunchScope scope(kKindRow, "softmax_bwd", (int)rows, T, 0, 0, (double)rows * (T * 6.0 + Tp * 2.0), stream);
  softmax_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(probs_bf16), dprobs, reinterpret_cast<__nv_bfloat16*>(dscores_bf16),
      rows, T, ld_dprobs, Tp, scale, drop);
  VQ_CHECK_CUDA(cudaGetLastError());
  return VQ_OK;
}

extern "C" int vq_layernorm_bwd(const void* dy_bf16, const void* z_bf16, const float* mean, const float* rstd,
                                const float* gamma, void* dz_bf16, float* dgamma, float* dbeta, int64_t rows,
                                int32_t D, const vq_dropout_desc* dropout, void* dz_dropped_bf16,
                                vq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VQ_CHECK_ARG(dy_bf16 && z_bf16 && mean && rstd && gamma && dz_bf16 && dgamma && dbeta, "null pointer");
  VQ_CHECK_DROPOUT(dropout);
  const DropoutParams drop = params_of(dropout);
  VQ_CHECK_ARG(!drop.thr || dz_dropped_bf16, "dropout needs dz_dropped");
  __nv_bfloat16* dzd = reinterpret_cast<__nv_bfloat16*>(dz_dropped_bf16);
  VQ_CHECK_ARG(rows > 0 && D % 128 == 0, "geometry");
  const __nv_bfloat16* dy = reinterpret_cast<const __nv_bfloat16*>(dy_bf16);
  const __nv_bfloat16* z = reinterpret_cast<const __nv_bfloat16*>(z_bf16);
  __nv_bfloat16* dz = reinterpret_cast<__nv_bfloat16*>(dz_bf16);
  {
    const unsigned grid = (unsigned)((rows + 7) / 8);
    LaunchScope scope(kKindRow, "layernorm_bwd_dz", (int)rows, D, 0, 0, (double)rows * D * 6.0, stream);
    const int g8 = D % 1024 == 0 ? D / 1024 : 0;
    if (g8 == 1 || g8 == 2 || g8 == 4 || g8 == 8) {
#define VQ_LNB_ROW(G)                                                                                                  \
  case G:                                                                                                              \
    if (drop.thr) layernorm_bwd_dz_row_kernel<G, true><<<(unsigned)rows, 128, 0, stream>>>(dy, z, mean, rstd, gamma, dz, D, drop, dzd); \
    else layernorm_bwd_dz_row_kernel<G, false><<<(unsigned)rows, 128, 0, stream>>>(dy, z, mean, rstd, gamma, dz, D, drop, dzd); \
    break;
      switch (g8) { VQ_LNB_ROW(1) VQ_LNB_ROW(2) VQ_LNB_ROW(4) VQ_LNB_ROW(8) }
#undef VQ_LNB_ROW
    } else {
#define VQ_LNB_CASE(V)                                                                                  \
  case V:                                                                                               \
    if (drop.thr) layernorm_bwd_dz_kernel<V, true><<<grid, 256, 0, stream>>>(dy, z, mean, rstd, gamma, dz, rows, D, drop, dzd); \
    else layernorm_bwd_dz_kernel<V, false><<<grid, 256, 0, stream>>>(dy, z, mean, rstd, gamma, dz, rows, D, drop, dzd); \
    break;
    switch (D / 128) {
      VQ_LNB_CASE(1) VQ_LNB_CASE(2) VQ_LNB_CASE(4) VQ_LNB_CASE(8) VQ_LNB_CASE(16) VQ_LNB_CASE(32)
      default:
        set_error("layernorm_bwd: unsupported width");
        return VQ_ERR_UNSUPPORTED;
    }
#undef VQ_LNB_CASE
    }
  }
  VQ_CHECK_CUDA(cudaGetLastError());
  {
    // enough CTAs to fill the machine: with 1024 rows x 4096 columns 32 rows per CTA gave 128 CTAs of 128 threads
    const int col_ctas = (D + kColGroups * 8 - 1) / (kColGroups * 8);
    int rows_per_cta = 64;   // 8 rows per thread; fewer while the grid would not fill the machine twice over
    while (rows_per_cta > 16 && (long long)col_ctas * ((rows + rows_per_cta - 1) / rows_per_cta) < 2LL * 148) rows_per_cta /= 2;
    dim3 grid(col_ctas, (unsigned)((rows + rows_per_cta - 1) / rows_per_cta));
    LaunchScope scope(kKindRow, "layernorm_bwd_affine", (int)rows, D, 0, 0, (double)rows * D * 4.0, stream);
    layernorm_bwd_affine_kernel<<<grid, kColGroups * kRowLanes, 0, stream>>>(dy, z, mean, rstd, dgamma, dbeta, rows, D,
                                                                            rows_per_cta);
  }
  VQ_CHECK_CUDA(cudaGetLastError());
  return VQ_OK;
}

extern "C" int vq_colsum(const void* x, int32_t dtype, float* out, int64_t rows, int32_t N, int64_t ld,
                         vq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VQ_CHECK_ARG(x && out && rows > 0 && N > 0 && ld >= N, "colsum arguments");
  VQ_CHECK_ARG(N % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) % 16) == 0, "colsum: 16-byte rows");
  const int per = dtype == VQ_F32 ? 4 : 8;
  const int col_ctas = (N + kColGroups * per - 1) / (kColGroups * per);
  int rows_per_cta = 64;   // 8 rows per thread; fewer while the grid would not fill the machine twice over
  while (rows_per_cta > 16 && (long long)col_ctas * ((rows + rows_per_cta - 1) / rows_per_cta) < 2LL * 148) rows_per_cta /= 2;
  dim3 grid(col_ctas, (unsigned)((rows + rows_per_cta - 1) / rows_per_cta));
  LaunchScope scope(kKindRow, "colsum", (int)rows, N, 0, 0, (double)rows * N * (dtype == VQ_F32 ? 4.0 : 2.0), stream);
  if (dtype == VQ_F32) colsum_kernel<true><<<grid, kColGroups * kRowLanes, 0, stream>>>(x, out, rows, N, ld, rows_per_cta);
  else colsum_kernel<false><<<grid, kColGroups * kRowLanes, 0, stream>>>(x, out, rows, N, ld, rows_per_cta);
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
