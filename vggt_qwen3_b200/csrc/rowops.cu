// HBM-bound row kernels of the aggregator path: normalise+im2col, token assembly, LayerNorm,
// DINOv2 final-norm + special-token swap, last-layer token gather, inputs_embeds splice.
// All use 128-bit (or widest aligned) accesses, one warp per row with shuffle reductions.
#include "common.cuh"
#include "ptx.cuh"

namespace vq {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_max_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------------------------------------
// K1 + K2 front: (x - mean) / std, then im2col of 14x14/14 patches into bf16 rows [c*ps*ps + dy*ps + dx].
// One CTA per (frame, patch row): the 3 x ps x W strip is read coalesced, staged in smem as normalised bf16
// ([c][dy][x]), and written out as whole patch rows with 16-byte stores (8 consecutive k per thread).
// SRC_U8 = false: f32 images [frames, 3, H, W] in [0, 1];  SRC_U8 = true: uint8 views [frames, H, W, 3] straight from the
// decoder / H2D copy -- ToTensor's x / 255 is applied here, bit-identical to torchvision (IEEE division), so the f32
// image is never materialised (4x less input traffic: 3.22 MB -> 0.80 MB per frame).
// ------------------------------------------------------------------------------------------------
template <bool SRC_U8>
__global__ void __launch_bounds__(256)
im2col_kernel(const void* __restrict__ src, __nv_bfloat16* __restrict__ out, int H, int W, int ps, int gh, int gw,
              int Kp) {
  extern __shared__ __nv_bfloat16 strip[];  // [3][ps][Wc], Wc = gw*ps
  const int f = blockIdx.x / gh, gy = blockIdx.x % gh;
  const int Wc = gw * ps;
  const float mean[3] = {0.485f, 0.456f, 0.406f};
  const float istd[3] = {1.0f / 0.229f, 1.0f / 0.224f, 1.0f / 0.225f};
  if constexpr (SRC_U8) {
    // the ps rows of the strip are one contiguous run of ps * W * 3 bytes in HWC order: read it as 32-bit words where
    // the alignment allows, scatter the bytes into the [c][dy][x] layout
    const unsigned char* base = reinterpret_cast<const unsigned char*>(src) + ((size_t)f * H + (size_t)gy * ps) * W * 3;
    const int nbytes = ps * W * 3;
    const int lead = (int)((4 - (reinterpret_cast<uintptr_t>(base) & 3)) & 3);   // bytes before the first aligned word
    auto put = [&](int byte_index, unsigned v) {
      const int pix = byte_index / 3, c = byte_index - pix * 3;
      const int dy = pix / W, x = pix - dy * W;
      if (x < Wc) strip[(c * ps + dy) * Wc + x] = __float2bfloat16(((float)v / 255.0f - mean[c]) * istd[c]);
    };
    for (int i = threadIdx.x; i < lead && i < nbytes; i += blockDim.x) put(i, base[i]);
    const int nwords = (nbytes - lead) / 4;
    const unsigned* words = reinterpret_cast<const unsigned*>(base + lead);
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) {
      const unsigned w = __ldg(words + i);
      const int b0 = lead + 4 * i;
      put(b0, w & 0xFFu); put(b0 + 1, (w >> 8) & 0xFFu); put(b0 + 2, (w >> 16) & 0xFFu); put(b0 + 3, w >> 24);
    }
    for (int i = lead + 4 * nwords + threadIdx.x; i < nbytes; i += blockDim.x) put(i, base[i]);
  } else {
    const float* img = reinterpret_cast<const float*>(src);
    const int n = 3 * ps * Wc;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int x = i % Wc;
      const int dy = (i / Wc) % ps;
      const int c = i / (Wc * ps);
      const float v = img[(((long long)f * 3 + c) * H + gy * ps + dy) * W + x];
      strip[i] = __float2bfloat16((v - mean[c]) * istd[c]);
    }
  }
  __syncthreads();
  const int K = 3 * ps * ps;
  const int k8 = Kp / 8;                       // 16-byte granules per patch row
  const int total = gw * k8;
  __nv_bfloat16* dst = out + ((long long)f * gh + gy) * gw * Kp;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int gx = i / k8, k0 = (i - gx * k8) * 8;
    __align__(16) __nv_bfloat16 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + j;
      __nv_bfloat16 e = __float2bfloat16(0.f);
      if (k < K) {
        const int c = k / (ps * ps);
        const int r = k - c * ps * ps;
        const int dy = r / ps, dx = r - dy * ps;
        e = strip[(c * ps + dy) * Wc + gx * ps + dx];
      }
      v[j] = e;
    }
    *reinterpret_cast<uint4*>(dst + (long long)gx * Kp + k0) = *reinterpret_cast<const uint4*>(v);
  }
}

// uint8 views with a compile-time patch size (14 for DINOv2 / VGGT): the byte strip of one patch row goes to smem RAW
// with 16-byte loads (no per-byte index arithmetic), ToTensor + normalisation is a 3 x 256-entry table (the same IEEE
// expression as above, evaluated once per CTA: bit-identical), and every thread assembles one 16-byte granule of a patch
// row from (byte, table) lookups with constant divisors.
template <int PS>
__global__ void __launch_bounds__(256)
im2col_u8_kernel(const unsigned char* __restrict__ src, __nv_bfloat16* __restrict__ out, int H, int W, int gh, int gw,
                 int Kp) {
  extern __shared__ __align__(16) unsigned char raw[];   // [16 slack][PS * W * 3 rounded up to 16][3 * 256 bf16]
  const int f = blockIdx.x / gh, gy = blockIdx.x % gh;
  const int nbytes = PS * W * 3;
  const unsigned char* base = src + ((size_t)f * H + (size_t)gy * PS) * W * 3;
  const int mis = (int)(reinterpret_cast<uintptr_t>(base) & 15);     // the strip starts `mis` bytes into a 16-byte line
  const uint4* lines = reinterpret_cast<const uint4*>(base - mis);
  const int nlines = (mis + nbytes + 15) >> 4;
  __nv_bfloat16* lut = reinterpret_cast<__nv_bfloat16*>(raw + ((PS * W * 3 + 31 + 15) & ~15));
  // whole 16-byte lines: the first / last one may reach up to 15 bytes outside the strip -- still inside the frame
  // buffer except at its very ends, where those lines are read bytewise
  const bool head_ok = mis == 0 || f > 0 || gy > 0;
  const bool tail_ok = ((mis + nbytes) & 15) == 0 || f + 1 < (int)(gridDim.x / gh) || gy + 1 < gh || H > gh * PS;
  for (int i = threadIdx.x; i < nlines; i += blockDim.x) {
    if ((i == 0 && !head_ok) || (i == nlines - 1 && !tail_ok)) {
      for (int b = 0; b < 16; ++b) {
        const int idx = 16 * i + b - mis;
        if (idx >= 0 && idx < nbytes) raw[16 * i + b] = base[idx];
      }
    } else {
      *reinterpret_cast<uint4*>(raw + 16 * i) = __ldg(lines + i);
    }
  }
  for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) {
    const int c = i >> 8;
    const float mean = c == 0 ? 0.485f : c == 1 ? 0.456f : 0.406f;
    const float istd = c == 0 ? 1.0f / 0.229f : c == 1 ? 1.0f / 0.224f : 1.0f / 0.225f;
    lut[i] = __float2bfloat16(((float)(i & 255) / 255.0f - mean) * istd);
  }
  __syncthreads();
  constexpr int K = 3 * PS * PS;
  const int k8 = Kp / 8;
  const int total = gw * k8;
  const unsigned char* px = raw + mis;
  __nv_bfloat16* dst = out + ((long long)f * gh + gy) * gw * Kp;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int gx = i / k8, k0 = (i - gx * k8) * 8;
    __align__(16) __nv_bfloat16 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + j;
      __nv_bfloat16 e = __float2bfloat16(0.f);
      if (k < K) {
        const int c = k / (PS * PS);
        const int r = k - c * (PS * PS);
        const int dy = r / PS, dx = r - dy * PS;
        e = lut[c * 256 + px[(dy * W + gx * PS + dx) * 3 + c]];
      }
      v[j] = e;
    }
    *reinterpret_cast<uint4*>(dst + (long long)gx * Kp + k0) = *reinterpret_cast<const uint4*>(v);
  }
}

// K3: rows 0..n_special-1 of every frame: cls + pos_embed[0], then the register tokens (f32 stream).
__global__ void init_special_kernel(float* __restrict__ x, const float* __restrict__ cls,
                                    const float* __restrict__ pos0, const float* __restrict__ regs,
                                    int P, int D, int n_special) {
  const int f = blockIdx.x;
  for (int i = threadIdx.x; i < n_special * D; i += blockDim.x) {
    const int r = i / D, c = i - r * D;
    const float v = r == 0 ? cls[c] + pos0[c] : regs[(r - 1) * D + c];
    x[((long long)f * P + r) * D + c] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over the last dimension, one warp per row, row held in registers (two-pass variance).
// IN_F32 / OUT_F32 select the stream dtypes; VEC = D / 128 (each lane owns VEC float4 groups).
// ------------------------------------------------------------------------------------------------
template <int VEC, bool IN_F32>
__device__ __forceinline__ void load_row(const void* in, long long row, int D, int lane, float (&v)[VEC][4]) {
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int col = (i * 32 + lane) * 4;
    if constexpr (IN_F32) {
      const float4 t = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(in) + row * D + col);
      v[i][0] = t.x; v[i][1] = t.y; v[i][2] = t.z; v[i][3] = t.w;
    } else {
      const uint2 t = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(in) + row * D + col);
      v[i][0] = bf16_lo(t.x); v[i][1] = bf16_hi(t.x); v[i][2] = bf16_lo(t.y); v[i][3] = bf16_hi(t.y);
    }
  }
}
template <int VEC>
__device__ __forceinline__ void row_stats(const float (&v)[VEC][4], int D, float eps, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) s += (v[i][0] + v[i][1]) + (v[i][2] + v[i][3]);
  mean = warp_sum(s) / D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float d = v[i][j] - mean;
      q += d * d;
    }
  }
  rstd = rsqrtf(warp_sum(q) / D + eps);
}
template <int VEC, bool OUT_F32>
__device__ __forceinline__ void norm_store_row(void* out, long long row, int D, int lane, const float (&v)[VEC][4],
                                               float mean, float rstd, const float* gamma, const float* beta) {
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int col = (i * 32 + lane) * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col));
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta + col));
    const float y0 = (v[i][0] - mean) * rstd * g.x + b.x;
    const float y1 = (v[i][1] - mean) * rstd * g.y + b.y;
    const float y2 = (v[i][2] - mean) * rstd * g.z + b.z;
    const float y3 = (v[i][3] - mean) * rstd * g.w + b.w;
    if constexpr (OUT_F32) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + row * D + col) = make_float4(y0, y1, y2, y3);
    } else {
      uint2 u;
      u.x = pack_bf16x2(y0, y1);
      u.y = pack_bf16x2(y2, y3);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + row * D + col) = u;
    }
  }
}

template <int VEC, bool IN_F32, bool OUT_F32>
__global__ void __launch_bounds__(256)
layernorm_kernel(const void* __restrict__ in, void* __restrict__ out, const float* __restrict__ gamma,
                 const float* __restrict__ beta, long long rows, int D, float eps,
                 float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  pdl_launch_dependents();
  pdl_wait();   // `in` is the previous kernel's output
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float v[VEC][4];
  load_row<VEC, IN_F32>(in, row, D, lane, v);
  float mean, rstd;
  row_stats<VEC>(v, D, eps, mean, rstd);
  norm_store_row<VEC, OUT_F32>(out, row, D, lane, v, mean, rstd, gamma, beta);
  if (mean_out && lane == 0) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
}

// Residual add fused with the LayerNorm that follows it: x += y (y = the bf16 branch output gamma_ls (acc + b) of the proj /
// fc2 GEMM), then xn = LN(x) for the next GEMM.  The GEMM epilogue then is a plain bf16 store (no f32 read-modify-write
// stalling its accumulator drain) and the f32 stream is touched by ONE streaming pass per half block:
// 4 B read + 2 B read + 4 B write + 2 B write per element.  xn == nullptr: add only (end of a block chain).
template <int VEC>
__global__ void __launch_bounds__(256)
residual_layernorm_kernel(float* __restrict__ x, const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ xn,
                          const float* __restrict__ gamma, const float* __restrict__ beta, long long rows, int D,
                          float eps) {
  pdl_launch_dependents();
  pdl_wait();   // x and y come from the previous kernels
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float v[VEC][4];
  load_row<VEC, true>(x, row, D, lane, v);
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int col = (i * 32 + lane) * 4;
    const uint2 t = *reinterpret_cast<const uint2*>(y + row * D + col);
    v[i][0] += bf16_lo(t.x); v[i][1] += bf16_hi(t.x); v[i][2] += bf16_lo(t.y); v[i][3] += bf16_hi(t.y);
    *reinterpret_cast<float4*>(x + row * D + col) = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
  }
  if (xn == nullptr) return;
  float mean, rstd;
  row_stats<VEC>(v, D, eps, mean, rstd);
  norm_store_row<VEC, false>(xn, row, D, lane, v, mean, rstd, gamma, beta);
}

// Entry of a LayerNorm-folded chain: per-row (sum, sum of squares) + bf16 copy, one warp per row.
template <int VEC>
__global__ void __launch_bounds__(256)
row_stats_cast_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xb, float* __restrict__ stats,
                      float* __restrict__ stats_zero, long long rows, int D) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float v[VEC][4];
  load_row<VEC, true>(x, row, D, lane, v);
  float s = 0.f, q = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    s += (v[i][0] + v[i][1]) + (v[i][2] + v[i][3]);
    q += (v[i][0] * v[i][0] + v[i][1] * v[i][1]) + (v[i][2] * v[i][2] + v[i][3] * v[i][3]);
    uint2 u;
    u.x = pack_bf16x2(v[i][0], v[i][1]);
    u.y = pack_bf16x2(v[i][2], v[i][3]);
    *reinterpret_cast<uint2*>(xb + row * D + (i * 32 + lane) * 4) = u;
  }
  s = warp_sum(s);
  q = warp_sum(q);
  if (lane == 0) {
    stats[2 * row] = s;
    stats[2 * row + 1] = q;
    if (stats_zero) {
      stats_zero[2 * row] = 0.f;
      stats_zero[2 * row + 1] = 0.f;
    }
  }
}

// K13: DINOv2 final LayerNorm on the patch rows (in place, f32) + camera/register tokens swapped into
// the special rows (token index 0 for the first frame of a scene, 1 for the others).
template <int VEC>
__global__ void __launch_bounds__(256)
stem_finalize_kernel(float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     const float* __restrict__ camera, const float* __restrict__ regs, long long rows, int D,
                     int P, int S, int n_special, float eps) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const long long frame = row / P;
  const int p = (int)(row - frame * P);
  if (p < n_special) {
    const int which = (frame % S) == 0 ? 0 : 1;
    // camera_token [1,2,1,D], register_token [1,2,R,D]
    const float* src = p == 0 ? camera + (long long)which * D
                              : regs + ((long long)which * (n_special - 1) + (p - 1)) * D;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int col = (i * 32 + lane) * 4;
      *reinterpret_cast<float4*>(x + row * D + col) = __ldg(reinterpret_cast<const float4*>(src + col));
    }
    return;
  }
  float v[VEC][4];
  load_row<VEC, true>(x, row, D, lane, v);
  float mean, rstd;
  row_stats<VEC>(v, D, eps, mean, rstd);
  norm_store_row<VEC, true>(x, row, D, lane, v, mean, rstd, gamma, beta);
}

// K14 + K15: out[b, t, col_off : col_off + D] = bf16(x[b * tokens_per_scene + t, :]) for t < n_take.
__global__ void __launch_bounds__(256)
gather_tokens_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int D, long long tokens_per_scene,
                     int n_take, int out_ld, int col_off, long long total_rows) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= total_rows) return;
  const long long b = r / n_take;
  const int t = (int)(r - b * n_take);
  const float* src = x + (b * tokens_per_scene + t) * D;
  __nv_bfloat16* dst = out + r * out_ld + col_off;
  for (int c = lane * 4; c < D; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(src + c);
    uint2 u;
    u.x = pack_bf16x2(v.x, v.y);
    u.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(dst + c) = u;
  }
}

// K22: training splice.  For every row r of sample b the last <image> position pos <= r with r - pos < n
// wins (the reference loops over positions in ascending order, later spans overwrite earlier ones).
// One warp per (b, r); no host synchronisation.  ELT = bytes per element of embeds/features.
template <typename T>
__global__ void __launch_bounds__(256)
splice_overwrite_kernel(T* __restrict__ embeds, const long long* __restrict__ ids, const T* __restrict__ feats,
                        int B, int L, int Hd, int n, long long image_id, int* __restrict__ overflow) {
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (long long)B * L) return;
  const int b = (int)(w / L), r = (int)(w - (long long)b * L);
  const long long* row_ids = ids + (long long)b * L;
  int best = -1;
  for (int p = r - lane; p > r - n && p >= 0; p -= 32)
    if (row_ids[p] == image_id) { best = p; break; }  // first hit per lane is that lane's largest
  best = warp_max_i(best);
  if (overflow && lane == 0 && row_ids[r] == image_id && r + n > L) atomicAdd(overflow, 1);
  if (best < 0) return;
  const T* src = feats + ((long long)b * n + (r - best)) * Hd;
  T* dst = embeds + ((long long)b * L + r) * Hd;
  constexpr int kVec = 16 / sizeof(T);
  for (int c = lane * kVec; c < Hd; c += 32 * kVec)
    *reinterpret_cast<uint4*>(dst + c) = *reinterpret_cast<const uint4*>(src + c);
}

}  // namespace vq

// ==================================================================================================
// C ABI
// ==================================================================================================
using namespace vq;

static int im2col_launch(const void* images, bool u8, void* out_bf16, int32_t frames, int32_t H, int32_t W, int32_t patch,
                         int32_t Kp, cudaStream_t stream) {
  VQ_CHECK_ARG(images && out_bf16, "null pointer");
  VQ_CHECK_ARG(frames > 0 && patch > 0 && H >= patch && W >= patch, "image geometry");
  const int gh = H / patch, gw = W / patch;
  VQ_CHECK_ARG(Kp >= 3 * patch * patch && Kp % 8 == 0, "Kp must be >= 3*patch^2 and a multiple of 8");
  VQ_CHECK_ARG((reinterpret_cast<uintptr_t>(out_bf16) % 16) == 0, "patch matrix must be 16-byte aligned");
  const size_t smem = (size_t)3 * patch * gw * patch * sizeof(__nv_bfloat16);
  VQ_CHECK_ARG(smem <= 200 * 1024, "image too wide for the im2col strip");
  const void* kern = u8 ? reinterpret_cast<const void*>(im2col_kernel<true>) : reinterpret_cast<const void*>(im2col_kernel<false>);
  if (smem > 48 * 1024) VQ_CHECK_CUDA(ensure_dynamic_smem(kern, (int)smem));
  LaunchScope scope(kKindRow, u8 ? "im2col_u8" : "im2col", frames, H, W, 0,
                    (double)frames * (3.0 * H * W * (u8 ? 1 : 4) + (double)gh * gw * Kp * 2), stream);
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  if (u8 && patch == 14) {
    // raw byte strip (+ up to 15 bytes either side of it) and the 3 x 256 normalisation table
    const size_t smem14 = (((size_t)14 * W * 3 + 31 + 15) & ~(size_t)15) + 3 * 256 * sizeof(__nv_bfloat16);
    VQ_CHECK_ARG(smem14 <= 200 * 1024, "image too wide for the im2col strip");
    if (smem14 > 48 * 1024)
      VQ_CHECK_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(im2col_u8_kernel<14>), (int)smem14));
    im2col_u8_kernel<14><<<frames * gh, 256, smem14, stream>>>(reinterpret_cast<const unsigned char*>(images), out, H, W,
                                                               gh, gw, Kp);
  } else if (u8) {
    im2col_kernel<true><<<frames * gh, 256, smem, stream>>>(images, out, H, W, patch, gh, gw, Kp);
  } else {
    im2col_kernel<false><<<frames * gh, 256, smem, stream>>>(images, out, H, W, patch, gh, gw, Kp);
  }
  VQ_CHECK_CUDA(cudaGetLastError());
  return VQ_OK;
}

extern "C" int vq_im2col_patches(const float* images, void* out_bf16, int32_t frames, int32_t H, int32_t W,
                                 int32_t patch, int32_t Kp, vq_stream_t stream_) {
  return im2col_launch(images, false, out_bf16, frames, H, W, patch, Kp, reinterpret_cast<cudaStream_t>(stream_));
}

extern "C" int vq_im2col_patches_u8(const uint8_t* views_hwc, void* out_bf16, int32_t frames, int32_t H, int32_t W,
                                    int32_t patch, int32_t Kp, vq_stream_t stream_) {
  return im2col_launch(views_hwc, true, out_bf16, frames, H, W, patch, Kp, reinterpret_cast<cudaStream_t>(stream_));
}

extern "C" int vq_init_special_tokens(float* x, const float* cls_token, const float* pos_embed0,
                                      const float* register_tokens, int32_t frames, int32_t P, int32_t D,
                                      int32_t n_special, vq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VQ_CHECK_ARG(x && cls_token && pos_embed0 && register_tokens, "null pointer");
  VQ_CHECK_ARG(frames > 0 && P >= n_special && n_special >= 1, "token geometry");
  LaunchScope scope(kKindRow, "init_special", frames, n_special, D, 0, (double)frames * n_special * D * 4, stream);
  init_special_kernel<<<frames, 256, 0, stream>>>(x, cls_token, pos_embed0, register_tokens, P, D, n_special);
  VQ_CHECK_CUDA(cudaGetLastError());
  return VQ_OK;
}

template <bool IN_F32, bool OUT_F32>
static int launch_ln(const void* in, void* out, const float* gamma, const float* beta, long long rows, int D,
                     float eps, float* mean_out, float* rstd_out, cudaStream_t stream) {
  const int warps = 8;
  const unsigned grid = (unsigned)((rows + warps - 1) / warps);
  LaunchScope scope(kKindRow, "layernorm", (int)rows, D, 0, 0, (double)rows * D * ((IN_F32 ? 4 : 2) + (OUT_F32 ? 4 : 2)), stream);
#define VQ_LN_CASE(V)                                                                                      \
  case V:                                                                                                  \
    VQ_CHECK_CUDA(launch_pdl(layernorm_kernel<V, IN_F32, OUT_F32>, dim3(grid), dim3(warps * 32), 0, stream, in, out, \
                             gamma, beta, rows, D, eps, mean_out, rstd_out));                                    \
    break;
  switch (D / 128) {
    VQ_LN_CASE(1) VQ_LN_CASE(2) VQ_LN_CASE(3) VQ_LN_CASE(4) VQ_LN_CASE(6) VQ_LN_CASE(8) VQ_LN_CASE(16)
    VQ_LN_CASE(32)
    default:
      set_error("layernorm: unsupported width (D/128 must be one of 1,2,3,4,6,8,16,32)");
      return VQ_ERR_UNSUPPORTED;
  }
#undef VQ_LN_CASE
  VQ_CHECK_CUDA(cudaGetLastError());
  return VQ_OK;
}

extern "C" int vq_layernorm(const void* in, int32_t in_dtype, void* out, int32_t out_dtype, const float* gamma,
                            const float* beta, int64_t rows, int32_t D, float eps, float* mean_out,
                            float* rstd_out, vq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VQ_CHECK_ARG(in && out && gamma && beta, "null pointer");
  VQ_CHECK_ARG(rows > 0 && D > 0 && D % 128 == 0, "rows > 0 and D % 128 == 0 required");
  VQ_CHECK_ARG((mean_out == nullptr) == (rstd_out == nullptr), "mean_out and rstd_out go together");
  const bool i32 = in_dtype == VQ_F32, o32 = out_dtype == VQ_F32;
  if (i32 && !o32) return launch_ln<true, false>(in, out, gamma, beta, rows, D, eps, mean_out, rstd_out, stream);
  if (i32 && o32) return launch_ln<true, true>(in, out, gamma, beta, rows, D, eps, mean_out, rstd_out, stream);
  if (!i32 && !o32) return launch_ln<false, false>(in, out, gamma, beta, rows, D, eps, mean_out, rstd_out, stream);
  return launch_ln<false, true>(in, out, gamma, beta, rows, D, eps, mean_out, rstd_out, stream);
}

extern "C" int vq_residual_layernorm(float* x, const void* y_bf16, void* xn_bf16, const float* gamma, const float* beta,
                                     int64_t rows, int32_t D, float eps, vq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VQ_CHECK_ARG(x && y_bf16, "null pointer");
  VQ_CHECK_ARG(!xn_bf16 || (gamma && beta), "LayerNorm output needs gamma and beta");
  VQ_CHECK_ARG(rows > 0 && D > 0 && D % 128 == 0, "rows > 0 and D % 128 == 0 required");
  const int warps = 8;
  const unsigned grid = (unsigned)((rows + warps - 1) / warps);
  const __nv_bfloat16* y = reinterpret_cast<const __nv_bfloat16*>(y_bf16);
  __nv_bfloat16* xn = reinterpret_cast<__nv_bfloat16*>(xn_bf16);
  LaunchScope scope(kKindRow, xn ? "residual_layernorm" : "residual_add", (int)rows, D, 0, 0,
                    (double)rows * D * (xn ? 12.0 : 10.0), stream);
#define VQ_RL_CASE(V)                                                                                            \
  case V:                                                                                                        \
    VQ_CHECK_CUDA(launch_pdl(residual_layernorm_kernel<V>, dim3(grid), dim3(warps * 32), 0, stream, x, y, xn,    \
                             gamma, beta, (long long)rows, D, eps));                                             \
    break;
  switch (D / 128) {
    VQ_RL_CASE(1) VQ_RL_CASE(2) VQ_RL_CASE(3) VQ_RL_CASE(4) VQ_RL_CASE(6) VQ_RL_CASE(8)
    default:
      set_error("residual_layernorm: unsupported width");
      return VQ_ERR_UNSUPPORTED;
  }
#undef VQ_RL_CASE
  VQ_CHECK_CUDA(cudaGetLastError());
  return VQ_OK;
}

extern "C" int vq_row_stats_cast(const float* x, void* xb_bf16, float* stats, float* stats_zero, int64_t rows, int32_t D,
                                 vq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VQ_CHECK_ARG(x && xb_bf16 && stats, "null pointer");
  VQ_CHECK_ARG(rows > 0 && D > 0 && D % 128 == 0, "rows > 0 and D % 128 == 0 required");
  const int warps = 8;
  const unsigned grid = (unsigned)((rows + warps - 1) / warps);
  __nv_bfloat16* xb = reinterpret_cast<__nv_bfloat16*>(xb_bf16);
  LaunchScope scope(kKindRow, "row_stats_cast", (int)rows, D, 0, 0, (double)rows * D * 6.0, stream);
#define VQ_RS_CASE(V)                                                                                              \
  case V:                                                                                                          \
    VQ_CHECK_CUDA(launch_pdl(row_stats_cast_kernel<V>, dim3(grid), dim3(warps * 32), 0, stream, x, xb, stats,      \
                             stats_zero, (long long)rows, D));                                                     \
    break;
  switch (D / 128) {
    VQ_RS_CASE(1) VQ_RS_CASE(2) VQ_RS_CASE(3) VQ_RS_CASE(4) VQ_RS_CASE(6) VQ_RS_CASE(8)
    default:
      set_error("row_stats_cast: unsupported width");
      return VQ_ERR_UNSUPPORTED;
  }
#undef VQ_RS_CASE
  VQ_CHECK_CUDA(cudaGetLastError());
  return VQ_OK;
}

extern "C" int vq_stem_finalize(float* x, const float* gamma, const float* beta, const float* camera_token,
                                const float* register_token, int64_t rows, int32_t D, int32_t P, int32_t S,
                                int32_t n_special, float eps, vq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VQ_CHECK_ARG(x && gamma && beta && camera_token && register_token, "null pointer");
  VQ_CHECK_ARG(rows > 0 && D % 128 == 0 && P > n_special && S > 0, "geometry");
  const int warps = 8;
  const unsigned grid = (unsigned)((rows + warps - 1) / warps);
  LaunchScope scope(kKindRow, "stem_finalize", (int)rows, D, 0, 0, (double)rows * D * 8, stream);
#define VQ_SF_CASE(V)                                                                                       \
  case V:                                                                                                   \
    stem_finalize_kernel<V><<<grid, warps * 32, 0, stream>>>(x, gamma, beta, camera_token, register_token, rows, \
                                                             D, P, S, n_special, eps);                      \
    break;
  switch (D / 128) {
    VQ_SF_CASE(1) VQ_SF_CASE(2) VQ_SF_CASE(3) VQ_SF_CASE(4) VQ_SF_CASE(6) VQ_SF_CASE(8)
    default:
      set_error("stem_finalize: unsupported width");
      return VQ_ERR_UNSUPPORTED;
  }
#undef VQ_SF_CASE
  VQ_CHECK_CUDA(cudaGetLastError());
  return VQ_OK;
}

extern "C" int vq_gather_tokens(const float* x, void* out_bf16, int32_t B, int64_t tokens_per_scene, int32_t n_take,
                                int32_t D, int32_t out_ld, int32_t col_off, vq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VQ_CHECK_ARG(x && out_bf16, "null pointer");
  VQ_CHECK_ARG(B > 0 && n_take > 0 && n_take <= tokens_per_scene && D % 4 == 0, "geometry");
  VQ_CHECK_ARG(out_ld % 4 == 0 && col_off % 4 == 0 && col_off + D <= out_ld, "output layout");
  const long long total = (long long)B * n_take;
  const int warps = 8;
  LaunchScope scope(kKindRow, "gather_tokens", B, n_take, D, 0, (double)total * D * 6, stream);
  gather_tokens_kernel<<<(unsigned)((total + warps - 1) / warps), warps * 32, 0, stream>>>(
      x, reinterpret_cast<__nv_bfloat16*>(out_bf16), D, tokens_per_scene, n_take, out_ld, col_off, total);
  VQ_CHECK_CUDA(cudaGetLastError());
  return VQ_OK;
}

extern "C" int vq_splice_overwrite(void* inputs_embeds, int32_t dtype, const int64_t* input_ids, const void* features,
                                   int32_t B, int32_t L, int32_t hidden, int32_t n_feat, int64_t image_id,
                                   int32_t* overflow_count, vq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VQ_CHECK_ARG(inputs_embeds && input_ids && features, "null pointer");
  VQ_CHECK_ARG(B > 0 && L > 0 && n_feat > 0, "geometry");
  const int elt = dtype == VQ_F32 ? 4 : 2;
  VQ_CHECK_ARG((hidden * elt) % 16 == 0, "hidden size must give 16-byte rows");
  VQ_CHECK_ARG((reinterpret_cast<uintptr_t>(inputs_embeds) % 16) == 0 && (reinterpret_cast<uintptr_t>(features) % 16) == 0,
               "16-byte alignment");
  const long long total = (long long)B * L;
  const int warps = 8;
  const unsigned grid = (unsigned)((total + warps - 1) / warps);
  LaunchScope scope(kKindRow, "splice_overwrite", B, L, n_feat, 0, (double)B * n_feat * hidden * elt * 2, stream);
  if (dtype == VQ_F32)
    splice_overwrite_kernel<float><<<grid, warps * 32, 0, stream>>>(
        reinterpret_cast<float*>(inputs_embeds), reinterpret_cast<const long long*>(input_ids),
        reinterpret_cast<const float*>(features), B, L, hidden, n_feat, image_id, overflow_count);
  else
    splice_overwrite_kernel<__nv_bfloat16><<<grid, warps * 32, 0, stream>>>(
        reinterpret_cast<__nv_bfloat16*>(inputs_embeds), reinterpret_cast<const long long*>(input_ids),
        reinterpret_cast<const __nv_bfloat16*>(features), B, L, hidden, n_feat, image_id, overflow_count);
  VQ_CHECK_CUDA(cudaGetLastError());
  return VQ_OK;
}
