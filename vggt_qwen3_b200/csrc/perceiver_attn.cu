// One-tile attention core of the Perceiver projector: nn.MultiheadAttention(latents, context, context) of
// src/models/projector_perceiver.py:33,45 and its autograd backward, one CTA per (sample, head).
//
// The whole score matrix of a (sample, head) is ONE 128 x 128 tile (num_latents <= 128 queries, <= 128 context
// tokens), so the five launches of the GEMM formulation (S = q k^T | softmax | P v, and dP | softmax' | dq | dk | dv
// backward) with S / dP materialised in f32 HBM collapse into one kernel each way:
//   forward : S (TMEM, 128 f32 columns) = sum over head_dim chunks of q_j k_j^T -> row softmax (+ attention-probability
//             dropout) in registers, one thread per latent row -> P (bf16) to smem as the A operand of P v and to HBM
//             for the backward -> O = P v in 128-column pieces through two TMEM accumulators -> bf16 `att`.
//   backward: dP (TMEM) = sum_j dO_j v_j^T -> dS = scale P (mask dP / (1-p) - rowsum) in registers -> dS and
//             Pd = dropout(P) (regenerated, bit-identical to the forward) to smem, where ONE image serves as the
//             K-major A operand of dq = dS k and as the MN-major A operand of dk = dS^T q, dv = Pd^T dO -> per 64-wide
//             head_dim chunk three 128 x 64 accumulators (double-buffered in TMEM) -> bf16 dq / dk / dv.
// A 128-row x 64-column bf16 box (SWIZZLE_128B) is at the same time a K-major operand chunk (rows = M/N index) and an
// MN-major one (rows = k index), so q_j, k_j, v_j, dO_j are each loaded by one TMA box whichever role they play.
// warps 0-3: one row each (softmax / epilogue), warp 4: TMA producer, warp 5: MMA issuer.
#include <math.h>
#include <stdlib.h>

#include <cuda_bf16.h>

#include "common.cuh"
#include "philox.cuh"
#include "ptx.cuh"

namespace vq {
namespace {

constexpr int kPaThreads = 192;
constexpr int kPaTile = 128;                      // latent rows / keys per tile
constexpr int kPaChunkBytes = kPaTile * 128;      // 128 rows x 64 bf16 = 16 KB
constexpr int kPaTmemCols = 512;

struct PaArgs {
  int Lq, T, Tp, H, hd, E;
  float scale;
  DropoutParams drop;
  __nv_bfloat16* P;         // [B*H*Lq, Tp]  softmax probabilities (fwd: out, bwd: in)
  __nv_bfloat16* Pd;        // [B*H*Lq, Tp]  dropout(P) (fwd out, train mode only)
  __nv_bfloat16* att;       // [B*Lq, E]     fwd out
  __nv_bfloat16* dq;        // [B*Lq, E]     bwd out
  __nv_bfloat16* dkv;       // [B*T, 2E]     bwd out: dk | dv
};

// 16-byte piece `c16` (8 bf16 along the 128-wide inner dimension) of row `r` of a [128 x 128] bf16 tile kept as two
// 64-column SWIZZLE_128B blocks of 16 KB: the layout TMA would have produced, readable K-major and MN-major.
__device__ __forceinline__ void st_tile16(uint8_t* tile, int r, int c16, uint4 v) {
  const int blk = c16 >> 3, c = c16 & 7;
  *reinterpret_cast<uint4*>(tile + blk * kPaChunkBytes + r * 128 + ((c ^ (r & 7)) << 4)) = v;
}

__device__ __forceinline__ uint32_t taddr_of(uint32_t tmem_base, int warp, int col) {
  return tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)col;
}

// ------------------------------------------------------------------------------------------------ forward
constexpr int kFwdStages = 5;                              // 32 KB each: (q_j, k_j) or one 128-column piece of v
constexpr int kFwdStageBytes = 2 * kPaChunkBytes;
constexpr int kFwdSmemBytes = 1024 + kFwdStages * kFwdStageBytes + 2 * kPaChunkBytes + 256;

template <bool DROP>
__global__ void __launch_bounds__(kPaThreads, 1)
perceiver_attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                          const PaArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t smem_p = smem_base + kFwdStages * kFwdStageBytes;
  uint8_t* p_tile = smem_gen + kFwdStages * kFwdStageBytes;
  const uint32_t bar_base = smem_p + 2 * kPaChunkBytes;
  const uint32_t bar_full = bar_base, bar_empty = bar_base + 8 * kFwdStages;
  const uint32_t bar_s = bar_base + 16 * kFwdStages, bar_p = bar_s + 8, bar_ofull = bar_s + 16, bar_oempty = bar_s + 32;
  const uint32_t tmem_slot = bar_s + 48;
  const uint32_t* tmem_slot_ptr = reinterpret_cast<const uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / a.H, h = blockIdx.x - b * a.H;
  const int nkb = a.hd >> 6;                 // 64-wide head_dim chunks
  const int npc = (a.hd + 127) >> 7;         // 128-wide output pieces

  pdl_launch_dependents();
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    for (int s = 0; s < kFwdStages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_s, 1);
    mbar_init(bar_p, kPaTile);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_ofull + 8 * s, 1);
      mbar_init(bar_oempty + 8 * s, kPaTile);
    }
    fence_barrier_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, kPaTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 128;   // S: 128 columns; O: 2 x 128 columns
  pdl_wait();

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nkb; ++j) {
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        const uint32_t full = bar_full + 8 * stage, dst = smem_base + stage * kFwdStageBytes;
        mbar_arrive_expect_tx(full, 2 * kPaChunkBytes);
        tma_load_2d(dst, &tmQ, full, h * a.hd + 64 * j, b * a.Lq);
        tma_load_2d(dst + kPaChunkBytes, &tmKV, full, h * a.hd + 64 * j, b * a.T);
        if (++stage == kFwdStages) { stage = 0; phase ^= 1; }
      }
      for (int q = 0; q < npc; ++q) {
        const int nch = (a.hd - 128 * q) >= 128 ? 2 : 1;
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        const uint32_t full = bar_full + 8 * stage, dst = smem_base + stage * kFwdStageBytes;
        mbar_arrive_expect_tx(full, nch * kPaChunkBytes);
        for (int i = 0; i < nch; ++i)
          tma_load_2d(dst + i * kPaChunkBytes, &tmKV, full, a.E + h * a.hd + 128 * q + 64 * i, b * a.T);
        if (++stage == kFwdStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t idesc_s = make_idesc_bf16(kPaTile, kPaTile, 0, 0);
    for (int j = 0; j < nkb; ++j) {
      mbar_wait(bar_full + 8 * stage, phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_base + stage * kFwdStageBytes;
        const uint64_t da = make_smem_desc(sa, 16, 1024), db = make_smem_desc(sa + kPaChunkBytes, 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tmem_s, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_s, (j | k) ? 1u : 0u);
        umma_commit(bar_empty + 8 * stage);
        if (j == nkb - 1) umma_commit(bar_s);
      }
      __syncwarp();
      if (++stage == kFwdStages) { stage = 0; phase ^= 1; }
    }
    mbar_wait(bar_p, 0);
    tc_fence_after();
    for (int q = 0; q < npc; ++q) {
      const int ncol = (a.hd - 128 * q) >= 128 ? 128 : 64;
      const uint32_t idesc_o = make_idesc_bf16(kPaTile, ncol, 0, 1);
      mbar_wait(bar_full + 8 * stage, phase);
      mbar_wait(bar_oempty + 8 * (q & 1), ((q >> 1) & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sv = smem_base + stage * kFwdStageBytes;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t da = make_smem_desc(smem_p + (ks >> 2) * kPaChunkBytes + (ks & 3) * 32, 16, 1024);
          const uint64_t db = make_smem_desc(sv + ks * 2048, kPaChunkBytes, 1024);
          umma_ss(tmem_o + (q & 1) * 128, da, db, idesc_o, ks ? 1u : 0u);
        }
        umma_commit(bar_empty + 8 * stage);
        umma_commit(bar_ofull + 8 * (q & 1));
      }
      __syncwarp();
      if (++stage == kFwdStages) { stage = 0; phase ^= 1; }
    }
  } else {
    // ===================== softmax + epilogue: thread = latent row =====================
    const int r = threadIdx.x;
    const bool row_ok = r < a.Lq;
    const long long prow = ((long long)b * a.H + h) * a.Lq + r;      // row of P / Pd, and of the dropout site
    mbar_wait(bar_s, 0);
    tc_fence_after();
    uint32_t sr[4][32];
#pragma unroll
    for (int c = 0; c < 4; ++c) tmem_ld32(taddr_of(tmem_s, warp, 32 * c), sr[c]);
    tmem_wait_ld();
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float v = (32 * c + i) < a.T ? __uint_as_float(sr[c][i]) * a.scale : -INFINITY;
        sr[c][i] = __float_as_uint(v);
        mx = fmaxf(mx, v);
      }
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float e = (32 * c + i) < a.T ? __expf(__uint_as_float(sr[c][i]) - mx) : 0.f;
        sr[c][i] = __float_as_uint(e);
        sum += e;
      }
    const float inv = 1.0f / sum;
    __nv_bfloat16* prow_p = a.P + prow * a.Tp;
    __nv_bfloat16* prow_pd = DROP ? a.Pd + prow * a.Tp : nullptr;
#pragma unroll
    for (int g = 0; g < 16; ++g) {            // 8 columns at a time
      const int c = g >> 2, i0 = (g & 3) * 8;
      uint32_t pk[4], dk[4];
#pragma unroll
      for (int t = 0; t < 4; ++t)
        pk[t] = pack_bf16x2(__uint_as_float(sr[c][i0 + 2 * t]) * inv, __uint_as_float(sr[c][i0 + 2 * t + 1]) * inv);
      const bool col_ok = 8 * g < a.Tp;
      if (row_ok && col_ok) *reinterpret_cast<uint4*>(prow_p + 8 * g) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      if constexpr (DROP) {
        // attention-probability dropout acts on the bf16 probability the P v product would otherwise read
        const uint32_t keep = col_ok ? dropout_keep8(a.drop, (uint64_t)prow * (uint64_t)(a.Tp >> 3) + (uint64_t)g) : 0u;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float lo = (keep >> (2 * t)) & 1u ? bf16_lo(pk[t]) * a.drop.scale : 0.f;
          const float hi = (keep >> (2 * t + 1)) & 1u ? bf16_hi(pk[t]) * a.drop.scale : 0.f;
          dk[t] = pack_bf16x2(lo, hi);
        }
        if (row_ok && col_ok) *reinterpret_cast<uint4*>(prow_pd + 8 * g) = make_uint4(dk[0], dk[1], dk[2], dk[3]);
      } else {
#pragma unroll
        for (int t = 0; t < 4; ++t) dk[t] = pk[t];
      }
      st_tile16(p_tile, r, g, make_uint4(dk[0], dk[1], dk[2], dk[3]));
    }
    fence_proxy_async_smem();
    mbar_arrive(bar_p);
    // O pieces -> att[b*Lq + r, h*hd + ...]
    __nv_bfloat16* orow = a.att + ((long long)b * a.Lq + r) * a.E + h * a.hd;
    for (int q = 0; q < npc; ++q) {
      const int ncol = (a.hd - 128 * q) >= 128 ? 128 : 64;
      mbar_wait(bar_ofull + 8 * (q & 1), (q >> 1) & 1);
      tc_fence_after();
      for (int c = 0; c < ncol; c += 32) {
        uint32_t o[32];
        tmem_ld32(taddr_of(tmem_o + (q & 1) * 128, warp, c), o);
        tmem_wait_ld();
        if (row_ok) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(o[8 * v + 0]), __uint_as_float(o[8 * v + 1]));
            w.y = pack_bf16x2(__uint_as_float(o[8 * v + 2]), __uint_as_float(o[8 * v + 3]));
            w.z = pack_bf16x2(__uint_as_float(o[8 * v + 4]), __uint_as_float(o[8 * v + 5]));
            w.w = pack_bf16x2(__uint_as_float(o[8 * v + 6]), __uint_as_float(o[8 * v + 7]));
            *reinterpret_cast<uint4*>(orow + 128 * q + c + 8 * v) = w;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_oempty + 8 * (q & 1));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, kPaTmemCols);
}

// ------------------------------------------------------------------------------------------------ backward
constexpr int kBwdStages = 3;                              // 48 KB each: (dO_j, v_j) then (k_j, q_j, dO_j)
constexpr int kBwdStageBytes = 3 * kPaChunkBytes;
constexpr int kBwdSmemBytes = 1024 + kBwdStages * kBwdStageBytes + 4 * kPaChunkBytes + 256;

template <bool DROP>
__global__ void __launch_bounds__(kPaThreads, 1)
perceiver_attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                          const __grid_constant__ CUtensorMap tmDO, const PaArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t smem_ds = smem_base + kBwdStages * kBwdStageBytes, smem_pd = smem_ds + 2 * kPaChunkBytes;
  uint8_t* ds_tile = smem_gen + kBwdStages * kBwdStageBytes;
  uint8_t* pd_tile = ds_tile + 2 * kPaChunkBytes;
  const uint32_t bar_base = smem_pd + 2 * kPaChunkBytes;
  const uint32_t bar_full = bar_base, bar_empty = bar_base + 8 * kBwdStages;
  const uint32_t bar_dp = bar_base + 16 * kBwdStages, bar_ds = bar_dp + 8, bar_afull = bar_dp + 16, bar_aempty = bar_dp + 32;
  const uint32_t tmem_slot = bar_dp + 48;
  const uint32_t* tmem_slot_ptr = reinterpret_cast<const uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / a.H, h = blockIdx.x - b * a.H;
  const int nkb = a.hd >> 6;

  pdl_launch_dependents();
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmDO);
    for (int s = 0; s < kBwdStages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_dp, 1);
    mbar_init(bar_ds, kPaTile);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_afull + 8 * s, 1);
      mbar_init(bar_aempty + 8 * s, kPaTile);
    }
    fence_barrier_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, kPaTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t tmem_dp = tmem_base, tmem_acc = tmem_base + 128;   // dP: 128 columns; 2 x [dq | dk | dv] x 64 columns
  pdl_wait();

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nkb; ++j) {
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        const uint32_t full = bar_full + 8 * stage, dst = smem_base + stage * kBwdStageBytes;
        mbar_arrive_expect_tx(full, 2 * kPaChunkBytes);
        tma_load_2d(dst, &tmDO, full, h * a.hd + 64 * j, b * a.Lq);
        tma_load_2d(dst + kPaChunkBytes, &tmKV, full, a.E + h * a.hd + 64 * j, b * a.T);
        if (++stage == kBwdStages) { stage = 0; phase ^= 1; }
      }
      for (int j = 0; j < nkb; ++j) {
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        const uint32_t full = bar_full + 8 * stage, dst = smem_base + stage * kBwdStageBytes;
        mbar_arrive_expect_tx(full, 3 * kPaChunkBytes);
        tma_load_2d(dst, &tmKV, full, h * a.hd + 64 * j, b * a.T);
        tma_load_2d(dst + kPaChunkBytes, &tmQ, full, h * a.hd + 64 * j, b * a.Lq);
        tma_load_2d(dst + 2 * kPaChunkBytes, &tmDO, full, h * a.hd + 64 * j, b * a.Lq);
        if (++stage == kBwdStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t idesc_dp = make_idesc_bf16(kPaTile, kPaTile, 0, 0);
    for (int j = 0; j < nkb; ++j) {      // dP += dO_j v_j^T
      mbar_wait(bar_full + 8 * stage, phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_base + stage * kBwdStageBytes;
        const uint64_t da = make_smem_desc(sa, 16, 1024), db = make_smem_desc(sa + kPaChunkBytes, 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tmem_dp, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_dp, (j | k) ? 1u : 0u);
        umma_commit(bar_empty + 8 * stage);
        if (j == nkb - 1) umma_commit(bar_dp);
      }
      __syncwarp();
      if (++stage == kBwdStages) { stage = 0; phase ^= 1; }
    }
    mbar_wait(bar_ds, 0);
    tc_fence_after();
    const uint32_t idesc_dq = make_idesc_bf16(kPaTile, 64, 0, 1);    // dS (K-major)    x k_j  (MN-major)
    const uint32_t idesc_dkv = make_idesc_bf16(kPaTile, 64, 1, 1);   // dS^T / Pd^T (MN-major) x q_j / dO_j (MN-major)
    for (int j = 0; j < nkb; ++j) {
      mbar_wait(bar_full + 8 * stage, phase);
      mbar_wait(bar_aempty + 8 * (j & 1), ((j >> 1) & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sk = smem_base + stage * kBwdStageBytes, sq = sk + kPaChunkBytes, so = sk + 2 * kPaChunkBytes;
        const uint32_t acc = tmem_acc + (j & 1) * 192;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_ss(acc, make_smem_desc(smem_ds + (ks >> 2) * kPaChunkBytes + (ks & 3) * 32, 16, 1024),
                  make_smem_desc(sk + ks * 2048, kPaChunkBytes, 1024), idesc_dq, ks ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_ss(acc + 64, make_smem_desc(smem_ds + ks * 2048, kPaChunkBytes, 1024),
                  make_smem_desc(sq + ks * 2048, kPaChunkBytes, 1024), idesc_dkv, ks ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_ss(acc + 128, make_smem_desc(smem_pd + ks * 2048, kPaChunkBytes, 1024),
                  make_smem_desc(so + ks * 2048, kPaChunkBytes, 1024), idesc_dkv, ks ? 1u : 0u);
        umma_commit(bar_empty + 8 * stage);
        umma_commit(bar_afull + 8 * (j & 1));
      }
      __syncwarp();
      if (++stage == kBwdStages) { stage = 0; phase ^= 1; }
    }
  } else {
    // ===================== softmax backward + epilogue: thread = latent row (then = key row) =====================
    const int r = threadIdx.x;
    const bool row_ok = r < a.Lq;
    const long long prow = ((long long)b * a.H + h) * a.Lq + r;
    // P row and keep bits: independent of dP, fetched while the dP products run
    uint32_t pk[64];
    uint32_t keep[4] = {0u, 0u, 0u, 0u};
    const __nv_bfloat16* prow_p = a.P + prow * a.Tp;
#pragma unroll
    for (int g = 0; g < 16; ++g) {
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      const bool ok = row_ok && 8 * g < a.Tp;
      if (ok) v = *reinterpret_cast<const uint4*>(prow_p + 8 * g);
      pk[4 * g + 0] = v.x; pk[4 * g + 1] = v.y; pk[4 * g + 2] = v.z; pk[4 * g + 3] = v.w;
      if constexpr (DROP) {
        const uint32_t k8 = ok ? dropout_keep8(a.drop, (uint64_t)prow * (uint64_t)(a.Tp >> 3) + (uint64_t)g) : 0u;
        keep[g >> 2] |= k8 << (8 * (g & 3));
      }
    }
    mbar_wait(bar_dp, 0);
    tc_fence_after();
    // two passes over the dP row in TMEM (32 columns at a time) instead of 128 live registers: rowsum, then dS
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t dr[32];
      tmem_ld32(taddr_of(tmem_dp, warp, 32 * c), dr);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int col = 32 * c + i;
        float dv = (row_ok && col < a.T) ? __uint_as_float(dr[i]) : 0.f;
        if constexpr (DROP) dv = (keep[c] >> i) & 1u ? dv * a.drop.scale : 0.f;
        const uint32_t w = pk[col >> 1];
        dot += ((col & 1) ? bf16_hi(w) : bf16_lo(w)) * dv;
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t dr[32];
      tmem_ld32(taddr_of(tmem_dp, warp, 32 * c), dr);
      tmem_wait_ld();
#pragma unroll
      for (int g4 = 0; g4 < 4; ++g4) {
        const int g = 4 * c + g4, i0 = 8 * g4;
        uint32_t ds[4], pd[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const uint32_t w = pk[4 * g + t];
          const float p0 = bf16_lo(w), p1 = bf16_hi(w);
          float d0 = (row_ok && 8 * g + 2 * t < a.T) ? __uint_as_float(dr[i0 + 2 * t]) : 0.f;
          float d1 = (row_ok && 8 * g + 2 * t + 1 < a.T) ? __uint_as_float(dr[i0 + 2 * t + 1]) : 0.f;
          if constexpr (DROP) {
            const uint32_t kb = keep[c] >> (i0 + 2 * t);
            d0 = kb & 1u ? d0 * a.drop.scale : 0.f;
            d1 = kb & 2u ? d1 * a.drop.scale : 0.f;
            pd[t] = pack_bf16x2(kb & 1u ? p0 * a.drop.scale : 0.f, kb & 2u ? p1 * a.drop.scale : 0.f);
          } else {
            pd[t] = w;
          }
          ds[t] = pack_bf16x2(a.scale * p0 * (d0 - dot), a.scale * p1 * (d1 - dot));
        }
        st_tile16(ds_tile, r, g, make_uint4(ds[0], ds[1], ds[2], ds[3]));
        st_tile16(pd_tile, r, g, make_uint4(pd[0], pd[1], pd[2], pd[3]));
      }
    }
    fence_proxy_async_smem();
    mbar_arrive(bar_ds);
    // per head_dim chunk: dq row r (latent), dk / dv row r (key)
    __nv_bfloat16* dq_row = a.dq + ((long long)b * a.Lq + r) * a.E + h * a.hd;
    __nv_bfloat16* dk_row = a.dkv + ((long long)b * a.T + r) * (2 * a.E) + h * a.hd;
    __nv_bfloat16* dv_row = dk_row + a.E;
    const bool key_ok = r < a.T;
    for (int j = 0; j < nkb; ++j) {
      mbar_wait(bar_afull + 8 * (j & 1), (j >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int w3 = 0; w3 < 3; ++w3) {
        __nv_bfloat16* dst = (w3 == 0 ? dq_row : w3 == 1 ? dk_row : dv_row) + 64 * j;
        const bool ok = w3 == 0 ? row_ok : key_ok;
#pragma unroll
        for (int c = 0; c < 64; c += 32) {
          uint32_t o[32];
          tmem_ld32(taddr_of(tmem_acc + (j & 1) * 192 + 64 * w3, warp, c), o);
          tmem_wait_ld();
          if (ok) {
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              uint4 w;
              w.x = pack_bf16x2(__uint_as_float(o[8 * v + 0]), __uint_as_float(o[8 * v + 1]));
              w.y = pack_bf16x2(__uint_as_float(o[8 * v + 2]), __uint_as_float(o[8 * v + 3]));
              w.z = pack_bf16x2(__uint_as_float(o[8 * v + 4]), __uint_as_float(o[8 * v + 5]));
              w.w = pack_bf16x2(__uint_as_float(o[8 * v + 6]), __uint_as_float(o[8 * v + 7]));
              *reinterpret_cast<uint4*>(dst + c + 8 * v) = w;
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_aempty + 8 * (j & 1));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, kPaTmemCols);
}

int g_fused = -1;   // -1: from the environment on first use

DropoutParams params_of(const vq_dropout_desc* d) {
  if (!d) return make_dropout_params(0.f, 0, 0, 0);
  return make_dropout_params(d->p, d->seed, d->offset, d->site);
}

int check_geometry(int B, int heads, int Lq, int T, int E) {
  VQ_CHECK_ARG(B > 0 && heads > 0 && E > 0 && E % heads == 0 && (E / heads) % 64 == 0, "head_dim must be a multiple of 64");
  VQ_CHECK_ARG((Lq == 64 || Lq == 128) && T > 0 && T <= 128 && T % 8 == 0,
               "one-tile attention: num_latents 64 or 128, T <= 128 and a multiple of 8");
  return VQ_OK;
}

bool fill_args(PaArgs& a, int B, int heads, int Lq, int T, int E, const vq_dropout_desc* dropout) {
  (void)B;
  a.Lq = Lq; a.T = T; a.Tp = (T + 63) / 64 * 64; a.H = heads; a.hd = E / heads; a.E = E;
  a.scale = 1.0f / sqrtf((float)a.hd);
  a.drop = params_of(dropout);
  a.P = a.Pd = a.att = a.dq = a.dkv = nullptr;
  return true;
}

}  // namespace
}  // namespace vq

using namespace vq;

extern "C" int vq_perceiver_attention_tune(int32_t fused) {
  if (fused >= 0) g_fused = fused ? 1 : 0;
  return VQ_OK;
}

extern "C" int vq_perceiver_attention_supported(int32_t Lq, int32_t T, int32_t E, int32_t heads) {
  if (g_fused < 0) {
    const char* e = getenv("VQ_PERCEIVER_FUSED_ATTN");
    g_fused = (e && atoi(e) == 0) ? 0 : 1;
  }
  if (!g_fused) return 0;
  return (heads > 0 && E > 0 && E % heads == 0 && (E / heads) % 64 == 0 && (Lq == 64 || Lq == 128) && T > 0 && T <= 128 &&
          T % 8 == 0)
             ? 1
             : 0;
}

extern "C" int vq_perceiver_attention_fwd(const void* q, const void* kv, void* probs, void* probs_dropped, void* att,
                                          int32_t B, int32_t heads, int32_t Lq, int32_t T, int32_t E,
                                          const vq_dropout_desc* dropout, vq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VQ_CHECK_ARG(q && kv && probs && att, "null pointer");
  VQ_CHECK_ARG(!dropout || (dropout->p >= 0.f && dropout->p < 1.f), "dropout p must be in [0, 1)");
  if (int rc = check_geometry(B, heads, Lq, T, E)) return rc;
  PaArgs a;
  fill_args(a, B, heads, Lq, T, E, dropout);
  VQ_CHECK_ARG(!a.drop.thr || probs_dropped, "dropout needs probs_dropped");
  a.P = reinterpret_cast<__nv_bfloat16*>(probs);
  a.Pd = reinterpret_cast<__nv_bfloat16*>(probs_dropped);
  a.att = reinterpret_cast<__nv_bfloat16*>(att);
  CUtensorMap mq, mkv;
  if (!get_tmap_2d_bf16(q, (uint64_t)E, (uint64_t)B * Lq, (uint64_t)E, 64, kPaTile, &mq)) return VQ_ERR_CUDA;
  if (!get_tmap_2d_bf16(kv, (uint64_t)2 * E, (uint64_t)B * T, (uint64_t)2 * E, 64, kPaTile, &mkv)) return VQ_ERR_CUDA;
  const double macs = (double)B * heads * Lq * (double)T * a.hd;
  LaunchScope scope(kKindAttention, "perceiver_attn_fwd", B * heads, Lq, T, 4.0 * macs,
                    2.0 * ((double)B * Lq * E * 2 + (double)B * T * 2 * E + (double)B * heads * Lq * a.Tp), stream);
  if (a.drop.thr) {
    VQ_CHECK_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(perceiver_attn_fwd_kernel<true>), kFwdSmemBytes));
    VQ_CHECK_CUDA(launch_pdl(perceiver_attn_fwd_kernel<true>, dim3(B * heads), dim3(kPaThreads), kFwdSmemBytes, stream,
                             mq, mkv, a));
  } else {
    VQ_CHECK_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(perceiver_attn_fwd_kernel<false>), kFwdSmemBytes));
    VQ_CHECK_CUDA(launch_pdl(perceiver_attn_fwd_kernel<false>, dim3(B * heads), dim3(kPaThreads), kFwdSmemBytes, stream,
                             mq, mkv, a));
  }
  return VQ_OK;
}

extern "C" int vq_perceiver_attention_bwd(const void* q, const void* kv, const void* probs, const void* datt, void* dq,
                                          void* dkv, int32_t B, int32_t heads, int32_t Lq, int32_t T, int32_t E,
                                          const vq_dropout_desc* dropout, vq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VQ_CHECK_ARG(q && kv && probs && datt && dq && dkv, "null pointer");
  VQ_CHECK_ARG(!dropout || (dropout->p >= 0.f && dropout->p < 1.f), "dropout p must be in [0, 1)");
  if (int rc = check_geometry(B, heads, Lq, T, E)) return rc;
  PaArgs a;
  fill_args(a, B, heads, Lq, T, E, dropout);
  a.P = reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(probs));
  a.dq = reinterpret_cast<__nv_bfloat16*>(dq);
  a.dkv = reinterpret_cast<__nv_bfloat16*>(dkv);
  CUtensorMap mq, mkv, mdo;
  if (!get_tmap_2d_bf16(q, (uint64_t)E, (uint64_t)B * Lq, (uint64_t)E, 64, kPaTile, &mq)) return VQ_ERR_CUDA;
  if (!get_tmap_2d_bf16(kv, (uint64_t)2 * E, (uint64_t)B * T, (uint64_t)2 * E, 64, kPaTile, &mkv)) return VQ_ERR_CUDA;
  if (!get_tmap_2d_bf16(datt, (uint64_t)E, (uint64_t)B * Lq, (uint64_t)E, 64, kPaTile, &mdo)) return VQ_ERR_CUDA;
  const double macs = (double)B * heads * Lq * (double)T * a.hd;
  LaunchScope scope(kKindAttention, "perceiver_attn_bwd", B * heads, Lq, T, 8.0 * macs,
                    2.0 * (2.0 * (double)B * Lq * E * 2 + 2.0 * (double)B * T * 2 * E + (double)B * heads * Lq * a.Tp),
                    stream);
  if (a.drop.thr) {
    VQ_CHECK_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(perceiver_attn_bwd_kernel<true>), kBwdSmemBytes));
    VQ_CHECK_CUDA(launch_pdl(perceiver_attn_bwd_kernel<true>, dim3(B * heads), dim3(kPaThreads), kBwdSmemBytes, stream,
                             mq, mkv, mdo, a));
  } else {
    VQ_CHECK_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(perceiver_attn_bwd_kernel<false>), kBwdSmemBytes));
    VQ_CHECK_CUDA(launch_pdl(perceiver_attn_bwd_kernel<false>, dim3(B * heads), dim3(kPaThreads), kBwdSmemBytes, stream,
                             mq, mkv, mdo, a));
  }
  return VQ_OK;
}
