// vq_gemm: persistent, warp-specialised bf16 GEMM on the 5th-gen tensor cores (sm_100a).
//
//   warps 0..7  epilogue       tcgen05.ld (one accumulator row per thread) -> fused epilogue -> global
//   warp 8      TMA producer   cp.async.bulk.tensor (SWIZZLE_128B boxes) -> smem ring, mbarrier full/empty
//   warp 9      MMA issuer     one elected lane issues tcgen05.mma (128/256 x BLOCK_N x 16) into a double-buffered
//                              TMEM accumulator; tcgen05.commit releases smem stages / publishes tiles
//
// Both operands may be K-major or MN-major (UMMA descriptor + instruction-descriptor major bits), so the
// same kernel serves forward (x W^T), dgrad (dy W) and wgrad (dy^T x) without materialised transposes.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "philox.cuh"
#include "ptx.cuh"

namespace vq {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;   // 64 bf16 = 128 B = one swizzle-128B row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 384;     // 8 epilogue warps (0..7) + TMA warp (8) + MMA warp (9) + 2 idle warps: setmaxnreg
                                      // is a warpgroup-wide instruction, so the third warpgroup is complete
constexpr int kTmaWarp = 8;
constexpr int kMmaWarp = 9;
constexpr int kRopeSmemPos = 64;       // rope table positions staged in smem (larger tables are read from global)
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KB

struct GemmArgs {
  int M, N, K;
  int a_mn, b_mn;
  int Z1;            // inner batch extent (z = z0 * Z1 + z1)
  int m_tiles, n_tiles, total_tiles;
  int a_r0, a_c0, a_r1, a_c1;
  int b_r0, b_c0, b_r1, b_c1;
  long long c_r0, c_c0, c_r1, c_c1;
  void* C;
  long long ldc;
  float alpha;
  const float* bias;
  const float* gamma;
  const void* add;
  long long ldadd;
  void* aux;
  long long ldaux;
  // QKV_ROPE / PATCH
  int tokens_per_frame, patch_start, grid_w, apply_qknorm_rope;
  const float *qn_w, *qn_b, *kn_w, *kn_b;
  float qk_eps, q_scale;
  const float *rope_cos, *rope_sin;
  int rope_positions;  // rows of the rope tables (max_pos + 1), 0 if unknown
  int n_patches;
  const float* pos_embed;
  DropoutParams drop;  // thr == 0: off
  // LayerNorm folding (see vq_gemm_desc): consumer side (A holds the RAW rows, the norm is applied in the epilogue) ...
  const float* ln_stats;   // [M][2] (sum x, sum x^2) of the rows of A, or null
  const float* ln_colsum;  // [N] row sums of the (gamma-folded) weight
  float ln_eps, ln_inv_dim;
  // ... and producer side (LINEAR f32 epilogue): row statistics and a bf16 copy of the rows it writes
  float* stats_out;        // [M][2], accumulated with atomics, or null
  float* stats_zero;       // [M][2] cleared by this launch after its grid dependency resolved, or null
  void* c_copy;            // bf16 [M, ldc] copy of C, or null
  // Tail split (see launch_gemm_cg): tiles [0, full_tiles) are processed whole, round-robin; each of the remaining
  // tiles is split along K into `tail_split` slices handled by different CTAs (pairs) in the last round.
  int full_tiles, tail_split;
  float* splitk_ws;         // partial accumulators of the slices > 0
  unsigned* splitk_flags;   // per tail tile: [arrivals of the partial writers, owner warps done]
};

// One unit of work of a persistent CTA (pair): a whole tile or one K slice of a tail tile.
struct GemmWork {
  int tile, kb0, kb1, slice;   // slice < 0: whole tile
};
__device__ __forceinline__ bool gemm_work(const GemmArgs& g, int it, int first, int stride, int num_kb, GemmWork& w) {
  const int n_full_iters = g.full_tiles / stride;
  if (it < n_full_iters) {
    w.tile = first + it * stride; w.kb0 = 0; w.kb1 = num_kb; w.slice = -1;
    return true;
  }
  const int n_tail = g.total_tiles - g.full_tiles;
  if (it > n_full_iters || first >= n_tail * g.tail_split) return false;
  w.tile = g.full_tiles + first / g.tail_split;
  if (g.tail_split == 1) { w.kb0 = 0; w.kb1 = num_kb; w.slice = -1; return true; }
  w.slice = first % g.tail_split;
  w.kb0 = w.slice * num_kb / g.tail_split;
  w.kb1 = (w.slice + 1) * num_kb / g.tail_split;
  return true;
}

// CG = CTAs per MMA (cta_group): 1 = one CTA owns a 128 x BLOCK_N tile; 2 = a CTA pair owns a 256 x BLOCK_N tile,
// each CTA loading its 128 rows of A and BLOCK_N/2 rows of B.
constexpr int kStageLd = 36;                          // floats per staged row: 32 columns + 4 pad (conflict-free 128-bit rows)
constexpr int kStageWarpFloats = 32 * kStageLd;       // one 32 x 32 fp32 transposition buffer per epilogue warp
constexpr int kSmemLimit = 232448;                    // 227 KB per CTA

template <int BLOCK_N, int CG>
struct GemmCfg {
  static constexpr int kBTileBytes = (BLOCK_N / CG) * kBlockK * 2;
  static constexpr int kStageBytes = kATileBytes + kBTileBytes;
  static constexpr int kTmemCols = 2 * BLOCK_N;  // two accumulator stages (power of two: 256 / 512)
  // per-warp [bias|gamma] vectors (8 warps x BLOCK_N floats), then EITHER the QKV extras (q/k-norm affine 4 x 64 and
  // the rope cos/sin tables) OR the per-warp transposition buffers of the other epilogues
  static constexpr int kQkvStageWarpBytes = 32 * 128;   // one 32-row x 64-column bf16 head block per epilogue warp
  static constexpr int kQkvExtraBytes = 4 * 64 * 4 + 2 * kRopeSmemPos * 16 * 4 + 8 * kQkvStageWarpBytes;
  static constexpr int kTransposeBytes = 8 * kStageWarpFloats * 4;
  static constexpr int kEpiVecBytes =
      8 * BLOCK_N * 4 + (kQkvExtraBytes > kTransposeBytes ? kQkvExtraBytes : kTransposeBytes);
  static constexpr int kFixedBytes = 1024 /*align slack*/ + 256 /*barriers*/ + kEpiVecBytes;
  static constexpr int kStagesFit = (kSmemLimit - kFixedBytes) / kStageBytes;
  static constexpr int kStages = kStagesFit > 6 ? 6 : kStagesFit;
  static constexpr int kSmemBytes = kStages * kStageBytes + kFixedBytes;
  static_assert(kStages >= 3, "smem ring too shallow");
};

// ---- erf without erff() (used by the GELU derivative): Phi(x) = 0.5 (1 + E(t)), t = clamp(x, -4.2, 4.2), E(t) = t P(t^2)
// with P a degree-8 minimax polynomial constrained to E(4.2) = 1.  Max abs error of x Phi(x) against the erf
// definition: 5.9e-5.  Evaluated two columns at a time with packed fp32 FMAs (FFMA2).
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{.reg .b64 ra, rb, rc, rd;\n"
      "mov.b64 ra, {%2, %3};\n"
      "mov.b64 rb, {%4, %5};\n"
      "mov.b64 rc, {%6, %7};\n"
      "fma.rn.f32x2 rd, ra, rb, rc;\n"
      "mov.b64 {%0, %1}, rd;}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float clamp_sym(float x, float c) {  // sign(x) * min(|x|, c) in one FMNMX
  float t;
  asm("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(t) : "f"(x), "f"(c));
  return t;
}
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }
// E(t) = erf(t / sqrt(2)) for |t| <= 4.2, both lanes
__device__ __forceinline__ float2 erf_sqrt2_poly2(float2 t) {
  const float2 u = ffma2(t, t, splat2(0.f));
  float2 p = splat2(9.847654636e-11f);
  p = ffma2(p, u, splat2(-9.591337502e-09f));
  p = ffma2(p, u, splat2(4.144957018e-07f));
  p = ffma2(p, u, splat2(-1.057913687e-05f));
  p = ffma2(p, u, splat2(1.796719589e-04f));
  p = ffma2(p, u, splat2(-2.172918525e-03f));
  p = ffma2(p, u, splat2(1.948188245e-02f));
  p = ffma2(p, u, splat2(-1.324441582e-01f));
  p = ffma2(p, u, splat2(7.976739407e-01f));
  return ffma2(p, t, splat2(0.f));
}
// Forward GELU (exact erf form) through the complementary normal CDF, whose logarithm is nearly polynomial:
//   gelu(x) = x Phi(x) = max(x, 0) - |x| Q(|x|),   Q(a) = 1 - Phi(a) = 0.5 erfc(a / sqrt(2)) ~= 2^p(a),
//   p(a) = -1 + a (c1 + a (c2 + a (c3 + a c4)))   (minimax fit of a Q(a) on [0, 9], p(0) = -1 exactly so that the slope at
//   the origin is exact).  Max abs error of gelu against the erf definition: 1.3e-5 (bf16 rounding of the result is 2e-3
//   relative).  4 FMAs + one MUFU.EX2 + 3 cheap ops per element instead of a degree-8 polynomial: the epilogue of the
//   K = 1024 fc1 GEMM is issue-bound, every instruction per element shows up in its run time.
__device__ __forceinline__ float gelu1(float x) {
  const float a = fminf(fabsf(x), 8.0f);        // 2^p(8) = 2^-47: the correction has vanished long before
  float p = fmaf(a, 0.0038706841245f, -0.044546267023f);
  p = fmaf(a, p, -0.46639128587f);
  p = fmaf(a, p, -1.14905086599f);
  p = fmaf(a, p, -1.0f);
  return fmaf(-fabsf(x), ex2_approx(p), fmaxf(x, 0.f));
}
#ifdef VQ_GELU_POLY   // A/B builds: the degree-8 erf polynomial form
__device__ __forceinline__ float2 gelu2(float2 x) {
  const float2 t = make_float2(clamp_sym(x.x, 4.2f), clamp_sym(x.y, 4.2f));
  const float2 e = erf_sqrt2_poly2(t);
  const float2 hx = ffma2(x, splat2(0.5f), splat2(0.f));
  return ffma2(hx, e, hx);
}
#else
__device__ __forceinline__ float2 gelu2(float2 x) { return make_float2(gelu1(x.x), gelu1(x.y)); }
#endif
// gelu'(x) = Phi(x) + x phi(x)
__device__ __forceinline__ float2 gelu_grad2(float2 x) {
  const float2 t = make_float2(clamp_sym(x.x, 4.2f), clamp_sym(x.y, 4.2f));
  const float2 e = erf_sqrt2_poly2(t);
  const float2 cdf = ffma2(e, splat2(0.5f), splat2(0.5f));
  const float2 a = ffma2(x, x, splat2(0.f));
  // phi(x) = exp(-x^2/2) / sqrt(2 pi) = 0.39894228 * 2^(-0.72134752 x^2)
  const float2 pdf = make_float2(0.39894228040143268f * ex2_approx(-0.72134752044448170f * a.x),
                                 0.39894228040143268f * ex2_approx(-0.72134752044448170f * a.y));
  return ffma2(x, pdf, cdf);
}

// 8 consecutive elements (QKV epilogue: bf16 only)
__device__ __forceinline__ void store8_bf16(void* base, long long off, const float* v) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]);
  u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]);
  u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + off) = u;
}
// One 16-byte access: 4 f32 or 8 bf16 consecutive elements.
template <bool F32>
__device__ __forceinline__ void storev(void* base, long long off, const float* v) {
  if constexpr (F32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + off) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    store8_bf16(base, off, v);
  }
}
template <bool F32>
__device__ __forceinline__ void loadv(const void* base, long long off, float* v) {
  if constexpr (F32) {
    const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  } else {
    const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + off);
    v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x);
    v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
    v[4] = bf16_lo(u.z); v[5] = bf16_hi(u.z);
    v[6] = bf16_lo(u.w); v[7] = bf16_hi(u.w);
  }
}

// MODE: VQ_EPI_* ; OUT_F32: C/add dtype ; EXL: which of the rarely used epilogue extras are compiled in -- 0 none (what the
// aggregator runs), 1 train-mode dropout only (the Perceiver's four sites), 2 everything (LayerNorm folding, row
// statistics / bf16 copy, K-sliced tail tiles, dropout).  Separate instantiations because the unrolled epilogue is
// instruction-cache bound: with everything compiled into one kernel the f32-residual instance was 173 KB of SASS and
// ncu's top stall after `barrier` was `no_instruction` (profiles/r02_gemm_fc1_gelu.txt).
template <int BLOCK_N, int MODE, bool OUT_F32, int CG, int EXL>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const GemmArgs g) {
  using Cfg = GemmCfg<BLOCK_N, CG>;
  constexpr bool EX = EXL == 2;      // LayerNorm folding, statistics, tail split
  constexpr bool DROPX = EXL >= 1;   // dropout
  constexpr int kStages = Cfg::kStages;
  constexpr int kTileM = kBlockM * CG;                       // rows of C per (pair-)tile
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0;  // 0 = leader (issues the MMAs)
  const int first_tile = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_stride = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + kStages * kATileBytes;
  const uint32_t bar_base = smem_base + kStages * Cfg::kStageBytes;
  // barrier layout (8 B each): full[kStages] | empty[kStages] | tmem_full[2] | tmem_empty[2] | tmem ptr
  const uint32_t bar_full = bar_base;
  const uint32_t bar_empty = bar_base + 8 * kStages;
  const uint32_t bar_tfull = bar_base + 16 * kStages;
  const uint32_t bar_tempty = bar_tfull + 16;
  const uint32_t tmem_slot = bar_tempty + 16;
  uint32_t* tmem_slot_ptr =
      reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  float* epi_vec = reinterpret_cast<float*>(smem_raw + (bar_base + 256 - smem_u32(smem_raw)));
  float* qk_vec = epi_vec + 8 * BLOCK_N;        // [q_w | q_b | k_w | k_b] x 64
  float* rope_vec = qk_vec + 4 * 64;            // [cos | sin] x kRopeSmemPos x 16
  float* stage_vec = epi_vec + 8 * BLOCK_N;     // aliases the QKV extras: 8 x [32][kStageLd] f32 transposition buffers
  uint8_t* qkv_stage = reinterpret_cast<uint8_t*>(rope_vec + 2 * kRopeSmemPos * 16);   // 8 x 4 KB, after the rope tables
  const bool rope_in_smem = g.rope_positions > 0 && g.rope_positions <= kRopeSmemPos;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  pdl_launch_dependents();   // the next kernel's prologue may overlap this grid's tail
  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tempty + 8 * s, 8 * CG);  // one elected lane per epilogue warp of the (pair of) CTA(s)
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    if constexpr (CG == 2) {
      tmem_alloc_pair(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish();
    }
  }
  if constexpr (MODE == VQ_EPI_QKV_ROPE) {
    if (g.apply_qknorm_rope) {
      for (int i = threadIdx.x; i < 64; i += blockDim.x) {
        qk_vec[i] = g.qn_w[i]; qk_vec[64 + i] = g.qn_b[i];
        qk_vec[128 + i] = g.kn_w[i]; qk_vec[192 + i] = g.kn_b[i];
      }
      if (rope_in_smem) {
        for (int i = threadIdx.x; i < g.rope_positions * 16; i += blockDim.x) {
          rope_vec[i] = g.rope_cos[i];
          rope_vec[kRopeSmemPos * 16 + i] = g.rope_sin[i];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();  // peer barriers are initialised before any remote arrive / TMA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // everything above touched only smem, TMEM and weights; A (and the residual / C) come from the previous kernels
  pdl_wait();

  const int num_kb = (g.K + kBlockK - 1) / kBlockK;
  const int tiles_per_z = g.m_tiles * g.n_tiles;

  // Register budget: 384 x 168 at launch; warps 8..11 shrink to 56 so that the 8 epilogue warps can grow to 224
  // (8 x 32 x 224 + 4 x 32 x 56 = 64512).
  if (warp >= kTmaWarp) setmaxnreg_dec<56>();
  if (warp == kTmaWarp) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto load = [&](uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
        if constexpr (CG == 2) tma_load_2d_pair(dst, map, bar, c0, c1);
        else tma_load_2d(dst, map, bar, c0, c1);
      };
      constexpr int kBRows = BLOCK_N / CG;  // rows (N indices) of B this CTA loads
      GemmWork wk;
      for (int wi = 0; gemm_work(g, wi, first_tile, tile_stride, num_kb, wk); ++wi) {
        const int tile = wk.tile;
        const int z = tile / tiles_per_z;
        const int rem = tile - z * tiles_per_z;
        const int m_blk = rem / g.n_tiles;
        const int n_blk = rem - m_blk * g.n_tiles;
        const int z0 = z / g.Z1, z1 = z - z0 * g.Z1;
        const int a_r = z0 * g.a_r0 + z1 * g.a_r1, a_c = z0 * g.a_c0 + z1 * g.a_c1;
        const int b_r = z0 * g.b_r0 + z1 * g.b_r1, b_c = z0 * g.b_c0 + z1 * g.b_c1;
        const int m0 = m_blk * kTileM + (int)cta_rank * kBlockM;       // first A row of this CTA
        const int n0 = n_blk * BLOCK_N + (int)cta_rank * kBRows;       // first B row of this CTA
        for (int kb = wk.kb0; kb < wk.kb1; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t full = bar_full + 8 * stage;
          // the leader's barrier collects the bytes of both CTAs
          if (cta_rank == 0) mbar_arrive_expect_tx(full, Cfg::kStageBytes * CG);
          const uint32_t sa = smem_a + stage * kATileBytes;
          const uint32_t sb = smem_b + stage * Cfg::kBTileBytes;
          if (!g.a_mn) {
            load(sa, &tmA, full, a_c + kb * kBlockK, a_r + m0);
          } else {
#pragma unroll
            for (int i = 0; i < kBlockM / 64; ++i)
              load(sa + i * (kBlockK * 128), &tmA, full, a_c + m0 + i * 64, a_r + kb * kBlockK);
          }
          if (!g.b_mn) {
            load(sb, &tmB, full, b_c + kb * kBlockK, b_r + n0);
          } else {
#pragma unroll
            for (int i = 0; i < kBRows / 64; ++i)
              load(sb + i * (kBlockK * 128), &tmB, full, b_c + n0 + i * 64, b_r + kb * kBlockK);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    // The whole warp runs the control flow so that waits and descriptor arithmetic stay warp-uniform; only the
    // tcgen05 instructions are predicated on one elected lane.  (Highest warp id: favoured by the scheduler.)
    if (cta_rank == 0) {
      const uint32_t idesc = make_idesc_bf16(kTileM, BLOCK_N, g.a_mn, g.b_mn);
      // K-major: 8-row groups 1024 B apart, k-step = 32 B inside the 128 B swizzle row.
      // MN-major: 8-k groups 1024 B apart, 64-wide M/N chunks kBlockK*128 B apart, k-step = 16 rows.
      const uint32_t a_lbo = g.a_mn ? kBlockK * 128 : 16, b_lbo = g.b_mn ? kBlockK * 128 : 16;
      // descriptor start-address field is in 16-byte units
      const uint32_t a_kstep = (g.a_mn ? (kUmmaK * 128) : (kUmmaK * 2)) >> 4;
      const uint32_t b_kstep = (g.b_mn ? (kUmmaK * 128) : (kUmmaK * 2)) >> 4;
      const uint64_t desc_a0 = make_smem_desc(smem_a, a_lbo, 1024);
      const uint64_t desc_b0 = make_smem_desc(smem_b, b_lbo, 1024);
      int stage = 0;
      uint32_t phase = 0;
      GemmWork wk;
      for (int it = 0; gemm_work(g, it, first_tile, tile_stride, num_kb, wk); ++it) {
        const int as = it & 1;
        const uint32_t as_phase = (it >> 1) & 1;
        mbar_wait(bar_tempty + 8 * as, as_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        for (int kb = wk.kb0; kb < wk.kb1; ++kb) {
          mbar_wait_fast(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint64_t desc_a = desc_a0 + (uint64_t)(stage * (kATileBytes >> 4));
          const uint64_t desc_b = desc_b0 + (uint64_t)(stage * (Cfg::kBTileBytes >> 4));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              const uint64_t adesc = desc_a + (uint64_t)(k * a_kstep);
              const uint64_t bdesc = desc_b + (uint64_t)(k * b_kstep);
              if constexpr (CG == 2) umma_ss_pair(d_tmem, adesc, bdesc, idesc, ((kb - wk.kb0) | k) ? 1u : 0u);
              else umma_ss(d_tmem, adesc, bdesc, idesc, ((kb - wk.kb0) | k) ? 1u : 0u);
            }
            // frees the smem stage (in both CTAs of a pair) when these MMAs retire
            if constexpr (CG == 2) umma_commit_pair(bar_empty + 8 * stage);
            else umma_commit(bar_empty + 8 * stage);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        // accumulator complete (published to both CTAs of a pair)
        if (elect_one()) {
          if constexpr (CG == 2) umma_commit_pair(bar_tfull + 8 * as);
          else umma_commit(bar_tfull + 8 * as);
        }
        __syncwarp();
      }
    }
  } else if (warp < kTmaWarp) {
    // ===================== epilogue (warps 0..7) =====================
    // Two warps share each TMEM lane quarter (hardware: warp w may touch lanes 32*(w%4)..+31) and split the
    // tile's columns in halves; one accumulator row per thread.  Per-column vectors (bias, LayerScale gamma)
    // are staged per warp in smem before the accumulator is waited for, TMEM loads are double-buffered and all
    // global loads of a 32-column chunk are issued before its math, so no latency sits on the critical path.
    setmaxnreg_inc<224>();
    const int quarter = warp & 3;
    const int half = warp >> 2;
    constexpr int kHalfN = BLOCK_N / 2;
    const int row_in_tile = quarter * 32 + lane;
    float* my_vec = epi_vec + warp * (2 * kHalfN);  // [bias | gamma] x kHalfN, private to this warp
    // The residual (`add`) block of a tile is known long before its accumulator: pull it into L2 one tile ahead so
    // that the epilogue's loads are L2 hits instead of HBM round trips (the f32 residual epilogue is latency-bound).
    auto prefetch_add = [&](int t) {
      if (MODE == VQ_EPI_QKV_ROPE || !g.add || t >= g.total_tiles) return;
      const int z = t / tiles_per_z;
      const int rem = t - z * tiles_per_z;
      const int m_blk = rem / g.n_tiles;
      const int n_blk = rem - m_blk * g.n_tiles;
      const int z0 = z / g.Z1, z1 = z - z0 * g.Z1;
      const long long c_r = z0 * g.c_r0 + z1 * g.c_r1, c_c = z0 * g.c_c0 + z1 * g.c_c1;
      const int m0 = m_blk * kTileM + (int)cta_rank * kBlockM + quarter * 32;
      const int n0 = n_blk * BLOCK_N + half * kHalfN;
      constexpr int kLineElts = OUT_F32 ? 32 : 64;   // elements per 128-byte line
      constexpr int kLinesPerRow = kHalfN / kLineElts;
      const char* base = reinterpret_cast<const char*>(g.add);
#pragma unroll
      for (int i = lane; i < 32 * kLinesPerRow; i += 32) {
        const int m = m0 + i / kLinesPerRow, n = n0 + (i % kLinesPerRow) * kLineElts;
        if (m < g.M && n < g.N) prefetch_l2(base + ((c_r + m) * g.ldadd + c_c + n) * (OUT_F32 ? 4 : 2));
      }
    };
    {
      GemmWork w0;
      if (gemm_work(g, 0, first_tile, tile_stride, num_kb, w0) && w0.slice <= 0) prefetch_add(w0.tile);
    }
    if constexpr (EX && MODE == VQ_EPI_LINEAR && OUT_F32) {
      // the statistics buffer of the NEXT producer: its last readers finished before this grid's dependency resolved
      if (g.stats_zero) {
        const int n2 = 2 * g.M;
        for (int i = blockIdx.x * 256 + threadIdx.x; i < n2; i += gridDim.x * 256) g.stats_zero[i] = 0.f;
      }
    }
    GemmWork wk, wk_next;
    for (int it = 0; gemm_work(g, it, first_tile, tile_stride, num_kb, wk); ++it) {
      const int tile = wk.tile;
      if (gemm_work(g, it + 1, first_tile, tile_stride, num_kb, wk_next) && wk_next.slice <= 0) prefetch_add(wk_next.tile);
      const int z = tile / tiles_per_z;
      const int rem = tile - z * tiles_per_z;
      const int m_blk = rem / g.n_tiles;
      const int n_blk = rem - m_blk * g.n_tiles;
      const int z0 = z / g.Z1, z1 = z - z0 * g.Z1;
      const long long c_r = z0 * g.c_r0 + z1 * g.c_r1, c_c = z0 * g.c_c0 + z1 * g.c_c1;
      const int as = it & 1;
      const uint32_t as_phase = (it >> 1) & 1;
      const int n_base = n_blk * BLOCK_N + half * kHalfN;  // first column this warp owns

      // stage bias / gamma for this warp's columns (overlaps the wait for the accumulator)
      __syncwarp();
      for (int c = lane * 4; c < kHalfN; c += 128) {
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f), gm = make_float4(1.f, 1.f, 1.f, 1.f);
        if (n_base + c < g.N) {
          if (g.bias) b = __ldg(reinterpret_cast<const float4*>(g.bias + n_base + c));
          if (g.gamma) gm = __ldg(reinterpret_cast<const float4*>(g.gamma + n_base + c));
          else if (g.ln_colsum) gm = __ldg(reinterpret_cast<const float4*>(g.ln_colsum + n_base + c));
        }
        *reinterpret_cast<float4*>(my_vec + c) = b;
        *reinterpret_cast<float4*>(my_vec + kHalfN + c) = gm;
      }
      __syncwarp();

      mbar_wait(bar_tfull + 8 * as, as_phase);
      tc_fence_after();

      const int m = m_blk * kTileM + (int)cta_rank * kBlockM + row_in_tile;
      const bool row_ok = m < g.M;
      const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + as * BLOCK_N + half * kHalfN;

      if constexpr (MODE == VQ_EPI_QKV_ROPE) {
        // One head (64 columns) at a time; the whole head row lives in this thread's registers.
        static_assert(kHalfN % 64 == 0, "QKV epilogue: whole heads per warp");
        const int D3 = g.N / 3;  // q | k | v column blocks
#pragma unroll 1
        for (int hh = 0; hh < kHalfN / 64; ++hh) {
          const int n_head = n_base + hh * 64;
          if (n_head >= g.N) break;
          float v[64];
          {
            // LayerNorm folded into this GEMM: A held the raw rows x, W the gamma-scaled weight W', so
            // LN(x) W^T + b = rstd (x W'^T - mean colsum(W')) + (W beta + b); the last term arrives as the bias
            float ln_mean = 0.f, ln_rstd = 1.f;
            if (EX && g.ln_stats && row_ok) {
              const float2 st = *reinterpret_cast<const float2*>(g.ln_stats + 2 * (long long)m);
              ln_mean = st.x * g.ln_inv_dim;
              ln_rstd = rsqrtf(fmaxf(st.y * g.ln_inv_dim - ln_mean * ln_mean, 0.f) + g.ln_eps);
            }
            const float* cs = my_vec + kHalfN + hh * 64;
            uint32_t r[32];
            tmem_ld32(taddr + hh * 64, r);
            tmem_wait_ld();
            if (EX && g.ln_stats) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                v[i] = (__uint_as_float(r[i]) - ln_mean * cs[i]) * ln_rstd + my_vec[hh * 64 + i];
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) + my_vec[hh * 64 + i];
            }
            tmem_ld32(taddr + hh * 64 + 32, r);
            tmem_wait_ld();
            if (EX && g.ln_stats) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                v[32 + i] = (__uint_as_float(r[i]) - ln_mean * cs[32 + i]) * ln_rstd + my_vec[hh * 64 + 32 + i];
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[32 + i] = __uint_as_float(r[i]) + my_vec[hh * 64 + 32 + i];
            }
          }
          const int which = n_head / D3;  // 0 = q, 1 = k, 2 = v
          if (which < 2 && g.apply_qknorm_rope) {
            const float* w = qk_vec + which * 128;  // [w | b] of q_norm or k_norm
            float mean = 0.f;
#pragma unroll
            for (int i = 0; i < 64; ++i) mean += v[i];
            mean *= (1.0f / 64.0f);
            float var = 0.f;
#pragma unroll
            for (int i = 0; i < 64; ++i) {
              const float d = v[i] - mean;
              var += d * d;
            }
            const float rstd = rsqrtf(var * (1.0f / 64.0f) + g.qk_eps);
#pragma unroll
            for (int i = 0; i < 64; ++i) v[i] = (v[i] - mean) * rstd * w[i] + w[64 + i];
            // 2-D RoPE: dims [0,32) rotate with the row (y) position, [32,64) with the column (x).
            int pos_y = 0, pos_x = 0;
            const int p = m % g.tokens_per_frame;
            if (p >= g.patch_start) {
              const int idx = p - g.patch_start;
              pos_y = idx / g.grid_w + 1;
              pos_x = idx % g.grid_w + 1;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int pos = h == 0 ? pos_y : pos_x;
              float co[16], si[16];
              if (rope_in_smem) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                  const float4 c4 = *reinterpret_cast<const float4*>(rope_vec + pos * 16 + j);
                  const float4 s4 = *reinterpret_cast<const float4*>(rope_vec + kRopeSmemPos * 16 + pos * 16 + j);
                  co[j] = c4.x; co[j + 1] = c4.y; co[j + 2] = c4.z; co[j + 3] = c4.w;
                  si[j] = s4.x; si[j + 1] = s4.y; si[j + 2] = s4.z; si[j + 3] = s4.w;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  co[j] = __ldg(g.rope_cos + pos * 16 + j);
                  si[j] = __ldg(g.rope_sin + pos * 16 + j);
                }
              }
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float a = v[32 * h + j], b2 = v[32 * h + 16 + j];
                v[32 * h + j] = a * co[j] - b2 * si[j];
                v[32 * h + 16 + j] = b2 * co[j] + a * si[j];
              }
            }
          }
          if (which == 0) {
#pragma unroll
            for (int i = 0; i < 64; ++i) v[i] *= g.q_scale;
          }
          // One row per thread would make every store instruction touch 32 different lines (16 B each): go through
          // a per-warp smem block (16-byte chunks XOR-swizzled by row, conflict-free both ways) and store 4 whole
          // 128-byte row segments per instruction instead.
          uint8_t* st = qkv_stage + warp * Cfg::kQkvStageWarpBytes;
          __syncwarp();   // the previous head's read-back is complete
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint4 u;
            u.x = pack_bf16x2(v[8 * c + 0], v[8 * c + 1]);
            u.y = pack_bf16x2(v[8 * c + 2], v[8 * c + 3]);
            u.z = pack_bf16x2(v[8 * c + 4], v[8 * c + 5]);
            u.w = pack_bf16x2(v[8 * c + 6], v[8 * c + 7]);
            *reinterpret_cast<uint4*>(st + lane * 128 + ((c ^ (lane & 7)) << 4)) = u;
          }
          __syncwarp();
          {
            const int cc = lane & 7, rsub = lane >> 3;
            const int m_warp0 = m_blk * kTileM + (int)cta_rank * kBlockM + quarter * 32;
#pragma unroll
            for (int ps = 0; ps < 8; ++ps) {
              const int r = ps * 4 + rsub;
              const uint4 u = *reinterpret_cast<const uint4*>(st + r * 128 + ((cc ^ (r & 7)) << 4));
              if (m_warp0 + r < g.M)
                *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(g.C) +
                                          (c_r + m_warp0 + r) * g.ldc + c_c + n_head + cc * 8) = u;
            }
          }
        }
      } else {
        // 32 columns at a time: the raw accumulators (one row per thread) go through a per-warp smem buffer and
        // come back transposed, kVec consecutive columns per thread and 32 / kVec threads per row, so that every
        // global access of the warp covers whole 128-byte (f32) / 64-byte (bf16) row segments instead of 32
        // different lines.  All epilogue math and every global load / store happens in that layout.
        constexpr int kVec = OUT_F32 ? 4 : 8;            // columns per thread = one 16-byte C access
        constexpr int kLanesPerRow = 32 / kVec;
        constexpr int kRowsPerPass = 32 / kLanesPerRow;
        constexpr int kPasses = 32 / kRowsPerPass;
        float* stage = stage_vec + warp * kStageWarpFloats;
        const int cb = (lane % kLanesPerRow) * kVec;     // this thread's first column inside the chunk
        const int rsub = lane / kLanesPerRow;
        const int m_warp0 = m_blk * kTileM + (int)cta_rank * kBlockM + quarter * 32;
        constexpr int kChunks = kHalfN / 32;
        // LayerNorm folding, consumer side: per-row mean / rstd of the rows this thread touches
        float ln_mean[kPasses], ln_rstd[kPasses];
        if constexpr (EX && (MODE == VQ_EPI_GELU || MODE == VQ_EPI_LINEAR)) {
          if (g.ln_stats) {
#pragma unroll
            for (int ps = 0; ps < kPasses; ++ps) {
              const int mr = m_warp0 + ps * kRowsPerPass + rsub;
              float2 st = make_float2(0.f, 1.f);
              if (mr < g.M) st = *reinterpret_cast<const float2*>(g.ln_stats + 2 * (long long)mr);
              ln_mean[ps] = st.x * g.ln_inv_dim;
              ln_rstd[ps] = rsqrtf(fmaxf(st.y * g.ln_inv_dim - ln_mean[ps] * ln_mean[ps], 0.f) + g.ln_eps);
            }
          }
        }
        // producer side: running (sum, sum of squares) of the rows this thread writes, over this warp's columns
        float rs[kPasses], rq[kPasses];
#pragma unroll
        for (int ps = 0; ps < kPasses; ++ps) rs[ps] = rq[ps] = 0.f;
        uint32_t r[32];   // the accumulator chunk in flight: dead once it is staged in smem, so one buffer suffices
        // K-sliced tail tile: slices > 0 park their raw accumulators in the workspace (thread-private layout: the owner's
        // thread at the same position reads them back; lane-interleaved 16-byte granules), slice 0 waits for them and adds them in slice order before its
        // epilogue -- a fixed summation order, so the result does not depend on timing.
        const int tail_idx = tile - g.full_tiles;
        constexpr int kWarpWs = kChunks * 1024;   // floats per warp and partial tile
        if (EX && wk.slice > 0) {
          float* wsp = g.splitk_ws +
                       ((((size_t)tail_idx * (g.tail_split - 1) + (wk.slice - 1)) * CG + cta_rank) * 8 + warp) * kWarpWs;
#pragma unroll 1
          for (int ch = 0; ch < kChunks; ++ch) {
            if (n_base + ch * 32 >= g.N) break;
            tmem_ld32(taddr + ch * 32, r);
            tmem_wait_ld();
#pragma unroll
            for (int k = 0; k < 8; ++k)
              *reinterpret_cast<uint4*>(wsp + ((ch * 8 + k) * 32 + lane) * 4) =   // lane-interleaved: 512 B per request
                  make_uint4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
          }
          __threadfence();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            atomicAdd(g.splitk_flags + 2 * tail_idx, 1u);
            if constexpr (CG == 2) mbar_arrive_leader(bar_tempty + 8 * as);
            else mbar_arrive(bar_tempty + 8 * as);
          }
          continue;
        }
        const bool owner = EX && wk.slice == 0;
        if (owner) {
          if (lane == 0) {
            const unsigned want = (unsigned)(g.tail_split - 1) * CG * 8;
            unsigned spins = 0;
            while (*reinterpret_cast<volatile unsigned*>(g.splitk_flags + 2 * tail_idx) < want) {
              __nanosleep(64);
              if (++spins > (1u << 24)) __trap();   // a partial writer never arrived
            }
            __threadfence();
          }
          __syncwarp();
        }
        if (n_base < g.N) tmem_ld32(taddr, r);
#pragma unroll 1   // ONE copy of the body: the fully unrolled epilogue did not fit the 32 KB L1.5 instruction cache
        for (int ch = 0; ch < kChunks; ++ch) {
          const int n0 = n_base + ch * 32;
          if (n0 >= g.N) break;  // warp-uniform
          tmem_wait_ld();
          if (owner) {
            for (int sl = 1; sl < g.tail_split; ++sl) {
              const float* wsp = g.splitk_ws +
                                 ((((size_t)tail_idx * (g.tail_split - 1) + (sl - 1)) * CG + cta_rank) * 8 + warp) * kWarpWs +
                                 ch * 1024 + lane * 4;
              float4 pvs[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) pvs[k] = __ldcg(reinterpret_cast<const float4*>(wsp + k * 128));
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const float4 pv = pvs[k];
                r[4 * k] = __float_as_uint(__uint_as_float(r[4 * k]) + pv.x);
                r[4 * k + 1] = __float_as_uint(__uint_as_float(r[4 * k + 1]) + pv.y);
                r[4 * k + 2] = __float_as_uint(__uint_as_float(r[4 * k + 2]) + pv.z);
                r[4 * k + 3] = __float_as_uint(__uint_as_float(r[4 * k + 3]) + pv.w);
              }
            }
          }
          __syncwarp();  // previous chunk's transposed reads are done
#pragma unroll
          for (int k = 0; k < 8; ++k)
            *reinterpret_cast<uint4*>(stage + lane * kStageLd + 4 * k) =
                make_uint4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
          __syncwarp();
          // the registers are free again: the next chunk's TMEM load overlaps this chunk's math and global traffic
          if (ch + 1 < kChunks && n0 + 32 < g.N) tmem_ld32(taddr + (ch + 1) * 32, r);
          const int ncol = n0 + cb;
          const bool col_ok = ncol < g.N;  // N % 8 == 0: a thread's kVec columns are all in or all out
          float bv[kVec], gv[kVec];
#pragma unroll
          for (int e = 0; e < kVec; e += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(my_vec + ch * 32 + cb + e);
            const float4 g4 = *reinterpret_cast<const float4*>(my_vec + kHalfN + ch * 32 + cb + e);
            bv[e] = b4.x; bv[e + 1] = b4.y; bv[e + 2] = b4.z; bv[e + 3] = b4.w;
            gv[e] = g4.x; gv[e + 1] = g4.y; gv[e + 2] = g4.z; gv[e + 3] = g4.w;
          }
          // ---- every global load of the chunk first
          float addv[kPasses][kVec];
          float auxv[kPasses][kVec];
          long long orow[kPasses];
          bool ok[kPasses];
#pragma unroll
          for (int ps = 0; ps < kPasses; ++ps) {
            const int m = m_warp0 + ps * kRowsPerPass + rsub;
            ok[ps] = col_ok && m < g.M;
            const long long io_row = c_r + m;
            orow[ps] = io_row;
            if constexpr (MODE == VQ_EPI_PATCH) {
              const int f = m / g.n_patches;
              const int idx = m - f * g.n_patches;
              orow[ps] = (long long)f * g.tokens_per_frame + g.patch_start + idx;
              if (ok[ps]) loadv<true>(g.pos_embed, (long long)(1 + idx) * g.N + ncol, auxv[ps]);
            }
            if (g.add && ok[ps]) loadv<OUT_F32>(g.add, io_row * g.ldadd + c_c + ncol, addv[ps]);
            if constexpr (MODE == VQ_EPI_GELU_BWD) {
              if (ok[ps]) loadv<false>(g.aux, io_row * g.ldaux + c_c + ncol, auxv[ps]);
            }
          }
          // ---- math + store
#pragma unroll
          for (int ps = 0; ps < kPasses; ++ps) {
            float v[kVec];
            const float* srow = stage + (ps * kRowsPerPass + rsub) * kStageLd + cb;
#pragma unroll
            for (int e = 0; e < kVec; e += 4) {
              const float4 a4 = *reinterpret_cast<const float4*>(srow + e);
              v[e] = a4.x * g.alpha + bv[e];
              v[e + 1] = a4.y * g.alpha + bv[e + 1];
              v[e + 2] = a4.z * g.alpha + bv[e + 2];
              v[e + 3] = a4.w * g.alpha + bv[e + 3];
            }
            if constexpr (EX && (MODE == VQ_EPI_GELU || MODE == VQ_EPI_LINEAR)) {
              if (g.ln_stats) {   // LN(x) W^T + b = rstd (x W'^T - mean colsum(W')) + b'   (gv = colsum, bv = b')
#pragma unroll
                for (int e = 0; e < kVec; ++e)
                  v[e] = (v[e] - bv[e] - ln_mean[ps] * gv[e]) * ln_rstd[ps] + bv[e];
              }
            }
            if (!ok[ps]) continue;
            if constexpr (MODE == VQ_EPI_GELU) {
              if (g.aux) storev<false>(g.aux, orow[ps] * g.ldaux + c_c + ncol, v);
#pragma unroll
              for (int e = 0; e < kVec; e += 2) {
                const float2 y = gelu2(make_float2(v[e], v[e + 1]));
                v[e] = y.x; v[e + 1] = y.y;
              }
            }
            if constexpr (MODE == VQ_EPI_GELU_BWD) {
#pragma unroll
              for (int e = 0; e < kVec; e += 2) {
                const float2 d = gelu_grad2(make_float2(auxv[ps][e], auxv[ps][e + 1]));
                v[e] *= d.x; v[e + 1] *= d.y;
              }
            }
            if constexpr (MODE == VQ_EPI_PATCH) {
#pragma unroll
              for (int e = 0; e < kVec; ++e) v[e] += auxv[ps][e];
            }
            if constexpr (DROPX && (MODE == VQ_EPI_LINEAR || MODE == VQ_EPI_GELU || MODE == VQ_EPI_GELU_BWD)) {
              if (g.drop.thr) {   // train-mode dropout: mask regenerated from (row, column), see philox.cuh
                const int m = m_warp0 + ps * kRowsPerPass + rsub;
                const uint32_t keep =
                    dropout_keep8(g.drop, (uint64_t)m * (uint64_t)(g.N >> 3) + (uint64_t)(ncol >> 3)) >> (ncol & 7);
#pragma unroll
                for (int e = 0; e < kVec; ++e) v[e] = ((keep >> e) & 1u) ? v[e] * g.drop.scale : 0.f;
              }
            }
            if (g.gamma) {
#pragma unroll
              for (int e = 0; e < kVec; ++e) v[e] *= gv[e];
            }
            if (g.add) {
#pragma unroll
              for (int e = 0; e < kVec; ++e) v[e] += addv[ps][e];
            }
            storev<OUT_F32>(g.C, orow[ps] * g.ldc + c_c + ncol, v);
            if constexpr (EX && MODE == VQ_EPI_LINEAR && OUT_F32) {
              if (g.c_copy) {      // bf16 copy of the rows: the A operand of the LayerNorm-folded GEMM that follows
                uint2 u;
                u.x = pack_bf16x2(v[0], v[1]);
                u.y = pack_bf16x2(v[2], v[3]);
                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(g.c_copy) + orow[ps] * g.ldc + c_c + ncol) = u;
              }
              if (g.stats_out) {
                rs[ps] += (v[0] + v[1]) + (v[2] + v[3]);
                rq[ps] += (v[0] * v[0] + v[1] * v[1]) + (v[2] * v[2] + v[3] * v[3]);
              }
            }
          }
        }
        if constexpr (EX && MODE == VQ_EPI_LINEAR && OUT_F32) {
          if (g.stats_out) {
            // the kLanesPerRow lanes that share a row combine, then one pair of atomics per row and warp
#pragma unroll
            for (int ps = 0; ps < kPasses; ++ps) {
              float a = rs[ps], b = rq[ps];
#pragma unroll
              for (int o = kLanesPerRow / 2; o > 0; o >>= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, o);
                b += __shfl_xor_sync(0xffffffffu, b, o);
              }
              const int mr = m_warp0 + ps * kRowsPerPass + rsub;
              if ((lane % kLanesPerRow) == 0 && mr < g.M && n_base < g.N) {
                atomicAdd(g.stats_out + 2 * (long long)mr, a);
                atomicAdd(g.stats_out + 2 * (long long)mr + 1, b);
              }
            }
          }
        }
      }
      if constexpr (EX && MODE != VQ_EPI_QKV_ROPE) {
        if (wk.slice == 0) {   // the last owner warp to finish re-arms the tile's flags for the next launch
          __syncwarp();
          if (lane == 0) {
            const int tail_idx = tile - g.full_tiles;
            if (atomicAdd(g.splitk_flags + 2 * tail_idx + 1, 1u) == (unsigned)(CG * 8 - 1)) {
              g.splitk_flags[2 * tail_idx] = 0u;
              g.splitk_flags[2 * tail_idx + 1] = 0u;
            }
          }
        }
      }
      // release the accumulator stage back to the (leader's) MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_leader(bar_tempty + 8 * as);
        else mbar_arrive(bar_tempty + 8 * as);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();  // neither CTA may exit (or free TMEM) while its peer still uses it
  if (warp == kMmaWarp) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// maximum number of K slices per tail tile (0 / 1 = no split); VQ_GEMM_TAIL_SPLIT or vq_gemm_tune.
// OFF by default -- measured on B200 (round 2, scripts/gemm_split_ab.py): the partially filled last round of the static
// schedule costs far less than a full round (the few busy CTA pairs run with less L2 / HBM contention and higher
// clocks under the power cap): splitting the tail mainloop 3 ways WITHOUT any fix-up leaves fc2 at M = 10 992
// unchanged (82.6 vs 82.2 us), and the fix-up adds 7-13 us.  The path stays as a tested, deterministic option.
static int g_gemm_tail_split = [] {
  const char* e = getenv("VQ_GEMM_TAIL_SPLIT");
  const int v = e ? atoi(e) : 0;
  return v < 0 ? 0 : v > 8 ? 8 : v;
}();

template <int BLOCK_N, int MODE, bool OUT_F32, int CG, int EXL>
static int launch_gemm_cg(const CUtensorMap* tmA, const CUtensorMap* tmB, GemmArgs& g, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N, CG>;
  constexpr bool EX = EXL == 2;
  auto kern = gemm_kernel<BLOCK_N, MODE, OUT_F32, CG, EXL>;
  VQ_CHECK_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(kern), Cfg::kSmemBytes));
  const int sms = num_sms();
  const int units = CG == 2 ? sms / 2 : sms;                     // CTAs (pairs) of a persistent grid
  int grid_units = g.total_tiles < units ? g.total_tiles : units;
  // Optional (see g_gemm_tail_split): the last round of a static round-robin schedule runs R = tiles mod units tiles
  // on `units` CTAs (pairs).  When R leaves at least half of them idle, every tail tile can be split along K into
  // floor(units / R) slices (>= 4 k-blocks each) that different CTAs work on concurrently; slice 0 adds the partial
  // accumulators of the others in slice order (deterministic) before its epilogue.
  g.full_tiles = g.total_tiles / units * units;
  g.tail_split = 1;
  g.splitk_ws = nullptr;
  g.splitk_flags = nullptr;
  const int n_tail = g.total_tiles - g.full_tiles;
  const int split_on = g_gemm_tail_split;
  if (EX && MODE != VQ_EPI_QKV_ROPE && split_on > 1 && n_tail > 0 && g.total_tiles == g.m_tiles * g.n_tiles) {
    const int num_kb = (g.K + kBlockK - 1) / kBlockK;
    int split = units / n_tail;
    if (split > split_on) split = split_on;
    while (split > 1 && num_kb / split < 4) --split;
    constexpr size_t kTileWsFloats = (size_t)CG * 8 * (BLOCK_N / 2 / 32) * 1024;
    const size_t need = (size_t)n_tail * (split - 1) * kTileWsFloats * sizeof(float);
    if (split > 1) {
      void* ws = nullptr;
      unsigned* flags = nullptr;
      if (splitk_workspace(stream, need, (size_t)n_tail * 2, &ws, &flags) == cudaSuccess) {
        g.tail_split = split;
        g.splitk_ws = reinterpret_cast<float*>(ws);
        g.splitk_flags = flags;
        if (n_tail * split > grid_units) grid_units = n_tail * split;   // fewer tiles than CTAs: still one slice each
      }
    }
  }
  const int grid = CG * grid_units;
  static const char* const kTags[] = {"gemm_linear", "gemm_gelu", "gemm_gelu_bwd", "gemm_qkv_rope", "gemm_patch"};
  const double Z = (double)g.total_tiles / ((double)g.m_tiles * g.n_tiles);
  const double mnk = (double)g.M * g.N * g.K;
  const double bytes = Z * (2.0 * ((double)g.M * g.K + (double)g.N * g.K) +
                            (OUT_F32 ? 4.0 : 2.0) * (double)g.M * g.N * (g.add ? 2.0 : 1.0));
  LaunchScope scope(kKindGemm, kTags[MODE], g.M, g.N, g.K, 2.0 * mnk * Z, bytes, stream);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  VQ_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, *tmA, *tmB, g));
  return VQ_OK;
}

template <int BLOCK_N, int MODE, bool OUT_F32>
static int launch_gemm(const CUtensorMap* tmA, const CUtensorMap* tmB, GemmArgs& g, cudaStream_t stream,
                       int cta_group) {
  // the lean instantiation unless this launch uses one of the epilogue extras (or tail splitting is switched on)
  const bool ex = g.ln_stats || g.stats_out || g.stats_zero || g.c_copy || g_gemm_tail_split > 1;
  if (ex) {
    if (cta_group == 2) return launch_gemm_cg<BLOCK_N, MODE, OUT_F32, 2, 2>(tmA, tmB, g, stream);
    return launch_gemm_cg<BLOCK_N, MODE, OUT_F32, 1, 2>(tmA, tmB, g, stream);
  }
  if constexpr (MODE == VQ_EPI_LINEAR || MODE == VQ_EPI_GELU || MODE == VQ_EPI_GELU_BWD) {
    if (g.drop.thr != 0) {
      if (cta_group == 2) return launch_gemm_cg<BLOCK_N, MODE, OUT_F32, 2, 1>(tmA, tmB, g, stream);
      return launch_gemm_cg<BLOCK_N, MODE, OUT_F32, 1, 1>(tmA, tmB, g, stream);
    }
  }
  if (cta_group == 2) return launch_gemm_cg<BLOCK_N, MODE, OUT_F32, 2, 0>(tmA, tmB, g, stream);
  return launch_gemm_cg<BLOCK_N, MODE, OUT_F32, 1, 0>(tmA, tmB, g, stream);
}

static inline bool fits_i32(long long v) { return v >= INT32_MIN && v <= INT32_MAX; }

}  // namespace vq

extern "C" int vq_gemm_tune(int32_t tail_split_max) {
  if (tail_split_max >= 0) vq::g_gemm_tail_split = tail_split_max > 8 ? 8 : tail_split_max;
  return VQ_OK;
}

extern "C" int vq_gemm(const vq_gemm_desc* d, vq_stream_t stream_) {
  using namespace vq;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VQ_CHECK_ARG(d != nullptr, "desc is null");
  VQ_CHECK_ARG(d->M > 0 && d->N > 0 && d->K > 0, "M, N, K must be positive");
  VQ_CHECK_ARG(d->N % 8 == 0, "N must be a multiple of 8");
  VQ_CHECK_ARG(d->A && d->B && d->C, "A, B, C must be non-null");
  VQ_CHECK_ARG(d->c_dtype == VQ_BF16 || d->c_dtype == VQ_F32, "c_dtype");
  const int Z0 = d->Z0 > 0 ? d->Z0 : 1, Z1 = d->Z1 > 0 ? d->Z1 : 1;
  VQ_CHECK_ARG(Z0 * Z1 == 1 || d->K % kBlockK == 0, "batched GEMM needs K % 64 == 0");
  const int elt_c = d->c_dtype == VQ_F32 ? 4 : 2;
  VQ_CHECK_ARG((d->ldc * elt_c) % 16 == 0 && (reinterpret_cast<uintptr_t>(d->C) % 16) == 0,
               "C rows must be 16-byte aligned");
  VQ_CHECK_ARG(!d->add || ((d->ldadd * elt_c) % 16 == 0 &&
                           (reinterpret_cast<uintptr_t>(d->add) % 16) == 0),
               "add rows must be 16-byte aligned");
  VQ_CHECK_ARG(!d->aux || ((d->ldaux * 2) % 16 == 0 && (reinterpret_cast<uintptr_t>(d->aux) % 16) == 0),
               "aux rows must be 16-byte aligned");
  VQ_CHECK_ARG((d->c_c0 % 8) == 0 && (d->c_c1 % 8) == 0, "C column offsets must be multiples of 8");
  VQ_CHECK_ARG(!d->bias || (reinterpret_cast<uintptr_t>(d->bias) % 16) == 0, "bias alignment");
  VQ_CHECK_ARG(!d->gamma || (reinterpret_cast<uintptr_t>(d->gamma) % 16) == 0, "gamma alignment");

  // Tiling.  Large unbatched problems run on CTA pairs (cta_group::2, 256 x 256 tiles: each CTA stages only half
  // of B, which keeps smem traffic within the 128 B/clk port at full tensor rate); batched / small-M problems
  // use single CTAs with 128 x 256 or 128 x 128 tiles.
  static const int force_cg = [] {
    const char* e = getenv("VQ_GEMM_CTA_GROUP");
    return e ? atoi(e) : 0;
  }();
  int cg = (Z0 * Z1 == 1 && d->M > kBlockM && d->N >= 256) ? 2 : 1;
  if (force_cg == 1 || force_cg == 2) cg = (force_cg == 2 && (Z0 * Z1 != 1 || d->N < 256)) ? 1 : force_cg;
  const int m_tiles = (d->M + kBlockM * cg - 1) / (kBlockM * cg);
  int block_n = 256;
  if (cg == 2 && d->N % 128 == 0 && d->K <= 1024) {
    // 256 x 128 tiles only when 256 x 256 tiles would not even fill ONE round of the persistent grid (small M: the
    // narrower tile doubles the number of busy pairs).  With more than one round the wide tile wins inside a step
    // although the narrow one looks better alone: c2 proj (M = 10992, N = 1024: 172 tiles = 3 rounds against 5
    // half-rounds) is 5 % faster in isolation with 128 and the whole c2 step 1.5 % SLOWER (A is re-read twice as often;
    // at the power cap the extra L2 -> smem traffic costs clocks everywhere) -- `scripts/knob_sweep.sh`,
    // profiles/r02_knob_sweep_in_step.txt.  K = 4096 (fc2) loses either way, hence the K limit.
    const long long pairs = (num_sms() > 1 ? num_sms() : 148) / 2;
    const long long t256 = (long long)m_tiles * ((d->N + 255) / 256), t128 = (long long)m_tiles * (d->N / 128);
    static const int force_bn = [] {
      const char* e = getenv("VQ_GEMM_BLOCK_N");
      return e ? atoi(e) : 0;
    }();
    if (force_bn == 128 || (force_bn == 0 && t256 <= pairs && t128 <= pairs)) block_n = 128;
  }
  if (cg == 1) {
    const long long tiles256 = (long long)m_tiles * ((d->N + 255) / 256) * Z0 * Z1;
    const int sms = num_sms() > 0 ? num_sms() : 148;
    if (d->N % 256 != 0 || tiles256 < 4LL * sms) block_n = 128;
    if (d->epilogue == VQ_EPI_QKV_ROPE && d->N % 256 != 0) block_n = 128;
  }
  const int n_tiles = (d->N + block_n - 1) / block_n;

  // extents of the 2-D tensors the operands live in
  const long long a_outer_need = (d->a_major ? d->K : d->M) + (long long)(Z0 - 1) * d->a_r0 +
                                 (long long)(Z1 - 1) * d->a_r1;
  const long long a_inner_need = (d->a_major ? d->M : d->K) + (long long)(Z0 - 1) * d->a_c0 +
                                 (long long)(Z1 - 1) * d->a_c1;
  const long long b_outer_need = (d->b_major ? d->K : d->N) + (long long)(Z0 - 1) * d->b_r0 +
                                 (long long)(Z1 - 1) * d->b_r1;
  const long long b_inner_need = (d->b_major ? d->N : d->K) + (long long)(Z0 - 1) * d->b_c0 +
                                 (long long)(Z1 - 1) * d->b_c1;
  const long long a_rows = d->a_rows > 0 ? d->a_rows : a_outer_need;
  const long long a_cols = d->a_cols > 0 ? d->a_cols : a_inner_need;
  const long long b_rows = d->b_rows > 0 ? d->b_rows : b_outer_need;
  const long long b_cols = d->b_cols > 0 ? d->b_cols : b_inner_need;
  VQ_CHECK_ARG(d->lda >= a_cols && d->ldb >= b_cols, "lda/ldb smaller than the row extent");

  CUtensorMap mapA, mapB;
  if (!get_tmap_2d_bf16(d->A, a_cols, a_rows, d->lda, 64, d->a_major ? kBlockK : kBlockM, &mapA)) return VQ_ERR_CUDA;
  if (!get_tmap_2d_bf16(d->B, b_cols, b_rows, d->ldb, 64, d->b_major ? kBlockK : block_n / cg, &mapB)) return VQ_ERR_CUDA;
  const CUtensorMap *tmA = &mapA, *tmB = &mapB;

  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.M = d->M; g.N = d->N; g.K = d->K;
  g.a_mn = d->a_major ? 1 : 0;
  g.b_mn = d->b_major ? 1 : 0;
  g.Z1 = Z1;
  g.m_tiles = m_tiles;
  g.n_tiles = n_tiles;
  const long long total = (long long)m_tiles * n_tiles * Z0 * Z1;
  VQ_CHECK_ARG(fits_i32(total), "too many tiles");
  g.total_tiles = (int)total;
  VQ_CHECK_ARG(fits_i32(a_rows + 512) && fits_i32(a_cols + 512) && fits_i32(b_rows + 512) &&
                   fits_i32(b_cols + 512),
               "operand extents exceed int32");
  g.a_r0 = (int)d->a_r0; g.a_c0 = (int)d->a_c0; g.a_r1 = (int)d->a_r1; g.a_c1 = (int)d->a_c1;
  g.b_r0 = (int)d->b_r0; g.b_c0 = (int)d->b_c0; g.b_r1 = (int)d->b_r1; g.b_c1 = (int)d->b_c1;
  g.c_r0 = d->c_r0; g.c_c0 = d->c_c0; g.c_r1 = d->c_r1; g.c_c1 = d->c_c1;
  g.C = d->C; g.ldc = d->ldc;
  g.alpha = d->alpha == 0.f ? 1.f : d->alpha;
  g.bias = d->bias; g.gamma = d->gamma;
  g.add = d->add; g.ldadd = d->ldadd;
  g.aux = d->aux; g.ldaux = d->ldaux;
  g.tokens_per_frame = d->tokens_per_frame; g.patch_start = d->patch_start; g.grid_w = d->grid_w;
  g.apply_qknorm_rope = d->apply_qknorm_rope;
  g.qn_w = d->qn_w; g.qn_b = d->qn_b; g.kn_w = d->kn_w; g.kn_b = d->kn_b;
  g.qk_eps = d->qk_eps; g.q_scale = d->q_scale == 0.f ? 1.f : d->q_scale;
  g.rope_cos = d->rope_cos; g.rope_sin = d->rope_sin;
  g.rope_positions = d->rope_positions;
  g.n_patches = d->n_patches; g.pos_embed = d->pos_embed;
  VQ_CHECK_ARG(d->dropout.p >= 0.f && d->dropout.p < 1.f, "dropout p must be in [0, 1)");
  VQ_CHECK_ARG(d->dropout.p == 0.f || (Z0 * Z1 == 1 && (d->epilogue == VQ_EPI_LINEAR || d->epilogue == VQ_EPI_GELU ||
                                                       d->epilogue == VQ_EPI_GELU_BWD)),
               "dropout: unbatched LINEAR / GELU / GELU_BWD epilogues only");
  g.drop = make_dropout_params(d->dropout.p, d->dropout.seed, d->dropout.offset, d->dropout.site);
  VQ_CHECK_ARG((d->ln_stats == nullptr) == (d->ln_colsum == nullptr), "ln_stats and ln_colsum go together");
  VQ_CHECK_ARG(!d->ln_stats || (Z0 * Z1 == 1 && !d->gamma && d->bias && !d->a_major &&
                                (d->epilogue == VQ_EPI_LINEAR || d->epilogue == VQ_EPI_GELU ||
                                 d->epilogue == VQ_EPI_QKV_ROPE)),
               "LayerNorm folding: unbatched K-major A, bias required, no gamma, LINEAR / GELU / QKV_ROPE epilogue");
  VQ_CHECK_ARG(!d->ln_colsum || (reinterpret_cast<uintptr_t>(d->ln_colsum) % 16) == 0, "ln_colsum alignment");
  VQ_CHECK_ARG(!d->ln_stats || (reinterpret_cast<uintptr_t>(d->ln_stats) % 8) == 0, "ln_stats alignment");
  VQ_CHECK_ARG(!(d->stats_out || d->stats_zero || d->c_copy_bf16) ||
                   (Z0 * Z1 == 1 && d->epilogue == VQ_EPI_LINEAR && d->c_dtype == VQ_F32),
               "row statistics / bf16 copy: unbatched LINEAR epilogue with f32 output only");
  VQ_CHECK_ARG(!d->c_copy_bf16 || (reinterpret_cast<uintptr_t>(d->c_copy_bf16) % 8) == 0, "c_copy_bf16 alignment");
  g.ln_stats = d->ln_stats; g.ln_colsum = d->ln_colsum;
  g.ln_eps = d->ln_eps; g.ln_inv_dim = 1.0f / (float)d->K;
  g.stats_out = d->stats_out; g.stats_zero = d->stats_zero; g.c_copy = d->c_copy_bf16;

  const bool f32 = d->c_dtype == VQ_F32;
  switch (d->epilogue) {
    case VQ_EPI_LINEAR:
      if (block_n == 256)
        return f32 ? launch_gemm<256, VQ_EPI_LINEAR, true>(tmA, tmB, g, stream, cg)
                   : launch_gemm<256, VQ_EPI_LINEAR, false>(tmA, tmB, g, stream, cg);
      return f32 ? launch_gemm<128, VQ_EPI_LINEAR, true>(tmA, tmB, g, stream, cg)
                 : launch_gemm<128, VQ_EPI_LINEAR, false>(tmA, tmB, g, stream, cg);
    case VQ_EPI_GELU:
      VQ_CHECK_ARG(!f32, "GELU epilogue writes bf16");
      VQ_CHECK_ARG(d->bias != nullptr, "GELU epilogue needs a bias");
      return block_n == 256 ? launch_gemm<256, VQ_EPI_GELU, false>(tmA, tmB, g, stream, cg)
                            : launch_gemm<128, VQ_EPI_GELU, false>(tmA, tmB, g, stream, cg);
    case VQ_EPI_GELU_BWD:
      VQ_CHECK_ARG(!f32, "GELU_BWD epilogue writes bf16");
      VQ_CHECK_ARG(d->aux != nullptr, "GELU_BWD epilogue needs aux (saved pre-activation)");
      return block_n == 256 ? launch_gemm<256, VQ_EPI_GELU_BWD, false>(tmA, tmB, g, stream, cg)
                            : launch_gemm<128, VQ_EPI_GELU_BWD, false>(tmA, tmB, g, stream, cg);
    case VQ_EPI_QKV_ROPE:
      VQ_CHECK_ARG(!f32, "QKV epilogue writes bf16");
      VQ_CHECK_ARG(d->bias != nullptr, "QKV epilogue needs a bias");
      VQ_CHECK_ARG(d->N % 192 == 0, "QKV epilogue: N = 3 * heads * 64");
      VQ_CHECK_ARG(d->tokens_per_frame > 0 && d->grid_w > 0, "QKV epilogue: token geometry");
      VQ_CHECK_ARG(!d->apply_qknorm_rope || (d->qn_w && d->qn_b && d->kn_w && d->kn_b &&
                                             d->rope_cos && d->rope_sin),
                   "QKV epilogue: q/k norm and rope tables required");
      return block_n == 256 ? launch_gemm<256, VQ_EPI_QKV_ROPE, false>(tmA, tmB, g, stream, cg)
                            : launch_gemm<128, VQ_EPI_QKV_ROPE, false>(tmA, tmB, g, stream, cg);
    case VQ_EPI_PATCH:
      VQ_CHECK_ARG(f32, "PATCH epilogue writes the f32 token stream");
      VQ_CHECK_ARG(d->pos_embed && d->n_patches > 0 && d->tokens_per_frame > 0, "PATCH epilogue args");
      VQ_CHECK_ARG(Z0 * Z1 == 1, "PATCH epilogue is not batched");
      return block_n == 256 ? launch_gemm<256, VQ_EPI_PATCH, true>(tmA, tmB, g, stream, cg)
                            : launch_gemm<128, VQ_EPI_PATCH, true>(tmA, tmB, g, stream, cg);
    default:
      set_error("unknown epilogue");
      return VQ_ERR_INVALID;
  }
}
