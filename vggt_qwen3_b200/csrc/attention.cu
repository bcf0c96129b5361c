// vq_attention_fwd: flash-style non-causal attention for head_dim 64 on tcgen05 / TMEM (sm_100a).
//
// One CTA = one (sequence, head, block of 256 query rows) = two 128-row query tiles that ping-pong
// on the tensor core while each tile's softmax warpgroup works:
//
//   warp 0        TMA producer: Q tiles once, then K_j / V_j (128 keys x 64) through a 4-stage ring
//   warp 1        MMA issuer (one lane):  S_t = Q_t K_j^T   (SS, 128x128x64, accumulator in TMEM)
//                                         O_t += P_t V_j    (TS: P read from TMEM, V MN-major in smem)
//   warps 4..7    softmax warpgroup of query tile 0 : one thread per query row
//   warps 8..11   softmax warpgroup of query tile 1
//
// TMEM columns (512): S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384) P0 [384,448) P1 [448,512).
// q is pre-scaled by log2(e)/sqrt(64) in the QKV GEMM epilogue, so probabilities are exp2(s - max).
#include "common.cuh"
#include "ptx.cuh"

namespace vq {

constexpr int kTileQ = 128;
constexpr int kTileKV = 128;
constexpr int kHeadDim = 64;
constexpr int kAttnThreads = 384;
constexpr int kKVStages = 4;
constexpr int kTileBytes = 128 * kHeadDim * 2;  // 16 KB: 128 rows x 128 B
constexpr int kAttnSmemBytes = 2 * kTileBytes + 2 * kKVStages * kTileBytes + 1024 + 256;

struct AttnArgs {
  void* out;
  int seq_len;
  int heads;
  int n_kv_tiles;
};

__global__ void __launch_bounds__(kAttnThreads, 1)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnArgs g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_q = smem_base;                               // 2 tiles
  const uint32_t smem_k = smem_base + 2 * kTileBytes;              // kKVStages tiles
  const uint32_t smem_v = smem_k + kKVStages * kTileBytes;         // kKVStages tiles
  const uint32_t bar_base = smem_v + kKVStages * kTileBytes;
  const uint32_t bar_qfull = bar_base;
  const uint32_t bar_kfull = bar_base + 8;
  const uint32_t bar_vfull = bar_kfull + 8 * kKVStages;
  const uint32_t bar_kvempty = bar_vfull + 8 * kKVStages;
  const uint32_t bar_sfull = bar_kvempty + 8 * kKVStages;  // [2]
  const uint32_t bar_pfull = bar_sfull + 16;               // [2]
  const uint32_t bar_odone = bar_pfull + 16;               // [2]
  const uint32_t tmem_slot = bar_odone + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qb = blockIdx.x, head = blockIdx.y, seq = blockIdx.z;
  const int L = g.seq_len;
  const int HD = g.heads * kHeadDim;
  const int seq_row0 = seq * L;
  const int q_row0 = qb * 2 * kTileQ;                        // within the sequence
  const int n_qtiles = (q_row0 + kTileQ < L) ? 2 : 1;        // second tile may be entirely past the end
  const int n_kv = g.n_kv_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(bar_qfull, 1);
    for (int s = 0; s < kKVStages; ++s) {
      mbar_init(bar_kfull + 8 * s, 1);
      mbar_init(bar_vfull + 8 * s, 1);
      mbar_init(bar_kvempty + 8 * s, 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(bar_sfull + 8 * t, 1);
      mbar_init(bar_pfull + 8 * t, 128);
      mbar_init(bar_odone + 8 * t, 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < 4) {
    setmaxnreg_dec<56>();
    if (warp == 0 && lane == 0) {
      // ===================== TMA producer =====================
      mbar_arrive_expect_tx(bar_qfull, n_qtiles * kTileBytes);
      for (int t = 0; t < n_qtiles; ++t)
        tma_load_2d(smem_q + t * kTileBytes, &tmQKV, bar_qfull, head * kHeadDim,
                    seq_row0 + q_row0 + t * kTileQ);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(bar_kvempty + 8 * stage, phase ^ 1);
        const int row = seq_row0 + j * kTileKV;
        mbar_arrive_expect_tx(bar_kfull + 8 * stage, kTileBytes);
        tma_load_2d(smem_k + stage * kTileBytes, &tmQKV, bar_kfull + 8 * stage,
                    HD + head * kHeadDim, row);
        mbar_arrive_expect_tx(bar_vfull + 8 * stage, kTileBytes);
        tma_load_2d(smem_v + stage * kTileBytes, &tmQKV, bar_vfull + 8 * stage,
                    2 * HD + head * kHeadDim, row);
        if (++stage == kKVStages) { stage = 0; phase ^= 1; }
      }
    } else if (warp == 1 && lane == 0) {
      // ===================== MMA issuer =====================
      const uint32_t idesc_s = make_idesc_bf16(kTileQ, kTileKV, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(kTileQ, kHeadDim, 0, 1);  // V is MN-major (d contiguous)
      auto issue_s = [&](int t, int stage) {
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k) {
          const uint64_t adesc = make_smem_desc(smem_q + t * kTileBytes + k * 32, 16, 1024);
          const uint64_t bdesc = make_smem_desc(smem_k + stage * kTileBytes + k * 32, 16, 1024);
          umma_ss(tmem_base + t * 128, adesc, bdesc, idesc_s, k ? 1u : 0u);
        }
        umma_commit(bar_sfull + 8 * t);
      };
      mbar_wait(bar_qfull, 0);
      mbar_wait(bar_kfull, 0);
      tc_fence_after();
      for (int t = 0; t < n_qtiles; ++t) issue_s(t, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_kv; ++j) {
        const int nstage = (stage + 1 == kKVStages) ? 0 : stage + 1;
        const uint32_t nphase = (stage + 1 == kKVStages) ? (phase ^ 1) : phase;
        mbar_wait(bar_vfull + 8 * stage, phase);
        for (int t = 0; t < n_qtiles; ++t) {
          mbar_wait(bar_pfull + 8 * t, j & 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < kTileKV / 16; ++k) {
            // 16 keys per MMA: 8 TMEM columns of packed bf16 P, 16 smem rows (2048 B) of V
            const uint64_t bdesc = make_smem_desc(smem_v + stage * kTileBytes + k * 2048, 8192, 1024);
            umma_ts(tmem_base + 256 + t * 64, tmem_base + 384 + t * 64 + k * 8, bdesc, idesc_o,
                    (j | k) ? 1u : 0u);
          }
          umma_commit(bar_odone + 8 * t);
          if (j + 1 < n_kv) {
            if (t == 0) {
              mbar_wait(bar_kfull + 8 * nstage, nphase);
              tc_fence_after();
            }
            issue_s(t, nstage);
          }
        }
        umma_commit(bar_kvempty + 8 * stage);
        stage = nstage;
        phase = nphase;
      }
    }
  } else {
    // ===================== softmax warpgroups =====================
    setmaxnreg_inc<216>();
    const int t = (warp - 4) >> 2;
    if (t < n_qtiles) {
      const int quarter = warp & 3;
      const int row = quarter * 32 + lane;                       // row within the query tile
      const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
      const uint32_t tS = tmem_base + lane_addr + t * 128;
      const uint32_t tO = tmem_base + lane_addr + 256 + t * 64;
      const uint32_t tP = tmem_base + lane_addr + 384 + t * 64;
      float m_run = -INFINITY, l_run = 0.f;

      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(bar_sfull + 8 * t, j & 1);
        tc_fence_after();
        float s[128];
        {
          uint32_t r[32];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            tmem_ld32(tS + c * 32, r);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) s[c * 32 + i] = __uint_as_float(r[i]);
          }
        }
        const int n_valid = L - j * kTileKV;  // keys of this tile inside the sequence
        if (n_valid < kTileKV) {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (i >= n_valid) s[i] = -INFINITY;
        }
        float mx = s[0];
#pragma unroll
        for (int i = 1; i < 128; ++i) mx = fmaxf(mx, s[i]);
        const float m_new = fmaxf(m_run, mx);
        const float alpha = ex2_approx(m_run - m_new);  // 0 on the first tile (m_run = -inf)
        float sum = 0.f;
        uint32_t p[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          const float p0 = ex2_approx(s[2 * i] - m_new);
          const float p1 = ex2_approx(s[2 * i + 1] - m_new);
          sum += p0 + p1;
          p[i] = pack_bf16x2(p0, p1);
        }
        l_run = l_run * alpha + sum;

        if (j > 0) {
          // PV_{j-1} must have retired before P is overwritten or O is rescaled
          mbar_wait(bar_odone + 8 * t, (j - 1) & 1);
          tc_fence_after();
          if (__any_sync(0xffffffffu, m_new > m_run)) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              uint32_t o[32];
              tmem_ld32(tO + c * 32, o);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st32(tO + c * 32, o);
            }
          }
        }
        {
          uint32_t(&p_lo)[32] = *reinterpret_cast<uint32_t(*)[32]>(&p[0]);
          uint32_t(&p_hi)[32] = *reinterpret_cast<uint32_t(*)[32]>(&p[32]);
          tmem_st32(tP, p_lo);
          tmem_st32(tP + 32, p_hi);
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(bar_pfull + 8 * t);
        m_run = m_new;
      }

      // ---- epilogue: O / l -> bf16 -> out[row, head*64 .. +64]
      mbar_wait(bar_odone + 8 * t, (n_kv - 1) & 1);
      tc_fence_after();
      const float inv_l = 1.0f / l_run;
      const int q_in_seq = q_row0 + t * kTileQ + row;
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(g.out) +
                           (long long)(seq_row0 + q_in_seq) * HD + head * kHeadDim;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t o[32];
        tmem_ld32(tO + c * 32, o);
        tmem_wait_ld();
        if (q_in_seq < L) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(o[i + 0]) * inv_l, __uint_as_float(o[i + 1]) * inv_l);
            u.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l);
            u.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l);
            u.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l);
            *reinterpret_cast<uint4*>(dst + c * 32 + i) = u;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace vq

extern "C" int vq_attention_fwd(const void* qkv, void* out, int32_t n_seq, int32_t seq_len,
                                int32_t heads, vq_stream_t stream_) {
  using namespace vq;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VQ_CHECK_ARG(qkv && out, "qkv/out must be non-null");
  VQ_CHECK_ARG(n_seq > 0 && seq_len > 0 && heads > 0, "n_seq, seq_len, heads must be positive");
  VQ_CHECK_ARG(n_seq <= 65535 && heads <= 65535, "grid limits");
  VQ_CHECK_ARG((reinterpret_cast<uintptr_t>(out) % 16) == 0, "out alignment");
  const long long rows = (long long)n_seq * seq_len;
  VQ_CHECK_ARG(rows + 512 < INT32_MAX, "too many rows");
  const int HD = heads * kHeadDim;
  const CUtensorMap* tm = get_tmap_2d_bf16(qkv, 3ull * HD, (uint64_t)rows, 3ull * HD, 64, 128);
  if (!tm) return VQ_ERR_CUDA;

  static bool configured = false;
  if (!configured) {
    VQ_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes));
    configured = true;
  }
  AttnArgs g;
  g.out = out;
  g.seq_len = seq_len;
  g.heads = heads;
  g.n_kv_tiles = (seq_len + kTileKV - 1) / kTileKV;
  dim3 grid((seq_len + 2 * kTileQ - 1) / (2 * kTileQ), heads, n_seq);
  {
    const double flops = 4.0 * n_seq * heads * (double)seq_len * seq_len * kHeadDim;
    const double bytes = 2.0 * (double)rows * HD * 4.0;  // q, k, v read once + out written (bf16)
    LaunchScope scope(kKindAttention, "attention_fwd", n_seq, seq_len, heads, flops, bytes, stream);
    attention_fwd_kernel<<<grid, kAttnThreads, kAttnSmemBytes, stream>>>(*tm, g);
  }
  VQ_CHECK_CUDA(cudaGetLastError());
  return VQ_OK;
}
