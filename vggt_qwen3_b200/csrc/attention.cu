// vq_attention_fwd: flash-style non-causal attention for head_dim 64 on tcgen05 / TMEM (sm_100a).
//
// One CTA = one (sequence, head, block of 256 query rows) = two 128-row query tiles that ping-pong
// on the tensor core while each tile's softmax warpgroup works:
//
//   warps 0..3    softmax warpgroup of query tile 0 : one thread per query row
//   warps 4..7    softmax warpgroup of query tile 1
//   warp 8        TMA producer: Q tiles once, then K_j / V_j (128 keys x 64) through a 4-stage ring
//   warp 9        MMA issuer (whole warp runs the loop, one elected lane issues tcgen05):
//                                         S_t = Q_t K_j^T   (SS, 128x128x64, accumulator in TMEM)
//                                         O_t += P_t V_j    (TS: P read from TMEM, V MN-major in smem)
//                 (highest active warp id: the scheduler favours it over the busy softmax warps)
//
// TMEM columns (512): S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384) P0 [384,448) P1 [448,512).
// q is pre-scaled by log2(e)/sqrt(64) in the QKV GEMM epilogue, so probabilities are exp2(s - max).
#include <stdlib.h>

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
  long long* trace;  // debug: clock64() timestamps of CTA (0,0,0) (nullptr in production)
  void* out;
  int seq_len;
  int heads;
  int n_kv_tiles;
};

// trace layout: [role (0 = MMA thread, 1 = softmax warp 4 lane 0, 2 = softmax warp 8 lane 0)][iteration < 32][event < 8]
#define VQ_TRACE(role, j, ev)                                                                     \
  do {                                                                                            \
    if (g.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (j) < 32)            \
      g.trace[((role) * 32 + (j)) * 8 + (ev)] = clock64();                                        \
  } while (0)

template <int DBG>
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
  const uint32_t bar_sfree = bar_odone + 16;               // [2] softmax has S_t in registers: S_t may be overwritten
  const uint32_t bar_token = bar_sfree + 16;               // [2] MUFU token: warpgroup t may run its exp phase
  const uint32_t tmem_slot = bar_token + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // q-block index is the slowest grid dimension: the (partly empty) last q-blocks are scheduled last
  const int head = blockIdx.x, seq = blockIdx.y, qb = blockIdx.z;
  const int L = g.seq_len;
  const int HD = g.heads * kHeadDim;
  const int seq_row0 = seq * L;
  const int q_row0 = qb * 2 * kTileQ;                        // within the sequence
  const int n_qtiles = (q_row0 + kTileQ < L) ? 2 : 1;        // second tile may be entirely past the end
  const int n_kv = g.n_kv_tiles;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(bar_qfull, 1);
    for (int s = 0; s < kKVStages; ++s) {
      mbar_init(bar_kfull + 8 * s, 1);
      mbar_init(bar_vfull + 8 * s, 1);
      mbar_init(bar_kvempty + 8 * s, 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(bar_sfull + 8 * t, 1);
      mbar_init(bar_pfull + 8 * t, 4);   // one elected lane per softmax warp
      mbar_init(bar_odone + 8 * t, 1);
      mbar_init(bar_sfree + 8 * t, 4);
      mbar_init(bar_token + 8 * t, 4);
    }
    fence_barrier_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp >= 8) {
    setmaxnreg_dec<56>();
    if (warp == 8 && lane == 0) {
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
    } else if (warp == 9) {
      // ===================== MMA issuer =====================
      // Every lane runs the control flow (waits, descriptor arithmetic stay warp-uniform); only the tcgen05
      // instructions themselves are predicated on one elected lane.
      const uint32_t idesc_s = make_idesc_bf16(kTileQ, kTileKV, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(kTileQ, kHeadDim, 0, 1);  // V is MN-major (d contiguous)
      const uint64_t desc_q = make_smem_desc(smem_q, 16, 1024);
      const uint64_t desc_k = make_smem_desc(smem_k, 16, 1024);
      const uint64_t desc_v = make_smem_desc(smem_v, 8192, 1024);
      auto issue_s = [&](int t, int stage) {
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kHeadDim / 16; ++k) {
            // descriptor start-address field is in 16-byte units: +2 per 32-byte k-step, +1024 per 16 KB tile
            const uint64_t adesc = desc_q + (uint64_t)(t * (kTileBytes >> 4) + k * 2);
            const uint64_t bdesc = desc_k + (uint64_t)(stage * (kTileBytes >> 4) + k * 2);
            if constexpr ((DBG & 32) == 0) umma_ss(tmem_base + t * 128, adesc, bdesc, idesc_s, k ? 1u : 0u);
          }
          umma_commit(bar_sfull + 8 * t);
        }
        __syncwarp();
      };
      mbar_wait(bar_qfull, 0);
      mbar_wait(bar_kfull, 0);
      tc_fence_after();
      for (int t = 0; t < n_qtiles; ++t) issue_s(t, 0);
      // Steady state, per kv tile j and query tile t:
      //   s_free[t](j)  (softmax t holds S_t(j) in registers)  ->  issue S_t(j+1) = Q_t K_{j+1}^T
      //   p_full[t](j)  (softmax t wrote P_t(j))               ->  issue O_t += P_t(j) V_j
      // so the next scores are ready before the softmax of the current tile finishes.
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_kv; ++j) {
        const int nstage = (stage + 1 == kKVStages) ? 0 : stage + 1;
        const uint32_t nphase = (stage + 1 == kKVStages) ? (phase ^ 1) : phase;
        VQ_TRACE(0, j, 0);
        if (j + 1 < n_kv) {
          mbar_wait_fast(bar_kfull + 8 * nstage, nphase);
          VQ_TRACE(0, j, 1);
          for (int t = 0; t < n_qtiles; ++t) {
            mbar_wait_fast(bar_sfree + 8 * t, j & 1);
            VQ_TRACE(0, j, 2 + t);
            tc_fence_after();
            issue_s(t, nstage);
          }
        }
        mbar_wait_fast(bar_vfull + 8 * stage, phase);
        VQ_TRACE(0, j, 4);
        for (int t = 0; t < n_qtiles; ++t) {
          mbar_wait_fast(bar_pfull + 8 * t, j & 1);
          VQ_TRACE(0, j, 5 + t);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kTileKV / 16; ++k) {
              // 16 keys per MMA: 8 TMEM columns of packed bf16 P, 16 smem rows (2048 B = 128 units) of V
              const uint64_t bdesc = desc_v + (uint64_t)(stage * (kTileBytes >> 4) + k * 128);
              if constexpr ((DBG & 16) == 0)
                umma_ts(tmem_base + 256 + t * 64, tmem_base + 384 + t * 64 + k * 8, bdesc, idesc_o,
                        (j | k) ? 1u : 0u);
            }
            umma_commit(bar_odone + 8 * t);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(bar_kvempty + 8 * stage);
        __syncwarp();
        stage = nstage;
        phase = nphase;
      }
    }
  } else {
    // ===================== softmax warpgroups =====================
    setmaxnreg_inc<216>();
    const int t = warp >> 2;
    if (t < n_qtiles) {
      const int quarter = warp & 3;
      const int row = quarter * 32 + lane;                       // row within the query tile
      const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
      const uint32_t tS = tmem_base + lane_addr + t * 128;
      const uint32_t tO = tmem_base + lane_addr + 256 + t * 64;
      const uint32_t tP = tmem_base + lane_addr + 384 + t * 64;
      float m_run = -INFINITY, l_run = 0.f;
      // Ping-pong of the two softmax warpgroups over the MUFU pipe (16 ex2/clk/SM bounds this kernel): a warpgroup
      // runs its exp phase only while holding the token (mbarrier bar_token[t]), then hands it to the other one,
      // so one warpgroup's loads / max / pack / TMEM stores overlap the other's exponentials instead of both
      // competing for the MUFU at the same time and idling it together afterwards.  (A named bar.sync does not
      // work here: BAR.SYNC.DEFER_BLOCKING lets register-only math run ahead of the barrier.)
      const bool pingpong = n_qtiles == 2;
      if (pingpong && t == 1 && lane == 0) mbar_arrive(bar_token);  // warpgroup 0 goes first

      const int trole = 1 + t;
      const bool tlane = (warp & 3) == 0 && lane == 0;
      for (int j = 0; j < n_kv; ++j) {
        if (tlane) VQ_TRACE(trole, j, 0);
        mbar_wait(bar_sfull + 8 * t, j & 1);
        if (tlane) VQ_TRACE(trole, j, 1);
        tc_fence_after();
        float s[128];
        {
          uint32_t r[32];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if ((DBG & 4) == 0 || c == 0) {  // timing experiment (DBG & 4): one TMEM load instead of four
              tmem_ld32(tS + c * 32, r);
              tmem_wait_ld();
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) s[c * 32 + i] = __uint_as_float(r[i]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_sfree + 8 * t);  // S_t is in registers: the MMA warp may start S_t(j+1)
        if (tlane) VQ_TRACE(trole, j, 2);
        const int n_valid = L - j * kTileKV;  // keys of this tile inside the sequence
        if (n_valid < kTileKV) {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (i >= n_valid) s[i] = -INFINITY;
        }
        // four independent max / sum chains: with two softmax warps per scheduler the dependent-issue latency
        // of a single 128-long chain would otherwise sit on the critical path
        float mx4[4] = {s[0], s[1], s[2], s[3]};
#pragma unroll
        for (int i = 4; i < 128; i += 4) {
          mx4[0] = fmaxf(mx4[0], s[i]);
          mx4[1] = fmaxf(mx4[1], s[i + 1]);
          mx4[2] = fmaxf(mx4[2], s[i + 2]);
          mx4[3] = fmaxf(mx4[3], s[i + 3]);
        }
        float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        if constexpr (DBG & 2) mx = s[0];  // timing experiment: no row max
        const float m_new = fmaxf(m_run, mx);
        const float alpha = ex2_approx(m_run - m_new);  // 0 on the first tile (m_run = -inf)
        if (tlane) VQ_TRACE(trole, j, 3);
        float m_use = m_new;
        if (pingpong) {
          mbar_wait_fast(bar_token + 8 * t, j & 1);
          asm volatile("" : "+f"(m_use));  // keep the exponentials behind the token wait
        }
        if (tlane) VQ_TRACE(trole, j, 4);
        float sum4[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t p[64];
        // One warp alone reaches only ~70 % of the MUFU rate, two together saturate it: the token is handed over
        // after kHandoff of the 64 pair-iterations so the other warpgroup's exponentials overlap the tail of ours.
        constexpr int kHandoff = 40;
#pragma unroll
        for (int i = 0; i < 64; i += 2) {
          float p0, p1, p2, p3;
          if constexpr (DBG & 1) {  // timing experiment: no MUFU
            p0 = s[2 * i] - m_use; p1 = s[2 * i + 1] - m_use; p2 = s[2 * i + 2] - m_use; p3 = s[2 * i + 3] - m_use;
          } else {
            p0 = ex2_approx(s[2 * i] - m_use);
            p1 = ex2_approx(s[2 * i + 1] - m_use);
            p2 = ex2_approx(s[2 * i + 2] - m_use);
            p3 = ex2_approx(s[2 * i + 3] - m_use);
          }
          sum4[0] += p0;
          sum4[1] += p1;
          sum4[2] += p2;
          sum4[3] += p3;
          p[i] = pack_bf16x2(p0, p1);
          p[i + 1] = pack_bf16x2(p2, p3);
          if (i + 2 == kHandoff && pingpong) {
            asm volatile("" ::"f"(sum4[0]), "f"(sum4[1]), "f"(sum4[2]), "f"(sum4[3]));
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_token + 8 * (t ^ 1));
          }
        }
        if (tlane) VQ_TRACE(trole, j, 5);
        const float sum = (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
        l_run = l_run * alpha + sum;

        if (j > 0) {
          // PV_{j-1} must have retired before P is overwritten or O is rescaled
          mbar_wait(bar_odone + 8 * t, (j - 1) & 1);
          if (tlane) VQ_TRACE(trole, j, 6);
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
          if constexpr ((DBG & 8) == 0) {  // timing experiment (DBG & 8): no P store
            tmem_st32(tP, p_lo);
            tmem_st32(tP + 32, p_hi);
          }
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_pfull + 8 * t);
        if (tlane) VQ_TRACE(trole, j, 7);
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
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace vq

static long long* g_attn_trace = nullptr;
extern "C" void vq_debug_set_attention_trace(long long* device_buf) { g_attn_trace = device_buf; }

extern "C" int vq_attention_fwd(const void* qkv, void* out, int32_t n_seq, int32_t seq_len,
                                int32_t heads, vq_stream_t stream_) {
  using namespace vq;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VQ_CHECK_ARG(qkv && out, "qkv/out must be non-null");
  VQ_CHECK_ARG(n_seq > 0 && seq_len > 0 && heads > 0, "n_seq, seq_len, heads must be positive");
  VQ_CHECK_ARG(n_seq <= 65535 && (seq_len + 2 * kTileQ - 1) / (2 * kTileQ) <= 65535, "grid limits");
  VQ_CHECK_ARG((reinterpret_cast<uintptr_t>(out) % 16) == 0, "out alignment");
  const long long rows = (long long)n_seq * seq_len;
  VQ_CHECK_ARG(rows + 512 < INT32_MAX, "too many rows");
  const int HD = heads * kHeadDim;
  const CUtensorMap* tm = get_tmap_2d_bf16(qkv, 3ull * HD, (uint64_t)rows, 3ull * HD, 64, 128);
  if (!tm) return VQ_ERR_CUDA;

  static const int dbg = [] {
    const char* e = getenv("VQ_ATTN_DBG");  // timing experiments only (results are wrong when non-zero)
    return e ? atoi(e) : 0;
  }();
  auto kern = dbg == 1 ? attention_fwd_kernel<1> : dbg == 4 ? attention_fwd_kernel<4>
            : dbg == 8 ? attention_fwd_kernel<8> : dbg == 15 ? attention_fwd_kernel<15>
            : dbg == 31 ? attention_fwd_kernel<31> : dbg == 47 ? attention_fwd_kernel<47>
            : dbg == 63 ? attention_fwd_kernel<63> : attention_fwd_kernel<0>;
  static bool configured = false;
  if (!configured) {
    VQ_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes));
    configured = true;
  }
  AttnArgs g;
  g.trace = g_attn_trace;
  g.out = out;
  g.seq_len = seq_len;
  g.heads = heads;
  g.n_kv_tiles = (seq_len + kTileKV - 1) / kTileKV;
  dim3 grid(heads, n_seq, (seq_len + 2 * kTileQ - 1) / (2 * kTileQ));
  {
    const double flops = 4.0 * n_seq * heads * (double)seq_len * seq_len * kHeadDim;
    const double bytes = 2.0 * (double)rows * HD * 4.0;  // q, k, v read once + out written (bf16)
    LaunchScope scope(kKindAttention, "attention_fwd", n_seq, seq_len, heads, flops, bytes, stream);
    kern<<<grid, kAttnThreads, kAttnSmemBytes, stream>>>(*tm, g);
  }
  VQ_CHECK_CUDA(cudaGetLastError());
  return VQ_OK;
}
