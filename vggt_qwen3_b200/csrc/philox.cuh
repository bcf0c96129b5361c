// Counter-based dropout masks (Philox4x32-10) for the Perceiver's train-mode dropout sites
// (src/models/projector_perceiver.py:33,37,42,46,49 of the reference).
//
// A mask bit is a pure function of (seed, offset, site, element index), so the forward kernels (GEMM epilogues,
// softmax) and the backward kernels (GELU-backward epilogue, LayerNorm backward, softmax backward) regenerate the
// same mask without storing it.  Elements are grouped by 8 along the row: one Philox call yields 4 x 32 bits =
// eight 16-bit lanes, element j of the group is DROPPED when lane j < thr with thr = round(p * 65536).
//
//   group  = row * (ld / 8) + col / 8            (ld = logical row length of the tensor the site acts on, ld % 8 == 0)
//   ctr    = { group lo, group hi, site tag, offset lo },   key = { seed lo ^ offset hi, seed hi }
//
// oracle/perceiver_ref.py restates the same function in numpy; tests compare both bit for bit (vq_dropout_mask).
#pragma once
#include <stdint.h>

namespace vq {

struct DropoutParams {
  uint32_t thr;      // 0 = dropout off
  float scale;       // 1 / (1 - p)
  uint32_t k0, k1;   // Philox key
  uint32_t c2, c3;   // counter words 2 and 3: site tag, offset
};

#if defined(__CUDACC__)
#define VQ_HD __host__ __device__ __forceinline__
#else
#define VQ_HD inline
#endif

VQ_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                         uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 8 keep bits (bit j = element j of the group is kept)
VQ_HD uint32_t dropout_keep8(const DropoutParams& d, uint64_t group) {
  uint32_t r[4];
  philox4x32_10((uint32_t)group, (uint32_t)(group >> 32), d.c2, d.c3, d.k0, d.k1, r);
  uint32_t bits = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t lane = (r[j >> 1] >> (16 * (j & 1))) & 0xFFFFu;
    bits |= (lane >= d.thr ? 1u : 0u) << j;
  }
  return bits;
}

inline DropoutParams make_dropout_params(float p, uint64_t seed, uint64_t offset, uint32_t site_tag) {
  DropoutParams d;
  d.thr = 0; d.scale = 1.f; d.k0 = d.k1 = d.c2 = d.c3 = 0;
  if (p > 0.f) {
    double t = (double)p * 65536.0 + 0.5;
    if (t > 65535.0) t = 65535.0;
    d.thr = (uint32_t)t;
    if (d.thr == 0) d.thr = 1;
    d.scale = 1.0f / (1.0f - p);
    d.k0 = (uint32_t)seed ^ (uint32_t)(offset >> 32);
    d.k1 = (uint32_t)(seed >> 32);
    d.c2 = site_tag;
    d.c3 = (uint32_t)offset;
  }
  return d;
}

}  // namespace vq
