// Pieces shared by the per-layer tower kernels (tower.cu) and the fused middle-of-the-tower
// kernel (tower_mid.cu): the counter-based dropout stream and the BN constants.
#pragma once
#include "common.cuh"

namespace ctr {

// ------------------------------------------------------------------------ dropout RNG
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const unsigned hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const unsigned hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
// keep-scale (0 or 1/(1-p)) of element (row, col): one Philox block covers 4 consecutive columns.
__device__ __forceinline__ float drop_scale(unsigned seed, unsigned layer, unsigned step, int row,
                                            int col, float p, float inv_keep) {
  const uint4 r = philox4x32_10(make_uint4(static_cast<unsigned>(row), static_cast<unsigned>(col >> 2),
                                           layer, step),
                                make_uint2(seed, 0x5EEDu));
  const unsigned w = (col & 3) == 0 ? r.x : (col & 3) == 1 ? r.y : (col & 3) == 2 ? r.z : r.w;
  const float u = (w >> 8) * (1.0f / 16777216.0f);
  return u >= p ? inv_keep : 0.f;
}

// keep-scales of the 4 consecutive columns 4*quad .. 4*quad+3 of `row` (one Philox block)
__device__ __forceinline__ float4 drop_scale4(unsigned seed, unsigned layer, unsigned step, int row,
                                              int quad, float p, float inv_keep) {
  const uint4 r = philox4x32_10(make_uint4(static_cast<unsigned>(row), static_cast<unsigned>(quad),
                                           layer, step),
                                make_uint2(seed, 0x5EEDu));
  const float k = 1.0f / 16777216.0f;
  return make_float4((r.x >> 8) * k >= p ? inv_keep : 0.f, (r.y >> 8) * k >= p ? inv_keep : 0.f,
                     (r.z >> 8) * k >= p ? inv_keep : 0.f, (r.w >> 8) * k >= p ? inv_keep : 0.f);
}

constexpr int kMaxBn = 256;

__device__ __forceinline__ void bn_consts(const float* sums, const float* mean, const float* var,
                                          int k, int K, float inv_B, float eps, float* mu,
                                          float* rstd) {
  if (sums != nullptr) {
    const float m = sums[k] * inv_B;
    const float v = fmaxf(sums[K + k] * inv_B - m * m, 0.f);   // biased batch variance
    *mu = m;
    *rstd = rsqrtf(v + eps);
  } else {
    *mu = mean[k];
    *rstd = rsqrtf(var[k] + eps);
  }
}

}  // namespace ctr
