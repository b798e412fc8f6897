// Shared fp32 SIMT GEMM tile core (used by tower.cu and din.cu).
#pragma once
#include "common.cuh"

namespace ctr {

// ----------------------------------------------------------------- tiled GEMM core
// C[32 x 128] += A[32 x nk] . B[nk x 128], operands produced element-wise by functors (so the
// BN/dropout prologue and the ReLU/BN-backward gradient source are applied on the way into
// shared memory).  Register-prefetch double buffering: the global loads of chunk i+1 are in
// flight while chunk i is multiplied; one __syncthreads per chunk.  256 threads, 4x4 micro-tile.
constexpr int kTwBM = 32, kTwBN = 128, kTwKC = 32;
constexpr int kTwAP = kTwBM + 4, kTwBP = kTwBN + 4;   // padded pitches (16-byte aligned rows)

struct GemmSmem {
  float A[2][kTwKC][kTwAP];
  float B[2][kTwKC][kTwBP];
};

template <bool A_KFAST, bool B_KFAST, typename FA, typename FB>
__device__ __forceinline__ void gemm_32x128(GemmSmem& sm, int nk, FA fa, FB fb, float (&acc)[4][4]) {
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  float ra[4], rb[16];
  auto load = [&](int k0) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int e = tid + 256 * t;
      const int kk = A_KFAST ? (e & 31) : (e >> 5), rr = A_KFAST ? (e >> 5) : (e & 31);
      ra[t] = (k0 + kk < nk) ? fa(rr, k0 + kk) : 0.f;
    }
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      const int e = tid + 256 * t;
      const int kk = B_KFAST ? (e & 31) : (e >> 7), c = B_KFAST ? (e >> 5) : (e & 127);
      rb[t] = (k0 + kk < nk) ? fb(k0 + kk, c) : 0.f;
    }
  };
  auto store = [&](int buf) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int e = tid + 256 * t;
      const int kk = A_KFAST ? (e & 31) : (e >> 5), rr = A_KFAST ? (e >> 5) : (e & 31);
      sm.A[buf][kk][rr] = ra[t];
    }
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      const int e = tid + 256 * t;
      const int kk = B_KFAST ? (e & 31) : (e >> 7), c = B_KFAST ? (e >> 5) : (e & 127);
      sm.B[buf][kk][c] = rb[t];
    }
  };
  load(0);
  store(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < nk; k0 += kTwKC) {
    const bool more = k0 + kTwKC < nk;
    if (more) load(k0 + kTwKC);
#pragma unroll
    for (int kk = 0; kk < kTwKC; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&sm.A[buf][kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&sm.B[buf][kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i][0] = fmaf(av[i], b.x, acc[i][0]);
        acc[i][1] = fmaf(av[i], b.y, acc[i][1]);
        acc[i][2] = fmaf(av[i], b.z, acc[i][2]);
        acc[i][3] = fmaf(av[i], b.w, acc[i][3]);
      }
    }
    if (more) store(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
}


}  // namespace ctr
