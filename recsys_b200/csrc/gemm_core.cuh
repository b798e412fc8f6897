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

// ------------------------------------------------------------------ vectorised variant
// Same pipeline, but the functors hand over 4 consecutive elements along the operand's
// contiguous axis (one LDG.128 when aligned), and the row tile is RT*8 rows (RT = 2 -> 16-row
// tiles: twice the CTAs when the batch alone cannot fill the machine).
//   fa4(rr, k)  A_KFAST : elements (rr, k..k+3)          !A_KFAST : elements (rr..rr+3, k)
//   fb4(k, c)   B_KFAST : elements (k..k+3, c)           !B_KFAST : elements (k, c..c+3)
template <int RT, bool A_KFAST, bool B_KFAST, typename FA, typename FB>
__device__ __forceinline__ void gemm_tile_v4(GemmSmem& sm, int nk, FA fa4, FB fb4,
                                             float (&acc)[RT][4]) {
  constexpr int BM = RT * 8;
  constexpr int NA4 = BM * kTwKC / 4;              // float4 loads for the A chunk
  constexpr int TA = (NA4 + 255) / 256;
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  float4 ra[TA], rb[4];
  auto a_pos = [&](int e, int& rr, int& kk) {
    if (A_KFAST) { rr = e >> 3; kk = (e & 7) * 4; }
    else { kk = e / (BM / 4); rr = (e % (BM / 4)) * 4; }
  };
  auto b_pos = [&](int e, int& kk, int& c) {
    if (B_KFAST) { c = e >> 3; kk = (e & 7) * 4; }
    else { kk = e >> 5; c = (e & 31) * 4; }
  };
  auto load = [&](int k0) {
#pragma unroll
    for (int t = 0; t < TA; ++t) {
      const int e = tid + 256 * t;
      if (e < NA4) {
        int rr, kk;
        a_pos(e, rr, kk);
        ra[t] = fa4(rr, k0 + kk);
      }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      int kk, c;
      b_pos(tid + 256 * t, kk, c);
      rb[t] = fb4(k0 + kk, c);
    }
  };
  auto store = [&](int buf) {
#pragma unroll
    for (int t = 0; t < TA; ++t) {
      const int e = tid + 256 * t;
      if (e < NA4) {
        int rr, kk;
        a_pos(e, rr, kk);
        if (A_KFAST) {
          sm.A[buf][kk][rr] = ra[t].x;
          sm.A[buf][kk + 1][rr] = ra[t].y;
          sm.A[buf][kk + 2][rr] = ra[t].z;
          sm.A[buf][kk + 3][rr] = ra[t].w;
        } else {
          *reinterpret_cast<float4*>(&sm.A[buf][kk][rr]) = ra[t];
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      int kk, c;
      b_pos(tid + 256 * t, kk, c);
      if (B_KFAST) {
        sm.B[buf][kk][c] = rb[t].x;
        sm.B[buf][kk + 1][c] = rb[t].y;
        sm.B[buf][kk + 2][c] = rb[t].z;
        sm.B[buf][kk + 3][c] = rb[t].w;
      } else {
        *reinterpret_cast<float4*>(&sm.B[buf][kk][c]) = rb[t];
      }
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
      float av[RT];
      if constexpr (RT == 4) {
        const float4 a = *reinterpret_cast<const float4*>(&sm.A[buf][kk][ty * 4]);
        av[0] = a.x; av[1] = a.y; av[2] = a.z; av[3] = a.w;
      } else {
        const float2 a = *reinterpret_cast<const float2*>(&sm.A[buf][kk][ty * 2]);
        av[0] = a.x; av[1] = a.y;
      }
      const float4 b = *reinterpret_cast<const float4*>(&sm.B[buf][kk][tx * 4]);
#pragma unroll
      for (int i = 0; i < RT; ++i) {
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

// 4 consecutive floats at p (elements beyond `valid` read as 0); LDG.128 when possible.
__device__ __forceinline__ float4 load4_guard(const float* p, int valid, bool aligned) {
  if (valid >= 4 && aligned) return *reinterpret_cast<const float4*>(p);
  float4 v = f4_zero();
  if (valid > 0) v.x = p[0];
  if (valid > 1) v.y = p[1];
  if (valid > 2) v.z = p[2];
  if (valid > 3) v.w = p[3];
  return v;
}

}  // namespace ctr

// ------------------------------------------------------------- tensor-core variant (3xTF32)
// Same staging pipeline; the multiply runs on the warp-level tensor-core path
// (mma.sync.m16n8k8 tf32, fp32 accumulate) with the error-compensated split
//   a.b ~= a_lo.b_hi + a_hi.b_lo + a_hi.b_hi,   x_hi = tf32(x), x_lo = tf32(x - x_hi)
// which keeps fp32-grade accuracy (the dense tower must match the fp32 reference to 1e-4).
// The split is done once per element on the way into shared memory (hi and lo tiles); the 8
// warps each own 16 of the 128 columns and all BM rows.  This is for the small dense-tower
// GEMMs only - the CIN contraction uses tcgen05 (cin_tc.cuh).
namespace ctr {

constexpr int kMmAP = 40, kMmBP = 136;     // pitches = 8 (mod 32): conflict-free fragment loads

struct MmaSmem {
  float Ah[2][kTwKC][kMmAP], Al[2][kTwKC][kMmAP];
  float Bh[2][kTwKC][kMmBP], Bl[2][kTwKC][kMmBP];
};

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  uint32_t h, l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(x - hi));
  lo = __uint_as_float(l);
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const float (&a)[4], const float (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])),
        "r"(__float_as_uint(a[3])), "r"(__float_as_uint(b[0])), "r"(__float_as_uint(b[1])));
}

// acc[mt][nt][4]: C fragments of m16-tile mt (rows mt*16 + g, +8) and n8-tile nt of this warp
// (columns warp*16 + nt*8 + 2t, +1), g = lane >> 2, t = lane & 3.   MT = BM / 16.
template <int MT, bool A_KFAST, bool B_KFAST, typename FA, typename FB>
__device__ __forceinline__ void gemm_tile_mma(MmaSmem& sm, int nk, FA fa4, FB fb4,
                                              float (&acc)[MT][2][4], int ncols = kTwBN) {
  constexpr int BM = MT * 16;
  constexpr int NA4 = BM * kTwKC / 4;
  constexpr int TA = (NA4 + 255) / 256;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  struct Regs { float4 a[TA]; float4 b[4]; };
  auto a_pos = [&](int e, int& rr, int& kk) {
    if (A_KFAST) { rr = e >> 3; kk = (e & 7) * 4; }
    else { kk = e / (BM / 4); rr = (e % (BM / 4)) * 4; }
  };
  auto b_pos = [&](int e, int& kk, int& c) {
    if (B_KFAST) { c = e >> 3; kk = (e & 7) * 4; }
    else { kk = e >> 5; c = (e & 31) * 4; }
  };
  auto load = [&](int k0, Regs& R) {
#pragma unroll
    for (int i = 0; i < TA; ++i) {
      const int e = tid + 256 * i;
      if (e < NA4) {
        int rr, kk;
        a_pos(e, rr, kk);
        R.a[i] = fa4(rr, k0 + kk);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int kk, c;
      b_pos(tid + 256 * i, kk, c);
      R.b[i] = fb4(k0 + kk, c);
    }
  };
  auto split4 = [&](const float4& v, float4& hi, float4& lo) {
    split_tf32(v.x, hi.x, lo.x);
    split_tf32(v.y, hi.y, lo.y);
    split_tf32(v.z, hi.z, lo.z);
    split_tf32(v.w, hi.w, lo.w);
  };
  auto store = [&](int buf, const Regs& R) {
#pragma unroll
    for (int i = 0; i < TA; ++i) {
      const int e = tid + 256 * i;
      if (e < NA4) {
        int rr, kk;
        a_pos(e, rr, kk);
        float4 hi, lo;
        split4(R.a[i], hi, lo);
        if (A_KFAST) {          // 4 consecutive k of one row: transposing scalar stores
          sm.Ah[buf][kk][rr] = hi.x;     sm.Al[buf][kk][rr] = lo.x;
          sm.Ah[buf][kk + 1][rr] = hi.y; sm.Al[buf][kk + 1][rr] = lo.y;
          sm.Ah[buf][kk + 2][rr] = hi.z; sm.Al[buf][kk + 2][rr] = lo.z;
          sm.Ah[buf][kk + 3][rr] = hi.w; sm.Al[buf][kk + 3][rr] = lo.w;
        } else {
          *reinterpret_cast<float4*>(&sm.Ah[buf][kk][rr]) = hi;
          *reinterpret_cast<float4*>(&sm.Al[buf][kk][rr]) = lo;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int kk, c;
      b_pos(tid + 256 * i, kk, c);
      float4 hi, lo;
      split4(R.b[i], hi, lo);
      if (B_KFAST) {
        sm.Bh[buf][kk][c] = hi.x;     sm.Bl[buf][kk][c] = lo.x;
        sm.Bh[buf][kk + 1][c] = hi.y; sm.Bl[buf][kk + 1][c] = lo.y;
        sm.Bh[buf][kk + 2][c] = hi.z; sm.Bl[buf][kk + 2][c] = lo.z;
        sm.Bh[buf][kk + 3][c] = hi.w; sm.Bl[buf][kk + 3][c] = lo.w;
      } else {
        *reinterpret_cast<float4*>(&sm.Bh[buf][kk][c]) = hi;
        *reinterpret_cast<float4*>(&sm.Bl[buf][kk][c]) = lo;
      }
    }
  };
  const int nb = warp * 16;
  auto compute = [&](int buf) {
    if (nb < ncols) {    // warps whose 16 columns lie beyond the live columns only help staging
#pragma unroll
      for (int ks = 0; ks < kTwKC; ks += 8) {
        float ah[MT][4], al[MT][4], bh[2][2], bl[2][2];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const int m = mt * 16 + g;
          ah[mt][0] = sm.Ah[buf][ks + t][m];         al[mt][0] = sm.Al[buf][ks + t][m];
          ah[mt][1] = sm.Ah[buf][ks + t][m + 8];     al[mt][1] = sm.Al[buf][ks + t][m + 8];
          ah[mt][2] = sm.Ah[buf][ks + t + 4][m];     al[mt][2] = sm.Al[buf][ks + t + 4][m];
          ah[mt][3] = sm.Ah[buf][ks + t + 4][m + 8]; al[mt][3] = sm.Al[buf][ks + t + 4][m + 8];
        }
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const int n = nb + nt * 8 + g;
          bh[nt][0] = sm.Bh[buf][ks + t][n];      bl[nt][0] = sm.Bl[buf][ks + t][n];
          bh[nt][1] = sm.Bh[buf][ks + t + 4][n];  bl[nt][1] = sm.Bl[buf][ks + t + 4][n];
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            mma_tf32(acc[mt][nt], al[mt], bh[nt]);
            mma_tf32(acc[mt][nt], ah[mt], bl[nt]);
            mma_tf32(acc[mt][nt], ah[mt], bh[nt]);
          }
      }
    }
  };
  // Prefetch distance 2: while chunk i is multiplied from shared memory, chunk i+1 sits in one
  // register set (loaded an iteration ago) and chunk i+2 is being loaded into the other.
  Regs R0, R1;
  load(0, R0);
  store(0, R0);
  if (kTwKC < nk) load(kTwKC, R0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < nk; k0 += 2 * kTwKC) {
    // even step: shared[buf] = chunk k0, R0 = chunk k0 + KC
    if (k0 + 2 * kTwKC < nk) load(k0 + 2 * kTwKC, R1);
    compute(buf);
    if (k0 + kTwKC < nk) store(buf ^ 1, R0);
    __syncthreads();
    buf ^= 1;
    if (k0 + kTwKC >= nk) break;
    // odd step: shared[buf] = chunk k0 + KC, R1 = chunk k0 + 2 KC
    if (k0 + 3 * kTwKC < nk) load(k0 + 3 * kTwKC, R0);
    compute(buf);
    if (k0 + 2 * kTwKC < nk) store(buf ^ 1, R1);
    __syncthreads();
    buf ^= 1;
  }
}

}  // namespace ctr
