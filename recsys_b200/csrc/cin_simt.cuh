// xDeepFM CIN layer on fp32 CUDA cores - the exact-parity mode (CTR_CIN_FP32) and
// the yardstick the tcgen05 path (cin_tc.cu) is checked against on the device.
//
// All three data contractions of a CIN layer (xdeepfm/xdeepfm.py:145-169) are
// one shape:   out[r, n] (+)= act( sum_{a<A} sum_{c<C} P[r,a] * Q[r,c] * Wg[a,c,n] + bias[n] )
//   forward : P = X0t (A = m),  Q = Xp   (C = Hp), Wg[i,j,h] = W[(i*Hp+j)*H + h],   n = h
//   dXp     : P = X0t (A = m),  Q = dpre (C = H),  Wg[i,h,j] = W[(i*Hp+j)*H + h],   n = j
//   dX0t    : P = Xp  (A = Hp), Q = dpre (C = H),  Wg[j,h,i] = W[(i*Hp+j)*H + h],   n = i
// so one tiled SIMT GEMM whose A operand (the outer product Z = P (x) Q) is formed
// on the fly from shared-memory tiles serves all of them; Z is never materialised.
// dW is the same product reduced over rows (split across CTAs, RED to global).
#pragma once
#include "common.cuh"

namespace ctr {

constexpr int kCinBM = 64, kCinBN = 64, kCinKC = 16;

__global__ void __launch_bounds__(256)
cin_contract_kernel(const float* __restrict__ P, int ldp, int A, const float* __restrict__ Q,
                    int ldq, int C, const float* __restrict__ Wg, long long sa, long long sc,
                    long long sn, const float* __restrict__ bias, float* __restrict__ out, int ldo,
                    int M, int N, int relu, int accumulate) {
  extern __shared__ float smem[];
  float* Ps = smem;                         // [A][BM]
  float* Qs = Ps + A * kCinBM;              // [C][BM]
  float* Ws = Qs + C * kCinBM;              // [2][KC][BN]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int r0 = blockIdx.x * kCinBM;
  const int n0 = blockIdx.y * kCinBN;
  const int K = A * C;

  for (int e = tid; e < A * kCinBM; e += 256) {
    const int rr = e / A, a = e % A;       // coalesced along a
    const int r = r0 + rr;
    Ps[a * kCinBM + rr] = r < M ? P[static_cast<size_t>(r) * ldp + a] : 0.f;
  }
  for (int e = tid; e < C * kCinBM; e += 256) {
    const int rr = e / C, c = e % C;
    const int r = r0 + rr;
    Qs[c * kCinBM + rr] = r < M ? Q[static_cast<size_t>(r) * ldq + c] : 0.f;
  }
  auto load_w = [&](int k0, float* regs) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int e = tid + 256 * t;
      const int n = n0 + (e & 63);
      const int k = k0 + (e >> 6);
      float w = 0.f;
      if (k < K && n < N) {
        const int a = k / C, c = k - a * C;
        w = __ldg(Wg + a * sa + c * sc + n * sn);
      }
      regs[t] = w;
    }
  };
  auto store_w = [&](int buf, const float* regs) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int e = tid + 256 * t;
      Ws[buf * kCinKC * kCinBN + (e >> 6) * kCinBN + (e & 63)] = regs[t];
    }
  };

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float wreg[4];
  load_w(0, wreg);
  store_w(0, wreg);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < K; k0 += kCinKC) {
    const bool more = k0 + kCinKC < K;
    if (more) load_w(k0 + kCinKC, wreg);
    int a = k0 / C, c = k0 - a * C;
    const float* wsb = Ws + buf * kCinKC * kCinBN;
#pragma unroll
    for (int kk = 0; kk < kCinKC; ++kk) {
      if (k0 + kk < K) {
        const float4 pa = *reinterpret_cast<const float4*>(Ps + a * kCinBM + ty * 4);
        const float4 qc = *reinterpret_cast<const float4*>(Qs + c * kCinBM + ty * 4);
        const float4 b = *reinterpret_cast<const float4*>(wsb + kk * kCinBN + tx * 4);
        const float z[4] = {pa.x * qc.x, pa.y * qc.y, pa.z * qc.z, pa.w * qc.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[i][0] = fmaf(z[i], b.x, acc[i][0]);
          acc[i][1] = fmaf(z[i], b.y, acc[i][1]);
          acc[i][2] = fmaf(z[i], b.z, acc[i][2]);
          acc[i][3] = fmaf(z[i], b.w, acc[i][3]);
        }
        if (++c == C) {
          c = 0;
          ++a;
        }
      }
    }
    if (more) store_w(buf ^ 1, wreg);
    __syncthreads();
    buf ^= 1;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty * 4 + i;
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias != nullptr) v += bias[n];
      if (relu) v = fmaxf(v, 0.f);
      float* o = out + static_cast<size_t>(r) * ldo + n;
      *o = accumulate ? (*o + v) : v;
    }
  }
}

// dW[q, h] += sum_r X0t[r, q/Hp] * Xp[r, q%Hp] * dpre[r, h]   (rows split over gridDim.z)
__global__ void __launch_bounds__(256)
cin_dw_kernel(const float* __restrict__ X0t, int ld0, const float* __restrict__ Xp, int ldp,
              const float* __restrict__ dpre, int M, int m, int Hp, int H,
              float* __restrict__ dW, int rows_per_split) {
  __shared__ __align__(16) float Zs[kCinKC][kCinBN];
  __shared__ __align__(16) float Ds[kCinKC][kCinBN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int q0 = blockIdx.x * 64, h0 = blockIdx.y * 64;
  const int Kq = m * Hp;
  const int rbeg = blockIdx.z * rows_per_split;
  const int rend = min(M, rbeg + rows_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int rc = rbeg; rc < rend; rc += kCinKC) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int e = tid + 256 * t;
      const int rr = e >> 6, cc = e & 63;
      const int r = rc + rr;
      const int q = q0 + cc, h = h0 + cc;
      float z = 0.f, d = 0.f;
      if (r < rend) {
        if (q < Kq) {
          const int i = q / Hp, j = q - i * Hp;
          z = __ldg(X0t + static_cast<size_t>(r) * ld0 + i) * __ldg(Xp + static_cast<size_t>(r) * ldp + j);
        }
        if (h < H) d = __ldg(dpre + static_cast<size_t>(r) * H + h);
      }
      Zs[rr][cc] = z;
      Ds[rr][cc] = d;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < kCinKC; ++rr) {
      const float4 zq = *reinterpret_cast<const float4*>(&Zs[rr][ty * 4]);
      const float4 dh = *reinterpret_cast<const float4*>(&Ds[rr][tx * 4]);
      const float z[4] = {zq.x, zq.y, zq.z, zq.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i][0] = fmaf(z[i], dh.x, acc[i][0]);
        acc[i][1] = fmaf(z[i], dh.y, acc[i][1]);
        acc[i][2] = fmaf(z[i], dh.z, acc[i][2]);
        acc[i][3] = fmaf(z[i], dh.w, acc[i][3]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + ty * 4 + i;
    if (q >= Kq) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int h = h0 + tx * 4 + j;
      if (h < H) red_add_f32(dW + static_cast<size_t>(q) * H + h, acc[i][j]);
    }
  }
}

// dbias[h] += sum_r dpre[r, h]
__global__ void __launch_bounds__(256)
cin_colsum_kernel(const float* __restrict__ dpre, int M, int H, float* __restrict__ dbias) {
  const int h = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ry = threadIdx.x >> 5;
  float s = 0.f;
  if (h < H)
    for (int r = blockIdx.y * 8 + ry; r < M; r += gridDim.y * 8) s += dpre[static_cast<size_t>(r) * H + h];
  __shared__ float red[8][33];
  red[ry][threadIdx.x & 31] = s;
  __syncthreads();
  if (ry == 0 && h < H) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x & 31];
    red_add_f32(dbias + h, t);
  }
}

static inline int cin_contract_launch(const float* P, int ldp, int A, const float* Q, int ldq, int C,
                                      const float* Wg, long long sa, long long sc, long long sn,
                                      const float* bias, float* out, int ldo, int M, int N,
                                      int relu, int accumulate, cudaStream_t st) {
  const size_t smem = (static_cast<size_t>(A + C) * kCinBM + 2 * kCinKC * kCinBN) * sizeof(float);
  if (smem > 200 * 1024) return fail_arg("ctr_cin_layer", "m / Hp / H too large for the fp32 path");
  static size_t configured = 0;
  if (smem > configured) {
    cudaFuncSetAttribute(cin_contract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         static_cast<int>(smem));
    configured = smem;
  }
  dim3 grid((M + kCinBM - 1) / kCinBM, (N + kCinBN - 1) / kCinBN);
  cin_contract_kernel<<<grid, 256, smem, st>>>(P, ldp, A, Q, ldq, C, Wg, sa, sc, sn, bias, out, ldo,
                                               M, N, relu, accumulate);
  return CTR_OK;
}

}  // namespace ctr
