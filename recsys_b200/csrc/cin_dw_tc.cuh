// CIN weight gradient on tcgen05:  dW[(i,j), h] = sum_r X0t[r,i] * Xp[r,j] * dpre[r,h].
//
// The reduction index is the row r = (b, d) (131 072 rows at BASELINE config 3), so this is a
// split-K GEMM  dW[q, h] += ZT[q, r] . dpreT[h, r]  with both operands K-major in r:
//   cin_build_zt_kernel   ZT[(i*Hp + j), r] = tf32(X0t[r,i] * Xp[r,j])   (scratch, read once)
//   cin_transpose_kernel  dpreT[h, r]       = tf32(dpre[r,h])
//   cin_dw_tc_kernel      one 128 x NT output tile per CTA in TMEM, K range = one split of r,
//                         A/B k-blocks [128|NT x 32] streamed by TMA through a 6-stage ring,
//                         tcgen05.mma.kind::tf32, epilogue = vector RED into dW.
// 3xTF32 keeps lo parts of both operands and runs (hi,hi), (lo,hi), (hi,lo).
#pragma once
#include "cin_tc.cuh"

namespace ctr {

constexpr int kDwStages = 6;
constexpr int kDwStageBytes = kTcABytes + kTcBStageBytes / 2;   // A 16 KB + B (<=128 rows) 16 KB

struct CinDwParams {
  float* dW;
  int Kq, H, NT, M, n_pass;
  int kb_per_split;     // k-blocks (of 32 rows of r) per split
  uint32_t idesc;
  uint64_t desc_hi;
};

// ZT[(i*Hp + j) * ldz + r] = round_tf32(X0t[r,i] * Xp[r,j]) (+ lo part).  CTA = 32 rows of r.
__global__ void __launch_bounds__(256)
cin_build_zt_kernel(const float* __restrict__ X0t, int ld0, const float* __restrict__ Xp, int ldp,
                    int M, int m, int Hp, float* __restrict__ ZT, float* __restrict__ ZT_lo,
                    long long ldz) {
  extern __shared__ float zs[];           // xp[Hp][33], x0[m][33]
  float* xp = zs;
  float* x0 = zs + Hp * 33;
  const int r0 = blockIdx.x * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int e = threadIdx.x; e < 32 * Hp; e += 256) {
    const int rr = e / Hp, j = e % Hp;
    xp[j * 33 + rr] = (r0 + rr < M) ? Xp[static_cast<size_t>(r0 + rr) * ldp + j] : 0.f;
  }
  for (int e = threadIdx.x; e < 32 * m; e += 256) {
    const int rr = e / m, i = e % m;
    x0[i * 33 + rr] = (r0 + rr < M) ? X0t[static_cast<size_t>(r0 + rr) * ld0 + i] : 0.f;
  }
  __syncthreads();
  const int total = m * Hp;
  for (int q = warp; q < total; q += 8) {
    const int i = q / Hp, j = q - i * Hp;
    const float v = x0[i * 33 + lane] * xp[j * 33 + lane];
    const float hi = round_tf32(v);
    if (r0 + lane < M) {
      ZT[static_cast<size_t>(q) * ldz + r0 + lane] = hi;
      if (ZT_lo != nullptr) ZT_lo[static_cast<size_t>(q) * ldz + r0 + lane] = round_tf32(v - hi);
    }
  }
}

// dst[c, r] = round_tf32(src[r, c]) (+ lo), src [M, C] row-major, dst pitch ldd.
__global__ void __launch_bounds__(256)
cin_transpose_kernel(const float* __restrict__ src, int M, int Cn, float* __restrict__ dst,
                     float* __restrict__ dst_lo, long long ldd) {
  __shared__ float t[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int k = ty; k < 32; k += 8) {
    const int r = r0 + k, c = c0 + tx;
    t[k][tx] = (r < M && c < Cn) ? src[static_cast<size_t>(r) * Cn + c] : 0.f;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int c = c0 + k, r = r0 + tx;
    if (c < Cn && r < M) {
      const float v = t[tx][k];
      const float hi = round_tf32(v);
      dst[static_cast<size_t>(c) * ldd + r] = hi;
      if (dst_lo != nullptr) dst_lo[static_cast<size_t>(c) * ldd + r] = round_tf32(v - hi);
    }
  }
}

__global__ void __launch_bounds__(256, 1)
cin_dw_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
                 const CinDwParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(kDwStages) * kDwStageBytes);
  uint64_t* full = bars;                 // [kDwStages]
  uint64_t* empty = bars + 8;            // [kDwStages]
  uint64_t* t_full = bars + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kTcBM;
  const int kb_total = (p.M + kTcKB - 1) / kTcKB;
  const int kb_beg = blockIdx.y * p.kb_per_split;
  const int kb_end = min(kb_total, kb_beg + p.kb_per_split);
  const int nkb = max(0, kb_end - kb_beg);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kDwStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(t_full, 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    uint32_t it = 0;
    for (int ps = 0; ps < p.n_pass; ++ps)
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const uint32_t st = it % kDwStages, ph = (it / kDwStages) & 1;
        mbar_wait(&empty[st], ph ^ 1);
        mbar_expect_tx(&full[st], static_cast<uint32_t>(kTcABytes + p.NT * kTcKB * 4));
        uint8_t* sa = smem + static_cast<size_t>(st) * kDwStageBytes;
        tma_load_2d(sa, ps == 1 ? &tmA1 : &tmA0, (kb_beg + kb) * kTcKB, q0, &full[st]);
        tma_load_2d(sa + kTcABytes, ps == 2 ? &tmB1 : &tmB0, (kb_beg + kb) * kTcKB, 0, &full[st]);
      }
  } else if (warp == 1 && lane == 0) {
    uint32_t it = 0, accum = 0;
    for (int ps = 0; ps < p.n_pass; ++ps)
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const uint32_t st = it % kDwStages, ph = (it / kDwStages) & 1;
        mbar_wait(&full[st], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + static_cast<size_t>(st) * kDwStageBytes);
        const uint32_t b_addr = a_addr + kTcABytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ad = p.desc_hi | static_cast<uint64_t>(((a_addr + k * 32) >> 4) & 0x3FFF);
          const uint64_t bd = p.desc_hi | static_cast<uint64_t>(((b_addr + k * 32) >> 4) & 0x3FFF);
          tc_mma_tf32(tmem_base, ad, bd, p.idesc, accum);
          accum = 1;
        }
        tc_commit(&empty[st]);
      }
    tc_commit(t_full);
  } else if (warp >= 4) {
    const int quarter = warp & 3;
    const int q = q0 + quarter * 32 + lane;
    if (nkb > 0) {
      mbar_wait(t_full, 0);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
      for (int c = 0; c < p.NT; c += 8) {
        float v[8];
        tc_ld<8>(taddr + c, v);
        if (q < p.Kq) {
          float* o = p.dW + static_cast<size_t>(q) * p.H + c;
          if ((p.H & 3) == 0 && c + 8 <= p.H) {
            red_add_v4(o, make_float4(v[0], v[1], v[2], v[3]));
            red_add_v4(o + 4, make_float4(v[4], v[5], v[6], v[7]));
          } else {
#pragma unroll
            for (int t = 0; t < 8; ++t)
              if (c + t < p.H) red_add_f32(o + t, v[t]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256)
                 : "memory");
  }
}

static int64_t cin_dw_tc_ws(int M, int m, int Hp, int H, int prec) {
  const int64_t ldz = (static_cast<int64_t>(M) + 3) / 4 * 4;
  const int64_t elems = (static_cast<int64_t>(m) * Hp + H) * ldz;
  return elems * 4 * (prec == CTR_CIN_TF32X3 ? 2 : 1) + 2048;
}

static bool cin_dw_tc_supported(int H) { return H <= 128 && H >= 8; }

static int cin_dw_tc(const float* X0t, int ld0, const float* Xp, int ldp, const float* dpre, int M,
                     int m, int Hp, int H, float* dW, int prec, void* ws, int64_t ws_bytes,
                     cudaStream_t st, const char* fn) {
  const bool split = prec == CTR_CIN_TF32X3;
  const long long ldz = (static_cast<long long>(M) + 3) / 4 * 4;
  const int Kq = m * Hp;
  CTR_REQUIRE(ws != nullptr && ws_bytes >= cin_dw_tc_ws(M, m, Hp, H, prec), fn, "workspace too small");
  CTR_REQUIRE(ldz < (1LL << 31), fn, "too many rows for the tensor-core dW path");
  auto align256 = [](void* q) {
    return reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(q) + 255) & ~uintptr_t(255));
  };
  float* ZT = align256(ws);
  float* dT = align256(ZT + static_cast<size_t>(Kq) * ldz);
  float* ZT_lo = split ? align256(dT + static_cast<size_t>(H) * ldz) : nullptr;
  float* dT_lo = split ? align256(ZT_lo + static_cast<size_t>(Kq) * ldz) : nullptr;
  {
    const size_t smem = static_cast<size_t>(Hp + m) * 33 * sizeof(float);
    CTR_REQUIRE(smem <= 100 * 1024, fn, "m + Hp too large for the ZT builder");
    static size_t configured = 0;
    if (smem > configured) {
      cudaFuncSetAttribute(cin_build_zt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(smem));
      configured = smem;
    }
    cin_build_zt_kernel<<<(M + 31) / 32, 256, smem, st>>>(X0t, ld0, Xp, ldp, M, m, Hp, ZT, ZT_lo, ldz);
    dim3 g((M + 31) / 32, (H + 31) / 32);
    cin_transpose_kernel<<<g, 256, 0, st>>>(dpre, M, H, dT, dT_lo, ldz);
  }
  const int NT = (H + 15) / 16 * 16;
  CUtensorMap tA0, tA1, tB0, tB1;
  int r = make_map(&tA0, ZT, Kq, M, static_cast<int>(ldz), kTcBM);
  if (r != CTR_OK) return r;
  r = make_map(&tA1, split ? ZT_lo : ZT, Kq, M, static_cast<int>(ldz), kTcBM);
  if (r != CTR_OK) return r;
  r = make_map(&tB0, dT, H, M, static_cast<int>(ldz), NT);
  if (r != CTR_OK) return r;
  r = make_map(&tB1, split ? dT_lo : dT, H, M, static_cast<int>(ldz), NT);
  if (r != CTR_OK) return r;
  CinDwParams p;
  p.dW = dW; p.Kq = Kq; p.H = H; p.NT = NT; p.M = M; p.n_pass = split ? 3 : 1;
  const int qtiles = (Kq + kTcBM - 1) / kTcBM;
  const int kb_total = (M + kTcKB - 1) / kTcKB;
  int splits = std::max(1, std::min(sm_count() / qtiles, kb_total / 64 + 1));
  p.kb_per_split = (kb_total + splits - 1) / splits;
  splits = (kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.idesc = cin_idesc(NT);
  p.desc_hi = cin_desc_hi();
  const size_t smem = static_cast<size_t>(kDwStages) * kDwStageBytes + 256 + 1024;
  static bool optin = false;
  if (!optin) {
    cudaFuncSetAttribute(cin_dw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         static_cast<int>(smem));
    optin = true;
  }
  dim3 grid(qtiles, splits);
  cin_dw_tc_kernel<<<grid, 256, smem, st>>>(tA0, tA1, tB0, tB1, p);
  return check_cuda(cudaGetLastError(), fn);
}

}  // namespace ctr
