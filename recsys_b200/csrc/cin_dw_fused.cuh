// CIN weight gradient on tcgen05 without materialising Z:
//     dW[(i,j), h] = sum_r X0t[r,i] * Xp[r,j] * dpre[r,h]            (xdeepfm/xdeepfm.py:145-169, backward)
// is the split-K GEMM  dW[q, h] += ZT[q, r] . dpreT[h, r]  with ZT[(i,j), r] = X0t[r,i] * Xp[r,j].
// The unfused path (cin_dw_tc.cuh) writes ZT to global memory first - [m*Hp, B*D] fp32, 2.6 GB
// per layer at BASELINE config 3, twice that with the 3xTF32 lo copy - and streams it back.  Here
// the A operand is BUILT IN SHARED MEMORY per k-block (32 rows r) from two small TMA tiles:
//   XP  = XpT[0..127, r-block]   (every j of the layer, K-major, 128-byte swizzle; Hp <= 128)
//   X0  = X0T[i_lo..i_lo+15, r-block]   (the <= 16 values of i a 128-row q-tile can span)
//   A_hi[(row), r] = X0[i(row)][r] * XP[j(row)][r]   (q0 + row = i*Hp + j),  A_lo = tcg_lo(A_hi)
// by six otherwise idle warps while the tensor core works on the previous k-block; B = dpreT
// arrives pre-split (hi / lo) by TMA.  Per k-block: (hi,hi), (lo,hi), (hi,lo) = 12 MMAs for
// 3xTF32, 4 for plain tf32.  Only the three transposed copies XpT, X0T, dpreT (64 MB each at config
// 3) touch global memory.
//   warp 0   TMA producer          warp 1   MMA issuer          warp 2   TMEM allocator
//   warps 2-7  A-tile builders      warps 4-7  epilogue: TMEM -> vector RED into dW
#pragma once
#include "cin_tc.cuh"
#include "tc_gemm.cuh"

namespace ctr {

constexpr int kDfStages = 3;
constexpr int kDfX0Rows = 16;
constexpr int kDfXpBytes = kTcABytes;                       // 128 rows x 128 B
constexpr int kDfX0Bytes = kDfX0Rows * kTcKB * 4;           // 2 KB
constexpr int kDfBBytes = 128 * kTcKB * 4;                  // NT <= 128 rows x 128 B
constexpr int kDfStageBytes = kDfXpBytes + kDfX0Bytes + 2 * kDfBBytes + 1024 - kDfX0Bytes % 1024;
constexpr int kDfASlotBytes = 2 * kTcABytes;                // A_hi | A_lo
constexpr int kDfBuilders = 192;

struct CinDwFusedParams {
  float* dW;
  int Kq, Hp, H, NT, M, n_pass;
  int kb_per_split;
  uint32_t idesc;
  uint64_t desc_hi;
};

// physical byte offset of 16-byte chunk c of row `row` in a K-major 128-byte-swizzled tile
__device__ __forceinline__ uint32_t sw128(int row, int c) {
  return static_cast<uint32_t>(row) * 128u + (static_cast<uint32_t>(c ^ (row & 7)) << 4);
}

__global__ void __launch_bounds__(256, 1)
cin_dw_fused_kernel(const __grid_constant__ CUtensorMap tmXp, const __grid_constant__ CUtensorMap tmX0,
                    const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
                    const CinDwFusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* a_ring = smem + static_cast<size_t>(kDfStages) * kDfStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_ring + 2 * kDfASlotBytes);
  uint64_t* full = bars;             // [kDfStages]  TMA landed
  uint64_t* empty = bars + 4;        // [kDfStages]  builders and MMAs are done with the stage
  uint64_t* conv = bars + 8;         // [2]          A slot built
  uint64_t* a_empty = bars + 10;     // [2]          MMAs that read the A slot retired
  uint64_t* t_full = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kTcBM;
  const int i_lo = q0 / p.Hp;
  const int kb_total = (p.M + kTcKB - 1) / kTcKB;
  const int kb_beg = blockIdx.y * p.kb_per_split;
  const int nkb = max(0, min(kb_total, kb_beg + p.kb_per_split) - kb_beg);
  const uint32_t b_bytes = static_cast<uint32_t>(p.NT) * kTcKB * 4;
  const bool split = p.n_pass == 3;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kDfStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&conv[s], kDfBuilders);
      mbar_init(&a_empty[s], 1);
    }
    mbar_init(t_full, 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(128)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------------ TMA producer
    for (int kb = 0; kb < nkb; ++kb) {
      const uint32_t st = kb % kDfStages, ph = (kb / kDfStages) & 1;
      mbar_wait(&empty[st], ph ^ 1);
      mbar_expect_tx(&full[st], kDfXpBytes + kDfX0Bytes + (split ? 2 : 1) * b_bytes);
      uint8_t* s = smem + static_cast<size_t>(st) * kDfStageBytes;
      const int rr = (kb_beg + kb) * kTcKB;
      tma_load_2d(s, &tmXp, rr, 0, &full[st]);
      tma_load_2d(s + kDfXpBytes, &tmX0, rr, i_lo, &full[st]);
      tma_load_2d(s + kDfXpBytes + 2048, &tmB0, rr, 0, &full[st]);
      if (split) tma_load_2d(s + kDfXpBytes + 2048 + kDfBBytes, &tmB1, rr, 0, &full[st]);
    }
  } else if (warp == 1 && lane == 0) {
    // -------------------------------------------------------------------- MMA issuer
    uint32_t accum = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      const uint32_t st = kb % kDfStages, ph = (kb / kDfStages) & 1;
      const uint32_t sl = kb & 1, lph = (kb >> 1) & 1;
      mbar_wait(&full[st], ph);
      mbar_wait(&conv[sl], lph);
      tc_fence_after();
      const uint32_t a_hi = smem_u32(a_ring + static_cast<size_t>(sl) * kDfASlotBytes);
      const uint32_t a_lo = a_hi + kTcABytes;
      const uint32_t b_hi = smem_u32(smem + static_cast<size_t>(st) * kDfStageBytes) + kDfXpBytes + 2048;
      const uint32_t b_lo = b_hi + kDfBBytes;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        tc_mma_tf32(tmem_base, p.desc_hi | (((a_hi + k * 32) >> 4) & 0x3FFF),
                    p.desc_hi | (((b_hi + k * 32) >> 4) & 0x3FFF), p.idesc, accum);
        accum = 1;
      }
      if (split) {
#pragma unroll
        for (int k = 0; k < 4; ++k)   // (lo, hi)
          tc_mma_tf32(tmem_base, p.desc_hi | (((a_lo + k * 32) >> 4) & 0x3FFF),
                      p.desc_hi | (((b_hi + k * 32) >> 4) & 0x3FFF), p.idesc, 1);
#pragma unroll
        for (int k = 0; k < 4; ++k)   // (hi, lo)
          tc_mma_tf32(tmem_base, p.desc_hi | (((a_hi + k * 32) >> 4) & 0x3FFF),
                      p.desc_hi | (((b_lo + k * 32) >> 4) & 0x3FFF), p.idesc, 1);
      }
      tc_commit(&a_empty[sl]);
      tc_commit(&empty[st]);
    }
    tc_commit(t_full);
  } else if (warp >= 2) {
    // ------------------------------------------ A-tile builders (warps 2-7), then the epilogue
    const int ctid = threadIdx.x - 64;
    // the chunks this thread builds are the same in every k-block: e = ctid + u*192 < 1024,
    // row = e / 8, chunk = e % 8; (i, j) of the row and the two source offsets are fixed
    constexpr int NU = (1024 + kDfBuilders - 1) / kDfBuilders;     // 6
    uint32_t off_a[NU], off_xp[NU], off_x0[NU];
    bool live[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const int e = ctid + u * kDfBuilders;
      const int row = e >> 3, c = e & 7;
      const int q = q0 + row;
      const int i = q / p.Hp, j = q - i * p.Hp;
      live[u] = e < 1024 && q < p.Kq;
      off_a[u] = e < 1024 ? sw128(row, c) : 0u;
      off_xp[u] = live[u] ? sw128(j, c) : 0u;
      off_x0[u] = live[u] ? sw128(i - i_lo, c) : 0u;
    }
    for (int kb = 0; kb < nkb; ++kb) {
      const uint32_t st = kb % kDfStages, ph = (kb / kDfStages) & 1;
      const uint32_t sl = kb & 1, lph = (kb >> 1) & 1;
      mbar_wait(&full[st], ph);
      mbar_wait(&a_empty[sl], lph ^ 1);
      const uint8_t* xp = smem + static_cast<size_t>(st) * kDfStageBytes;
      const uint8_t* x0 = xp + kDfXpBytes;
      uint8_t* ah = a_ring + static_cast<size_t>(sl) * kDfASlotBytes;
      uint8_t* al = ah + kTcABytes;
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        if (ctid + u * kDfBuilders >= 1024) continue;
        float4 a = f4_zero();
        if (live[u]) {
          const float4 x = *reinterpret_cast<const float4*>(xp + off_xp[u]);
          const float4 s = *reinterpret_cast<const float4*>(x0 + off_x0[u]);
          a = make_float4(x.x * s.x, x.y * s.y, x.z * s.z, x.w * s.w);
        }
        if (split) {
          // tcgen05 kind::tf32 reads the top 19 bits: the fp32 product is its own hi operand
          *reinterpret_cast<float4*>(ah + off_a[u]) = a;
          *reinterpret_cast<float4*>(al + off_a[u]) = tcg_lo4(a);
        } else {
          *reinterpret_cast<float4*>(ah + off_a[u]) =
              make_float4(round_tf32(a.x), round_tf32(a.y), round_tf32(a.z), round_tf32(a.w));
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&conv[sl]);
    }
    if (warp >= 4 && nkb > 0) {
      const int quarter = warp & 3;
      const int q = q0 + quarter * 32 + lane;
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128)
                 : "memory");
  }
}

// dst[c, r] = src[r, c] (raw fp32), src [M, C] with row pitch lds, dst pitch ldd.
__global__ void __launch_bounds__(256)
cin_transpose_raw_kernel(const float* __restrict__ src, int lds, int M, int Cn,
                         float* __restrict__ dst, long long ldd) {
  __shared__ float t[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int k = ty; k < 32; k += 8) {
    const int r = r0 + k, c = c0 + tx;
    t[k][tx] = (r < M && c < Cn) ? src[static_cast<size_t>(r) * lds + c] : 0.f;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int c = c0 + k, r = r0 + tx;
    if (c < Cn && r < M) dst[static_cast<size_t>(c) * ldd + r] = t[tx][k];
  }
}

static bool cin_dw_fused_supported(int Hp, int H) {
  // every j of the layer fits the 128-row XP tile; a 128-row q-tile spans <= 16 values of i
  return H <= 128 && H >= 8 && Hp <= 128 && (127 / Hp) + 2 <= kDfX0Rows;
}

static int64_t cin_dw_fused_ws(int M, int m, int Hp, int H, int prec) {
  const int64_t ldz = (static_cast<int64_t>(M) + 3) / 4 * 4;
  const int64_t rows = static_cast<int64_t>(Hp) + m + static_cast<int64_t>(H) * (prec == CTR_CIN_TF32X3 ? 2 : 1);
  return rows * ldz * 4 + 4096;
}

static int cin_dw_fused(const float* X0t, int ld0, const float* Xp, int ldp, const float* dpre, int M,
                        int m, int Hp, int H, float* dW, int prec, void* ws, int64_t ws_bytes,
                        cudaStream_t st, const char* fn) {
  const bool split = prec == CTR_CIN_TF32X3;
  const long long ldz = (static_cast<long long>(M) + 3) / 4 * 4;
  const int Kq = m * Hp;
  CTR_REQUIRE(ws != nullptr && ws_bytes >= cin_dw_fused_ws(M, m, Hp, H, prec), fn, "workspace too small");
  CTR_REQUIRE(ldz < (1LL << 31), fn, "too many rows for the tensor-core dW path");
  auto align256 = [](void* q) {
    return reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(q) + 255) & ~uintptr_t(255));
  };
  float* XpT = align256(ws);
  float* X0T = align256(XpT + static_cast<size_t>(Hp) * ldz);
  float* dT = align256(X0T + static_cast<size_t>(m) * ldz);
  float* dT_lo = split ? align256(dT + static_cast<size_t>(H) * ldz) : nullptr;
  {
    dim3 g1((M + 31) / 32, (Hp + 31) / 32), g2((M + 31) / 32, (m + 31) / 32), g3((M + 31) / 32, (H + 31) / 32);
    cin_transpose_raw_kernel<<<g1, 256, 0, st>>>(Xp, ldp, M, Hp, XpT, ldz);
    cin_transpose_raw_kernel<<<g2, 256, 0, st>>>(X0t, ld0, M, m, X0T, ldz);
    cin_transpose_kernel<<<g3, 256, 0, st>>>(dpre, M, H, dT, dT_lo, ldz);
  }
  const int NT = (H + 15) / 16 * 16;
  CUtensorMap tXp, tX0, tB0, tB1;
  int r = make_map(&tXp, XpT, Hp, M, static_cast<int>(ldz), kTcBM);
  if (r != CTR_OK) return r;
  r = make_map(&tX0, X0T, m, M, static_cast<int>(ldz), kDfX0Rows);
  if (r != CTR_OK) return r;
  r = make_map(&tB0, dT, H, M, static_cast<int>(ldz), NT);
  if (r != CTR_OK) return r;
  r = make_map(&tB1, split ? dT_lo : dT, H, M, static_cast<int>(ldz), NT);
  if (r != CTR_OK) return r;
  CinDwFusedParams p;
  p.dW = dW; p.Kq = Kq; p.Hp = Hp; p.H = H; p.NT = NT; p.M = M; p.n_pass = split ? 3 : 1;
  const int qtiles = (Kq + kTcBM - 1) / kTcBM;
  const int kb_total = (M + kTcKB - 1) / kTcKB;
  int splits = std::max(1, std::min(sm_count() / qtiles, kb_total / 64 + 1));
  p.kb_per_split = (kb_total + splits - 1) / splits;
  splits = (kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.idesc = cin_idesc(NT);
  p.desc_hi = cin_desc_hi();
  const size_t smem = static_cast<size_t>(kDfStages) * kDfStageBytes + 2 * kDfASlotBytes + 256 + 1024;
  static bool optin = false;
  if (!optin) {
    cudaFuncSetAttribute(cin_dw_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         static_cast<int>(smem));
    optin = true;
  }
  dim3 grid(qtiles, splits);
  cin_dw_fused_kernel<<<grid, 256, smem, st>>>(tXp, tX0, tB0, tB1, p);
  return check_cuda(cudaGetLastError(), fn);
}

}  // namespace ctr
