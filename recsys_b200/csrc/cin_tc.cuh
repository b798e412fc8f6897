// xDeepFM CIN layer on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// Formulation: the CIN contraction
//     out[r,h] = sum_i X0t[r,i] * ( sum_j Xp[r,j] * W[i,j,h] )
// is a plain GEMM  T[r,(i,g)] = A[r,:] . Bp[(i,g),:]   (M = B*D rows, K = Hp, N = m*G)
// followed by a per-row contraction of T that never leaves the SM:
//   MODE_SCALE: out[r,g] (+)= act( sum_i x[r,i] * T[r,i,g] + bias[g] )   forward (A=Xp, x=X0t)
//                                                                      and dXp (A=dpre, x=X0t)
//   MODE_DOT  : out[r,i] (+)=      sum_g x[r,g] * T[r,i,g]              dX0t  (A=Xp, x=dpre)
// so the [M, m*G] product (2.6 GB at BASELINE config 3) is never written.
//
// One persistent CTA per SM, warp-specialised:
//   warp 0   TMA producer: A tile [128 x K] resident per M-tile, B k-blocks [NT x 32] through a ring
//   warp 1   MMA issuer:   tcgen05.mma.kind::tf32, M=128, N=NT<=256, accumulators in TMEM,
//                          two 256-column accumulator stages (all 512 TMEM columns)
//   warp 2   TMEM allocator
//   warps 4-7 epilogue:    tcgen05.ld 32x32b (thread = row), FMA contraction in registers
// 3xTF32 (prec 2) runs the passes (hi,hi),(lo,hi),(hi,lo) into the same accumulator.
#pragma once
#include "tc_common.cuh"

namespace ctr {

constexpr int kTcMaxG = 128;        // widest per-field group the epilogue keeps in registers
constexpr int MODE_SCALE = 0, MODE_DOT = 1;

struct CinTcParams {
  const float* xvec;  // MODE_SCALE: [M, ldx] with m values/row; MODE_DOT: [M, ldx] with G values/row
  int ldx;
  float* out;
  int ldo;
  const float* bias;
  int M, m, G, Gp, FPT, NT, n_ntiles, nkb, K, relu, accumulate, n_pass, NA, S;
  uint32_t idesc;
  uint64_t desc_hi;   // smem descriptor without the start address
};

// ------------------------------------------------------------------------ kernel
template <int MODE, int CW>
__global__ void __launch_bounds__(256, 1)
cin_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
              const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
              const CinTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;                                              // [NA][nkb] x 16 KB
  uint8_t* sB = sA + static_cast<size_t>(p.NA) * p.nkb * kTcABytes;  // [S] x 32 KB
  float* sX = reinterpret_cast<float*>(sB + static_cast<size_t>(p.S) * kTcBStageBytes);  // [m][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(
      reinterpret_cast<uint8_t*>(sX) + (MODE == MODE_SCALE ? static_cast<size_t>(p.m) * kTcBM * 4 : 0));
  uint64_t* full = bars;               // [S]
  uint64_t* empty = bars + 8;          // [S]   (S <= 8)
  uint64_t* a_full = bars + 16;
  uint64_t* a_empty = bars + 17;
  uint64_t* t_full = bars + 18;        // [2]
  uint64_t* t_empty = bars + 20;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_mtiles = (p.M + kTcBM - 1) / kTcBM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&t_full[s], 1);
      mbar_init(&t_empty[s], 4);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      uint32_t kbi = 0, mt = 0;
      for (int tile = blockIdx.x; tile < n_mtiles; tile += gridDim.x, ++mt) {
        mbar_wait(a_empty, (mt & 1) ^ 1);
        mbar_expect_tx(a_full, static_cast<uint32_t>(p.NA * p.nkb * kTcABytes));
        for (int pa = 0; pa < p.NA; ++pa)
          for (int kb = 0; kb < p.nkb; ++kb)
            tma_load_2d(sA + static_cast<size_t>(pa * p.nkb + kb) * kTcABytes, pa == 0 ? &tmA0 : &tmA1,
                        kb * kTcKB, tile * kTcBM, a_full);
        for (int nt = 0; nt < p.n_ntiles; ++nt)
          for (int ps = 0; ps < p.n_pass; ++ps)
            for (int kb = 0; kb < p.nkb; ++kb, ++kbi) {
              const uint32_t st = kbi % p.S, ph = (kbi / p.S) & 1;
              mbar_wait(&empty[st], ph ^ 1);
              mbar_expect_tx(&full[st], static_cast<uint32_t>(p.NT * kTcKB * 4));
              tma_load_2d(sB + static_cast<size_t>(st) * kTcBStageBytes, ps == 2 ? &tmB1 : &tmB0,
                          kb * kTcKB, nt * p.NT, &full[st]);
            }
      }
    }
  } else if (warp == 1) {
    // ======================================================= MMA issuer
    if (lane == 0) {
      uint32_t kbi = 0, nti = 0, mt = 0;
      for (int tile = blockIdx.x; tile < n_mtiles; tile += gridDim.x, ++mt) {
        mbar_wait(a_full, mt & 1);
        tc_fence_after();
        for (int nt = 0; nt < p.n_ntiles; ++nt, ++nti) {
          const uint32_t as = nti & 1, aph = (nti >> 1) & 1;
          mbar_wait(&t_empty[as], aph ^ 1);
          tc_fence_after();
          const uint32_t tacc = tmem_base + as * 256;
          uint32_t accum = 0;
          for (int ps = 0; ps < p.n_pass; ++ps) {
            const int pa = (ps == 1) ? 1 : 0;   // passes: (hi,hi) (lo,hi) (hi,lo)
            for (int kb = 0; kb < p.nkb; ++kb, ++kbi) {
              const uint32_t st = kbi % p.S, ph = (kbi / p.S) & 1;
              mbar_wait(&full[st], ph);
              tc_fence_after();
              const uint32_t a_addr = smem_u32(sA + static_cast<size_t>(pa * p.nkb + kb) * kTcABytes);
              const uint32_t b_addr = smem_u32(sB + static_cast<size_t>(st) * kTcBStageBytes);
              const int nk = min(4, (p.K - kb * kTcKB + 7) / 8);
              for (int k = 0; k < nk; ++k) {
                const uint64_t ad = p.desc_hi | static_cast<uint64_t>(((a_addr + k * 32) >> 4) & 0x3FFF);
                const uint64_t bd = p.desc_hi | static_cast<uint64_t>(((b_addr + k * 32) >> 4) & 0x3FFF);
                tc_mma_tf32(tacc, ad, bd, p.idesc, accum);
                accum = 1;
              }
              tc_commit(&empty[st]);   // frees the B stage once these MMAs have read it
            }
          }
          tc_commit(&t_full[as]);      // accumulator stage complete -> epilogue
        }
        tc_commit(a_empty);            // A tile no longer needed
      }
    }
  } else if (warp >= 4) {
    // ========================================================= epilogue
    const int quarter = warp & 3;
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    uint32_t nti = 0;
    for (int tile = blockIdx.x; tile < n_mtiles; tile += gridDim.x) {
      const int rl = quarter * 32 + lane;
      const int r = tile * kTcBM + rl;
      const bool valid = r < p.M;
      float acc[kTcMaxG];   // MODE_SCALE: accumulators; MODE_DOT: the row of x
#pragma unroll
      for (int g = 0; g < kTcMaxG; ++g) acc[g] = 0.f;
      if (MODE == MODE_SCALE) {
        for (int i = 0; i < p.m; ++i)
          sX[i * kTcBM + rl] = valid ? __ldg(p.xvec + static_cast<size_t>(r) * p.ldx + i) : 0.f;
      } else {
#pragma unroll
        for (int g = 0; g < kTcMaxG; ++g)
          if (g < p.G && valid) acc[g] = __ldg(p.xvec + static_cast<size_t>(r) * p.ldx + g);
      }
      for (int nt = 0; nt < p.n_ntiles; ++nt, ++nti) {
        const uint32_t as = nti & 1, aph = (nti >> 1) & 1;
        mbar_wait(&t_full[as], aph);
        tc_fence_after();
        for (int fl = 0; fl < p.FPT; ++fl) {
          const int i = nt * p.FPT + fl;
          if (i >= p.m) break;
          const uint32_t tcol = tmem_base + lane_base + as * 256 + fl * p.Gp;
          const float xv = MODE == MODE_SCALE ? sX[i * kTcBM + rl] : 0.f;
          float dot = 0.f;
#pragma unroll
          for (int c = 0; c < kTcMaxG / CW; ++c) {
            if (c * CW < p.Gp) {
              float v[CW];
              tc_ld<CW>(tcol + c * CW, v);
#pragma unroll
              for (int t = 0; t < CW; ++t) {
                if (MODE == MODE_SCALE) acc[c * CW + t] = fmaf(xv, v[t], acc[c * CW + t]);
                else dot = fmaf(acc[c * CW + t], v[t], dot);
              }
            }
          }
          if (MODE == MODE_DOT && valid) {
            float* o = p.out + static_cast<size_t>(r) * p.ldo + i;
            *o = p.accumulate ? (*o + dot) : dot;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[as]);
      }
      if (MODE == MODE_SCALE && valid) {
        float* o = p.out + static_cast<size_t>(r) * p.ldo;
#pragma unroll
        for (int g = 0; g < kTcMaxG; ++g) {
          if (g < p.G) {
            float v = acc[g];
            if (p.bias != nullptr) v += __ldg(p.bias + g);
            if (p.relu) v = fmaxf(v, 0.f);
            o[g] = p.accumulate ? (o[g] + v) : v;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512)
                 : "memory");
  }
}

// ------------------------------------------------------------------ prep kernels
// The tensor core reads fp32 operands as tf32 by dropping the low 13 mantissa bits
// (truncation, biased).  Operands are therefore pre-rounded to nearest tf32 (cvt.rna) so
// that what the MMA truncates is already zero; the 3xTF32 split keeps the remainder.
// Bp[(i*Gp + g)*Kp + k] = (g < G && k < K) ? W[i*slab + g*sg + k*sk] : 0, for i < m_pad fields.
__global__ void cin_prep_b_kernel(const float* __restrict__ W, float* __restrict__ Bp,
                                  float* __restrict__ Bp_lo, int m, int m_pad, int G, int Gp, int K,
                                  int Kp, long long slab, long long sg, long long sk) {
  const long long n = static_cast<long long>(m_pad) * Gp * Kp;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < n;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(e % Kp);
    const long long ig = e / Kp;
    const int g = static_cast<int>(ig % Gp);
    const int i = static_cast<int>(ig / Gp);
    float w = 0.f;
    if (i < m && g < G && k < K) w = __ldg(W + i * slab + g * sg + k * sk);
    const float hi = round_tf32(w);
    Bp[e] = hi;
    if (Bp_lo != nullptr) Bp_lo[e] = round_tf32(w - hi);
  }
}
// A [M, K] pitch lda -> hi (and lo) [M, Kp] pitch Kp, zero padded: 16-byte aligned rows for TMA.
__global__ void cin_prep_a_kernel(const float* __restrict__ A, int lda, int K, float* __restrict__ hi,
                                  float* __restrict__ lo, int Kp, long long M) {
  const long long n = M * Kp;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < n;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = e / Kp;
    const int k = static_cast<int>(e - r * Kp);
    const float v = k < K ? A[r * lda + k] : 0.f;
    const float h = round_tf32(v);
    hi[e] = h;
    if (lo != nullptr) lo[e] = round_tf32(v - h);
  }
}

// ------------------------------------------------------------------- host side
static bool cin_dw_tc_supported(int H);
static int64_t cin_dw_tc_ws(int M, int m, int Hp, int H, int prec);
static bool cin_dw_use_fused(int Hp, int H);
static int64_t cin_dw_fused_ws(int M, int m, int Hp, int H, int prec);
struct CinTcPlan {
  int Gp, FPT, NT, n_ntiles, m_pad, nkb, Kp, NA, S, n_pass;
  size_t smem;
  bool ok;
};

// G = per-field group width of the GEMM's N axis, K = reduction length, scale = MODE_SCALE.
static CinTcPlan cin_tc_plan(int m, int G, int K, bool scale, int prec) {
  CinTcPlan pl{};
  pl.Gp = (G + 7) / 8 * 8;
  pl.ok = pl.Gp <= kTcMaxG && K <= 256;
  if (!pl.ok) return pl;
  int fpt = 256 / pl.Gp;
  while (fpt > 1 && (fpt * pl.Gp) % 16 != 0) --fpt;
  if ((fpt * pl.Gp) % 16 != 0) {   // single odd-multiple-of-8 group: pad the group itself
    pl.Gp = (pl.Gp + 15) / 16 * 16;
    fpt = 256 / pl.Gp;
  }
  fpt = std::min(fpt, m);
  while (fpt > 1 && (fpt * pl.Gp) % 16 != 0) --fpt;
  pl.FPT = fpt;
  pl.NT = fpt * pl.Gp;
  pl.ok = pl.Gp <= kTcMaxG && (pl.NT % 16) == 0 && pl.NT >= 16;
  pl.n_ntiles = (m + fpt - 1) / fpt;
  pl.m_pad = pl.n_ntiles * fpt;
  pl.nkb = (K + kTcKB - 1) / kTcKB;
  pl.Kp = (K + 3) / 4 * 4;
  pl.NA = prec == CTR_CIN_TF32X3 ? 2 : 1;
  pl.n_pass = prec == CTR_CIN_TF32X3 ? 3 : 1;
  const size_t fixed = static_cast<size_t>(pl.NA) * pl.nkb * kTcABytes +
                       (scale ? static_cast<size_t>(m) * kTcBM * 4 : 0) + 256 + 1024;
  const size_t budget = 227 * 1024;
  int S = fixed < budget ? static_cast<int>((budget - fixed) / kTcBStageBytes) : 0;
  pl.S = std::min(S, 8);
  pl.ok = pl.ok && pl.S >= 2;
  pl.smem = fixed + static_cast<size_t>(pl.S) * kTcBStageBytes;
  return pl;
}

// One GEMM+epilogue pass.  A [M, K] (pitch lda), W viewed as Wg[i][g][k] with strides, x per MODE.
static int cin_tc_pass(int mode, const float* A, int lda, int K, const float* W, long long slab,
                       long long sg, long long sk, int m, int G, const float* xvec, int ldx,
                       const float* bias, int relu, int accumulate, float* out, int ldo, int M,
                       int prec, void* ws, int64_t ws_bytes, cudaStream_t st, const char* fn) {
  const CinTcPlan pl = cin_tc_plan(m, G, K, mode == MODE_SCALE, prec);
  if (!pl.ok) return fail_arg(fn, "shape not supported by the tensor-core path (need G<=128, K<=256)");
  const size_t bp_elems = static_cast<size_t>(pl.m_pad) * pl.Gp * pl.Kp;
  const size_t a_elems = static_cast<size_t>(M) * pl.Kp;
  const bool split = prec == CTR_CIN_TF32X3;
  const size_t need = (bp_elems + a_elems) * 4 * (split ? 2 : 1) + 1024;
  CTR_REQUIRE(ws != nullptr && static_cast<size_t>(ws_bytes) >= need, fn, "workspace too small");
  auto align256 = [](void* q) {
    return reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(q) + 255) & ~uintptr_t(255));
  };
  float* Bp = align256(ws);
  float* Bp_lo = split ? align256(Bp + bp_elems) : nullptr;
  float* A_hi = align256((split ? Bp_lo : Bp) + bp_elems);
  float* A_lo = split ? align256(A_hi + a_elems) : nullptr;
  {
    const long long n = static_cast<long long>(bp_elems);
    const int grid = static_cast<int>(std::min<long long>((n + 255) / 256, sm_count() * 8LL));
    cin_prep_b_kernel<<<grid, 256, 0, st>>>(W, Bp, Bp_lo, m, pl.m_pad, G, pl.Gp, K, pl.Kp, slab, sg, sk);
    const long long na = static_cast<long long>(a_elems);
    const int g2 = static_cast<int>(std::min<long long>((na + 255) / 256, sm_count() * 8LL));
    cin_prep_a_kernel<<<g2, 256, 0, st>>>(A, lda, K, A_hi, A_lo, pl.Kp, M);
  }
  CUtensorMap tA0, tA1, tB0, tB1;
  int r = make_map(&tA0, A_hi, M, K, pl.Kp, kTcBM);
  if (r != CTR_OK) return r;
  r = make_map(&tA1, split ? A_lo : A_hi, M, K, pl.Kp, kTcBM);
  if (r != CTR_OK) return r;
  r = make_map(&tB0, Bp, static_cast<long long>(pl.m_pad) * pl.Gp, K, pl.Kp, pl.NT);
  if (r != CTR_OK) return r;
  r = make_map(&tB1, split ? Bp_lo : Bp, static_cast<long long>(pl.m_pad) * pl.Gp, K, pl.Kp, pl.NT);
  if (r != CTR_OK) return r;

  CinTcParams p;
  p.xvec = xvec; p.ldx = ldx; p.out = out; p.ldo = ldo; p.bias = bias;
  p.M = M; p.m = m; p.G = G; p.Gp = pl.Gp; p.FPT = pl.FPT; p.NT = pl.NT; p.n_ntiles = pl.n_ntiles;
  p.nkb = pl.nkb; p.K = K; p.relu = relu; p.accumulate = accumulate; p.n_pass = pl.n_pass;
  p.NA = pl.NA; p.S = pl.S; p.idesc = cin_idesc(pl.NT); p.desc_hi = cin_desc_hi();
  const int n_mtiles = (M + kTcBM - 1) / kTcBM;
  const int grid = std::min(n_mtiles, sm_count());
  const bool cw32 = (pl.Gp % 32) == 0;
#define CTR_TC_LAUNCH(MODE_, CW_)                                                              \
  {                                                                                            \
    cudaFuncSetAttribute(cin_tc_kernel<MODE_, CW_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                         static_cast<int>(pl.smem));                                           \
    cin_tc_kernel<MODE_, CW_><<<grid, 256, pl.smem, st>>>(tA0, tA1, tB0, tB1, p);              \
  }
  if (mode == MODE_SCALE) {
    if (cw32) CTR_TC_LAUNCH(MODE_SCALE, 32) else CTR_TC_LAUNCH(MODE_SCALE, 8)
  } else {
    if (cw32) CTR_TC_LAUNCH(MODE_DOT, 32) else CTR_TC_LAUNCH(MODE_DOT, 8)
  }
#undef CTR_TC_LAUNCH
  return check_cuda(cudaGetLastError(), fn);
}

static int64_t cin_tc_pass_ws(int M, int m, int G, int K, bool scale, int prec) {
  const CinTcPlan pl = cin_tc_plan(m, G, K, scale, prec);
  if (!pl.ok) return 0;
  const size_t bp = static_cast<size_t>(pl.m_pad) * pl.Gp * pl.Kp;
  const size_t a = static_cast<size_t>(M) * pl.Kp;
  return static_cast<int64_t>((bp + a) * 4 * (prec == CTR_CIN_TF32X3 ? 2 : 1) + 1024);
}

static int64_t cin_tc_workspace_bytes(int B, int D, int m, int Hp, int H, int prec) {
  const int M = B * D;
  int64_t a = cin_tc_pass_ws(M, m, H, Hp, true, prec);   // fwd / dX0t: A = Xp, K = Hp, G = H
  int64_t b = cin_tc_pass_ws(M, m, Hp, H, true, prec);   // dXp: A = dpre, K = H, G = Hp
  int64_t c = 0;                                          // dW: XpT + X0T + dpreT (fused) or ZT + dpreT
  if (cin_dw_use_fused(Hp, H)) c = cin_dw_fused_ws(M, m, Hp, H, prec);
  else if (cin_dw_tc_supported(H)) c = cin_dw_tc_ws(M, m, Hp, H, prec);
  return std::max(std::max(a, b), c) + 1024;
}

static int cin_tc_layer_fwd(const float* X0t, int ld0, const float* Xp, int ldp, const float* W,
                            const float* bias, int B, int D, int m, int Hp, int H, float* out,
                            int prec, void* ws, int64_t ws_bytes, cudaStream_t st) {
  // T[r,(i,h)] = sum_j Xp[r,j] * W[(i*Hp+j)*H + h]: g = h (stride 1), k = j (stride H)
  return cin_tc_pass(MODE_SCALE, Xp, ldp, Hp, W, static_cast<long long>(Hp) * H, 1, H, m, H, X0t,
                     ld0, bias, 1, 0, out, H, B * D, prec, ws, ws_bytes, st, "ctr_cin_layer_fwd");
}

static bool cin_dw_tc_supported(int H);
static int64_t cin_dw_tc_ws(int M, int m, int Hp, int H, int prec);
static int cin_dw_tc(const float* X0t, int ld0, const float* Xp, int ldp, const float* dpre, int M,
                     int m, int Hp, int H, float* dW, int prec, void* ws, int64_t ws_bytes,
                     cudaStream_t st, const char* fn);
// the same GEMM with the A operand built in shared memory (cin_dw_fused.cuh): no Z in global memory
static bool cin_dw_fused_supported(int Hp, int H);
static int64_t cin_dw_fused_ws(int M, int m, int Hp, int H, int prec);
static int cin_dw_fused(const float* X0t, int ld0, const float* Xp, int ldp, const float* dpre, int M,
                        int m, int Hp, int H, float* dW, int prec, void* ws, int64_t ws_bytes,
                        cudaStream_t st, const char* fn);
static bool cin_dw_use_fused(int Hp, int H) {
  const char* e = ctr_knob("CTR_CIN_DW_FUSED");          // developer builds: 0 = the ZT path
  if (e != nullptr && atoi(e) == 0) return false;
  return cin_dw_fused_supported(Hp, H);
}

static int cin_tc_layer_bwd(const float* X0t, int ld0, const float* Xp, int ldp, const float* W,
                            const float* dpre, int B, int D, int m, int Hp, int H, float* dX0t,
                            float* dXp, float* dW, float* dbias, int prec, void* ws,
                            int64_t ws_bytes, cudaStream_t st) {
  const int M = B * D;
  const long long HpH = static_cast<long long>(Hp) * H;
  int r;
  if (dXp != nullptr) {
    // U[r,(i,j)] = sum_h dpre[r,h] * W[(i*Hp+j)*H + h]: g = j (stride H), k = h (stride 1);
    // dXp[r,j] += sum_i X0t[r,i] * U[r,i,j]
    r = cin_tc_pass(MODE_SCALE, dpre, H, H, W, HpH, H, 1, m, Hp, X0t, ld0, nullptr, 0, 1, dXp, ldp, M,
                    prec, ws, ws_bytes, st, "ctr_cin_layer_bwd");
    if (r != CTR_OK) return r;
  }
  if (dX0t != nullptr) {
    // T as in the forward; dX0t[r,i] += sum_h dpre[r,h] * T[r,i,h]
    r = cin_tc_pass(MODE_DOT, Xp, ldp, Hp, W, HpH, 1, H, m, H, dpre, H, nullptr, 0, 1, dX0t, ld0, M,
                    prec, ws, ws_bytes, st, "ctr_cin_layer_bwd");
    if (r != CTR_OK) return r;
  }
  // dW: split-K tcgen05 GEMM over the rows (cin_dw_tc.cuh); fp32 CUDA cores when H > 128
  if (cin_dw_use_fused(Hp, H)) {
    r = cin_dw_fused(X0t, ld0, Xp, ldp, dpre, M, m, Hp, H, dW, prec, ws, ws_bytes, st, "ctr_cin_layer_bwd");
    if (r != CTR_OK) return r;
  } else if (cin_dw_tc_supported(H)) {
    r = cin_dw_tc(X0t, ld0, Xp, ldp, dpre, M, m, Hp, H, dW, prec, ws, ws_bytes, st, "ctr_cin_layer_bwd");
    if (r != CTR_OK) return r;
  } else {
    const int Kq = m * Hp;
    int splits = std::max(1, std::min(64, (sm_count() * 4) / (((Kq + 63) / 64) * ((H + 63) / 64))));
    int rps = (M + splits - 1) / splits;
    rps = (rps + kCinKC - 1) / kCinKC * kCinKC;
    splits = (M + rps - 1) / rps;
    dim3 grid((Kq + 63) / 64, (H + 63) / 64, splits);
    cin_dw_kernel<<<grid, 256, 0, st>>>(X0t, ld0, Xp, ldp, dpre, M, m, Hp, H, dW, rps);
  }
  if (dbias != nullptr) {
    dim3 grid((H + 31) / 32, std::max(1, std::min(1024, (M + 63) / 64)));
    cin_colsum_kernel<<<grid, 256, 0, st>>>(dpre, M, H, dbias);
  }
  return check_cuda(cudaGetLastError(), "ctr_cin_layer_bwd");
}

}  // namespace ctr
