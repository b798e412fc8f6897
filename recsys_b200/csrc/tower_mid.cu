// The middle of the dense tower as ONE persistent kernel.
//
// At batch 4096 the tower of the CTR models (deepfm/deepfm.py:100-129) is a chain of tiny,
// strictly dependent pieces: dense -> BN(batch statistics) -> dropout per layer, dense(1),
// the logit/loss head, and the same chain backwards.  Every BN needs a whole-batch column
// reduction before the next piece can start, so as separate launches the chain costs one
// kernel boundary per reduction and per piece (14 launches + 5 memsets, ~70 us of ~200).
// Here everything between the first layer's GEMM (624 -> H0, tc_gemm.cuh) and the first layer's
// backward GEMMs runs in one cooperative launch; the whole-batch reductions are grid barriers:
//
//   F_l  (l = 1..L-1)  a_l = relu( drop(BN(a_{l-1})) . W_l + b_l ),  stats_l += colsums(a_l, a_l^2)
//   O                  y = relu( drop(BN(a_{L-1})) . w_out + b_out ); logit/prob/BCE head;
//                      d loss / d head inputs; d w_out, d b_out, d head params;
//                      dn_{L-1} = dpre_out * w_out * keep;  dbeta/dgamma_{L-1} += colsums
//   B_l  (l = L-1..1)  dpre_l = BN-backward(dn_l) * 1[a_l > 0]  (written out for the dW kernels),
//                      db_l += colsums;  dn_{l-1} = (dpre_l . W_l^T) * keep;  dbeta/dgamma_{l-1}
//   D_0                dpre_0, db_0
//
// A CTA owns 32-row tiles of the batch (the same tiles in every phase, so everything a phase
// reads from an earlier one - except the column sums - it wrote itself).  The hidden GEMMs
// (H <= 128) run on mma.sync.m16n8k8 TF32 with the 3xTF32 split (fp32-grade): the
// operands sit in shared memory as fp32 and are split while the fragments are loaded.  The
// weights of the next phase are staged before waiting on the barrier.
#include <algorithm>
#include <cstdlib>

#include "gemm_core.cuh"
#include "tower_common.cuh"

namespace ctr {

constexpr int kMidMaxL = CTR_TOWER_MID_MAX_LAYERS;
constexpr int kMidMaxH = 128;
constexpr int kMidBM = 32;
constexpr int kMidThreads = 512;
constexpr int kMidWarps = kMidThreads / 32;
constexpr int kMidPA = 132;   // activation-tile pitch = 4 (mod 32): conflict-free A fragments
constexpr int kMidPB = 136;   // weight pitch = 8 (mod 32): conflict-free B fragments, W and W^T

// Work split.  Element-wise passes: warp w owns rows w and w+16 of the 32-row tile, lane q owns
// columns 4q..4q+3 (one float4 and one Philox block per row; lanes with 4q >= H idle).  GEMMs:
// warp = (row half mh, 16-column slab nw), m16n8k8 fragments; the product goes through shared
// memory (To) so that the epilogue is again a float4 pass with coalesced global accesses.
struct MidSmem {
  float Ws[kMidMaxH * kMidPB];
  float Ta[kMidBM * kMidPA];
  float To[kMidBM * kMidPA];
  float mu[kMidMaxH], rs[kMidMaxH], sc[kMidMaxH], sh[kMidMaxH];   // BN of the input-side layer
  float gmu[kMidMaxH], grs[kMidMaxH], gc1[kMidMaxH], gc2[kMidMaxH], gsc[kMidMaxH];  // grad source
  float s0[2 * kMidMaxH];                    // layer-0 column sums assembled from stats0_part
  float vec[kMidMaxH];                       // bias of the layer / w_out
  float red[3][kMidWarps][kMidMaxH];         // per-warp column partials
  float scal[kMidWarps][8];
};

__device__ __forceinline__ float ldcg1(const float* p) { return __ldcg(p); }
__device__ __forceinline__ float4 ldcg4(const float* p) {
  return __ldcg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void sts4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 f4_mul(float4 a, float4 b) {
  return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
}
__device__ __forceinline__ float4 f4_fma4(float4 a, float4 b, float4 c) {   // a*b + c
  return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}

// Grid barrier on {count, generation}; every CTA of the (cooperative) grid calls it.  Release /
// acquire operations at gpu scope instead of __threadfence() (a sequentially consistent MEMBAR):
// bar.sync makes the CTA's writes visible to thread 0, whose release-add publishes them; the
// waiters' acquire-load of the generation word, followed by bar.sync, hands them to the CTA.
__device__ __forceinline__ unsigned mid_atom_add_acq_rel(unsigned* p, unsigned v) {
  unsigned old;
  asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ unsigned mid_ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// (Measured and rejected: a two-level arrival - 8 group counters, then a top counter - to avoid 128
// same-address atomics; the extra dependent L2 round trip cost more, 3.3 us against 2.4 us.)
__device__ __forceinline__ void mid_grid_barrier(unsigned* bar, unsigned nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned gen = mid_ld_acquire(bar + 1);
    if (mid_atom_add_acq_rel(bar, 1u) == nblocks - 1) {
      asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(bar), "r"(0u) : "memory");
      mid_atom_add_acq_rel(bar + 1, 1u);
    } else {
      while (mid_ld_acquire(bar + 1) == gen) {
      }
    }
  }
  __syncthreads();
}

// Ws[k][n] = W[k][n] (row-major [K, N]), zero padded to 16-row / 16-column multiples.  The global
// loads are issued in batches of 4 per thread before the first shared store, so that the staging
// costs ~2 L2 round trips instead of one per float4.
__device__ __forceinline__ void mid_load_weights(MidSmem& sm, const float* __restrict__ W, int K,
                                                 int N) {
  const int KR = (K + 15) & ~15, NQ = ((N + 15) & ~15) >> 2;
  const int total = KR * NQ;
  for (int e0 = threadIdx.x; e0 < total; e0 += 4 * kMidThreads) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * kMidThreads;
      const int k = e / NQ, n = (e % NQ) * 4;
      v[u] = (e < total && k < K && n < N) ? ldg4(W + static_cast<size_t>(k) * N + n) : f4_zero();
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * kMidThreads;
      if (e < total) sts4(&sm.Ws[(e / NQ) * kMidPB + (e % NQ) * 4], v[u]);
    }
  }
}

struct MidCtx {
  float inv_B, inv_keep, p;
  unsigned seed, step;
  int training, B;
};

__device__ __forceinline__ void mid_bn_consts(const MidSmem& sm, const ctr_tower_mid_args& A,
                                              const MidCtx& c, int l, int k, float* mu, float* rstd) {
  const int H = A.H[l];
  if (c.training) {
    const bool parts = l == 0 && A.stats0_part != nullptr;
    const float m = (parts ? sm.s0[k] : ldcg1(A.stats[l] + k)) * c.inv_B;
    const float v = fmaxf((parts ? sm.s0[H + k] : ldcg1(A.stats[l] + H + k)) * c.inv_B - m * m,
                          0.f);  // biased variance
    *mu = m;
    *rstd = rsqrtf(v + A.eps);
  } else {
    *mu = A.mean[l][k];
    *rstd = rsqrtf(A.var[l][k] + A.eps);
  }
}
__device__ __forceinline__ void mid_fill_bn(MidSmem& sm, const ctr_tower_mid_args& A,
                                            const MidCtx& c, int l) {
  for (int k = threadIdx.x; k < A.H[l]; k += kMidThreads) {
    float mu, rstd;
    mid_bn_consts(sm, A, c, l, k, &mu, &rstd);
    sm.mu[k] = mu;
    sm.rs[k] = rstd;
    sm.sc[k] = rstd * A.gamma[l][k];
    sm.sh[k] = A.beta[l][k];
  }
}
__device__ __forceinline__ void mid_fill_grad(MidSmem& sm, const ctr_tower_mid_args& A,
                                              const MidCtx& c, int l) {
  // threads from the top of the block, so that it overlaps mid_fill_bn (threads from the bottom)
  for (int k = kMidThreads - 1 - threadIdx.x; k < A.H[l]; k += kMidThreads) {
    float mu, rstd;
    mid_bn_consts(sm, A, c, l, k, &mu, &rstd);
    sm.gmu[k] = mu;
    sm.grs[k] = rstd;
    sm.gc1[k] = ldcg1(A.dbeta[l] + k) * c.inv_B;
    sm.gc2[k] = ldcg1(A.dgamma[l] + k) * c.inv_B;
    sm.gsc[k] = rstd * A.gamma[l][k];
  }
}

// To[32 x ncols] = Ta[32 x kdim] . Bop, 3xTF32 (a.b ~ a_lo.b_hi + a_hi.b_lo + a_hi.b_hi, operands split
// while the fragments are loaded).  TRANS = false: Bop[k][n] = Ws[k][n]; true: Bop[k][n] = Ws[n][k].
// Warp (mh, nw) computes rows 16 mh.. and columns 16 nw.. (idle when they lie beyond ncols).
template <bool TRANS>
__device__ __forceinline__ void mid_gemm(MidSmem& sm, int kdim8, int ncols) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int mh = warp >> 3, nb = (warp & 7) * 16;
  if (nb >= ncols) return;
  float acc[2][4];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[j][c] = 0.f;
  const float* arow = &sm.Ta[(mh * 16 + g) * kMidPA + t];
#pragma unroll 2
  for (int ks = 0; ks < kdim8; ks += 8) {
    float ah[4], al[4], bh[2][2], bl[2][2];
    split_tf32(arow[ks], ah[0], al[0]);
    split_tf32(arow[ks + 8 * kMidPA], ah[1], al[1]);
    split_tf32(arow[ks + 4], ah[2], al[2]);
    split_tf32(arow[ks + 8 * kMidPA + 4], ah[3], al[3]);
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int n = nb + nt * 8 + g;
      const float b0 = TRANS ? sm.Ws[n * kMidPB + ks + t] : sm.Ws[(ks + t) * kMidPB + n];
      const float b1 = TRANS ? sm.Ws[n * kMidPB + ks + t + 4] : sm.Ws[(ks + t + 4) * kMidPB + n];
      split_tf32(b0, bh[nt][0], bl[nt][0]);
      split_tf32(b1, bh[nt][1], bl[nt][1]);
    }
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      mma_tf32(acc[nt], al, bh[nt]);
      mma_tf32(acc[nt], ah, bl[nt]);
      mma_tf32(acc[nt], ah, bh[nt]);
    }
  }
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    float* o = &sm.To[(mh * 16 + g) * kMidPA + nb + nt * 8 + 2 * t];
    *reinterpret_cast<float2*>(o) = make_float2(acc[nt][0], acc[nt][1]);
    *reinterpret_cast<float2*>(o + 8 * kMidPA) = make_float2(acc[nt][2], acc[nt][3]);
  }
}

// x' = drop(BN(a)) for the float4 at columns k..k+3 of row r of layer l's stored activation
__device__ __forceinline__ float4 mid_bn_drop4(const MidSmem& sm, const MidCtx& c, int l, int r, int k,
                                               float4 a, float4* keep) {
  float4 v;
  v.x = fmaf(a.x - sm.mu[k], sm.sc[k], sm.sh[k]);
  v.y = fmaf(a.y - sm.mu[k + 1], sm.sc[k + 1], sm.sh[k + 1]);
  v.z = fmaf(a.z - sm.mu[k + 2], sm.sc[k + 2], sm.sh[k + 2]);
  v.w = fmaf(a.w - sm.mu[k + 3], sm.sc[k + 3], sm.sh[k + 3]);
  if (c.p > 0.f) {
    *keep = drop_scale4(c.seed, static_cast<unsigned>(l), c.step, r, k >> 2, c.p, c.inv_keep);
    v = f4_mul(v, *keep);
  } else {
    *keep = make_float4(1.f, 1.f, 1.f, 1.f);
  }
  return v;
}
__device__ __forceinline__ float4 mid_xhat4(const MidSmem& sm, int k, float4 a) {
  return make_float4((a.x - sm.mu[k]) * sm.rs[k], (a.y - sm.mu[k + 1]) * sm.rs[k + 1],
                     (a.z - sm.mu[k + 2]) * sm.rs[k + 2], (a.w - sm.mu[k + 3]) * sm.rs[k + 3]);
}
// dpre = BN-backward(dn) * 1[a > 0] for the float4 at columns n..n+3 (gradient-source tables)
__device__ __forceinline__ float4 mid_dpre4(const MidSmem& sm, int n, float4 a, float4 dn) {
  float4 v;
#define CTR_MID_DP(cc, o)                                                          \
  {                                                                                \
    const float xhat = (a.cc - sm.gmu[n + o]) * sm.grs[n + o];                     \
    v.cc = (dn.cc - sm.gc1[n + o] - xhat * sm.gc2[n + o]) * sm.gsc[n + o];         \
    v.cc = a.cc > 0.f ? v.cc : 0.f;                                                \
  }
  CTR_MID_DP(x, 0) CTR_MID_DP(y, 1) CTR_MID_DP(z, 2) CTR_MID_DP(w, 3)
#undef CTR_MID_DP
  return v;
}

// Sum the per-thread float4 column partials over the 16 warps and RED them into global memory.
// v[i] goes to dst[i][0..H) (nullptr: skipped).  Ends with a __syncthreads().
template <int NV>
__device__ __forceinline__ void mid_reduce_cols(MidSmem& sm, const float4 (&v)[NV], float* const (&dst)[NV],
                                                int H) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) sts4(&sm.red[i][warp][4 * lane], v[i]);
  __syncthreads();
  for (int e = threadIdx.x; e < NV * H; e += kMidThreads) {
    const int i = e / H, n = e % H;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kMidWarps; ++w) s += sm.red[i][w][n];
    if (dst[i] != nullptr && s != 0.f) red_add_f32(dst[i] + n, s);
  }
  __syncthreads();
}

// Optional phase profile: block 0 / thread 0 stamps %globaltimer (ns) at the phase boundaries.
__device__ __forceinline__ void mid_stamp(const ctr_tower_mid_args& A, int slot) {
  if (A.timing != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    A.timing[slot] = t;
  }
}

__global__ void __launch_bounds__(kMidThreads, 1)
tower_mid_kernel(const __grid_constant__ ctr_tower_mid_args A, const int B) {
  extern __shared__ __align__(16) uint8_t mid_smem[];
  MidSmem& sm = *reinterpret_cast<MidSmem*>(mid_smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ntiles = (B + kMidBM - 1) / kMidBM;
  const int L = A.L;
  MidCtx c;
  c.B = B;
  c.inv_B = 1.f / static_cast<float>(B);
  c.training = A.training;
  c.p = A.training ? A.p_drop : 0.f;
  c.inv_keep = c.p > 0.f ? 1.f / (1.f - c.p) : 1.f;
  c.seed = A.seed;
  c.step = A.state != nullptr ? adam_step_of(A.state) : 0u;
  int ws_layer = -1;
  const int k4 = 4 * lane;                      // this lane's columns in the element-wise passes
  // One tile per CTA (the usual case: B / 32 <= #SMs): everything a phase needs from an earlier
  // one is carried in registers (a_c: the stored activation of the "current" layer for this
  // thread's 2 rows x 4 columns, dn_c: the gradient arriving at it); the global copies are still
  // written (the weight-gradient kernels read them) but never read back here.
  const bool single = gridDim.x >= static_cast<unsigned>(ntiles);
  float4 a_c[2] = {f4_zero(), f4_zero()}, dn_c[2] = {f4_zero(), f4_zero()};
  float4 ap[2] = {f4_zero(), f4_zero()};
  mid_stamp(A, 0);
  if (c.training && A.stats0_part != nullptr) {
    // layer 0 came with its column sums as one block per producer cluster (ctr_embed_tower_fwd):
    // add them up in a fixed order (no atomics upstream, no barrier here)
    const int H0 = A.H[0];
    for (int k = tid; k < 2 * H0; k += kMidThreads) {
      float s = 0.f;
      for (int q = 0; q < A.n_stats0_part; ++q) s += ldcg1(A.stats0_part + static_cast<size_t>(q) * 2 * H0 + k);
      sm.s0[k] = s;
      // the totals also go where the per-layer kernels of the backward expect them
      if (blockIdx.x == 0 && A.stats[0] != nullptr) A.stats[0][k] = s;
    }
    __syncthreads();
  }
  if (A.pre0 != nullptr) {
    // ---------------------------------------------- layer 0 epilogue (split-K GEMM partial sums)
    const int N = A.H[0];
    float* __restrict__ out = A.act[0];
    float4 acc[2] = {f4_zero(), f4_zero()};
    if (k4 < N) {
      const float4 bias = ldg4(A.b[0] + k4);
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int r = tile * kMidBM + warp + 16 * h;
          if (r < B) {
            float4 v = f4_add(ldcg4(A.pre0 + static_cast<size_t>(r) * N + k4), bias);
            v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
            *reinterpret_cast<float4*>(out + static_cast<size_t>(r) * N + k4) = v;
            acc[0] = f4_add(acc[0], v);
            acc[1] = f4_fma4(v, v, acc[1]);
            a_c[h] = v;
          }
        }
      }
    }
    if (c.training) {
      float* const dst[2] = {A.stats[0], A.stats[0] + N};
      mid_reduce_cols<2>(sm, acc, dst, N);
      if (L > 1) {                       // the next phase stages its weights before waiting
        mid_load_weights(sm, A.W[1], A.H[0], A.H[1]);
        ws_layer = 1;
      }
      mid_grid_barrier(A.barrier, gridDim.x);                      // stats_0 complete
    }
  } else if (single) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = blockIdx.x * kMidBM + warp + 16 * h;
      if (r < B && k4 < A.H[0]) a_c[h] = ldcg4(A.act[0] + static_cast<size_t>(r) * A.H[0] + k4);
    }
  }

  // ------------------------------------------------------------- forward hidden layers
  for (int l = 1; l < L; ++l) {
    const int K = A.H[l - 1], N = A.H[l];
    __syncthreads();
    if (ws_layer != l) {
      mid_load_weights(sm, A.W[l], K, N);
      ws_layer = l;
    }
    for (int n = tid; n < N; n += kMidThreads) sm.vec[n] = A.b[l][n];
    if (c.training && l > 1) mid_grid_barrier(A.barrier, gridDim.x);   // stats_{l-1} complete
    mid_fill_bn(sm, A, c, l - 1);
    __syncthreads();
    const float* __restrict__ in = A.act[l - 1];
    float* __restrict__ out = A.act[l];
    float4 acc[2] = {f4_zero(), f4_zero()};                // column sums of a_l and a_l^2
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int r0 = tile * kMidBM;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int rr = warp + 16 * h, r = r0 + rr;
        if (k4 < ((K + 7) & ~7)) {
          float4 v = f4_zero(), keep;
          if (r < B && k4 < K)
            v = mid_bn_drop4(sm, c, l - 1, r, k4,
                             single ? a_c[h] : ldcg4(in + static_cast<size_t>(r) * K + k4), &keep);
          sts4(&sm.Ta[rr * kMidPA + k4], v);
        }
      }
      __syncthreads();
      mid_gemm<false>(sm, (K + 7) & ~7, N);
      __syncthreads();
      a_c[0] = a_c[1] = f4_zero();
      if (k4 < N) {
        const float4 bias = lds4(&sm.vec[k4]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int rr = warp + 16 * h, r = r0 + rr;
          if (r < B) {
            float4 v = f4_add(lds4(&sm.To[rr * kMidPA + k4]), bias);
            v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
            *reinterpret_cast<float4*>(out + static_cast<size_t>(r) * N + k4) = v;
            acc[0] = f4_add(acc[0], v);
            acc[1] = f4_fma4(v, v, acc[1]);
            a_c[h] = v;
          }
        }
      }
    }
    if (c.training) {
      float* const dst[2] = {A.stats[l], A.stats[l] + N};
      mid_reduce_cols<2>(sm, acc, dst, N);
    }
  }
  mid_stamp(A, 1);
  if (c.training && L > 1) mid_grid_barrier(A.barrier, gridDim.x);       // stats_{L-1} complete
  mid_stamp(A, 2);

  // ----------------------------------------------------- final dense(1, relu) + loss head
  {
    const int l = L - 1, K = A.H[l];
    const int C = A.C;
    // head inputs of this thread's rows (first tile): fetched ahead of the table fill
    float zpre[2][3], lpre[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = blockIdx.x * kMidBM + warp + 16 * h;
      lpre[h] = r < B ? A.labels[r] : 0.f;
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) zpre[h][cc] = (r < B && cc < C - 1) ? A.z[cc][r] : 0.f;
    }
    __syncthreads();
    mid_fill_bn(sm, A, c, l);
    for (int k = tid; k < K; k += kMidThreads) sm.vec[k] = A.w_out[k];
    __syncthreads();
    const float b_out = A.b_out[0];
    float hw[4] = {0.f, 0.f, 0.f, 0.f};
    for (int cc = 0; cc < C; ++cc) hw[cc] = A.hw[cc];
    const float hb = A.hb[0];
    const float b1 = (A.relu0 && C > 1) ? A.b1[0] : 0.f;
    const float* __restrict__ act = A.act[l];
    float* __restrict__ dn = A.dn[l];
    const bool on = k4 < K;
    const float4 w4 = on ? lds4(&sm.vec[k4]) : f4_zero();
    float4 acc[3] = {f4_zero(), f4_zero(), f4_zero()};     // d w_out, d beta, d gamma
    float s_loss = 0.f, s_hb = 0.f, s_b1 = 0.f, s_bout = 0.f, s_hw[4] = {0.f, 0.f, 0.f, 0.f};
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int r0 = tile * kMidBM;
      float4 a4[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = r0 + warp + 16 * h;
        a4[h] = single ? a_c[h]
                       : (on && r < B) ? ldcg4(act + static_cast<size_t>(r) * K + k4) : f4_zero();
        dn_c[h] = f4_zero();
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = r0 + warp + 16 * h;
        if (r >= B) continue;                               // warp-uniform
        float4 hv = f4_zero(), xh = f4_zero(), keep = f4_zero();
        if (on) {
          xh = mid_xhat4(sm, k4, a4[h]);
          hv = mid_bn_drop4(sm, c, l, r, k4, a4[h], &keep);
        }
        const float y = fmaxf(warp_sum(f4_dot(hv, w4)) + b_out, 0.f);
        float av[4];
        float logit = hb;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          av[cc] = 0.f;
          if (cc < C) {
            float v = y;
            if (cc < C - 1) {
              v = (tile == static_cast<int>(blockIdx.x) && cc < 3) ? zpre[h][cc] : A.z[cc][r];
              if (cc == 0 && A.relu0) v = fmaxf(v + b1, 0.f);
            }
            av[cc] = v;
            logit = fmaf(hw[cc], v, logit);
          }
        }
        const float zl = tile == static_cast<int>(blockIdx.x) ? lpre[h] : A.labels[r];
        const float pr = 1.f / (1.f + expf(-logit));
        // tf.nn.sigmoid_cross_entropy_with_logits: max(x,0) - x z + log1p(exp(-|x|))
        const float bce = fmaxf(logit, 0.f) - logit * zl + log1pf(expf(-fabsf(logit)));
        if (lane == 0) {
          if (A.y_out != nullptr) A.y_out[r] = y;
          if (A.logits != nullptr) A.logits[r] = logit;
          if (A.prob != nullptr) A.prob[r] = pr;
        }
        s_loss += bce;
        if (c.training) {
          const float dl = (pr - zl) * A.grad_scale;
          s_hb += dl;
          float dy = 0.f;
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            if (cc < C) {
              s_hw[cc] = fmaf(dl, av[cc], s_hw[cc]);
              float gg = dl * hw[cc];
              if (cc == 0 && A.relu0 && C > 1) {
                gg = av[0] > 0.f ? gg : 0.f;
                s_b1 += gg;
              }
              if (cc < C - 1) {
                if (lane == 0) A.dz[cc][r] = gg;
              } else {
                dy = gg;
              }
            }
          }
          const float dpo = y > 0.f ? dy : 0.f;
          s_bout += dpo;
          if (on) {
            const float4 d = make_float4(dpo * w4.x * keep.x, dpo * w4.y * keep.y, dpo * w4.z * keep.z,
                                         dpo * w4.w * keep.w);
            *reinterpret_cast<float4*>(dn + static_cast<size_t>(r) * K + k4) = d;
            dn_c[h] = d;
            acc[0] = f4_fma(dpo, hv, acc[0]);
            acc[1] = f4_add(acc[1], d);
            acc[2] = f4_fma4(d, xh, acc[2]);
          }
        }
      }
    }
    if (lane == 0) {
      sm.scal[warp][0] = s_loss; sm.scal[warp][1] = s_hb; sm.scal[warp][2] = s_b1;
      sm.scal[warp][3] = s_bout;
      sm.scal[warp][4] = s_hw[0]; sm.scal[warp][5] = s_hw[1]; sm.scal[warp][6] = s_hw[2];
      sm.scal[warp][7] = s_hw[3];
    }
    if (c.training) {
      float* const dst[3] = {A.dw_out, A.dbeta[l], A.dgamma[l]};
      mid_reduce_cols<3>(sm, acc, dst, K);
    } else {
      __syncthreads();
    }
    if (tid < 8) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kMidWarps; ++w) s += sm.scal[w][tid];
      if (tid == 0) {
        if (A.loss != nullptr) red_add_f32(A.loss, s * c.inv_B);
      } else if (c.training) {
        if (tid == 1) red_add_f32(A.dhb, s);
        else if (tid == 2) { if (A.relu0 && C > 1) red_add_f32(A.db1, s); }
        else if (tid == 3) red_add_f32(A.db_out, s);
        else if (tid - 4 < C) red_add_f32(A.dhw + (tid - 4), s);
      }
    }
  }
  mid_stamp(A, 3);
  if (!c.training) return;
  // the activation below the next backward layer is final: fetch it while waiting on the barrier
  auto prefetch_prev = [&](int lp) {
    if (!single || lp < 0) return;
    const int Kp = A.H[lp];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = blockIdx.x * kMidBM + warp + 16 * h;
      ap[h] = (r < B && k4 < Kp) ? ldcg4(A.act[lp] + static_cast<size_t>(r) * Kp + k4) : f4_zero();
    }
  };
  prefetch_prev(L - 2);
  mid_grid_barrier(A.barrier, gridDim.x);                          // dbeta/dgamma_{L-1} complete
  mid_stamp(A, 4);

  // ------------------------------------------------------------ backward hidden layers
  for (int l = L - 1; l >= 1; --l) {
    const int N = A.H[l], K = A.H[l - 1];
    __syncthreads();
    if (ws_layer != l) {
      mid_load_weights(sm, A.W[l], K, N);
      ws_layer = l;
    }
    mid_fill_grad(sm, A, c, l);
    mid_fill_bn(sm, A, c, l - 1);
    __syncthreads();
    const float* __restrict__ act = A.act[l];
    const float* __restrict__ dnl = A.dn[l];
    float* __restrict__ dpre = A.dpre[l];
    const float* __restrict__ aprev = A.act[l - 1];
    float* __restrict__ dnp = A.dn[l - 1];
    float4 acc[3] = {f4_zero(), f4_zero(), f4_zero()};     // d b_l, d beta_{l-1}, d gamma_{l-1}
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int r0 = tile * kMidBM;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int rr = warp + 16 * h, r = r0 + rr;
        if (!single)
          ap[h] = (r < B && k4 < K) ? ldcg4(aprev + static_cast<size_t>(r) * K + k4) : f4_zero();
        if (k4 < ((N + 7) & ~7)) {
          float4 v = f4_zero();
          if (r < B && k4 < N) {
            v = mid_dpre4(sm, k4, single ? a_c[h] : ldcg4(act + static_cast<size_t>(r) * N + k4),
                          single ? dn_c[h] : ldcg4(dnl + static_cast<size_t>(r) * N + k4));
            *reinterpret_cast<float4*>(dpre + static_cast<size_t>(r) * N + k4) = v;
            acc[0] = f4_add(acc[0], v);
          }
          sts4(&sm.Ta[rr * kMidPA + k4], v);
        }
      }
      __syncthreads();
      mid_gemm<true>(sm, (N + 7) & ~7, K);
      __syncthreads();
      dn_c[0] = dn_c[1] = f4_zero();
      if (k4 < K) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int rr = warp + 16 * h, r = r0 + rr;
          if (r < B) {
            float4 v = lds4(&sm.To[rr * kMidPA + k4]);
            if (c.p > 0.f)
              v = f4_mul(v, drop_scale4(c.seed, static_cast<unsigned>(l - 1), c.step, r, lane, c.p, c.inv_keep));
            *reinterpret_cast<float4*>(dnp + static_cast<size_t>(r) * K + k4) = v;
            acc[1] = f4_add(acc[1], v);
            acc[2] = f4_fma4(v, mid_xhat4(sm, k4, ap[h]), acc[2]);
            dn_c[h] = v;
          }
        }
      }
      a_c[0] = ap[0];                   // the layer below becomes the current one
      a_c[1] = ap[1];
    }
    float* const dst[3] = {A.dbias[l], A.dbeta[l - 1], A.dgamma[l - 1]};
    mid_reduce_cols<3>(sm, acc, dst, N > K ? N : K);
    prefetch_prev(l - 2);
    mid_stamp(A, 5);
    mid_grid_barrier(A.barrier, gridDim.x);                        // dbeta/dgamma_{l-1} complete
    mid_stamp(A, 6);
  }

  // ------------------------------------------------------------------ dpre of layer 0
  {
    const int N = A.H[0];
    __syncthreads();
    mid_fill_grad(sm, A, c, 0);
    __syncthreads();
    const float* __restrict__ act = A.act[0];
    const float* __restrict__ dnl = A.dn[0];
    float* __restrict__ dpre = A.dpre[0];
    float* __restrict__ dpre_lo = A.dpre0_lo;
    float4 acc[1] = {f4_zero()};
    if (k4 < N) {
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int r = tile * kMidBM + warp + 16 * h;
          if (r < B) {
            const float4 v = mid_dpre4(sm, k4,
                                       single ? a_c[h] : ldcg4(act + static_cast<size_t>(r) * N + k4),
                                       single ? dn_c[h] : ldcg4(dnl + static_cast<size_t>(r) * N + k4));
            *reinterpret_cast<float4*>(dpre + static_cast<size_t>(r) * N + k4) = v;
            if (dpre_lo != nullptr)
              *reinterpret_cast<float4*>(dpre_lo + static_cast<size_t>(r) * N + k4) = tcg_lo4(v);
            acc[0] = f4_add(acc[0], v);
          }
        }
      }
    }
    float* const dst[1] = {A.dbias[0]};
    mid_reduce_cols<1>(sm, acc, dst, N);
  }
  mid_stamp(A, 7);
}

}  // namespace ctr

using namespace ctr;

extern "C" {

int ctr_tower_mid(const ctr_tower_mid_args* a, int B, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(a != nullptr && B >= 0, "ctr_tower_mid", "null args");
  CTR_REQUIRE(a->L >= 1 && a->L <= kMidMaxL, "ctr_tower_mid", "need 1 <= L <= 4 hidden layers");
  CTR_REQUIRE(a->C >= 1 && a->C <= 4, "ctr_tower_mid", "need 1 <= C <= 4 head columns");
  for (int l = 0; l < a->L; ++l) {
    CTR_REQUIRE(a->H[l] >= 4 && a->H[l] <= kMidMaxH && (a->H[l] & 3) == 0, "ctr_tower_mid",
                "hidden widths must be multiples of 4 in [4, 128]");
    CTR_REQUIRE(a->act[l] && a->gamma[l] && a->beta[l], "ctr_tower_mid", "null act/gamma/beta");
    CTR_REQUIRE(aligned16(a->act[l]), "ctr_tower_mid", "activations must be 16-byte aligned");
    CTR_REQUIRE(l == 0 || (a->W[l] && a->b[l] && aligned16(a->W[l])), "ctr_tower_mid",
                "null / unaligned W or b");
    CTR_REQUIRE(l != 0 || a->pre0 == nullptr || (a->b[0] && aligned16(a->b[0]) && aligned16(a->pre0)),
                "ctr_tower_mid", "pre0 needs an aligned b[0]");
    if (a->training) {
      CTR_REQUIRE((a->stats[l] || (l == 0 && a->stats0_part)) && a->dn[l] && a->dpre[l] && a->dbeta[l] && a->dgamma[l] && a->dbias[l],
                  "ctr_tower_mid", "training needs stats/dn/dpre/dbeta/dgamma/dbias per layer");
      CTR_REQUIRE(aligned16(a->dn[l]) && aligned16(a->dpre[l]), "ctr_tower_mid",
                  "dn/dpre must be 16-byte aligned");
    } else {
      CTR_REQUIRE(a->mean[l] && a->var[l], "ctr_tower_mid", "eval needs moving mean/var");
    }
  }
  CTR_REQUIRE(a->w_out && a->b_out && a->hw && a->hb && a->labels, "ctr_tower_mid",
              "null w_out/b_out/hw/hb/labels");
  CTR_REQUIRE(!a->relu0 || a->C == 1 || a->b1, "ctr_tower_mid", "relu0 needs b1");
  for (int c = 0; c + 1 < a->C; ++c)
    CTR_REQUIRE(a->z[c] && (!a->training || a->dz[c]), "ctr_tower_mid", "null head column");
  if (a->training)
    CTR_REQUIRE(a->barrier && a->dhw && a->dhb && a->dw_out && a->db_out &&
                    (!a->relu0 || a->C == 1 || a->db1),
                "ctr_tower_mid", "training needs barrier and head/out gradient buffers");
  CTR_REQUIRE(a->stats0_part == nullptr || (a->pre0 == nullptr && a->n_stats0_part >= 1), "ctr_tower_mid",
              "stats0_part excludes pre0 and needs n_stats0_part >= 1");
  if (B == 0) return CTR_OK;
  static bool optin = false;
  const int smem = static_cast<int>(sizeof(MidSmem));
  if (!optin) {
    cudaFuncSetAttribute(tower_mid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    optin = true;
  }
  const int ntiles = (B + kMidBM - 1) / kMidBM;
  const int grid = std::min(ntiles, sm_count());      // 1 CTA per SM: all resident
  int Bv = B;
  ctr_tower_mid_args args = *a;
  void* kargs[] = {&args, &Bv};
  cudaError_t e;
  // Training needs every CTA resident (grid barriers): cooperative launch.  CTR_MID_COOP=0 uses a
  // plain launch instead, which is equally safe here (grid <= #SMs at 1 CTA/SM, and nothing that
  // could occupy an SM ever waits on this kernel), for drivers that cannot capture cooperative
  // launches into a CUDA graph.
  const bool coop = option_get("mid_coop", 1) != 0;
  if (a->training && coop) {
    e = cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(tower_mid_kernel), dim3(grid),
                                    dim3(kMidThreads), kargs, smem, static_cast<cudaStream_t>(stream));
  } else {
    e = cudaLaunchKernel(reinterpret_cast<const void*>(tower_mid_kernel), dim3(grid), dim3(kMidThreads),
                         kargs, smem, static_cast<cudaStream_t>(stream));
  }
  return check_cuda(e, "ctr_tower_mid");
}

}  // extern "C"
