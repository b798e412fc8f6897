// Fused dense tower + loss head of the CTR models, fp32 CUDA cores.
//
// The reference's tower is `dense(relu) -> batch_normalization -> dropout` per layer and a
// final `dense(1, relu)` (deepfm/deepfm.py:100-108, xdeepfm/xdeepfm.py:184-192,
// dcn/dcn.py:144-149), followed by `dense(concat[...], 1)`, sigmoid and the mean
// sigmoid-cross-entropy (deepfm/deepfm.py:110-129).  TF (and torch) run it as ~80 tiny
// kernels per step at batch 4096; here it is 3 + 1 launches forward and 6 backward:
//
//   tower_layer_fwd   out = relu( P(X) . W + b ),  column sums of out / out^2 for the next BN,
//                     P = identity | BN(batch stats or moving stats) + dropout of the previous
//                     activation, recomputed on the fly (the normalised/dropped tensors are
//                     never written)
//   loss_head         logit = sum_c hw[c] * act_c(z_c) + hb, prob, mean BCE, and - because it is
//                     the end of the graph - d(loss)/dz_c, d hw, d hb, d b1 in the same launch
//   tower_layer_bwd_data     dX' = dpre . W^T -> dropout/BN bookkeeping for the layer below
//   tower_layer_bwd_weights  dW += P(X)^T . dpre,  db += colsum(dpre)   (split over rows, RED)
//
// Dropout masks are counter based (Philox4x32-10 keyed by seed, layer and the device-side step
// counter), regenerated in the backward instead of stored.
#include <cstdlib>

#include "gemm_core.cuh"
#include "tc_gemm.cuh"
#include "tower_common.cuh"

namespace ctr {

// ------------------------------------------------------------------- prologue descriptors
// Normalise + drop a stored post-ReLU activation A[r,k] (k < K <= kMaxBn):
//   xhat = (A - mu) * rstd;  x' = (xhat * gamma + beta) * keep
struct BnDrop {
  const float* sums;    // train: [2][K] column sums of A and A^2 over the batch; else nullptr
  const float* mean;    // eval: moving mean / variance (sums == nullptr)
  const float* var;
  const float* gamma;
  const float* beta;
  const float* state;   // device step counter {t, lr_t} (dropout stream), nullable
  float inv_B, eps, p, inv_keep;
  unsigned seed, layer;
  int enabled;          // 0: identity
};
// Gradient source for a layer's post-ReLU output a[r,n]:
//   kind 0: g = G[r*ldg + n]                                     (given directly)
//   kind 1: g = BN-backward of the stored dn[r,n] through the BN that follows a:
//           g = gamma*rstd * (dn - dbeta/B - xhat*dgamma/B)      (train)   or gamma*rstd*dn (eval)
// then dpre = g * 1[a > 0].
struct GradSrc {
  const float* G;
  int ldg;
  const float* a;       // the layer's stored output (post-ReLU), [B, N]
  int lda;
  const float* sums;    // stats of a (train) / nullptr
  const float* mean;
  const float* var;
  const float* gamma;
  const float* dbeta;   // column sums of dn and dn*xhat (train)
  const float* dgamma;
  float inv_B, eps;
  int kind, train;
};

struct TowerSmem : MmaSmem {
  float mu[kMaxBn], sc[kMaxBn], sh[kMaxBn];                              // prologue tables
  float gmu[kMaxBn], grs[kMaxBn], gc1[kMaxBn], gc2[kMaxBn], gsc[kMaxBn];  // gradient-source tables
  float red[2][8][kTwBN];
};

__device__ __forceinline__ void fill_pro_tables(TowerSmem& sm, const BnDrop& pro, int K) {
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float mu, rstd;
    bn_consts(pro.sums, pro.mean, pro.var, k, K, pro.inv_B, pro.eps, &mu, &rstd);
    sm.mu[k] = mu;
    sm.sc[k] = rstd * pro.gamma[k];
    sm.sh[k] = pro.beta[k];
  }
}
__device__ __forceinline__ void fill_gs_tables(TowerSmem& sm, const GradSrc& g, int N) {
  if (g.kind != 1) return;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float mu, rstd;
    bn_consts(g.sums, g.mean, g.var, n, N, g.inv_B, g.eps, &mu, &rstd);
    sm.gmu[n] = mu;
    sm.grs[n] = rstd;
    sm.gc1[n] = g.train ? g.dbeta[n] * g.inv_B : 0.f;
    sm.gc2[n] = g.train ? g.dgamma[n] * g.inv_B : 0.f;
    sm.gsc[n] = rstd * g.gamma[n];
  }
}
// x' = P(X)[r,k] with the tables in shared memory
__device__ __forceinline__ float pro_value(const TowerSmem& sm, const BnDrop& pro, unsigned step,
                                           float v, int r, int k) {
  v = fmaf(v - sm.mu[k], sm.sc[k], sm.sh[k]);
  if (pro.p > 0.f) v *= drop_scale(pro.seed, pro.layer, step, r, k, pro.p, pro.inv_keep);
  return v;
}
// dpre[r,n] = g * 1[a > 0] with the tables in shared memory
__device__ __forceinline__ float dpre_value(const TowerSmem& sm, const GradSrc& g, int r, int n) {
  if (g.kind == 2) return g.G[static_cast<size_t>(r) * g.ldg + n];
  const float a = g.a[static_cast<size_t>(r) * g.lda + n];
  if (!(a > 0.f)) return 0.f;
  float v = g.G[static_cast<size_t>(r) * g.ldg + n];
  if (g.kind == 1) {
    const float xhat = (a - sm.gmu[n]) * sm.grs[n];
    v = (v - sm.gc1[n] - xhat * sm.gc2[n]) * sm.gsc[n];
  }
  return v;
}

__device__ __forceinline__ bool aligned16_dev(const void* p) {
  return (reinterpret_cast<uintptr_t>(p) & 15) == 0;
}
// dpre for 4 consecutive columns n..n+3 of row r (columns >= N read as 0)
__device__ __forceinline__ float4 dpre4(const TowerSmem& sm, const GradSrc& g, int r, int n, int N) {
  if (g.kind == 2)
    return load4_guard(g.G + static_cast<size_t>(r) * g.ldg + n, N - n,
                       (g.ldg & 3) == 0 && aligned16_dev(g.G));
  const bool al = (g.lda & 3) == 0 && (g.ldg & 3) == 0 && aligned16_dev(g.a) && aligned16_dev(g.G);
  const float4 a = load4_guard(g.a + static_cast<size_t>(r) * g.lda + n, N - n, al);
  float4 v = load4_guard(g.G + static_cast<size_t>(r) * g.ldg + n, N - n, al);
  if (g.kind == 1) {
#define CTR_DP1(c, o)                                                          \
    if (n + o < N) {                                                           \
      const float xhat = (a.c - sm.gmu[n + o]) * sm.grs[n + o];                \
      v.c = (v.c - sm.gc1[n + o] - xhat * sm.gc2[n + o]) * sm.gsc[n + o];      \
    }
    CTR_DP1(x, 0) CTR_DP1(y, 1) CTR_DP1(z, 2) CTR_DP1(w, 3)
#undef CTR_DP1
  }
  v.x = a.x > 0.f ? v.x : 0.f;
  v.y = a.y > 0.f ? v.y : 0.f;
  v.z = a.z > 0.f ? v.z : 0.f;
  v.w = a.w > 0.f ? v.w : 0.f;
  return v;
}

// --------------------------------------------------------------------------- forward
template <bool PRO, int RT>
__global__ void __launch_bounds__(256)
tower_layer_fwd_kernel(const float* __restrict__ X, int ldx, int K, const BnDrop pro,
                       const float* __restrict__ W, const float* __restrict__ bias, int N,
                       float* __restrict__ out, int ldo, float* __restrict__ stats, int relu, int B) {
  extern __shared__ __align__(16) uint8_t tw_smem[];
  TowerSmem& sm = *reinterpret_cast<TowerSmem*>(tw_smem);
  constexpr int MT = RT / 2;                 // RT = 2 -> 16-row tile, RT = 4 -> 32-row tile
  constexpr int BM = MT * 16;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int r0 = blockIdx.x * BM, n0 = blockIdx.y * kTwBN;
  unsigned step = 0;
  if (PRO) {
    fill_pro_tables(sm, pro, K);
    if (pro.state != nullptr) step = adam_step_of(pro.state);
    __syncthreads();
  }
  float acc[MT][2][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;
  const bool xal = (ldx & 3) == 0 && aligned16_dev(X);
  const bool wal = (N & 3) == 0 && aligned16_dev(W);
  auto fa4 = [&](int rr, int k) -> float4 {
    const int r = r0 + rr;
    if (r >= B || k >= K) return f4_zero();
    float4 v = load4_guard(X + static_cast<size_t>(r) * ldx + k, K - k, xal);
    if (PRO) {
      v.x = pro_value(sm, pro, step, v.x, r, k);
      if (k + 1 < K) v.y = pro_value(sm, pro, step, v.y, r, k + 1);
      if (k + 2 < K) v.z = pro_value(sm, pro, step, v.z, r, k + 2);
      if (k + 3 < K) v.w = pro_value(sm, pro, step, v.w, r, k + 3);
    }
    return v;
  };
  auto fb4 = [&](int k, int c) -> float4 {
    const int n = n0 + c;
    if (k >= K || n >= N) return f4_zero();
    return load4_guard(W + static_cast<size_t>(k) * N + n, N - n, wal);
  };
  gemm_tile_mma<MT, true, false>(sm, K, fa4, fb4, acc);

  // epilogue on the C fragments: rows r0 + mt*16 + g (+8), columns n0 + warp*16 + nt*8 + 2t (+1)
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    float cs[2] = {0.f, 0.f}, cq[2] = {0.f, 0.f};
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        const int r = r0 + mt * 16 + g + (ci >> 1) * 8;
        const int n = n0 + warp * 16 + nt * 8 + 2 * t + (ci & 1);
        if (r < B && n < N) {
          float v = acc[mt][nt][ci] + (bias != nullptr ? __ldg(bias + n) : 0.f);
          if (relu) v = fmaxf(v, 0.f);
          out[static_cast<size_t>(r) * ldo + n] = v;
          cs[ci & 1] += v;
          cq[ci & 1] = fmaf(v, v, cq[ci & 1]);
        }
      }
    if (stats != nullptr) {
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        cs[0] += __shfl_xor_sync(0xffffffffu, cs[0], o);
        cs[1] += __shfl_xor_sync(0xffffffffu, cs[1], o);
        cq[0] += __shfl_xor_sync(0xffffffffu, cq[0], o);
        cq[1] += __shfl_xor_sync(0xffffffffu, cq[1], o);
      }
      if (g == 0) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int n = n0 + warp * 16 + nt * 8 + 2 * t + j;
          if (n < N) {
            red_add_f32(stats + n, cs[j]);
            red_add_f32(stats + N + n, cq[j]);
          }
        }
      }
    }
  }
}

// N == 1 (the final dense(1, relu)): one warp per row.
__global__ void __launch_bounds__(256)
tower_out_fwd_kernel(const float* __restrict__ X, int ldx, int K, const BnDrop pro,
                     const float* __restrict__ W, const float* __restrict__ bias,
                     float* __restrict__ out, int ldo, int relu, int B) {
  __shared__ float s_mu[kMaxBn], s_sc[kMaxBn], s_sh[kMaxBn], s_w[kMaxBn];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool use_tab = K <= kMaxBn;
  if (use_tab) {
    for (int k = threadIdx.x; k < K; k += 256) {
      s_w[k] = W[k];
      if (pro.enabled) {
        float mu, rstd;
        bn_consts(pro.sums, pro.mean, pro.var, k, K, pro.inv_B, pro.eps, &mu, &rstd);
        s_mu[k] = mu;
        s_sc[k] = rstd * pro.gamma[k];
        s_sh[k] = pro.beta[k];
      }
    }
    __syncthreads();
  }
  const unsigned step = pro.state != nullptr ? adam_step_of(pro.state) : 0u;
  for (int r = blockIdx.x * 8 + warp; r < B; r += gridDim.x * 8) {
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) {
      float v = X[static_cast<size_t>(r) * ldx + k];
      if (pro.enabled) {
        v = fmaf(v - s_mu[k], s_sc[k], s_sh[k]);
        if (pro.p > 0.f) v *= drop_scale(pro.seed, pro.layer, step, r, k, pro.p, pro.inv_keep);
      }
      acc = fmaf(v, use_tab ? s_w[k] : W[k], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      float v = acc + (bias != nullptr ? bias[0] : 0.f);
      if (relu) v = fmaxf(v, 0.f);
      out[static_cast<size_t>(r) * ldo] = v;
    }
  }
}

// BN + dropout of a stored activation written out (towers without a final dense layer, DCN).
__global__ void __launch_bounds__(256)
bn_drop_apply_kernel(const float* __restrict__ A, int K, const BnDrop pro, float* __restrict__ out,
                     int B) {
  const unsigned step = pro.state != nullptr ? adam_step_of(pro.state) : 0u;
  const long long n = static_cast<long long>(B) * K;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < n; e += gridDim.x * 256LL) {
    const int r = static_cast<int>(e / K), k = static_cast<int>(e % K);
    float mu, rstd;
    bn_consts(pro.sums, pro.mean, pro.var, k, K, pro.inv_B, pro.eps, &mu, &rstd);
    float v = fmaf((A[e] - mu) * rstd, pro.gamma[k], pro.beta[k]);
    if (pro.p > 0.f) v *= drop_scale(pro.seed, pro.layer, step, r, k, pro.p, pro.inv_keep);
    out[e] = v;
  }
}

// Backward of bn_drop_apply (the last BN + dropout of a tower that ends without a dense layer,
// dcn/dcn.py:146-149): dn[r,k] = dout[r,k] * keep(r,k), dbeta[k] += sum_r dn, dgamma[k] += sum_r
// dn * xhat(A[r,k]) - the same bookkeeping tower_layer_bwd_data leaves for the layer below, so
// the layer's dpre follows from a kind-1 gradient source over dn.  CTA = 32 rows, thread = column.
__global__ void __launch_bounds__(256)
bn_drop_apply_bwd_kernel(const float* __restrict__ dout, int ldd, const float* __restrict__ A, int K,
                         const BnDrop pro, float* __restrict__ dn, float* __restrict__ dbeta,
                         float* __restrict__ dgamma, int B) {
  const unsigned step = pro.state != nullptr ? adam_step_of(pro.state) : 0u;
  const int r0 = blockIdx.x * 32, r1 = min(B, r0 + 32);
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float mu, rstd;
    bn_consts(pro.sums, pro.mean, pro.var, k, K, pro.inv_B, pro.eps, &mu, &rstd);
    float cb = 0.f, cg = 0.f;
    for (int r = r0; r < r1; ++r) {
      float v = dout[static_cast<size_t>(r) * ldd + k];
      if (pro.p > 0.f) v *= drop_scale(pro.seed, pro.layer, step, r, k, pro.p, pro.inv_keep);
      dn[static_cast<size_t>(r) * K + k] = v;
      cb += v;
      cg = fmaf(v, (A[static_cast<size_t>(r) * K + k] - mu) * rstd, cg);
    }
    if (dbeta != nullptr && cb != 0.f) red_add_f32(dbeta + k, cb);
    if (dgamma != nullptr && cg != 0.f) red_add_f32(dgamma + k, cg);
  }
}

// ------------------------------------------------------------------------- DCN head
// dcn/dcn.py:151-153 + the loss block :166-169 in one launch: logit[b] = [h[b,:H] | xl[b,:W]] . w
// + hb, prob, mean BCE, and (training) dh = dl * w[:H], dxl = dl * w[H:], dw += sum_b dl * [h|xl],
// dhb += sum_b dl with dl = (prob - z) * grad_scale.  Warp per sample; lane l owns the float4
// columns l, l+32, ... of the concatenated row in every sample its warp visits, so the weight
// gradient accumulates in registers and is flushed once per warp.
constexpr int kDcnHeadNIT = 12;      // (H + W) / 4 <= 32 * 12
struct DcnHeadParams {
  const float* h; const float* xl; const float* w; const float* hb; const float* labels;
  float* logits; float* prob; float* loss; float* dh; float* dxl; float* dw; float* dhb;
  float loss_scale, grad_scale;
  int H, W, B, want_grad;
};
__global__ void __launch_bounds__(256) dcn_head_kernel(const DcnHeadParams p) {
  __shared__ float s_red[8][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int HQ = p.H >> 2, NQ = (p.H + p.W) >> 2;
  float4 wv[kDcnHeadNIT], dwv[kDcnHeadNIT];
#pragma unroll
  for (int it = 0; it < kDcnHeadNIT; ++it) {
    const int q = it * 32 + lane;
    wv[it] = q < NQ ? ldg4(p.w + q * 4) : f4_zero();
    dwv[it] = f4_zero();
  }
  const float hb = p.hb[0];
  float a_loss = 0.f, a_hb = 0.f;
  for (int b = blockIdx.x * 8 + warp; b < p.B; b += gridDim.x * 8) {
    float4 x[kDcnHeadNIT];
    float dot = 0.f;
#pragma unroll
    for (int it = 0; it < kDcnHeadNIT; ++it) {
      const int q = it * 32 + lane;
      x[it] = f4_zero();
      if (q < HQ) x[it] = ldg4(p.h + static_cast<size_t>(b) * p.H + q * 4);
      else if (q < NQ) x[it] = ldg4(p.xl + static_cast<size_t>(b) * p.W + (q - HQ) * 4);
      dot += f4_dot(x[it], wv[it]);
    }
    const float logit = warp_sum(dot) + hb;
    const float z = p.labels[b];
    const float pr = 1.f / (1.f + expf(-logit));
    if (lane == 0) {
      if (p.logits != nullptr) p.logits[b] = logit;
      if (p.prob != nullptr) p.prob[b] = pr;
      a_loss += fmaxf(logit, 0.f) - logit * z + log1pf(expf(-fabsf(logit)));
    }
    if (p.want_grad) {
      const float dl = (pr - z) * p.grad_scale;
      if (lane == 0) a_hb += dl;
#pragma unroll
      for (int it = 0; it < kDcnHeadNIT; ++it) {
        const int q = it * 32 + lane;
        if (q >= NQ) continue;
        dwv[it] = f4_fma(dl, x[it], dwv[it]);
        const float4 g = make_float4(dl * wv[it].x, dl * wv[it].y, dl * wv[it].z, dl * wv[it].w);
        if (q < HQ) *reinterpret_cast<float4*>(p.dh + static_cast<size_t>(b) * p.H + q * 4) = g;
        else *reinterpret_cast<float4*>(p.dxl + static_cast<size_t>(b) * p.W + (q - HQ) * 4) = g;
      }
    }
  }
  if (p.want_grad) {
#pragma unroll
    for (int it = 0; it < kDcnHeadNIT; ++it) {
      const int q = it * 32 + lane;
      if (q < NQ) red_add_v4(p.dw + q * 4, dwv[it]);
    }
  }
  if (lane == 0) {
    s_red[warp][0] = a_loss;
    s_red[warp][1] = a_hb;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_red[w][threadIdx.x];
    if (threadIdx.x == 0) red_add_f32(p.loss, t * p.loss_scale);
    else if (p.want_grad) red_add_f32(p.dhb, t);
  }
}

// -------------------------------------------------------------------------- backward
// dXin[r,k] = sum_n dpre[r,n] * W[k,n]; then through the dropout + BN that produced Xin from the
// stored activation Aprev (pro): dn = dXin * keep is stored, and its column sums
// dbeta_prev += sum_r dn, dgamma_prev += sum_r dn * xhat_prev are accumulated for the layer below.
// With pro.enabled == 0 (first layer) dXin is stored as is.
template <int RT>
__global__ void __launch_bounds__(256)
tower_layer_bwd_data_kernel(const GradSrc gs, int N, const float* __restrict__ W, int K,
                            const BnDrop pro, const float* __restrict__ Aprev,
                            float* __restrict__ dn_out, int ldn, float* __restrict__ dbeta_prev,
                            float* __restrict__ dgamma_prev, int B) {
  extern __shared__ __align__(16) uint8_t tw_smem[];
  TowerSmem& sm = *reinterpret_cast<TowerSmem*>(tw_smem);
  constexpr int MT = RT / 2;
  constexpr int BM = MT * 16;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int r0 = blockIdx.x * BM, k0 = blockIdx.y * kTwBN;
  const unsigned step = pro.state != nullptr ? adam_step_of(pro.state) : 0u;
  fill_gs_tables(sm, gs, N);
  if (pro.enabled) fill_pro_tables(sm, pro, K);
  __syncthreads();
  float acc[MT][2][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;
  const bool wal = (N & 3) == 0 && aligned16_dev(W);
  auto fa4 = [&](int rr, int n) -> float4 {
    const int r = r0 + rr;
    if (r >= B || n >= N) return f4_zero();
    return dpre4(sm, gs, r, n, N);
  };
  auto fb4 = [&](int n, int c) -> float4 {          // 4 consecutive n of W row (k0 + c)
    const int k = k0 + c;
    if (k >= K || n >= N) return f4_zero();
    return load4_guard(W + static_cast<size_t>(k) * N + n, N - n, wal);
  };
  gemm_tile_mma<MT, true, true>(sm, N, fa4, fb4, acc);

#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    float cb[2] = {0.f, 0.f}, cg[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int k = k0 + warp * 16 + nt * 8 + 2 * t + j;
      if (k >= K) continue;
      float mu = 0.f, rstd = 0.f;
      if (pro.enabled) bn_consts(pro.sums, pro.mean, pro.var, k, K, pro.inv_B, pro.eps, &mu, &rstd);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int r = r0 + mt * 16 + g + h * 8;
          if (r >= B) continue;
          float v = acc[mt][nt][h * 2 + j];
          if (pro.enabled) {
            if (pro.p > 0.f) v *= drop_scale(pro.seed, pro.layer, step, r, k, pro.p, pro.inv_keep);
            const float xhat = (Aprev[static_cast<size_t>(r) * K + k] - mu) * rstd;
            cb[j] += v;
            cg[j] = fmaf(v, xhat, cg[j]);
          }
          dn_out[static_cast<size_t>(r) * ldn + k] = v;
        }
    }
    if (pro.enabled && dbeta_prev != nullptr) {
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        cb[0] += __shfl_xor_sync(0xffffffffu, cb[0], o);
        cb[1] += __shfl_xor_sync(0xffffffffu, cb[1], o);
        cg[0] += __shfl_xor_sync(0xffffffffu, cg[0], o);
        cg[1] += __shfl_xor_sync(0xffffffffu, cg[1], o);
      }
      if (g == 0) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int k = k0 + warp * 16 + nt * 8 + 2 * t + j;
          if (k < K) {
            red_add_f32(dbeta_prev + k, cb[j]);
            red_add_f32(dgamma_prev + k, cg[j]);
          }
        }
      }
    }
  }
}

// dpre[r,n] = g * 1[a > 0] written out once for both backward GEMMs of a layer (GradSrc kind 2),
// db[n] += column sums.  Thread = column, CTA = kDpreRows rows.
constexpr int kDpreRows = 32;
__global__ void __launch_bounds__(256)
tower_dpre_kernel(const GradSrc gs, int N, float* __restrict__ dpre, int ldd,
                  float* __restrict__ db, int B) {
  const int r0 = blockIdx.x * kDpreRows, r1 = min(B, r0 + kDpreRows);
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float mu = 0.f, rstd = 0.f, c1 = 0.f, c2 = 0.f, gsc = 1.f;
    if (gs.kind == 1) {
      bn_consts(gs.sums, gs.mean, gs.var, n, N, gs.inv_B, gs.eps, &mu, &rstd);
      c1 = gs.train ? gs.dbeta[n] * gs.inv_B : 0.f;
      c2 = gs.train ? gs.dgamma[n] * gs.inv_B : 0.f;
      gsc = rstd * gs.gamma[n];
    }
    float acc = 0.f;
#pragma unroll 4
    for (int r = r0; r < r1; ++r) {
      const float a = gs.a[static_cast<size_t>(r) * gs.lda + n];
      float v = gs.G[static_cast<size_t>(r) * gs.ldg + n];
      if (gs.kind == 1) v = (v - c1 - (a - mu) * rstd * c2) * gsc;
      v = a > 0.f ? v : 0.f;
      dpre[static_cast<size_t>(r) * ldd + n] = v;
      acc += v;
    }
    if (db != nullptr && acc != 0.f) red_add_f32(db + n, acc);
  }
}

// Same, 4 columns per thread with 16-byte accesses (N % 4 == 0, pitches % 4 == 0, aligned):
// 8 row groups x 32 column quads per CTA, 4 rows per thread all in flight at once.
__global__ void __launch_bounds__(256)
tower_dpre_v4_kernel(const GradSrc gs, int N, float* __restrict__ dpre, int ldd,
                     float* __restrict__ db, int B) {
  __shared__ float4 s_sum[8][64];
  const int cq = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int r0 = blockIdx.x * kDpreRows;
  const int nq = N >> 2;
  for (int q = cq; q < nq; q += 32) {
    const int n = q * 4;
    float mu[4] = {0.f, 0.f, 0.f, 0.f}, rstd[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f},
          c2[4] = {0.f, 0.f, 0.f, 0.f}, gsc[4] = {1.f, 1.f, 1.f, 1.f};
    if (gs.kind == 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        bn_consts(gs.sums, gs.mean, gs.var, n + j, N, gs.inv_B, gs.eps, &mu[j], &rstd[j]);
        c1[j] = gs.train ? gs.dbeta[n + j] * gs.inv_B : 0.f;
        c2[j] = gs.train ? gs.dgamma[n + j] * gs.inv_B : 0.f;
        gsc[j] = rstd[j] * gs.gamma[n + j];
      }
    }
    float4 av[4], gv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + rg + 8 * i;
      if (r < B) {
        av[i] = ldg4(gs.a + static_cast<size_t>(r) * gs.lda + n);
        gv[i] = ldg4(gs.G + static_cast<size_t>(r) * gs.ldg + n);
      } else {
        av[i] = f4_zero();
        gv[i] = f4_zero();
      }
    }
    float4 acc = f4_zero();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + rg + 8 * i;
      float4 v = gv[i];
      const float4 a = av[i];
      if (gs.kind == 1) {
        v.x = (v.x - c1[0] - (a.x - mu[0]) * rstd[0] * c2[0]) * gsc[0];
        v.y = (v.y - c1[1] - (a.y - mu[1]) * rstd[1] * c2[1]) * gsc[1];
        v.z = (v.z - c1[2] - (a.z - mu[2]) * rstd[2] * c2[2]) * gsc[2];
        v.w = (v.w - c1[3] - (a.w - mu[3]) * rstd[3] * c2[3]) * gsc[3];
      }
      v.x = a.x > 0.f ? v.x : 0.f;
      v.y = a.y > 0.f ? v.y : 0.f;
      v.z = a.z > 0.f ? v.z : 0.f;
      v.w = a.w > 0.f ? v.w : 0.f;
      if (r < B) *reinterpret_cast<float4*>(dpre + static_cast<size_t>(r) * ldd + n) = v;
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (db != nullptr) {
      s_sum[rg][q] = acc;
    }
  }
  if (db != nullptr) {
    __syncthreads();
    for (int n = threadIdx.x; n < N; n += 256) {
      float t = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) t += reinterpret_cast<const float*>(&s_sum[g][0])[n];
      if (t != 0.f) red_add_f32(db + n, t);
    }
  }
}

// db[n] += sum_r D[r*ldd + n]
__global__ void __launch_bounds__(256)
tower_colsum_kernel(const float* __restrict__ D, int ldd, int N, float* __restrict__ db, int B) {
  const int r0 = blockIdx.x * kDpreRows, r1 = min(B, r0 + kDpreRows);
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float acc = 0.f;
    for (int r = r0; r < r1; ++r) acc += D[static_cast<size_t>(r) * ldd + n];
    if (acc != 0.f) red_add_f32(db + n, acc);
  }
}

// N == 1 (the final dense(1, relu)), both halves of its backward in one pass over the rows:
//   dpre[r] = G[r] * 1[y[r] > 0]
//   dW[k] += sum_r P(A)[r,k] * dpre[r];  db += sum_r dpre[r]                   (weights == true)
//   dn[r,k] = dpre[r] * W[k] * keep;  dbeta[k] += sum_r dn;  dgamma[k] += sum_r dn * xhat  (data)
// CTA = 16 rows x K columns (thread = column), grid = B/16: enough CTAs to hide the latency.
constexpr int kOutRows = 16;
template <bool WEIGHTS>
__global__ void __launch_bounds__(128)
tower_out_bwd_kernel(const GradSrc gs, const float* __restrict__ X, int ldx,
                     const float* __restrict__ W, int K, const BnDrop pro, float* __restrict__ dW,
                     float* __restrict__ db, float* __restrict__ dn_out, int ldn,
                     float* __restrict__ dbeta_prev, float* __restrict__ dgamma_prev, int B) {
  __shared__ float s_dp[kOutRows];
  const int r0 = blockIdx.x * kOutRows;
  const unsigned step = pro.state != nullptr ? adam_step_of(pro.state) : 0u;
  if (threadIdx.x < kOutRows) {
    const int r = r0 + threadIdx.x;
    float v = 0.f;
    if (r < B) {
      const float a = gs.a[static_cast<size_t>(r) * gs.lda];
      v = a > 0.f ? gs.G[static_cast<size_t>(r) * gs.ldg] : 0.f;
    }
    s_dp[threadIdx.x] = v;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float mu = 0.f, rstd = 0.f, sc = 1.f, sh = 0.f;
    if (pro.enabled) {
      bn_consts(pro.sums, pro.mean, pro.var, k, K, pro.inv_B, pro.eps, &mu, &rstd);
      sc = rstd * pro.gamma[k];
      sh = pro.beta[k];
    }
    const float w = WEIGHTS ? 0.f : W[k];
    const bool need_x = WEIGHTS || pro.enabled;
    float xv[kOutRows];
#pragma unroll
    for (int rr = 0; rr < kOutRows; ++rr)
      xv[rr] = (need_x && r0 + rr < B) ? X[static_cast<size_t>(r0 + rr) * ldx + k] : 0.f;
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int rr = 0; rr < kOutRows; ++rr) {
      const int r = r0 + rr;
      const float d = s_dp[rr];
      if (r >= B || d == 0.f) {
        if (!WEIGHTS && r < B) dn_out[static_cast<size_t>(r) * ldn + k] = 0.f;
        continue;
      }
      const float keep = (pro.enabled && pro.p > 0.f)
                             ? drop_scale(pro.seed, pro.layer, step, r, k, pro.p, pro.inv_keep) : 1.f;
      if (WEIGHTS) {
        const float v = pro.enabled ? fmaf(xv[rr] - mu, sc, sh) * keep : xv[rr];
        a0 = fmaf(v, d, a0);
      } else {
        const float v = d * w * keep;
        dn_out[static_cast<size_t>(r) * ldn + k] = v;
        if (pro.enabled) {
          a0 += v;
          a1 = fmaf(v, (xv[rr] - mu) * rstd, a1);
        }
      }
    }
    if (WEIGHTS) {
      if (a0 != 0.f) red_add_f32(dW + k, a0);
    } else if (pro.enabled && dbeta_prev != nullptr) {
      if (a0 != 0.f) red_add_f32(dbeta_prev + k, a0);
      if (a1 != 0.f) red_add_f32(dgamma_prev + k, a1);
    }
  }
  if (WEIGHTS && db != nullptr && threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int rr = 0; rr < kOutRows; ++rr) t += s_dp[rr];
    if (t != 0.f) red_add_f32(db, t);
  }
}

// dW[k,n] += sum_r P(X)[r,k] * dpre[r,n];  db[n] += sum_r dpre[r,n]  (rows split over gridDim.z)
// Output tile 32 (k) x 128 (n); the reduction runs over the CTA's row range.
__global__ void __launch_bounds__(256)
tower_layer_bwd_weights_kernel(const float* __restrict__ X, int ldx, int K, const BnDrop pro,
                               const GradSrc gs, int N, float* __restrict__ dW,
                               float* __restrict__ db, int B, int rows_per_split) {
  extern __shared__ __align__(16) uint8_t tw_smem[];
  TowerSmem& sm = *reinterpret_cast<TowerSmem*>(tw_smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int k0 = blockIdx.x * kTwBM, n0 = blockIdx.y * kTwBN;
  const int rbeg = blockIdx.z * rows_per_split, rend = min(B, rbeg + rows_per_split);
  const unsigned step = pro.state != nullptr ? adam_step_of(pro.state) : 0u;
  fill_gs_tables(sm, gs, N);
  if (pro.enabled) fill_pro_tables(sm, pro, K);
  __syncthreads();
  float acc[2][2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;
  const bool xal = (ldx & 3) == 0 && aligned16_dev(X);
  const int nrows = rend - rbeg;
  auto fa4 = [&](int m, int rr) -> float4 {      // 4 consecutive k of row (rbeg + rr)
    const int k = k0 + m, r = rbeg + rr;
    if (k >= K || rr >= nrows) return f4_zero();
    float4 v = load4_guard(X + static_cast<size_t>(r) * ldx + k, K - k, xal);
    if (pro.enabled) {
      v.x = pro_value(sm, pro, step, v.x, r, k);
      if (k + 1 < K) v.y = pro_value(sm, pro, step, v.y, r, k + 1);
      if (k + 2 < K) v.z = pro_value(sm, pro, step, v.z, r, k + 2);
      if (k + 3 < K) v.w = pro_value(sm, pro, step, v.w, r, k + 3);
    }
    return v;
  };
  auto fb4 = [&](int rr, int c) -> float4 {
    const int n = n0 + c;
    if (rr >= nrows || n >= N) return f4_zero();
    return dpre4(sm, gs, rbeg + rr, n, N);
  };
  gemm_tile_mma<2, false, false>(sm, nrows, fa4, fb4, acc);
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        const int k = k0 + mt * 16 + g + (ci >> 1) * 8;
        const int n = n0 + warp * 16 + nt * 8 + 2 * t + (ci & 1);
        if (k < K && n < N) red_add_f32(dW + static_cast<size_t>(k) * N + n, acc[mt][nt][ci]);
      }
  if (db != nullptr) {   // column sums of dpre: this split's rows are shared out over blockIdx.x
    const int len = rend - rbeg, gx = gridDim.x;
    const int per = (len + gx - 1) / gx;
    const int rb = rbeg + blockIdx.x * per, re = min(rend, rb + per);
    const int c = tid & 127, half = tid >> 7;
    const int n = n0 + c;
    float sacc = 0.f;
    if (n < N)
      for (int r = rb + half; r < re; r += 2) sacc += dpre_value(sm, gs, r, n);
    sm.red[0][half][c] = sacc;
    __syncthreads();
    if (half == 0 && n < N) {
      const float t = sm.red[0][0][c] + sm.red[0][1][c];
      if (t != 0.f) red_add_f32(db + n, t);
    }
  }
}

// ------------------------------------------------------------------------- loss head
// logit[b] = sum_{c<C} hw[c] * act_c(z_c[b]) + hb, act_0 = relu(. + b1) when relu0, identity else.
// loss += mean BCE; grads: dz_c[b], dhw[c], dhb, db1 accumulated (RED).  scale = 1/(B*world).
struct HeadParams {
  const float* z[4];
  float* dz[4];
  const float* hw;
  const float* hb;
  const float* b1;
  const float* labels;
  float* logits;
  float* prob;
  float* loss;       // scalar, += sum_b bce_b * loss_scale
  float* dhw;
  float* dhb;
  float* db1;
  float loss_scale;  // 1/B
  float grad_scale;  // 1/(B * world)
  int C, relu0, B, want_grad;
};

__global__ void __launch_bounds__(256) loss_head_kernel(const HeadParams p) {
  __shared__ float s_part[8][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float hw[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c = 0; c < p.C; ++c) hw[c] = p.hw[c];
  const float hb = p.hb[0];
  const float b1 = p.relu0 ? p.b1[0] : 0.f;
  float a_loss = 0.f, a_hb = 0.f, a_b1 = 0.f, a_hw[4] = {0.f, 0.f, 0.f, 0.f};
  for (int b = blockIdx.x * 256 + threadIdx.x; b < p.B; b += gridDim.x * 256) {
    float act[4];
    float logit = hb;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      act[c] = 0.f;
      if (c < p.C) {
        float v = p.z[c][b];
        if (c == 0 && p.relu0) v = fmaxf(v + b1, 0.f);
        act[c] = v;
        logit = fmaf(hw[c], v, logit);
      }
    }
    const float z = p.labels[b];
    const float pr = 1.f / (1.f + expf(-logit));
    // tf.nn.sigmoid_cross_entropy_with_logits: max(x,0) - x z + log1p(exp(-|x|))
    const float bce = fmaxf(logit, 0.f) - logit * z + log1pf(expf(-fabsf(logit)));
    if (p.logits != nullptr) p.logits[b] = logit;
    if (p.prob != nullptr) p.prob[b] = pr;
    a_loss += bce;
    if (p.want_grad) {
      const float dl = (pr - z) * p.grad_scale;
      a_hb += dl;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < p.C) {
          a_hw[c] = fmaf(dl, act[c], a_hw[c]);
          float g = dl * hw[c];
          if (c == 0 && p.relu0) {
            g = act[0] > 0.f ? g : 0.f;
            a_b1 += g;
          }
          p.dz[c][b] = g;
        }
      }
    }
  }
  float vals[8] = {a_loss, a_hb, a_b1, a_hw[0], a_hw[1], a_hw[2], a_hw[3], 0.f};
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const float t = warp_sum(vals[i]);
    if (lane == 0) s_part[warp][i] = t;
  }
  __syncthreads();
  if (threadIdx.x < 7) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_part[w][threadIdx.x];
    const int i = threadIdx.x;
    if (i == 0) red_add_f32(p.loss, t * p.loss_scale);
    else if (p.want_grad) {
      if (i == 1) red_add_f32(p.dhb, t);
      else if (i == 2) { if (p.relu0) red_add_f32(p.db1, t); }
      else if (i - 3 < p.C) red_add_f32(p.dhw + (i - 3), t);
    }
  }
}

// lo[i] = tcg_lo(x[i])
__global__ void __launch_bounds__(256) split_lo_kernel(const float* __restrict__ x,
                                                       float* __restrict__ lo, long long n) {
  const long long n4 = n >> 2, stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long i0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  for (long long i = i0; i < n4; i += stride)
    reinterpret_cast<float4*>(lo)[i] = tcg_lo4(__ldg(reinterpret_cast<const float4*>(x) + i));
  for (long long i = (n4 << 2) + i0; i < n; i += stride) lo[i] = tcg_lo(x[i]);
}

}  // namespace ctr

using namespace ctr;

extern "C" {

static void tower_smem_optin() {
  static bool done = false;
  if (done) return;
  const int bytes = static_cast<int>(sizeof(TowerSmem));
  cudaFuncSetAttribute(tower_layer_fwd_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  cudaFuncSetAttribute(tower_layer_fwd_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  cudaFuncSetAttribute(tower_layer_fwd_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  cudaFuncSetAttribute(tower_layer_fwd_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  cudaFuncSetAttribute(tower_layer_bwd_data_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  cudaFuncSetAttribute(tower_layer_bwd_data_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  cudaFuncSetAttribute(tower_layer_bwd_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  done = true;
}

// tcgen05 path for wide layers without a BN prologue (CTR_TOWER_TC=0 forces the mma.sync path)
static bool tower_tc() {
  const char* e = getenv("CTR_TOWER_TC");
  return e == nullptr || atoi(e) != 0;
}
static int round16(int x) { return (x + 15) / 16 * 16; }

static BnDrop make_pro(const ctr_bn_drop* d, int B) {
  BnDrop p{};
  if (d == nullptr || !d->enabled) return p;
  p.sums = d->sums; p.mean = d->mean; p.var = d->var; p.gamma = d->gamma; p.beta = d->beta;
  p.state = d->state; p.inv_B = 1.f / static_cast<float>(B); p.eps = d->eps; p.p = d->p_drop;
  p.inv_keep = d->p_drop > 0.f ? 1.f / (1.f - d->p_drop) : 1.f;
  p.seed = d->seed; p.layer = d->layer; p.enabled = 1;
  return p;
}
static GradSrc make_gs(const ctr_grad_src* s, int B) {
  GradSrc g{};
  g.G = s->G; g.ldg = s->ldg; g.a = s->a; g.lda = s->lda; g.sums = s->sums; g.mean = s->mean;
  g.var = s->var; g.gamma = s->gamma; g.dbeta = s->dbeta; g.dgamma = s->dgamma;
  g.inv_B = 1.f / static_cast<float>(B); g.eps = s->eps; g.kind = s->kind; g.train = s->train;
  return g;
}

int ctr_tower_layer_fwd(const float* X, int ldx, int K, const ctr_bn_drop* pro, const float* W,
                        const float* bias, int N, float* out, int ldo, float* stats, int relu,
                        int B, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(X && W && out && B >= 0 && K > 0 && N > 0 && ldx >= K && ldo >= N,
              "ctr_tower_layer_fwd", "bad argument");
  const bool has_pro = pro != nullptr && pro->enabled;
  CTR_REQUIRE(!has_pro || K <= kMaxBn, "ctr_tower_layer_fwd", "BN prologue needs K <= 256");
  CTR_REQUIRE(!has_pro || (pro->gamma && pro->beta && (pro->sums || (pro->mean && pro->var))),
              "ctr_tower_layer_fwd", "BN prologue needs gamma/beta and stats");
  if (B == 0) return CTR_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const BnDrop p = make_pro(pro, B);
  if (N == 1 && stats == nullptr && (!has_pro || K <= kMaxBn)) {
    const int g1 = std::min((B + 7) / 8, sm_count() * 8);
    tower_out_fwd_kernel<<<g1, 256, 0, st>>>(X, ldx, K, p, W, bias, out, ldo, relu, B);
    CTR_LAUNCH_CHECK("ctr_tower_layer_fwd");
  }
  if (!has_pro && tower_tc() && B >= 256 && K >= 64 && N >= 16 && (K & 3) == 0 && (N & 3) == 0 &&
      (ldx & 3) == 0 && aligned16(X) && aligned16(W)) {
    const int mtiles = (B + kTcBM - 1) / kTcBM;
    // few row tiles: split N down to 32 columns per CTA to put every SM to work (the A tiles
    // are then re-read from L2, which is cheaper than idle SMs)
    int ntiles = std::max((N + 127) / 128, std::min(sm_count() / mtiles, (N + 31) / 32));
    if (const char* e = ctr_knob("CTR_TCG_NTILES")) ntiles = std::max(1, atoi(e));
    const int NT = std::min(128, round16((N + ntiles - 1) / ntiles));
    return tc_gemm_launch<TCG_EPI_FWD>(X, ldx, false, W, N, true, B, N, K, NT, 1, out, ldo, bias,
                                       stats, relu, st, "ctr_tower_layer_fwd");
  }
  tower_smem_optin();
  const int ny = (N + kTwBN - 1) / kTwBN;
  bool small = static_cast<long long>((B + 31) / 32) * ny * 2 < sm_count();   // 16-row tiles
  if (const char* e = ctr_knob("CTR_TOWER_RT")) small = atoi(e) == 2;
  dim3 grid(small ? (B + 15) / 16 : (B + 31) / 32, ny);
  const size_t sb = sizeof(TowerSmem);
  if (has_pro) {
    if (small) tower_layer_fwd_kernel<true, 2><<<grid, 256, sb, st>>>(X, ldx, K, p, W, bias, N, out, ldo, stats, relu, B);
    else tower_layer_fwd_kernel<true, 4><<<grid, 256, sb, st>>>(X, ldx, K, p, W, bias, N, out, ldo, stats, relu, B);
  } else {
    if (small) tower_layer_fwd_kernel<false, 2><<<grid, 256, sb, st>>>(X, ldx, K, p, W, bias, N, out, ldo, stats, relu, B);
    else tower_layer_fwd_kernel<false, 4><<<grid, 256, sb, st>>>(X, ldx, K, p, W, bias, N, out, ldo, stats, relu, B);
  }
  CTR_LAUNCH_CHECK("ctr_tower_layer_fwd");
}

int ctr_split_lo(const float* x, float* lo, int64_t n, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(x && lo && n >= 0 && aligned16(x) && aligned16(lo), "ctr_split_lo",
              "null / unaligned pointer");
  if (n == 0) return CTR_OK;
  const int grid = static_cast<int>(std::min<long long>((n / 4 + 255) / 256 + 1, sm_count() * 4LL));
  split_lo_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, lo, n);
  CTR_LAUNCH_CHECK("ctr_split_lo");
}

int ctr_tower_gemm_presplit(int kind, const float* A, const float* A_lo, const float* Bm,
                            const float* B_lo, int B, int K, int N, float* out, const float* bias,
                            float* stats, int relu, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(kind >= 0 && kind <= 3 && A && A_lo && Bm && B_lo && out, "ctr_tower_gemm_presplit",
              "bad kind / null pointer");
  CTR_REQUIRE(B >= 256 && K >= 32 && N >= 16 && (K & 3) == 0 && (N & 3) == 0, "ctr_tower_gemm_presplit",
              "needs B >= 256, K >= 32, N >= 16, K % 4 == N % 4 == 0");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const char* fn = "ctr_tower_gemm_presplit";
  if (kind == 0) {          // out[B,N] = act(X[B,K] . W[K,N] + bias): W is MN-major (n contiguous)
    const int mtiles = (B + kTcBM - 1) / kTcBM;
    int ntiles = std::max((N + 127) / 128, std::min(sm_count() / mtiles, (N + 31) / 32));
    if (const char* e = ctr_knob("CTR_TCG_NTILES")) ntiles = std::max(1, atoi(e));
    const int NT = std::min(128, round16((N + ntiles - 1) / ntiles));
    return tc_gemm_launch<TCG_EPI_FWD>(A, K, false, Bm, N, true, B, N, K, NT, 1, out, N, bias, stats,
                                       relu, st, fn, A_lo, B_lo);
  }
  if (kind == 3) {
    // out[B,N] += X . W, reduction split over the CTAs.  A tf32 MMA costs ~160 cycles of operand
    // fetch whatever its N, so the cheapest schedule has the fewest MMAs per CTA: one N tile and
    // the k-blocks shared out over sm_count / mtiles CTAs, partial sums RED'ed into `out`.
    CTR_REQUIRE(N <= 256 && aligned16(out), fn, "split-K forward needs N <= 256 and an aligned output");
    const int mtiles = (B + kTcBM - 1) / kTcBM;
    int splits = std::max(1, sm_count() / mtiles);
    if (const char* e = ctr_knob("CTR_TCG_FWD_SPLITS")) splits = std::max(1, atoi(e));
    return tc_gemm_launch<TCG_EPI_RED>(A, K, false, Bm, N, true, B, N, K, round16(N), splits, out, N,
                                       nullptr, nullptr, 0, st, fn, A_lo, B_lo);
  }
  if (kind == 1) {          // out[B,K] = dpre[B,N] . W[K,N]^T: both K-major in n
    const int mtiles = (B + kTcBM - 1) / kTcBM;
    int ntiles = std::max((K + 255) / 256, std::min(sm_count() / mtiles, (K + 63) / 64));
    if (const char* e = ctr_knob("CTR_TCG_NTILES")) ntiles = std::max((K + 255) / 256, atoi(e));
    const int NT = round16((K + ntiles - 1) / ntiles);
    return tc_gemm_launch<TCG_EPI_STORE>(A, N, false, Bm, N, false, B, K, N, NT, 1, out, K, nullptr,
                                         nullptr, 0, st, fn, A_lo, B_lo);
  }
  // out[K,N] += X[B,K]^T . dpre[B,N]: both MN-major, reduction over the rows, split-K with REDs
  CTR_REQUIRE(N <= 256 && aligned16(out), fn, "weights GEMM needs N <= 256 and an aligned output");
  const int mtiles = (K + kTcBM - 1) / kTcBM;
  int splits = std::max(1, std::min(sm_count() / mtiles, (B / kTcKB) / 4));
  if (const int o = option_get("tcg_dw_splits", 0)) splits = std::max(1, o);
  if (const char* e = ctr_knob("CTR_TCG_SPLITS")) splits = std::max(1, atoi(e));
  // runs on a side stream beside the embedding scatter: a 2-stage ring (128 KB) leaves the SM's
  // shared memory to the scatter kernel's CTAs ("tcg_dw_stages")
  return tc_gemm_launch<TCG_EPI_RED>(A, K, true, Bm, N, true, K, N, B, round16(N), splits, out, N,
                                     nullptr, nullptr, 0, st, fn, A_lo, B_lo,
                                     std::max(2, option_get("tcg_dw_stages", 2)));
}

int ctr_bn_drop_apply(const float* A, int K, const ctr_bn_drop* pro, float* out, int B,
                      ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(A && out && pro && pro->enabled && K > 0 && B >= 0, "ctr_bn_drop_apply", "bad argument");
  if (B == 0) return CTR_OK;
  const long long n = static_cast<long long>(B) * K;
  const int grid = static_cast<int>(std::min<long long>((n + 255) / 256, sm_count() * 8LL));
  bn_drop_apply_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(A, K, make_pro(pro, B), out, B);
  CTR_LAUNCH_CHECK("ctr_bn_drop_apply");
}

int ctr_bn_drop_apply_bwd(const float* dout, int ldd, const float* A, int K, const ctr_bn_drop* pro,
                          float* dn, float* dbeta, float* dgamma, int B, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(dout && A && dn && pro && pro->enabled && K > 0 && B >= 0 && ldd >= K,
              "ctr_bn_drop_apply_bwd", "bad argument");
  CTR_REQUIRE(pro->gamma && pro->beta && (pro->sums || (pro->mean && pro->var)),
              "ctr_bn_drop_apply_bwd", "BN needs gamma/beta and stats");
  if (B == 0) return CTR_OK;
  const int threads = K <= 32 ? 32 : K <= 64 ? 64 : K <= 128 ? 128 : 256;
  bn_drop_apply_bwd_kernel<<<(B + 31) / 32, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      dout, ldd, A, K, make_pro(pro, B), dn, dbeta, dgamma, B);
  CTR_LAUNCH_CHECK("ctr_bn_drop_apply_bwd");
}

int ctr_dcn_head(const float* h, int H, const float* xl, int W, const float* w, const float* hb,
                 const float* labels, int B, float* logits, float* prob, float* loss, float* dh,
                 float* dxl, float* dw, float* dhb, float grad_scale, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(h && xl && w && hb && labels && loss && B >= 0, "ctr_dcn_head", "null pointer");
  CTR_REQUIRE(H > 0 && W > 0 && (H & 3) == 0 && (W & 3) == 0 && (H + W) / 4 <= 32 * kDcnHeadNIT,
              "ctr_dcn_head", "need H % 4 == W % 4 == 0 and H + W <= 1536");
  CTR_REQUIRE(aligned16(h) && aligned16(xl) && aligned16(w) && aligned16(dh) && aligned16(dxl) &&
                  aligned16(dw),
              "ctr_dcn_head", "pointers must be 16-byte aligned");
  const bool want_grad = dh != nullptr;
  CTR_REQUIRE(!want_grad || (dxl && dw && dhb), "ctr_dcn_head", "missing gradient outputs");
  if (B == 0) return CTR_OK;
  DcnHeadParams p{};
  p.h = h; p.xl = xl; p.w = w; p.hb = hb; p.labels = labels; p.logits = logits; p.prob = prob;
  p.loss = loss; p.dh = dh; p.dxl = dxl; p.dw = dw; p.dhb = dhb;
  p.loss_scale = 1.f / static_cast<float>(B); p.grad_scale = grad_scale;
  p.H = H; p.W = W; p.B = B; p.want_grad = want_grad ? 1 : 0;
  const int grid = std::min((B + 7) / 8, sm_count() * 2);
  dcn_head_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  CTR_LAUNCH_CHECK("ctr_dcn_head");
}

int ctr_tower_layer_bwd_data(const ctr_grad_src* gs, int N, const float* W, int K,
                             const ctr_bn_drop* pro, const float* Aprev, float* dn_out, int ldn,
                             float* dbeta_prev, float* dgamma_prev, int B, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(gs && gs->G && (gs->a || gs->kind == 2) && W && dn_out && K > 0 && N > 0 && B >= 0 &&
                  ldn >= K,
              "ctr_tower_layer_bwd_data", "bad argument");
  const bool has_pro = pro != nullptr && pro->enabled;
  CTR_REQUIRE(!has_pro || (Aprev && K <= kMaxBn), "ctr_tower_layer_bwd_data",
              "BN bookkeeping needs Aprev and K <= 256");
  CTR_REQUIRE(gs->kind == 0 || N <= kMaxBn, "ctr_tower_layer_bwd_data", "BN gradient source needs N <= 256");
  if (B == 0) return CTR_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (N == 1 && gs->kind == 0) {
    tower_out_bwd_kernel<false><<<(B + kOutRows - 1) / kOutRows, 128, 0, st>>>(
        make_gs(gs, B), Aprev != nullptr ? Aprev : W, K, W, K, make_pro(pro, B), nullptr, nullptr, dn_out,
        ldn, dbeta_prev, dgamma_prev, B);
    CTR_LAUNCH_CHECK("ctr_tower_layer_bwd_data");
  }
  if (gs->kind == 2 && !has_pro && tower_tc() && B >= 256 && K >= 64 && N >= 32 && (N & 3) == 0 &&
      (gs->ldg & 3) == 0 && aligned16(gs->G) && aligned16(W)) {
    // out[B, K] = dpre[B, N] . W[K, N]^T: both operands K-major in n
    const int mtiles = (B + kTcBM - 1) / kTcBM;
    int ntiles = std::max((K + 255) / 256, std::min(sm_count() / mtiles, (K + 63) / 64));
    if (const char* e = ctr_knob("CTR_TCG_NTILES")) ntiles = std::max((K + 255) / 256, atoi(e));
    const int NT = round16((K + ntiles - 1) / ntiles);
    return tc_gemm_launch<TCG_EPI_STORE>(gs->G, gs->ldg, false, W, N, false, B, K, N, NT, 1, dn_out,
                                         ldn, nullptr, nullptr, 0, st, "ctr_tower_layer_bwd_data");
  }
  tower_smem_optin();
  const int ny = (K + kTwBN - 1) / kTwBN;
  bool small = static_cast<long long>((B + 31) / 32) * ny * 2 < sm_count();
  if (const char* e = ctr_knob("CTR_TOWER_RT")) small = atoi(e) == 2;
  dim3 grid(small ? (B + 15) / 16 : (B + 31) / 32, ny);
  if (small)
    tower_layer_bwd_data_kernel<2><<<grid, 256, sizeof(TowerSmem), st>>>(
        make_gs(gs, B), N, W, K, make_pro(pro, B), Aprev, dn_out, ldn, dbeta_prev, dgamma_prev, B);
  else
    tower_layer_bwd_data_kernel<4><<<grid, 256, sizeof(TowerSmem), st>>>(
        make_gs(gs, B), N, W, K, make_pro(pro, B), Aprev, dn_out, ldn, dbeta_prev, dgamma_prev, B);
  CTR_LAUNCH_CHECK("ctr_tower_layer_bwd_data");
}

int ctr_tower_dpre(const ctr_grad_src* gs, int N, float* dpre, int ldd, float* db, int B,
                   ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(gs && gs->G && gs->a && dpre && N > 0 && B >= 0 && ldd >= N, "ctr_tower_dpre",
              "bad argument");
  CTR_REQUIRE(gs->kind == 0 || gs->kind == 1, "ctr_tower_dpre", "gradient source must be kind 0 or 1");
  CTR_REQUIRE(gs->kind == 0 || (gs->gamma && (gs->sums || (gs->mean && gs->var))), "ctr_tower_dpre",
              "BN gradient source needs gamma and stats");
  CTR_REQUIRE(gs->kind == 0 || !gs->train || (gs->dbeta && gs->dgamma), "ctr_tower_dpre",
              "BN gradient source needs dbeta / dgamma in train mode");
  if (B == 0) return CTR_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = (B + kDpreRows - 1) / kDpreRows;
  if ((N & 3) == 0 && N <= 256 && (ldd & 3) == 0 && (gs->ldg & 3) == 0 && (gs->lda & 3) == 0 &&
      aligned16(dpre) && aligned16(gs->G) && aligned16(gs->a)) {
    tower_dpre_v4_kernel<<<grid, 256, 0, st>>>(make_gs(gs, B), N, dpre, ldd, db, B);
  } else {
    const int threads = N <= 32 ? 32 : N <= 64 ? 64 : N <= 128 ? 128 : 256;
    tower_dpre_kernel<<<grid, threads, 0, st>>>(make_gs(gs, B), N, dpre, ldd, db, B);
  }
  CTR_LAUNCH_CHECK("ctr_tower_dpre");
}

int ctr_tower_layer_bwd_weights(const float* X, int ldx, int K, const ctr_bn_drop* pro,
                                const ctr_grad_src* gs, int N, float* dW, float* db, int B,
                                ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(X && gs && gs->G && (gs->a || gs->kind == 2) && dW && K > 0 && N > 0 && B >= 0 &&
                  ldx >= K,
              "ctr_tower_layer_bwd_weights", "bad argument");
  const bool has_pro = pro != nullptr && pro->enabled;
  CTR_REQUIRE(!has_pro || K <= kMaxBn, "ctr_tower_layer_bwd_weights", "BN prologue needs K <= 256");
  CTR_REQUIRE(gs->kind == 0 || N <= kMaxBn, "ctr_tower_layer_bwd_weights", "BN gradient source needs N <= 256");
  if (B == 0) return CTR_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (N == 1 && gs->kind == 0) {
    tower_out_bwd_kernel<true><<<(B + kOutRows - 1) / kOutRows, 128, 0, st>>>(
        make_gs(gs, B), X, ldx, nullptr, K, make_pro(pro, B), dW, db, nullptr, 0, nullptr, nullptr, B);
    CTR_LAUNCH_CHECK("ctr_tower_layer_bwd_weights");
  }
  if (gs->kind == 2 && !has_pro && tower_tc() && B >= 1024 && K >= 64 && N >= 16 && N <= 256 &&
      (K & 3) == 0 && (N & 3) == 0 && (ldx & 3) == 0 && (gs->ldg & 3) == 0 && aligned16(X) &&
      aligned16(gs->G) && aligned16(dW)) {
    // dW[K, N] += X[B, K]^T . dpre[B, N]: both operands MN-major, reduction over the rows
    if (db != nullptr)
      tower_colsum_kernel<<<(B + kDpreRows - 1) / kDpreRows, 256, 0, st>>>(gs->G, gs->ldg, N, db, B);
    const int mtiles = (K + kTcBM - 1) / kTcBM;
    int splits = std::max(1, std::min(sm_count() / mtiles, (B / kTcKB) / 4));
    if (const char* e = ctr_knob("CTR_TCG_SPLITS")) splits = std::max(1, atoi(e));
    return tc_gemm_launch<TCG_EPI_RED>(X, ldx, true, gs->G, gs->ldg, true, K, N, B, round16(N),
                                       splits, dW, N, nullptr, nullptr, 0, st,
                                       "ctr_tower_layer_bwd_weights");
  }
  tower_smem_optin();
  const int tiles = ((K + kTwBM - 1) / kTwBM) * ((N + kTwBN - 1) / kTwBN);
  int splits = std::max(1, std::min((sm_count() * 3) / tiles, (B + 63) / 64));
  // off the critical path (side stream): fewer, longer CTAs leave the SMs to the scatter / row
  // optimiser kernels it runs beside ("tower_dw_splits", 0 = automatic)
  if (const int o = option_get("tower_dw_splits", 0)) splits = std::max(1, std::min(splits, o));
  int rps = (B + splits - 1) / splits;
  rps = (rps + kTwKC - 1) / kTwKC * kTwKC;
  splits = (B + rps - 1) / rps;
  dim3 grid((K + kTwBM - 1) / kTwBM, (N + kTwBN - 1) / kTwBN, splits);
  tower_layer_bwd_weights_kernel<<<grid, 256, sizeof(TowerSmem), st>>>(
      X, ldx, K, make_pro(pro, B), make_gs(gs, B), N, dW, db, B, rps);
  CTR_LAUNCH_CHECK("ctr_tower_layer_bwd_weights");
}

int ctr_loss_head(const float* const* z, float* const* dz, int C, int relu0, const float* hw,
                  const float* hb, const float* b1, const float* labels, int B, float* logits,
                  float* prob, float* loss, float* dhw, float* dhb, float* db1, float grad_scale,
                  ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(z && hw && hb && labels && loss && C >= 1 && C <= 4 && B >= 0, "ctr_loss_head",
              "bad argument");
  CTR_REQUIRE(!relu0 || b1, "ctr_loss_head", "relu0 needs b1");
  const bool want_grad = dz != nullptr;
  CTR_REQUIRE(!want_grad || (dhw && dhb && (!relu0 || db1)), "ctr_loss_head", "missing gradient outputs");
  if (B == 0) return CTR_OK;
  HeadParams p{};
  for (int c = 0; c < C; ++c) {
    p.z[c] = z[c];
    p.dz[c] = want_grad ? dz[c] : nullptr;
    CTR_REQUIRE(p.z[c] && (!want_grad || p.dz[c]), "ctr_loss_head", "null column");
  }
  p.hw = hw; p.hb = hb; p.b1 = b1; p.labels = labels; p.logits = logits; p.prob = prob; p.loss = loss;
  p.dhw = dhw; p.dhb = dhb; p.db1 = db1; p.loss_scale = 1.f / static_cast<float>(B);
  p.grad_scale = grad_scale; p.C = C; p.relu0 = relu0; p.B = B; p.want_grad = want_grad ? 1 : 0;
  const int grid = std::min((B + 255) / 256, sm_count());
  loss_head_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  CTR_LAUNCH_CHECK("ctr_loss_head");
}

}  // extern "C"
