// xDeepFM CIN layer: C ABI + layout transposes.  prec 0 -> fp32 CUDA cores
// (cin_simt.cuh); prec 1/2 -> tcgen05 tensor cores (cin_tc.cuh).
#include "cin_simt.cuh"
#include "cin_tc.cuh"
#include "cin_dw_tc.cuh"
#include "cin_dw_fused.cuh"

namespace ctr {

// E[b, f, d] -> Xt[(b*D+d)*ld + f] (zero padded to ld): one warp per sample.
__global__ void __launch_bounds__(256)
transpose_fd_kernel(const float* __restrict__ E, int B, int F, int D, float* __restrict__ Xt,
                    int ld) {
  extern __shared__ float s[];  // [8 warps][F*(D+1)]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* t = s + warp * F * (D + 1);
  for (int b = blockIdx.x * 8 + warp; b < B; b += gridDim.x * 8) {
    for (int e = lane; e < F * D; e += 32) t[(e / D) * (D + 1) + (e % D)] = E[static_cast<size_t>(b) * F * D + e];
    __syncwarp();
    for (int e = lane; e < D * ld; e += 32) {
      const int d = e / ld, f = e % ld;
      Xt[(static_cast<size_t>(b) * D + d) * ld + f] = f < F ? t[f * (D + 1) + d] : 0.f;
    }
    __syncwarp();
  }
}

// dE[b, f, d] += dXt[(b*D+d)*ld + f]
__global__ void __launch_bounds__(256)
transpose_df_add_kernel(const float* __restrict__ dXt, int ld, int B, int F, int D,
                        float* __restrict__ dE) {
  extern __shared__ float s[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* t = s + warp * F * (D + 1);
  for (int b = blockIdx.x * 8 + warp; b < B; b += gridDim.x * 8) {
    for (int e = lane; e < D * ld; e += 32) {
      const int d = e / ld, f = e % ld;
      if (f < F) t[f * (D + 1) + d] = dXt[(static_cast<size_t>(b) * D + d) * ld + f];
    }
    __syncwarp();
    for (int e = lane; e < F * D; e += 32)
      dE[static_cast<size_t>(b) * F * D + e] += t[(e / D) * (D + 1) + (e % D)];
    __syncwarp();
  }
}

// pooled[b, h] = sum_d out[(b*D+d), h]   (xdeepfm/xdeepfm.py:180-181): thread per (b, 4 columns).
__global__ void __launch_bounds__(256)
cin_pool_kernel(const float* __restrict__ out, int B, int D, int H, float* __restrict__ pooled,
                int ldp) {
  const int hq = H >> 2;
  const long long n = static_cast<long long>(B) * hq;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < n; e += gridDim.x * 256LL) {
    const int b = static_cast<int>(e / hq), q = static_cast<int>(e % hq);
    float4 a = f4_zero();
    for (int d = 0; d < D; ++d)
      a = f4_add(a, ldg4(out + (static_cast<size_t>(b) * D + d) * H + q * 4));
    *reinterpret_cast<float4*>(pooled + static_cast<size_t>(b) * ldp + q * 4) = a;
  }
}

// dpre[(b*D+d), h] = (dpool[b, h] + dacc[(b*D+d), h]) * 1[out[(b*D+d), h] > 0]: the gradient of the
// sum-pool broadcast over d, plus what the next layer sent back (dacc, nullable), through the ReLU.
__global__ void __launch_bounds__(256)
cin_dpre_kernel(const float* __restrict__ dpool, int ldp, const float* __restrict__ dacc,
                const float* __restrict__ out, int B, int D, int H, float* __restrict__ dpre) {
  const int hq = H >> 2;
  const long long n = static_cast<long long>(B) * D * hq;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < n; e += gridDim.x * 256LL) {
    const long long r = e / hq;
    const int q = static_cast<int>(e % hq);
    const int b = static_cast<int>(r / D);
    float4 g = ldg4(dpool + static_cast<size_t>(b) * ldp + q * 4);
    if (dacc != nullptr) g = f4_add(g, ldg4(dacc + static_cast<size_t>(r) * H + q * 4));
    const float4 o = ldg4(out + static_cast<size_t>(r) * H + q * 4);
    g.x = o.x > 0.f ? g.x : 0.f;
    g.y = o.y > 0.f ? g.y : 0.f;
    g.z = o.z > 0.f ? g.z : 0.f;
    g.w = o.w > 0.f ? g.w : 0.f;
    *reinterpret_cast<float4*>(dpre + static_cast<size_t>(r) * H + q * 4) = g;
  }
}

}  // namespace ctr

using namespace ctr;

extern "C" {

int ctr_cin_pool(const float* out, int B, int D, int H, float* pooled, int ld_pooled,
                 ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(out && pooled && B >= 0 && D > 0 && H > 0 && (H & 3) == 0 && (ld_pooled & 3) == 0 &&
                  ld_pooled >= H && aligned16(out) && aligned16(pooled),
              "ctr_cin_pool", "bad argument (H and ld_pooled multiples of 4, 16-byte aligned)");
  if (B == 0) return CTR_OK;
  const long long n = static_cast<long long>(B) * (H / 4);
  const int grid = static_cast<int>(std::min<long long>((n + 255) / 256, sm_count() * 8LL));
  cin_pool_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(out, B, D, H, pooled, ld_pooled);
  CTR_LAUNCH_CHECK("ctr_cin_pool");
}

int ctr_cin_dpre(const float* dpool, int ld_dpool, const float* dacc, const float* out, int B, int D,
                 int H, float* dpre, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(dpool && out && dpre && B >= 0 && D > 0 && H > 0 && (H & 3) == 0 && (ld_dpool & 3) == 0 &&
                  aligned16(dpool) && aligned16(dacc) && aligned16(out) && aligned16(dpre),
              "ctr_cin_dpre", "bad argument (H and ld_dpool multiples of 4, 16-byte aligned)");
  if (B == 0) return CTR_OK;
  const long long n = static_cast<long long>(B) * D * (H / 4);
  const int grid = static_cast<int>(std::min<long long>((n + 255) / 256, sm_count() * 16LL));
  cin_dpre_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(dpool, ld_dpool, dacc, out, B, D, H, dpre);
  CTR_LAUNCH_CHECK("ctr_cin_dpre");
}

int ctr_transpose_fd(const float* E, int B, int F, int D, float* Xt, int ld, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(E && Xt && B >= 0 && F > 0 && D > 0 && ld >= F, "ctr_transpose_fd", "bad argument");
  const size_t smem = static_cast<size_t>(8) * F * (D + 1) * sizeof(float);
  CTR_REQUIRE(smem <= 48 * 1024, "ctr_transpose_fd", "F*D too large");
  if (B == 0) return CTR_OK;
  const int grid = std::min((B + 7) / 8, sm_count() * 8);
  transpose_fd_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(E, B, F, D, Xt, ld);
  CTR_LAUNCH_CHECK("ctr_transpose_fd");
}

int ctr_transpose_df_add(const float* dXt, int ld, int B, int F, int D, float* dE,
                         ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(dXt && dE && B >= 0 && F > 0 && D > 0 && ld >= F, "ctr_transpose_df_add",
              "bad argument");
  const size_t smem = static_cast<size_t>(8) * F * (D + 1) * sizeof(float);
  CTR_REQUIRE(smem <= 48 * 1024, "ctr_transpose_df_add", "F*D too large");
  if (B == 0) return CTR_OK;
  const int grid = std::min((B + 7) / 8, sm_count() * 8);
  transpose_df_add_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(dXt, ld, B, F, D,
                                                                                 dE);
  CTR_LAUNCH_CHECK("ctr_transpose_df_add");
}

int64_t ctr_cin_workspace_bytes(int B, int D, int m, int Hp, int H, int prec) {
  if (prec == CTR_CIN_FP32) return 0;
  return cin_tc_workspace_bytes(B, D, m, Hp, H, prec);
}

int ctr_cin_layer_fwd(const float* X0t, int ld0, const float* Xp, int ldp, const float* W,
                      const float* bias, int B, int D, int m, int Hp, int H, float* out, int prec,
                      void* workspace, int64_t workspace_bytes, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(X0t && Xp && W && out, "ctr_cin_layer_fwd", "null pointer");
  CTR_REQUIRE(B >= 0 && D > 0 && m > 0 && Hp > 0 && H > 0 && ld0 >= m && ldp >= Hp,
              "ctr_cin_layer_fwd", "bad shape");
  if (B == 0) return CTR_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int M = B * D;
  if (prec == CTR_CIN_FP32) {
    int r = cin_contract_launch(X0t, ld0, m, Xp, ldp, Hp, W, static_cast<long long>(Hp) * H, H, 1,
                                bias, out, H, M, H, 1, 0, st);
    if (r != CTR_OK) return r;
    CTR_LAUNCH_CHECK("ctr_cin_layer_fwd");
  }
  return cin_tc_layer_fwd(X0t, ld0, Xp, ldp, W, bias, B, D, m, Hp, H, out, prec, workspace,
                          workspace_bytes, st);
}

int ctr_cin_layer_bwd(const float* X0t, int ld0, const float* Xp, int ldp, const float* W,
                      const float* dpre, int B, int D, int m, int Hp, int H, float* dX0t,
                      float* dXp, float* dW, float* dbias, int prec, void* workspace,
                      int64_t workspace_bytes, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(X0t && Xp && W && dpre && dW, "ctr_cin_layer_bwd", "null pointer");
  CTR_REQUIRE(B >= 0 && D > 0 && m > 0 && Hp > 0 && H > 0 && ld0 >= m && ldp >= Hp,
              "ctr_cin_layer_bwd", "bad shape");
  if (B == 0) return CTR_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int M = B * D;
  if (prec != CTR_CIN_FP32)
    return cin_tc_layer_bwd(X0t, ld0, Xp, ldp, W, dpre, B, D, m, Hp, H, dX0t, dXp, dW, dbias, prec,
                            workspace, workspace_bytes, st);
  const long long HpH = static_cast<long long>(Hp) * H;
  int r;
  if (dXp != nullptr) {   // n = j: Wg[i, h, j] = W[(i*Hp + j)*H + h]
    r = cin_contract_launch(X0t, ld0, m, dpre, H, H, W, HpH, 1, H, nullptr, dXp, ldp, M, Hp, 0, 1, st);
    if (r != CTR_OK) return r;
  }
  if (dX0t != nullptr) {  // n = i: Wg[j, h, i] = W[(i*Hp + j)*H + h]
    r = cin_contract_launch(Xp, ldp, Hp, dpre, H, H, W, H, 1, HpH, nullptr, dX0t, ld0, M, m, 0, 1, st);
    if (r != CTR_OK) return r;
  }
  {
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
  CTR_LAUNCH_CHECK("ctr_cin_layer_bwd");
}

}  // extern "C"
