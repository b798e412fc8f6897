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

}  // namespace ctr

using namespace ctr;

extern "C" {

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
