// Dense-tower GEMMs on tcgen05:  out[M, N] (op)= A[M, K] . B[N, K]^T  in 3xTF32.
//
// One kernel serves the three layer GEMMs of the tower without any transposed or split copies
// of the operands in global memory:
//   forward        out = X . W          A = X  (K-major),   B = W    (MN-major: n contiguous)
//   backward data  dX  = dpre . W^T     A = dpre (K-major), B = W    (K-major: reduction over n)
//   backward wts   dW += X^T . dpre     A = X  (MN-major),  B = dpre (MN-major), split over rows
// Operand k-blocks ([128 | NT] x 32 fp32) are streamed by TMA (128-byte swizzles) through a
// 2-3 stage ring.  tcgen05.mma.kind::tf32 reads the top 19 bits of each fp32 word, so the raw
// tile doubles as the "hi" operand; the epilogue warps compute the "lo" tiles
// (x - trunc_tf32(x), elementwise, so the swizzled layout is irrelevant) into a second smem
// buffer while the tensor core runs the (hi, hi) pass, then (lo, hi) and (hi, lo) follow:
//   warp 0   TMA producer            warp 1   MMA issuer           warp 2   TMEM allocator
//   warps 4-7  hi/lo splitter during the main loop, then the epilogue (TMEM -> registers ->
//              bias / ReLU / BN column statistics / store, or vector RED for the split-K dW)
#pragma once
#include "tc_common.cuh"

namespace ctr {

enum { TCG_EPI_FWD = 0, TCG_EPI_STORE = 1, TCG_EPI_RED = 2 };

struct TcGemmParams {
  int M, N, K;            // out is [M, N]; K = reduction length
  int NT;                 // output columns per CTA (multiple of 16, <= 256)
  int a_mn, b_mn;         // operand is MN-major (else K-major)
  int kb_per_split;       // k-blocks (32 of K) per blockIdx.z
  int n_pass;             // 1 = plain tf32, 3 = 3xTF32
  int stages;
  uint32_t b_bytes;       // bytes of one B k-block in shared memory
  uint32_t lo_off;        // offset of the lo tiles inside a stage (= 16 KB + max B bytes)
  uint32_t stage_bytes;   // 2 * lo_off
  uint32_t tmem_cols;
  uint32_t idesc;
  uint64_t desc_k, desc_mn;
  float* out;
  int ldo;
  const float* bias;      // EPI_FWD
  float* stats;           // EPI_FWD: [2][N] column sums of out and out^2 (nullable)
  int relu;
};

constexpr uint32_t kTcgMnBox = 32 * kTcKB * 4;   // one MN-major TMA box: 32 k-rows x 128 B

__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Column sums over the 32 lanes of a warp for 32 columns held as v[0..31] in every lane:
// 31 shuffles, lane c ends up with the sum of column c.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = upper ? v[i] : v[i + o];
      const float keep = upper ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}

template <int EPI>
__global__ void __launch_bounds__(256, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const TcGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(p.stages) * p.stage_bytes);
  uint64_t* full = bars;             // [stages] TMA landed
  uint64_t* conv = bars + 4;         // [stages] lo tiles written
  uint64_t* empty = bars + 8;        // [stages] MMAs retired
  uint64_t* t_full = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kTcBM, n0 = blockIdx.y * p.NT;
  const int kb_total = (p.K + kTcKB - 1) / kTcKB;
  const int kb_beg = blockIdx.z * p.kb_per_split;
  const int nkb = max(0, min(kb_total, kb_beg + p.kb_per_split) - kb_beg);
  const uint32_t ab_bytes = kTcABytes + p.b_bytes;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&conv[s], 128);
      mbar_init(&empty[s], 1);
    }
    mbar_init(t_full, 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(p.tmem_cols)
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
      const uint32_t st = kb % p.stages, ph = (kb / p.stages) & 1;
      mbar_wait(&empty[st], ph ^ 1);
      mbar_expect_tx(&full[st], ab_bytes);
      uint8_t* sa = smem + static_cast<size_t>(st) * p.stage_bytes;
      uint8_t* sb = sa + kTcABytes;
      const int kk = (kb_beg + kb) * kTcKB;
      if (p.a_mn) {
#pragma unroll
        for (int j = 0; j < 4; ++j) tma_load_2d(sa + j * kTcgMnBox, &tmA, m0 + 32 * j, kk, &full[st]);
      } else {
        tma_load_2d(sa, &tmA, kk, m0, &full[st]);
      }
      if (p.b_mn) {
        const int nbox = static_cast<int>(p.b_bytes / kTcgMnBox);
        for (int j = 0; j < nbox; ++j) tma_load_2d(sb + j * kTcgMnBox, &tmB, n0 + 32 * j, kk, &full[st]);
      } else {
        tma_load_2d(sb, &tmB, kk, n0, &full[st]);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // -------------------------------------------------------------------- MMA issuer
    const uint64_t da = p.a_mn ? p.desc_mn : p.desc_k, db = p.b_mn ? p.desc_mn : p.desc_k;
    const uint32_t sa_step = p.a_mn ? 1024u : 32u, sb_step = p.b_mn ? 1024u : 32u;
    uint32_t accum = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      const uint32_t st = kb % p.stages, ph = (kb / p.stages) & 1;
      const uint32_t a_addr = smem_u32(smem + static_cast<size_t>(st) * p.stage_bytes);
      const uint32_t b_addr = a_addr + kTcABytes;
      mbar_wait(&full[st], ph);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        tc_mma_tf32(tmem_base, da | (((a_addr + k * sa_step) >> 4) & 0x3FFF),
                    db | (((b_addr + k * sb_step) >> 4) & 0x3FFF), p.idesc, accum);
        accum = 1;
      }
      if (p.n_pass == 3) {
        mbar_wait(&conv[st], ph);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)   // (lo, hi)
          tc_mma_tf32(tmem_base, da | (((a_addr + p.lo_off + k * sa_step) >> 4) & 0x3FFF),
                      db | (((b_addr + k * sb_step) >> 4) & 0x3FFF), p.idesc, 1);
#pragma unroll
        for (int k = 0; k < 4; ++k)   // (hi, lo)
          tc_mma_tf32(tmem_base, da | (((a_addr + k * sa_step) >> 4) & 0x3FFF),
                      db | (((b_addr + p.lo_off + k * sb_step) >> 4) & 0x3FFF), p.idesc, 1);
      }
      tc_commit(&empty[st]);
    }
    tc_commit(t_full);
  } else if (warp >= 4) {
    // ------------------------------------------------------- hi/lo splitter, then epilogue
    const int ctid = threadIdx.x - 128;
    if (p.n_pass == 3) {
      const int n16 = static_cast<int>(ab_bytes >> 4);
      for (int kb = 0; kb < nkb; ++kb) {
        const uint32_t st = kb % p.stages, ph = (kb / p.stages) & 1;
        mbar_wait(&full[st], ph);
        const float4* hi = reinterpret_cast<const float4*>(smem + static_cast<size_t>(st) * p.stage_bytes);
        float4* lo = reinterpret_cast<float4*>(smem + static_cast<size_t>(st) * p.stage_bytes + p.lo_off);
        for (int i = ctid; i < n16; i += 128) {
          const float4 x = hi[i];
          float4 l;
          l.x = round_tf32(x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u));
          l.y = round_tf32(x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u));
          l.z = round_tf32(x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u));
          l.w = round_tf32(x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u));
          lo[i] = l;
        }
        fence_proxy_async_smem();
        mbar_arrive(&conv[st]);
      }
    }
    if (nkb > 0) {
      const int quarter = warp & 3;
      const int r = m0 + quarter * 32 + lane;
      mbar_wait(t_full, 0);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
      const bool vec = (p.ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
      for (int c = 0; c < p.NT; c += 32) {
        float v[32];
        tc_ld<32>(taddr + c, v);
        const int nb = n0 + c;
        if (EPI == TCG_EPI_FWD) {
#pragma unroll
          for (int t = 0; t < 32; ++t) {
            const int n = nb + t;
            float x = 0.f;
            if (r < p.M && n < p.N && c + t < p.NT) {
              x = v[t] + (p.bias != nullptr ? __ldg(p.bias + n) : 0.f);
              if (p.relu) x = fmaxf(x, 0.f);
            }
            v[t] = x;
          }
        }
        if (r < p.M) {
          float* o = p.out + static_cast<size_t>(r) * p.ldo + nb;
#pragma unroll
          for (int t = 0; t < 32; t += 4) {
            const int lim = min(p.N - nb, p.NT - c);     // valid columns in this chunk
            if (t + 4 <= lim && vec) {
              const float4 q = make_float4(v[t], v[t + 1], v[t + 2], v[t + 3]);
              if (EPI == TCG_EPI_RED) red_add_v4(o + t, q);
              else *reinterpret_cast<float4*>(o + t) = q;
            } else {
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (t + u < lim) {
                  if (EPI == TCG_EPI_RED) red_add_f32(o + t + u, v[t + u]);
                  else o[t + u] = v[t + u];
                }
            }
          }
        }
        if (EPI == TCG_EPI_FWD && p.stats != nullptr) {
          float q[32];
#pragma unroll
          for (int t = 0; t < 32; ++t) q[t] = v[t] * v[t];
          const float s1 = warp_colsum32(v, lane);
          const float s2 = warp_colsum32(q, lane);
          const int n = nb + lane;
          if (n < p.N && c + lane < p.NT) {
            red_add_f32(p.stats + n, s1);
            red_add_f32(p.stats + p.N + n, s2);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(p.tmem_cols)
                 : "memory");
  }
}

// MN-major tf32 operands only exist in the SWIZZLE_128B_BASE32B layout (type 1; TMA mode
// 128B_ATOM_32B): 32 floats contiguous along M/N per 128-byte row, 32-byte chunks XORed with the
// row index mod 4, 4 k-rows per 512-byte atom.  LBO = distance between 32-float M/N groups (one
// TMA box), SBO = distance between 4-row k groups.
static uint64_t tcg_desc_mn() {
  uint64_t lbo = kTcgMnBox >> 4, sbo = 512 >> 4, ver = 1, lay = 1;
  if (const char* e = getenv("CTR_TCG_MN_LAYOUT")) lay = strtoull(e, nullptr, 0);
  if (const char* e = getenv("CTR_TCG_MN_LBO")) lbo = strtoull(e, nullptr, 0);
  if (const char* e = getenv("CTR_TCG_MN_SBO")) sbo = strtoull(e, nullptr, 0);
  return (lbo << 16) | (sbo << 32) | (ver << 46) | (lay << 61);
}

// A: K-major [M, K] pitch lda, or MN-major stored [K, M] pitch lda.  B likewise with N.
// splits > 1 only with TCG_EPI_RED.
template <int EPI>
static int tc_gemm_launch(const float* A, int lda, bool a_mn, const float* Bm, int ldb, bool b_mn,
                          int M, int N, int K, int NT, int splits, float* out, int ldo,
                          const float* bias, float* stats, int relu, cudaStream_t st,
                          const char* fn) {
  CTR_REQUIRE((lda & 3) == 0 && (ldb & 3) == 0 && aligned16(A) && aligned16(Bm), fn,
              "tensor-core path needs 16-byte aligned operands");
  CTR_REQUIRE(NT % 16 == 0 && NT >= 16 && NT <= 256, fn, "bad NT");
  CUtensorMap tA, tB;
  int r = a_mn ? make_map(&tA, A, K, M, lda, 32, true) : make_map(&tA, A, M, K, lda, kTcBM);
  if (r != CTR_OK) return r;
  r = b_mn ? make_map(&tB, Bm, K, N, ldb, 32, true) : make_map(&tB, Bm, N, K, ldb, NT);
  if (r != CTR_OK) return r;
  TcGemmParams p{};
  p.M = M; p.N = N; p.K = K; p.NT = NT; p.a_mn = a_mn; p.b_mn = b_mn; p.n_pass = 3;
  if (const char* e = getenv("CTR_TCG_PASSES")) p.n_pass = atoi(e) == 1 ? 1 : 3;
  const int kb_total = (K + kTcKB - 1) / kTcKB;
  splits = std::max(1, std::min(splits, kb_total));
  p.kb_per_split = (kb_total + splits - 1) / splits;
  splits = (kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.b_bytes = b_mn ? static_cast<uint32_t>((NT + 31) / 32) * kTcgMnBox
                   : static_cast<uint32_t>(NT) * kTcKB * 4;
  const uint32_t bmax = NT <= 128 ? 128 * kTcKB * 4 : 256 * kTcKB * 4;
  p.lo_off = kTcABytes + bmax;
  p.stage_bytes = 2 * p.lo_off;
  p.stages = NT <= 128 ? 3 : 2;
  p.tmem_cols = NT <= 32 ? 32 : NT <= 64 ? 64 : NT <= 128 ? 128 : 256;
  p.idesc = cin_idesc(NT) | (a_mn ? 1u << 15 : 0u) | (b_mn ? 1u << 16 : 0u);
  p.desc_k = cin_desc_hi();
  p.desc_mn = tcg_desc_mn();
  p.out = out; p.ldo = ldo; p.bias = bias; p.stats = stats; p.relu = relu;
  const size_t smem = static_cast<size_t>(p.stages) * p.stage_bytes + 256 + 1024;
  static bool optin = false;
  if (!optin) {
    cudaFuncSetAttribute(tc_gemm_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         static_cast<int>(3 * 65536 + 256 + 1024));
    optin = true;
  }
  dim3 grid((M + kTcBM - 1) / kTcBM, (N + NT - 1) / NT, splits);
  tc_gemm_kernel<EPI><<<grid, 256, smem, st>>>(tA, tB, p);
  return check_cuda(cudaGetLastError(), fn);
}

}  // namespace ctr
