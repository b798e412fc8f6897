// Dense-tower GEMMs on tcgen05:  out[M, N] (op)= A[M, K] . B[N, K]^T  in 3xTF32.
//
// One kernel serves the three layer GEMMs of the tower without any transposed or split copies
// of the operands in global memory:
//   forward        out = X . W          A = X  (K-major),   B = W    (MN-major: n contiguous)
//   backward data  dX  = dpre . W^T     A = dpre (K-major), B = W    (K-major: reduction over n)
//   backward wts   dW += X^T . dpre     A = X  (MN-major),  B = dpre (MN-major), split over rows
// Operand k-blocks ([128 | NT] x 32 fp32) are streamed by TMA (128-byte swizzles) through a
// 2-7 stage ring.  tcgen05.mma.kind::tf32 reads the top 19 bits of each fp32 word, so the raw
// tile doubles as the "hi" operand; the epilogue warps compute the "lo" tiles
// (x - trunc_tf32(x), elementwise, so the swizzled layout is irrelevant) into a 2-slot smem
// ring while the tensor core runs the (hi, hi) pass, then (lo, hi) and (hi, lo) follow:
//   warp 0   TMA producer            warp 1   MMA issuer           warp 2   TMEM allocator
//   warps 2-7  hi/lo splitter during the main loop
//   warps 4-7  epilogue: TMEM -> registers -> bias / ReLU / BN column statistics -> smem
//              transpose -> 128-byte row segments stored (or vector RED for the split-K dW)
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
  int stages;             // raw (hi) ring depth, <= 8
  int lo_slots;           // lo ring depth, 1 or 2 (0 when the lo operands arrive pre-split)
  int presplit;           // lo tiles are TMA-loaded from global (A_lo / B_lo), no splitter warps
  uint32_t b_bytes;       // bytes of one B k-block in shared memory
  uint32_t stage_bytes;   // 16 KB (A) + b_bytes rounded up to 1 KB; raw and lo slots alike
  uint32_t tmem_cols;
  int n_acc;              // independent TMEM accumulators the MMAs rotate over (summed in the epilogue)
  int acc_stride;         // TMEM columns between accumulators (NT rounded up to 32)
  uint32_t idesc;
  uint64_t desc_k, desc_mn;
  float* out;
  int ldo;
  const float* bias;      // EPI_FWD
  float* stats;           // EPI_FWD: [2][N] column sums of out and out^2 (nullable)
  int relu;
};

constexpr uint32_t kTcgMnBox = 32 * kTcKB * 4;   // one MN-major TMA box: 32 k-rows x 128 B
constexpr int kTcgSplitThreads = 192;            // warps 2-7 compute the lo tiles

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
               const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmBlo,
               const TcGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* lo_ring = smem + static_cast<size_t>(p.stages) * p.stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(lo_ring + static_cast<size_t>(p.lo_slots) * p.stage_bytes);
  uint64_t* full = bars;             // [stages]   TMA landed
  uint64_t* empty = bars + 8;        // [stages]   MMAs that read the raw stage retired
  uint64_t* conv = bars + 16;        // [lo_slots] lo tiles written
  uint64_t* lo_empty = bars + 18;    // [lo_slots] MMAs that read the lo slot retired
  uint64_t* t_full = bars + 20;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kTcBM, n0 = blockIdx.y * p.NT;
  const int kb_total = (p.K + kTcKB - 1) / kTcKB;
  const int kb_beg = blockIdx.z * p.kb_per_split;
  const int nkb = max(0, min(kb_total, kb_beg + p.kb_per_split) - kb_beg);
  const uint32_t ab_bytes = kTcABytes + p.b_bytes;
  // pre-split operands: a stage is [A | B | A_lo | B_lo], the lo half `lo_off` bytes in
  const uint32_t lo_off = p.stage_bytes >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < p.lo_slots; ++s) {
      mbar_init(&conv[s], kTcgSplitThreads);
      mbar_init(&lo_empty[s], 1);
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
      mbar_expect_tx(&full[st], p.presplit ? 2 * ab_bytes : ab_bytes);
      const int kk = (kb_beg + kb) * kTcKB;
      for (int half = 0; half <= p.presplit; ++half) {
        uint8_t* sa = smem + static_cast<size_t>(st) * p.stage_bytes + half * lo_off;
        uint8_t* sb = sa + kTcABytes;
        const CUtensorMap* mA = half ? &tmAlo : &tmA;
        const CUtensorMap* mB = half ? &tmBlo : &tmB;
        if (p.a_mn) {
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_2d(sa + j * kTcgMnBox, mA, m0 + 32 * j, kk, &full[st]);
        } else {
          tma_load_2d(sa, mA, kk, m0, &full[st]);
        }
        if (p.b_mn) {
          const int nbox = static_cast<int>(p.b_bytes / kTcgMnBox);
          for (int j = 0; j < nbox; ++j) tma_load_2d(sb + j * kTcgMnBox, mB, n0 + 32 * j, kk, &full[st]);
        } else {
          tma_load_2d(sb, mB, kk, n0, &full[st]);
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // -------------------------------------------------------------------- MMA issuer
    const uint64_t da = p.a_mn ? p.desc_mn : p.desc_k, db = p.b_mn ? p.desc_mn : p.desc_k;
    const uint32_t sa_step = p.a_mn ? 1024u : 32u, sb_step = p.b_mn ? 1024u : 32u;
    // MMAs may rotate over n_acc independent TMEM accumulators (summed in the epilogue); with
    // n_acc = 1 (default) this is the usual single-accumulator k loop.
    int mi = 0;
    auto dst = [&]() -> uint32_t { return tmem_base + static_cast<uint32_t>((mi % p.n_acc) * p.acc_stride); };
    for (int kb = 0; kb < nkb; ++kb) {
      const uint32_t st = kb % p.stages, ph = (kb / p.stages) & 1;
      const uint32_t a_addr = smem_u32(smem + static_cast<size_t>(st) * p.stage_bytes);
      const uint32_t b_addr = a_addr + kTcABytes;
      mbar_wait(&full[st], ph);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        tc_mma_tf32(dst(), da | (((a_addr + k * sa_step) >> 4) & 0x3FFF),
                    db | (((b_addr + k * sb_step) >> 4) & 0x3FFF), p.idesc, mi >= p.n_acc);
        ++mi;
      }
      if (p.n_pass == 3) {
        uint32_t sl = 0, a_lo;
        if (p.presplit) {
          a_lo = a_addr + lo_off;            // landed together with the hi tiles
        } else {
          sl = kb % p.lo_slots;
          a_lo = smem_u32(lo_ring + static_cast<size_t>(sl) * p.stage_bytes);
          mbar_wait(&conv[sl], (kb / p.lo_slots) & 1);
          tc_fence_after();
        }
        const uint32_t b_lo = a_lo + kTcABytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // (lo, hi)
          tc_mma_tf32(dst(), da | (((a_lo + k * sa_step) >> 4) & 0x3FFF),
                      db | (((b_addr + k * sb_step) >> 4) & 0x3FFF), p.idesc, mi >= p.n_acc);
          ++mi;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // (hi, lo)
          tc_mma_tf32(dst(), da | (((a_addr + k * sa_step) >> 4) & 0x3FFF),
                      db | (((b_lo + k * sb_step) >> 4) & 0x3FFF), p.idesc, mi >= p.n_acc);
          ++mi;
        }
        if (!p.presplit) tc_commit(&lo_empty[sl]);
      }
      tc_commit(&empty[st]);
    }
    tc_commit(t_full);
  } else if (warp >= 2) {
    // ------------------------------------ hi/lo splitter (warps 2-7), then epilogue (warps 4-7)
    if (p.n_pass == 3 && !p.presplit) {
      const int ctid = threadIdx.x - 64;
      const int n16 = static_cast<int>(ab_bytes >> 4);
      for (int kb = 0; kb < nkb; ++kb) {
        const uint32_t st = kb % p.stages, ph = (kb / p.stages) & 1;
        const uint32_t sl = kb % p.lo_slots, lph = (kb / p.lo_slots) & 1;
        mbar_wait(&full[st], ph);
        mbar_wait(&lo_empty[sl], lph ^ 1);
        const float4* hi = reinterpret_cast<const float4*>(smem + static_cast<size_t>(st) * p.stage_bytes);
        float4* lo = reinterpret_cast<float4*>(lo_ring + static_cast<size_t>(sl) * p.stage_bytes);
        for (int i = ctid; i < n16; i += 4 * kTcgSplitThreads) {
          float4 x[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = i + u * kTcgSplitThreads;
            if (j < n16) x[u] = hi[j];
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = i + u * kTcgSplitThreads;
            if (j < n16) {
              float4 l;
              l.x = tcg_lo(x[u].x);
              l.y = tcg_lo(x[u].y);
              l.z = tcg_lo(x[u].z);
              l.w = tcg_lo(x[u].w);
              lo[j] = l;
            }
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(&conv[sl]);
      }
    }
    if (warp >= 4 && nkb > 0) {
      const int quarter = warp & 3;
      mbar_wait(t_full, 0);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
      const bool vec = (p.ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
      const int r_own = m0 + quarter * 32 + lane;          // this lane's TMEM row
      const int nlim = min(p.N, n0 + p.NT);                // first invalid output column
      // all MMAs have retired: the raw ring is free, use it to turn the row-per-lane TMEM
      // fragments into 128-byte row segments (4 rows x 128 B per store instruction)
      float4* sw = reinterpret_cast<float4*>(smem + quarter * (32 * 9 * 16));
      const int n_used = min(p.n_acc, nkb * 4 * p.n_pass);   // accumulators that were written
      for (int c = 0; c < p.NT; c += 32) {
        float v[32];
        tc_ld<32>(taddr + c, v);
        for (int a = 1; a < n_used; ++a) {
          float w[32];
          tc_ld<32>(taddr + static_cast<uint32_t>(a * p.acc_stride) + c, w);
#pragma unroll
          for (int t = 0; t < 32; ++t) v[t] += w[t];
        }
        const int nb = n0 + c;
        if (EPI == TCG_EPI_FWD) {
#pragma unroll
          for (int t = 0; t < 32; ++t) {
            const int n = nb + t;
            float x = 0.f;
            if (r_own < p.M && n < nlim) {
              x = v[t] + (p.bias != nullptr ? __ldg(p.bias + n) : 0.f);
              if (p.relu) x = fmaxf(x, 0.f);
            }
            v[t] = x;
          }
        }
#pragma unroll
        for (int t = 0; t < 8; ++t)
          sw[lane * 9 + t] = make_float4(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]);
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int row = it * 4 + (lane >> 3);
          const float4 q = sw[row * 9 + (lane & 7)];
          const int rr = m0 + quarter * 32 + row, n = nb + (lane & 7) * 4;
          if (rr < p.M && n < nlim) {
            float* o = p.out + static_cast<size_t>(rr) * p.ldo + n;
            if (vec && n + 4 <= nlim) {
              if (EPI == TCG_EPI_RED) red_add_v4(o, q);
              else *reinterpret_cast<float4*>(o) = q;
            } else {
              const float e[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (n + u < nlim) {
                  if (EPI == TCG_EPI_RED) red_add_f32(o + u, e[u]);
                  else o[u] = e[u];
                }
            }
          }
        }
        __syncwarp();
        if (EPI == TCG_EPI_FWD && p.stats != nullptr) {
          float q[32];
#pragma unroll
          for (int t = 0; t < 32; ++t) q[t] = v[t] * v[t];
          const float s1 = warp_colsum32(v, lane);
          const float s2 = warp_colsum32(q, lane);
          const int n = nb + lane;
          if (n < nlim) {
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
  if (const char* e = ctr_knob("CTR_TCG_MN_LAYOUT")) lay = strtoull(e, nullptr, 0);
  if (const char* e = ctr_knob("CTR_TCG_MN_LBO")) lbo = strtoull(e, nullptr, 0);
  if (const char* e = ctr_knob("CTR_TCG_MN_SBO")) sbo = strtoull(e, nullptr, 0);
  return (lbo << 16) | (sbo << 32) | (ver << 46) | (lay << 61);
}

// A: K-major [M, K] pitch lda, or MN-major stored [K, M] pitch lda.  B likewise with N.
// splits > 1 only with TCG_EPI_RED.
// A_lo / B_lo (both or neither): the lo halves of the 3xTF32 split, same shapes and pitches as A / B,
// produced with tcg_lo() by whoever wrote A / B; the kernel then has no CUDA-core work
// in its main loop.
template <int EPI>
static int tc_gemm_launch(const float* A, int lda, bool a_mn, const float* Bm, int ldb, bool b_mn,
                          int M, int N, int K, int NT, int splits, float* out, int ldo,
                          const float* bias, float* stats, int relu, cudaStream_t st,
                          const char* fn, const float* A_lo = nullptr, const float* B_lo = nullptr,
                          int max_stages = 8) {
  CTR_REQUIRE((lda & 3) == 0 && (ldb & 3) == 0 && aligned16(A) && aligned16(Bm), fn,
              "tensor-core path needs 16-byte aligned operands");
  CTR_REQUIRE(NT % 16 == 0 && NT >= 16 && NT <= 256, fn, "bad NT");
  const bool presplit = A_lo != nullptr && B_lo != nullptr;
  CTR_REQUIRE(!presplit || (aligned16(A_lo) && aligned16(B_lo)), fn, "lo operands must be 16-byte aligned");
  CUtensorMap tA, tB, tAlo, tBlo;
  int r = a_mn ? make_map(&tA, A, K, M, lda, 32, true) : make_map(&tA, A, M, K, lda, kTcBM);
  if (r != CTR_OK) return r;
  r = b_mn ? make_map(&tB, Bm, K, N, ldb, 32, true) : make_map(&tB, Bm, N, K, ldb, NT);
  if (r != CTR_OK) return r;
  tAlo = tA;
  tBlo = tB;
  if (presplit) {
    r = a_mn ? make_map(&tAlo, A_lo, K, M, lda, 32, true) : make_map(&tAlo, A_lo, M, K, lda, kTcBM);
    if (r != CTR_OK) return r;
    r = b_mn ? make_map(&tBlo, B_lo, K, N, ldb, 32, true) : make_map(&tBlo, B_lo, N, K, ldb, NT);
    if (r != CTR_OK) return r;
  }
  TcGemmParams p{};
  p.M = M; p.N = N; p.K = K; p.NT = NT; p.a_mn = a_mn; p.b_mn = b_mn; p.n_pass = 3;
  if (const char* e = ctr_knob("CTR_TCG_PASSES")) p.n_pass = atoi(e) == 1 ? 1 : 3;
  const int kb_total = (K + kTcKB - 1) / kTcKB;
  splits = std::max(1, std::min(splits, kb_total));
  p.kb_per_split = (kb_total + splits - 1) / splits;
  splits = (kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.b_bytes = b_mn ? static_cast<uint32_t>((NT + 31) / 32) * kTcgMnBox
                   : static_cast<uint32_t>(NT) * kTcKB * 4;
  p.stage_bytes = kTcABytes + ((p.b_bytes + 1023u) & ~1023u);
  p.lo_slots = 2;
  p.presplit = (presplit && p.n_pass == 3) ? 1 : 0;
  if (p.presplit) {
    p.stage_bytes *= 2;
    p.lo_slots = 0;
  }
  p.stages = std::max(2, std::min(8, static_cast<int>((200u * 1024u) / p.stage_bytes) - p.lo_slots));
  // a CTA never has more loads in flight than k-blocks; off-critical-path callers cap the ring
  // further so that the kernels they run beside keep their shared memory (max_stages)
  p.stages = std::max(2, std::min(p.stages, std::min(max_stages, p.kb_per_split)));
  if (const char* e = ctr_knob("CTR_TCG_STAGES")) p.stages = std::max(1, std::min(p.stages, atoi(e)));
  p.acc_stride = (NT + 31) / 32 * 32;
  // measured: one accumulator is as fast as many (the MMAs are operand-fetch bound, ~160 cycles
  // each whatever N is), and the epilogue then has nothing to add up; CTR_TCG_NACC > 1 rotates
  p.n_acc = 1;
  if (const char* e = ctr_knob("CTR_TCG_NACC"))
    p.n_acc = std::max(1, std::min(std::min(12, 512 / p.acc_stride), atoi(e)));
  p.tmem_cols = 32;
  while (p.tmem_cols < static_cast<uint32_t>(p.n_acc * p.acc_stride)) p.tmem_cols <<= 1;
  p.idesc = cin_idesc(NT) | (a_mn ? 1u << 15 : 0u) | (b_mn ? 1u << 16 : 0u);
  p.desc_k = cin_desc_hi();
  p.desc_mn = tcg_desc_mn();
  p.out = out; p.ldo = ldo; p.bias = bias; p.stats = stats; p.relu = relu;
  const size_t smem = static_cast<size_t>(p.stages + p.lo_slots) * p.stage_bytes + 256 + 1024;
  static bool optin = false;
  if (!optin) {
    cudaFuncSetAttribute(tc_gemm_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         static_cast<int>(200 * 1024 + 256 + 1024));
    optin = true;
  }
  dim3 grid((M + kTcBM - 1) / kTcBM, (N + NT - 1) / NT, splits);
  tc_gemm_kernel<EPI><<<grid, 256, smem, st>>>(tA, tB, tAlo, tBlo, p);
  return check_cuda(cudaGetLastError(), fn);
}

}  // namespace ctr
