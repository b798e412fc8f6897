// Fused lookup + first tower layer:  ids -> gather -> FM terms -> act0 = relu(E . W0 + b0)
// in ONE kernel; the concatenated embedding never makes the round trip through global memory
// between the lookup and the layer that consumes it (deepfm/deepfm.py:76-101, fm/fm.py:76-129).
//
// The unfused step runs embed_fwd (13-15 us at B = 4096) and then the split-K tcgen05 GEMM
// (11.5 us + 2 us of launch gap): the GEMM's CTAs re-read E / E_lo, which the lookup's CTAs have
// just written, and RED their partial sums into a zeroed [B, H0] buffer that the next kernel
// reads back.  Here a cluster of 4 CTAs owns 128 samples; CTA kg of the cluster owns the k-group
// of fields [kg*FPC, (kg+1)*FPC):
//   warps 0-7   id pipeline for the CTA's (sample, field) pairs, then ALL row loads of the CTA
//               at once (LPR lanes x 16 B per row, 20 loads in flight per lane); the rows go
//               straight into the K-major 128B-swizzled A tiles of the MMA (hi = the fp32 word,
//               lo = tcg_lo) - and out to E / E_lo for the weight-gradient GEMM of the backward,
//               fire and forget; FM partial sums in registers
//   warp 8      TMA producer of the W0 k-blocks (hi / lo, MN-major boxes), 2-stage ring
//   warp 9      tcgen05.mma issuer: k-block kb is multiplied as soon as its two fields have
//               landed (3xTF32: 12 MMAs per k-block), accumulator [128, H0] in TMEM
//   warps 8-11  TMEM -> shared memory (the partial product of this k-group)
// and the 4 partial products / FM partial sums are added over DISTRIBUTED SHARED MEMORY: CTA r
// reduces rows 32r..32r+31 of all four, adds the bias, applies the ReLU and writes act0, S, y1,
// y2; the BN column sums of act0 leave as one partial block per cluster (no atomics, no zeroed
// buffer, nothing for the next kernel to wait on but the kernel boundary itself).
#include <cuda.h>

#include <algorithm>

#include "criteo_ids.cuh"
#include "tc_gemm.cuh"

namespace ctr {

constexpr int kEtGather = 256;        // threads of warps 0-7
constexpr int kEtThreads = 384;
constexpr int kEtKS = 4;              // CTAs per cluster = k-groups
constexpr int kEtMaxFPC = 10;         // fields per CTA (even): <= 5 k-blocks of 2 fields
constexpr int kEtMaxKB = kEtMaxFPC / 2;
constexpr int kEtD = 16;
constexpr int kEtPartPitch = 20;      // floats per row of the FM partials: S[16] | q | y1 | pad

struct EtParams {
  const float* table;
  const float* w1;
  long long ld, ld1;
  unsigned long long w1_fields;
  const float* xcont;
  const long long* xcat;
  const ctr_field_desc* fields;
  const float* bnd;
  int n_cont, n_cat, n_bnd;
  const int* rows_in;     // nullable: ids computed earlier (ctr_criteo_rows), the id stage is skipped
  int* rows_out;
  int* status;
  float* E;
  float* E_lo;
  float* S;
  float* y1;
  float* y2;
  const float* bias;
  float* act0;
  float* stats_part;      // nullable: [mtiles][2][N] column sums of act0 and act0^2 per cluster
  float* zero_buf;
  long long zero_n4;
  int B, F, N, NT, fpc, kbc;
  uint32_t b_bytes;       // one B k-block tile (hi or lo)
  uint32_t bar_off;       // byte offset of the barrier block (behind the tiles and the reuse region)
  uint32_t idesc;
  uint64_t desc_k, desc_mn;
  unsigned long long* timing;   // nullable: 10 words, %globaltimer of CTA (0,0) at the phase boundaries
};

// Optional phase profile (ctr_embed_tower_timing): thread 0 of CTA (0, 0) stamps %globaltimer.
__device__ __forceinline__ void et_stamp(const EtParams& p, int slot) {
  if (p.timing != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0)
    p.timing[slot] = gtime_ns();
}

__device__ __forceinline__ uint32_t et_sw128(int row, int c) {
  return static_cast<uint32_t>(row) * 128u + (static_cast<uint32_t>(c ^ (row & 7)) << 4);
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ float4 ld_cluster4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr)
               : "memory");
  return v;
}
__device__ __forceinline__ float ld_cluster1(uint32_t addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

__global__ void __launch_bounds__(kEtThreads, 1)
embed_tower_fwd_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmWlo,
                       const EtParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  // [B: kbc x (hi | lo), all resident] [A ring: 2 x (hi | lo)] [barriers]
  // W0's k-blocks are requested by TMA when the kernel starts - they arrive while the ids are
  // computed and the rows are in flight - and the gathered rows wait in REGISTERS for their A stage,
  // so no TMA latency sits between two k-blocks of MMAs.
  uint8_t* b_all = smem;
  uint8_t* a_ring = smem + static_cast<size_t>(p.kbc) * 2 * p.b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.bar_off);
  uint64_t* a_full = bars;            // [2] both fields of the k-block are in the A stage
  uint64_t* a_empty = bars + 2;       // [2] the MMAs that read the A stage have retired
  uint64_t* b_full = bars + 4;        // [kEtMaxKB]
  uint64_t* t_full = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
  // after the MMAs the B region is reused: the partial product [128][PD] and the FM partials
  const int PD = p.NT + 4;
  float* dump = reinterpret_cast<float*>(smem);
  float* part = reinterpret_cast<float*>(smem + 72 * 1024);        // [128][kEtPartPitch]
  float* colp = part + 128 * kEtPartPitch;                          // [2][128] column partials
  float* tile = colp + 256;                                         // [32][PD] this CTA's act0 rows

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * kTcBM;
  const int kg = blockIdx.y;
  const int B = p.B, F = p.F;
  et_stamp(p, 0);

  if (tid == 0) {
    for (int s = 0; s < kEtMaxKB; ++s) mbar_init(&b_full[s], 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_full[s], kEtGather);
      mbar_init(&a_empty[s], 1);
    }
    mbar_init(t_full, 1);
    mbar_fence_init();
  }
  if (warp == 10) {
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
  et_stamp(p, 8);

  float4 Sp[2] = {f4_zero(), f4_zero()};
  float qp[2] = {0.f, 0.f}, y1p[2] = {0.f, 0.f};
  const int r8 = lane & 7, c = lane >> 3;

  if (warp < 8) {
    // ------------------------------------------------------------ ids, gather, A tiles
    if (p.zero_buf != nullptr && blockIdx.x == 0 && kg == 0)
      for (long long i = tid; i < p.zero_n4; i += kEtGather)
        reinterpret_cast<float4*>(p.zero_buf)[i] = f4_zero();
    const int f0 = kg * p.fpc;
    int rid[2 * kEtMaxFPC];
    if (p.rows_in != nullptr) {
      // the ids were computed ahead of the step (on the copy stream, beside the previous step)
#pragma unroll
      for (int i = 0; i < 2 * kEtMaxFPC; ++i) {
        const int j = i >> 1, bb = m0 + 16 * warp + 8 * (i & 1) + r8, f = f0 + j;
        rid[i] = (j < p.fpc && bb < B && f < F) ? __ldg(p.rows_in + static_cast<size_t>(bb) * F + f) : -1;
      }
    } else {
      // ids of the CTA's 128 x fpc (sample, field) pairs: every thread computes its share with
      // independent loads (no shuffle between them), staged in the not yet used tail of the A
      // ring, then each lane picks up the ids of its rows
      int* ids_s = reinterpret_cast<int*>(a_ring + 4 * kTcABytes) - 128 * kEtMaxFPC;
      ctr_field_desc* s_fields = reinterpret_cast<ctr_field_desc*>(reinterpret_cast<uint8_t*>(ids_s) - 4096);
      float* s_bnd = reinterpret_cast<float*>(s_fields + kEtMaxFPC + 2);       // <= 512 boundaries
      {
        // the CTA's field descriptors and the bucket boundaries first (one L2 round trip), so that
        // the id arithmetic itself only waits for the feature values
        const int nf = min(p.fpc, F - f0);
        const int* src = reinterpret_cast<const int*>(p.fields + f0);
        for (int i = tid; i < nf * 8; i += kEtGather) reinterpret_cast<int*>(s_fields)[i] = __ldg(src + i);
        for (int i = tid; i < p.n_bnd; i += kEtGather) s_bnd[i] = __ldg(p.bnd + i);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      et_stamp(p, 9);
      {
        // all feature loads of the thread's (sample, field) pairs first, then the id arithmetic
        constexpr int NU = (128 * kEtMaxFPC + kEtGather - 1) / kEtGather;
        const int npair = 128 * p.fpc;
        CriteoRaw raw[NU];
  #pragma unroll
        for (int u = 0; u < NU; ++u) {
          const int i = tid + u * kEtGather;
          const int bl = i / p.fpc, j = i - bl * p.fpc;
          raw[u].xc = 0.f;
          raw[u].xk = 0;
          if (i < npair && m0 + bl < B && f0 + j < F)
            raw[u] = criteo_load_raw(s_fields[j], p.xcont, p.n_cont, p.xcat, p.n_cat, m0 + bl);
        }
  #pragma unroll
        for (int u = 0; u < NU; ++u) {
          const int i = tid + u * kEtGather;
          if (i < npair) {
            const int bl = i / p.fpc, j = i - bl * p.fpc;
            const int b = m0 + bl, f = f0 + j;
            int id = -1;
            if (b < B && f < F) {
              id = criteo_id_of(s_fields[j], s_bnd, raw[u], nullptr, p.status);
              p.rows_out[static_cast<size_t>(b) * F + f] = id;
            }
            ids_s[i] = id;
          }
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      et_stamp(p, 1);
  #pragma unroll
      for (int i = 0; i < 2 * kEtMaxFPC; ++i) {
        const int j = i >> 1, row = 16 * warp + 8 * (i & 1) + r8;
        rid[i] = j < p.fpc ? ids_s[row * p.fpc + j] : -1;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");     // the staging area becomes an A tile again
    }
    float4 v[2 * kEtMaxFPC];
#pragma unroll
    for (int i = 0; i < 2 * kEtMaxFPC; ++i)
      v[i] = rid[i] >= 0 ? ldg4(p.table + static_cast<size_t>(rid[i]) * p.ld + c * 4) : f4_zero();
    // first-order weights: requested now, summed after the A tiles are out (a sum here would wait
    // for the LAST load of the CTA before the first k-block could be handed to the tensor core)
    float w1v[2 * kEtMaxFPC];
#pragma unroll
    for (int i = 0; i < 2 * kEtMaxFPC; ++i)
      w1v[i] = (p.y1 != nullptr && c == 1 && rid[i] >= 0 && ((p.w1_fields >> (f0 + (i >> 1))) & 1ull))
                   ? __ldg(p.w1 + static_cast<size_t>(rid[i]) * p.ld1) : 0.f;
    et_stamp(p, 2);
#pragma unroll
    for (int kb = 0; kb < kEtMaxKB; ++kb) {
      if (kb < p.kbc) {                            // block-uniform
        const int st = kb & 1;
        if (kb >= 2) mbar_wait(&a_empty[st], ((kb >> 1) - 1) & 1);
        uint8_t* ah = a_ring + static_cast<size_t>(st) * 2 * kTcABytes;
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          const int i = 4 * kb + ii, j = i >> 1, half = i & 1;
          const int row = 16 * warp + 8 * half + r8;
          const uint32_t off = et_sw128(row, (j & 1) * 4 + c);
          const float4 x = v[i];
          *reinterpret_cast<float4*>(ah + off) = x;
          *reinterpret_cast<float4*>(ah + kTcABytes + off) = tcg_lo4(x);
          Sp[half] = f4_add(Sp[half], x);
          qp[half] += f4_dot(x, x);
        }
        fence_proxy_async_smem();
        mbar_arrive(&a_full[st]);
      }
    }
    et_stamp(p, 3);
#pragma unroll
    for (int i = 0; i < 2 * kEtMaxFPC; ++i) y1p[i & 1] += w1v[i];
    // E / E_lo for the backward's weight-gradient GEMM: written behind the A tiles, so that the
    // 20 MB of stores overlap the MMAs instead of delaying them
#pragma unroll
    for (int i = 0; i < 2 * kEtMaxFPC; ++i) {
      const int j = i >> 1, row = 16 * warp + 8 * (i & 1) + r8;
      if (j < p.fpc && rid[i] >= 0) {
        const size_t o = static_cast<size_t>(m0 + row) * F * kEtD + (f0 + j) * kEtD + c * 4;
        *reinterpret_cast<float4*>(p.E + o) = v[i];
        if (p.E_lo != nullptr) *reinterpret_cast<float4*>(p.E_lo + o) = tcg_lo4(v[i]);
      }
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      qp[half] += __shfl_xor_sync(0xffffffffu, qp[half], 8);
      qp[half] += __shfl_xor_sync(0xffffffffu, qp[half], 16);
    }
  } else if (warp == 8 && lane == 0) {
    // ------------------------------------------------------------------ TMA producer (W0)
    const int nbox = static_cast<int>(p.b_bytes / kTcgMnBox);
    for (int kb = 0; kb < p.kbc; ++kb) {
      mbar_expect_tx(&b_full[kb], 2 * p.b_bytes);
      const int kk = (kg * p.fpc + 2 * kb) * kEtD;
      uint8_t* sb = b_all + static_cast<size_t>(kb) * 2 * p.b_bytes;
      for (int j = 0; j < nbox; ++j) {
        tma_load_2d(sb + j * kTcgMnBox, &tmW, 32 * j, kk, &b_full[kb]);
        tma_load_2d(sb + p.b_bytes + j * kTcgMnBox, &tmWlo, 32 * j, kk, &b_full[kb]);
      }
    }
  } else if (warp == 9 && lane == 0) {
    // -------------------------------------------------------------------- MMA issuer
    uint32_t accum = 0;
    for (int kb = 0; kb < p.kbc; ++kb) {
      const uint32_t st = kb & 1, ph = (kb >> 1) & 1;
      mbar_wait(&b_full[kb], 0);
      mbar_wait(&a_full[st], ph);
      tc_fence_after();
      const uint32_t ah = smem_u32(a_ring + static_cast<size_t>(st) * 2 * kTcABytes);
      const uint32_t al = ah + kTcABytes;
      const uint32_t bh = smem_u32(b_all + static_cast<size_t>(kb) * 2 * p.b_bytes);
      const uint32_t bl = bh + p.b_bytes;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        tc_mma_tf32(tmem_base, p.desc_k | (((ah + k * 32) >> 4) & 0x3FFF),
                    p.desc_mn | (((bh + k * 1024) >> 4) & 0x3FFF), p.idesc, accum);
        accum = 1;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)   // (lo, hi)
        tc_mma_tf32(tmem_base, p.desc_k | (((al + k * 32) >> 4) & 0x3FFF),
                    p.desc_mn | (((bh + k * 1024) >> 4) & 0x3FFF), p.idesc, 1);
#pragma unroll
      for (int k = 0; k < 4; ++k)   // (hi, lo)
        tc_mma_tf32(tmem_base, p.desc_k | (((ah + k * 32) >> 4) & 0x3FFF),
                    p.desc_mn | (((bl + k * 1024) >> 4) & 0x3FFF), p.idesc, 1);
      tc_commit(&a_empty[st]);
    }
    tc_commit(t_full);
  }
  __syncwarp();

  // ------------------------------------------- every MMA has retired: the B region is free
  mbar_wait(t_full, 0);
  tc_fence_after();
  et_stamp(p, 4);
  if (warp < 8) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int row = 16 * warp + 8 * half + r8;
      *reinterpret_cast<float4*>(part + row * kEtPartPitch + c * 4) = Sp[half];
      if (c == 0) part[row * kEtPartPitch + 16] = qp[half];
      if (c == 1) part[row * kEtPartPitch + 17] = y1p[half];
    }
  }
  {
    // TMEM -> shared memory by all 12 warps: warp w reads the lanes of its quarter (w & 3), the
    // 8-column granules are dealt out over the three warps of a quarter
    const int quarter = warp & 3, third = warp >> 2;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    float* drow = dump + (quarter * 32 + lane) * PD;
    for (int c0 = third * 8; c0 < p.NT; c0 += 24) {
      float x[8];
      tc_ld<8>(taddr + c0, x);
      *reinterpret_cast<float4*>(drow + c0) = make_float4(x[0], x[1], x[2], x[3]);
      *reinterpret_cast<float4*>(drow + c0 + 4) = make_float4(x[4], x[5], x[6], x[7]);
    }
  }
  tc_fence_before();
  cluster_sync_all();                 // all four partial products and FM partials are visible
  et_stamp(p, 5);

  // ---------------------------------------- cluster reduction: this CTA takes rows 32r..32r+31
  const uint32_t rank = cluster_rank();
  const int N = p.N, nq = N >> 2;
  {
    const uint32_t dump_s = smem_u32(dump);
    uint32_t peer[kEtKS];
#pragma unroll
    for (int k = 0; k < kEtKS; ++k) peer[k] = map_to_rank(dump_s, static_cast<uint32_t>(k));
    // all remote loads of the thread's (<= 3) items first, then the epilogue arithmetic
    constexpr int NI = (32 * 32 + kEtThreads - 1) / kEtThreads;
    float4 acc[NI][kEtKS];
#pragma unroll
    for (int u = 0; u < NI; ++u) {
      const int e = tid + u * kEtThreads;
      if (e < 32 * nq) {
        const int rr = e / nq, cq = (e - rr * nq) * 4;
        const uint32_t off = static_cast<uint32_t>((static_cast<int>(rank) * 32 + rr) * PD + cq) * 4u;
#pragma unroll
        for (int k = 0; k < kEtKS; ++k) acc[u][k] = ld_cluster4(peer[k] + off);
      }
    }
#pragma unroll
    for (int u = 0; u < NI; ++u) {
      const int e = tid + u * kEtThreads;
      if (e < 32 * nq) {
        const int rr = e / nq, cq = (e - rr * nq) * 4;
        const int row = static_cast<int>(rank) * 32 + rr;
        float4 s = acc[u][0];
#pragma unroll
        for (int k = 1; k < kEtKS; ++k) s = f4_add(s, acc[u][k]);
        const float4 bb = ldg4(p.bias + cq);
        s = make_float4(fmaxf(s.x + bb.x, 0.f), fmaxf(s.y + bb.y, 0.f), fmaxf(s.z + bb.z, 0.f),
                        fmaxf(s.w + bb.w, 0.f));
        const int b = m0 + row;
        if (b < B) {
          *reinterpret_cast<float4*>(p.act0 + static_cast<size_t>(b) * N + cq) = s;
        } else {
          s = f4_zero();                 // rows past the batch do not count in the column sums
        }
        *reinterpret_cast<float4*>(tile + rr * PD + cq) = s;
      }
    }
    // FM terms of the same rows: S, y1, y2
    if (tid < 128) {
      const int rr = tid >> 2, cc = tid & 3;
      const int row = static_cast<int>(rank) * 32 + rr;
      const uint32_t part_s = smem_u32(part);
      float4 s = f4_zero();
      float q = 0.f, y = 0.f;
#pragma unroll
      for (int k = 0; k < kEtKS; ++k) {
        const uint32_t pa = map_to_rank(part_s, static_cast<uint32_t>(k)) +
                            static_cast<uint32_t>(row * kEtPartPitch) * 4u;
        s = f4_add(s, ld_cluster4(pa + cc * 16));
        if (cc == 0) {
          q += ld_cluster1(pa + 64);
          y += ld_cluster1(pa + 68);
        }
      }
      float t = f4_dot(s, s);
      t += __shfl_xor_sync(0xffffffffu, t, 1);
      t += __shfl_xor_sync(0xffffffffu, t, 2);
      const int b = m0 + row;
      if (b < B) {
        if (p.S != nullptr) *reinterpret_cast<float4*>(p.S + static_cast<size_t>(b) * kEtD + cc * 4) = s;
        if (cc == 0) {
          if (p.y2 != nullptr) p.y2[b] = 0.5f * (t - q);
          if (p.y1 != nullptr) p.y1[b] = y;
        }
      }
    }
  }
  __syncthreads();
  et_stamp(p, 6);
  if (p.stats_part != nullptr) {
    // column sums of this CTA's 32 rows, then one block per cluster through rank 0
    if (tid < 2 * 128) {
      const int n = tid & 127, sq = tid >> 7;
      float s = 0.f;
      if (n < N)
        for (int rr = 0; rr < 32; ++rr) {
          const float x = tile[rr * PD + n];
          s += sq ? x * x : x;
        }
      colp[sq * 128 + n] = s;
    }
    cluster_sync_all();
    if (rank == 0 && tid < 2 * 128) {
      const int n = tid & 127, sq = tid >> 7;
      if (n < N) {
        const uint32_t cs = smem_u32(colp + sq * 128 + n);
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < kEtKS; ++k) s += ld_cluster1(map_to_rank(cs, static_cast<uint32_t>(k)));
        p.stats_part[(static_cast<size_t>(blockIdx.x) * 2 + sq) * N + n] = s;
      }
    }
  }
  cluster_sync_all();                 // no CTA leaves while a peer may still read its shared memory
  et_stamp(p, 7);
  if (warp == 10) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128)
                 : "memory");
  }
}


// ======================================================================== backward
// Fused first-layer data gradient + embedding scatter-add:  dE = dpre0 . W0^T is formed per
// (128 samples, 10 fields) tile in TMEM and consumed from there - row gradient dE + dy2 * (S - E)
// (fm/fm.py:123-129), warp-aggregated vector REDs into the table's gradient accumulator, first-order
// gradient - so dE never exists in global memory and the GEMM -> scatter kernel boundary is gone
// (unfused: 11.7 us GEMM + 18.3 us scatter at B = 4096).  One CTA per (M tile, field group), no
// cross-CTA reduction:
//   TMA         dpre0 / dpre0_lo k-blocks (all resident) and the W0 / W0_lo rows of the CTA's
//               fields (K-major, 2-stage ring)
//   MMA         tcgen05.mma, 3xTF32; the last k-block only issues the k-steps that hold data
//   epilogue    8 warps, two per TMEM lane quarter, lane = sample: per field 16 accumulator
//               columns -> registers; fields with more than 32 rows: duplicates inside the warp are
//               summed into one lane (__match_any_sync), 4 vector REDs per distinct row; fields with
//               <= 32 rows (every sample of the warp hits a handful of rows): the 32 slots are summed
//               per row in a warp-private shared-memory tile first, one RED group per row that was hit
// Lanes that hold the same key, from `nbits` ballots (one per key bit) instead of match.any: the
// match instruction walks the warp's DISTINCT values one by one (~2000 cycles when all 32 differ,
// the common case for a large field), and an epilogue warp would pay that once per field.
// Lanes with valid == false match nobody.
__device__ __forceinline__ unsigned warp_peers(int key, bool valid, int nbits, int lane) {
  unsigned peers = __ballot_sync(0xffffffffu, valid);
  for (int bit = 0; bit < nbits; ++bit) {
    const bool one = (key >> bit) & 1;
    const unsigned bal = __ballot_sync(0xffffffffu, one);
    peers &= one ? bal : ~bal;
  }
  return valid ? peers : (1u << lane);
}

constexpr int kEbThreads = 256;
constexpr int kEbFPC = 10;                       // fields per CTA
constexpr int kEbNT = kEbFPC * kEtD;             // 160 accumulator columns
constexpr int kEbBStage = 2 * kEbNT * kTcKB * 4; // W0 hi | lo k-block: 2 x 20 KB
constexpr int kEbStage = 2 * kTcABytes + kEbBStage;   // one k-block of both operands: 72 KB
constexpr int kEbStages = 2;                     // 144 KB: leaves room for a second resident kernel
constexpr int kEbMaxKB = 4;                      // H0 <= 128
constexpr int kEbFPW = kEbFPC / 2;               // fields per epilogue warp

struct EbParams {
  const int* rows;
  const float* E;
  const float* S;
  const float* dy2;
  const float* dy1;
  float* dtable;
  float* dw1;
  long long ld_g, ld_w;
  unsigned long long w1_fields;
  unsigned long long tiny_fields;     // bit f: field f has <= 32 rows
  int off[CTR_MAX_FIELDS + 1];
  int B, F, H, nkb, last_ksteps;
  uint32_t idesc;
  uint64_t desc_k;
  unsigned long long* timing;
  int debug;          // profiling aid ("eb_debug"): 1 = no epilogue at all, 2 = epilogue without REDs,
                      // 4 = every field through the large-field path
};

// All 8 warps are epilogue warps (two per TMEM lane quarter, lane = sample); lane 0 of warp 0
// first issues every TMA load, lane 0 of warp 1 issues the MMAs.  Everything the epilogue needs
// that does not depend on the GEMM - row ids, S, dy, the E rows of the FM term - is requested
// before the roles split, so it is in registers when the accumulator is ready.
__global__ void __launch_bounds__(kEbThreads, 2)     // <= 128 registers: a second kernel fits beside it
tower_embed_bwd_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
                       const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmWlo,
                       const EbParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  // [ring: 2 x (A hi | A lo | B hi | B lo)] [barriers]; the epilogue's tiles reuse the ring
  uint8_t* ring = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + kEbStages * kEbStage);
  uint64_t* full = bars;              // [2]
  uint64_t* empty = bars + 2;         // [2]
  uint64_t* t_full = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * kTcBM;
  const int kg = blockIdx.y;
  const int f0 = kg * kEbFPC;
  const bool stamp = p.timing != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && tid == 64;
  if (stamp) p.timing[0] = gtime_ns();

  if (tid == 0) {
    for (int s = 0; s < kEbStages; ++s) {
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

  auto load_stage = [&](int kb) {
    const uint32_t st = kb % kEbStages;
    uint8_t* sa = ring + static_cast<size_t>(st) * kEbStage;
    mbar_expect_tx(&full[st], kEbStage);
    tma_load_2d(sa, &tmA, kb * kTcKB, m0, &full[st]);
    tma_load_2d(sa + kTcABytes, &tmAlo, kb * kTcKB, m0, &full[st]);
    tma_load_2d(sa + 2 * kTcABytes, &tmW, kb * kTcKB, f0 * kEtD, &full[st]);
    tma_load_2d(sa + 2 * kTcABytes + kEbBStage / 2, &tmWlo, kb * kTcKB, f0 * kEtD, &full[st]);
  };
  if (tid == 0) {
    // the first two k-blocks of both operands: in flight before anything else
    for (int kb = 0; kb < min(kEbStages, p.nkb); ++kb) load_stage(kb);
  }
  __syncwarp();

  // ---------------------------------------------- epilogue operands that do not need the GEMM
  const int quarter = warp & 3, sub = warp >> 2;             // two warps per lane quarter
  const int b = m0 + quarter * 32 + lane;
  const bool live = b < p.B;
  const bool fm = p.dy2 != nullptr && !(p.debug & 8);
  int rid[kEbFPW];
#pragma unroll
  for (int t = 0; t < kEbFPW; ++t) {
    const int f = f0 + sub + 2 * t;
    rid[t] = (live && f < p.F) ? __ldg(p.rows + static_cast<size_t>(b) * p.F + f) : -1;
  }
  // the E row of the FM term: one field ahead of the one being scattered
  float4 En[4];
  auto load_e = [&](int t) {
    const int f = f0 + sub + 2 * t;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      En[c] = (fm && live && f < p.F)
                  ? ldg4(p.E + (static_cast<size_t>(b) * p.F + f) * kEtD + 4 * c) : f4_zero();
  };
  load_e(0);
  float4 Sv[4];
  float d2 = 0.f, d1 = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) Sv[c] = (fm && live) ? ldg4(p.S + static_cast<size_t>(b) * kEtD + 4 * c) : f4_zero();
  if (live) {
    if (fm) d2 = __ldg(p.dy2 + b);
    if (p.dy1 != nullptr) d1 = __ldg(p.dy1 + b);
  }

  if (tid == 0) {
    // ------------------------------------------------ TMA producer: the remaining k-blocks
    for (int kb = kEbStages; kb < p.nkb; ++kb) {
      const uint32_t st = kb % kEbStages, ph = (kb / kEbStages) & 1;
      mbar_wait(&empty[st], ph ^ 1);
      load_stage(kb);
    }
  } else if (tid == 32) {
    // -------------------------------------------------------------------- MMA issuer
    uint32_t accum = 0;
    for (int kb = 0; kb < p.nkb; ++kb) {
      const uint32_t st = kb % kEbStages, ph = (kb / kEbStages) & 1;
      mbar_wait(&full[st], ph);
      tc_fence_after();
      const uint32_t ah = smem_u32(ring + static_cast<size_t>(st) * kEbStage);
      const uint32_t al = ah + kTcABytes;
      const uint32_t bh = ah + 2 * kTcABytes;
      const uint32_t bl = bh + kEbBStage / 2;
      const int ks = kb == p.nkb - 1 ? p.last_ksteps : 4;
      for (int k = 0; k < ks; ++k) {
        tc_mma_tf32(tmem_base, p.desc_k | (((ah + k * 32) >> 4) & 0x3FFF),
                    p.desc_k | (((bh + k * 32) >> 4) & 0x3FFF), p.idesc, accum);
        accum = 1;
      }
      for (int k = 0; k < ks; ++k)   // (lo, hi)
        tc_mma_tf32(tmem_base, p.desc_k | (((al + k * 32) >> 4) & 0x3FFF),
                    p.desc_k | (((bh + k * 32) >> 4) & 0x3FFF), p.idesc, 1);
      for (int k = 0; k < ks; ++k)   // (hi, lo)
        tc_mma_tf32(tmem_base, p.desc_k | (((ah + k * 32) >> 4) & 0x3FFF),
                    p.desc_k | (((bl + k * 32) >> 4) & 0x3FFF), p.idesc, 1);
      tc_commit(&empty[st]);
    }
    tc_commit(t_full);
  }
  __syncwarp();

  // ----------------------------------------------------------------------- epilogue
  if (stamp) p.timing[1] = gtime_ns();
  mbar_wait(t_full, 0);
  tc_fence_after();
  if (stamp) p.timing[2] = gtime_ns();
  // warp-private tile: [32 rows][16 + 4] floats, in the retired ring
  float* tileb = reinterpret_cast<float*>(ring) + warp * (32 * 20 + 64);
  float* tw1 = tileb + 32 * 20;
  int* tcnt = reinterpret_cast<int*>(tw1 + 32);
  const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll
  for (int t = 0; t < kEbFPW; ++t) {
    const int j = sub + 2 * t, f = f0 + j;
    if (f < p.F && !(p.debug & 1)) {                           // warp-uniform
      float4 Ec[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) Ec[c] = En[c];
      if (t + 1 < kEbFPW) load_e(t + 1);
      float x[16];
      tc_ld<16>(taddr + j * kEtD, x);
      float4 g[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        g[c] = make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
        const float4 ev = Ec[c];
        g[c].x = fmaf(d2, Sv[c].x - ev.x, g[c].x);
        g[c].y = fmaf(d2, Sv[c].y - ev.y, g[c].y);
        g[c].z = fmaf(d2, Sv[c].z - ev.z, g[c].z);
        g[c].w = fmaf(d2, Sv[c].w - ev.w, g[c].w);
      }
      const bool has_w1 = p.dw1 != nullptr && p.dy1 != nullptr && ((p.w1_fields >> f) & 1ull);
      const int r = rid[t];
      float gw = has_w1 ? d1 : 0.f;
      if (((p.tiny_fields >> f) & 1ull) && !(p.debug & 4)) {
        // every row of the field fits 32 tile rows.  The warp's 32 slots go to shared memory once;
        // then 4 lanes per row (8 rows per pass) add up the slots of that row - found through the
        // row's peer mask - and RED the row's 64 bytes straight from registers.  No read-modify-write
        // chain on shared memory: a row that all 32 samples hit costs 32 independent loads, not 32
        // dependent turns.
        const int lid = r >= 0 ? r - p.off[f] : -1;
        const unsigned peers = warp_peers(lid, lid >= 0, 5, lane);
#pragma unroll
        for (int c = 0; c < 4; ++c) *reinterpret_cast<float4*>(tileb + lane * 20 + 4 * c) = g[c];
        tileb[lane * 20 + 16] = gw;
        unsigned* pm = reinterpret_cast<unsigned*>(tcnt);
        pm[lane] = 0u;
        __syncwarp();
        if (lid >= 0) pm[lid] = peers;            // every peer writes the same mask
        __syncwarp();
        const int nrow = (p.debug & 2) ? 0 : p.off[f + 1] - p.off[f];
        for (int k = 0; 8 * k < nrow; ++k) {
          const int lr = 8 * k + (lane >> 2), cc = lane & 3;
          unsigned m = lr < nrow ? pm[lr] : 0u;
          if (m != 0u) {
            float4 acc = f4_zero();
            float aw = 0.f;
            while (m != 0u) {
              const int i = __ffs(m) - 1;
              m &= m - 1u;
              acc = f4_add(acc, *reinterpret_cast<const float4*>(tileb + i * 20 + 4 * cc));
              if (cc == 0) aw += tileb[i * 20 + 16];
            }
            red_add_v4(p.dtable + static_cast<size_t>(p.off[f] + lr) * p.ld_g + 4 * cc, acc);
            if (has_w1 && cc == 0) red_add_f32(p.dw1 + static_cast<size_t>(p.off[f] + lr) * p.ld_w, aw);
          }
        }
        __syncwarp();
      } else {
        // large field: duplicates inside the warp are rare - sum them into the first lane of the group
        const int nbits = 32 - __clz(max(1, p.off[f + 1] - p.off[f] - 1));
        const unsigned peers = (p.debug & 16) ? (1u << lane) : warp_peers(r - p.off[f], r >= 0, nbits, lane);
        const bool leader = (peers & ((1u << lane) - 1u)) == 0u;
        if (!(p.debug & 16) && __any_sync(0xffffffffu, __popc(peers) > 1)) {
          unsigned rest = peers & ~(1u << lane);
          while (__ballot_sync(0xffffffffu, rest != 0u) != 0u) {
            const int src = rest != 0u ? __ffs(rest) - 1 : lane;
            const bool ok = rest != 0u && leader;
            rest &= rest - 1u;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float4 o = f4_shfl(g[c], src);
              if (ok) g[c] = f4_add(g[c], o);
            }
            const float ow = __shfl_sync(0xffffffffu, gw, src);
            if (ok) gw += ow;
          }
        }
        // through the warp's tile, so that 4 lanes RED one row's 64 bytes in ONE instruction (one
        // L2 transaction per row instead of four half-sector ones)
#pragma unroll
        for (int c = 0; c < 4; ++c) *reinterpret_cast<float4*>(tileb + lane * 20 + 4 * c) = g[c];
        __syncwarp();
        const int r_eff = (r >= 0 && leader && !(p.debug & 2)) ? r : -1;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int src = 8 * k + (lane >> 2), cc = lane & 3;
          const int rk = __shfl_sync(0xffffffffu, r_eff, src);
          const float gwk = __shfl_sync(0xffffffffu, gw, src);
          if (rk >= 0) {
            red_add_v4(p.dtable + static_cast<size_t>(rk) * p.ld_g + 4 * cc,
                       *reinterpret_cast<const float4*>(tileb + src * 20 + 4 * cc));
            if (has_w1 && cc == 0) red_add_f32(p.dw1 + static_cast<size_t>(rk) * p.ld_w, gwk);
          }
        }
        __syncwarp();
      }
    }
  }
  if (stamp) p.timing[3] = gtime_ns();
  tc_fence_before();
  __syncthreads();
  if (stamp) p.timing[4] = gtime_ns();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256)
                 : "memory");
  }
}

}  // namespace ctr

using namespace ctr;

static unsigned long long* g_et_timing = nullptr;

extern "C" {

int ctr_embed_tower_timing(uint64_t* timing_dev) {
  g_et_timing = reinterpret_cast<unsigned long long*>(timing_dev);
  return CTR_OK;
}

int ctr_embed_tower_fwd(const float* table, const float* w1, const float* xcont, int n_cont,
                        const int64_t* xcat, int n_cat, const ctr_field_desc* fields_dev,
                        const float* boundaries_dev, int n_boundaries, const int32_t* rows_in,
                        int32_t* rows_out, int32_t* status, int B,
                        int F, int D, uint64_t w1_fields, float* E, float* E_lo, float* S, float* y1,
                        float* y2, int64_t row_stride, int64_t w1_stride, const float* W0,
                        const float* W0_lo, const float* b0, int N, float* act0, float* stats_part,
                        float* zero_buf, int64_t zero_n, ctr_stream_t stream) {
  const char* fn = "ctr_embed_tower_fwd";
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(table && (rows_in || (fields_dev && rows_out)) && E && W0 && W0_lo && b0 && act0, fn,
              "null pointer");
  CTR_REQUIRE(D == kEtD, fn, "D must be 16");
  CTR_REQUIRE(B >= 0 && F > 0 && F <= kEtKS * kEtMaxFPC, fn, "need 0 < F <= 40");
  CTR_REQUIRE(N >= 16 && N <= 128 && (N & 3) == 0, fn, "need 16 <= N <= 128, N % 4 == 0");
  CTR_REQUIRE(rows_in || n_cont == 0 || (xcont && boundaries_dev), fn, "null xcont/boundaries");
  CTR_REQUIRE(rows_in || n_cat == 0 || xcat, fn, "null xcat");
  CTR_REQUIRE(n_boundaries >= 0 && n_boundaries <= 512, fn, "at most 512 bucket boundaries in total");
  CTR_REQUIRE(y1 == nullptr || w1 != nullptr, fn, "y1 requested without w1");
  CTR_REQUIRE(aligned16(table) && aligned16(E) && aligned16(E_lo) && aligned16(S) && aligned16(W0) &&
                  aligned16(W0_lo) && aligned16(b0) && aligned16(act0),
              fn, "pointers must be 16-byte aligned");
  CTR_REQUIRE(zero_buf == nullptr || (aligned16(zero_buf) && zero_n >= 0 && (zero_n & 3) == 0), fn,
              "zero_buf must be 16-byte aligned, zero_n a multiple of 4");
  if (row_stride <= 0) row_stride = D;
  if (w1_stride <= 0) w1_stride = 1;
  CTR_REQUIRE(row_stride >= D && (row_stride & 3) == 0, fn,
              "row_stride must be >= D and a multiple of 4 floats");
  if (B == 0) return CTR_OK;
  EtParams p{};
  p.table = table; p.w1 = w1; p.ld = row_stride; p.ld1 = w1_stride; p.w1_fields = w1_fields;
  p.xcont = xcont; p.xcat = reinterpret_cast<const long long*>(xcat); p.fields = fields_dev;
  p.bnd = boundaries_dev; p.n_cont = n_cont; p.n_cat = n_cat; p.n_bnd = n_boundaries; p.rows_in = rows_in; p.rows_out = rows_out; p.status = status;
  p.E = E; p.E_lo = E_lo; p.S = S; p.y1 = y1; p.y2 = y2; p.bias = b0; p.act0 = act0;
  p.stats_part = stats_part; p.zero_buf = zero_n > 0 ? zero_buf : nullptr; p.zero_n4 = zero_n >> 2;
  p.B = B; p.F = F; p.N = N; p.NT = (N + 15) / 16 * 16;
  p.fpc = 2 * ((F + 2 * kEtKS - 1) / (2 * kEtKS));
  p.kbc = p.fpc / 2;
  p.b_bytes = static_cast<uint32_t>((p.NT + 31) / 32) * kTcgMnBox;
  p.idesc = cin_idesc(p.NT) | (1u << 16);          // B is MN-major
  p.desc_k = cin_desc_hi();
  p.desc_mn = tcg_desc_mn();
  p.timing = g_et_timing;
  const int K = F * D;
  CUtensorMap tW, tWlo;
  int r = make_map(&tW, W0, K, N, N, 32, true);
  if (r != CTR_OK) return r;
  r = make_map(&tWlo, W0_lo, K, N, N, 32, true);
  if (r != CTR_OK) return r;
  const size_t tiles = static_cast<size_t>(p.kbc) * 2 * p.b_bytes + 4 * static_cast<size_t>(kTcABytes);
  // the region the epilogue reuses: partial product at 0, FM partials / column partials / act0 tile
  // from 72 KB on
  const size_t reuse = 72 * 1024 + (128 * kEtPartPitch + 256 + 32 * (p.NT + 4)) * sizeof(float);
  static_assert(128 * (128 + 4) * sizeof(float) <= 72 * 1024, "partial product must fit below 72 KB");
  p.bar_off = static_cast<uint32_t>((std::max(tiles, reuse) + 1023) & ~static_cast<size_t>(1023));
  const size_t smem = p.bar_off + 256 + 1024;
  CTR_REQUIRE(smem <= 227 * 1024, fn, "internal: shared memory budget exceeded");
  static bool optin = false;
  if (!optin) {
    cudaFuncSetAttribute(embed_tower_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         227 * 1024);
    optin = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((B + kTcBM - 1) / kTcBM, kEtKS, 1);
  cfg.blockDim = dim3(kEtThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = kEtKS;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return check_cuda(cudaLaunchKernelEx(&cfg, embed_tower_fwd_kernel, tW, tWlo, p), fn);
}

int ctr_tower_embed_bwd(const float* dpre0, const float* dpre0_lo, const float* W0, const float* W0_lo,
                        int N, const int32_t* rows, const float* E, const float* S, const float* dy2,
                        const float* dy1, uint64_t w1_fields, const int64_t* row_offsets_host, int B,
                        int F, int D, float* dtable, float* dw1, int64_t row_stride, int64_t w1_stride,
                        ctr_stream_t stream) {
  const char* fn = "ctr_tower_embed_bwd";
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(dpre0 && dpre0_lo && W0 && W0_lo && rows && dtable && row_offsets_host, fn, "null pointer");
  CTR_REQUIRE(D == kEtD, fn, "D must be 16");
  CTR_REQUIRE(B >= 0 && F > 0 && F <= CTR_MAX_FIELDS, fn, "need 0 < F <= 64");
  CTR_REQUIRE(N >= 16 && N <= 32 * kEbMaxKB && (N & 3) == 0, fn, "need 16 <= N <= 128, N % 4 == 0");
  CTR_REQUIRE(!dy2 || (S && E), fn, "dy2 needs S and E");
  CTR_REQUIRE(aligned16(dpre0) && aligned16(dpre0_lo) && aligned16(W0) && aligned16(W0_lo) &&
                  aligned16(E) && aligned16(S) && aligned16(dtable),
              fn, "pointers must be 16-byte aligned");
  CTR_REQUIRE(row_offsets_host[F] < (1LL << 31), fn, "table too large for int32 rows");
  if (row_stride <= 0) row_stride = D;
  if (w1_stride <= 0) w1_stride = 1;
  CTR_REQUIRE(row_stride >= D && (row_stride & 3) == 0, fn,
              "row_stride must be >= D and a multiple of 4 floats");
  if (B == 0) return CTR_OK;
  EbParams p{};
  p.rows = rows; p.E = E; p.S = S; p.dy2 = dy2; p.dy1 = dy1; p.dtable = dtable; p.dw1 = dw1;
  p.ld_g = row_stride; p.ld_w = w1_stride; p.w1_fields = w1_fields; p.B = B; p.F = F; p.H = N;
  for (int f = 0; f <= F; ++f) p.off[f] = static_cast<int>(row_offsets_host[f]);
  for (int f = 0; f < F; ++f)
    if (p.off[f + 1] - p.off[f] <= 32) p.tiny_fields |= 1ull << f;
  p.nkb = (N + kTcKB - 1) / kTcKB;
  p.last_ksteps = (N - (p.nkb - 1) * kTcKB + 7) / 8;
  p.idesc = cin_idesc(kEbNT);
  p.desc_k = cin_desc_hi();
  p.timing = g_et_timing;
  p.debug = option_get("eb_debug", 0);
  CUtensorMap tA, tAlo, tW, tWlo;
  int r = make_map(&tA, dpre0, B, N, N, kTcBM);
  if (r == CTR_OK) r = make_map(&tAlo, dpre0_lo, B, N, N, kTcBM);
  if (r == CTR_OK) r = make_map(&tW, W0, static_cast<long long>(F) * D, N, N, kEbNT);
  if (r == CTR_OK) r = make_map(&tWlo, W0_lo, static_cast<long long>(F) * D, N, N, kEbNT);
  if (r != CTR_OK) return r;
  const size_t smem = static_cast<size_t>(kEbStages) * kEbStage + 256 + 1024;
  static bool optin = false;
  if (!optin) {
    cudaFuncSetAttribute(tower_embed_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         227 * 1024);
    optin = true;
  }
  dim3 grid((B + kTcBM - 1) / kTcBM, (F + kEbFPC - 1) / kEbFPC);
  tower_embed_bwd_kernel<<<grid, kEbThreads, smem, static_cast<cudaStream_t>(stream)>>>(tA, tAlo, tW, tWlo,
                                                                                         p);
  return check_cuda(cudaGetLastError(), fn);
}

}  // extern "C"
