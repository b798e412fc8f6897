// Backward scatter-add and the row optimiser in ONE pass over the touched rows.
//
// The unfused step costs every touched row three trips: RED the gradient into the record's
// accumulator (embed_bwd), then fetch the whole record again to apply Adam and clear the
// accumulator (adam_rows).  Here the slot that completes a row's gradient applies the update on the
// spot, while the record is still in L2:
//
//   ctr_count_rows       cnt[row] += 1 per lookup of the batch (before the backward; runs beside
//                        the tower's backward GEMMs, off the critical path)
//   ctr_embed_bwd_adam   per (sample, field) slot: RED g[row] += dE + dy2*S, c[row] += dy2,
//                        g1[row] += dy1; fence; old = atomicSub(cnt[row], n); the slot that takes the
//                        count to zero is the row's LAST ARRIVER: it reads g, c, theta, m, v back
//                        (L2 hits), forms the FM gradient g - c*theta (so E is never re-read:
//                        sum_slots dy2*(S - E) = sum dy2*S - (sum dy2)*theta), applies TF-Adam, and
//                        clears g / c / g1.  Exactly one update per distinct row, no claim pass.
//
// Slots of one warp instruction that hit the same row are summed in registers first
// (__match_any_sync; the leader commits for the group and subtracts the multiplicity), fields with
// <= 32 rows are summed per row in a warp-private shared-memory tile over a whole sample chunk
// (_c5 puts every sample in one row) - the "warp-aggregated atomic scatter-add" of
// BASELINE north_star.  Works on the row-record layout only (include/ctr_b200.h, "Row strides"):
//   record[r] = { theta[D] | m[D] | v[D] | g[D] | theta1 m1 v1 g1 | claim cnt c pad }.
#include "row_commit.cuh"

namespace ctr {

struct BwdAdamParams {
  const int* rows;
  const float* dE;
  const float* S;
  const float* dy2;
  const float* dy1;
  float* rec;
  long long ld;
  unsigned long long w1_fields;
  const float* state;
  float lr_t;
  AdamK k;
  int B, F;
  // unfused scatter (ctr_embed_bwd): plain accumulation targets and the -dy2*E term's source
  const float* E;           // nullable: re-gather the row from `table` instead
  const float* table;
  float* dtable;
  float* dw1;
  long long ld_t, ld_g, ld_w;
  int n_fields;             // fields of this launch (tiny or big), listed in `order`
  int chunk, nchunks;       // samples per task
  int off[CTR_MAX_FIELDS + 1];
  unsigned char order[CTR_MAX_FIELDS];
};

constexpr int kTinyRowsA = 32;

// Gradient of one (sample, field) slot.  FUSED: dE + dy2*S, the -dy2*E term is applied per row by
// the optimiser through c = sum dy2 (see the header).  Unfused: the whole dE + dy2*(S - E)
// (fm/fm.py:123-129), E read back or re-gathered from the table.
template <int D, bool FUSED>
__device__ __forceinline__ void slot_grad(const BwdAdamParams& p, int b, int f, int rid, int q,
                                          bool has_w1, float4& g, float& gw, float& gc) {
  const size_t eo = (static_cast<size_t>(b) * p.F + f) * D + q * 4;
  g = p.dE != nullptr ? ld4_stream(p.dE + eo) : f4_zero();
  gc = 0.f;
  if (p.dy2 != nullptr) {
    float4 sv = ldg4(p.S + static_cast<size_t>(b) * D + q * 4);
    gc = __ldg(p.dy2 + b);
    if (!FUSED) {
      const float4 e = p.E != nullptr ? ldg4(p.E + eo)
                                      : ldg4(p.table + static_cast<size_t>(rid) * p.ld_t + q * 4);
      sv = make_float4(sv.x - e.x, sv.y - e.y, sv.z - e.z, sv.w - e.w);
    }
    g.x = fmaf(gc, sv.x, g.x); g.y = fmaf(gc, sv.y, g.y);
    g.z = fmaf(gc, sv.z, g.z); g.w = fmaf(gc, sv.w, g.w);
  }
  gw = has_w1 ? __ldg(p.dy1 + b) : 0.f;
}

// Fields with more than 32 rows: warp task = (field, chunk of samples); every warp instruction
// covers RPW rows x UNR, the row's theta / m / v are requested together with the slot's gradient.
template <int D>
__global__ void __launch_bounds__(128, 5) embed_bwd_adam_big_kernel(const BwdAdamParams p) {
  constexpr int LPR = D / 4;
  constexpr int RPW = 32 / LPR;
  constexpr int UNR = 2;
  const int lane = threadIdx.x & 31;
  const int r = lane / LPR;
  const int q = lane % LPR;
  const int wpb = blockDim.x >> 5;
  const int ntask = p.n_fields * p.nchunks;
  const float lr_t = p.state != nullptr ? p.state[1] : p.lr_t;
  const bool has_c = p.dy2 != nullptr;
  for (int task = blockIdx.x * wpb + (threadIdx.x >> 5); task < ntask; task += gridDim.x * wpb) {
    const int f = p.order[task % p.n_fields];
    const int b_begin = (task / p.n_fields) * p.chunk;
    const int b_end = min(p.B, b_begin + p.chunk);
    const bool has_w1 = p.dy1 != nullptr && ((p.w1_fields >> f) & 1ull);
    for (int bb = b_begin; bb < b_end; bb += RPW * UNR) {
      int rid[UNR], n[UNR];
      float4 g[UNR];
      float gw[UNR], gc[UNR];
      float* row[UNR];
      RowPre pre[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int b = bb + u * RPW + r;
        rid[u] = b < b_end ? __ldg(p.rows + static_cast<size_t>(b) * p.F + f) : -1;
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        g[u] = f4_zero();
        gw[u] = gc[u] = 0.f;
        if (rid[u] >= 0) {
          row_preload<D>(p.rec + static_cast<size_t>(rid[u]) * p.ld, q, pre[u]);
          slot_grad<D, true>(p, bb + u * RPW + r, f, rid[u], q, has_w1, g[u], gw[u], gc[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        bool leader;
        n[u] = merge_duplicates<LPR>(rid[u] >= 0 ? rid[u] : -1 - r, g[u], gw[u], gc[u], leader, lane, q);
        row[u] = (rid[u] >= 0 && leader) ? p.rec + static_cast<size_t>(rid[u]) * p.ld : nullptr;
      }
      commit_rows<D, UNR, true>(row, g, gw, gc, n, pre, has_w1, has_c, p.k, lr_t, lane, q);
    }
  }
}

// The unfused scatter-add for fields with more than 32 rows (ctr_embed_bwd): the same short warp
// tasks, gradients summed over the duplicates of a warp instruction (AGG), one vector RED per
// distinct row into the gradient accumulator.
template <int D, bool AGG>
__device__ __forceinline__ void bwd_big_body(const BwdAdamParams& p, int cta, int ncta) {
  constexpr int LPR = D / 4;
  constexpr int RPW = 32 / LPR;
  constexpr int UNR = 4;
  const int lane = threadIdx.x & 31;
  const int r = lane / LPR;
  const int q = lane % LPR;
  const int wpb = blockDim.x >> 5;
  const int ntask = p.n_fields * p.nchunks;
  for (int task = cta * wpb + (threadIdx.x >> 5); task < ntask; task += ncta * wpb) {
    const int f = p.order[task % p.n_fields];
    const int b_begin = (task / p.n_fields) * p.chunk;
    const int b_end = min(p.B, b_begin + p.chunk);
    const bool has_w1 = p.dw1 != nullptr && p.dy1 != nullptr && ((p.w1_fields >> f) & 1ull);
    for (int bb = b_begin; bb < b_end; bb += RPW * UNR) {
      int rid[UNR];
      float4 g[UNR];
      float gw[UNR], gc[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int b = bb + u * RPW + r;
        rid[u] = b < b_end ? __ldg(p.rows + static_cast<size_t>(b) * p.F + f) : -1;
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        g[u] = f4_zero();
        gw[u] = gc[u] = 0.f;
        if (rid[u] >= 0) slot_grad<D, false>(p, bb + u * RPW + r, f, rid[u], q, has_w1, g[u], gw[u], gc[u]);
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        bool leader = true;
        if (AGG) merge_duplicates<LPR>(rid[u] >= 0 ? rid[u] : -1 - r, g[u], gw[u], gc[u], leader, lane, q);
        if (rid[u] >= 0 && leader) {        // negative ids (sharded overflow slots) are skipped
          red_add_v4(p.dtable + static_cast<size_t>(rid[u]) * p.ld_g + q * 4, g[u]);
          if (has_w1 && q == 0) red_add_f32(p.dw1 + static_cast<size_t>(rid[u]) * p.ld_w, gw[u]);
        }
      }
    }
  }
}

// Fields with <= 32 rows (13 bucketised numerics, the small hashed fields; SURVEY H3: one of them
// puts every sample in one row): CTA task = (field, chunk of samples).  Every warp sums its share
// of the chunk per row in a private shared-memory tile (a slot adds its float4 with a plain
// read-modify-write; slots of one warp instruction that hit the same row take turns by their rank
// in a __match_any_sync group), the CTA adds the eight tiles up, and each row that was hit is
// committed once per CTA, with its multiplicity.
template <int D>
struct TinySmem {
  float acc[8][kTinyRowsA * (D + 4)];
  float w[8][kTinyRowsA], c[8][kTinyRowsA];
  int n[8][kTinyRowsA];
};

template <int D, bool FUSED>
__device__ __forceinline__ void bwd_tiny_body(const BwdAdamParams& p, TinySmem<D>& sm, int cta, int ncta) {
  constexpr int LPR = D / 4;
  constexpr int RPW = 32 / LPR;
  constexpr int TU = 4;
  constexpr int PT = D + 4;           // tile pitch (floats): conflict-free float4 rows
  auto& s_acc = sm.acc;
  auto& s_w = sm.w;
  auto& s_c = sm.c;
  auto& s_n = sm.n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = lane / LPR;
  const int q = lane % LPR;
  const int wpb = blockDim.x >> 5;
  const int ntask = p.n_fields * p.nchunks;
  const float lr_t = p.state != nullptr ? p.state[1] : p.lr_t;
  const bool has_c = p.dy2 != nullptr;
  float* sacc = s_acc[warp];
  for (int task = cta; task < ntask; task += ncta) {
    const int f = p.order[task % p.n_fields];
    const int b_begin = (task / p.n_fields) * p.chunk;
    const int b_end = min(p.B, b_begin + p.chunk);
    const int off = p.off[f];
    const int nrow = p.off[f + 1] - off;
    const bool has_w1 = p.dy1 != nullptr && (FUSED || p.dw1 != nullptr) && ((p.w1_fields >> f) & 1ull);
    for (int i = lane; i < kTinyRowsA * PT / 4; i += 32) reinterpret_cast<float4*>(sacc)[i] = f4_zero();
    s_w[warp][lane] = 0.f;
    s_c[warp][lane] = 0.f;
    s_n[warp][lane] = 0;
    __syncwarp();
    for (int bb = b_begin + warp * RPW * TU; bb < b_end; bb += wpb * RPW * TU) {
      int lid[TU];
      float4 g[TU];
      float gw[TU], gc[TU];
#pragma unroll
      for (int u = 0; u < TU; ++u) {
        const int b = bb + u * RPW + r;
        lid[u] = b < b_end ? __ldg(p.rows + static_cast<size_t>(b) * p.F + f) - off : -1;
      }
#pragma unroll
      for (int u = 0; u < TU; ++u) {
        g[u] = f4_zero();
        gw[u] = gc[u] = 0.f;
        if (lid[u] >= 0)
          slot_grad<D, FUSED>(p, bb + u * RPW + r, f, lid[u] + off, q, has_w1, g[u], gw[u], gc[u]);
      }
#pragma unroll
      for (int u = 0; u < TU; ++u) {
        const bool live = lid[u] >= 0 && lid[u] < kTinyRowsA;
        const unsigned peers = __match_any_sync(0xffffffffu, live ? lid[u] : -1 - r);
        const int rank = live ? __popc(peers & ((1u << (lane - q)) - 1u)) / LPR : 0;
        const int rounds = __reduce_max_sync(0xffffffffu, rank) + 1;
        for (int k = 0; k < rounds; ++k) {
          if (live && rank == k) {
            float4* dst = reinterpret_cast<float4*>(&sacc[lid[u] * PT + q * 4]);
            *dst = f4_add(*dst, g[u]);
            if (q == 0) {
              s_w[warp][lid[u]] += gw[u];
              s_c[warp][lid[u]] += gc[u];
              s_n[warp][lid[u]] += 1;
            }
          }
          __syncwarp();
        }
      }
    }
    __syncthreads();
    // warp w adds up rows w*RPW .. w*RPW+RPW-1 of the eight tiles: one group of LPR lanes per row
    if (warp * RPW < nrow) {
      const int lr = warp * RPW + r;
      float* row[1];
      float4 a[1];
      float aw[1], ac[1];
      int n[1];
      RowPre pre[1];
      a[0] = f4_zero();
      aw[0] = ac[0] = 0.f;
      n[0] = 0;
      if (lr < nrow) {
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          a[0] = f4_add(a[0], *reinterpret_cast<const float4*>(&s_acc[w][lr * PT + q * 4]));
          aw[0] += s_w[w][lr];
          ac[0] += s_c[w][lr];
          n[0] += s_n[w][lr];
        }
      }
      if (FUSED) {
        row[0] = n[0] > 0 ? p.rec + static_cast<size_t>(off + lr) * p.ld : nullptr;
        commit_rows<D, 1, false>(row, a, aw, ac, n, pre, has_w1, has_c, p.k, lr_t, lane, q);
      } else if (n[0] > 0) {
        red_add_v4(p.dtable + static_cast<size_t>(off + lr) * p.ld_g + q * 4, a[0]);
        if (has_w1 && q == 0) red_add_f32(p.dw1 + static_cast<size_t>(off + lr) * p.ld_w, aw[0]);
      }
    }
    __syncthreads();
  }
}

template <int D, bool FUSED>
__global__ void __launch_bounds__(256) embed_bwd_tiny_kernel(const BwdAdamParams p) {
  __shared__ __align__(16) TinySmem<D> sm;
  bwd_tiny_body<D, FUSED>(p, sm, blockIdx.x, gridDim.x);
}

// The unfused scatter-add in ONE launch: the first `tiny_ctas` CTAs take the <= 32-row fields (the
// longer tasks, scheduled first), the others the large fields.
struct BwdPair {
  BwdAdamParams tiny, big;
  int tiny_ctas;
};
template <int D, bool AGG>
__global__ void __launch_bounds__(256) embed_bwd_kernel(const __grid_constant__ BwdPair pp) {
  __shared__ __align__(16) TinySmem<D> sm;
  if (static_cast<int>(blockIdx.x) < pp.tiny_ctas)
    bwd_tiny_body<D, false>(pp.tiny, sm, blockIdx.x, pp.tiny_ctas);
  else
    bwd_big_body<D, AGG>(pp.big, blockIdx.x - pp.tiny_ctas, gridDim.x - pp.tiny_ctas);
}

// cnt[row] += lookups of the row in rows[n]; duplicates inside a warp instruction are combined.
// Every lookup also pulls the row's whole record into L2 (prefetch), so that the scatter + optimiser
// pass that follows - its REDs, its count round trip and its read-back - runs at L2 latency.
__global__ void __launch_bounds__(256)
count_rows_kernel(const int* __restrict__ rows, long long n, float* __restrict__ rec, long long ld,
                  int cnt_off) {
  const int lane = threadIdx.x & 31;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long n_pad = (n + 31) / 32 * 32;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n_pad;
       i += stride) {
    const int rid = i < n ? __ldg(rows + i) : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, rid >= 0 ? rid : -1 - lane);
    if (rid >= 0 && (peers & ((1u << lane) - 1u)) == 0u) {
      float* row = rec + static_cast<size_t>(rid) * ld;
      atomicAdd(reinterpret_cast<int*>(row + cnt_off), __popc(peers));
      const char* b0 = reinterpret_cast<const char*>(row);
      const char* b1 = reinterpret_cast<const char*>(row + cnt_off);       // last sector of the record
      for (const char* l = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(b0) & ~uintptr_t(127));
           l < b1; l += 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(l));
    }
  }
}

}  // namespace ctr

using namespace ctr;

// Split the fields into the two task spaces and launch both kernels side by side (the tiny-field
// kernel on the library's forked stream).  fused: scatter + optimiser; else plain scatter-add.
static int launch_bwd(BwdAdamParams& p, int D, bool fused, bool aggregate, cudaStream_t st, const char* fn) {
  const int B = p.B, F = p.F;
  const int rpw = 32 / (D / 4);
  BwdAdamParams pt = p;
  pt.n_fields = 0;
  for (int f = 0; f < F; ++f)
    if (p.off[f + 1] - p.off[f] <= kTinyRowsA) pt.order[pt.n_fields++] = static_cast<unsigned char>(f);
  BwdAdamParams pb = p;
  pb.n_fields = 0;
  for (int f = 0; f < F; ++f)
    if (p.off[f + 1] - p.off[f] > kTinyRowsA) pb.order[pb.n_fields++] = static_cast<unsigned char>(f);
  if (pt.n_fields > 0) {
    int chunk = 8 * rpw * 4;      // one load round per warp
    while (chunk < 4096 &&
           static_cast<long long>(pt.n_fields) * ((B + chunk - 1) / chunk) > sm_count() * 3LL)
      chunk <<= 1;
    pt.chunk = chunk;
    pt.nchunks = (B + chunk - 1) / chunk;
  }
  if (pb.n_fields > 0) {
    // short tasks (one load round each) so that the grid balances
    int chunk = rpw * (fused ? 2 : 4);
    while (chunk < 1024 &&
           static_cast<long long>(pb.n_fields) * ((B + chunk - 1) / chunk) > sm_count() * 192LL)
      chunk <<= 1;
    pb.chunk = chunk;
    pb.nchunks = (B + chunk - 1) / chunk;
  }
  const int tiny_ctas = pt.n_fields * (pt.n_fields > 0 ? pt.nchunks : 0);
  const long long big_tasks = static_cast<long long>(pb.n_fields) * (pb.n_fields > 0 ? pb.nchunks : 0);
  if (!fused) {
    BwdPair pp;
    pp.tiny = pt; pp.big = pb; pp.tiny_ctas = tiny_ctas;
    const int big_ctas = static_cast<int>(std::min<long long>((big_tasks + 7) / 8, sm_count() * 8LL));
    const int grid = tiny_ctas + big_ctas;
    if (grid == 0) return CTR_OK;
#define CTR_BWD(DD)                                                          \
  if (aggregate) embed_bwd_kernel<DD, true><<<grid, 256, 0, st>>>(pp);       \
  else embed_bwd_kernel<DD, false><<<grid, 256, 0, st>>>(pp);
    switch (D) {
      case 8: CTR_BWD(8) break;
      case 16: CTR_BWD(16) break;
      default: CTR_BWD(32) break;
    }
#undef CTR_BWD
    return check_cuda(cudaGetLastError(), fn);
  }
  // fused scatter + optimiser: two kernels (different register budgets), side by side
  cudaStream_t aux = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  const bool both = pt.n_fields > 0 && pb.n_fields > 0;
  if (both) {
    if (!aux_stream(&aux, &ev_fork, &ev_join)) return check_cuda(cudaGetLastError(), fn);
    cudaEventRecord(ev_fork, st);
    cudaStreamWaitEvent(aux, ev_fork, 0);
  }
  if (pt.n_fields > 0) {
    cudaStream_t s2 = both ? aux : st;
    switch (D) {
      case 8: embed_bwd_tiny_kernel<8, true><<<tiny_ctas, 256, 0, s2>>>(pt); break;
      case 16: embed_bwd_tiny_kernel<16, true><<<tiny_ctas, 256, 0, s2>>>(pt); break;
      default: embed_bwd_tiny_kernel<32, true><<<tiny_ctas, 256, 0, s2>>>(pt); break;
    }
  }
  if (pb.n_fields > 0) {
    const int grid = static_cast<int>(std::min<long long>((big_tasks + 3) / 4, sm_count() * 5LL));
    switch (D) {
      case 8: embed_bwd_adam_big_kernel<8><<<grid, 128, 0, st>>>(pb); break;
      case 16: embed_bwd_adam_big_kernel<16><<<grid, 128, 0, st>>>(pb); break;
      default: embed_bwd_adam_big_kernel<32><<<grid, 128, 0, st>>>(pb); break;
    }
  }
  if (both) {
    cudaEventRecord(ev_join, aux);
    cudaStreamWaitEvent(st, ev_join, 0);
  }
  return check_cuda(cudaGetLastError(), fn);
}

extern "C" {

int ctr_count_rows(const int32_t* rows, int64_t n, int D, float* rec, int64_t row_stride,
                   ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(rows && rec && n >= 0, "ctr_count_rows", "null pointer");
  CTR_REQUIRE((D == 8 || D == 16 || D == 32) && row_stride >= 4 * D + 8, "ctr_count_rows",
              "needs the row-record layout (row_stride >= 4D+8), D in {8,16,32}");
  if (n == 0) return CTR_OK;
  const int grid = static_cast<int>(std::min<long long>((n + 255) / 256, sm_count() * 8LL));
  count_rows_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(rows, n, rec, row_stride,
                                                                         4 * D + 5);
  CTR_LAUNCH_CHECK("ctr_count_rows");
}

int ctr_embed_bwd_adam(const int32_t* rows, const float* dE, const float* S, const float* dy2,
                       const float* dy1, uint64_t w1_fields, const int64_t* row_offsets_host, int B,
                       int F, int D, float* rec, int64_t row_stride, float lr_t, float beta1,
                       float beta2, float eps, const float* state_dev, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(rows && rec && row_offsets_host, "ctr_embed_bwd_adam", "null rows/rec/row_offsets");
  CTR_REQUIRE(B >= 0 && F > 0 && F <= CTR_MAX_FIELDS, "ctr_embed_bwd_adam", "need 0 < F <= 64");
  CTR_REQUIRE(dE || dy2, "ctr_embed_bwd_adam", "nothing to scatter: dE and dy2 both null");
  CTR_REQUIRE(!dy2 || S, "ctr_embed_bwd_adam", "dy2 needs S");
  CTR_REQUIRE(aligned16(dE) && aligned16(S) && aligned16(rec), "ctr_embed_bwd_adam",
              "pointers must be 16-byte aligned");
  CTR_REQUIRE((D == 8 || D == 16 || D == 32) && row_stride >= 4 * D + 8 && (row_stride & 3) == 0,
              "ctr_embed_bwd_adam", "needs the row-record layout (row_stride >= 4D+8, multiple of 4)");
  CTR_REQUIRE(row_offsets_host[F] < (1LL << 31), "ctr_embed_bwd_adam", "table too large for int32 rows");
  if (B == 0) return CTR_OK;
  BwdAdamParams p{};
  p.rows = rows; p.dE = dE; p.S = S; p.dy2 = dy2; p.dy1 = dy1; p.rec = rec; p.ld = row_stride;
  p.w1_fields = w1_fields; p.state = state_dev; p.lr_t = lr_t; p.k = AdamK{beta1, beta2, eps};
  p.B = B; p.F = F;
  for (int f = 0; f <= F; ++f) p.off[f] = static_cast<int>(row_offsets_host[f]);
  return launch_bwd(p, D, true, true, static_cast<cudaStream_t>(stream), "ctr_embed_bwd_adam");
}

int ctr_embed_bwd(const int32_t* rows, const float* dE, const float* E, const float* table,
                  const float* S, const float* dy2, const float* dy1, uint64_t w1_fields,
                  const int64_t* row_offsets_host, int B, int F, int D, float* dtable, float* dw1,
                  int64_t row_stride, int64_t w1_stride, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(rows && dtable && row_offsets_host, "ctr_embed_bwd", "null rows/dtable/row_offsets");
  CTR_REQUIRE(B >= 0 && F > 0 && F <= CTR_MAX_FIELDS, "ctr_embed_bwd", "need 0 < F <= 64");
  CTR_REQUIRE(dE || dy2, "ctr_embed_bwd", "nothing to scatter: dE and dy2 both null");
  CTR_REQUIRE(!dy2 || (S && (E || table)), "ctr_embed_bwd", "dy2 needs S and E (or table)");
  CTR_REQUIRE(aligned16(dE) && aligned16(E) && aligned16(table) && aligned16(S) && aligned16(dtable),
              "ctr_embed_bwd", "pointers must be 16-byte aligned");
  CTR_REQUIRE(D == 8 || D == 16 || D == 32, "ctr_embed_bwd", "D must be 8, 16 or 32");
  CTR_REQUIRE(row_offsets_host[F] < (1LL << 31), "ctr_embed_bwd", "table too large for int32 rows");
  if (row_stride <= 0) row_stride = D;
  if (w1_stride <= 0) w1_stride = 1;
  CTR_REQUIRE(row_stride >= D && (row_stride & 3) == 0, "ctr_embed_bwd",
              "row_stride must be >= D and a multiple of 4 floats");
  if (B == 0) return CTR_OK;
  BwdAdamParams p{};
  p.rows = rows; p.dE = dE; p.S = S; p.dy2 = dy2; p.dy1 = dy1; p.E = E; p.table = table;
  p.dtable = dtable; p.dw1 = dw1; p.ld_t = row_stride; p.ld_g = row_stride; p.ld_w = w1_stride;
  p.w1_fields = w1_fields; p.B = B; p.F = F;
  for (int f = 0; f <= F; ++f) p.off[f] = static_cast<int>(row_offsets_host[f]);
  return launch_bwd(p, D, false, option_get("bwd_aggregate", 1) != 0, static_cast<cudaStream_t>(stream),
                    "ctr_embed_bwd");
}

}  // extern "C"
