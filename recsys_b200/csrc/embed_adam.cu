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
#include "common.cuh"

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
  float lr_t, b1, b2, eps;
  int B, F, chunk, nchunks;
  int off[CTR_MAX_FIELDS + 1];
};

constexpr int kTinyRowsA = 32;

__device__ __forceinline__ float4 ldcg4(const float* p) {
  return __ldcg(reinterpret_cast<const float4*>(p));
}

template <int D>
struct Rec {
  static constexpr int TH = 0, M = D, V = 2 * D, G = 3 * D, TH1 = 4 * D, M1 = 4 * D + 1,
                       V1 = 4 * D + 2, G1 = 4 * D + 3, CNT = 4 * D + 5, CC = 4 * D + 6;
};

// U independent commits per call, phase by phase, so that their REDs, atomics and read-backs are
// in flight together.  In each, one group of LPR lanes owns one row (`row[u]` null = idle group;
// group-uniform): add the group's partial sums into the record, retire `n[u]` lookups of the row
// and - if that completes the row - apply the optimiser.  Warp-uniform call (warp collectives).
template <int D, int U>
__device__ __forceinline__ void commit_rows(float* const (&row)[U], const float4 (&acc)[U],
                                            const float (&gw)[U], const float (&gc)[U],
                                            const int (&n)[U], bool has_w1, bool has_c,
                                            const BwdAdamParams& p, float lr_t, int lane, int q) {
  using R = Rec<D>;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (row[u] != nullptr) {
      red_add_v4(row[u] + R::G + q * 4, acc[u]);
      if (q == 0) {
        if (has_w1) red_add_f32(row[u] + R::G1, gw[u]);
        if (has_c) red_add_f32(row[u] + R::CC, gc[u]);
      }
    }
  }
  __threadfence();      // this lane's REDs are performed before its group retires the lookups
  __syncwarp();
  int last[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    last[u] = 0;
    if (row[u] != nullptr && q == 0)
      last[u] = atomicSub(reinterpret_cast<int*>(row[u] + R::CNT), n[u]) == n[u] ? 1 : 0;
  }
  bool any = false;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    last[u] = __shfl_sync(0xffffffffu, last[u], lane - q);
    any |= last[u] != 0;
  }
  if (!any) return;
  __threadfence();      // every other arriver's REDs happened before its decrement
  // read the completed records back and update them, two rows at a time (register budget)
  constexpr int H = U >= 2 ? 2 : 1;
#pragma unroll
  for (int u0 = 0; u0 < U; u0 += H) {
    float4 G[H], T[H], Mv[H], Vv[H], f1[H];
    float c[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const int u = u0 + h;
      if (!last[u]) continue;
      G[h] = ldcg4(row[u] + R::G + q * 4);
      T[h] = ldcg4(row[u] + R::TH + q * 4);
      Mv[h] = ldcg4(row[u] + R::M + q * 4);
      Vv[h] = ldcg4(row[u] + R::V + q * 4);
      c[h] = has_c ? __ldcg(row[u] + R::CC) : 0.f;
      if (has_w1 && q == 0) f1[h] = ldcg4(row[u] + R::TH1);     // theta1 m1 v1 g1
    }
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const int u = u0 + h;
      if (!last[u]) continue;
      float4 g = G[h], t = T[h], m = Mv[h], v = Vv[h];
      if (has_c) {    // FM: sum_slots dy2 * (S - E) = sum dy2 * S - (sum dy2) * theta
        g.x = fmaf(-c[h], t.x, g.x); g.y = fmaf(-c[h], t.y, g.y);
        g.z = fmaf(-c[h], t.z, g.z); g.w = fmaf(-c[h], t.w, g.w);
      }
#define CTR_ADAM4(k)                                   \
  m.k = p.b1 * m.k + (1.f - p.b1) * g.k;               \
  v.k = p.b2 * v.k + (1.f - p.b2) * g.k * g.k;         \
  t.k -= lr_t * m.k / (sqrtf(v.k) + p.eps);
      CTR_ADAM4(x) CTR_ADAM4(y) CTR_ADAM4(z) CTR_ADAM4(w)
#undef CTR_ADAM4
      *reinterpret_cast<float4*>(row[u] + R::M + q * 4) = m;
      *reinterpret_cast<float4*>(row[u] + R::V + q * 4) = v;
      *reinterpret_cast<float4*>(row[u] + R::TH + q * 4) = t;
      *reinterpret_cast<float4*>(row[u] + R::G + q * 4) = f4_zero();
      if (q == 0) {
        if (has_w1) {   // the row's first-order weight rides on the same completion
          const float Mn = p.b1 * f1[h].y + (1.f - p.b1) * f1[h].w;
          const float Vn = p.b2 * f1[h].z + (1.f - p.b2) * f1[h].w * f1[h].w;
          *reinterpret_cast<float4*>(row[u] + R::TH1) =
              make_float4(f1[h].x - lr_t * Mn / (sqrtf(Vn) + p.eps), Mn, Vn, 0.f);
        }
        if (has_c) row[u][R::CC] = 0.f;
      }
    }
  }
}

template <int D>
__global__ void __launch_bounds__(256, 2) embed_bwd_adam_kernel(const BwdAdamParams p) {
  using R = Rec<D>;
  constexpr int LPR = D / 4;
  constexpr int RPW = 32 / LPR;
  constexpr int J = kTinyRowsA / RPW;
  constexpr int PT = D + 4;
  __shared__ __align__(16) float s_tiny[8 * kTinyRowsA * PT];
  __shared__ float s_tinyw[8 * kTinyRowsA];
  __shared__ float s_tinyc[8 * kTinyRowsA];
  __shared__ int s_tinyn[8 * kTinyRowsA];
  const int lane = threadIdx.x & 31;
  const int r = lane / LPR;
  const int q = lane % LPR;
  const int F = p.F;
  const int wpb = blockDim.x >> 5;
  const int ntask = F * p.nchunks;
  const float lr_t = p.state != nullptr ? p.state[1] : p.lr_t;
  const bool has_c = p.dy2 != nullptr;
  const unsigned qmask = (LPR == 4 ? 0x11111111u : LPR == 2 ? 0x55555555u : 0x01010101u) << q;

  for (int task = blockIdx.x * wpb + (threadIdx.x >> 5); task < ntask; task += gridDim.x * wpb) {
    const int f = task % F;
    const int c = task / F;
    const int b_begin = c * p.chunk;
    const int b_end = min(p.B, b_begin + p.chunk);
    const int off = p.off[f];
    const int nrow = p.off[f + 1] - off;
    const bool tiny = nrow <= kTinyRowsA;
    const bool has_w1 = p.dy1 != nullptr && ((p.w1_fields >> f) & 1ull);

    // partial gradient of one (sample, field) slot without the -dy2*E term (see the header)
    auto slot_grad = [&](int b, float4& g, float& gw, float& gc) {
      const size_t eo = (static_cast<size_t>(b) * F + f) * D + q * 4;
      g = p.dE != nullptr ? ld4_stream(p.dE + eo) : f4_zero();
      gc = 0.f;
      if (has_c) {
        const float4 sv = ldg4(p.S + static_cast<size_t>(b) * D + q * 4);
        gc = __ldg(p.dy2 + b);
        g.x = fmaf(gc, sv.x, g.x); g.y = fmaf(gc, sv.y, g.y);
        g.z = fmaf(gc, sv.z, g.z); g.w = fmaf(gc, sv.w, g.w);
      }
      gw = has_w1 ? __ldg(p.dy1 + b) : 0.f;
    };

    if (!tiny) {
      constexpr int UNR = 4;
      for (int bb = b_begin; bb < b_end; bb += RPW * UNR) {
        int rid[UNR], n[UNR];
        float4 g[UNR];
        float gw[UNR], gc[UNR];
        float* row[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int b = bb + u * RPW + r;
          rid[u] = b < b_end ? __ldg(p.rows + static_cast<size_t>(b) * F + f) : -1;
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          g[u] = f4_zero();
          gw[u] = gc[u] = 0.f;
          if (rid[u] >= 0) slot_grad(bb + u * RPW + r, g[u], gw[u], gc[u]);
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          // groups of this instruction that hit the same row: the first one commits their sum
          const unsigned peers = __match_any_sync(0xffffffffu, rid[u] >= 0 ? rid[u] : -1 - r) & qmask;
          n[u] = __popc(peers);
          const bool leader = (peers & ((1u << lane) - 1u)) == 0u;
          if (__reduce_max_sync(0xffffffffu, n[u]) > 1) {      // warp-uniform, rare for large tables
            unsigned rest = peers & ~(1u << lane);
            for (int it = 1; it < RPW; ++it) {
              const int src = rest != 0u ? __ffs(rest) - 1 : lane;
              const bool ok = rest != 0u;
              rest &= rest - 1u;
              const float4 o = f4_shfl(g[u], src);
              const float ow = __shfl_sync(0xffffffffu, gw[u], src);
              const float oc = __shfl_sync(0xffffffffu, gc[u], src);
              if (ok && leader) {
                g[u] = f4_add(g[u], o);
                gw[u] += ow;
                gc[u] += oc;
              }
              if (__ballot_sync(0xffffffffu, rest != 0u) == 0u) break;
            }
          }
          row[u] = (rid[u] >= 0 && leader) ? p.rec + static_cast<size_t>(rid[u]) * p.ld : nullptr;
        }
        commit_rows<D, UNR>(row, g, gw, gc, n, has_w1, has_c, p, lr_t, lane, q);
      }
      continue;
    }

    // Tiny field: sum the chunk per row in a warp-private shared-memory tile, then one commit per
    // row that was hit (with its multiplicity).
    const int wslot = threadIdx.x >> 5;
    float* sacc = &s_tiny[wslot * kTinyRowsA * PT];
    float* saccw = &s_tinyw[wslot * kTinyRowsA];
    float* saccc = &s_tinyc[wslot * kTinyRowsA];
    int* saccn = &s_tinyn[wslot * kTinyRowsA];
    for (int i = lane; i < kTinyRowsA * PT / 4; i += 32) reinterpret_cast<float4*>(sacc)[i] = f4_zero();
    saccw[lane] = 0.f;
    saccc[lane] = 0.f;
    saccn[lane] = 0;
    __syncwarp();
    constexpr int TU = 4;
    for (int bb = b_begin; bb < b_end; bb += RPW * TU) {
      int lid[TU];
      float4 g[TU];
      float gw[TU], gc[TU];
#pragma unroll
      for (int u = 0; u < TU; ++u) {
        const int b = bb + u * RPW + r;
        lid[u] = b < b_end ? __ldg(p.rows + static_cast<size_t>(b) * F + f) : -1;
      }
#pragma unroll
      for (int u = 0; u < TU; ++u) {
        g[u] = f4_zero();
        gw[u] = gc[u] = 0.f;
        if (lid[u] >= 0) slot_grad(bb + u * RPW + r, g[u], gw[u], gc[u]);
        lid[u] = lid[u] >= 0 ? lid[u] - off : -1;
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
              saccw[lid[u]] += gw[u];
              saccc[lid[u]] += gc[u];
              saccn[lid[u]] += 1;
            }
          }
          __syncwarp();
        }
      }
    }
    {
      float* row[J];
      float4 a[J];
      float aw[J], ac[J];
      int n[J];
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int lr = r + RPW * j;
        n[j] = lr < nrow ? saccn[lr] : 0;
        row[j] = n[j] > 0 ? p.rec + static_cast<size_t>(off + lr) * p.ld : nullptr;
        a[j] = n[j] > 0 ? *reinterpret_cast<const float4*>(&sacc[lr * PT + q * 4]) : f4_zero();
        aw[j] = n[j] > 0 ? saccw[lr] : 0.f;
        ac[j] = n[j] > 0 ? saccc[lr] : 0.f;
      }
      commit_rows<D, J>(row, a, aw, ac, n, has_w1, has_c, p, lr_t, lane, q);
    }
    __syncwarp();
  }
}

// cnt[row] += lookups of the row in rows[n]; duplicates inside a warp instruction are combined.
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
    if (rid >= 0 && (peers & ((1u << lane) - 1u)) == 0u)
      atomicAdd(reinterpret_cast<int*>(rec + static_cast<size_t>(rid) * ld + cnt_off), __popc(peers));
  }
}

}  // namespace ctr

using namespace ctr;

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
  BwdAdamParams p;
  p.rows = rows; p.dE = dE; p.S = S; p.dy2 = dy2; p.dy1 = dy1; p.rec = rec; p.ld = row_stride;
  p.w1_fields = w1_fields; p.state = state_dev; p.lr_t = lr_t; p.b1 = beta1; p.b2 = beta2; p.eps = eps;
  p.B = B; p.F = F;
  for (int f = 0; f <= F; ++f) p.off[f] = static_cast<int>(row_offsets_host[f]);
  int chunk = 64;
  while (chunk < 1024 && static_cast<long long>(F) * ((B + chunk - 1) / chunk) > sm_count() * 64LL)
    chunk <<= 1;
  p.chunk = chunk;
  p.nchunks = (B + chunk - 1) / chunk;
  const long long ntask = static_cast<long long>(F) * p.nchunks;
  const int grid = static_cast<int>(std::min<long long>((ntask + 7) / 8, sm_count() * 8LL));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (D) {
    case 8: embed_bwd_adam_kernel<8><<<grid, 256, 0, st>>>(p); break;
    case 16: embed_bwd_adam_kernel<16><<<grid, 256, 0, st>>>(p); break;
    default: embed_bwd_adam_kernel<32><<<grid, 256, 0, st>>>(p); break;
  }
  CTR_LAUNCH_CHECK("ctr_embed_bwd_adam");
}

}  // extern "C"
