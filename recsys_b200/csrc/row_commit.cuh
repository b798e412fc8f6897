// Row-record commit shared by the fused scatter + optimiser kernels (embed_adam.cu, p2p.cu):
// add a lookup's partial gradient sums into its record, retire the lookup(s), and - on the lookup
// that completes the row - apply TF-Adam in place.  See embed_adam.cu for the scheme.
#pragma once
#include "common.cuh"

namespace ctr {

struct AdamK {
  float b1, b2, eps;
};

__device__ __forceinline__ float4 ldcg4(const float* p) {
  return __ldcg(reinterpret_cast<const float4*>(p));
}

// record[r] = { theta[D] | m[D] | v[D] | g[D] | theta1 m1 v1 g1 | claim cnt c arr }
//   cnt: lookups of the row in the current batch (ctr_count_rows / K2 of the peer exchange),
//   arr: lookups retired so far by the scatter pass, c: sum of dy2 (FM term), all zero between steps.
template <int D>
struct Rec {
  static constexpr int TH = 0, M = D, V = 2 * D, G = 3 * D, TH1 = 4 * D, M1 = 4 * D + 1,
                       V1 = 4 * D + 2, G1 = 4 * D + 3, CLAIM = 4 * D + 4, CNT = 4 * D + 5,
                       CC = 4 * D + 6, ARR = 4 * D + 7;
};

// A row's theta / m / v quarter and its tail, requested together with the slot's gradient (PRE):
// nobody writes theta, m, v or cnt before the row's completion, so an early copy is valid.
struct RowPre {
  float4 T, M, V;
  float4 F1;      // theta1 m1 v1 g1        (lane q == 0 of the group)
  int cnt;        // lookups of the row in this batch
};
template <int D>
__device__ __forceinline__ void row_preload(const float* row, int q, RowPre& o) {
  using R = Rec<D>;
  o.T = ldcg4(row + R::TH + q * 4);
  o.M = ldcg4(row + R::M + q * 4);
  o.V = ldcg4(row + R::V + q * 4);
  if (q == 0) o.F1 = ldcg4(row + R::TH1);
  o.cnt = __ldcg(reinterpret_cast<const int*>(row + R::CNT));
}

__device__ __forceinline__ void adam4(float4& t, float4& m, float4& v, const float4& g,
                                      const AdamK& p, float lr_t) {
#define CTR_ADAM4(k)                                   \
  m.k = p.b1 * m.k + (1.f - p.b1) * g.k;               \
  v.k = p.b2 * v.k + (1.f - p.b2) * g.k * g.k;         \
  t.k -= lr_t * m.k / (sqrtf(v.k) + p.eps);
  CTR_ADAM4(x) CTR_ADAM4(y) CTR_ADAM4(z) CTR_ADAM4(w)
#undef CTR_ADAM4
}
// theta1 block {theta1 m1 v1 g1} after one Adam step on gradient g1
__device__ __forceinline__ float4 adam1(const float4& f, float g1, const AdamK& p, float lr_t) {
  const float Mn = p.b1 * f.y + (1.f - p.b1) * g1;
  const float Vn = p.b2 * f.z + (1.f - p.b2) * g1 * g1;
  return make_float4(f.x - lr_t * Mn / (sqrtf(Vn) + p.eps), Mn, Vn, 0.f);
}

// U independent commits per call, phase by phase, so that their memory operations are in flight
// together.  In each, one group of LPR lanes owns one row (`row[u]` null = idle group;
// group-uniform) and holds the sum `acc/gw/gc` of n[u] of the row's lookups (gw / gc identical
// in the group's lanes).
//   PRE and cnt == n[u]  (the group holds ALL lookups of the row - the common case in a large
//     table): the update is computed from registers and stored; no atomic, no fence.
//   otherwise: RED the sums into the record, fence, arr += n[u]; the group that takes arr to cnt is
//     the row's last arriver: it reads the accumulators back, updates, and clears them.
// Warp-uniform call (warp collectives).
template <int D, int U, bool PRE>
__device__ __forceinline__ void commit_rows(float* const (&row)[U], const float4 (&acc)[U],
                                            const float (&gw)[U], const float (&gc)[U],
                                            const int (&n)[U], const RowPre (&pre)[U], bool has_w1,
                                            bool has_c, const AdamK& p, float lr_t, int lane, int q) {
  using R = Rec<D>;
  bool slow[U];
  bool any_slow = false;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    slow[u] = row[u] != nullptr;
    if (PRE) {
      if (row[u] != nullptr && pre[u].cnt == n[u]) {
        slow[u] = false;
        float4 g = acc[u], t = pre[u].T, m = pre[u].M, v = pre[u].V;
        if (has_c) {    // FM: sum_slots dy2 * (S - E) = sum dy2 * S - (sum dy2) * theta
          const float c = gc[u];          // group-uniform (every lane of the group formed it)
          g.x = fmaf(-c, t.x, g.x); g.y = fmaf(-c, t.y, g.y);
          g.z = fmaf(-c, t.z, g.z); g.w = fmaf(-c, t.w, g.w);
        }
        adam4(t, m, v, g, p, lr_t);
        *reinterpret_cast<float4*>(row[u] + R::M + q * 4) = m;
        *reinterpret_cast<float4*>(row[u] + R::V + q * 4) = v;
        *reinterpret_cast<float4*>(row[u] + R::TH + q * 4) = t;
        if (q == 0) {
          if (has_w1) *reinterpret_cast<float4*>(row[u] + R::TH1) = adam1(pre[u].F1, gw[u], p, lr_t);
          *reinterpret_cast<int*>(row[u] + R::CNT) = 0;
        }
      }
    }
    any_slow |= slow[u];
  }
  if (__ballot_sync(0xffffffffu, any_slow) == 0u) return;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (slow[u]) {
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
    if (slow[u] && q == 0) {
      const int cnt = PRE ? pre[u].cnt : __ldcg(reinterpret_cast<const int*>(row[u] + R::CNT));
      last[u] = atomicAdd(reinterpret_cast<int*>(row[u] + R::ARR), n[u]) + n[u] == cnt ? 1 : 0;
    }
  }
  bool any = false;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    last[u] = __shfl_sync(0xffffffffu, last[u], lane - q);
    any |= last[u] != 0;
  }
  if (!any) return;
  __threadfence();      // every other arriver's REDs happened before its arrival count
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (!last[u]) continue;
    float4 g = ldcg4(row[u] + R::G + q * 4);
    float4 t, m, v;
    if (PRE) {
      t = pre[u].T; m = pre[u].M; v = pre[u].V;
    } else {
      t = ldcg4(row[u] + R::TH + q * 4);
      m = ldcg4(row[u] + R::M + q * 4);
      v = ldcg4(row[u] + R::V + q * 4);
    }
    if (has_c) {
      const float c = __ldcg(row[u] + R::CC);
      g.x = fmaf(-c, t.x, g.x); g.y = fmaf(-c, t.y, g.y);
      g.z = fmaf(-c, t.z, g.z); g.w = fmaf(-c, t.w, g.w);
    }
    adam4(t, m, v, g, p, lr_t);
    *reinterpret_cast<float4*>(row[u] + R::M + q * 4) = m;
    *reinterpret_cast<float4*>(row[u] + R::V + q * 4) = v;
    *reinterpret_cast<float4*>(row[u] + R::TH + q * 4) = t;
    *reinterpret_cast<float4*>(row[u] + R::G + q * 4) = f4_zero();
    if (q == 0) {
      if (has_w1) {     // the row's first-order weight rides on the same completion
        const float4 f = ldcg4(row[u] + R::TH1);
        *reinterpret_cast<float4*>(row[u] + R::TH1) = adam1(f, f.w, p, lr_t);
      }
      // claim cnt c arr: the claim word belongs to the unfused optimiser, the rest is cleared
      const float claim = __ldcg(row[u] + R::CLAIM);
      *reinterpret_cast<float4*>(row[u] + R::CLAIM) = make_float4(claim, 0.f, 0.f, 0.f);
    }
  }
}

// Sum, over the groups of one warp instruction that hit the same row (`key` equal; idle groups
// pass a unique negative key), their partial gradients into the first such group.  Returns the
// multiplicity; `leader` says whether this group commits.  Warp-uniform call.
template <int LPR>
__device__ __forceinline__ int merge_duplicates(int key, float4& g, float& gw, float& gc, bool& leader,
                                                int lane, int q) {
  constexpr int RPW = 32 / LPR;
  const unsigned qmask = (LPR == 4 ? 0x11111111u : LPR == 2 ? 0x55555555u : 0x01010101u) << q;
  const unsigned peers = __match_any_sync(0xffffffffu, key) & qmask;
  const int n = __popc(peers);
  leader = (peers & ((1u << lane) - 1u)) == 0u;
  if (__reduce_max_sync(0xffffffffu, n) > 1) {        // warp-uniform, rare for large tables
    unsigned rest = peers & ~(1u << lane);
    for (int it = 1; it < RPW; ++it) {
      const int src = rest != 0u ? __ffs(rest) - 1 : lane;
      const bool ok = rest != 0u;
      rest &= rest - 1u;
      const float4 o = f4_shfl(g, src);
      const float ow = __shfl_sync(0xffffffffu, gw, src);
      const float oc = __shfl_sync(0xffffffffu, gc, src);
      if (ok && leader) {
        g = f4_add(g, o);
        gw += ow;
        gc += oc;
      }
      if (__ballot_sync(0xffffffffu, rest != 0u) == 0u) break;
    }
  }
  return n;
}

}  // namespace ctr
