// Row-sharded embedding table, device-initiated exchange over NVLink peer memory.
//
// One process per GPU.  Every rank owns a same-shaped ARENA in its HBM and maps every peer's arena
// (cudaIpc); the kernels of a step store straight into the arena of the rank that consumes the
// data and raise a flag there - no NCCL launch, no staging copy, no host involvement, the whole
// step stays one CUDA graph per rank:
//
//   K1 p2p_bucket_send     requester: bucket the batch's lookups by owner (row % G) and store the
//                          owner-local ids into the OWNER's req_ids[parity][me][pos]
//   K2 p2p_gather_reply    owner: for every requester, wait for its ids, gather row | w1 from the
//                          local row records and store them into the REQUESTER's resp[me][pos];
//                          also counts the lookups per row (for the fused optimiser of K5)
//   K3 ctr_embed_fwd       requester: the ordinary fused lookup + interaction kernel over the resp
//                          slab (slots as row ids), after waiting for every owner's reply flag
//   K4 p2p_grad_send       requester: per slot g = dE + dy2*S, tail {dy1, dy2} stored into the
//                          OWNER's grad[me][pos]
//   K5 p2p_scatter_adam    owner: wait for every requester's gradients; scatter-add and TF-Adam in
//                          one pass (row_commit.cuh: the lookup that completes a row updates it)
//   K6 p2p_dense_push / p2p_adam_dense   replicated dense weights: every rank stores its gradient
//                          into every peer's dense[me]; the optimiser sums the G copies in rank
//                          order (bitwise identical on every rank) - the all-reduce without NCCL
//
// Flags carry the step's sequence number (monotonic), written with a system-scope release after
// the data, read with a system-scope acquire; every wait is bounded (a peer that never arrives
// sets arena.err instead of hanging the GPU).  Slabs are sized for the worst case (capacity =
// lookups per rank), so nothing can overflow.  req_ids / req_cnt are double buffered by step
// parity: a fast requester may post step s+1 while the owner still scatters step s.
#include "row_commit.cuh"

namespace ctr {

struct P2P {
  char* peer[CTR_P2P_MAX_RANKS];
  int me, G, cap, P;
  long long off_req_flag, off_req_cnt, off_resp_flag, off_grad_flag, off_dense_flag;
  long long off_req_ids, off_resp, off_grad, off_dense, off_counts, off_done;
  long long off_inv, off_sent;      // requester side: inv[G][cap] lookup index per slab position, sent[G]
  long long n_dense;
  long long spin_limit_ns;
};
// arena header: {step, err, pad...}
__device__ __forceinline__ int* hdr(const P2P& c, int r) { return reinterpret_cast<int*>(c.peer[r]); }
__device__ __forceinline__ int* iptr(const P2P& c, int r, long long off) {
  return reinterpret_cast<int*>(c.peer[r] + off);
}
__device__ __forceinline__ float* fptr(const P2P& c, int r, long long off) {
  return reinterpret_cast<float*>(c.peer[r] + off);
}

// Wait until *flag >= seq; on timeout sets arena.err and returns.
__device__ __forceinline__ void wait_flag(const P2P& c, const int* flag, int seq) {
  wait_flag_bounded(flag, seq, c.spin_limit_ns, hdr(c, c.me) + 1);
}

// Last-block detection: every block calls this once after its last global store; returns true in
// exactly one block (thread 0's value broadcast through shared memory), after all blocks arrived.
__device__ __forceinline__ bool last_block(int* counter, int nblocks) {
  __shared__ int s_last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int old = atomicAdd(counter, 1);
    s_last = old == nblocks - 1 ? 1 : 0;
    if (s_last) {
      *counter = 0;
      __threadfence();
    }
  }
  __syncthreads();
  return s_last != 0;
}

// ----------------------------------------------------------------------------- K1
// CTA = a contiguous run of kBktPerThread x 256 lookups.  Positions are handed out in two levels:
// lanes of a warp that target the same owner take consecutive ranks with one shared-memory atomic,
// the CTA reserves its range of every owner's slab with ONE global atomic per owner.  Besides the
// owner-local id stored into the owner's arena, the requester keeps inv[owner][pos] = lookup index
// (so that K4 can stream the gradients out in slab order) and slot[i] = owner * cap + pos (what K3
// indexes the reply slab with).
constexpr int kBktPerThread = 4;
__global__ void __launch_bounds__(256)
p2p_bucket_send_kernel(const int* __restrict__ rows, long long n, const P2P c, int* __restrict__ slot) {
  __shared__ int s_cnt[CTR_P2P_MAX_RANKS], s_base[CTR_P2P_MAX_RANKS];
  const int seq = hdr(c, c.me)[0] + 1;
  const int par = seq & 1;
  int* counts = iptr(c, c.me, c.off_counts);
  int* inv = iptr(c, c.me, c.off_inv);
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < CTR_P2P_MAX_RANKS) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const long long i0 = static_cast<long long>(blockIdx.x) * (256 * kBktPerThread) + threadIdx.x;
  int row[kBktPerThread], owner[kBktPerThread], rank[kBktPerThread];
#pragma unroll
  for (int u = 0; u < kBktPerThread; ++u) {
    const long long i = i0 + u * 256;
    row[u] = i < n ? __ldg(rows + i) : -1;
  }
#pragma unroll
  for (int u = 0; u < kBktPerThread; ++u) {
    owner[u] = row[u] >= 0 ? row[u] % c.G : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, owner[u]);
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (lane == leader && owner[u] >= 0) base = atomicAdd(&s_cnt[owner[u]], __popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    rank[u] = base + __popc(peers & ((1u << lane) - 1u));
  }
  __syncthreads();
  if (threadIdx.x < c.G) s_base[threadIdx.x] = atomicAdd(counts + threadIdx.x, s_cnt[threadIdx.x]);
  __syncthreads();
#pragma unroll
  for (int u = 0; u < kBktPerThread; ++u) {
    if (owner[u] < 0) continue;
    const long long i = i0 + u * 256;
    const int pos = s_base[owner[u]] + rank[u];          // < cap by construction (cap >= n)
    int* dst = iptr(c, owner[u], c.off_req_ids) +
               (static_cast<long long>(par) * c.G + c.me) * c.cap + pos;
    *dst = row[u] / c.G;
    inv[static_cast<long long>(owner[u]) * c.cap + pos] = static_cast<int>(i);
    slot[i] = owner[u] * c.cap + pos;
  }
  if (last_block(iptr(c, c.me, c.off_done), gridDim.x)) {
    if (threadIdx.x < c.G) {
      const int o = threadIdx.x;
      const int cnt = *reinterpret_cast<volatile int*>(counts + o);
      *reinterpret_cast<volatile int*>(iptr(c, o, c.off_req_cnt) + par * c.G + c.me) = cnt;
      __threadfence_system();
      st_release_sys(iptr(c, o, c.off_req_flag) + par * c.G + c.me, seq);
      iptr(c, c.me, c.off_sent)[o] = cnt;               // K4 of this step reads it
      counts[o] = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) hdr(c, c.me)[0] = seq;       // the step is now `seq` for K2..K6
  }
}

// ----------------------------------------------------------------------------- K2
// CTA b serves requester src = b % G (share b / G of gridDim.x / G).  A record travels as
// P = D + 4 floats: row | w1 | pad.
template <int D>
__global__ void __launch_bounds__(256)
p2p_gather_reply_kernel(float* __restrict__ rec, long long ld, int with_w1, int count, const P2P c) {
  using R = Rec<D>;
  constexpr int LPR = D / 4;
  const int seq = hdr(c, c.me)[0];
  const int par = seq & 1;
  const int src = blockIdx.x % c.G;
  const int share = blockIdx.x / c.G, nshare = gridDim.x / c.G;
  if (share >= nshare) return;                  // gridDim.x is a multiple of G (launcher)
  __shared__ int s_cnt;
  if (threadIdx.x == 0) {
    wait_flag(c, iptr(c, c.me, c.off_req_flag) + par * c.G + src, seq);
    s_cnt = *reinterpret_cast<volatile int*>(iptr(c, c.me, c.off_req_cnt) + par * c.G + src);
  }
  __syncthreads();
  const int cnt = min(s_cnt, c.cap);
  const int* ids = iptr(c, c.me, c.off_req_ids) + (static_cast<long long>(par) * c.G + src) * c.cap;
  float* out = fptr(c, src, c.off_resp) + static_cast<long long>(c.me) * c.cap * c.P;
  const int gpb = blockDim.x / LPR;
  const int q = threadIdx.x % LPR;
  for (int pos = share * gpb + threadIdx.x / LPR; pos < cnt; pos += nshare * gpb) {
    const int id = __ldcg(ids + pos);
    const float* row = rec + static_cast<size_t>(id) * ld;
    const float4 v = ldg4(row + R::TH + q * 4);
    float* o = out + static_cast<long long>(pos) * c.P;
    *reinterpret_cast<float4*>(o + q * 4) = v;
    if (q == 0) {
      *reinterpret_cast<float4*>(o + D) = make_float4(with_w1 ? __ldg(row + R::TH1) : 0.f, 0.f, 0.f, 0.f);
      (void)count;
    }
  }
  // per-requester completion: the last of this requester's CTAs raises its reply flag
  if (last_block(iptr(c, c.me, c.off_done) + 1 + src, nshare)) {
    if (threadIdx.x == 0) st_release_sys(iptr(c, src, c.off_resp_flag) + c.me, seq);
  }
}

// Wait for every owner's reply (or every requester's gradients): 1 CTA, used when the consumer
// kernel is not one of ours to extend.  The fused lookup kernel waits itself (ctr_embed_fwd_wait).
__global__ void p2p_wait_kernel(const P2P c, long long off_flags) {
  if (threadIdx.x < c.G) wait_flag(c, iptr(c, c.me, off_flags) + threadIdx.x, hdr(c, c.me)[0]);
}

// ----------------------------------------------------------------------------- K4
// Slab order: CTA b serves owner o = b % G; five lanes per record (four gradient quarters + the
// tail {dy1, dy2}), six records per warp instruction, so the remote stores of a warp are one
// contiguous 480-byte run of the owner's grad[me] slab (NVLink likes long writes; the scattered
// side - dE[b,f,:] through inv[] - is local).
template <int D>
__global__ void __launch_bounds__(256)
p2p_grad_send_kernel(const float* __restrict__ dE, const float* __restrict__ S,
                     const float* __restrict__ dy2, const float* __restrict__ dy1,
                     unsigned long long w1_fields, int F, const P2P c) {
  constexpr int LPR = D / 4;
  constexpr int GL = LPR + 1;                 // lanes per record
  constexpr int RPW = 32 / GL;                // records per warp instruction
  const int seq = hdr(c, c.me)[0];
  const int o = blockIdx.x % c.G;
  const int share = blockIdx.x / c.G, nshare = gridDim.x / c.G;
  const int lane = threadIdx.x & 31, r = lane / GL, q = lane % GL;
  const int wpb = blockDim.x >> 5;
  if (share < nshare) {
    const int cnt = min(iptr(c, c.me, c.off_sent)[o], c.cap);
    const int* inv = iptr(c, c.me, c.off_inv) + static_cast<long long>(o) * c.cap;
    float* out = fptr(c, o, c.off_grad) + static_cast<long long>(c.me) * c.cap * c.P;
    const int w = share * wpb + (threadIdx.x >> 5), nw = nshare * wpb;
    constexpr int U = 4;                                   // records in flight per lane group
    for (int p0 = w * RPW * U; p0 < cnt; p0 += nw * RPW * U) {
      int idx[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pos = p0 + u * RPW + r;
        idx[u] = (r < RPW && pos < cnt) ? __ldg(inv + pos) : -1;
      }
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (idx[u] < 0) continue;
        const int i = idx[u];
        const int b = i / F, f = i - b * F;
        if (q < LPR) {
          v[u] = dE != nullptr ? ld4_stream(dE + static_cast<size_t>(i) * D + q * 4) : f4_zero();
          if (dy2 != nullptr) {
            const float cdy = __ldg(dy2 + b);
            const float4 sv = ldg4(S + static_cast<size_t>(b) * D + q * 4);
            v[u].x = fmaf(cdy, sv.x, v[u].x); v[u].y = fmaf(cdy, sv.y, v[u].y);
            v[u].z = fmaf(cdy, sv.z, v[u].z); v[u].w = fmaf(cdy, sv.w, v[u].w);
          }
        } else {
          v[u] = make_float4((dy1 != nullptr && ((w1_fields >> f) & 1ull)) ? __ldg(dy1 + b) : 0.f,
                             dy2 != nullptr ? __ldg(dy2 + b) : 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (idx[u] < 0) continue;
        const int pos = p0 + u * RPW + r;
        *reinterpret_cast<float4*>(out + static_cast<long long>(pos) * c.P + q * 4) = v[u];
      }
    }
  }
  if (last_block(iptr(c, c.me, c.off_done), gridDim.x)) {
    if (threadIdx.x < c.G) st_release_sys(iptr(c, threadIdx.x, c.off_grad_flag) + c.me, seq);
  }
}

// ----------------------------------------------------------------------------- K5
// Owner side, two passes over the received (id, gradient record) pairs of every requester - the
// measured-faster form of the scatter + optimiser (a one-pass variant with per-row completion
// counts ran 53 us against 36 us for this pair at 160 K lookups):
//   K5a p2p_scatter        wait for the requester's gradients; g[row] += record, g1 += dy1,
//                          c += dy2 (vector REDs; duplicates inside a warp instruction are summed
//                          in registers first)
//   K5b p2p_adam_rows      exactly one update per distinct row (claim word tagged with the step
//                          number): gradient g - c*theta (FM term without re-reading E), TF-Adam on
//                          theta and theta1, accumulators cleared
template <int D>
__global__ void __launch_bounds__(256)
p2p_scatter_kernel(float* __restrict__ rec, long long ld, int with_w1, int has_c, const P2P c) {
  using R = Rec<D>;
  constexpr int LPR = D / 4;
  constexpr int RPW = 32 / LPR;
  constexpr int U = 4;
  const int seq = hdr(c, c.me)[0];
  const int par = seq & 1;
  const int src = blockIdx.x % c.G;
  const int share = blockIdx.x / c.G, nshare = gridDim.x / c.G;
  if (share >= nshare) return;
  __shared__ int s_cnt;
  if (threadIdx.x == 0) {
    wait_flag(c, iptr(c, c.me, c.off_grad_flag) + src, seq);
    s_cnt = *reinterpret_cast<volatile int*>(iptr(c, c.me, c.off_req_cnt) + par * c.G + src);
  }
  __syncthreads();
  const int cnt = min(s_cnt, c.cap);
  const int* ids = iptr(c, c.me, c.off_req_ids) + (static_cast<long long>(par) * c.G + src) * c.cap;
  const float* gin = fptr(c, c.me, c.off_grad) + static_cast<long long>(src) * c.cap * c.P;
  const int lane = threadIdx.x & 31, r = lane / LPR, q = lane % LPR;
  const int wpb = blockDim.x >> 5;
  const int per_iter = RPW * U;
  const int w = share * wpb + (threadIdx.x >> 5), nw = nshare * wpb;
  for (int p0 = w * per_iter; p0 < cnt; p0 += nw * per_iter) {      // warp-uniform trip count
    int id[U];
    float4 g[U];
    float gw[U], gc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pos = p0 + u * RPW + r;
      id[u] = pos < cnt ? __ldcg(ids + pos) : -1;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pos = p0 + u * RPW + r;
      g[u] = f4_zero();
      gw[u] = gc[u] = 0.f;
      if (id[u] >= 0) {
        const float* in = gin + static_cast<long long>(pos) * c.P;
        g[u] = ldcg4(in + q * 4);
        const float4 tail = ldcg4(in + D);
        gw[u] = tail.x;
        gc[u] = tail.y;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      bool leader;
      merge_duplicates<LPR>(id[u] >= 0 ? id[u] : -1 - r, g[u], gw[u], gc[u], leader, lane, q);
      if (id[u] >= 0 && leader) {
        float* row = rec + static_cast<size_t>(id[u]) * ld;
        red_add_v4(row + R::G + q * 4, g[u]);
        if (q == 0) {
          if (with_w1) red_add_f32(row + R::G1, gw[u]);
          if (has_c) red_add_f32(row + R::CC, gc[u]);
        }
      }
    }
  }
}

template <int D>
__global__ void __launch_bounds__(256)
p2p_adam_rows_kernel(float* __restrict__ rec, long long ld, int with_w1, int has_c,
                     const float* __restrict__ state, float lr_t_arg, AdamK k, const P2P c) {
  using R = Rec<D>;
  constexpr int LPR = D / 4;
  constexpr int RPW = 32 / LPR;
  constexpr int U = 2;
  const int seq = hdr(c, c.me)[0];
  const int par = seq & 1;
  const float lr_t = state != nullptr ? state[1] : lr_t_arg;
  const int src = blockIdx.x % c.G;
  const int share = blockIdx.x / c.G, nshare = gridDim.x / c.G;
  if (share >= nshare) return;
  const int cnt = min(*reinterpret_cast<volatile int*>(iptr(c, c.me, c.off_req_cnt) + par * c.G + src), c.cap);
  const int* ids = iptr(c, c.me, c.off_req_ids) + (static_cast<long long>(par) * c.G + src) * c.cap;
  const int lane = threadIdx.x & 31, r = lane / LPR, q = lane % LPR;
  const int wpb = blockDim.x >> 5;
  const int per_iter = RPW * U;
  const int w = share * wpb + (threadIdx.x >> 5), nw = nshare * wpb;
  for (int p0 = w * per_iter; p0 < cnt; p0 += nw * per_iter) {      // warp-uniform trip count
    int id[U], won[U];
    float4 G[U], M[U], V[U], T[U], F1[U];
    float cc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pos = p0 + u * RPW + r;
      id[u] = pos < cnt ? __ldcg(ids + pos) : -1;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      won[u] = 0;
      cc[u] = 0.f;
      if (id[u] >= 0) {
        float* row = rec + static_cast<size_t>(id[u]) * ld;
        // a row that is already claimed - every repeat of a hot row - is dropped without an atomic
        if (q == 0) won[u] = __ldcg(reinterpret_cast<const int*>(row + R::CLAIM)) != seq ? 1 : 0;
        G[u] = ld4_plain(row + R::G + q * 4);
        M[u] = ld4_plain(row + R::M + q * 4);
        V[u] = ld4_plain(row + R::V + q * 4);
        T[u] = ld4_plain(row + R::TH + q * 4);
        if (q == 0 && with_w1) F1[u] = ld4_plain(row + R::TH1);
        if (has_c) cc[u] = ld1_plain(row + R::CC);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)     // candidates race for the claim; exactly one wins per row
      if (won[u])
        won[u] = atomicExch(reinterpret_cast<int*>(rec + static_cast<size_t>(id[u]) * ld + R::CLAIM), seq) != seq ? 1 : 0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int wn = __shfl_sync(0xffffffffu, won[u], lane - q);
      if (!wn) continue;
      // only the claim winner writes this row in this launch, and what it loaded above was written
      // by earlier kernels (the scatter pass): the speculative loads are race free
      float* row = rec + static_cast<size_t>(id[u]) * ld;
      float4 g = G[u], t = T[u], m = M[u], v = V[u];
      if (has_c) {
        g.x = fmaf(-cc[u], t.x, g.x); g.y = fmaf(-cc[u], t.y, g.y);
        g.z = fmaf(-cc[u], t.z, g.z); g.w = fmaf(-cc[u], t.w, g.w);
      }
      adam4(t, m, v, g, k, lr_t);
      *reinterpret_cast<float4*>(row + R::M + q * 4) = m;
      *reinterpret_cast<float4*>(row + R::V + q * 4) = v;
      *reinterpret_cast<float4*>(row + R::TH + q * 4) = t;
      *reinterpret_cast<float4*>(row + R::G + q * 4) = f4_zero();
      if (q == 0) {
        if (with_w1) *reinterpret_cast<float4*>(row + R::TH1) = adam1(F1[u], F1[u].w, k, lr_t);
        if (has_c) row[R::CC] = 0.f;
      }
    }
  }
}

// ----------------------------------------------------------------------------- K6
__global__ void __launch_bounds__(256)
p2p_dense_push_kernel(const float* __restrict__ grad, long long n, const P2P c) {
  const int seq = hdr(c, c.me)[0];
  const long long n4 = n >> 2;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldcg(reinterpret_cast<const float4*>(grad) + i);
    for (int r = 0; r < c.G; ++r)
      reinterpret_cast<float4*>(fptr(c, r, c.off_dense) + static_cast<long long>(c.me) * c.n_dense)[i] = v;
  }
  if (last_block(iptr(c, c.me, c.off_done), gridDim.x)) {
    if (threadIdx.x < c.G) st_release_sys(iptr(c, threadIdx.x, c.off_dense_flag) + c.me, seq);
  }
}

// theta -= Adam(sum_r dense[r]); zeroes the local gradient buffer; the last block advances the
// device Adam schedule (see adam_dense_kernel in embed.cu).
__device__ __forceinline__ void adam_advance_p2p(float* __restrict__ state, float b1, float b2) {
  const unsigned t = adam_step_of(state) + 1u;     // integer step counter (bit pattern of state[0])
  const float tn = static_cast<float>(t) + 1.f;
  state[0] = __uint_as_float(t);
  state[1] = state[2] * sqrtf(1.f - powf(b2, tn)) / (1.f - powf(b1, tn));
}
__global__ void __launch_bounds__(256)
p2p_adam_dense_kernel(float* __restrict__ th, float* __restrict__ m, float* __restrict__ v,
                      float* __restrict__ g_local, long long n, float lr_t, AdamK k,
                      float* __restrict__ state, int advance, const P2P c) {
  const int seq = hdr(c, c.me)[0];
  if (state != nullptr) lr_t = state[1];
  if (threadIdx.x < c.G) wait_flag(c, iptr(c, c.me, c.off_dense_flag) + threadIdx.x, seq);
  __syncthreads();
  const float* in = fptr(c, c.me, c.off_dense);
  const long long n4 = n >> 2;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4; i += stride) {
    float4 G = f4_zero();
    for (int r = 0; r < c.G; ++r)
      G = f4_add(G, __ldcg(reinterpret_cast<const float4*>(in + static_cast<long long>(r) * c.n_dense) + i));
    float4 M = reinterpret_cast<float4*>(m)[i];
    float4 V = reinterpret_cast<float4*>(v)[i];
    float4 T = reinterpret_cast<float4*>(th)[i];
#define CTR_ADAM1(x)                                   \
  M.x = k.b1 * M.x + (1.f - k.b1) * G.x;               \
  V.x = k.b2 * V.x + (1.f - k.b2) * G.x * G.x;         \
  T.x -= lr_t * M.x / (sqrtf(V.x) + k.eps);
    CTR_ADAM1(x) CTR_ADAM1(y) CTR_ADAM1(z) CTR_ADAM1(w)
#undef CTR_ADAM1
    reinterpret_cast<float4*>(m)[i] = M;
    reinterpret_cast<float4*>(v)[i] = V;
    reinterpret_cast<float4*>(th)[i] = T;
    reinterpret_cast<float4*>(g_local)[i] = f4_zero();
  }
  if (advance && state != nullptr) {
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned* cntr = reinterpret_cast<unsigned*>(state + 3);
      __threadfence();
      if (atomicAdd(cntr, 1u) == gridDim.x - 1) {
        *cntr = 0u;
        adam_advance_p2p(state, k.b1, k.b2);
      }
    }
  }
}

static P2P make_ctx(const ctr_p2p_ctx* x) {
  P2P c{};
  for (int r = 0; r < CTR_P2P_MAX_RANKS; ++r) c.peer[r] = static_cast<char*>(x->peer[r]);
  c.me = x->me; c.G = x->G; c.cap = x->capacity; c.P = x->record_floats;
  c.off_req_flag = x->off_req_flag; c.off_req_cnt = x->off_req_cnt; c.off_resp_flag = x->off_resp_flag;
  c.off_grad_flag = x->off_grad_flag; c.off_dense_flag = x->off_dense_flag;
  c.off_req_ids = x->off_req_ids; c.off_resp = x->off_resp; c.off_grad = x->off_grad;
  c.off_dense = x->off_dense; c.off_counts = x->off_counts; c.off_done = x->off_done;
  c.off_inv = x->off_inv; c.off_sent = x->off_sent;
  c.n_dense = x->n_dense;
  c.spin_limit_ns = x->spin_limit_ms > 0 ? static_cast<long long>(x->spin_limit_ms) * 1000000LL
                                         : 10000000000LL;
  return c;
}
static bool ctx_ok(const ctr_p2p_ctx* x) {
  if (x == nullptr || x->G < 1 || x->G > CTR_P2P_MAX_RANKS || x->me < 0 || x->me >= x->G) return false;
  for (int r = 0; r < x->G; ++r)
    if (x->peer[r] == nullptr) return false;
  return x->capacity > 0 && x->record_floats >= 8;
}

}  // namespace ctr

using namespace ctr;

extern "C" {

int ctr_p2p_alloc(int64_t bytes, void** ptr, void* ipc_handle_out) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(bytes > 0 && ptr && ipc_handle_out, "ctr_p2p_alloc", "bad argument");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, static_cast<size_t>(bytes));
  if (e != cudaSuccess) return check_cuda(e, "ctr_p2p_alloc");
  e = cudaMemset(p, 0, static_cast<size_t>(bytes));
  if (e == cudaSuccess) {
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, p);
    if (e == cudaSuccess) memcpy(ipc_handle_out, &h, sizeof(h));
  }
  if (e != cudaSuccess) {
    cudaFree(p);
    return check_cuda(e, "ctr_p2p_alloc");
  }
  *ptr = p;
  return CTR_OK;
}

int ctr_p2p_open(const void* ipc_handle, void** ptr) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(ipc_handle && ptr, "ctr_p2p_open", "bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, sizeof(h));
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return check_cuda(e, "ctr_p2p_open");
  *ptr = p;
  return CTR_OK;
}

int ctr_p2p_close(void* ptr) {
  if (ptr == nullptr) return CTR_OK;
  return check_cuda(cudaIpcCloseMemHandle(ptr), "ctr_p2p_close");
}

int ctr_p2p_free(void* ptr) {
  if (ptr == nullptr) return CTR_OK;
  return check_cuda(cudaFree(ptr), "ctr_p2p_free");
}

int ctr_p2p_bucket_send(const int32_t* rows, int64_t n, const ctr_p2p_ctx* ctx, int32_t* slot,
                        ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(rows && slot && n >= 0 && ctx_ok(ctx), "ctr_p2p_bucket_send", "bad argument");
  CTR_REQUIRE(n <= ctx->capacity, "ctr_p2p_bucket_send",
              "capacity must cover all lookups of a rank (worst case: one owner)");
  const long long per_cta = 256LL * kBktPerThread;
  const int grid = static_cast<int>(std::max<long long>(1, (n + per_cta - 1) / per_cta));
  p2p_bucket_send_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(rows, n, make_ctx(ctx), slot);
  CTR_LAUNCH_CHECK("ctr_p2p_bucket_send");
}

int ctr_p2p_gather_reply(float* rec, int64_t row_stride, int D, int with_w1, int count_lookups,
                         const ctr_p2p_ctx* ctx, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(rec && ctx_ok(ctx), "ctr_p2p_gather_reply", "bad argument");
  CTR_REQUIRE((D == 8 || D == 16 || D == 32) && row_stride >= 4 * D + 8 && ctx->record_floats == D + 4,
              "ctr_p2p_gather_reply", "needs the row-record layout and record_floats == D + 4");
  const int G = ctx->G;
  const int grid = std::max(1, (sm_count() * 4) / G) * G;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const P2P c = make_ctx(ctx);
  switch (D) {
    case 8: p2p_gather_reply_kernel<8><<<grid, 256, 0, st>>>(rec, row_stride, with_w1, count_lookups, c); break;
    case 16: p2p_gather_reply_kernel<16><<<grid, 256, 0, st>>>(rec, row_stride, with_w1, count_lookups, c); break;
    default: p2p_gather_reply_kernel<32><<<grid, 256, 0, st>>>(rec, row_stride, with_w1, count_lookups, c); break;
  }
  CTR_LAUNCH_CHECK("ctr_p2p_gather_reply");
}

int ctr_p2p_wait(const ctr_p2p_ctx* ctx, int what, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(ctx_ok(ctx) && what >= 0 && what <= 2, "ctr_p2p_wait", "bad argument");
  const long long off = what == 0 ? ctx->off_resp_flag : what == 1 ? ctx->off_grad_flag : ctx->off_dense_flag;
  p2p_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(make_ctx(ctx), off);
  CTR_LAUNCH_CHECK("ctr_p2p_wait");
}

int ctr_p2p_grad_send(const int32_t* slot, const float* dE, const float* S, const float* dy2,
                      const float* dy1, uint64_t w1_fields, int B, int F, int D,
                      const ctr_p2p_ctx* ctx, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  (void)slot;       // the slab order comes from the arena's inv[] (written by ctr_p2p_bucket_send)
  CTR_REQUIRE((dE || dy2) && B >= 0 && F > 0 && F <= CTR_MAX_FIELDS && ctx_ok(ctx),
              "ctr_p2p_grad_send", "bad argument");
  CTR_REQUIRE(!dy2 || S, "ctr_p2p_grad_send", "dy2 needs S");
  CTR_REQUIRE(aligned16(dE) && aligned16(S), "ctr_p2p_grad_send", "pointers must be 16-byte aligned");
  CTR_REQUIRE((D == 8 || D == 16 || D == 32) && ctx->record_floats == D + 4, "ctr_p2p_grad_send",
              "D in {8,16,32}, record_floats == D + 4");
  const int G = ctx->G;
  const int grid = std::max(1, (sm_count() * 4) / G) * G;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const P2P c = make_ctx(ctx);
  switch (D) {
    case 8: p2p_grad_send_kernel<8><<<grid, 256, 0, st>>>(dE, S, dy2, dy1, w1_fields, F, c); break;
    case 16: p2p_grad_send_kernel<16><<<grid, 256, 0, st>>>(dE, S, dy2, dy1, w1_fields, F, c); break;
    default: p2p_grad_send_kernel<32><<<grid, 256, 0, st>>>(dE, S, dy2, dy1, w1_fields, F, c); break;
  }
  CTR_LAUNCH_CHECK("ctr_p2p_grad_send");
}

int ctr_p2p_scatter_adam(float* rec, int64_t row_stride, int D, int with_w1, int has_c, float lr_t,
                         float beta1, float beta2, float eps, const float* state_dev,
                         const ctr_p2p_ctx* ctx, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(rec && ctx_ok(ctx), "ctr_p2p_scatter_adam", "bad argument");
  CTR_REQUIRE((D == 8 || D == 16 || D == 32) && row_stride >= 4 * D + 8 && ctx->record_floats == D + 4,
              "ctr_p2p_scatter_adam", "needs the row-record layout and record_floats == D + 4");
  const int G = ctx->G;
  const int grid = std::max(1, (sm_count() * 6) / G) * G;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const P2P c = make_ctx(ctx);
  const AdamK k{beta1, beta2, eps};
#define CTR_K5(DD)                                                                              \
  p2p_scatter_kernel<DD><<<grid, 256, 0, st>>>(rec, row_stride, with_w1, has_c, c);              \
  p2p_adam_rows_kernel<DD><<<grid, 256, 0, st>>>(rec, row_stride, with_w1, has_c, state_dev, lr_t, k, c);
  switch (D) {
    case 8: CTR_K5(8) break;
    case 16: CTR_K5(16) break;
    default: CTR_K5(32) break;
  }
#undef CTR_K5
  CTR_LAUNCH_CHECK("ctr_p2p_scatter_adam");
}

int ctr_p2p_dense_push(const float* grad, int64_t n, const ctr_p2p_ctx* ctx, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(grad && n > 0 && (n & 3) == 0 && ctx_ok(ctx) && n <= ctx->n_dense && aligned16(grad),
              "ctr_p2p_dense_push", "bad argument (n % 4 == 0, n <= ctx.n_dense)");
  const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>((n / 4 + 255) / 256, sm_count())));
  p2p_dense_push_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(grad, n, make_ctx(ctx));
  CTR_LAUNCH_CHECK("ctr_p2p_dense_push");
}

int ctr_p2p_adam_dense(float* theta, float* m, float* v, float* g_local, int64_t n, float lr_t,
                       float beta1, float beta2, float eps, float* state_dev, int advance_state,
                       const ctr_p2p_ctx* ctx, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(theta && m && v && g_local && n > 0 && (n & 3) == 0 && ctx_ok(ctx) && n <= ctx->n_dense,
              "ctr_p2p_adam_dense", "bad argument (n % 4 == 0, n <= ctx.n_dense)");
  const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>((n / 4 + 255) / 256, sm_count() * 4LL)));
  p2p_adam_dense_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      theta, m, v, g_local, n, lr_t, AdamK{beta1, beta2, eps}, state_dev, advance_state, make_ctx(ctx));
  CTR_LAUNCH_CHECK("ctr_p2p_adam_dense");
}

int ctr_p2p_status(const ctr_p2p_ctx* ctx, int32_t* step_out, int32_t* err_out) {
  CTR_REQUIRE(ctx_ok(ctx) && step_out && err_out, "ctr_p2p_status", "bad argument");
  int h[2] = {0, 0};
  cudaError_t e = cudaMemcpy(h, ctx->peer[ctx->me], sizeof(h), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return check_cuda(e, "ctr_p2p_status");
  *step_out = h[0];
  *err_out = h[1];
  return CTR_OK;
}

}  // extern "C"
