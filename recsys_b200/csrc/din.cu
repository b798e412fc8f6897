// DIN activation unit (din/din.py:103-125 `_attention`), fused: gather the history
// rows, run the 4E->80->40->1 attention MLP per position in registers, masked
// weighted sum - one warp per sample, one lane per history position.
//
// Algebra used to shrink layer 1 (a = [h, q, h*q, h-q], W1 = [W1a; W1b; W1c; W1d]):
//   a.W1 = h.(W1a + W1d) + q.(W1b - W1d) + (h*q).W1c = cq + h.Weff,
//   cq = b1 + q.Wq,  Weff[e][k] = Wh[e][k] + q[e] * Wp[e][k]   (per sample, in shared memory)
// which takes the per-position cost from 5120+3200 to 1280+3200 MACs.
//
// Backward = kernel 1 (same structure): recompute, d(out) -> dw, dh2, dh1, dh (RED into the
// table gradient), dq; it also leaves h1/dh1/dh2/h/h*q rows in scratch so that kernel 2
// (`xtx`: C += A^T B over the valid rows) forms dW1/dW2 as tall-skinny reductions.
// FP32 CUDA-core work, compute-bound: padding positions are skipped.
#include "gemm_core.cuh"
#include "tower_common.cuh"

namespace ctr {

constexpr int kH1 = 80, kH2 = 40;   // attention_layers = [80, 40] is hard-coded in din/din.py:85

struct DinParams {
  const float* table;
  const int* hist;
  const float* query;
  const float* W1;
  const float* b1;
  const float* W2;
  const float* b2;
  const float* W3;
  const float* b3;
  int B, P;
  // forward
  float* out;
  float* att_w;
  // backward
  const float* dout;
  float* dtable;
  float* dquery;
  float* sH1;    // [B*P, 80] relu(h1)
  float* sdH1;   // [B*P, 80]
  float* sdH2;   // [B*P, 40]
  float* sHh;    // [B*P, E]
  float* sHQ;    // [B*P, E]
  float* sSd;    // [B, 80]  sum_p dh1
  const int* rowbase;   // [B+1] exclusive prefix of valid positions per sample (compact rows)
  float* dW3;
  float* db3;
  // dropout after each of the two hidden attention layers (din/din.py:118)
  const float* dstate;      // device Adam schedule: [0] = t selects the dropout stream (nullable -> 0)
  float inv_keep;
  unsigned thr;             // keep iff a 16-bit uniform >= thr = p * 65536
  unsigned seed, unit;      // unit: which attention unit (0 = att_iid, 1 = att_cat)
  int n_rows;               // rows of `table`: ids outside [0, n_rows) are treated as padding ...
  int* status;              // ... and set bit 1 of *status (nullable)
};

__device__ __forceinline__ bool din_valid_id(const DinParams& p, int id) {
  if (id > 0 && id < p.n_rows) return true;      // din/din.py:107 mask: id > 0
  if (id != 0 && p.status != nullptr) atomicOr(p.status, 2);
  return false;
}

// Keep bits of the 8 consecutive columns [8*blk, 8*blk+8) of attention layer `layer` (0: the 80-wide
// one, 1: the 40-wide one) at history position `row` = b*P + pos: one Philox4x32-10 block =
// 8 x 16-bit uniforms, counter (row, blk, 0x100 + 2*unit + layer, step), key (seed, 0xD1A7).  The
// backward regenerates the same bits; nothing is stored.
__device__ __forceinline__ unsigned din_keep8(const DinParams& p, unsigned step, unsigned row,
                                              unsigned layer, unsigned blk) {
  const uint4 r = philox4x32_10(make_uint4(row, blk, 0x100u + 2u * p.unit + layer, step),
                                make_uint2(p.seed, 0xD1A7u));
  unsigned bits = 0u;
  bits |= (r.x & 0xFFFFu) >= p.thr ? 1u : 0u;
  bits |= (r.x >> 16) >= p.thr ? 2u : 0u;
  bits |= (r.y & 0xFFFFu) >= p.thr ? 4u : 0u;
  bits |= (r.y >> 16) >= p.thr ? 8u : 0u;
  bits |= (r.z & 0xFFFFu) >= p.thr ? 16u : 0u;
  bits |= (r.z >> 16) >= p.thr ? 32u : 0u;
  bits |= (r.w & 0xFFFFu) >= p.thr ? 64u : 0u;
  bits |= (r.w >> 16) >= p.thr ? 128u : 0u;
  return bits;
}

template <int E>
struct DinSmem {
  float W2[kH1 * kH2];
  float Wh[kH1 * E];     // [k][e] = W1a + W1d
  float Wp[kH1 * E];     // [k][e] = W1c
  float Wq[kH1 * E];     // [k][e] = W1b - W1d
  float b1[kH1];
  float b2[kH2];
  float W3[kH2];
  float Weff[8][kH1 * E];   // per warp
  float cq[8][kH1];
  float sd[8][kH1];         // per warp: sum_p dh1 (backward)
};

template <int E>
__device__ __forceinline__ void din_load_weights(DinSmem<E>& s, const DinParams& p) {
  for (int i = threadIdx.x; i < kH1 * kH2; i += blockDim.x) s.W2[i] = p.W2[i];
  for (int i = threadIdx.x; i < kH1 * E; i += blockDim.x) {
    const int k = i / E, e = i % E;
    const float a = p.W1[(0 * E + e) * kH1 + k], b = p.W1[(1 * E + e) * kH1 + k];
    const float c = p.W1[(2 * E + e) * kH1 + k], d = p.W1[(3 * E + e) * kH1 + k];
    s.Wh[i] = a + d;
    s.Wp[i] = c;
    s.Wq[i] = b - d;
  }
  for (int i = threadIdx.x; i < kH1; i += blockDim.x) s.b1[i] = p.b1[i];
  for (int i = threadIdx.x; i < kH2; i += blockDim.x) {
    s.b2[i] = p.b2[i];
    s.W3[i] = p.W3[i];
  }
}

// per-sample layer-1 folding, by the whole warp
template <int E>
__device__ __forceinline__ void din_fold_sample(DinSmem<E>& s, int warp, int lane, const float* q) {
  for (int i = lane; i < kH1 * E; i += 32) s.Weff[warp][i] = fmaf(q[i % E], s.Wp[i], s.Wh[i]);
  for (int k = lane; k < kH1; k += 32) {
    float a = s.b1[k];
#pragma unroll
    for (int e = 0; e < E; ++e) a = fmaf(q[e], s.Wq[k * E + e], a);
    s.cq[warp][k] = a;
  }
}

template <int E, bool DROP>
__global__ void __launch_bounds__(256) din_att_fwd_kernel(const DinParams p) {
  extern __shared__ __align__(16) uint8_t din_smem_raw[];
  DinSmem<E>& s = *reinterpret_cast<DinSmem<E>*>(din_smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  din_load_weights<E>(s, p);
  __syncthreads();
  const float b3 = p.b3[0];
  const unsigned step = (DROP && p.dstate != nullptr) ? adam_step_of(p.dstate) : 0u;

  for (int b = blockIdx.x * 8 + warp; b < p.B; b += gridDim.x * 8) {
    float q[E];
#pragma unroll
    for (int e = 0; e < E; e += 4) {
      const float4 t = ldg4(p.query + static_cast<size_t>(b) * E + e);
      q[e] = t.x; q[e + 1] = t.y; q[e + 2] = t.z; q[e + 3] = t.w;
    }
    __syncwarp();
    din_fold_sample<E>(s, warp, lane, q);
    __syncwarp();
    float o[E];
#pragma unroll
    for (int e = 0; e < E; ++e) o[e] = 0.f;

    for (int p0 = 0; p0 < p.P; p0 += 32) {
      const int pos = p0 + lane;
      const int id = pos < p.P ? __ldg(p.hist + static_cast<size_t>(b) * p.P + pos) : 0;
      const bool valid = din_valid_id(p, id);
      if (__ballot_sync(0xffffffffu, valid) == 0u) continue;
      float w = 0.f;
      if (valid) {
        float h[E];
#pragma unroll
        for (int e = 0; e < E; e += 4) {
          const float4 t = ldg4(p.table + static_cast<size_t>(id) * E + e);
          h[e] = t.x; h[e + 1] = t.y; h[e + 2] = t.z; h[e + 3] = t.w;
        }
        float h2[kH2];
#pragma unroll
        for (int j = 0; j < kH2; ++j) h2[j] = s.b2[j];
        const unsigned row = static_cast<unsigned>(b) * p.P + pos;
        for (int kb = 0; kb < kH1 / 8; ++kb) {
          const unsigned keep = DROP ? din_keep8(p, step, row, 0u, kb) : 0xFFu;
#pragma unroll 2
          for (int u = 0; u < 8; ++u) {
            const int k = kb * 8 + u;
            float a = s.cq[warp][k];
            const float* we = &s.Weff[warp][k * E];
#pragma unroll
            for (int e = 0; e < E; ++e) a = fmaf(h[e], we[e], a);
            a = fmaxf(a, 0.f);
            if (DROP) a = ((keep >> u) & 1u) ? a * p.inv_keep : 0.f;
            const float* w2 = &s.W2[k * kH2];
#pragma unroll
            for (int j = 0; j < kH2; ++j) h2[j] = fmaf(a, w2[j], h2[j]);
          }
        }
        w = b3;
#pragma unroll
        for (int jb = 0; jb < kH2 / 8; ++jb) {
          const unsigned keep = DROP ? din_keep8(p, step, row, 1u, jb) : 0xFFu;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            float r = fmaxf(h2[jb * 8 + u], 0.f);
            if (DROP) r = ((keep >> u) & 1u) ? r * p.inv_keep : 0.f;
            w = fmaf(r, s.W3[jb * 8 + u], w);
          }
        }
#pragma unroll
        for (int e = 0; e < E; ++e) o[e] = fmaf(w, h[e], o[e]);
      }
      if (p.att_w != nullptr && pos < p.P) p.att_w[static_cast<size_t>(b) * p.P + pos] = w;
    }
#pragma unroll
    for (int e = 0; e < E; ++e) o[e] = warp_sum(o[e]);
    if (lane == 0) {
#pragma unroll
      for (int e = 0; e < E; e += 4)
        *reinterpret_cast<float4*>(p.out + static_cast<size_t>(b) * E + e) =
            make_float4(o[e], o[e + 1], o[e + 2], o[e + 3]);
    }
  }
}

template <int E, bool DROP>
__global__ void __launch_bounds__(256) din_att_bwd_kernel(const DinParams p) {
  extern __shared__ __align__(16) uint8_t din_smem_raw[];
  DinSmem<E>& s = *reinterpret_cast<DinSmem<E>*>(din_smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  din_load_weights<E>(s, p);
  __syncthreads();
  const float b3 = p.b3[0];
  float dW3acc[kH2];
#pragma unroll
  for (int j = 0; j < kH2; ++j) dW3acc[j] = 0.f;
  float db3acc = 0.f;
  const unsigned step = (DROP && p.dstate != nullptr) ? adam_step_of(p.dstate) : 0u;

  for (int b = blockIdx.x * 8 + warp; b < p.B; b += gridDim.x * 8) {
    float q[E], g[E], dq[E];
#pragma unroll
    for (int e = 0; e < E; e += 4) {
      const float4 t = ldg4(p.query + static_cast<size_t>(b) * E + e);
      q[e] = t.x; q[e + 1] = t.y; q[e + 2] = t.z; q[e + 3] = t.w;
      const float4 u = ldg4(p.dout + static_cast<size_t>(b) * E + e);
      g[e] = u.x; g[e + 1] = u.y; g[e + 2] = u.z; g[e + 3] = u.w;
    }
#pragma unroll
    for (int e = 0; e < E; ++e) dq[e] = 0.f;
    __syncwarp();
    din_fold_sample<E>(s, warp, lane, q);
    for (int k = lane; k < kH1; k += 32) s.sd[warp][k] = 0.f;
    __syncwarp();

    int row_cursor = __ldg(p.rowbase + b);
    for (int p0 = 0; p0 < p.P; p0 += 32) {
      const int pos = p0 + lane;
      const int id = pos < p.P ? __ldg(p.hist + static_cast<size_t>(b) * p.P + pos) : 0;
      const bool valid = din_valid_id(p, id);
      const unsigned vmask = __ballot_sync(0xffffffffu, valid);
      if (vmask == 0u) continue;
      // compact scratch row of this position: valid positions only, in (sample, position) order
      const size_t n = static_cast<size_t>(row_cursor + __popc(vmask & ((1u << lane) - 1u)));
      row_cursor += __popc(vmask);
      float h[E], dh[E], tq[E];
      float h2[kH2];
      unsigned m1a = 0u, m1b = 0u, m1c = 0u;   // relu (and dropout keep) mask of h1 (80 bits)
      unsigned long long keep2 = ~0ull;        // dropout keep bits of h2 (40 bits)
      const unsigned row = static_cast<unsigned>(b) * p.P + pos;
      float w = 0.f, dw = 0.f;
      if (valid) {
#pragma unroll
        for (int e = 0; e < E; e += 4) {
          const float4 t = ldg4(p.table + static_cast<size_t>(id) * E + e);
          h[e] = t.x; h[e + 1] = t.y; h[e + 2] = t.z; h[e + 3] = t.w;
        }
#pragma unroll
        for (int j = 0; j < kH2; ++j) h2[j] = s.b2[j];
        float* oh1 = p.sH1 + n * kH1;
        unsigned keep8 = 0xFFu;
        for (int k4 = 0; k4 < kH1; k4 += 4) {
          float a4[4];
          if (DROP && (k4 & 4) == 0) keep8 = din_keep8(p, step, row, 0u, k4 >> 3);
          const unsigned keep = (keep8 >> (k4 & 4)) & 0xFu;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int k = k4 + u;
            float a = s.cq[warp][k];
            const float* we = &s.Weff[warp][k * E];
#pragma unroll
            for (int e = 0; e < E; ++e) a = fmaf(h[e], we[e], a);
            a = fmaxf(a, 0.f);
            if (DROP) a = ((keep >> u) & 1u) ? a * p.inv_keep : 0.f;
            a4[u] = a;
            const float* w2 = &s.W2[k * kH2];
#pragma unroll
            for (int j = 0; j < kH2; ++j) h2[j] = fmaf(a, w2[j], h2[j]);
          }
          const unsigned bits = (a4[0] > 0.f ? 1u : 0u) | (a4[1] > 0.f ? 2u : 0u) |
                                (a4[2] > 0.f ? 4u : 0u) | (a4[3] > 0.f ? 8u : 0u);
          if (k4 < 32) m1a |= bits << k4;
          else if (k4 < 64) m1b |= bits << (k4 - 32);
          else m1c |= bits << (k4 - 64);
          *reinterpret_cast<float4*>(oh1 + k4) = make_float4(a4[0], a4[1], a4[2], a4[3]);
        }
        if (DROP) {
          keep2 = 0ull;
#pragma unroll
          for (int jb = 0; jb < kH2 / 8; ++jb)
            keep2 |= static_cast<unsigned long long>(din_keep8(p, step, row, 1u, jb)) << (8 * jb);
        }
        const float ik = DROP ? p.inv_keep : 1.f;
        w = b3;
#pragma unroll
        for (int j = 0; j < kH2; ++j)
          w = fmaf(((keep2 >> j) & 1ull) ? fmaxf(h2[j], 0.f) * ik : 0.f, s.W3[j], w);
        // out = sum_p w h  ->  dw = g . h
#pragma unroll
        for (int e = 0; e < E; ++e) dw = fmaf(g[e], h[e], dw);
        db3acc += dw;
        float* odh2 = p.sdH2 + n * kH2;
#pragma unroll
        for (int j = 0; j < kH2; ++j) {
          const bool kp = (keep2 >> j) & 1ull;
          const float r = kp ? fmaxf(h2[j], 0.f) * ik : 0.f;
          dW3acc[j] = fmaf(dw, r, dW3acc[j]);
          h2[j] = (h2[j] > 0.f && kp) ? dw * s.W3[j] * ik : 0.f;    // h2 now holds dh2
        }
#pragma unroll
        for (int j = 0; j < kH2; j += 4)
          *reinterpret_cast<float4*>(odh2 + j) = make_float4(h2[j], h2[j + 1], h2[j + 2], h2[j + 3]);
#pragma unroll
        for (int e = 0; e < E; ++e) {
          dh[e] = w * g[e];
          tq[e] = 0.f;
        }
      }
      // second sweep over k: dh1, dh, dq (Wp part) and the warp-wide sum of dh1
      float* odh1 = p.sdH1 + n * kH1;
      for (int k4 = 0; k4 < kH1; k4 += 4) {
        const unsigned mw = k4 < 32 ? (m1a >> k4) : k4 < 64 ? (m1b >> (k4 - 32)) : (m1c >> (k4 - 64));
        float d4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int k = k4 + u;
          float d1 = 0.f;
          if (valid && ((mw >> u) & 1u)) {
            const float* w2 = &s.W2[k * kH2];
#pragma unroll
            for (int j = 0; j < kH2; ++j) d1 = fmaf(h2[j], w2[j], d1);
            if (DROP) d1 *= p.inv_keep;      // the mask bit already includes the keep decision
            const float* we = &s.Weff[warp][k * E];
            const float* wp = &s.Wp[k * E];
#pragma unroll
            for (int e = 0; e < E; ++e) {
              dh[e] = fmaf(d1, we[e], dh[e]);
              tq[e] = fmaf(d1, wp[e], tq[e]);
            }
          }
          d4[u] = d1;
          const float tot = warp_sum(d1);
          if (lane == 0) s.sd[warp][k] += tot;
        }
        if (valid) *reinterpret_cast<float4*>(odh1 + k4) = make_float4(d4[0], d4[1], d4[2], d4[3]);
      }
      if (valid) {
#pragma unroll
        for (int e = 0; e < E; ++e) dq[e] = fmaf(h[e], tq[e], dq[e]);
      }
      if (valid) {
        float* oh = p.sHh + n * E;
        float* ohq = p.sHQ + n * E;
#pragma unroll
        for (int e = 0; e < E; e += 4) {
          *reinterpret_cast<float4*>(oh + e) = make_float4(h[e], h[e + 1], h[e + 2], h[e + 3]);
          *reinterpret_cast<float4*>(ohq + e) =
              make_float4(h[e] * q[e], h[e + 1] * q[e + 1], h[e + 2] * q[e + 2], h[e + 3] * q[e + 3]);
          red_add_v4(p.dtable + static_cast<size_t>(id) * E + e,
                     make_float4(dh[e], dh[e + 1], dh[e + 2], dh[e + 3]));
        }
      }
    }
    __syncwarp();
    // dq += sum_k sd[k] * Wq[k][:]; write per-sample sums
#pragma unroll
    for (int e = 0; e < E; ++e) dq[e] = warp_sum(dq[e]);
    for (int k = lane; k < kH1; k += 32) p.sSd[static_cast<size_t>(b) * kH1 + k] = s.sd[warp][k];
    if (lane < E) {
      float a = 0.f;
      for (int k = 0; k < kH1; ++k) a = fmaf(s.sd[warp][k], s.Wq[k * E + lane], a);
      float mine = 0.f;
#pragma unroll
      for (int e = 0; e < E; ++e)
        if (e == lane) mine = dq[e];
      p.dquery[static_cast<size_t>(b) * E + lane] = a + mine;
    }
    __syncwarp();
  }
#pragma unroll
  for (int j = 0; j < kH2; ++j) {
    const float t = warp_sum(dW3acc[j]);
    if (lane == 0 && t != 0.f) red_add_f32(p.dW3 + j, t);
  }
  const float t3 = warp_sum(db3acc);
  if (lane == 0 && t3 != 0.f) red_add_f32(p.db3, t3);
}

// counts[b] = number of valid (id > 0) history positions of sample b: one warp per sample.
__global__ void __launch_bounds__(256)
din_count_kernel(const int* __restrict__ hist, int B, int P, int n_rows, int* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  for (int b = blockIdx.x * 8 + (threadIdx.x >> 5); b < B; b += gridDim.x * 8) {
    int c = 0;
    for (int p = lane; p < P; p += 32) {
      const int id = __ldg(hist + static_cast<size_t>(b) * P + p);
      c += (id > 0 && id < n_rows) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) counts[b + 1] = c;
  }
}
// in place: rowbase[0] = 0, rowbase[b+1] = sum_{i <= b} counts[i]  (single CTA, B <= 1M).
__global__ void __launch_bounds__(1024)
din_scan_kernel(int* __restrict__ rowbase, int B) {
  __shared__ int s_part[1024];
  const int tid = threadIdx.x;
  const int per = (B + 1023) / 1024;
  int cnt = 0;
  for (int i = 0; i < per; ++i) {
    const int b = tid * per + i;
    if (b < B) cnt += rowbase[b + 1];
  }
  s_part[tid] = cnt;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int v = tid >= o ? s_part[tid - o] : 0;
    __syncthreads();
    s_part[tid] += v;
    __syncthreads();
  }
  int run = s_part[tid] - cnt;
  for (int i = 0; i < per; ++i) {
    const int b = tid * per + i;
    if (b < B) {
      run += rowbase[b + 1];
      rowbase[b + 1] = run;
    }
  }
  if (tid == 0) rowbase[0] = 0;
}

// C[a, c] += sum_{n < N} A[n, a] * Bm[n, c]   (N read from *n_dev when given: compact rows).
// colsum (nullable): colsum[c] += sum_n Bm[n, c].  Tile 16|32 (a) x 128 (c), rows split over
// gridDim.z, 3xTF32 tensor-core tile core (gemm_core.cuh).
template <int MT>
__global__ void __launch_bounds__(256)
xtx_kernel(const float* __restrict__ A, int lda, int Ka, const float* __restrict__ Bm, int ldb,
           int Kb, const int* __restrict__ n_dev, long long N, float* __restrict__ C, int ldc,
           float* __restrict__ colsum, long long rows_per_split) {
  extern __shared__ __align__(16) uint8_t xtx_smem[];
  MmaSmem& sm = *reinterpret_cast<MmaSmem*>(xtx_smem);
  __shared__ float s_red[2][kTwBN];
  if (n_dev != nullptr) {
    // the host sized the grid for the upper bound B * P; share out the rows that exist (about
    // half of them at the benchmark's history lengths), so that no split is left idle
    N = min(N, static_cast<long long>(*n_dev));
    rows_per_split = ((N + gridDim.z - 1) / gridDim.z + kTwKC - 1) / kTwKC * kTwKC;
  }
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int a0 = blockIdx.x * (MT * 16), c0 = blockIdx.y * kTwBN;
  const long long rbeg = blockIdx.z * rows_per_split;
  const long long rend = min(N, rbeg + rows_per_split);
  if (rbeg >= rend) return;
  const int nrows = static_cast<int>(rend - rbeg);
  float acc[MT][2][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;
  const bool aal = (lda & 3) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0;
  const bool bal = (ldb & 3) == 0 && (reinterpret_cast<uintptr_t>(Bm) & 15) == 0;
  auto fa4 = [&](int m, int rr) -> float4 {         // 4 consecutive a of row rbeg + rr
    const int a = a0 + m;
    if (a >= Ka || rr >= nrows) return f4_zero();
    return load4_guard(A + (rbeg + rr) * lda + a, Ka - a, aal);
  };
  auto fb4 = [&](int rr, int c) -> float4 {
    const int cc = c0 + c;
    if (cc >= Kb || rr >= nrows) return f4_zero();
    return load4_guard(Bm + (rbeg + rr) * ldb + cc, Kb - cc, bal);
  };
  gemm_tile_mma<MT, false, false>(sm, nrows, fa4, fb4, acc, Kb - c0);
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        const int a = a0 + mt * 16 + g + (ci >> 1) * 8;
        const int c = c0 + warp * 16 + nt * 8 + 2 * t + (ci & 1);
        if (a < Ka && c < Kb && acc[mt][nt][ci] != 0.f)
          red_add_f32(C + static_cast<size_t>(a) * ldc + c, acc[mt][nt][ci]);
      }
  if (colsum != nullptr) {   // this split's rows shared out over blockIdx.x
    const long long per = (nrows + gridDim.x - 1) / gridDim.x;
    const long long rb = rbeg + blockIdx.x * per, re = min(rend, rb + per);
    const int c = tid & 127, half = tid >> 7;
    const int cc = c0 + c;
    float sacc = 0.f;
    if (cc < Kb) {
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      long long r = rb + half;
      for (; r + 6 < re; r += 8) {
        s0 += Bm[r * ldb + cc];
        s1 += Bm[(r + 2) * ldb + cc];
        s2 += Bm[(r + 4) * ldb + cc];
        s3 += Bm[(r + 6) * ldb + cc];
      }
      for (; r < re; r += 2) s0 += Bm[r * ldb + cc];
      sacc = (s0 + s1) + (s2 + s3);
    }
    s_red[half][c] = sacc;
    __syncthreads();
    if (half == 0 && cc < Kb) {
      const float v = s_red[0][c] + s_red[1][c];
      if (v != 0.f) red_add_f32(colsum + cc, v);
    }
  }
}

// dW1[4E, 80] += [dWh; dWq; dWp; dWh - dWq] from the three [E, 80] partials in tmp.
__global__ void din_assemble_dw1_kernel(const float* __restrict__ tmp, int E, float* __restrict__ dW1) {
  const int n = E * kH1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float wh = tmp[i], wp = tmp[n + i], wq = tmp[2 * n + i];
    dW1[0 * n + i] += wh;
    dW1[1 * n + i] += wq;
    dW1[2 * n + i] += wp;
    dW1[3 * n + i] += wh - wq;
  }
}

static void xtx_launch(const float* A, int lda, int Ka, const float* Bm, int ldb, int Kb,
                       const int* n_dev, long long N, float* C, int ldc, float* colsum,
                       cudaStream_t st) {
  static bool optin = false;
  if (!optin) {
    cudaFuncSetAttribute(xtx_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         static_cast<int>(sizeof(MmaSmem)));
    cudaFuncSetAttribute(xtx_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         static_cast<int>(sizeof(MmaSmem)));
    optin = true;
  }
  const int mt = Ka <= 16 ? 1 : 2;
  const int bm = mt * 16;
  const int tiles = ((Ka + bm - 1) / bm) * ((Kb + kTwBN - 1) / kTwBN);
  long long splits = std::max<long long>(1, std::min<long long>((sm_count() * 4) / tiles, (N + 255) / 256));
  long long rps = (N + splits - 1) / splits;
  rps = (rps + kTwKC - 1) / kTwKC * kTwKC;
  splits = (N + rps - 1) / rps;
  dim3 grid((Ka + bm - 1) / bm, (Kb + kTwBN - 1) / kTwBN, static_cast<unsigned>(splits));
  if (mt == 1)
    xtx_kernel<1><<<grid, 256, sizeof(MmaSmem), st>>>(A, lda, Ka, Bm, ldb, Kb, n_dev, N, C, ldc, colsum, rps);
  else
    xtx_kernel<2><<<grid, 256, sizeof(MmaSmem), st>>>(A, lda, Ka, Bm, ldb, Kb, n_dev, N, C, ldc, colsum, rps);
}

template <int E, bool DROP>
static int din_fwd_launch2(const DinParams& p, cudaStream_t st) {
  const size_t smem = sizeof(DinSmem<E>);
  cudaFuncSetAttribute(din_att_fwd_kernel<E, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       static_cast<int>(smem));
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, din_att_fwd_kernel<E, DROP>, 256, smem);
  const int grid = std::min((p.B + 7) / 8, sm_count() * std::max(1, occ));
  din_att_fwd_kernel<E, DROP><<<grid, 256, smem, st>>>(p);
  return CTR_OK;
}
template <int E>
static int din_fwd_launch(const DinParams& p, cudaStream_t st) {
  return p.thr > 0u ? din_fwd_launch2<E, true>(p, st) : din_fwd_launch2<E, false>(p, st);
}

template <int E, bool DROP>
static int din_bwd_launch2(const DinParams& p, cudaStream_t st) {
  const size_t smem = sizeof(DinSmem<E>);
  cudaFuncSetAttribute(din_att_bwd_kernel<E, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       static_cast<int>(smem));
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, din_att_bwd_kernel<E, DROP>, 256, smem);
  const int grid = std::min((p.B + 7) / 8, sm_count() * std::max(1, occ));
  din_att_bwd_kernel<E, DROP><<<grid, 256, smem, st>>>(p);
  return CTR_OK;
}
template <int E>
static int din_bwd_launch(const DinParams& p, cudaStream_t st) {
  return p.thr > 0u ? din_bwd_launch2<E, true>(p, st) : din_bwd_launch2<E, false>(p, st);
}

// keep scale (0 or 1/(1-p)) of every (position row, column) of one attention layer, for tests that
// inject the kernel's masks into the oracle.
__global__ void din_dropout_mask_kernel(const DinParams p, unsigned layer, long long n_rows, int H,
                                        float* __restrict__ out) {
  const unsigned step = p.dstate != nullptr ? adam_step_of(p.dstate) : 0u;
  const int nb = (H + 7) / 8;
  const long long n = n_rows * nb;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / nb;
    const int blk = static_cast<int>(i % nb);
    const unsigned keep = din_keep8(p, step, static_cast<unsigned>(row), layer, blk);
    for (int u = 0; u < 8 && blk * 8 + u < H; ++u)
      out[row * H + blk * 8 + u] = ((keep >> u) & 1u) ? p.inv_keep : 0.f;
  }
}

static bool din_set_opts(DinParams& p, const ctr_din_opts* d) {
  p.thr = 0u; p.inv_keep = 1.f; p.dstate = nullptr; p.seed = 0u; p.unit = 0u;
  p.n_rows = 0x7FFFFFFF; p.status = nullptr;
  if (d == nullptr) return true;
  if (d->table_rows > 0) p.n_rows = d->table_rows;
  p.status = d->status;
  if (!(d->p_drop > 0.f)) return true;
  if (!(d->p_drop < 1.f)) return false;
  p.thr = static_cast<unsigned>(d->p_drop * 65536.f);
  if (p.thr == 0u) p.thr = 1u;
  p.inv_keep = 1.f / (1.f - d->p_drop);
  p.dstate = d->state; p.seed = d->seed; p.unit = d->unit;
  return true;
}

}  // namespace ctr

using namespace ctr;

extern "C" {

int64_t ctr_din_workspace_bytes(int B, int P, int E) {
  const int64_t n = static_cast<int64_t>(B) * P;
  return (n * (kH1 + kH1 + kH2 + E + E) + static_cast<int64_t>(B) * kH1 + 3LL * E * kH1 + B + 8) * 4 + 1024;
}

int ctr_din_att_fwd(const float* table, const int32_t* hist, const float* query, int B, int P,
                    int E, const float* W1, const float* b1, int H1, const float* W2,
                    const float* b2, int H2, const float* W3, const float* b3, float* out,
                    float* att_w, const ctr_din_opts* opts, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(table && hist && query && W1 && b1 && W2 && b2 && W3 && b3 && out, "ctr_din_att_fwd",
              "null pointer");
  CTR_REQUIRE(H1 == kH1 && H2 == kH2, "ctr_din_att_fwd",
              "attention layers must be [80, 40] (din/din.py:85 hard-codes them)");
  CTR_REQUIRE(B >= 0 && P > 0, "ctr_din_att_fwd", "bad B/P");
  CTR_REQUIRE(aligned16(table) && aligned16(query) && aligned16(out), "ctr_din_att_fwd",
              "pointers must be 16-byte aligned");
  if (B == 0) return CTR_OK;
  DinParams p{};
  p.table = table; p.hist = hist; p.query = query; p.W1 = W1; p.b1 = b1; p.W2 = W2; p.b2 = b2;
  p.W3 = W3; p.b3 = b3; p.B = B; p.P = P; p.out = out; p.att_w = att_w;
  CTR_REQUIRE(din_set_opts(p, opts), "ctr_din_att_fwd", "p_drop must be < 1");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (E) {
    case 8: din_fwd_launch<8>(p, st); break;
    case 16: din_fwd_launch<16>(p, st); break;
    case 32: din_fwd_launch<32>(p, st); break;
    default: return fail_arg("ctr_din_att_fwd", "E must be 8, 16 or 32");
  }
  CTR_LAUNCH_CHECK("ctr_din_att_fwd");
}

int ctr_din_att_bwd(const float* table, const int32_t* hist, const float* query, int B, int P,
                    int E, const float* W1, const float* b1, int H1, const float* W2,
                    const float* b2, int H2, const float* W3, const float* b3, const float* dout,
                    float* dtable, float* dquery, float* dW1, float* db1, float* dW2, float* db2,
                    float* dW3, float* db3, void* workspace, int64_t workspace_bytes,
                    const ctr_din_opts* opts, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(table && hist && query && W1 && b1 && W2 && b2 && W3 && b3 && dout && dtable &&
                  dquery && dW1 && db1 && dW2 && db2 && dW3 && db3,
              "ctr_din_att_bwd", "null pointer");
  CTR_REQUIRE(H1 == kH1 && H2 == kH2, "ctr_din_att_bwd", "attention layers must be [80, 40]");
  CTR_REQUIRE(B >= 0 && P > 0 && (E == 8 || E == 16 || E == 32), "ctr_din_att_bwd", "bad B/P/E");
  CTR_REQUIRE(workspace && workspace_bytes >= ctr_din_workspace_bytes(B, P, E), "ctr_din_att_bwd",
              "workspace too small (ctr_din_workspace_bytes)");
  CTR_REQUIRE(aligned16(table) && aligned16(query) && aligned16(dout) && aligned16(dtable),
              "ctr_din_att_bwd", "pointers must be 16-byte aligned");
  if (B == 0) return CTR_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long N = static_cast<long long>(B) * P;
  float* ws = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  DinParams p{};
  p.table = table; p.hist = hist; p.query = query; p.W1 = W1; p.b1 = b1; p.W2 = W2; p.b2 = b2;
  p.W3 = W3; p.b3 = b3; p.B = B; p.P = P; p.dout = dout; p.dtable = dtable; p.dquery = dquery;
  CTR_REQUIRE(din_set_opts(p, opts), "ctr_din_att_bwd", "p_drop must be < 1");
  p.sH1 = ws;
  p.sdH1 = p.sH1 + N * kH1;
  p.sdH2 = p.sdH1 + N * kH1;
  p.sHh = p.sdH2 + N * kH2;
  p.sHQ = p.sHh + N * E;
  p.sSd = p.sHQ + N * E;
  float* tmp = p.sSd + static_cast<long long>(B) * kH1;   // [3][E][80]
  int* rowbase = reinterpret_cast<int*>(tmp + 3 * E * kH1);   // [B+1]
  p.rowbase = rowbase;
  din_count_kernel<<<std::min((B + 7) / 8, sm_count() * 8), 256, 0, st>>>(hist, B, P, p.n_rows, rowbase);
  din_scan_kernel<<<1, 1024, 0, st>>>(rowbase, B);
  p.dW3 = dW3; p.db3 = db3;
  cudaError_t e = cudaMemsetAsync(tmp, 0, sizeof(float) * 3 * E * kH1, st);
  if (e != cudaSuccess) return check_cuda(e, "ctr_din_att_bwd");
  switch (E) {
    case 8: din_bwd_launch<8>(p, st); break;
    case 16: din_bwd_launch<16>(p, st); break;
    default: din_bwd_launch<32>(p, st); break;
  }
  // weight gradients: tall-skinny reductions over the valid positions; the bias gradients are
  // the column sums of the same B operands
  const int* nvalid = rowbase + B;     // rows actually written (compact), read on the device
  xtx_launch(p.sH1, kH1, kH1, p.sdH2, kH2, kH2, nvalid, N, dW2, kH2, db2, st);             // dW2 = H1^T dH2
  xtx_launch(p.sHh, E, E, p.sdH1, kH1, kH1, nvalid, N, tmp, kH1, nullptr, st);             // dWh = H^T dH1
  xtx_launch(p.sHQ, E, E, p.sdH1, kH1, kH1, nvalid, N, tmp + E * kH1, kH1, nullptr, st);   // dWp = (H*q)^T dH1
  xtx_launch(query, E, E, p.sSd, kH1, kH1, nullptr, B, tmp + 2 * E * kH1, kH1, db1, st);   // dWq = Q^T sum dH1
  din_assemble_dw1_kernel<<<(E * kH1 + 255) / 256, 256, 0, st>>>(tmp, E, dW1);
  CTR_LAUNCH_CHECK("ctr_din_att_bwd");
}

int ctr_din_dropout_mask(const ctr_din_opts* opts, int layer, int64_t n_rows, int H, float* out,
                         ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(opts && out && (layer == 0 || layer == 1) && n_rows >= 0 && H > 0,
              "ctr_din_dropout_mask", "bad argument");
  DinParams p{};
  CTR_REQUIRE(din_set_opts(p, opts), "ctr_din_dropout_mask", "p_drop must be < 1");
  if (n_rows == 0) return CTR_OK;
  const long long n = n_rows * ((H + 7) / 8);
  const int grid = static_cast<int>(std::min<long long>((n + 255) / 256, sm_count() * 8LL));
  din_dropout_mask_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      p, static_cast<unsigned>(layer), n_rows, H, out);
  CTR_LAUNCH_CHECK("ctr_din_dropout_mask");
}

}  // extern "C"
