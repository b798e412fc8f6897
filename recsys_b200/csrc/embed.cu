// Fused multi-field embedding lookup + in-register interaction (FM second order,
// first-order sum, DCN cross stack) and its field-major scatter-add backward.
//
// Forward (sample-major): a warp owns one sample at a time.  With D floats per
// row a row is LPR = D/4 lanes x float4; one warp-wide load instruction
// fetches RPW = 32/LPR rows (512 contiguous bytes of E[b,:]), so lane `l` of
// iteration `it` holds float4 number it*32+l of the sample's concatenated
// embedding - the same layout the coalesced E store and the cross stack want.
// The tile's row ids ([TB, F] int32, contiguous) are staged into shared memory
// by a 1-D TMA bulk copy (cp.async.bulk + mbarrier), double buffered across the
// persistent CTA's tiles.
//
// Backward (field-major): a warp owns (field f, chunk of samples).  Fields with
// <= 32 rows (13 bucketised numerics + the tiny hashed ones; H3 in SURVEY.md:
// one of them puts every sample in one row) are accumulated one-hot in
// registers and flushed with a handful of vector REDs; all other fields go
// straight to red.global.add.v4.f32.
#include "criteo_ids.cuh"
#include "row_commit.cuh"

namespace ctr {

// ------------------------------------------------------------------ forward
struct EmbedFwdParams {
  const float* table;
  const float* w1;
  const int* rows;
  float* E;
  float* E_lo;      // nullable: tcg_lo(E), the pre-split lo operand of the first tower GEMM
  float* S;
  float* y1;
  float* y2;
  const float* cross_w;
  const float* cross_b;
  float* xl;
  unsigned long long w1_fields;
  // fused id stage (RAW kernels): raw features in, row ids out (rows_out feeds the backward)
  const float* xcont;
  const long long* xcat;
  const ctr_field_desc* fields;
  const float* bnd;
  int* rows_out;
  float* logx;
  int* status;
  float* zero_buf;        // nullable: a buffer this launch also clears (the tower's per-step sums)
  long long zero_n4;      // its length in float4
  int n_cont, n_cat, n_bnd;
  long long ld;     // floats between consecutive table rows (D: planar table; 4D+4: row records)
  long long ld1;    // floats between consecutive first-order weights (1 or the record stride)
  int cross_layers;
  int B;
  int F;
  // peer-memory exchange (ctr_embed_fwd_p2p): flags to wait for before the first table read -
  // every owner's reply flag must have reached the step number *wait_step
  const int* wait_flags;
  const int* wait_step;
  int* wait_err;
  long long wait_ns;
  int wait_n;
  // row records: bytes of the record behind the embedding row (m | v | g | first-order group) to
  // pull into L2 while the tower runs - the scatter-add and the row optimiser of the same step
  // then find their records in L2 instead of paying a DRAM page each (0 = off)
  int prefetch_bytes;
};

constexpr int kFwdWarps = 8;
constexpr int kFwdSPW = 2;                      // samples per warp per tile
constexpr int kFwdTB = kFwdWarps * kFwdSPW;     // samples per tile (multiple of 4)

constexpr int kFwdMaxBnd = 512;      // boundaries staged in shared memory by the RAW kernels

template <int D, int NIT, bool CROSS>
__global__ void __launch_bounds__(kFwdWarps * 32)
embed_fwd_kernel(const EmbedFwdParams p) {
  const bool RAW = p.fields != nullptr;     // fused id stage (ctr_embed_fwd_raw), block-uniform
  constexpr int LPR = D / 4;
  constexpr int RPW = 32 / LPR;
  constexpr int SPW = kFwdSPW;
  constexpr int TB = kFwdTB;
  __shared__ __align__(128) int s_rows[2][TB * CTR_MAX_FIELDS];
  __shared__ __align__(8) uint64_t s_bar[2];
  __shared__ ctr_field_desc s_fields[CTR_MAX_FIELDS];
  __shared__ float s_bnd[kFwdMaxBnd];

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int r = lane / LPR;
  const int q = lane % LPR;
  const int F = p.F;
  const int B = p.B;
  const int ntiles = (B + TB - 1) / TB;

  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    mbar_fence_init();
  }
  if (p.wait_flags != nullptr && tid < p.wait_n)
    wait_flag_bounded(p.wait_flags + tid, *reinterpret_cast<const volatile int*>(p.wait_step),
                      p.wait_ns, p.wait_err);
  if (RAW) {
    for (int i = tid; i < F; i += blockDim.x) s_fields[i] = p.fields[i];
    for (int i = tid; i < p.n_bnd; i += blockDim.x) s_bnd[i] = p.bnd[i];
    if (p.zero_buf != nullptr)
      for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + tid; i < p.zero_n4;
           i += static_cast<long long>(gridDim.x) * blockDim.x)
        reinterpret_cast<float4*>(p.zero_buf)[i] = f4_zero();
  }
  __syncthreads();

  // A tile can be bulk-copied when its byte count is a multiple of 16 (always
  // true except possibly for the ragged last tile).
  auto tile_bulk = [&](int tile) -> bool {
    const int nb = min(TB, B - tile * TB);
    return ((nb * F) & 3) == 0;
  };
  auto prefetch = [&](int tile, int buf) {
    if (tid == 0 && tile_bulk(tile)) {
      const int nb = min(TB, B - tile * TB);
      const uint32_t bytes = static_cast<uint32_t>(nb * F) * 4u;
      mbar_expect_tx(&s_bar[buf], bytes);
      tma_load_1d(&s_rows[buf][0], p.rows + static_cast<size_t>(tile) * TB * F, bytes, &s_bar[buf]);
    }
  };

  uint32_t phase0 = 0, phase1 = 0;
  int buf = 0;
  int tile = blockIdx.x;
  if (!RAW && tile < ntiles) prefetch(tile, 0);

  for (; tile < ntiles; tile += gridDim.x) {
    const int nxt = tile + gridDim.x;
    if (!RAW && nxt < ntiles) prefetch(nxt, buf ^ 1);
    const int b0 = tile * TB;
    if (RAW) {
      // fused id stage: this tile's ids straight from the raw features into shared memory (and
      // out to rows_out for the backward / optimiser) - no separate id kernel, no id round trip
      const int n = min(TB, B - b0) * F;
      for (int i = tid; i < n; i += blockDim.x) {
        const int bl = i / F, f = i - bl * F;
        const int id = criteo_row_id(s_fields[f], s_bnd, p.xcont, p.n_cont, p.xcat, p.n_cat, b0 + bl,
                                     p.logx, p.status);
        s_rows[buf][i] = id;
        p.rows_out[static_cast<size_t>(b0) * F + i] = id;
      }
      __syncthreads();
    } else if (tile_bulk(tile)) {
      if (buf == 0) {
        mbar_wait(&s_bar[0], phase0);
        phase0 ^= 1;
      } else {
        mbar_wait(&s_bar[1], phase1);
        phase1 ^= 1;
      }
    } else {
      const int n = min(TB, B - b0) * F;
      for (int i = tid; i < n; i += blockDim.x)
        s_rows[buf][i] = p.rows[static_cast<size_t>(b0) * F + i];
      __syncthreads();
    }
    const int* srow = &s_rows[buf][0];

    float4 v[SPW][NIT];
    int rid[SPW][NIT];
#pragma unroll
    for (int s = 0; s < SPW; ++s) {
      const int sidx = warp * SPW + s;
      const bool sv = (b0 + sidx) < B;
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int f = it * RPW + r;
        const bool ok = sv && f < F;
        rid[s][it] = ok ? srow[sidx * F + f] : -1;
      }
    }
#pragma unroll
    for (int s = 0; s < SPW; ++s) {
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        v[s][it] = rid[s][it] >= 0
                       ? ldg4(p.table + static_cast<size_t>(rid[s][it]) * p.ld + q * 4)
                       : f4_zero();
      }
    }
    if (p.prefetch_bytes > 0) {
      // 64 B per lane (the default L2 fetch granularity), LPR lanes per row
#pragma unroll
      for (int s = 0; s < SPW; ++s)
#pragma unroll
        for (int it = 0; it < NIT; ++it)
          if (rid[s][it] >= 0)
            for (int o = D * 4 + q * 64; o < D * 4 + p.prefetch_bytes; o += LPR * 64)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(
                  reinterpret_cast<const char*>(p.table + static_cast<size_t>(rid[s][it]) * p.ld) + o));
    }
    float y1p[SPW];
#pragma unroll
    for (int s = 0; s < SPW; ++s) {
      y1p[s] = 0.f;
      if (p.y1 != nullptr && q == 0) {
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
          const int f = it * RPW + r;
          if (rid[s][it] >= 0 && ((p.w1_fields >> f) & 1ull)) y1p[s] += __ldg(p.w1 + static_cast<size_t>(rid[s][it]) * p.ld1);
        }
      }
    }

#pragma unroll
    for (int s = 0; s < SPW; ++s) {
      const int b = b0 + warp * SPW + s;
      if (b >= B) continue;  // warp-uniform
      if (p.E != nullptr) {
        float* e = p.E + static_cast<size_t>(b) * F * D;
#pragma unroll
        for (int it = 0; it < NIT; ++it)
          if (it * RPW + r < F)     // a negative row id (sharded overflow slot) reads as zeros
            *reinterpret_cast<float4*>(e + (it * 32 + lane) * 4) = v[s][it];
        if (p.E_lo != nullptr) {
          float* el = p.E_lo + static_cast<size_t>(b) * F * D;
#pragma unroll
          for (int it = 0; it < NIT; ++it)
            if (it * RPW + r < F)
              *reinterpret_cast<float4*>(el + (it * 32 + lane) * 4) = tcg_lo4(v[s][it]);
        }
      }
      if (p.y1 != nullptr) {
        const float t = warp_sum(y1p[s]);
        if (lane == 0) p.y1[b] = t;
      }
      if (p.S != nullptr || p.y2 != nullptr) {
        float4 sm = f4_zero(), sq = f4_zero();
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
          const float4 x = v[s][it];
          sm = f4_add(sm, x);
          sq = make_float4(fmaf(x.x, x.x, sq.x), fmaf(x.y, x.y, sq.y), fmaf(x.z, x.z, sq.z),
                           fmaf(x.w, x.w, sq.w));
        }
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) {
          sm.x += __shfl_xor_sync(0xffffffffu, sm.x, o);
          sm.y += __shfl_xor_sync(0xffffffffu, sm.y, o);
          sm.z += __shfl_xor_sync(0xffffffffu, sm.z, o);
          sm.w += __shfl_xor_sync(0xffffffffu, sm.w, o);
          sq.x += __shfl_xor_sync(0xffffffffu, sq.x, o);
          sq.y += __shfl_xor_sync(0xffffffffu, sq.y, o);
          sq.z += __shfl_xor_sync(0xffffffffu, sq.z, o);
          sq.w += __shfl_xor_sync(0xffffffffu, sq.w, o);
        }
        if (p.S != nullptr && r == 0)
          *reinterpret_cast<float4*>(p.S + static_cast<size_t>(b) * D + q * 4) = sm;
        if (p.y2 != nullptr) {
          float t = (sm.x * sm.x - sq.x) + (sm.y * sm.y - sq.y) + (sm.z * sm.z - sq.z) +
                    (sm.w * sm.w - sq.w);
#pragma unroll
          for (int o = 1; o < LPR; o <<= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
          if (lane == 0) p.y2[b] = 0.5f * t;
        }
      }
      if (CROSS) {
        // dcn/dcn.py:138-142: xl <- (xl . w_l) * x0 + xl + b_l, row in registers.
        const int W = F * D;
        float4 xl[NIT];
#pragma unroll
        for (int it = 0; it < NIT; ++it) xl[it] = v[s][it];
        for (int l = 0; l < p.cross_layers; ++l) {
          float4 bb[NIT];
          float dot = 0.f;
#pragma unroll
          for (int it = 0; it < NIT; ++it) {
            const int c = (it * 32 + lane) * 4;
            if (c < W) {
              dot += f4_dot(xl[it], ldg4(p.cross_w + static_cast<size_t>(l) * W + c));
              bb[it] = ldg4(p.cross_b + static_cast<size_t>(l) * W + c);
            } else {
              bb[it] = f4_zero();
            }
          }
          dot = warp_sum(dot);
#pragma unroll
          for (int it = 0; it < NIT; ++it) xl[it] = f4_add(f4_fma(dot, v[s][it], xl[it]), bb[it]);
        }
        float* o = p.xl + static_cast<size_t>(b) * W;
#pragma unroll
        for (int it = 0; it < NIT; ++it)
          if ((it * 32 + lane) * 4 < W) *reinterpret_cast<float4*>(o + (it * 32 + lane) * 4) = xl[it];
      }
    }
    __syncthreads();  // everyone is done with s_rows[buf] before it is refilled
    buf ^= 1;
  }
}

template <int D, int NIT, bool CROSS>
static int launch_fwd(const EmbedFwdParams& p, cudaStream_t st) {
  static int occ = 0;
  if (occ == 0) {
    int o = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, embed_fwd_kernel<D, NIT, CROSS>,
                                                  kFwdWarps * 32, 0);
    occ = o > 0 ? o : 1;
  }
  const int ntiles = (p.B + kFwdTB - 1) / kFwdTB;
  const int grid = min(ntiles, sm_count() * occ);
  embed_fwd_kernel<D, NIT, CROSS><<<grid, kFwdWarps * 32, 0, st>>>(p);
  return CTR_OK;
}

template <int D, bool CROSS>
static int dispatch_fwd_nit(const EmbedFwdParams& p, cudaStream_t st) {
  constexpr int RPW = 32 / (D / 4);
  const int need = (p.F + RPW - 1) / RPW;
  if (need <= 1) return launch_fwd<D, 1, CROSS>(p, st);
  if (need <= 2) return launch_fwd<D, 2, CROSS>(p, st);
  if (need <= 3) return launch_fwd<D, 3, CROSS>(p, st);
  if (need <= 4) return launch_fwd<D, 4, CROSS>(p, st);
  if (need <= 5) return launch_fwd<D, 5, CROSS>(p, st);
  if (need <= 6) return launch_fwd<D, 6, CROSS>(p, st);
  if (need <= 8) return launch_fwd<D, 8, CROSS>(p, st);
  if (need <= 10) return launch_fwd<D, 10, CROSS>(p, st);
  return fail_arg("ctr_embed_fwd", "F*D too large (max 1280 floats per sample)");
}

template <bool CROSS>
static int dispatch_fwd(const EmbedFwdParams& p, int D, cudaStream_t st) {
  switch (D) {
    case 8: return dispatch_fwd_nit<8, CROSS>(p, st);
    case 16: return dispatch_fwd_nit<16, CROSS>(p, st);
    case 32: return dispatch_fwd_nit<32, CROSS>(p, st);
    default: return fail_arg("ctr_embed_fwd", "D must be 8, 16 or 32");
  }
}

// (the scatter-add backward lives in embed_adam.cu)

// ------------------------------------------------------------ stand-alone cross
// xl forward on an existing x0[B,W] (used when E comes from somewhere else, and
// as the recompute inside the backward).  One warp per sample, NIT float4/lane.
template <int NIT, int LMAX>
__global__ void __launch_bounds__(256)
cross_bwd_kernel(const float* __restrict__ x0g, const float* __restrict__ w,
                 const float* __restrict__ bvec, int L, int B, int W,
                 const float* __restrict__ dxl, float* __restrict__ dx0, float* __restrict__ dw,
                 float* __restrict__ db) {
  extern __shared__ float s_acc[];  // [2][L][W]: dw then db, CTA-private accumulation
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int i = threadIdx.x; i < 2 * L * W; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  float* s_dw = s_acc;
  float* s_db = s_acc + L * W;

  for (int b = blockIdx.x * wpb + (threadIdx.x >> 5); b < B; b += gridDim.x * wpb) {
    float4 x0[NIT], xs[LMAX][NIT], d[NIT];
    float sdot[LMAX];
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int c = (it * 32 + lane) * 4;
      x0[it] = c < W ? ldg4(x0g + static_cast<size_t>(b) * W + c) : f4_zero();
      d[it] = c < W ? ldg4(dxl + static_cast<size_t>(b) * W + c) : f4_zero();
      xs[0][it] = x0[it];
    }
    // recompute forward, keeping every layer input xs[l] and its dot s_l
#pragma unroll
    for (int l = 0; l < LMAX; ++l) {
      if (l < L) {
        float dot = 0.f;
        float4 bb[NIT];
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
          const int c = (it * 32 + lane) * 4;
          if (c < W) {
            dot += f4_dot(xs[l][it], ldg4(w + static_cast<size_t>(l) * W + c));
            bb[it] = ldg4(bvec + static_cast<size_t>(l) * W + c);
          } else {
            bb[it] = f4_zero();
          }
        }
        dot = warp_sum(dot);
        sdot[l] = dot;
        if (l + 1 < LMAX) {
#pragma unroll
          for (int it = 0; it < NIT; ++it)
            xs[l + 1][it] = f4_add(f4_fma(dot, x0[it], xs[l][it]), bb[it]);
        }
      }
    }
    // backward: d = dL/dx_{l+1}
    float4 dx0acc[NIT];
#pragma unroll
    for (int it = 0; it < NIT; ++it) dx0acc[it] = f4_zero();
#pragma unroll
    for (int l = LMAX - 1; l >= 0; --l) {
      if (l < L) {
        float ds = 0.f;
#pragma unroll
        for (int it = 0; it < NIT; ++it) ds += f4_dot(d[it], x0[it]);
        ds = warp_sum(ds);
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
          const int c = (it * 32 + lane) * 4;
          if (c < W) {
            // db_l += d ; dw_l += ds * x_l ; dx0 += s_l * d ; d <- d + ds * w_l
            float* pdb = s_db + l * W + c;
            float* pdw = s_dw + l * W + c;
            atomicAdd(pdb + 0, d[it].x);
            atomicAdd(pdb + 1, d[it].y);
            atomicAdd(pdb + 2, d[it].z);
            atomicAdd(pdb + 3, d[it].w);
            atomicAdd(pdw + 0, ds * xs[l][it].x);
            atomicAdd(pdw + 1, ds * xs[l][it].y);
            atomicAdd(pdw + 2, ds * xs[l][it].z);
            atomicAdd(pdw + 3, ds * xs[l][it].w);
            dx0acc[it] = f4_fma(sdot[l], d[it], dx0acc[it]);
            d[it] = f4_fma(ds, ldg4(w + static_cast<size_t>(l) * W + c), d[it]);
          }
        }
      }
    }
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int c = (it * 32 + lane) * 4;
      if (c < W)
        *reinterpret_cast<float4*>(dx0 + static_cast<size_t>(b) * W + c) = f4_add(dx0acc[it], d[it]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < L * W; i += blockDim.x) {
    red_add_f32(dw + i, s_dw[i]);
    red_add_f32(db + i, s_db[i]);
  }
}

template <int NIT>
__global__ void __launch_bounds__(256)
cross_fwd_kernel(const float* __restrict__ x0g, const float* __restrict__ w,
                 const float* __restrict__ bvec, int L, int B, int W, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int b = blockIdx.x * wpb + (threadIdx.x >> 5); b < B; b += gridDim.x * wpb) {
    float4 x0[NIT], xl[NIT];
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int c = (it * 32 + lane) * 4;
      x0[it] = c < W ? ldg4(x0g + static_cast<size_t>(b) * W + c) : f4_zero();
      xl[it] = x0[it];
    }
    for (int l = 0; l < L; ++l) {
      float dot = 0.f;
      float4 bb[NIT];
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int c = (it * 32 + lane) * 4;
        if (c < W) {
          dot += f4_dot(xl[it], ldg4(w + static_cast<size_t>(l) * W + c));
          bb[it] = ldg4(bvec + static_cast<size_t>(l) * W + c);
        } else {
          bb[it] = f4_zero();
        }
      }
      dot = warp_sum(dot);
#pragma unroll
      for (int it = 0; it < NIT; ++it) xl[it] = f4_add(f4_fma(dot, x0[it], xl[it]), bb[it]);
    }
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int c = (it * 32 + lane) * 4;
      if (c < W) *reinterpret_cast<float4*>(out + static_cast<size_t>(b) * W + c) = xl[it];
    }
  }
}

// ------------------------------------------------------------------ id pipeline
__global__ void criteo_rows_kernel(const float* __restrict__ xcont, int n_cont,
                                   const long long* __restrict__ xcat, int n_cat,
                                   const ctr_field_desc* __restrict__ fields,
                                   const float* __restrict__ bnd, int B, int F,
                                   int* __restrict__ rows, float* __restrict__ logx,
                                   int* __restrict__ status) {
  const long long n = static_cast<long long>(B) * F;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / F);
    const int f = static_cast<int>(i % F);
    const ctr_field_desc fd = fields[f];
    rows[i] = criteo_row_id(fd, bnd, xcont, n_cont, xcat, n_cat, b, logx, status);
  }
}

// The same ids from a FEW CTAs (ctr_criteo_rows_bg): the variant that runs on a copy stream beside a
// training step.  The step's kernels want whole SMs (one CTA per SM, all of its registers); a
// full-width id kernel that reaches an SM first keeps such a CTA waiting, so this one stays on
// `max_ctas` SMs and makes up for it with memory-level parallelism: tables in shared memory, four
// feature loads in flight per thread.
__global__ void __launch_bounds__(256)
criteo_rows_bg_kernel(const float* __restrict__ xcont, int n_cont, const long long* __restrict__ xcat,
                      int n_cat, const ctr_field_desc* __restrict__ fields,
                      const float* __restrict__ bnd, int n_bnd, int B, int F, int* __restrict__ rows,
                      int* __restrict__ status) {
  __shared__ ctr_field_desc s_fields[CTR_MAX_FIELDS];
  __shared__ float s_bnd[kFwdMaxBnd];
  for (int i = threadIdx.x; i < F * 8; i += blockDim.x)
    reinterpret_cast<int*>(s_fields)[i] = reinterpret_cast<const int*>(fields)[i];
  for (int i = threadIdx.x; i < n_bnd; i += blockDim.x) s_bnd[i] = bnd[i];
  __syncthreads();
  const long long n = static_cast<long long>(B) * F;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i0 < n;
       i0 += 4 * stride) {
    CriteoRaw raw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + u * stride;
      raw[u].xc = 0.f;
      raw[u].xk = 0;
      if (i < n)
        raw[u] = criteo_load_raw(s_fields[i % F], xcont, n_cont, xcat, n_cat, static_cast<int>(i / F));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + u * stride;
      if (i < n) rows[i] = criteo_id_of(s_fields[i % F], s_bnd, raw[u], nullptr, status);
    }
  }
}

// ----------------------------------------------------------------------- Adam
// Device-side Adam schedule (nullable `state`, 4 floats) so that a captured CUDA graph advances it
// on replay:  [0] = t, optimiser steps completed so far;  [1] = lr_t of the step in progress
// (step t+1: lr * sqrt(1 - b2^(t+1)) / (1 - b1^(t+1)));  [2] = lr;  [3] = block counter (as u32).
// The kernels of step t+1 read tag = t+1 and lr_t = [1]; the LAST optimiser kernel of the step
// (ctr_adam_dense with advance_state) moves the schedule on, so no launch is spent on it.
__device__ __forceinline__ void adam_advance(float* __restrict__ state, float b1, float b2) {
  const unsigned t = adam_step_of(state) + 1u;     // integer step counter (bit pattern of state[0])
  const float tn = static_cast<float>(t) + 1.f;
  state[0] = __uint_as_float(t);
  state[1] = state[2] * sqrtf(1.f - powf(b2, tn)) / (1.f - powf(b1, tn));
}
// End of a step with several concurrent optimiser kernels: the state block continues with
// [4] = parties that have finished, [5] = block counter of the row kernel ([3] is the dense
// kernel's).  One thread per block calls this after the block's work; the last block of the last
// party moves the schedule on (every block of every party read t / lr_t when it started).
__device__ __forceinline__ void adam_party_done(float* __restrict__ state, int counter_word,
                                                int parties, float b1, float b2) {
  unsigned* cnt = reinterpret_cast<unsigned*>(state + counter_word);
  __threadfence();
  if (atomicAdd(cnt, 1u) == gridDim.x - 1) {
    *cnt = 0u;
    unsigned* pc = reinterpret_cast<unsigned*>(state + 4);
    if (parties <= 1) {
      adam_advance(state, b1, b2);
    } else if (atomicAdd(pc, 1u) == static_cast<unsigned>(parties) - 1u) {
      *pc = 0u;
      adam_advance(state, b1, b2);
    }
  }
}
// ctr_adam_tick: set lr and advance once (from t = -1 this initialises the schedule at t = 0).
__global__ void adam_tick_kernel(float* __restrict__ state, float lr, float b1, float b2) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    state[2] = lr;
    adam_advance(state, b1, b2);
  }
}

__global__ void adam_dense_kernel(float* __restrict__ th, float* __restrict__ m,
                                  float* __restrict__ v, float* __restrict__ g, long long n,
                                  float lr_t, float b1, float b2, float eps, int zero_g,
                                  float* __restrict__ state, int advance,
                                  float* __restrict__ lo_dst, long long lo_beg, long long lo_n) {
  if (state != nullptr) lr_t = state[1];
  const long long n4 = n >> 2;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += stride) {
    float4 G = reinterpret_cast<float4*>(g)[i];
    float4 M = reinterpret_cast<float4*>(m)[i];
    float4 V = reinterpret_cast<float4*>(v)[i];
    float4 T = reinterpret_cast<float4*>(th)[i];
#define CTR_ADAM1(c)                               \
  M.c = b1 * M.c + (1.f - b1) * G.c;               \
  V.c = b2 * V.c + (1.f - b2) * G.c * G.c;         \
  T.c -= lr_t * M.c / (sqrtf(V.c) + eps);
    CTR_ADAM1(x) CTR_ADAM1(y) CTR_ADAM1(z) CTR_ADAM1(w)
    reinterpret_cast<float4*>(m)[i] = M;
    reinterpret_cast<float4*>(v)[i] = V;
    reinterpret_cast<float4*>(th)[i] = T;
    if (zero_g) reinterpret_cast<float4*>(g)[i] = f4_zero();
    if (lo_dst != nullptr) {      // the 3xTF32 lo half of a slice (the first tower layer's weights)
      const long long e = (i << 2) - lo_beg;
      if (e >= 0 && e + 3 < lo_n) {
        *reinterpret_cast<float4*>(lo_dst + e) = tcg_lo4(T);
      } else if (e > -4 && e < lo_n) {
        const float t4[4] = {T.x, T.y, T.z, T.w};
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (e + u >= 0 && e + u < lo_n) lo_dst[e + u] = tcg_lo(t4[u]);
      }
    }
  }
  for (long long i = (n4 << 2) + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
       i < n; i += stride) {
    const float G = g[i];
    const float M = b1 * m[i] + (1.f - b1) * G;
    const float V = b2 * v[i] + (1.f - b2) * G * G;
    m[i] = M;
    v[i] = V;
    const float Tn = th[i] - lr_t * M / (sqrtf(V) + eps);
    th[i] = Tn;
    if (zero_g) g[i] = 0.f;
    if (lo_dst != nullptr && i >= lo_beg && i - lo_beg < lo_n) lo_dst[i - lo_beg] = tcg_lo(Tn);
  }
  if (advance && state != nullptr) {
    // every block read lr_t when it started; the last one to finish moves the schedule on
    // (advance = number of optimiser kernels that end the step together)
    __syncthreads();
    if (threadIdx.x == 0) adam_party_done(state, 3, advance, b1, b2);
  }
}

// One group of LPR lanes per lookup; the first group to tag claim[row] this step
// owns the row's update (exactly once per distinct row).  The row's g/m/v/theta
// loads are issued together with the claim exchange (not after it) and U lookups
// are in flight per group, so a pass costs two dependent memory round trips
// (row id -> {claim, row data}) instead of three; losers only cost L2 hits.
template <int D, int U>
__global__ void __launch_bounds__(256)
adam_rows_kernel(const int* __restrict__ rows, long long n, float* __restrict__ th,
                 float* __restrict__ m, float* __restrict__ v, float* __restrict__ g,
                 float* __restrict__ th1, float* __restrict__ m1, float* __restrict__ v1,
                 float* __restrict__ g1, int* __restrict__ claim, int tag, float lr_t, float b1,
                 float b2, float eps, float* __restrict__ state, long long ld, long long ld1, long long ldc,
                 int parties) {
  if (state != nullptr) {
    tag = static_cast<int>(adam_step_of(state)) + 1;      // the step in progress
    lr_t = state[1];
  }
  constexpr int LPR = D >= 4 ? D / 4 : 1;
  const int lane = threadIdx.x & 31;
  const int q = lane % LPR;
  const int leader = (lane / LPR) * LPR;
  const long long groups_per_block = blockDim.x / LPR;
  const long long stride = static_cast<long long>(gridDim.x) * groups_per_block;
  const long long nr = (n + stride * U - 1) / (stride * U);
  long long i0 = blockIdx.x * groups_per_block + threadIdx.x / LPR;
  for (long long k = 0; k < nr; ++k, i0 += stride * U) {
    int rid[U], won[U];
    float4 G[U], M[U], V[U], T[U];
    float G1[U], M1[U], V1[U], T1[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      rid[u] = i < n ? __ldg(rows + i) : -1;    // negative ids are padding (sharded exchange slabs)
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      won[u] = 0;
      if (rid[u] >= 0) {
        // claim word read with a plain L2 load first (same sector as the first-order group in the
        // record layout): a row that is already claimed - every later lookup of a hot row - is
        // dropped without an atomic, so same-address atomics no longer serialise at the L2.
        if (q == 0) won[u] = __ldcg(claim + static_cast<size_t>(rid[u]) * ldc) != tag ? 1 : 0;
        if (D >= 4) {
          const size_t o = static_cast<size_t>(rid[u]) * ld + q * 4;
          G[u] = ld4_plain(g + o);
          M[u] = ld4_plain(m + o);
          V[u] = ld4_plain(v + o);
          T[u] = ld4_plain(th + o);
          if (th1 != nullptr && q == 0) {
            const size_t o1 = static_cast<size_t>(rid[u]) * ld1;
            G1[u] = ld1_plain(g1 + o1);
            M1[u] = ld1_plain(m1 + o1);
            V1[u] = ld1_plain(v1 + o1);
            T1[u] = ld1_plain(th1 + o1);
          }
        } else {
          const size_t o1 = static_cast<size_t>(rid[u]) * ld;
          G1[u] = ld1_plain(g + o1);
          M1[u] = ld1_plain(m + o1);
          V1[u] = ld1_plain(v + o1);
          T1[u] = ld1_plain(th + o1);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)     // candidates race for the claim; exactly one wins per row
      if (won[u]) won[u] = atomicExch(claim + static_cast<size_t>(rid[u]) * ldc, tag) != tag ? 1 : 0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int w = __shfl_sync(0xffffffffu, won[u], leader);
      if (!w) continue;
      // Only the claim winner ever writes this row in this launch, and the values it loaded
      // above were written by earlier kernels: the speculative loads are race free.
      if (D >= 4) {
        const size_t o = static_cast<size_t>(rid[u]) * ld + q * 4;
        float4 Gu = G[u], Mu = M[u], Vu = V[u], Tu = T[u];
#define CTR_ADAM4(c)                                  \
  Mu.c = b1 * Mu.c + (1.f - b1) * Gu.c;               \
  Vu.c = b2 * Vu.c + (1.f - b2) * Gu.c * Gu.c;        \
  Tu.c -= lr_t * Mu.c / (sqrtf(Vu.c) + eps);
        CTR_ADAM4(x) CTR_ADAM4(y) CTR_ADAM4(z) CTR_ADAM4(w)
#undef CTR_ADAM4
        *reinterpret_cast<float4*>(m + o) = Mu;
        *reinterpret_cast<float4*>(v + o) = Vu;
        *reinterpret_cast<float4*>(th + o) = Tu;
        *reinterpret_cast<float4*>(g + o) = f4_zero();
        if (th1 != nullptr && q == 0) {   // the row's first-order weight rides on the same claim
          const float Mn = b1 * M1[u] + (1.f - b1) * G1[u];
          const float Vn = b2 * V1[u] + (1.f - b2) * G1[u] * G1[u];
          const size_t o1 = static_cast<size_t>(rid[u]) * ld1;
          m1[o1] = Mn;
          v1[o1] = Vn;
          th1[o1] = T1[u] - lr_t * Mn / (sqrtf(Vn) + eps);
          g1[o1] = 0.f;
        }
      } else {
        const float Mn = b1 * M1[u] + (1.f - b1) * G1[u];
        const float Vn = b2 * V1[u] + (1.f - b2) * G1[u] * G1[u];
        const size_t o1 = static_cast<size_t>(rid[u]) * ld;
        m[o1] = Mn;
        v[o1] = Vn;
        th[o1] = T1[u] - lr_t * Mn / (sqrtf(Vn) + eps);
        g[o1] = 0.f;
      }
    }
  }
  if (parties > 0 && state != nullptr) {
    __syncthreads();
    if (threadIdx.x == 0) adam_party_done(state, 5, parties, b1, b2);
  }
}
#undef CTR_ADAM1

// The same update for a [B, F] id matrix, ONE WAVE deep.  adam_rows_kernel carries the whole
// record of every lookup in registers (78 registers, 3 CTAs per SM), so the 159 744 lookups of a
// 4096 x 39 batch take ~2.7 waves and every wave pays the full dependent chain id -> claim / record
// -> exchange -> store.  Here the election costs one lane and a handful of registers per lookup:
//   * warp = (field f, 32 consecutive samples): the ids of one field side by side, so lookups of a
//     hot row (the <= 32-row fields put hundreds of samples on one row) meet in the same warp and
//     __match_any_sync leaves one candidate per distinct row - a row that every sample hits costs
//     B/32 exchanges instead of B;
//   * candidates read the claim word (plain load: rows already claimed by another warp drop out
//     without an atomic) and race with atomicExch; exactly one lookup per row wins;
//   * the warp's winners are compacted through shared memory and only THEIR records are fetched,
//     LPR lanes per record and U records in flight per lane group.
// Every lookup of the batch is resident at once (5 CTAs x 8 warps per SM >= F * B / 32 warps at
// the benchmark shape), so the kernel is one dependent chain long, not one per wave.
template <int D, int U>
__global__ void __launch_bounds__(256, U == 1 ? 5 : 4)
adam_rows_bf_kernel(const int* __restrict__ rows, int B, int F, float* __restrict__ th,
                    float* __restrict__ m, float* __restrict__ v, float* __restrict__ g,
                    float* __restrict__ th1, float* __restrict__ m1, float* __restrict__ v1,
                    float* __restrict__ g1, int* __restrict__ claim, int tag, float lr_t, float b1,
                    float b2, float eps, float* __restrict__ state, long long ld,
                    long long ld1, long long ldc, int parties) {
  if (state != nullptr) {
    tag = static_cast<int>(adam_step_of(state)) + 1;      // the step in progress
    lr_t = state[1];
  }
  constexpr int LPR = D / 4;
  constexpr int GPW = 32 / LPR;          // records per warp instruction
  __shared__ int s_win[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = lane % LPR, grp = lane / LPR;
  const int wpf = (B + 31) >> 5;
  const long long nwarps = static_cast<long long>(F) * wpf;
  for (long long w = blockIdx.x * 8LL + warp; w < nwarps; w += gridDim.x * 8LL) {
    const int f = static_cast<int>(w / wpf);
    const int b = static_cast<int>(w - static_cast<long long>(f) * wpf) * 32 + lane;
    const int rid = b < B ? __ldg(rows + static_cast<size_t>(b) * F + f) : -1;   // < 0: padding
    const unsigned peers = __match_any_sync(0xffffffffu, rid);
    bool won = rid >= 0 && (__ffs(peers) - 1) == lane;
    if (won) won = __ldcg(claim + static_cast<size_t>(rid) * ldc) != tag;
    if (won) won = atomicExch(claim + static_cast<size_t>(rid) * ldc, tag) != tag;
    const unsigned wm = __ballot_sync(0xffffffffu, won);
    const int nw = __popc(wm);
    if (nw == 0) continue;                 // warp-uniform
    if (won) s_win[warp][__popc(wm & ((1u << lane) - 1u))] = rid;
    __syncwarp();
    for (int k0 = 0; k0 < nw; k0 += GPW * U) {
      int r[U];
      float4 G[U], M[U], V[U], T[U];
      float G1[U], M1[U], V1[U], T1[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int k = k0 + u * GPW + grp;
        r[u] = k < nw ? s_win[warp][k] : -1;
        if (r[u] >= 0) {
          // only the claim winner ever touches this row in this launch, and what it loads was
          // written by earlier kernels
          const size_t o = static_cast<size_t>(r[u]) * ld + q * 4;
          G[u] = ld4_plain(g + o);
          M[u] = ld4_plain(m + o);
          V[u] = ld4_plain(v + o);
          T[u] = ld4_plain(th + o);
          if (th1 != nullptr && q == 0) {
            const size_t o1 = static_cast<size_t>(r[u]) * ld1;
            G1[u] = ld1_plain(g1 + o1);
            M1[u] = ld1_plain(m1 + o1);
            V1[u] = ld1_plain(v1 + o1);
            T1[u] = ld1_plain(th1 + o1);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (r[u] < 0) continue;
        const size_t o = static_cast<size_t>(r[u]) * ld + q * 4;
        float4 Gu = G[u], Mu = M[u], Vu = V[u], Tu = T[u];
#define CTR_ADAM4(c)                                  \
  Mu.c = b1 * Mu.c + (1.f - b1) * Gu.c;               \
  Vu.c = b2 * Vu.c + (1.f - b2) * Gu.c * Gu.c;        \
  Tu.c -= lr_t * Mu.c / (sqrtf(Vu.c) + eps);
        CTR_ADAM4(x) CTR_ADAM4(y) CTR_ADAM4(z) CTR_ADAM4(w)
#undef CTR_ADAM4
        *reinterpret_cast<float4*>(m + o) = Mu;
        *reinterpret_cast<float4*>(v + o) = Vu;
        *reinterpret_cast<float4*>(th + o) = Tu;
        *reinterpret_cast<float4*>(g + o) = f4_zero();
        if (th1 != nullptr && q == 0) {   // the row's first-order weight rides on the same claim
          const float Mn = b1 * M1[u] + (1.f - b1) * G1[u];
          const float Vn = b2 * V1[u] + (1.f - b2) * G1[u] * G1[u];
          const size_t o1 = static_cast<size_t>(r[u]) * ld1;
          m1[o1] = Mn;
          v1[o1] = Vn;
          th1[o1] = T1[u] - lr_t * Mn / (sqrtf(Vn) + eps);
          g1[o1] = 0.f;
        }
      }
    }
    __syncwarp();                          // s_win[warp] is rewritten by the next task
  }
  if (parties > 0 && state != nullptr) {
    __syncthreads();
    if (threadIdx.x == 0) adam_party_done(state, 5, parties, b1, b2);
  }
}

}  // namespace ctr

// =========================================================================== C ABI
using namespace ctr;

template <int NIT>
static int launch_cross_bwd(const float* x0, const float* w, const float* b, int L, int B, int W,
                            const float* dxl, float* dx0, float* dw, float* db, cudaStream_t st) {
  const size_t smem = static_cast<size_t>(2) * L * W * sizeof(float);
  const int grid = std::min((B + 7) / 8, sm_count());
#define CTR_XB(LM)                                                                            \
  {                                                                                           \
    cudaFuncSetAttribute(cross_bwd_kernel<NIT, LM>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                         static_cast<int>(smem));                                             \
    cross_bwd_kernel<NIT, LM><<<grid, 256, smem, st>>>(x0, w, b, L, B, W, dxl, dx0, dw, db);  \
  }
  if (L <= 2) CTR_XB(2) else if (L <= 4) CTR_XB(4) else CTR_XB(6)
#undef CTR_XB
  return CTR_OK;
}

extern "C" {

int ctr_criteo_rows(const float* xcont, int n_cont, const int64_t* xcat, int n_cat,
                    const ctr_field_desc* fields_dev, const float* boundaries_dev, int B, int F,
                    int32_t* rows, float* logx, int32_t* status, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(rows && fields_dev, "ctr_criteo_rows", "null rows/fields");
  CTR_REQUIRE(B >= 0 && F > 0 && F <= CTR_MAX_FIELDS, "ctr_criteo_rows", "bad B/F");
  CTR_REQUIRE(n_cont == 0 || (xcont && boundaries_dev), "ctr_criteo_rows", "null xcont/boundaries");
  CTR_REQUIRE(n_cat == 0 || xcat, "ctr_criteo_rows", "null xcat");
  if (B == 0) return CTR_OK;
  const long long n = static_cast<long long>(B) * F;
  const int grid = static_cast<int>(std::min<long long>((n + 255) / 256, sm_count() * 8LL));
  criteo_rows_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      xcont, n_cont, reinterpret_cast<const long long*>(xcat), n_cat, fields_dev, boundaries_dev, B,
      F, rows, logx, status);
  CTR_LAUNCH_CHECK("ctr_criteo_rows");
}

int ctr_criteo_rows_bg(const float* xcont, int n_cont, const int64_t* xcat, int n_cat,
                       const ctr_field_desc* fields_dev, const float* boundaries_dev, int n_boundaries,
                       int B, int F, int32_t* rows, int32_t* status, int max_ctas,
                       ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(fields_dev && rows && B >= 0 && F > 0 && F <= CTR_MAX_FIELDS, "ctr_criteo_rows_bg",
              "null fields/rows or bad B/F");
  CTR_REQUIRE(n_cont == 0 || (xcont && boundaries_dev), "ctr_criteo_rows_bg", "null xcont/boundaries");
  CTR_REQUIRE(n_cat == 0 || xcat, "ctr_criteo_rows_bg", "null xcat");
  CTR_REQUIRE(n_boundaries >= 0 && n_boundaries <= kFwdMaxBnd, "ctr_criteo_rows_bg",
              "at most 512 bucket boundaries in total");
  if (B == 0) return CTR_OK;
  const long long n = static_cast<long long>(B) * F;
  if (max_ctas <= 0) max_ctas = 16;
  const int grid = static_cast<int>(std::min<long long>((n + 1023) / 1024, max_ctas));
  criteo_rows_bg_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      xcont, n_cont, reinterpret_cast<const long long*>(xcat), n_cat, fields_dev, boundaries_dev,
      n_boundaries, B, F, rows, status);
  CTR_LAUNCH_CHECK("ctr_criteo_rows_bg");
}

static int embed_fwd_impl(const char* fn, const float* table, const float* w1, const int32_t* rows,
                          int B, int F, int D, uint64_t w1_fields, float* E, float* S, float* y1,
                          float* y2, const float* cross_w, const float* cross_b, int cross_layers,
                          float* xl, float* E_lo, int64_t row_stride, int64_t w1_stride,
                          const EmbedFwdParams* raw, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(table && (rows || raw), fn, "null table/rows");
  CTR_REQUIRE(E_lo == nullptr || (E != nullptr && aligned16(E_lo)), fn,
              "E_lo needs E and 16-byte alignment");
  CTR_REQUIRE(B >= 0 && F > 0 && F <= CTR_MAX_FIELDS, fn, "need 0 < F <= 64");
  CTR_REQUIRE(aligned16(table) && aligned16(rows) && aligned16(E) && aligned16(S) && aligned16(xl) &&
                  aligned16(cross_w) && aligned16(cross_b),
              fn, "pointers must be 16-byte aligned");
  CTR_REQUIRE(y1 == nullptr || w1 != nullptr, fn, "y1 requested without w1");
  if (row_stride <= 0) row_stride = D;
  if (w1_stride <= 0) w1_stride = 1;
  CTR_REQUIRE(row_stride >= D && (row_stride & 3) == 0, fn,
              "row_stride must be >= D and a multiple of 4 floats");
  const bool cross = xl != nullptr;
  CTR_REQUIRE(!cross || (cross_w && cross_b && cross_layers >= 0), fn,
              "xl requested without cross_w/cross_b");
  if (B == 0) return CTR_OK;
  EmbedFwdParams p{};
  if (raw != nullptr) p = *raw;
  p.table = table; p.w1 = w1; p.rows = rows; p.E = E; p.E_lo = E_lo; p.S = S; p.y1 = y1; p.y2 = y2;
  p.cross_w = cross_w; p.cross_b = cross_b; p.xl = xl; p.w1_fields = w1_fields;
  p.cross_layers = cross_layers; p.B = B; p.F = F; p.ld = row_stride; p.ld1 = w1_stride;
  // row-record layout (stride 4D+8 floats, include/ctr_b200.h "Row strides") in a training step
  // (the RAW entry point with rows_out): prefetch the rest of each record
  if (raw != nullptr && raw->rows_out != nullptr && row_stride == 4 * D + 8 &&
      option_get("fwd_prefetch_record", 0) != 0)
    p.prefetch_bytes = static_cast<int>(row_stride - D) * 4;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int r = cross ? dispatch_fwd<true>(p, D, st) : dispatch_fwd<false>(p, D, st);
  if (r != CTR_OK) return r;
  return check_cuda(cudaGetLastError(), fn);
}

int ctr_embed_fwd(const float* table, const float* w1, const int32_t* rows, int B, int F, int D,
                  uint64_t w1_fields, float* E, float* S, float* y1, float* y2,
                  const float* cross_w, const float* cross_b, int cross_layers, float* xl,
                  float* E_lo, int64_t row_stride, int64_t w1_stride, ctr_stream_t stream) {
  return embed_fwd_impl("ctr_embed_fwd", table, w1, rows, B, F, D, w1_fields, E, S, y1, y2, cross_w,
                        cross_b, cross_layers, xl, E_lo, row_stride, w1_stride, nullptr, stream);
}

int ctr_embed_fwd_p2p(const int32_t* slot, int B, int F, int D, uint64_t w1_fields, int with_w1,
                      float* E, float* S, float* y1, float* y2, const float* cross_w,
                      const float* cross_b, int cross_layers, float* xl, float* E_lo,
                      const ctr_p2p_ctx* ctx, ctr_stream_t stream) {
  CTR_REQUIRE(ctx && ctx->G >= 1 && ctx->G <= CTR_P2P_MAX_RANKS && ctx->me >= 0 && ctx->me < ctx->G &&
                  ctx->peer[ctx->me] && ctx->record_floats == D + 4,
              "ctr_embed_fwd_p2p", "bad exchange context");
  char* base = static_cast<char*>(ctx->peer[ctx->me]);
  const float* resp = reinterpret_cast<const float*>(base + ctx->off_resp);
  EmbedFwdParams w{};
  w.wait_flags = reinterpret_cast<const int*>(base + ctx->off_resp_flag);
  w.wait_step = reinterpret_cast<const int*>(base);
  w.wait_err = reinterpret_cast<int*>(base) + 1;
  w.wait_n = ctx->G;
  w.wait_ns = ctx->spin_limit_ms > 0 ? static_cast<long long>(ctx->spin_limit_ms) * 1000000LL : 10000000000LL;
  return embed_fwd_impl("ctr_embed_fwd_p2p", resp, with_w1 ? resp + D : nullptr, slot, B, F, D,
                        w1_fields, E, S, y1, y2, cross_w, cross_b, cross_layers, xl, E_lo,
                        ctx->record_floats, ctx->record_floats, &w, stream);
}

int ctr_embed_fwd_raw(const float* table, const float* w1, const float* xcont, int n_cont,
                      const int64_t* xcat, int n_cat, const ctr_field_desc* fields_dev,
                      const float* boundaries_dev, int n_boundaries, int32_t* rows_out, float* logx,
                      int32_t* status, int B, int F, int D, uint64_t w1_fields, float* E, float* S,
                      float* y1, float* y2, const float* cross_w, const float* cross_b,
                      int cross_layers, float* xl, float* E_lo, int64_t row_stride,
                      int64_t w1_stride, float* zero_buf, int64_t zero_n, ctr_stream_t stream) {
  CTR_REQUIRE(fields_dev && rows_out, "ctr_embed_fwd_raw", "null fields/rows_out");
  CTR_REQUIRE(zero_buf == nullptr || (aligned16(zero_buf) && zero_n >= 0 && (zero_n & 3) == 0),
              "ctr_embed_fwd_raw", "zero_buf must be 16-byte aligned, zero_n a multiple of 4");
  CTR_REQUIRE(n_cont == 0 || (xcont && boundaries_dev), "ctr_embed_fwd_raw", "null xcont/boundaries");
  CTR_REQUIRE(n_cat == 0 || xcat, "ctr_embed_fwd_raw", "null xcat");
  CTR_REQUIRE(n_boundaries >= 0 && n_boundaries <= kFwdMaxBnd, "ctr_embed_fwd_raw",
              "at most 512 bucket boundaries in total");
  EmbedFwdParams raw{};
  raw.xcont = xcont; raw.xcat = reinterpret_cast<const long long*>(xcat); raw.fields = fields_dev;
  raw.bnd = boundaries_dev; raw.rows_out = rows_out; raw.logx = logx; raw.status = status;
  raw.n_cont = n_cont; raw.n_cat = n_cat; raw.n_bnd = n_boundaries;
  raw.zero_buf = zero_n > 0 ? zero_buf : nullptr; raw.zero_n4 = zero_n >> 2;
  return embed_fwd_impl("ctr_embed_fwd_raw", table, w1, nullptr, B, F, D, w1_fields, E, S, y1, y2,
                        cross_w, cross_b, cross_layers, xl, E_lo, row_stride, w1_stride, &raw, stream);
}

int ctr_dcn_cross_fwd(const float* x0, const float* w, const float* b, int L, int B, int W,
                      float* xl, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(x0 && w && b && xl, "ctr_dcn_cross_fwd", "null pointer");
  CTR_REQUIRE(L >= 0 && B >= 0 && W > 0 && (W & 3) == 0 && W <= 1280, "ctr_dcn_cross_fwd",
              "need W%4==0, W<=1280");
  CTR_REQUIRE(aligned16(x0) && aligned16(w) && aligned16(b) && aligned16(xl), "ctr_dcn_cross_fwd",
              "pointers must be 16-byte aligned");
  if (B == 0) return CTR_OK;
  const int need = (W / 4 + 31) / 32;
  const int grid = std::min((B + 7) / 8, sm_count() * 4);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (need <= 2) cross_fwd_kernel<2><<<grid, 256, 0, st>>>(x0, w, b, L, B, W, xl);
  else if (need <= 5) cross_fwd_kernel<5><<<grid, 256, 0, st>>>(x0, w, b, L, B, W, xl);
  else cross_fwd_kernel<10><<<grid, 256, 0, st>>>(x0, w, b, L, B, W, xl);
  CTR_LAUNCH_CHECK("ctr_dcn_cross_fwd");
}

int ctr_dcn_cross_bwd(const float* x0, const float* w, const float* b, int L, int B, int W,
                      const float* dxl, float* dx0, float* dw, float* db, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(x0 && w && b && dxl && dx0 && dw && db, "ctr_dcn_cross_bwd", "null pointer");
  CTR_REQUIRE(L >= 1 && L <= 6, "ctr_dcn_cross_bwd", "need 1 <= L <= 6");
  CTR_REQUIRE(B >= 0 && W > 0 && (W & 3) == 0 && W <= 1280, "ctr_dcn_cross_bwd",
              "need W%4==0, W<=1280");
  CTR_REQUIRE(static_cast<size_t>(2) * L * W * 4 <= 200 * 1024, "ctr_dcn_cross_bwd",
              "L*W too large for the shared-memory accumulators");
  CTR_REQUIRE(aligned16(x0) && aligned16(w) && aligned16(b) && aligned16(dxl) && aligned16(dx0),
              "ctr_dcn_cross_bwd", "pointers must be 16-byte aligned");
  if (B == 0) return CTR_OK;
  const int need = (W / 4 + 31) / 32;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (need <= 2) launch_cross_bwd<2>(x0, w, b, L, B, W, dxl, dx0, dw, db, st);
  else if (need <= 5) launch_cross_bwd<5>(x0, w, b, L, B, W, dxl, dx0, dw, db, st);
  else launch_cross_bwd<10>(x0, w, b, L, B, W, dxl, dx0, dw, db, st);
  CTR_LAUNCH_CHECK("ctr_dcn_cross_bwd");
}

int ctr_adam_tick(float* state_dev, float lr, float beta1, float beta2, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(state_dev, "ctr_adam_tick", "null state");
  adam_tick_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(state_dev, lr, beta1, beta2);
  CTR_LAUNCH_CHECK("ctr_adam_tick");
}

int ctr_adam_dense_ex(float* theta, float* m, float* v, float* g, int64_t n, float lr_t, float beta1,
                      float beta2, float eps, int zero_g, float* state_dev, int advance_parties,
                      float* lo_dst, int64_t lo_begin, int64_t lo_n, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(theta && m && v && g && n >= 0, "ctr_adam_dense", "null pointer / negative n");
  CTR_REQUIRE(aligned16(theta) && aligned16(m) && aligned16(v) && aligned16(g), "ctr_adam_dense",
              "pointers must be 16-byte aligned");
  CTR_REQUIRE(lo_dst == nullptr || (lo_begin >= 0 && lo_n >= 0 && lo_begin + lo_n <= n &&
                                    (lo_begin & 3) == 0 && aligned16(lo_dst)),
              "ctr_adam_dense", "lo slice must lie inside theta, start at a multiple of 4, dst aligned");
  if (n == 0) return CTR_OK;
  const int grid = static_cast<int>(std::min<long long>((n / 4 + 255) / 256 + 1, sm_count() * 8LL));
  adam_dense_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      theta, m, v, g, n, lr_t, beta1, beta2, eps, zero_g, state_dev, advance_parties, lo_dst, lo_begin,
      lo_n);
  CTR_LAUNCH_CHECK("ctr_adam_dense");
}

int ctr_adam_dense(float* theta, float* m, float* v, float* g, int64_t n, float lr_t, float beta1,
                   float beta2, float eps, int zero_g, float* state_dev, int advance_state,
                   ctr_stream_t stream) {
  return ctr_adam_dense_ex(theta, m, v, g, n, lr_t, beta1, beta2, eps, zero_g, state_dev,
                           advance_state != 0 ? 1 : 0, nullptr, 0, 0, stream);
}

int ctr_adam_rows(const int32_t* rows, int64_t n, int D, float* theta, float* m, float* v,
                  float* g, float* theta1, float* m1, float* v1, float* g1, int32_t* claim,
                  int32_t tag, float lr_t, float beta1, float beta2, float eps,
                  const float* state_dev, int64_t row_stride, int64_t w1_stride, int64_t claim_stride,
                  ctr_stream_t stream) {
  return ctr_adam_rows_ex(rows, n, D, theta, m, v, g, theta1, m1, v1, g1, claim, tag, lr_t, beta1, beta2,
                          eps, const_cast<float*>(state_dev), row_stride, w1_stride, claim_stride, 0,
                          stream);
}

int ctr_adam_rows_ex(const int32_t* rows, int64_t n, int D, float* theta, float* m, float* v,
                     float* g, float* theta1, float* m1, float* v1, float* g1, int32_t* claim,
                     int32_t tag, float lr_t, float beta1, float beta2, float eps, float* state_dev,
                     int64_t row_stride, int64_t w1_stride, int64_t claim_stride,
                     int advance_parties, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(rows && theta && m && v && g && claim && n >= 0, "ctr_adam_rows", "null pointer");
  CTR_REQUIRE(theta1 == nullptr || (m1 && v1 && g1 && D >= 4), "ctr_adam_rows",
              "first-order vector needs m1/v1/g1 and a D >= 4 table");
  CTR_REQUIRE(aligned16(theta) && aligned16(m) && aligned16(v) && aligned16(g), "ctr_adam_rows",
              "pointers must be 16-byte aligned");
  if (row_stride <= 0) row_stride = D;
  if (w1_stride <= 0) w1_stride = 1;
  if (claim_stride <= 0) claim_stride = 1;
  CTR_REQUIRE(row_stride >= D && (D < 4 || (row_stride & 3) == 0), "ctr_adam_rows",
              "row_stride must be >= D and a multiple of 4 floats");
  if (n == 0) return CTR_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int lpr = D >= 4 ? D / 4 : 1;
  const long long gpb = 256 / lpr;
  constexpr int U = 2;                      // lookups in flight per lane group
  const long long per_block = gpb * U;
  const int grid = static_cast<int>(std::min<long long>((n + per_block - 1) / per_block, sm_count() * 8LL));
#define CTR_AR(DD) adam_rows_kernel<DD, U><<<grid, 256, 0, st>>>(rows, n, theta, m, v, g, theta1, m1, v1, g1, claim, tag, lr_t, beta1, beta2, eps, state_dev, row_stride, w1_stride, claim_stride, advance_parties)
  switch (D) {
    case 1: CTR_AR(1); break;
    case 8: CTR_AR(8); break;
    case 16: CTR_AR(16); break;
    case 32: CTR_AR(32); break;
    default: return fail_arg("ctr_adam_rows", "D must be 1, 8, 16 or 32");
  }
#undef CTR_AR
  CTR_LAUNCH_CHECK("ctr_adam_rows");
}

int ctr_adam_rows_bf(const int32_t* rows, int B, int F, int D, float* theta, float* m, float* v,
                     float* g, float* theta1, float* m1, float* v1, float* g1, int32_t* claim,
                     int32_t tag, float lr_t, float beta1, float beta2, float eps,
                     float* state_dev, int64_t row_stride, int64_t w1_stride,
                     int64_t claim_stride, int advance_parties, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(rows && theta && m && v && g && claim && B >= 0 && F >= 0, "ctr_adam_rows_bf",
              "null pointer");
  CTR_REQUIRE(D == 8 || D == 16 || D == 32, "ctr_adam_rows_bf", "D must be 8, 16 or 32");
  CTR_REQUIRE(theta1 == nullptr || (m1 && v1 && g1), "ctr_adam_rows_bf",
              "first-order vector needs m1/v1/g1");
  CTR_REQUIRE(aligned16(theta) && aligned16(m) && aligned16(v) && aligned16(g), "ctr_adam_rows_bf",
              "pointers must be 16-byte aligned");
  if (row_stride <= 0) row_stride = D;
  if (w1_stride <= 0) w1_stride = 1;
  if (claim_stride <= 0) claim_stride = 1;
  CTR_REQUIRE(row_stride >= D && (row_stride & 3) == 0, "ctr_adam_rows_bf",
              "row_stride must be >= D and a multiple of 4 floats");
  if (B == 0 || F == 0) return CTR_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long nwarps = static_cast<long long>(F) * ((B + 31) / 32);
  const int U = option_get("adam_rows_inflight", 1) >= 2 ? 2 : 1;
  const int grid = static_cast<int>(std::min<long long>((nwarps + 7) / 8, sm_count() * (U == 1 ? 5LL : 4LL)));
#define CTR_ARB(DD, UU) adam_rows_bf_kernel<DD, UU><<<grid, 256, 0, st>>>(rows, B, F, theta, m, v, g, theta1, m1, v1, g1, claim, tag, lr_t, beta1, beta2, eps, state_dev, row_stride, w1_stride, claim_stride, advance_parties)
#define CTR_ARB_D(DD) \
  if (U == 2) CTR_ARB(DD, 2); else CTR_ARB(DD, 1)
  switch (D) {
    case 8: CTR_ARB_D(8); break;
    case 16: CTR_ARB_D(16); break;
    default: CTR_ARB_D(32); break;
  }
#undef CTR_ARB_D
#undef CTR_ARB
  CTR_LAUNCH_CHECK("ctr_adam_rows_bf");
}

}  // extern "C"
