// Row-sharded embedding table: the three device primitives around the NCCL
// all-to-all.  owner(row) = row % G, local index = row / G.
//   requester:  bucket the batch's lookups by owner into fixed-capacity send slabs
//   owner:      gather the requested rows / scatter-add the returned gradients
// The requester-side interaction forward/backward re-use ctr_embed_fwd / ctr_embed_bwd
// with the received vectors as the "table" and the slab slots as the "row ids".
#include "common.cuh"

namespace ctr {

// One lookup per thread.  Threads of a warp that target the same owner take
// consecutive slots with one atomic per (warp, owner).
__global__ void __launch_bounds__(256)
shard_bucket_kernel(const int* __restrict__ rows, long long n, int G, int capacity,
                    int* __restrict__ send_local, int* __restrict__ slot, int* __restrict__ counts) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long nr = (n + stride - 1) / stride;
  long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const int lane = threadIdx.x & 31;
  for (long long k = 0; k < nr; ++k, i += stride) {
    const bool ok = i < n;
    const int row = ok ? __ldg(rows + i) : -1;
    const int owner = ok ? row % G : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, owner);
    if (!ok) continue;
    const int leader = __ffs(peers) - 1;
    const int rank_in = __popc(peers & ((1u << lane) - 1u));
    int base = 0;
    if (lane == leader) base = atomicAdd(counts + owner, __popc(peers));
    base = __shfl_sync(peers, base, leader);
    const int pos = base + rank_in;
    if (pos < capacity) {
      send_local[owner * capacity + pos] = row / G;
      slot[i] = owner * capacity + pos;
    } else {
      slot[i] = -1;   // overflow: counts[owner] > capacity is reported by the caller
    }
  }
}

template <int D>
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ table, const float* __restrict__ w1,
                   const int* __restrict__ ids, long long n, float* __restrict__ out,
                   float* __restrict__ out_w1, long long lt, long long lw, long long lo,
                   long long low) {
  constexpr int LPR = D / 4;
  const long long gpb = blockDim.x / LPR;
  const int q = threadIdx.x % LPR;
  for (long long i = blockIdx.x * gpb + threadIdx.x / LPR; i < n; i += gridDim.x * gpb) {
    const int id = __ldg(ids + i);
    float4 v = f4_zero();
    if (id >= 0) v = ldg4(table + static_cast<size_t>(id) * lt + q * 4);
    *reinterpret_cast<float4*>(out + static_cast<size_t>(i) * lo + q * 4) = v;
    if (out_w1 != nullptr && q == 0)
      out_w1[static_cast<size_t>(i) * low] = id >= 0 ? __ldg(w1 + static_cast<size_t>(id) * lw) : 0.f;
  }
}

template <int D>
__global__ void __launch_bounds__(256)
scatter_add_rows_kernel(const int* __restrict__ ids, const float* __restrict__ g,
                        const float* __restrict__ gw1, long long n, float* __restrict__ dtable,
                        float* __restrict__ dw1, long long lg, long long lgw, long long ld,
                        long long ldw) {
  constexpr int LPR = D / 4;
  const long long gpb = blockDim.x / LPR;
  const int q = threadIdx.x % LPR;
  for (long long i = blockIdx.x * gpb + threadIdx.x / LPR; i < n; i += gridDim.x * gpb) {
    const int id = __ldg(ids + i);
    if (id < 0) continue;
    red_add_v4(dtable + static_cast<size_t>(id) * ld + q * 4,
               ld4_stream(g + static_cast<size_t>(i) * lg + q * 4));
    if (dw1 != nullptr && q == 0)
      red_add_f32(dw1 + static_cast<size_t>(id) * ldw, gw1[static_cast<size_t>(i) * lgw]);
  }
}

}  // namespace ctr

using namespace ctr;

extern "C" {

int ctr_shard_bucket(const int32_t* rows, int64_t n, int G, int capacity, int32_t* send_local,
                     int32_t* slot, int32_t* counts, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(rows && send_local && slot && counts, "ctr_shard_bucket", "null pointer");
  CTR_REQUIRE(n >= 0 && G >= 1 && G <= 64 && capacity >= 1, "ctr_shard_bucket", "bad n/G/capacity");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(send_local, 0xFF, sizeof(int32_t) * static_cast<size_t>(G) * capacity, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(counts, 0, sizeof(int32_t) * G, st);
  if (e != cudaSuccess) return check_cuda(e, "ctr_shard_bucket");
  if (n == 0) return CTR_OK;
  const int grid = static_cast<int>(std::min<long long>((n + 255) / 256, sm_count() * 8LL));
  shard_bucket_kernel<<<grid, 256, 0, st>>>(rows, n, G, capacity, send_local, slot, counts);
  CTR_LAUNCH_CHECK("ctr_shard_bucket");
}

int ctr_gather_rows(const float* table, const float* w1, const int32_t* ids, int64_t n, int D,
                    float* out, float* out_w1, int64_t table_stride, int64_t w1_stride,
                    int64_t out_stride, int64_t out_w1_stride, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(table && ids && out && n >= 0, "ctr_gather_rows", "null pointer");
  CTR_REQUIRE(out_w1 == nullptr || w1 != nullptr, "ctr_gather_rows", "out_w1 without w1");
  CTR_REQUIRE(aligned16(table) && aligned16(out), "ctr_gather_rows", "pointers must be 16-byte aligned");
  const long long lt = table_stride > 0 ? table_stride : D, lw = w1_stride > 0 ? w1_stride : 1;
  const long long lo = out_stride > 0 ? out_stride : D, low = out_w1_stride > 0 ? out_w1_stride : 1;
  CTR_REQUIRE((lt & 3) == 0 && (lo & 3) == 0 && lt >= D && lo >= D, "ctr_gather_rows",
              "row strides must be >= D and multiples of 4 floats");
  if (n == 0) return CTR_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long gpb = 256 / (D / 4 > 0 ? D / 4 : 1);
  const int grid = static_cast<int>(std::min<long long>((n + gpb - 1) / gpb, sm_count() * 8LL));
  switch (D) {
    case 8: gather_rows_kernel<8><<<grid, 256, 0, st>>>(table, w1, ids, n, out, out_w1, lt, lw, lo, low); break;
    case 16: gather_rows_kernel<16><<<grid, 256, 0, st>>>(table, w1, ids, n, out, out_w1, lt, lw, lo, low); break;
    case 32: gather_rows_kernel<32><<<grid, 256, 0, st>>>(table, w1, ids, n, out, out_w1, lt, lw, lo, low); break;
    default: return fail_arg("ctr_gather_rows", "D must be 8, 16 or 32");
  }
  CTR_LAUNCH_CHECK("ctr_gather_rows");
}

int ctr_scatter_add_rows(const int32_t* ids, const float* g, const float* gw1, int64_t n, int D,
                         float* dtable, float* dw1, int64_t g_stride, int64_t gw1_stride,
                         int64_t dtable_stride, int64_t dw1_stride, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(ids && g && dtable && n >= 0, "ctr_scatter_add_rows", "null pointer");
  CTR_REQUIRE(dw1 == nullptr || gw1 != nullptr, "ctr_scatter_add_rows", "dw1 without gw1");
  CTR_REQUIRE(aligned16(g) && aligned16(dtable), "ctr_scatter_add_rows",
              "pointers must be 16-byte aligned");
  const long long lg = g_stride > 0 ? g_stride : D, lgw = gw1_stride > 0 ? gw1_stride : 1;
  const long long ld = dtable_stride > 0 ? dtable_stride : D, ldw = dw1_stride > 0 ? dw1_stride : 1;
  CTR_REQUIRE((lg & 3) == 0 && (ld & 3) == 0 && lg >= D && ld >= D, "ctr_scatter_add_rows",
              "row strides must be >= D and multiples of 4 floats");
  if (n == 0) return CTR_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long gpb = 256 / (D / 4 > 0 ? D / 4 : 1);
  const int grid = static_cast<int>(std::min<long long>((n + gpb - 1) / gpb, sm_count() * 8LL));
  switch (D) {
    case 8: scatter_add_rows_kernel<8><<<grid, 256, 0, st>>>(ids, g, gw1, n, dtable, dw1, lg, lgw, ld, ldw); break;
    case 16: scatter_add_rows_kernel<16><<<grid, 256, 0, st>>>(ids, g, gw1, n, dtable, dw1, lg, lgw, ld, ldw); break;
    case 32: scatter_add_rows_kernel<32><<<grid, 256, 0, st>>>(ids, g, gw1, n, dtable, dw1, lg, lgw, ld, ldw); break;
    default: return fail_arg("ctr_scatter_add_rows", "D must be 8, 16 or 32");
  }
  CTR_LAUNCH_CHECK("ctr_scatter_add_rows");
}

}  // extern "C"
