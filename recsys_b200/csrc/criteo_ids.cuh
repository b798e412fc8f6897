// The Criteo id pipeline as a device function, shared by ctr_criteo_rows and the fused id stages
// of the lookup kernels (embed.cu, embed_tower.cu), so that all of them produce bit-identical ids.
#pragma once
#include "common.cuh"

namespace ctr {

// One row id of the Criteo id pipeline (fm/fm.py:76-80,89): numeric fields log -> Bucketize
// (upper_bound), categorical fields offset + range check.  Shared by criteo_rows_kernel and the
// lookup kernels' fused id stages, so all of them produce bit-identical ids.  In two halves, so
// that a caller with several ids to compute can issue all of its feature loads first.
struct CriteoRaw {
  float xc;          // kind 0: the numeric feature
  long long xk;      // kind 1: the (hashed) categorical id
};
__device__ __forceinline__ CriteoRaw criteo_load_raw(const ctr_field_desc& fd,
                                                     const float* __restrict__ xcont, int n_cont,
                                                     const long long* __restrict__ xcat, int n_cat,
                                                     int b) {
  CriteoRaw r;
  r.xc = 0.f;
  r.xk = 0;
  if (fd.kind == 0) r.xc = xcont[static_cast<size_t>(b) * n_cont + fd.src];
  else r.xk = xcat[static_cast<size_t>(b) * n_cat + fd.src];
  return r;
}
__device__ __forceinline__ int criteo_id_of(const ctr_field_desc& fd, const float* __restrict__ bnd,
                                            const CriteoRaw& raw_in, float* __restrict__ logx_slot,
                                            int* __restrict__ status) {
  int id;
  if (fd.kind == 0) {
    // fm/fm.py:76-79: tf.log(x + off) in fp32, then Bucketize == upper_bound.
    const float v = logf(raw_in.xc + fd.log_offset);
    if (logx_slot != nullptr) *logx_slot = v;
    id = 0;
    for (int k = 0; k < fd.bnd_count; ++k) id += (bnd[fd.bnd_begin + k] <= v) ? 1 : 0;
    if (v != v) id = fd.bnd_count;
  } else {
    long long raw = raw_in.xk;
    if (raw < 0 || raw >= fd.n_rows) {
      if (status != nullptr) atomicOr(status, 1);
      raw %= fd.n_rows;
      if (raw < 0) raw += fd.n_rows;
    }
    id = static_cast<int>(raw);
  }
  return fd.row_offset + id;
}
__device__ __forceinline__ int criteo_row_id(const ctr_field_desc& fd, const float* __restrict__ bnd,
                                             const float* __restrict__ xcont, int n_cont,
                                             const long long* __restrict__ xcat, int n_cat, int b,
                                             float* __restrict__ logx, int* __restrict__ status) {
  const CriteoRaw raw = criteo_load_raw(fd, xcont, n_cont, xcat, n_cat, b);
  return criteo_id_of(fd, bnd, raw,
                      (logx != nullptr && fd.kind == 0) ? logx + static_cast<size_t>(b) * n_cont + fd.src
                                                        : nullptr,
                      status);
}

}  // namespace ctr
