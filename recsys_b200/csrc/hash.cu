// FarmHash Fingerprint64 (farmhashna::Hash64) on the device: one thread per string.
// Replaces TF's StringToHashBucketFast for categorical_column_with_hash_bucket
// (fm/fm.py:89, deepfm/deepfm.py:41,46).  Criteo values are 8 hex characters and
// the default b"NULL" is 4, so the <= 16-byte branches carry the traffic; the long
// branches are implemented for completeness.  Byte/integer work, HBM-bound.
#include "common.cuh"

namespace ctr {

constexpr unsigned long long K0 = 0xc3a5c85c97cb3127ULL;
constexpr unsigned long long K1 = 0xb492b66fbe98f273ULL;
constexpr unsigned long long K2 = 0x9ae16a3b2f90404fULL;
typedef unsigned long long u64;

__device__ __forceinline__ u64 fetch64(const uint8_t* p) {
  u64 v = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) v |= static_cast<u64>(p[i]) << (8 * i);
  return v;
}
__device__ __forceinline__ u64 fetch32(const uint8_t* p) {
  u64 v = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) v |= static_cast<u64>(p[i]) << (8 * i);
  return v;
}
__device__ __forceinline__ u64 rot(u64 v, int s) { return s == 0 ? v : (v >> s) | (v << (64 - s)); }
__device__ __forceinline__ u64 smix(u64 v) { return v ^ (v >> 47); }
__device__ __forceinline__ u64 hl16(u64 u, u64 v, u64 mul) {
  u64 a = (u ^ v) * mul;
  a ^= (a >> 47);
  u64 b = (v ^ a) * mul;
  b ^= (b >> 47);
  return b * mul;
}
__device__ u64 len0to16(const uint8_t* s, int n) {
  if (n >= 8) {
    const u64 mul = K2 + static_cast<u64>(n) * 2;
    const u64 a = fetch64(s) + K2;
    const u64 b = fetch64(s + n - 8);
    const u64 c = rot(b, 37) * mul + a;
    const u64 d = (rot(a, 25) + b) * mul;
    return hl16(c, d, mul);
  }
  if (n >= 4) {
    const u64 mul = K2 + static_cast<u64>(n) * 2;
    const u64 a = fetch32(s);
    return hl16(static_cast<u64>(n) + (a << 3), fetch32(s + n - 4), mul);
  }
  if (n > 0) {
    const uint8_t a = s[0], b = s[n >> 1], c = s[n - 1];
    const uint32_t y = static_cast<uint32_t>(a) + (static_cast<uint32_t>(b) << 8);
    const uint32_t z = static_cast<uint32_t>(n) + (static_cast<uint32_t>(c) << 2);
    return smix(static_cast<u64>(y) * K2 ^ static_cast<u64>(z) * K0) * K2;
  }
  return K2;
}
__device__ u64 len17to32(const uint8_t* s, int n) {
  const u64 mul = K2 + static_cast<u64>(n) * 2;
  const u64 a = fetch64(s) * K1;
  const u64 b = fetch64(s + 8);
  const u64 c = fetch64(s + n - 8) * mul;
  const u64 d = fetch64(s + n - 16) * K2;
  return hl16(rot(a + b, 43) + rot(c, 30) + d, a + rot(b + K2, 18) + c, mul);
}
__device__ u64 len33to64(const uint8_t* s, int n) {
  const u64 mul = K2 + static_cast<u64>(n) * 2;
  const u64 a = fetch64(s) * K2;
  const u64 b = fetch64(s + 8);
  const u64 c = fetch64(s + n - 8) * mul;
  const u64 d = fetch64(s + n - 16) * K2;
  const u64 y = rot(a + b, 43) + rot(c, 30) + d;
  const u64 z = hl16(y, a + rot(b + K2, 18) + c, mul);
  const u64 e = fetch64(s + 16) * mul;
  const u64 f = fetch64(s + 24);
  const u64 g = (y + fetch64(s + n - 32)) * mul;
  const u64 h = (z + fetch64(s + n - 24)) * mul;
  return hl16(rot(e + f, 43) + rot(g, 30) + h, e + rot(f + a, 18) + g, mul);
}
__device__ __forceinline__ void weak32(const uint8_t* s, u64 a, u64 b, u64* o1, u64* o2) {
  const u64 w = fetch64(s), x = fetch64(s + 8), y = fetch64(s + 16), z = fetch64(s + 24);
  a += w;
  b = rot(b + a + z, 21);
  const u64 c = a;
  a += x;
  a += y;
  b += rot(a, 44);
  *o1 = a + z;
  *o2 = b + c;
}
__device__ u64 fingerprint64(const uint8_t* s, int n) {
  if (n <= 16) return len0to16(s, n);
  if (n <= 32) return len17to32(s, n);
  if (n <= 64) return len33to64(s, n);
  const u64 seed = 81;
  u64 x = seed;
  u64 y = seed * K1 + 113;
  u64 z = smix(y * K2 + 113) * K2;
  u64 v1 = 0, v2 = 0, w1 = 0, w2 = 0;
  x = x * K2 + fetch64(s);
  const uint8_t* end = s + ((n - 1) / 64) * 64;
  const uint8_t* last64 = end + ((n - 1) & 63) - 63;
  do {
    x = rot(x + y + v1 + fetch64(s + 8), 37) * K1;
    y = rot(y + v2 + fetch64(s + 48), 42) * K1;
    x ^= w2;
    y += v1 + fetch64(s + 40);
    z = rot(z + w1, 33) * K1;
    weak32(s, v2 * K1, x + w1, &v1, &v2);
    weak32(s + 32, z + w2, y + fetch64(s + 16), &w1, &w2);
    const u64 t = z;
    z = x;
    x = t;
    s += 64;
  } while (s != end);
  const u64 mul = K1 + ((z & 0xff) << 1);
  s = last64;
  w1 += ((n - 1) & 63);
  v1 += w1;
  w1 += v1;
  x = rot(x + y + v1 + fetch64(s + 8), 37) * mul;
  y = rot(y + v2 + fetch64(s + 48), 42) * mul;
  x ^= w2 * 9;
  y += v1 * 9 + fetch64(s + 40);
  z = rot(z + w1, 33) * mul;
  weak32(s, v2 * mul, x + w1, &v1, &v2);
  weak32(s + 32, z + w2, y + fetch64(s + 16), &w1, &w2);
  const u64 t = z;
  z = x;
  x = t;
  return hl16(hl16(v1, w1, mul) + smix(y) * K0 + z, hl16(v2, w2, mul) + x, mul);
}

__global__ void hash_strings_kernel(const uint8_t* __restrict__ bytes,
                                    const int* __restrict__ offsets, long long N,
                                    const int* __restrict__ field_of,
                                    const int* __restrict__ n_buckets,
                                    const int* __restrict__ row_offset, int* __restrict__ out) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < N;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int o = offsets[i];
    const int n = offsets[i + 1] - o;
    const int f = field_of != nullptr ? field_of[i] : 0;
    const u64 h = fingerprint64(bytes + o, n);
    out[i] = row_offset[f] + static_cast<int>(h % static_cast<u64>(n_buckets[f]));
  }
}

// categorical_column_with_hash_bucket(dtype=int64) (deepfm/deepfm.py:41,46) [TF-sem]: the integer key
// is formatted with as_string ("%lld") and hashed like any other string.
__global__ void hash_int64_kernel(const long long* __restrict__ ids, long long N, int n_buckets,
                                  long long* __restrict__ out) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < N;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    uint8_t buf[24];
    const long long v = ids[i];
    u64 mag = v < 0 ? 0ull - static_cast<u64>(v) : static_cast<u64>(v);
    int n = 0;
    uint8_t rev[20];
    do {
      rev[n++] = static_cast<uint8_t>('0' + mag % 10ull);
      mag /= 10ull;
    } while (mag != 0ull);
    int len = 0;
    if (v < 0) buf[len++] = '-';
    while (n > 0) buf[len++] = rev[--n];
    out[i] = static_cast<long long>(fingerprint64(buf, len) % static_cast<u64>(n_buckets));
  }
}

// Fixed-width string slots as the record parser (records.cu) lays them out: string i occupies
// bytes[i*slot, i*slot + lens[i]), belongs to field i % n_fields; out = local id as int64, i.e.
// directly the `xcat` operand of the id pipeline.
__global__ void hash_slots_kernel(const uint8_t* __restrict__ bytes, int slot,
                                  const int* __restrict__ lens, long long N, int n_fields,
                                  const int* __restrict__ n_buckets, long long* __restrict__ out) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < N;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int f = static_cast<int>(i % n_fields);
    const u64 h = fingerprint64(bytes + i * slot, min(max(lens[i], 0), slot));
    out[i] = static_cast<long long>(h % static_cast<u64>(n_buckets[f]));
  }
}

}  // namespace ctr

extern "C" int ctr_hash_slots(const uint8_t* bytes, int slot, const int32_t* lens, int64_t N,
                              int n_fields, const int32_t* n_buckets_dev, int64_t* out,
                              ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(bytes && lens && n_buckets_dev && out && N >= 0 && slot > 0 && n_fields > 0,
              "ctr_hash_slots", "bad argument");
  if (N == 0) return CTR_OK;
  const int grid =
      static_cast<int>(std::min<long long>((N + 255) / 256, ctr::sm_count() * 8LL));
  ctr::hash_slots_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      bytes, slot, lens, N, n_fields, n_buckets_dev, reinterpret_cast<long long*>(out));
  CTR_LAUNCH_CHECK("ctr_hash_slots");
}

extern "C" int ctr_hash_int64(const int64_t* ids, int64_t N, int32_t n_buckets, int64_t* out,
                              ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(ids && out && N >= 0 && n_buckets > 0, "ctr_hash_int64", "bad argument");
  if (N == 0) return CTR_OK;
  const int grid =
      static_cast<int>(std::min<long long>((N + 255) / 256, ctr::sm_count() * 8LL));
  ctr::hash_int64_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(ids), N, n_buckets, reinterpret_cast<long long*>(out));
  CTR_LAUNCH_CHECK("ctr_hash_int64");
}

extern "C" int ctr_hash_strings(const uint8_t* bytes, const int32_t* offsets, int64_t N,
                                const int32_t* field_of, const int32_t* n_buckets_dev,
                                const int32_t* row_offset_dev, int32_t* out, ctr_stream_t stream) {
  CTR_ARCH_OR_RETURN();
  CTR_REQUIRE(bytes && offsets && n_buckets_dev && row_offset_dev && out && N >= 0,
              "ctr_hash_strings", "null pointer / negative N");
  if (N == 0) return CTR_OK;
  const int grid =
      static_cast<int>(std::min<long long>((N + 255) / 256, ctr::sm_count() * 8LL));
  ctr::hash_strings_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      bytes, offsets, N, field_of, n_buckets_dev, row_offset_dev, out);
  CTR_LAUNCH_CHECK("ctr_hash_strings");
}
