// Host side of the input pipeline: TFRecord framing + tf.train.Example decoding for the two
// schemas on the hot path, multi-threaded, straight into caller-owned (pinned) batch buffers.
//
//   fm/fm.py:100-112     TFRecordDataset -> parse_single_example(feature_description) -> batch
//   fm/fm.py:39-44       Criteo schema: _c0.._c13 float32 [1], _c14.._c39 string [1] default 'NULL'
//   din/din.py:43-60     DIN schema: label, i_id, i_cate int64 scalars; u_iid_seq, u_icat_seq
//                        var-len int64, densified per record
//
// No CUDA here (the file is .cu only so that the one nvcc build picks it up): these entry points
// run on host threads and work without a GPU.  The categorical strings are NOT hashed on the host:
// they are laid out in fixed-width slots that ctr_hash_slots (hash.cu) fingerprints on the device.
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "common.cuh"

namespace ctr {

// ------------------------------------------------------------------ crc32c (Castagnoli)
struct Crc32cTable {
  uint32_t t[8][256];
  Crc32cTable() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xFFu];
  }
};
static const Crc32cTable& crc_table() {
  static const Crc32cTable tab;
  return tab;
}
static uint32_t crc32c(const uint8_t* p, size_t n) {
  const Crc32cTable& T = crc_table();
  uint32_t c = 0xFFFFFFFFu;
  while (n >= 8) {
    uint32_t lo, hi;
    memcpy(&lo, p, 4);
    memcpy(&hi, p + 4, 4);
    lo ^= c;
    c = T.t[7][lo & 0xFF] ^ T.t[6][(lo >> 8) & 0xFF] ^ T.t[5][(lo >> 16) & 0xFF] ^ T.t[4][lo >> 24] ^
        T.t[3][hi & 0xFF] ^ T.t[2][(hi >> 8) & 0xFF] ^ T.t[1][(hi >> 16) & 0xFF] ^ T.t[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) c = T.t[0][(c ^ *p++) & 0xFFu] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}
// TFRecord stores crcs "masked": rotate right by 15 and add a constant.
static uint32_t masked_crc(const uint8_t* p, size_t n) {
  const uint32_t c = crc32c(p, n);
  return ((c >> 15) | (c << 17)) + 0xa282ead8u;
}

// ------------------------------------------------------------------ protobuf wire helpers
struct Span {
  const uint8_t* p;
  const uint8_t* e;
  bool ok() const { return p != nullptr; }
};
static inline bool rd_varint(const uint8_t*& p, const uint8_t* e, uint64_t* v) {
  uint64_t r = 0;
  for (int s = 0; s < 70 && p < e; s += 7) {
    const uint8_t c = *p++;
    r |= static_cast<uint64_t>(c & 0x7F) << s;
    if (c < 0x80) {
      *v = r;
      return true;
    }
  }
  return false;
}
static inline bool rd_len(const uint8_t*& p, const uint8_t* e, Span* out) {
  uint64_t n;
  if (!rd_varint(p, e, &n) || n > static_cast<uint64_t>(e - p)) return false;
  out->p = p;
  out->e = p + n;
  p += n;
  return true;
}
static inline bool skip_field(const uint8_t*& p, const uint8_t* e, uint32_t wire) {
  uint64_t v;
  Span s;
  switch (wire) {
    case 0: return rd_varint(p, e, &v);
    case 1: if (e - p < 8) return false; p += 8; return true;
    case 2: return rd_len(p, e, &s);
    case 5: if (e - p < 4) return false; p += 4; return true;
    default: return false;
  }
}

// One map entry of Features.feature: key + the Feature's one-of list.
struct Entry {
  Span key;
  int kind;     // 1 = bytes_list, 2 = float_list, 3 = int64_list, 0 = empty Feature
  Span list;    // the *List message body
};
// Calls fn(entry) for every feature of a serialized tf.train.Example.  false = malformed.
template <class Fn>
static bool for_each_feature(const uint8_t* p, const uint8_t* e, Fn&& fn) {
  while (p < e) {
    uint64_t tag;
    if (!rd_varint(p, e, &tag)) return false;
    if (tag != 0x0A) {            // Example.features = 1
      if (!skip_field(p, e, tag & 7)) return false;
      continue;
    }
    Span feats;
    if (!rd_len(p, e, &feats)) return false;
    const uint8_t* q = feats.p;
    while (q < feats.e) {
      if (!rd_varint(q, feats.e, &tag)) return false;
      if (tag != 0x0A) {          // Features.feature = 1 (map entries)
        if (!skip_field(q, feats.e, tag & 7)) return false;
        continue;
      }
      Span ent;
      if (!rd_len(q, feats.e, &ent)) return false;
      Entry en{{nullptr, nullptr}, 0, {nullptr, nullptr}};
      const uint8_t* r = ent.p;
      while (r < ent.e) {
        if (!rd_varint(r, ent.e, &tag)) return false;
        if (tag == 0x0A) {        // key
          if (!rd_len(r, ent.e, &en.key)) return false;
        } else if (tag == 0x12) { // value: Feature
          Span feat;
          if (!rd_len(r, ent.e, &feat)) return false;
          const uint8_t* s = feat.p;
          while (s < feat.e) {
            uint64_t t2;
            if (!rd_varint(s, feat.e, &t2)) return false;
            if ((t2 & 7) == 2 && (t2 >> 3) >= 1 && (t2 >> 3) <= 3) {
              if (!rd_len(s, feat.e, &en.list)) return false;
              en.kind = static_cast<int>(t2 >> 3);
            } else if (!skip_field(s, feat.e, t2 & 7)) {
              return false;
            }
          }
        } else if (!skip_field(r, ent.e, tag & 7)) {
          return false;
        }
      }
      if (en.key.ok() && !fn(en)) return false;
    }
  }
  return true;
}
// first value of a FloatList (packed or not)
static bool first_float(Span l, float* out) {
  const uint8_t* p = l.p;
  while (p < l.e) {
    uint64_t tag;
    if (!rd_varint(p, l.e, &tag)) return false;
    if (tag == 0x0A) {            // packed
      Span pk;
      if (!rd_len(p, l.e, &pk)) return false;
      if (pk.e - pk.p >= 4) {
        memcpy(out, pk.p, 4);
        return true;
      }
    } else if (tag == 0x0D) {     // fixed32
      if (l.e - p < 4) return false;
      memcpy(out, p, 4);
      return true;
    } else if (!skip_field(p, l.e, tag & 7)) {
      return false;
    }
  }
  return false;
}
static bool first_bytes(Span l, Span* out) {
  const uint8_t* p = l.p;
  while (p < l.e) {
    uint64_t tag;
    if (!rd_varint(p, l.e, &tag)) return false;
    if (tag == 0x0A) return rd_len(p, l.e, out);
    if (!skip_field(p, l.e, tag & 7)) return false;
  }
  return false;
}
// all values of an Int64List appended to out (max_n bounds the write); returns count or -1
static int64_t all_int64(Span l, int64_t* out, int64_t max_n) {
  int64_t n = 0;
  const uint8_t* p = l.p;
  while (p < l.e) {
    uint64_t tag, v;
    if (!rd_varint(p, l.e, &tag)) return -1;
    if (tag == 0x0A) {            // packed
      Span pk;
      if (!rd_len(p, l.e, &pk)) return -1;
      const uint8_t* q = pk.p;
      while (q < pk.e) {
        if (!rd_varint(q, pk.e, &v)) return -1;
        if (out != nullptr && n < max_n) out[n] = static_cast<int64_t>(v);
        ++n;
      }
    } else if (tag == 0x08) {
      if (!rd_varint(p, l.e, &v)) return -1;
      if (out != nullptr && n < max_n) out[n] = static_cast<int64_t>(v);
      ++n;
    } else if (!skip_field(p, l.e, tag & 7)) {
      return -1;
    }
  }
  return n;
}

// "_c<digits>" -> the number, else -1
static inline int criteo_key(Span k) {
  const int64_t n = k.e - k.p;
  if (n < 3 || n > 4 || k.p[0] != '_' || k.p[1] != 'c') return -1;
  int v = 0;
  for (const uint8_t* p = k.p + 2; p < k.e; ++p) {
    if (*p < '0' || *p > '9') return -1;
    v = v * 10 + (*p - '0');
  }
  return v;
}
static inline bool key_is(Span k, const char* s) {
  const size_t n = strlen(s);
  return static_cast<size_t>(k.e - k.p) == n && memcmp(k.p, s, n) == 0;
}

template <class Fn>
static int run_parallel(int64_t n, int n_threads, Fn&& fn) {
  if (n_threads <= 0) n_threads = static_cast<int>(std::thread::hardware_concurrency());
  n_threads = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(n_threads, (n + 255) / 256)));
  std::atomic<int64_t> bad(-1);
  auto work = [&](int t) {
    const int64_t lo = n * t / n_threads, hi = n * (t + 1) / n_threads;
    for (int64_t i = lo; i < hi; ++i)
      if (!fn(i)) {
        int64_t exp = -1;
        bad.compare_exchange_strong(exp, i);
        return;
      }
  };
  if (n_threads == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
    for (auto& x : th) x.join();
  }
  const int64_t b = bad.load();
  if (b >= 0) {
    set_error("malformed tf.train.Example at record " + std::to_string(b));
    return CTR_ERR_DATA;
  }
  return CTR_OK;
}

}  // namespace ctr

using namespace ctr;

extern "C" {

int64_t ctr_tfrecord_scan(const uint8_t* buf, int64_t nbytes, int verify_crc, int64_t* payload_off,
                          int32_t* payload_len, int64_t max_records) {
  if (buf == nullptr || nbytes < 0) return fail_arg("ctr_tfrecord_scan", "null buffer");
  int64_t pos = 0, n = 0;
  while (pos < nbytes) {
    if (nbytes - pos < 12) {
      set_error("ctr_tfrecord_scan: truncated record header at byte " + std::to_string(pos));
      return CTR_ERR_DATA;
    }
    uint64_t len;
    uint32_t crc;
    memcpy(&len, buf + pos, 8);
    memcpy(&crc, buf + pos + 8, 4);
    if (verify_crc && masked_crc(buf + pos, 8) != crc) {
      set_error("ctr_tfrecord_scan: length crc mismatch at record " + std::to_string(n));
      return CTR_ERR_DATA;
    }
    if (len > static_cast<uint64_t>(nbytes - pos - 16) || len > 0x7FFFFFFFull) {
      set_error("ctr_tfrecord_scan: truncated record payload at record " + std::to_string(n));
      return CTR_ERR_DATA;
    }
    if (verify_crc) {
      uint32_t pc;
      memcpy(&pc, buf + pos + 12 + len, 4);
      if (masked_crc(buf + pos + 12, len) != pc) {
        set_error("ctr_tfrecord_scan: payload crc mismatch at record " + std::to_string(n));
        return CTR_ERR_DATA;
      }
    }
    if (payload_off != nullptr && n < max_records) {
      payload_off[n] = pos + 12;
      payload_len[n] = static_cast<int32_t>(len);
    }
    ++n;
    pos += 16 + static_cast<int64_t>(len);
  }
  return n;
}

int ctr_criteo_parse(const uint8_t* buf, const int64_t* payload_off, const int32_t* payload_len,
                     int64_t n, int n_threads, float* labels, float* cont, uint8_t* cat_bytes,
                     int32_t* cat_len, int slot) {
  CTR_REQUIRE(buf && payload_off && payload_len && labels && cont && cat_bytes && cat_len && n >= 0,
              "ctr_criteo_parse", "null pointer");
  CTR_REQUIRE(slot >= 8 && slot <= 64 && (slot & 7) == 0, "ctr_criteo_parse",
              "slot must be a multiple of 8 in [8, 64]");
  std::atomic<int> too_long(0), missing(0);
  const int rc = run_parallel(n, n_threads, [&](int64_t i) -> bool {
    const uint8_t* p = buf + payload_off[i];
    float* c = cont + i * 13;
    uint8_t* sb = cat_bytes + i * 26 * slot;
    int32_t* sl = cat_len + i * 26;
    uint64_t seen = 0;
    const bool ok = for_each_feature(p, p + payload_len[i], [&](const Entry& en) -> bool {
      const int k = criteo_key(en.key);
      if (k < 0 || k > 39) return true;                // not in feature_description: ignored
      if (k <= 13) {
        float v;
        if (en.kind != 2 || !first_float(en.list, &v)) return true;   // stays "missing"
        if (k == 0) labels[i] = v; else c[k - 1] = v;
        seen |= 1ull << k;
      } else {
        Span s;
        if (en.kind != 1 || !first_bytes(en.list, &s)) return true;   // empty -> default 'NULL'
        const int64_t len = s.e - s.p;
        if (len > slot) {
          too_long.store(1);
          return true;
        }
        uint8_t* d = sb + (k - 14) * slot;
        memcpy(d, s.p, len);
        memset(d + len, 0, slot - len);
        sl[k - 14] = static_cast<int32_t>(len);
        seen |= 1ull << k;
      }
      return true;
    });
    if (!ok) return false;
    if ((seen & 0x3FFFull) != 0x3FFFull) missing.store(1);   // FixedLenFeature float without default
    for (int k = 14; k < 40; ++k)
      if (!((seen >> k) & 1ull)) {                     // default_value 'NULL' (fm/fm.py:44)
        uint8_t* d = sb + (k - 14) * slot;
        memset(d, 0, slot);
        memcpy(d, "NULL", 4);
        sl[k - 14] = 4;
      }
    return true;
  });
  if (rc != CTR_OK) return rc;
  if (missing.load()) {
    set_error("ctr_criteo_parse: a record lacks one of the float features _c0.._c13 (no default "
              "in feature_description, fm/fm.py:43)");
    return CTR_ERR_DATA;
  }
  if (too_long.load()) {
    set_error("ctr_criteo_parse: a categorical string is longer than the slot");
    return CTR_ERR_DATA;
  }
  return CTR_OK;
}

int64_t ctr_din_parse(const uint8_t* buf, const int64_t* payload_off, const int32_t* payload_len,
                      int64_t n, int n_threads, int64_t P, int64_t* labels, int64_t* i_id,
                      int64_t* i_cate, int64_t* u_iid_seq, int64_t* u_icat_seq) {
  CTR_REQUIRE(buf && payload_off && payload_len && n >= 0, "ctr_din_parse", "null pointer");
  if (n == 0) return 0;
  if (labels == nullptr) {
    // query: history length of the first record (din/din.py:73 batches densified var-len
    // features with .batch(), so every record of a batch must have this length)
    const uint8_t* p = buf + payload_off[0];
    int64_t len = 0;
    const bool ok = for_each_feature(p, p + payload_len[0], [&](const Entry& en) -> bool {
      if (key_is(en.key, "u_iid_seq") && en.kind == 3) len = all_int64(en.list, nullptr, 0);
      return len >= 0;
    });
    if (!ok) {
      set_error("ctr_din_parse: malformed first record");
      return CTR_ERR_DATA;
    }
    return len;
  }
  CTR_REQUIRE(i_id && i_cate && u_iid_seq && u_icat_seq && P >= 0, "ctr_din_parse", "null output");
  std::atomic<int> ragged(0), missing(0);
  const int rc = run_parallel(n, n_threads, [&](int64_t i) -> bool {
    const uint8_t* p = buf + payload_off[i];
    int seen = 0;
    int64_t n1 = 0, n2 = 0;
    const bool ok = for_each_feature(p, p + payload_len[i], [&](const Entry& en) -> bool {
      if (en.kind != 3) return true;
      int64_t v;
      if (key_is(en.key, "label")) { if (all_int64(en.list, &v, 1) >= 1) { labels[i] = v; seen |= 1; } }
      else if (key_is(en.key, "i_id")) { if (all_int64(en.list, &v, 1) >= 1) { i_id[i] = v; seen |= 2; } }
      else if (key_is(en.key, "i_cate")) { if (all_int64(en.list, &v, 1) >= 1) { i_cate[i] = v; seen |= 4; } }
      else if (key_is(en.key, "u_iid_seq")) n1 = all_int64(en.list, u_iid_seq + i * P, P);
      else if (key_is(en.key, "u_icat_seq")) n2 = all_int64(en.list, u_icat_seq + i * P, P);
      return n1 >= 0 && n2 >= 0;
    });
    if (!ok) return false;
    if (seen != 7) missing.store(1);
    if (n1 != P || n2 != P) ragged.store(1);
    return true;
  });
  if (rc != CTR_OK) return rc;
  if (missing.load()) {
    set_error("ctr_din_parse: a record lacks label / i_id / i_cate (FixedLenFeature without default)");
    return CTR_ERR_DATA;
  }
  if (ragged.load()) {
    set_error("ctr_din_parse: history lengths differ inside a batch (din/din.py:73 uses .batch(), "
              "not padded_batch: every record must carry the same number of history ids)");
    return CTR_ERR_DATA;
  }
  return P;
}

uint32_t ctr_masked_crc32c(const uint8_t* data, int64_t n) {
  return masked_crc(data, static_cast<size_t>(n));
}

}  // extern "C"
