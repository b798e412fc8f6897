/* ctr_b200.h - C ABI of libctr_b200.so: the B200 (sm_100a) CTR embedding +
 * feature-interaction hot path.
 *
 * The reference (wangruichens/recsys) has no FFI: its hot path is the TensorFlow
 * graph each model_fn builds.  Every entry point below replaces the TF ops of one
 * reference call site (cited per function, paths relative to the reference root);
 * INTEGRATION.md shows the ctypes / tf.load_op_library stubs that bind them.
 *
 * Conventions
 *  - plain C: raw device pointers, sizes, a cudaStream_t passed as void*.
 *  - returns 0 (CTR_OK) or a negative code; ctr_last_error() has the text
 *    (thread-local).  Never throws, never allocates device memory, never
 *    synchronises: work is enqueued on `stream`; the caller owns every buffer.
 *  - no CPU fallback: on a device that is not compute capability 10.x every
 *    compute entry point returns CTR_ERR_ARCH.
 *  - all float tensors fp32, row-major, 16-byte aligned; row ids int32 (global
 *    row of the concatenated table [R, D]); "F axis" = the reference's
 *    input_layer order (columns sorted by name).
 */
#ifndef CTR_B200_H_
#define CTR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTR_OK 0
#define CTR_ERR_ARG (-1)   /* bad argument (null, misaligned, unsupported size) */
#define CTR_ERR_CUDA (-2)  /* CUDA runtime error, text in ctr_last_error()      */
#define CTR_ERR_ARCH (-3)  /* current device is not sm_100                      */
#define CTR_ERR_DATA (-4)  /* corrupt / malformed input records (host-side input pipeline) */

#define CTR_MAX_FIELDS 64

typedef void* ctr_stream_t; /* cudaStream_t */

int ctr_version(void);
const char* ctr_last_error(void);
/* 0 when the current CUDA device can run this library (cc 10.x). */
int ctr_device_check(void);
/* Process-wide tuning options (results never depend on them):
 *   "bwd_aggregate" (default 1): ctr_embed_bwd sums the slots of one warp instruction that hit the
 *   same row in registers (__match_any_sync) and issues one RED per distinct row - the
 *   warp-aggregated scatter-add; 0 = one RED per slot (the A/B switch of bench.py --dist zipf).
 *   "adam_rows_bf" (reserved, default 1).
 *   "fwd_prefetch_record" (default 0): ctr_embed_fwd_raw over a row-record table also prefetches
 *   the rest of every looked-up record (moments, gradient accumulator) into L2, for the
 *   scatter-add and the row optimiser that follow in the same step (measured: no gain, the
 *   scatter and the optimiser are not DRAM bound at batch 4096).
 *   "mid_coop" (default 1): ctr_tower_mid training launches are cooperative; 0 = plain launch
 *   (equally safe while nothing resident on the device waits on that kernel).
 *   "tcg_dw_stages" (default 2), "tcg_dw_splits" (default 0 = automatic): TMA ring depth and
 *   split-K factor of the first tower layer's weight-gradient GEMM (it runs beside the scatter);
 *   "tower_dw_splits" (default 0 = automatic): row splits of ctr_tower_layer_bwd_weights.
 *   "adam_rows_inflight" (default 1): records in flight per lane group in ctr_adam_rows_bf
 *   (1 or 2). */
int ctr_set_option(const char* name, int value);

/* ---------------------------------------------------------------- id pipeline
 * Replaces the id-producing half of tf.feature_column.input_layer:
 *   numeric_column(normalizer_fn=log(x+off)) -> bucketized_column(boundaries)
 *   (fm/fm.py:76-80) and categorical_column_with_hash_bucket (fm/fm.py:89) for
 *   already-hashed local ids.
 * One descriptor per field, in F-axis order. */
typedef struct {
  int32_t kind;       /* 0 = log-bucketised numeric, 1 = pre-hashed categorical */
  int32_t src;        /* column of xcont (kind 0) / xcat (kind 1)               */
  int32_t n_rows;     /* rows of this field's table                             */
  int32_t row_offset; /* first row of the field in the concatenated table       */
  int32_t bnd_begin;  /* kind 0: first boundary in `boundaries`                 */
  int32_t bnd_count;  /* kind 0: number of boundaries (n_rows - 1)              */
  float log_offset;   /* kind 0: 1.0 (4.0 for _c2, fm/fm.py:77-78)              */
  int32_t pad_;
} ctr_field_desc;

/* rows[b,f] = row_offset_f + (kind0: upper_bound(boundaries_f, logf(x+off)) | kind1: xcat id).
 * logx (nullable): [B, n_cont] the log-normalised numerics (xdeepfm linear part,
 * xdeepfm/xdeepfm.py:82).  status (nullable, device int): bit0 set if any
 * categorical id was outside [0, n_rows) (it is wrapped by modulo). */
int ctr_criteo_rows(const float* xcont, int n_cont, const int64_t* xcat, int n_cat,
                    const ctr_field_desc* fields_dev, const float* boundaries_dev, int B, int F,
                    int32_t* rows, float* logx, int32_t* status, ctr_stream_t stream);

/* The same ids from at most max_ctas CTAs (0 = 16): the variant for a copy stream beside a running
 * training step, whose kernels want whole SMs - a full-width id kernel that reaches an SM first
 * keeps them waiting (estimator.GraphedTrainStep runs the ids of batch s+1 beside step s). */
int ctr_criteo_rows_bg(const float* xcont, int n_cont, const int64_t* xcat, int n_cat,
                       const ctr_field_desc* fields_dev, const float* boundaries_dev, int n_boundaries,
                       int B, int F, int32_t* rows, int32_t* status, int max_ctas,
                       ctr_stream_t stream);

/* FarmHash Fingerprint64(bytes) mod n_buckets for N strings (TF StringToHashBucketFast,
 * fm/fm.py:89).  bytes: concatenated strings; offsets[N+1]; field_of[N] (nullable) selects
 * n_buckets[field] and row_offset[field]; out[i] = row_offset + hash % n_buckets. */
int ctr_hash_strings(const uint8_t* bytes, const int32_t* offsets, int64_t N,
                     const int32_t* field_of, const int32_t* n_buckets_dev,
                     const int32_t* row_offset_dev, int32_t* out, ctr_stream_t stream);

/* categorical_column_with_hash_bucket(dtype=int64) (deepfm/deepfm.py:41,46) [TF-sem]: the key is
 * formatted as_string ("%lld") and hashed: out[i] = Fingerprint64(decimal(ids[i])) mod n_buckets. */
int ctr_hash_int64(const int64_t* ids, int64_t N, int32_t n_buckets, int64_t* out,
                   ctr_stream_t stream);

/* The same hash over fixed-width slots, the layout ctr_criteo_parse produces: string i occupies
 * bytes[i*slot, i*slot + lens[i]) and belongs to field i % n_fields;
 * out[i] = Fingerprint64 mod n_buckets[field] as int64 - directly the xcat operand [B, n_fields] of
 * ctr_criteo_rows / ctr_embed_fwd_raw. */
int ctr_hash_slots(const uint8_t* bytes, int slot, const int32_t* lens, int64_t N, int n_fields,
                   const int32_t* n_buckets_dev, int64_t* out, ctr_stream_t stream);

/* ------------------------------------------------- input pipeline (HOST side, no GPU needed)
 * Replaces TFRecordDataset + parse_single_example + batch of `input_fn` (fm/fm.py:100-112,
 * din/din.py:52-80) with a multi-threaded decoder that fills caller-owned (pinned) batch buffers.
 * All pointers are HOST pointers.
 *
 * ctr_tfrecord_scan: walk a buffer of TFRecord frames {u64 len | u32 masked crc32c(len) | payload |
 *   u32 masked crc32c(payload)}; payload_off/payload_len (nullable) receive the first max_records
 *   frames.  Returns the number of frames, or CTR_ERR_DATA (truncated frame; crc mismatch when
 *   verify_crc != 0).  ctr_masked_crc32c is the checksum it verifies.
 * ctr_criteo_parse: decode n tf.train.Example payloads with the Criteo feature_description
 *   (fm/fm.py:39-44): labels[n] = _c0; cont[n,13] = _c1.._c13; the 26 strings _c14.._c39 go to
 *   cat_bytes[n,26,slot] (zero padded) with cat_len[n,26]; a missing / empty string feature gets the
 *   default 'NULL' (fm/fm.py:44); a missing float feature is an error, as in TF.  slot: multiple
 *   of 8 in [8,64].  n_threads <= 0: all hardware threads.
 * ctr_din_parse: decode n payloads with DIN's feature_description (din/din.py:43-50).  With
 *   labels == NULL it only returns the history length of the first record; otherwise fills labels,
 *   i_id, i_cate [n] and the densified histories u_iid_seq, u_icat_seq [n,P]; every record must
 *   carry exactly P ids in both histories (the reference uses .batch(), not padded_batch,
 *   din/din.py:73), else CTR_ERR_DATA. */
int64_t ctr_tfrecord_scan(const uint8_t* buf, int64_t nbytes, int verify_crc, int64_t* payload_off,
                          int32_t* payload_len, int64_t max_records);
uint32_t ctr_masked_crc32c(const uint8_t* data, int64_t n);
int ctr_criteo_parse(const uint8_t* buf, const int64_t* payload_off, const int32_t* payload_len,
                     int64_t n, int n_threads, float* labels, float* cont, uint8_t* cat_bytes,
                     int32_t* cat_len, int slot);
int64_t ctr_din_parse(const uint8_t* buf, const int64_t* payload_off, const int32_t* payload_len,
                      int64_t n, int n_threads, int64_t P, int64_t* labels, int64_t* i_id,
                      int64_t* i_cate, int64_t* u_iid_seq, int64_t* u_icat_seq);

/* --------------------------------------------------- fused multi-field lookup
 * Forward.  Replaces input_layer(embedding columns) + input_layer(indicator
 * columns) + the FM second-order block + (optionally) the DCN cross stack:
 *   fm/fm.py:117-129, deepfm/deepfm.py:84-98, xdeepfm/xdeepfm.py:127-128,185,
 *   dcn/dcn.py:122-142.
 *   E[b, f*D:(f+1)*D] = table[rows[b,f]]                                 (nullable)
 *   S[b, :]  = sum_f table[rows[b,f]]                                    (nullable)
 *   y1[b]    = sum_{f in w1_fields} w1[rows[b,f]]   (pre-bias, pre-ReLU) (nullable)
 *   y2[b]    = 0.5 * sum_d[(sum_f E)^2 - sum_f E^2]                      (nullable)
 *   xl[b,:]  = cross stack on x0 = E[b,:]: xl <- (xl.w_l) x0 + xl + b_l  (nullable)
 *   E_lo     = the lo half of E's 3xTF32 split (see ctr_split_lo)        (nullable)
 * D in {8,16,32}; F <= 64 and F*D <= 1280; w1_fields bit f = field f has a first-order weight. */
int ctr_embed_fwd(const float* table, const float* w1, const int32_t* rows, int B, int F, int D,
                  uint64_t w1_fields, float* E, float* S, float* y1, float* y2,
                  const float* cross_w, const float* cross_b, int cross_layers, float* xl, float* E_lo,
                  int64_t row_stride, int64_t w1_stride, ctr_stream_t stream);

/* Same, with the id pipeline of ctr_criteo_rows fused in front (one launch instead of two): every
 * CTA computes the row ids of its sample tile from the raw features straight into shared memory
 * and writes them to rows_out [B,F] for the backward / optimiser.  boundaries: n_boundaries <= 512.
 * zero_buf (nullable) / zero_n floats: a buffer the launch clears on the side (the tower's
 * per-step accumulators: BN column sums, loss, split-K target), sparing the step a memset. */
int ctr_embed_fwd_raw(const float* table, const float* w1, const float* xcont, int n_cont,
                      const int64_t* xcat, int n_cat, const ctr_field_desc* fields_dev,
                      const float* boundaries_dev, int n_boundaries, int32_t* rows_out, float* logx,
                      int32_t* status, int B, int F, int D, uint64_t w1_fields, float* E, float* S,
                      float* y1, float* y2, const float* cross_w, const float* cross_b,
                      int cross_layers, float* xl, float* E_lo, int64_t row_stride,
                      int64_t w1_stride, float* zero_buf, int64_t zero_n, ctr_stream_t stream);

/* ctr_embed_fwd_raw and the first dense layer of the tower in ONE launch (D = 16, F <= 40,
 * 16 <= N <= 128): act0[B, N] = relu(E . W0 + b0) (deepfm/deepfm.py:101) is formed from the
 * gathered rows while they sit in shared memory - a cluster of 4 CTAs per 128 samples, each CTA
 * gathers a quarter of the fields straight into the swizzled A tiles of a tcgen05 3xTF32 GEMM
 * against TMA-fed k-blocks of W0 / W0_lo (the ctr_split_lo pair), and the four partial products
 * and FM partial sums are added over distributed shared memory.  E / E_lo are still written (the
 * weight-gradient GEMM of the backward reads them), S / y1 / y2 as by ctr_embed_fwd.
 * stats_part (nullable): [ceil(B/128)][2][N] column sums of act0 and act0^2 per cluster, for
 * ctr_tower_mid_args.stats0_part.  zero_buf as in ctr_embed_fwd_raw.
 * rows_in (nullable): the [B, F] ids when they were computed ahead of the step (ctr_criteo_rows on a
 * copy stream, beside the previous step); the id stage and rows_out are then skipped. */
int ctr_embed_tower_fwd(const float* table, const float* w1, const float* xcont, int n_cont,
                        const int64_t* xcat, int n_cat, const ctr_field_desc* fields_dev,
                        const float* boundaries_dev, int n_boundaries, const int32_t* rows_in,
                        int32_t* rows_out, int32_t* status, int B,
                        int F, int D, uint64_t w1_fields, float* E, float* E_lo, float* S, float* y1,
                        float* y2, int64_t row_stride, int64_t w1_stride, const float* W0,
                        const float* W0_lo, const float* b0, int N, float* act0, float* stats_part,
                        float* zero_buf, int64_t zero_n, ctr_stream_t stream);

/* The backward counterpart: the first layer's data gradient dE = dpre0 . W0^T (deepfm/deepfm.py:101,
 * backward) and ctr_embed_bwd in ONE launch.  Each CTA forms dE for 128 samples x 10 fields in TMEM
 * (tcgen05 3xTF32 from the ctr_split_lo pairs dpre0 / dpre0_lo [B, N] and W0 / W0_lo [F*D, N]) and
 * scatters it from there: g[b,f,:] = dE[b,f,:] + dy2[b] * (S[b,:] - E[b,f,:]), warp-aggregated
 * vector REDs into dtable / dw1 exactly as ctr_embed_bwd (fields with <= 32 rows are summed per row
 * in shared memory first).  dE never touches global memory.  D = 16, 16 <= N <= 128. */
int ctr_tower_embed_bwd(const float* dpre0, const float* dpre0_lo, const float* W0, const float* W0_lo,
                        int N, const int32_t* rows, const float* E, const float* S, const float* dy2,
                        const float* dy1, uint64_t w1_fields, const int64_t* row_offsets_host, int B,
                        int F, int D, float* dtable, float* dw1, int64_t row_stride, int64_t w1_stride,
                        ctr_stream_t stream);

/* Profiling aid: later ctr_embed_tower_fwd launches stamp %globaltimer (ns) of CTA (0,0) into
 * timing_dev[0..7] at the phase boundaries (start | ids staged | loads issued | A tiles written |
 * MMAs retired | partials visible to the cluster | act0 written | end); NULL switches it off. */
int ctr_embed_tower_timing(uint64_t* timing_dev);

/* Backward: scatter-add of the row gradients into dtable / dw1 (the IndexedSlices
 * gradient of the gathers, fm/fm.py:162-163), field-major, contention-free for
 * fields with <= 32 rows.
 *   g[b,f,:] = dE[b,f,:] (if given) + dy2[b] * (S[b,:] - E[b,f,:]) (if dy2 given)
 *   dtable[rows[b,f], :] += g[b,f,:]  ;  dw1[rows[b,f]] += dy1[b] for f in w1_fields
 * E nullable (rows are re-gathered from `table`).  row_offsets: HOST int64[F+1]. */
int ctr_embed_bwd(const int32_t* rows, const float* dE, const float* E, const float* table,
                  const float* S, const float* dy2, const float* dy1, uint64_t w1_fields,
                  const int64_t* row_offsets_host, int B, int F, int D, float* dtable, float* dw1,
                  int64_t row_stride, int64_t w1_stride, ctr_stream_t stream);

/* DCN cross stack, stand-alone (dcn/dcn.py:132-142) on x0[B,W] (W = F*D, W%4==0, W<=1280). */
int ctr_dcn_cross_fwd(const float* x0, const float* w, const float* b, int L, int B, int W,
                      float* xl, ctr_stream_t stream);
/* dx0 = d(loss)/d(x0) through the cross stack (written, not accumulated);
 * dw[L,W], db[L,W] accumulated (+=). */
int ctr_dcn_cross_bwd(const float* x0, const float* w, const float* b, int L, int B, int W,
                      const float* dxl, float* dx0, float* dw, float* db, ctr_stream_t stream);

/* ------------------------------------------------------------------ optimiser
 * tf.train.AdamOptimizer semantics (fm/fm.py:162): eps outside the sqrt,
 *   lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t).
 * The schedule can live on the device so that a captured CUDA graph advances it on replay:
 * state_dev = float[4] {t = optimiser steps completed (a uint32 BIT PATTERN in the first word: a
 * float counter would saturate at 2^24 steps), lr_t of the step in progress (step t+1), lr, block
 * counter}.  When state_dev is non-null the kernels of a step read lr_t (and the claim
 * tag = t+1) from it and ignore the by-value arguments; the step's LAST optimiser launch -
 * ctr_adam_dense with advance_state != 0 - moves the schedule on (its last block to finish does
 * t += 1 and recomputes lr_t), so advancing costs no launch.  ctr_adam_tick(state, lr, ...) sets lr
 * and advances once: from t = 0xFFFFFFFF (and zeros) it initialises the schedule at t = 0; it is also the way to
 * advance for a caller whose step does not end with ctr_adam_dense.
 * ctr_adam_dense: every element (TF's sparse apply decays m, v of every row [TF-sem]).
 * g is zeroed afterwards when zero_g != 0. */
int ctr_adam_tick(float* state_dev, float lr, float beta1, float beta2, ctr_stream_t stream);
int ctr_adam_dense(float* theta, float* m, float* v, float* g, int64_t n, float lr_t, float beta1,
                   float beta2, float eps, int zero_g, float* state_dev, int advance_state,
                   ctr_stream_t stream);
/* The step's closing optimiser kernels may run CONCURRENTLY (the dense weights on one stream, the
 * touched rows on another): each takes advance_parties = the number of kernels that end the step
 * together (0 = takes no part, 1 = alone, as ctr_adam_dense's advance_state), and the last block
 * of the last one to finish moves the schedule on.  state_dev then has 8 words:
 * {t, lr_t, lr, dense block counter, parties finished, row block counter, 0, 0}.
 * ctr_adam_dense_ex also writes the 3xTF32 lo half (ctr_split_lo rule) of the updated slice
 * theta[lo_begin, lo_begin + lo_n) to lo_dst (nullable) - the pre-split operand of the first
 * tower layer for the NEXT step, which then needs no split launch. */
int ctr_adam_dense_ex(float* theta, float* m, float* v, float* g, int64_t n, float lr_t, float beta1,
                      float beta2, float eps, int zero_g, float* state_dev, int advance_parties,
                      float* lo_dst, int64_t lo_begin, int64_t lo_n, ctr_stream_t stream);
/* Lazy variant: exactly one update per distinct row in rows[n] (claim[R] int32
 * scratch, tag must differ from the previous call's), then zeroes the row of g.  Negative
 * row ids are skipped.  theta1/m1/v1/g1 (nullable): a per-row scalar parameter indexed by the
 * same rows (the first-order weights w1) updated under the same claim.
 *
 * Row strides (ctr_embed_fwd / ctr_embed_bwd / ctr_adam_rows; 0 = planar defaults D, 1, 1): the
 * distance in floats between consecutive rows of table / m / v / g (row_stride), of the
 * first-order arrays (w1_stride) and of claim (claim_stride).  With the ROW-RECORD layout
 *   record[r] = { theta[D] | m[D] | v[D] | g[D] | theta1 m1 v1 g1 | claim, cnt, c, pad }   (4D+8 floats;
 *   cnt / c: lookup count and sum of dy2 of the fused scatter + optimiser pass, ctr_embed_bwd_adam)
 * all pointers address one array with one stride, so everything the optimiser touches for a row
 * sits in one DRAM page (1 activate per row instead of 9: random row access is bounded by the
 * HBM activate rate long before its bandwidth), and the lookup's first-order weight shares the
 * page of the row it just read. */
int ctr_adam_rows(const int32_t* rows, int64_t n, int D, float* theta, float* m, float* v,
                  float* g, float* theta1, float* m1, float* v1, float* g1, int32_t* claim,
                  int32_t tag, float lr_t, float beta1, float beta2, float eps,
                  const float* state_dev, int64_t row_stride, int64_t w1_stride,
                  int64_t claim_stride, ctr_stream_t stream);
int ctr_adam_rows_ex(const int32_t* rows, int64_t n, int D, float* theta, float* m, float* v,
                     float* g, float* theta1, float* m1, float* v1, float* g1, int32_t* claim,
                     int32_t tag, float lr_t, float beta1, float beta2, float eps, float* state_dev,
                     int64_t row_stride, int64_t w1_stride, int64_t claim_stride,
                     int advance_parties, ctr_stream_t stream);
/* The same update for the id matrix rows[B, F] of one batch (what ctr_embed_fwd consumed), one
 * wave deep: a warp takes 32 consecutive samples of ONE field, so the lookups of a hot row meet
 * in one warp and are de-duplicated in registers (__match_any_sync) before the claim exchange,
 * and only the records of the claim winners are fetched.  Same arguments and results as
 * ctr_adam_rows_ex(rows, B*F, ...); D in {8, 16, 32}. */
int ctr_adam_rows_bf(const int32_t* rows, int B, int F, int D, float* theta, float* m, float* v,
                     float* g, float* theta1, float* m1, float* v1, float* g1, int32_t* claim,
                     int32_t tag, float lr_t, float beta1, float beta2, float eps,
                     float* state_dev, int64_t row_stride, int64_t w1_stride,
                     int64_t claim_stride, int advance_parties, ctr_stream_t stream);

/* Scatter-add and row optimiser in ONE pass (row-record layout only; rec = the record array,
 * row_stride >= 4D+8).  The unfused pair ctr_embed_bwd + ctr_adam_rows visits every touched record
 * twice; here the lookup that completes a row's gradient applies the update while the record is
 * still in L2:
 *   ctr_count_rows      cnt[row] += 1 for every entry of rows[n] (the record's `cnt` word; zero
 *                       before and after a step).  Launch it before the backward, e.g. on a side
 *                       stream beside the tower's backward GEMMs.
 *   ctr_embed_bwd_adam  per slot: g[row] += dE[b,f,:] + dy2[b]*S[b,:], c[row] += dy2[b],
 *                       g1[row] += dy1[b]; then cnt[row] -= multiplicity; the slot that takes cnt
 *                       to zero reads the record back, forms the gradient g - c*theta (which equals
 *                       sum dE + dy2*(S - E), fm/fm.py:123-129, without re-reading E), applies
 *                       TF-Adam to theta (and theta1), and clears g / c / g1.  Duplicates inside a
 *                       warp instruction are summed in registers (__match_any_sync), fields with
 *                       <= 32 rows per sample chunk in shared memory.  lr_t etc. as ctr_adam_rows. */
int ctr_count_rows(const int32_t* rows, int64_t n, int D, float* rec, int64_t row_stride,
                   ctr_stream_t stream);
int ctr_embed_bwd_adam(const int32_t* rows, const float* dE, const float* S, const float* dy2,
                       const float* dy1, uint64_t w1_fields, const int64_t* row_offsets_host, int B,
                       int F, int D, float* rec, int64_t row_stride, float lr_t, float beta1,
                       float beta2, float eps, const float* state_dev, ctr_stream_t stream);

/* -------------------------------------------------------- DIN activation unit
 * din/din.py:103-125 `_attention`: for each sample b and position p with hist[b,p] > 0
 *   h = table[hist[b,p]]; a = [h, q_b, h*q_b, h-q_b]; w = W3.relu(W2.relu(W1.a+b1)+b2)+b3
 *   out[b,:] = sum_p w * h        (no softmax; padding id 0 is masked out)
 * W1 [4E,H1], W2 [H1,H2], W3 [H2] row-major as tf.layers.dense kernels.  E in {8,16,32};
 * H1 <= 128, H2 <= 64.  att_w (nullable) [B,P] receives the raw position weights. */
/* ctr_din_opts (nullable everywhere = no dropout, no range check):
 *  - dropout after each of the two hidden attention layers (din/din.py:118, default rate 0.5 at :15):
 *    inverted dropout, keep iff a 16-bit uniform >= p_drop * 65536, uniforms from Philox4x32-10 with
 *    counter (b*P + pos, column / 8, 0x100 + 2*unit + layer, step) and key (seed, 0xD1A7); step =
 *    state[0] (the device Adam schedule, so that a captured CUDA graph draws a fresh mask on every
 *    replay); the backward regenerates the same bits.  ctr_din_dropout_mask writes the keep scales
 *    (0 or 1/(1-p)) of one layer for n_rows positions, [n_rows, H] (H = 80 for layer 0, 40 for
 *    layer 1) - what the parity tests inject into the oracle.
 *  - table_rows > 0: history ids outside [0, table_rows) read as padding instead of out of bounds
 *    and set bit 1 of *status (nullable device int) - tf.gather raises on CPU and zero-fills on GPU. */
typedef struct {
  const float* state;
  float p_drop;
  uint32_t seed;
  uint32_t unit;        /* which attention unit: 0 = item history, 1 = category history */
  int32_t table_rows;
  int32_t* status;
} ctr_din_opts;
int ctr_din_att_fwd(const float* table, const int32_t* hist, const float* query, int B, int P,
                    int E, const float* W1, const float* b1, int H1, const float* W2,
                    const float* b2, int H2, const float* W3, const float* b3, float* out,
                    float* att_w, const ctr_din_opts* opts, ctr_stream_t stream);
int ctr_din_dropout_mask(const ctr_din_opts* opts, int layer, int64_t n_rows, int H, float* out,
                         ctr_stream_t stream);
/* Backward of the above: dtable rows += (RED scatter), dquery[B,E] written, weight and bias
 * gradients (dW1..dW3, db1..db3) accumulated (+=).  workspace: ctr_din_workspace_bytes(B,P,E)
 * bytes of scratch (per-position h1/dh1/dh2/h rows feeding the tall-skinny dW reductions). */
int64_t ctr_din_workspace_bytes(int B, int P, int E);
int ctr_din_att_bwd(const float* table, const int32_t* hist, const float* query, int B, int P,
                    int E, const float* W1, const float* b1, int H1, const float* W2,
                    const float* b2, int H2, const float* W3, const float* b3, const float* dout,
                    float* dtable, float* dquery, float* dW1, float* db1, float* dW2, float* db2,
                    float* dW3, float* db3, void* workspace, int64_t workspace_bytes,
                    const ctr_din_opts* opts, ctr_stream_t stream);

/* --------------------------------------------------------------- xDeepFM CIN
 * One CIN layer (xdeepfm/xdeepfm.py:145-169):
 *   out[b,d,h] = relu( sum_{i<m, j<Hp} X0[b,d,i] * Xp[b,d,j] * W[i*Hp+j, h] + bias[h] )
 * Internal layouts are "d-major rows": X0t [B*D, m_pad], Xp [B*D, Hp_pad], out [B*D, H]
 * (row = b*D+d, feature contiguous; *_pad = leading dimension).
 * prec: 0 = fp32 CUDA cores (exact-parity mode), 1 = TF32 tcgen05 tensor cores,
 *       2 = 3xTF32 split on tcgen05 (fp32-grade accuracy). */
#define CTR_CIN_FP32 0
#define CTR_CIN_TF32 1
#define CTR_CIN_TF32X3 2
int64_t ctr_cin_workspace_bytes(int B, int D, int m, int Hp, int H, int prec);
int ctr_cin_layer_fwd(const float* X0t, int ld0, const float* Xp, int ldp, const float* W,
                      const float* bias, int B, int D, int m, int Hp, int H, float* out, int prec,
                      void* workspace, int64_t workspace_bytes, ctr_stream_t stream);
/* dpre[B*D,H] = d(loss)/d(pre-activation) (ReLU mask already applied by the caller).
 * dX0t, dXp accumulated (+=) - X0 feeds every layer; dW [m*Hp,H], dbias [H] accumulated. */
int ctr_cin_layer_bwd(const float* X0t, int ld0, const float* Xp, int ldp, const float* W,
                      const float* dpre, int B, int D, int m, int Hp, int H, float* dX0t,
                      float* dXp, float* dW, float* dbias, int prec, void* workspace,
                      int64_t workspace_bytes, ctr_stream_t stream);
/* The sum-pool over the embedding axis (xdeepfm/xdeepfm.py:180-181) on d-major rows and its
 * backward folded with the ReLU of the layer it feeds back into:
 *   pooled[b, h] = sum_d out[b*D+d, h]                                       (ld_pooled >= H)
 *   dpre[b*D+d, h] = (dpool[b, h] + dacc[b*D+d, h]) * 1[out[b*D+d, h] > 0]   (dacc nullable: what the
 *                    next layer's backward accumulated for this layer's output)
 * H, ld_pooled, ld_dpool multiples of 4. */
int ctr_cin_pool(const float* out, int B, int D, int H, float* pooled, int ld_pooled,
                 ctr_stream_t stream);
int ctr_cin_dpre(const float* dpool, int ld_dpool, const float* dacc, const float* out, int B, int D,
                 int H, float* dpre, ctr_stream_t stream);
/* [B, F, D] (E layout) <-> [B*D, ld] d-major rows (zero padded to ld). */
int ctr_transpose_fd(const float* E, int B, int F, int D, float* Xt, int ld, ctr_stream_t stream);
int ctr_transpose_df_add(const float* dXt, int ld, int B, int F, int D, float* dE,
                         ctr_stream_t stream);

/* ------------------------------------------------------------- dense tower + loss head
 * The reference's tower is `dense(relu) -> batch_normalization -> dropout` per layer plus a
 * final dense(1, relu) (deepfm/deepfm.py:100-108, xdeepfm/xdeepfm.py:184-192,
 * dcn/dcn.py:144-149), then dense(concat[...], 1), sigmoid and the mean
 * sigmoid-cross-entropy (deepfm/deepfm.py:110-129).  fp32 results from 3xTF32 tensor-core
 * GEMMs (tcgen05 for wide first layers, mma.sync for the rest); BN and dropout are
 * never materialised: they are a prologue of the consuming GEMM, recomputed in the backward
 * (dropout masks are counter based: Philox4x32-10 over (seed, layer, step, row, col)).
 *
 * ctr_bn_drop: how to turn a stored post-ReLU activation A[r,k] into the next layer's input:
 *   x' = ((A - mu) * rstd * gamma + beta) * keep/(1-p).  Train: mu / biased var from the column
 *   sums `sums` = [2][K] (sum A, sum A^2) over the B rows; eval: moving `mean` / `var`. */
typedef struct {
  const float* sums;   /* train: device [2][K]; NULL in eval                         */
  const float* mean;   /* eval: device [K] moving mean / variance (sums == NULL)     */
  const float* var;
  const float* gamma;  /* device [K]                                                 */
  const float* beta;
  const float* state;  /* device Adam schedule (see "optimiser"): [0] = t selects the dropout stream; NULL -> 0 */
  float eps;           /* 1e-3 (tf.layers.batch_normalization default)               */
  float p_drop;        /* dropout rate, 0 disables                                   */
  uint32_t seed;
  uint32_t layer;
  int32_t enabled;     /* 0: identity prologue                                       */
  int32_t pad_;
} ctr_bn_drop;
/* ctr_grad_src: gradient arriving at a layer's stored post-ReLU output a[r,n]:
 *   kind 0: g = G[r*ldg+n];  kind 1: g = BN-backward of the stored dn = G through the BN that
 *   follows a (needs the column sums dbeta = sum dn, dgamma = sum dn*xhat in train mode).
 *   The kernels then use dpre = g * 1[a > 0].
 *   kind 2: G already is dpre (written by ctr_tower_dpre); `a` is not read. */
typedef struct {
  const float* G;
  const float* a;
  const float* sums;
  const float* mean;
  const float* var;
  const float* gamma;
  const float* dbeta;
  const float* dgamma;
  int32_t ldg;
  int32_t lda;
  float eps;
  int32_t kind;
  int32_t train;
  int32_t pad_;
} ctr_grad_src;
/* out[B,N] = act(P(X)[B,K] . W[K,N] + bias); stats (nullable) [2][N] += column sums of out, out^2. */
int ctr_tower_layer_fwd(const float* X, int ldx, int K, const ctr_bn_drop* pro, const float* W,
                        const float* bias, int N, float* out, int ldo, float* stats, int relu,
                        int B, ctr_stream_t stream);
/* out[B,K] = P(A): the last BN+dropout of a tower that does not end in a dense layer (DCN). */
int ctr_bn_drop_apply(const float* A, int K, const ctr_bn_drop* pro, float* out, int B,
                      ctr_stream_t stream);
/* Backward of ctr_bn_drop_apply (dcn/dcn.py:146-149, the tower that ends in BN + dropout):
 * dn[B,K] = dout * keep; dbeta[K] += colsum(dn); dgamma[K] += colsum(dn * xhat(A)) (both nullable).
 * The layer's dpre then follows from a kind-1 gradient source over dn. */
int ctr_bn_drop_apply_bwd(const float* dout, int ldd, const float* A, int K, const ctr_bn_drop* pro,
                          float* dn, float* dbeta, float* dgamma, int B, ctr_stream_t stream);
/* DCN head + loss in one launch (dcn/dcn.py:151-153,166-169): logit = [h | xl] . w + hb with
 * h [B,H] the tower output and xl [B,W] the cross output (H % 4 == W % 4 == 0, H + W <= 1536);
 * logits / prob (nullable) written, *loss += mean BCE.  With dh != NULL the gradients come out of
 * the same launch: dh [B,H], dxl [B,W] written; dw [H+W], dhb accumulated; all scaled by
 * grad_scale * B (grad_scale = 1/(B*world)). */
int ctr_dcn_head(const float* h, int H, const float* xl, int W, const float* w, const float* hb,
                 const float* labels, int B, float* logits, float* prob, float* loss, float* dh,
                 float* dxl, float* dw, float* dhb, float grad_scale, ctr_stream_t stream);
/* dn_out[B,K] = (dpre[B,N] . W[K,N]^T) * keep  (keep from `pro`, the prologue that produced this
 * layer's input from Aprev); dbeta_prev/dgamma_prev [K] += column sums of dn, dn*xhat(Aprev).
 * pro disabled (first layer): dn_out = dpre . W^T, nothing else. */
int ctr_tower_layer_bwd_data(const ctr_grad_src* gs, int N, const float* W, int K,
                             const ctr_bn_drop* pro, const float* Aprev, float* dn_out, int ldn,
                             float* dbeta_prev, float* dgamma_prev, int B, ctr_stream_t stream);
/* dpre[B,N] = g * 1[a > 0] for a kind 0/1 gradient source, written out once so that both
 * backward GEMMs of the layer can read it as a plain tensor (kind 2);
 * db[N] (nullable) += column sums of dpre. */
int ctr_tower_dpre(const ctr_grad_src* gs, int N, float* dpre, int ldd, float* db, int B,
                   ctr_stream_t stream);
/* dW[K,N] += P(X)^T . dpre;  db[N] (nullable) += column sums of dpre. */
int ctr_tower_layer_bwd_weights(const float* X, int ldx, int K, const ctr_bn_drop* pro,
                                const ctr_grad_src* gs, int N, float* dW, float* db, int B,
                                ctr_stream_t stream);
/* Loss head (deepfm/deepfm.py:110-129): logit = sum_{c<C} hw[c]*act_c(z_c) + hb with
 * act_0 = relu(. + b1) when relu0 (the ReLU'd first-order term), identity otherwise; C <= 4.
 * logits/prob (nullable) written; *loss += mean BCE.  With dz != NULL the gradients are produced in
 * the same launch: dz[c][b] written, dhw[C], dhb, db1 accumulated, all times grad_scale*B
 * (grad_scale = 1/(B*world)).  z / dz are HOST arrays of C device pointers. */
int ctr_loss_head(const float* const* z, float* const* dz, int C, int relu0, const float* hw,
                  const float* hb, const float* b1, const float* labels, int B, float* logits,
                  float* prob, float* loss, float* dhw, float* dhb, float* db1, float grad_scale,
                  ctr_stream_t stream);

/* 3xTF32 with pre-split operands.  The tower's wide first layer (624 -> H) is the one GEMM of a
 * CTR step that belongs on tcgen05; its fp32-grade accuracy comes from
 *   a.b ~ a_lo.b_hi + a_hi.b_lo + a_hi.b_hi,   hi = the top 19 bits of the fp32 word (what
 *   tcgen05.mma.kind::tf32 reads), lo = tf32(x - hi).
 * Computing lo inside the GEMM costs CUDA-core work per operand tile and per CTA; here the
 * producer of each operand writes lo next to it (ctr_embed_fwd: E_lo; ctr_tower_mid: dpre0_lo;
 * ctr_split_lo for the weights) and the GEMM only streams tiles by TMA into tcgen05.
 * ctr_split_lo: lo[i] = tf32(x[i] - trunc_tf32(x[i])), n floats. */
int ctr_split_lo(const float* x, float* lo, int64_t n, ctr_stream_t stream);
/* The three GEMMs of a dense(relu) layer out = relu(X . W + b), X [B,K], W [K,N] (deepfm.py:101):
 *   kind 0  forward   out[B,N] = act(X . W + bias); stats (nullable) [2][N] += colsums(out, out^2)
 *                     A = X, B = W
 *   kind 1  data      out[B,K] = dpre[B,N] . W^T                     A = dpre, B = W
 *   kind 2  weights   out[K,N] += X^T . dpre   (split over the rows)  A = X,    B = dpre
 *   kind 3  forward, split-K: out[B,N] += X . W only (out zero on entry; bias, ReLU and the
 *                     column statistics are then applied by ctr_tower_mid, field pre0)
 * A_lo / B_lo: the lo halves (same shapes and pitches).  Needs B >= 256, K % 4 == N % 4 == 0. */
int ctr_tower_gemm_presplit(int kind, const float* A, const float* A_lo, const float* Bm,
                            const float* B_lo, int B, int K, int N, float* out, const float* bias,
                            float* stats, int relu, ctr_stream_t stream);

/* ctr_tower_mid: everything between the first layer's GEMM and the first layer's backward
 * GEMMs in ONE cooperative launch (deepfm/deepfm.py:100-129, xdeepfm/xdeepfm.py:184-212):
 * hidden layers 1..L-1 forward (BN of the previous layer's stored output + dropout as the GEMM
 * prologue), the final dense(1, relu), the logit / sigmoid / mean-BCE head, and - in training -
 * the whole backward down to dpre_0 = d loss / d (pre-activation of layer 0).  The whole-batch
 * BN reductions between the pieces are grid barriers instead of kernel boundaries.
 *   given:   act[0] [B,H[0]] (post-ReLU output of layer 0) and, training, its column sums
 *            stats[0] = [2][H[0]]; the C-1 external head columns z[c] [B]; labels [B]
 *   written: act[l], l >= 1; y_out / logits / prob [B] (nullable); *loss += mean BCE;
 *   training (all accumulated with +=, buffers zero on entry except where noted):
 *            stats[l] l >= 1 (zero on entry), dz[c] [B] (plain store), dhw[C], dhb, db1,
 *            dw_out[H[L-1]], db_out, dbeta[l] / dgamma[l] / dbias[l] for every l,
 *            dpre[l] [B,H[l]] (plain store; feeds ctr_tower_layer_bwd_weights / _bwd_data as a
 *            kind-2 gradient source), dn[l] [B,H[l]] scratch.
 * Weight gradients dW_l are NOT computed here (they are off the critical path: the caller
 * runs ctr_tower_layer_bwd_weights on a side stream).  Hidden widths: multiples of 4, <= 128.
 * `barrier`: CTR_TOWER_MID_BARRIER_WORDS device words, zero-initialised once, reusable across
 * launches ({arrival count, generation}; the rest is reserved). */
#define CTR_TOWER_MID_BARRIER_WORDS 320
#define CTR_TOWER_MID_MAX_LAYERS 4
typedef struct {
  int32_t L, C, relu0, training;
  int32_t H[CTR_TOWER_MID_MAX_LAYERS];
  const float* W[CTR_TOWER_MID_MAX_LAYERS];      /* W[l]: [H[l-1], H[l]], l >= 1 (W[0] unused) */
  const float* b[CTR_TOWER_MID_MAX_LAYERS];      /* b[l]: [H[l]], l >= 1 (b[0] only with pre0) */
  const float* gamma[CTR_TOWER_MID_MAX_LAYERS];
  const float* beta[CTR_TOWER_MID_MAX_LAYERS];
  const float* mean[CTR_TOWER_MID_MAX_LAYERS];   /* eval only */
  const float* var[CTR_TOWER_MID_MAX_LAYERS];
  float* act[CTR_TOWER_MID_MAX_LAYERS];
  float* stats[CTR_TOWER_MID_MAX_LAYERS];
  const float* w_out;                            /* [H[L-1]] */
  const float* b_out;
  const float* state;                            /* device Adam schedule: [0] = t selects the dropout stream */
  float eps, p_drop;
  uint32_t seed;
  float grad_scale;                              /* 1 / (B * world) */
  const float* z[3];
  const float* hw;                               /* [C], the tower's column is the last one */
  const float* hb;
  const float* b1;
  const float* labels;
  float* y_out;
  float* logits;
  float* prob;
  float* loss;
  float* dz[3];
  float* dhw;
  float* dhb;
  float* db1;
  float* dw_out;
  float* db_out;
  float* dgamma[CTR_TOWER_MID_MAX_LAYERS];
  float* dbeta[CTR_TOWER_MID_MAX_LAYERS];
  float* dbias[CTR_TOWER_MID_MAX_LAYERS];
  float* dn[CTR_TOWER_MID_MAX_LAYERS];
  float* dpre[CTR_TOWER_MID_MAX_LAYERS];
  float* dpre0_lo;                               /* nullable: lo half of dpre[0] (ctr_split_lo rule) */
  const float* pre0;                             /* nullable: X . W_0 without bias (split-K GEMM, kind 3):
                                                    act[0] = relu(pre0 + b[0]) and stats[0] are then
                                                    produced here (one more phase and barrier) */
  uint32_t* barrier;
  unsigned long long* timing;                    /* nullable: 8 words, %globaltimer (ns) of block 0 at
                                                    the phase boundaries (profiling aid) */
  const float* stats0_part;                      /* nullable: [n_stats0_part][2][H[0]] partial column sums
                                                    of act[0] / act[0]^2 (ctr_embed_tower_fwd); they
                                                    replace stats[0], which is then not read */
  int32_t n_stats0_part;
  int32_t pad_;
} ctr_tower_mid_args;
int ctr_tower_mid(const ctr_tower_mid_args* args, int B, ctr_stream_t stream);

/* ------------------------------------------------------- row-sharded table (multi-GPU)
 * The reference only replicates (tf.distribute.MirroredStrategy, fm/fm.py:184-194); row
 * sharding is the north-star extension for tables larger than one GPU's HBM.  owner(row) =
 * row % G, local index = row / G.  These are the device halves of the exchange; the
 * all-to-all itself is NCCL (torch.distributed) in recsys_b200/sharded.py.
 *
 * ctr_shard_bucket: requester side.  For lookup i: slot[i] = owner*capacity + pos and
 *   send_local[slot[i]] = rows[i] / G; unused slab entries are -1; counts[owner] = number of
 *   lookups for that owner (> capacity means overflow: those lookups got slot -1). */
int ctr_shard_bucket(const int32_t* rows, int64_t n, int G, int capacity, int32_t* send_local,
                     int32_t* slot, int32_t* counts, ctr_stream_t stream);
/* Owner side: out[i,:] = table[ids[i],:] (zeros for ids[i] < 0); out_w1[i] = w1[ids[i]].
 * Strides in floats, 0 = planar (D, 1): with out_stride = out_w1_stride = D+4 and out_w1 = out + D the
 * row and its first-order weight travel in ONE exchange slab (one all-to-all instead of two); the
 * table side takes the row-record stride. */
int ctr_gather_rows(const float* table, const float* w1, const int32_t* ids, int64_t n, int D,
                    float* out, float* out_w1, int64_t table_stride, int64_t w1_stride,
                    int64_t out_stride, int64_t out_w1_stride, ctr_stream_t stream);
/* Owner side: dtable[ids[i],:] += g[i,:]; dw1[ids[i]] += gw1[i]; ids < 0 skipped. */
int ctr_scatter_add_rows(const int32_t* ids, const float* g, const float* gw1, int64_t n, int D,
                         float* dtable, float* dw1, int64_t g_stride, int64_t gw1_stride,
                         int64_t dtable_stride, int64_t dw1_stride, ctr_stream_t stream);

/* ------------------------------------- row-sharded table, device-initiated exchange (NVLink P2P)
 * The NCCL-free version of the exchange above (one process per GPU on one NVSwitch box): every
 * rank owns a same-shaped ARENA in its HBM and maps every peer's arena (cudaIpc); the kernels of a
 * step store straight into the consumer's arena and raise a sequence-numbered flag there.
 *
 *   arena := { int step; int err; ... | req_flag[2][G] req_cnt[2][G] resp_flag[G] grad_flag[G]
 *              dense_flag[G] counts[G] sent[G] done[1+G] | req_ids[2][G][capacity] (int32, owner side)
 *              | inv[G][capacity] (int32, requester side: lookup index of every slab position)
 *              | resp[G][capacity][P] (fp32, requester side) | grad[G][capacity][P] (owner side)
 *              | dense[G][n_dense] }        P = record_floats = D + 4: row | w1 | pad, and on the
 *                                           way back gradient | dy1 dy2 pad
 *   capacity >= the lookups of one rank per step (worst case: all for one owner) - no overflow.
 * The layout (byte offsets below) is chosen by the caller and must be identical on every rank.
 *
 * ctr_p2p_alloc / _open / _close / _free: the ONLY entry points of this library that allocate:
 * cudaMalloc'ed, zero-filled arena + its 64-byte cudaIpcMemHandle_t; map a peer's handle.
 * Per step (all async on `stream`, graph-capturable; K1 first - it opens the step):
 *   K1 ctr_p2p_bucket_send   rows[n] (global rows) -> owner = row % G; owner-local ids stored into
 *                            the owner's req_ids[step parity][me][pos]; slot[i] = owner*capacity+pos
 *   K2 ctr_p2p_gather_reply  owner: per requester, wait for its ids, gather row | w1 from the local
 *                            row records into the requester's resp[me][pos] (count_lookups is
 *                            reserved and ignored)
 *   K3 ctr_embed_fwd_p2p     the fused lookup + interaction kernel over the reply slab (slots as
 *                            row ids); waits for every owner's reply flag inside the kernel
 *   K4 ctr_p2p_grad_send     in slab order (inv[]): dE + dy2*S | dy1 dy2 stored into the owner's
 *                            grad[me][pos] as contiguous runs (`slot` is unused, kept for symmetry)
 *   K5 ctr_p2p_scatter_adam  owner, two launches: (a) per requester, wait for its gradients and
 *                            RED them into the records' accumulators (g, g1, c = sum dy2; duplicates
 *                            of a warp instruction summed in registers); (b) one TF-Adam update per
 *                            distinct row (claim word tagged with the step number), gradient
 *                            g - c*theta, accumulators cleared
 *   K6 ctr_p2p_dense_push + ctr_p2p_adam_dense   replicated dense weights: gradients stored into
 *                            every peer's dense[me]; Adam over their sum in rank order (bitwise
 *                            identical on every rank); zeroes g_local; advance_state as ctr_adam_dense
 * Every wait is bounded by spin_limit_ms (0 = 10 s): a peer that never arrives sets arena.err
 * (ctr_p2p_status) instead of hanging the GPU. */
#define CTR_P2P_MAX_RANKS 8
typedef struct {
  void* peer[CTR_P2P_MAX_RANKS]; /* arena base of every rank as mapped in THIS process */
  int32_t me, G, capacity, record_floats;
  int64_t off_req_flag, off_req_cnt, off_resp_flag, off_grad_flag, off_dense_flag;
  int64_t off_req_ids, off_resp, off_grad, off_dense, off_counts, off_done;
  int64_t off_inv, off_sent; /* requester side: inv[G][capacity] int32 lookup index per slab position; sent[G] */
  int64_t n_dense;
  int32_t spin_limit_ms, pad_;
} ctr_p2p_ctx;
int ctr_p2p_alloc(int64_t bytes, void** ptr, void* ipc_handle_out);
int ctr_p2p_open(const void* ipc_handle, void** ptr);
int ctr_p2p_close(void* ptr);
int ctr_p2p_free(void* ptr);
int ctr_p2p_bucket_send(const int32_t* rows, int64_t n, const ctr_p2p_ctx* ctx, int32_t* slot,
                        ctr_stream_t stream);
int ctr_p2p_gather_reply(float* rec, int64_t row_stride, int D, int with_w1, int count_lookups,
                         const ctr_p2p_ctx* ctx, ctr_stream_t stream);
/* K3: ctr_embed_fwd over this rank's reply slab (rows = slot [B,F], row | w1 records of
 * record_floats), preceded IN THE SAME KERNEL by the wait for every owner's reply flag. */
int ctr_embed_fwd_p2p(const int32_t* slot, int B, int F, int D, uint64_t w1_fields, int with_w1,
                      float* E, float* S, float* y1, float* y2, const float* cross_w,
                      const float* cross_b, int cross_layers, float* xl, float* E_lo,
                      const ctr_p2p_ctx* ctx, ctr_stream_t stream);
/* what: 0 = every owner's reply, 1 = every requester's gradients, 2 = every rank's dense gradients */
int ctr_p2p_wait(const ctr_p2p_ctx* ctx, int what, ctr_stream_t stream);
int ctr_p2p_grad_send(const int32_t* slot, const float* dE, const float* S, const float* dy2,
                      const float* dy1, uint64_t w1_fields, int B, int F, int D,
                      const ctr_p2p_ctx* ctx, ctr_stream_t stream);
int ctr_p2p_scatter_adam(float* rec, int64_t row_stride, int D, int with_w1, int has_c, float lr_t,
                         float beta1, float beta2, float eps, const float* state_dev,
                         const ctr_p2p_ctx* ctx, ctr_stream_t stream);
int ctr_p2p_dense_push(const float* grad, int64_t n, const ctr_p2p_ctx* ctx, ctr_stream_t stream);
int ctr_p2p_adam_dense(float* theta, float* m, float* v, float* g_local, int64_t n, float lr_t,
                       float beta1, float beta2, float eps, float* state_dev, int advance_state,
                       const ctr_p2p_ctx* ctx, ctr_stream_t stream);
/* Synchronous: the arena's step counter and error word (bit 0: a bounded wait timed out). */
int ctr_p2p_status(const ctr_p2p_ctx* ctx, int32_t* step_out, int32_t* err_out);

#ifdef __cplusplus
}
#endif
#endif /* CTR_B200_H_ */
