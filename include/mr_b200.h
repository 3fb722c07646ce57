/* mr_b200.h — C ABI of libmrgpu.so, the B200-native replacement for the /rank
 * inference hot path of metarank/metarank (feature-vector assembly + LambdaMART
 * GBDT scoring).  Plain C: pointers and sizes only, no C++/torch types.
 *
 * Metarank has no native plugin ABI of its own (it is 100 % JVM); each entry point
 * below replaces one JVM seam of the hot path and cites it.  Paths are relative to
 * the reference tree, S/ = src/main/scala/ai/metarank/.
 *
 * Conventions
 *   - every function returns an mr_status; nothing throws or aborts across the ABI;
 *     the message for the last failure on the calling thread is mr_last_error().
 *   - the caller owns every input/output buffer; the library copies what it keeps.
 *   - doubles in, doubles out; NaN = missing (S/model/MValue.scala:37-39,56-61);
 *     rows/cols are int32 (JVM array limit, S/ml/rank/LambdaMARTRanker.scala:304-306).
 *   - handles are safe for concurrent calls from any thread (http4s fibers call
 *     Ranker.rerank concurrently, S/main/command/Serve.scala:116-126).
 *   - there is NO CPU fallback: without a usable sm_100 device mr_init fails with
 *     MR_ERR_NO_DEVICE and nothing else can be called.
 */
#ifndef MR_B200_H
#define MR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MR_API __attribute__((visibility("default")))
#else
#define MR_API
#endif

typedef int32_t mr_status;
enum {
  MR_OK = 0,
  MR_ERR_INVALID_ARG = 1,      /* null pointer, negative size, wrong dimension */
  MR_ERR_PARSE = 2,            /* model / schema / state bytes malformed */
  MR_ERR_CUDA = 3,             /* a CUDA runtime call or kernel failed */
  MR_ERR_CLOSED = 4,           /* handle used after close() */
  MR_ERR_UNSUPPORTED = 5,      /* valid input the library does not implement */
  MR_ERR_FEATURE_MISMATCH = 6, /* model blob's feature list != config's (LambdaMARTRanker.scala:208-213) */
  MR_ERR_ARITHMETIC = 7,       /* java.lang.ArithmeticException analogue: Long / 0 in normalized rate
                                  (S/feature/RateFeature.scala:346-348) */
  MR_ERR_NO_DEVICE = 8,        /* no sm_100 GPU / driver */
  MR_ERR_NOT_FOUND = 9         /* unknown feature / table / model name */
};

typedef struct mr_ctx mr_ctx;       /* one per process per GPU */
typedef struct mr_model mr_model;   /* a loaded booster == ltrlib Booster[_] */
typedef struct mr_schema mr_schema; /* FeatureMapping for one model */
typedef struct mr_state mr_state;   /* device-resident Persistence.values */

/* ------------------------------------------------------------------ lifecycle */

/* Binds the calling process to CUDA device `device` (one process per GPU; LOCAL_RANK
 * under torchrun).  Fails with MR_ERR_NO_DEVICE when no compute-capability-10.x GPU
 * is present.  Replaces nothing in the reference (the JVM loads lightgbm4j /
 * xgboost4j natives lazily); it is where Serve.api would create the backend
 * (S/main/command/Serve.scala:72-128). */
MR_API mr_status mr_init(int32_t device, mr_ctx **out);
MR_API mr_status mr_shutdown(mr_ctx *ctx);
/* Thread-local, valid until the next failing call on the same thread. */
MR_API const char *mr_last_error(void);
/* "libmrgpu <version> sm_100a" */
MR_API const char *mr_version(void);
/* Number of this library's kernels launched so far by this process (bench.py's
 * gpu_launches claim). */
MR_API int64_t mr_kernel_launches(void);
/* Per-kernel device time for bench.py's roofline block.  Between begin and end every kernel this library
 * launches (any context of the process) is bracketed by two CUDA events on its stream; end waits for the
 * device and writes a JSON array [{"kernel": name, "launches": n, "ms": total device time}] (NUL-terminated;
 * *len = its length without the NUL; pass json = NULL to size the buffer — that call also closes the
 * profile, so size generously instead).  Not for the serving path: the events serialise nothing but cost a
 * few microseconds per launch. */
MR_API mr_status mr_profile_begin(void);
MR_API mr_status mr_profile_end(char *json, size_t cap, size_t *len);

/* ------------------------------------------------------------------ scorer (Booster) */

enum { MR_BOOSTER_LIGHTGBM = 0, MR_BOOSTER_XGBOOST = 1 };

/* LightGBMBooster(bytes) / XGBoostBooster(bytes)
 * (S/ml/rank/LambdaMARTRanker.scala:228-232).  kind 0: LightGBM model text as
 * written by booster.save(); kind 1: XGBoost model bytes (JSON or UBJSON).
 * n_features > 0 additionally checks the model's feature count (DatasetDescriptor.dim);
 * pass 0 to skip the check. */
MR_API mr_status mr_model_load(mr_ctx *ctx, int32_t kind, const uint8_t *blob, size_t len, int32_t n_features,
                        mr_model **out);

/* LambdaMARTPredictor.load (S/ml/rank/LambdaMARTRanker.scala:192-236): parses
 * Metarank's own model framing (byte version 2|3, int nFeatures, writeUTF names,
 * byte boosterType, int size, booster bytes, v3 warm-up requests which are skipped).
 * feature_names/n_names = the config's `features:` list; a differing list fails with
 * MR_ERR_FEATURE_MISMATCH exactly as :208-213 does.  n_names < 0 skips the check. */
MR_API mr_status mr_model_load_metarank(mr_ctx *ctx, const uint8_t *blob, size_t len, const char *const *feature_names,
                                 int32_t n_names, mr_model **out);

/* Booster.predictMat(values: Array[Double], rows: Int, cols: Int): Array[Double]
 * — the one call LambdaMARTModel.predict makes (S/ml/rank/LambdaMARTRanker.scala:348).
 * values: row-major rows×cols doubles in HOST memory (ltrlib Query.values,
 * S/flow/ClickthroughQuery.scala:50-74); out_scores: rows doubles in host memory.
 * Host<->device copies are part of the call.  rows == 0 is a no-op. */
MR_API mr_status mr_model_predict_mat(mr_model *m, const double *values, int32_t rows, int32_t cols, double *out_scores);

/* Same computation on DEVICE buffers, enqueued on `cuda_stream` (a cudaStream_t; 0 =
 * the legacy default stream) without synchronising.  Used by the fused rank path and
 * by bench.py's HBM-resident `value`. */
MR_API mr_status mr_model_predict_mat_device(mr_model *m, const double *d_values, int32_t rows, int32_t cols,
                                      double *d_out_scores, void *cuda_stream);

/* The two halves of the binned scorer as separate device-side steps (what mr_model_predict_mat_device
 * runs back to back when the model can be binned): values -> exact u16 rank codes
 * ([group of 32 rows][tile column][lane] layout, mr_model_codes_bytes(rows) bytes; a tile has one column per
 * feature plus a second one for every feature whose splits send NaN both ways — opaque to the caller, only
 * valid for the model and options it was produced with), then the tree traversal on the codes.
 * mr_model_codes_bytes returns 0 when the model is scored by the f64/f32 kernel
 * instead (categorical XGBoost splits, zero-as-missing LightGBM nodes, > 4095 columns). */
MR_API size_t mr_model_codes_bytes(mr_model *m, int32_t rows);
MR_API mr_status mr_model_bin_device(mr_model *m, const double *d_values, int32_t rows, int32_t cols, void *d_codes,
                                     void *cuda_stream);
MR_API mr_status mr_model_score_codes_device(mr_model *m, const void *d_codes, int32_t rows, double *d_out_scores,
                                             void *cuda_stream);

/* Booster.save(): the original booster bytes (S/ml/rank/LambdaMARTRanker.scala:373).
 * The pointer stays valid until the model is freed. */
MR_API mr_status mr_model_save(mr_model *m, const uint8_t **blob, size_t *len);

/* Booster.weights(): per-column split counts (S/ml/rank/LambdaMARTRanker.scala:392).
 * out has n doubles, n >= the model's feature count. */
MR_API mr_status mr_model_weights(mr_model *m, double *out, int32_t n);

typedef struct mr_model_info {
  int32_t kind;          /* MR_BOOSTER_* */
  int32_t n_features;    /* columns the model reads */
  int32_t n_trees;
  int32_t max_leaves;    /* largest tree */
  int32_t n_chunks;      /* TMA-staged tree chunks (1 = model resident in shared memory) */
  int32_t has_categorical;
  int64_t n_internal_nodes;
  int64_t device_bytes;  /* packed model size in HBM */
} mr_model_info;
MR_API mr_status mr_model_get_info(mr_model *m, mr_model_info *out);

/* Host-only: parse + pack a booster blob without touching the GPU and report its shape.
 * Lets a JVM validate a model at config time (and lets CPU-only CI exercise the
 * parsers).  It does not score anything. */
MR_API mr_status mr_model_inspect(int32_t kind, const uint8_t *blob, size_t len, int32_t chunk_kb, mr_model_info *out);

/* Host-only consistency check of the throughput scorer's packing (4-byte entries, root table, in-entry bitsets): `samples`
 * random code vectors are walked through the parsed trees and through the packed bytes the way the kernel reads them.
 * max_tile: largest CTA tile to pack for (512 | 256 | 128; 0 = 512) — the library keeps one form per tile size.
 * *form: 0 = the model has no such packing (other scorers serve it), else 1 | 2 * (root table present) |
 * 4 * (small-categorical codes) | tile size << 8; *mismatches: (sample, tree) pairs whose leaf differs — 0 for a sound
 * packing.  Validates layout on CPU-only CI; it scores nothing and is not a fallback. */
MR_API mr_status mr_model_selfcheck(int32_t kind, const uint8_t *blob, size_t len, int32_t samples, int32_t max_tile, int32_t *form,
                                    int64_t *mismatches);

/* Mean evaluated path length (internal nodes visited per item per tree) of the last
 * `rows` scored by mr_model_count_path(): the d̄ of SURVEY.md §8(d)'s B_item. */
MR_API mr_status mr_model_count_path(mr_model *m, const double *values, int32_t rows, int32_t cols, double *mean_path);

/* What the lock-step scorers execute on a device-resident matrix (compact / slim scorers only): lane_levels = the sum of
 * the rows' path lengths (internal nodes visited), warp_levels = the sum over (warp of 32 consecutive rows, tree) of the
 * DEEPEST lane's path — the level steps a warp really issues — warp_trees = the number of (warp, tree) pairs.
 * lane_levels / (32 * warp_levels) is the fraction of lanes doing useful work; bench.py's shared-memory roofline is built
 * from these counts.  Synchronous. */
MR_API mr_status mr_model_walk_stats(mr_model *m, const double *d_values, int32_t rows, int32_t cols, double *lane_levels,
                                     double *warp_levels, double *warp_trees, void *cuda_stream);

/* Booster.close() / isClosed() (S/ml/rank/LambdaMARTRanker.scala:361-365).  close is
 * idempotent; in-flight predicts finish (CachedModelStore may dispose a model that is
 * still in use, S/fstore/cache/CachedModelStore.scala:39-42).  free releases the handle. */
MR_API mr_status mr_model_close(mr_model *m);
MR_API int32_t mr_model_is_closed(mr_model *m);
MR_API mr_status mr_model_free(mr_model *m);

/* Tuning knobs for experiments (bench.py / tests); defaults are chosen per model.  Not synchronised
 * against concurrent predicts on the same handle: set them before serving.
 * key: "threads" (items per CTA, 0 = auto), "chunk_kb", "ilp" (trees in flight per thread, exact kernel: 2 | 4),
 * "variant" (-1 = auto, 0 = exact f64/f32 kernel, 2 = generic binned kernel, 4 = compact binned kernel with
 * 8-byte nodes, 5 = slim kernel with 4-byte nodes where the model has that form; any other value is MR_ERR_INVALID_ARG), "latency_rows" (largest batch the tree-parallel low-latency path takes,
 * 0 = auto). */
MR_API mr_status mr_model_set_option(mr_model *m, const char *key, int32_t value);

/* ------------------------------------------------------------------ final ordering */

/* Ranker.rerank's `sortBy(-_.score)` (S/ml/Ranker.scala:52-67): stable, descending,
 * java.lang.Double.compare on the negated score (NaN last; 0.0 before -0.0).
 * Sorts `n_requests` independent requests laid out back to back: request r owns
 * scores[offsets[r] .. offsets[r+1]); order receives item indices relative to the
 * request start.  Host buffers. */
MR_API mr_status mr_rank_order(mr_ctx *ctx, const double *scores, const int32_t *offsets, int32_t n_requests,
                        int32_t *order);

/* ------------------------------------------------------------------ feature assembly */

/* All strings that identify things at run time (item / user / session ids, field values,
 * tags) cross the ABI as 64-bit hashes produced by mr_hash64 (the JVM hashes while it
 * decodes the request JSON).  0 is reserved for "absent" (None). */
MR_API uint64_t mr_hash64(const void *bytes, size_t len);

/* WordCountFeature.tokenCount (S/feature/WordCountFeature.scala:73-76):
 * "\\s+".split(s).length with java.util.regex semantics (a leading whitespace run yields
 * an empty first token, trailing empty tokens are dropped, "" -> 1).  Host helper for the
 * ranking-scoped word_count input. */
MR_API int32_t mr_token_count(const char *utf8, size_t len);

/* FeatureMapping for one model (S/FeatureMapping.scala:56-99).  `json` is
 *   {"features": [<Metarank FeatureSchema JSON, polymorphic on "type">, ...],
 *    "model_features": ["name", ...]}
 * i.e. the `features:` section of the Metarank config (YAML -> JSON unchanged) and the
 * model's `features:` list.  The dense column layout is DatasetDescriptor's: model feature
 * order, widths = each extractor's dim (S/FeatureMapping.scala:89-99).  Supported types:
 * number, word_count, string (index|onehot), interaction_count, window_count, rate,
 * interacted_with, relevancy, position, diversity, field_match/bi-encoder, boolean, vector,
 * item_age, local_time.  Anything
 * else fails with MR_ERR_UNSUPPORTED naming the feature. */
MR_API mr_status mr_schema_create(mr_ctx *ctx, const char *json, size_t len, mr_schema **out);
MR_API mr_status mr_schema_free(mr_schema *s);
/* DatasetDescriptor.dim */
MR_API int32_t mr_schema_dim(const mr_schema *s);
/* DatasetDescriptor.offsets(feature): first column of a model feature, -1 if unknown;
 * *dim_out (optional) receives its width. */
MR_API int32_t mr_schema_feature_offset(const mr_schema *s, const char *feature, int32_t *dim_out);

/* Request-side inputs the schema needs (what the extractors read from the RankingEvent
 * itself rather than from state).  kind: */
enum {
  MR_IN_REQ_F64 = 0,  /* one double per request  (number/word_count with scope ranking, string on a
                         ranking field -> encoded category index, local_time -> the mapped value) */
  MR_IN_REQ_U64 = 1,  /* one u64 per request     (rate scoped ranking.<field>: hash of the field value;
                         item_age: the RankingEvent timestamp in epoch millis)          */
  MR_IN_REQ_VEC = 2,  /* one f32[dim] per request (bi-encoder query embedding)          */
  MR_IN_ITEM_F64 = 3, /* one double per item, NaN = absent (relevancy; per-item field overrides of
                         number / string-index features, S/feature/NumberFeature.scala:84-93) */
  MR_IN_REQ_TOKENS = 4 /* one token LIST per request (field_match ngram / term / bm25,
                         S/feature/FieldMatchFeature.scala:60-93): mr_hash64 of each token the matcher's
                         tokenize() returned for the ranking field, in that order — the analyzers are Lucene's
                         and stay with the caller — plus, for bm25, the token's IDF
                         (S/feature/matcher/BM25Matcher.scala:27) as its weight; an absent field = an empty list */
};
/* Slot index of `feature` within the inputs of `kind`, or -1 when the feature reads no such
 * input.  n_out (optional) receives the total number of slots of that kind. */
MR_API int32_t mr_schema_input_slot(const mr_schema *s, int32_t kind, const char *feature, int32_t *n_out);
/* Total f32 elements per request across all MR_IN_REQ_VEC slots, and the element offset of a slot. */
MR_API int32_t mr_schema_vec_stride(const mr_schema *s);
MR_API int32_t mr_schema_vec_offset(const mr_schema *s, int32_t slot, int32_t *dim_out);

/* Device-resident replacement for Persistence.values: KVStore[Key, FeatureValue]
 * (S/fstore/Persistence.scala:39,85-89) restricted to the state the schema's extractors
 * read.  Tables live in HBM; mr_state_upsert is KVStore.put, the two batched gets of
 * FeatureValueLoader.fromStateBackend (S/fstore/FeatureValueLoader.scala:11-25) become
 * hash probes + row gathers inside the assemble kernel. */
MR_API mr_status mr_state_create(mr_ctx *ctx, mr_schema *schema, mr_state **out);
MR_API mr_status mr_state_free(mr_state *st);

/* KVStore.put(Map[Key, FeatureValue]) in a packed little-endian wire format, records
 * back to back:
 *   u16 name_len, name bytes     feature-state name == Key.feature, e.g. "ctr_click_norm"
 *   u8  scope                    0 global | 1 item | 2 user | 3 session | 4 field | 5 irf | 6 ranking
 *   scope payload                1/2/3/6: u64 id hash | 4: u64 value hash |
 *                                5: u64 value hash, u64 item hash | 0: nothing
 *   u8  kind                     0 SDouble | 1 SString | 2 SStringList | 3 SDoubleList |
 *                                4 Counter | 5 PeriodicCounter | 6 BoundedList | 7 SBoolean
 *   payload                      0: f64 | 1: u64 hash | 2: u32 n, n x u64 hash | 3: u32 n, n x f64 | 7: u8 |
 *                                4: i64 | 5: u32 n, n x i64 (PeriodicValue.value, S/model/FeatureValue.scala:30-43) |
 *                                6: u32 n, n x u64 item-id hash, newest first (BoundedListValue.values)
 * Records whose name no extractor of the schema reads are skipped (counted in *skipped).
 * Visible to mr_rank after mr_state_flush. */
MR_API mr_status mr_state_upsert(mr_state *st, const uint8_t *packed, size_t len, int64_t *applied, int64_t *skipped);
/* The write path natively: instead of refreshed FeatureValues the caller forwards the
 * extractors' raw writes (BaseFeature.writes, the files under S/feature/) and the library keeps the state:
 * FeatureValueFlow.commitWrite + computeValue (S/flow/FeatureValueFlow.scala:24-92) over the Mem* state
 * semantics (S/fstore/memory/Mem{ScalarFeature,Counter,PeriodicCounter,BoundedList}.scala,
 * PeriodicCounterFeature.fromMap S/model/Feature.scala:140-162).  Every write refreshes its value
 * (refresh interval 0), so reads are never staler than the reference's.  Records, little-endian:
 *   u16 name_len, name | u8 scope | scope payload (as in mr_state_upsert) | u8 op | i64 ts (epoch ms) | payload
 *   op 0 Put: u8 kind + value as in mr_state_upsert (kinds 0,1,2,3,7) | op 1 Increment: i64 inc |
 *   op 2 PeriodicIncrement: i64 inc | op 3 Append: u64 mr_hash64 of the appended string (an item id)
 * Visible to mr_rank after mr_state_flush. */
MR_API mr_status mr_state_apply_writes(mr_state *st, const uint8_t *packed, size_t len, int64_t *applied, int64_t *skipped);
/* Bulk load from the reference's BINARY store format: `bytes` is a stream of delimited
 * FeatureValues exactly as BinaryVCodec(compress = false, FeatureValueCodec).encodeDelimited writes them
 * (S/fstore/codec/values/BinaryVCodec.scala:45-50, S/fstore/codec/impl/FeatureValueCodec.scala:41-237 — the
 * value bytes of Persistence.values in Redis / FileKVStore): big-endian i32 length, then tag, binary Key
 * (scope tag + strings + feature name, DataOutput.writeUTF), varlong timestamp, the value, varlong expire
 * (absent for the legacy tags 0-6).  Each record is what mr_state_upsert would have received for the same
 * FeatureValue: strings become mr_hash64 of their UTF-8 bytes; ScalarValue / CounterValue /
 * PeriodicCounterValue (the PeriodicValue.value column) / BoundedListValue of SStrings are stored,
 * NumStatsValue / MapValue / FrequencyValue (no supported extractor reads them) and records whose name the
 * schema does not read count as skipped.  Like decodeDelimited (:52-62) a truncated trailing record ends
 * the stream: *consumed (optional) receives the bytes of whole records read, so consumed < len says so.
 * An unknown tag or a record shorter than its fields is MR_ERR_PARSE and nothing is applied.
 * Visible to mr_rank after mr_state_flush. */
MR_API mr_status mr_state_load_feature_values(mr_state *st, const uint8_t *bytes, size_t len, int64_t *applied,
                                              int64_t *skipped, size_t *consumed);
/* The same decoder without a state (host only): writes the mr_state_upsert records to `out` (capacity
 * out_cap; pass NULL / 0 to size the buffer) and their byte length to *out_len.  n_records counts decoded
 * FeatureValues, n_unsupported the classes dropped (see above). */
MR_API mr_status mr_feature_values_transcode(const uint8_t *bytes, size_t len, uint8_t *out, size_t out_cap,
                                             size_t *out_len, int64_t *n_records, int64_t *n_unsupported,
                                             size_t *consumed);
/* Uploads pending upserts / writes to HBM (synchronous; waits for in-flight mr_rank calls).  Until then
 * mr_rank keeps reading the snapshot of the previous flush, so writers and rankers may run concurrently.  A handful of
 * touched rows are packed and scattered by a kernel; bulk loads copy the touched row range. */
MR_API mr_status mr_state_flush(mr_state *st);
typedef struct mr_state_info {
  int64_t rows[6];        /* global, item, user, session, field, irf */
  int64_t device_bytes;
  int64_t item_row_bytes; /* fixed-width bytes gathered per item row */
} mr_state_info;
MR_API mr_status mr_state_get_info(mr_state *st, mr_state_info *out);

/* A batch of RankingEvents (S/model/Event.scala:44-67) reduced to what the hot path reads. */
typedef struct mr_rank_batch {
  int32_t n_requests;
  const int32_t *item_offsets; /* n_requests + 1; request r owns items [off[r], off[r+1]) */
  const uint64_t *item_ids;    /* mr_hash64(RankItem.id) */
  const uint64_t *user_ids;    /* per request, 0 = None; may be NULL */
  const uint64_t *session_ids; /* per request, 0 = None; may be NULL */
  const double *req_f64;       /* [n_requests x n(MR_IN_REQ_F64)], NaN = absent */
  const uint64_t *req_u64;     /* [n_requests x n(MR_IN_REQ_U64)], 0 = absent */
  const float *req_vec;        /* [n_requests x mr_schema_vec_stride] */
  const uint8_t *req_vec_present; /* [n_requests x n(MR_IN_REQ_VEC)] */
  const double *item_f64;      /* [total_items x n(MR_IN_ITEM_F64)], NaN = absent */
  /* MR_IN_REQ_TOKENS (all three may be NULL when the schema has no such slot): request r, slot s owns the
   * tokens [req_tok_offsets[r * n + s], req_tok_offsets[r * n + s + 1]) of req_tok_hashes / req_tok_weights,
   * n = n(MR_IN_REQ_TOKENS); offsets are non-decreasing, n_requests * n + 1 of them, the first is 0. */
  const int32_t *req_tok_offsets;
  const uint64_t *req_tok_hashes;
  const double *req_tok_weights; /* bm25 only; may be NULL otherwise */
  /* Optional hint for the device-batch API, where the library cannot read item_offsets on the host: the item
   * count of the largest request (0 = unknown; then every size class of the ordering kernels is enqueued and
   * returns at once where it does not apply).  Ignored by mr_rank, which reads the offsets. */
  int32_t max_items_per_request;
} mr_rank_batch;

/* Native request decoder: the body of POST /rank — one RankingEvent JSON object, or an array
 * of them for a batch — decoded with the rules of the reference's circe codecs (S/model/Event.scala:44-99:
 * id, timestamp as long | numeric string | ISO date-time with a zone, user?, session?, fields?[{name, value}],
 * items[{id, relevancy?, fields?, label?}]; S/model/Field.scala:36-58) and packed into the arrays of
 * mr_rank_batch exactly as the extractors read the request: ids / users / sessions through mr_hash64, the
 * MR_IN_* inputs the schema asks for (ranking-scoped number / word_count / string, rate scoped ranking.<field>,
 * item_age, local_time, relevancy, per-item overrides, field_match).  Two optional keys carry what stays with the
 * caller's models: "embeddings": {feature: [f32...]} (bi-encoder query vector) and "tokens": {feature: [...]}
 * (the analyzer's tokens for field_match ngram / term / bm25; the `whitespace` analyzer is built in).
 * A body the reference's decoder rejects is MR_ERR_PARSE.  The returned handle owns every array. */
typedef struct mr_requests mr_requests;
MR_API mr_status mr_requests_decode(const mr_schema *schema, const char *json, size_t len, mr_requests **out);
/* The packed batch (valid until mr_requests_free) and its item count: pass them to mr_rank. */
MR_API const mr_rank_batch *mr_requests_batch(const mr_requests *r, int32_t *total_items);
/* For the response: the id string of item `index` (0 .. total_items-1, batch order) and a request's timestamp. */
MR_API const char *mr_requests_item_id(const mr_requests *r, int32_t index, size_t *len);
MR_API int64_t mr_requests_timestamp(const mr_requests *r, int32_t request);
MR_API mr_status mr_requests_free(mr_requests *r);

/* Ranker.rerank for a batch of requests (S/ml/Ranker.scala:27-83): makeQuery
 * (FeatureValueLoader + ItemValue.fromState + ClickthroughQuery) -> model.predict ->
 * sortBy(-score).  All buffers are HOST memory; copies are part of the call.
 *   out_scores   [total_items]  score per item in request order (required)
 *   out_order    [total_items]  optional: per request, item indices (relative to the request)
 *                               in response order (stable descending)
 *   out_features [total_items x dim] optional: the assembled dense matrix (explain=true)
 * model may be NULL: then only features are assembled (TrainBuffer / explain use). */
MR_API mr_status mr_rank(mr_state *st, mr_model *model, const mr_rank_batch *batch, double *out_scores,
                         int32_t *out_order, double *out_features);

/* Same, with every pointer in `batch` and the outputs being DEVICE memory and the work
 * enqueued on `cuda_stream` without synchronising (bench.py's HBM-resident `value`).
 * Arithmetic errors (MR_ERR_ARITHMETIC) are reported by mr_rank_device_status.
 * Unlike mr_rank (which checks out a private lane per call and may be called concurrently), the
 * device-batch API keeps ONE scratch area and error flag per mr_state: use one stream at a time per
 * state, or one mr_state per stream. */
MR_API mr_status mr_rank_device(mr_state *st, mr_model *model, const mr_rank_batch *d_batch, int32_t total_items,
                                double *d_out_scores, int32_t *d_out_order, double *d_out_features,
                                void *cuda_stream);
MR_API mr_status mr_rank_device_status(mr_state *st, void *cuda_stream);

/* ------------------------------------------------------------------ mega-request sharding (multi-GPU)
 *
 * Ordinary traffic shards by REQUEST: one mr_ctx / mr_state / mr_model per GPU, each ranking its own requests
 * with mr_rank, nothing exchanged.  Only a request too large for one GPU's latency budget (10 000 items x
 * 2000 trees) is split by ITEM across the GPUs of the box: an mr_group has one member per GPU; every member
 * assembles and scores a contiguous item range, its scoring kernel stores each score straight into every
 * member's exchange buffer over NVLink (peer memory), and each member then orders the full score vector.
 * The reference has no counterpart (it scores a request on one JVM thread, S/ml/Ranker.scala:27-83); the
 * result is defined to be identical to mr_rank on one GPU, bit for bit. */
typedef struct mr_group mr_group;
#define MR_GROUP_HANDLE_BYTES 64
/* Member `rank` of a group of `world` (1..8) GPUs, for requests of up to max_items items.  The member lives on
 * ctx's device.  State and model are replicated by the caller (one mr_state / mr_model per member). */
MR_API mr_status mr_group_create(mr_ctx *ctx, int32_t rank, int32_t world, int32_t max_items, mr_group **out);
/* One process per GPU (torchrun): each member exports the CUDA IPC handle of its exchange buffer
 * (MR_GROUP_HANDLE_BYTES bytes); the caller all-gathers the handles (any transport) and hands every member the
 * world x MR_GROUP_HANDLE_BYTES table, in rank order. */
MR_API mr_status mr_group_export(mr_group *g, uint8_t *handle);
MR_API mr_status mr_group_connect(mr_group *g, const uint8_t *handles);
/* One process driving several GPUs (a JVM): members[i] is the member of rank i; enables peer access between
 * their devices.  Members may share a device (tests on a single GPU). */
MR_API mr_status mr_group_connect_local(mr_group *const *members, int32_t world);
MR_API mr_status mr_group_free(mr_group *g);
/* The contiguous item range [*lo, *hi) member `rank` owns of an n_items request: ceil(n_items / world) rounded up
 * to whole 128-item scorer tiles. */
MR_API void mr_group_slice(int32_t n_items, int32_t world, int32_t rank, int32_t *lo, int32_t *hi);
/* Collective: EVERY member calls it with the same single-request batch (n_requests == 1), concurrently — from
 * its own process, or from one thread per member.  Returns the scores of ALL items in request order and,
 * optionally, the response order, on every member.  Host buffers; copies are part of the call.  A member that
 * does not show up within 2 s fails the others with MR_ERR_CUDA instead of hanging them. */
MR_API mr_status mr_group_rank(mr_group *g, mr_state *st, mr_model *model, const mr_rank_batch *batch,
                               double *out_scores, int32_t *out_order);
/* Same with device pointers, enqueued on `cuda_stream` without synchronising (errors via mr_rank_device_status). */
MR_API mr_status mr_group_rank_device(mr_group *g, mr_state *st, mr_model *model, const mr_rank_batch *d_batch,
                                      int32_t total_items, double *d_out_scores, int32_t *d_out_order, void *cuda_stream);

/* ------------------------------------------------------------------ bi-encoder query forward
 * OnnxBiEncoder.embed (S/ml/onnx/sbert/OnnxBiEncoder.scala:13-36): a BERT-shaped sentence encoder
 * (sentence-transformers all-MiniLM-L6-v2 and relatives) run over the tokenized query, then avgpool (:38-60) over the
 * first sum(attention_mask) tokens.  Tokenization stays with the caller (the reference's HuggingFaceTokenizer): the
 * three int64 tensors are exactly what OnnxBiEncoder hands to OrtSession.run.  Dense layers run on the tcgen05 tensor
 * cores in binary16 with f32 accumulation; the residual stream, LayerNorm, softmax and pooling are f32 (pooling sums in
 * f64 like the reference).  Tolerance against an f32 evaluation of the same weights: 1e-3 on the cosine of two
 * embeddings, the bar of the reference's own test (T/ml/onnx/sbert/OnnxBiencoderTest.scala:22-25). */
typedef struct mr_encoder mr_encoder;

/* Weights: the bytes of a HuggingFace `model.safetensors` of a BertModel (tensor names
 * `embeddings.word_embeddings.weight`, `encoder.layer.<i>.attention.self.query.weight`, ...; an optional `bert.` prefix
 * and the legacy LayerNorm.gamma / beta spellings are accepted; F32 or F16 tensors).  Shapes give hidden size, layer
 * count, intermediate size, vocabulary and position count; `n_heads` and `layer_norm_eps` come from config.json
 * (12 and 1e-12 for MiniLM-L6).  hidden % 64 == 0, intermediate % 64 == 0, hidden / n_heads == 32 or 64. */
MR_API mr_status mr_encoder_load(mr_ctx *ctx, const uint8_t *safetensors, size_t len, int32_t n_heads, double layer_norm_eps,
                                 mr_encoder **out);
/* dim = OnnxSession.dim (the hidden size) */
MR_API mr_status mr_encoder_info(const mr_encoder *e, int32_t *dim, int32_t *layers, int32_t *max_tokens, int32_t *vocab);
/* embed(batch): input_ids / token_type_ids / attention_mask are [batch x seq] int64, row-major, padded to the longest
 * sequence (`padding = true`); out = [batch x dim] f32.  Host buffers; copies are part of the call. */
MR_API mr_status mr_encoder_embed(mr_encoder *e, const int64_t *input_ids, const int64_t *token_type_ids,
                                  const int64_t *attention_mask, int32_t batch, int32_t seq, float *out);
/* Same with device pointers, enqueued on `cuda_stream` without synchronising.  d_out_f64, when not null, also receives
 * the embeddings widened to f64 — the query-embedding operand of mr_rank_batch.  A token / type id outside the tables,
 * which fails the host entry point with MR_ERR_INVALID_ARG (ONNX Runtime's Gather fails the run), reads as id 0 here. */
MR_API mr_status mr_encoder_embed_device(mr_encoder *e, const int64_t *d_input_ids, const int64_t *d_token_type_ids,
                                         const int64_t *d_attention_mask, int32_t batch, int32_t seq, float *d_out,
                                         double *d_out_f64, void *cuda_stream);
MR_API mr_status mr_encoder_free(mr_encoder *e);
/* The dense layer of the forward on its own (diagnostics, tests, bench.py's tensor roofline): device pointers,
 * C[M x N] = A[M x K] W[N x K]^T + bias, then either exact-erf GELU (`gelu`) or + residual, with A, W binary16
 * (K contiguous), bias / residual / out_f32 f32, out_f16 binary16; any of bias, residual, out_f32, out_f16 may be null.
 * K % 64 == 0, N % 64 == 0; GELU together with a residual is MR_ERR_INVALID_ARG. */
MR_API mr_status mr_encoder_gemm_f16(mr_ctx *ctx, const void *d_a, const void *d_w, const float *d_bias, const float *d_residual,
                                     float *d_out_f32, void *d_out_f16, int32_t m, int32_t n, int32_t k, int32_t gelu,
                                     void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* MR_B200_H */
