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

/* Mean evaluated path length (internal nodes visited per item per tree) of the last
 * `rows` scored by mr_model_count_path(): the d̄ of SURVEY.md §8(d)'s B_item. */
MR_API mr_status mr_model_count_path(mr_model *m, const double *values, int32_t rows, int32_t cols, double *mean_path);

/* Booster.close() / isClosed() (S/ml/rank/LambdaMARTRanker.scala:361-365).  close is
 * idempotent; in-flight predicts finish (CachedModelStore may dispose a model that is
 * still in use, S/fstore/cache/CachedModelStore.scala:39-42).  free releases the handle. */
MR_API mr_status mr_model_close(mr_model *m);
MR_API int32_t mr_model_is_closed(mr_model *m);
MR_API mr_status mr_model_free(mr_model *m);

/* Tuning knobs for experiments (bench.py / tests); defaults are chosen per model.
 * key: "threads" (items per CTA), "chunk_kb", "variant" (0 = lock-step, 1 = free-running). */
MR_API mr_status mr_model_set_option(mr_model *m, const char *key, int32_t value);

/* ------------------------------------------------------------------ final ordering */

/* Ranker.rerank's `sortBy(-_.score)` (S/ml/Ranker.scala:52-67): stable, descending,
 * java.lang.Double.compare on the negated score (NaN last; 0.0 before -0.0).
 * Sorts `n_requests` independent requests laid out back to back: request r owns
 * scores[offsets[r] .. offsets[r+1]); order receives item indices relative to the
 * request start.  Host buffers. */
MR_API mr_status mr_rank_order(mr_ctx *ctx, const double *scores, const int32_t *offsets, int32_t n_requests,
                        int32_t *order);

#ifdef __cplusplus
}
#endif
#endif /* MR_B200_H */
