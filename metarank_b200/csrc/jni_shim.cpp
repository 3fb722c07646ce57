// JNI shim for libmrgpu (compiled only with -DWITH_JNI -I$JAVA_HOME/include -I$JAVA_HOME/include/linux;
// there is no JDK / jni.h in the build image, so the default build leaves this file empty —
// tests/test_jni_shim_cpu.py compiles it against the test double in tests/c/jni_stub/jni.h, checks the exported
// Java_ai_metarank_b200_Native_* set against INTEGRATION.md's @native list and drives it with a fake JNIEnv).
// One native method per C-ABI entry point that the Scala adapter in INTEGRATION.md binds.
// Rules: no JNI critical section is held across a GPU call; a non-zero mr_status becomes a
// java.lang.RuntimeException carrying mr_last_error() (-> IO.raiseError -> HTTP 500, exactly how
// the reference surfaces booster failures, S/main/command/Serve.scala:101-103).
#ifdef WITH_JNI
#include <jni.h>

#include "../../include/mr_b200.h"

namespace {
void throw_status(JNIEnv *env, mr_status s) {
  if (s == MR_OK) return;
  jclass cls = env->FindClass(s == MR_ERR_ARITHMETIC ? "java/lang/ArithmeticException" : "java/lang/RuntimeException");
  env->ThrowNew(cls, mr_last_error());
}
}  // namespace

extern "C" {

JNIEXPORT jlong JNICALL Java_ai_metarank_b200_Native_init(JNIEnv *env, jclass, jint device) {
  mr_ctx *ctx = nullptr;
  throw_status(env, mr_init(device, &ctx));
  return (jlong)ctx;
}

JNIEXPORT jlong JNICALL Java_ai_metarank_b200_Native_modelLoad(JNIEnv *env, jclass, jlong ctx, jint kind, jbyteArray blob,
                                                               jint nFeatures) {
  jsize n = env->GetArrayLength(blob);
  jbyte *p = env->GetByteArrayElements(blob, nullptr);
  mr_model *m = nullptr;
  mr_status s = mr_model_load((mr_ctx *)ctx, kind, (const uint8_t *)p, (size_t)n, nFeatures, &m);
  env->ReleaseByteArrayElements(blob, p, JNI_ABORT);
  throw_status(env, s);
  return (jlong)m;
}

// Booster.predictMat(values: Array[Double], rows: Int, cols: Int): Array[Double] — the signature ltrlib's Booster has,
// so the adapter can be a one-line override.  Get/ReleaseDoubleArrayElements COPY on HotSpot (no pinning of
// primitive arrays outside critical sections, and no critical section may span a GPU call): 8 * rows * cols bytes each
// way through the JVM heap before the H2D copy.  predictMatDirect below is the zero-copy variant.
JNIEXPORT jdoubleArray JNICALL Java_ai_metarank_b200_Native_predictMat(JNIEnv *env, jclass, jlong model, jdoubleArray values,
                                                                        jint rows, jint cols) {
  jdoubleArray out = env->NewDoubleArray(rows);
  jdouble *v = env->GetDoubleArrayElements(values, nullptr);
  jdouble *o = env->GetDoubleArrayElements(out, nullptr);
  mr_status s = mr_model_predict_mat((mr_model *)model, v, rows, cols, o);
  env->ReleaseDoubleArrayElements(values, v, JNI_ABORT);
  env->ReleaseDoubleArrayElements(out, o, 0);
  throw_status(env, s);
  return out;
}

// Same on direct ByteBuffers (native order, what ByteBuffer.allocateDirect().order(nativeOrder()).asDoubleBuffer()
// views): nothing is copied in JNI, and a buffer the caller registered with cudaHostRegister is DMA'd in place.
JNIEXPORT void JNICALL Java_ai_metarank_b200_Native_predictMatDirect(JNIEnv *env, jclass, jlong model, jobject values,
                                                                    jint rows, jint cols, jobject outScores) {
  const double *v = (const double *)env->GetDirectBufferAddress(values);
  double *o = (double *)env->GetDirectBufferAddress(outScores);
  if (rows > 0 && (!v || !o || env->GetDirectBufferCapacity(values) < (jlong)rows * cols * 8 ||
                   env->GetDirectBufferCapacity(outScores) < (jlong)rows * 8)) {
    env->ThrowNew(env->FindClass("java/lang/IllegalArgumentException"), "predictMatDirect: direct buffers too small");
    return;
  }
  throw_status(env, mr_model_predict_mat((mr_model *)model, v, rows, cols, o));
}

JNIEXPORT void JNICALL Java_ai_metarank_b200_Native_modelClose(JNIEnv *env, jclass, jlong model) {
  throw_status(env, mr_model_close((mr_model *)model));
}
JNIEXPORT jboolean JNICALL Java_ai_metarank_b200_Native_modelIsClosed(JNIEnv *, jclass, jlong model) {
  return mr_model_is_closed((mr_model *)model) ? JNI_TRUE : JNI_FALSE;
}

JNIEXPORT jlong JNICALL Java_ai_metarank_b200_Native_schemaCreate(JNIEnv *env, jclass, jlong ctx, jbyteArray json) {
  jsize n = env->GetArrayLength(json);
  jbyte *p = env->GetByteArrayElements(json, nullptr);
  mr_schema *sc = nullptr;
  mr_status s = mr_schema_create((mr_ctx *)ctx, (const char *)p, (size_t)n, &sc);
  env->ReleaseByteArrayElements(json, p, JNI_ABORT);
  throw_status(env, s);
  return (jlong)sc;
}

JNIEXPORT jlong JNICALL Java_ai_metarank_b200_Native_stateCreate(JNIEnv *env, jclass, jlong ctx, jlong schema) {
  mr_state *st = nullptr;
  throw_status(env, mr_state_create((mr_ctx *)ctx, (mr_schema *)schema, &st));
  return (jlong)st;
}

// KVStore.put: `packed` is a direct ByteBuffer in the wire format of mr_state_upsert
JNIEXPORT void JNICALL Java_ai_metarank_b200_Native_stateUpsert(JNIEnv *env, jclass, jlong state, jobject packed, jint len,
                                                                jboolean flush) {
  const uint8_t *p = (const uint8_t *)env->GetDirectBufferAddress(packed);
  mr_status s = mr_state_upsert((mr_state *)state, p, (size_t)len, nullptr, nullptr);
  if (s == MR_OK && flush) s = mr_state_flush((mr_state *)state);
  throw_status(env, s);
}

// Raw extractor writes (Put / Increment / PeriodicIncrement / Append) in the wire format of mr_state_apply_writes
JNIEXPORT void JNICALL Java_ai_metarank_b200_Native_stateApplyWrites(JNIEnv *env, jclass, jlong state, jobject packed,
                                                                     jint len, jboolean flush) {
  const uint8_t *p = (const uint8_t *)env->GetDirectBufferAddress(packed);
  mr_status s = mr_state_apply_writes((mr_state *)state, p, (size_t)len, nullptr, nullptr);
  if (s == MR_OK && flush) s = mr_state_flush((mr_state *)state);
  throw_status(env, s);
}

// Warm start: a stream of BinaryStoreFormat.featureValue.encodeDelimited(...) records (direct ByteBuffer).
// Returns the number of bytes consumed (whole records).
JNIEXPORT jlong JNICALL Java_ai_metarank_b200_Native_stateLoadFeatureValues(JNIEnv *env, jclass, jlong state, jobject bytes,
                                                                            jlong len, jboolean flush) {
  const uint8_t *p = (const uint8_t *)env->GetDirectBufferAddress(bytes);
  size_t consumed = 0;
  mr_status s = mr_state_load_feature_values((mr_state *)state, p, (size_t)len, nullptr, nullptr, &consumed);
  if (s == MR_OK && flush) s = mr_state_flush((mr_state *)state);
  throw_status(env, s);
  return (jlong)consumed;
}

// Ranker.rerank for one request: item-id hashes in, scores + order out (direct buffers, no copies in JNI).
// tokOffsets / tokHashes / tokWeights: the request's MR_IN_REQ_TOKENS lists (field_match ngram/term/bm25) or null.
JNIEXPORT void JNICALL Java_ai_metarank_b200_Native_rank(JNIEnv *env, jclass, jlong state, jlong model, jint nItems,
                                                         jobject itemIds, jlong user, jlong session, jobject reqF64,
                                                         jobject reqU64, jobject reqVec, jobject reqVecPresent,
                                                         jobject itemF64, jobject tokOffsets, jobject tokHashes,
                                                         jobject tokWeights, jobject outScores, jobject outOrder,
                                                         jobject outFeatures) {
  auto addr = [&](jobject b) { return b ? env->GetDirectBufferAddress(b) : nullptr; };
  int32_t offs[2] = {0, nItems};
  uint64_t u = (uint64_t)user, se = (uint64_t)session;
  mr_rank_batch b{};
  b.n_requests = 1;
  b.item_offsets = offs;
  b.item_ids = (const uint64_t *)addr(itemIds);
  b.user_ids = &u;
  b.session_ids = &se;
  b.req_f64 = (const double *)addr(reqF64);
  b.req_u64 = (const uint64_t *)addr(reqU64);
  b.req_vec = (const float *)addr(reqVec);
  b.req_vec_present = (const uint8_t *)addr(reqVecPresent);
  b.item_f64 = (const double *)addr(itemF64);
  b.req_tok_offsets = (const int32_t *)addr(tokOffsets);
  b.req_tok_hashes = (const uint64_t *)addr(tokHashes);
  b.req_tok_weights = (const double *)addr(tokWeights);
  throw_status(env, mr_rank((mr_state *)state, (mr_model *)model, &b, (double *)addr(outScores), (int32_t *)addr(outOrder),
                            (double *)addr(outFeatures)));
}

// ---- one mega-request across the GPUs of a box (mr_group_*)
JNIEXPORT jlong JNICALL Java_ai_metarank_b200_Native_groupCreate(JNIEnv *env, jclass, jlong ctx, jint rank, jint world, jint maxItems) {
  mr_group *g = nullptr;
  throw_status(env, mr_group_create((mr_ctx *)ctx, rank, world, maxItems, &g));
  return (jlong)g;
}

// members(i) = the group member of rank i, all created in this process
JNIEXPORT void JNICALL Java_ai_metarank_b200_Native_groupConnectLocal(JNIEnv *env, jclass, jlongArray members) {
  const jsize n = env->GetArrayLength(members);
  jlong *p = env->GetLongArrayElements(members, nullptr);
  mr_group *gs[8] = {};
  for (jsize k = 0; k < n && k < 8; k++) gs[k] = (mr_group *)p[k];
  env->ReleaseLongArrayElements(members, p, JNI_ABORT);
  throw_status(env, n > 8 ? MR_ERR_INVALID_ARG : mr_group_connect_local(gs, n));
}

// Collective: every member calls it with the same request, concurrently (one blocking call per member).  Direct buffers;
// every member receives the full scores and order.
JNIEXPORT void JNICALL Java_ai_metarank_b200_Native_groupRank(JNIEnv *env, jclass, jlong group, jlong state, jlong model,
                                                              jint nItems, jobject itemIds, jlong user, jlong session,
                                                              jobject reqF64, jobject reqU64, jobject reqVec,
                                                              jobject reqVecPresent, jobject itemF64, jobject outScores,
                                                              jobject outOrder) {
  auto addr = [&](jobject b) { return b ? env->GetDirectBufferAddress(b) : nullptr; };
  int32_t offs[2] = {0, nItems};
  uint64_t u = (uint64_t)user, se = (uint64_t)session;
  mr_rank_batch b{};
  b.n_requests = 1;
  b.item_offsets = offs;
  b.item_ids = (const uint64_t *)addr(itemIds);
  b.user_ids = &u;
  b.session_ids = &se;
  b.req_f64 = (const double *)addr(reqF64);
  b.req_u64 = (const uint64_t *)addr(reqU64);
  b.req_vec = (const float *)addr(reqVec);
  b.req_vec_present = (const uint8_t *)addr(reqVecPresent);
  b.item_f64 = (const double *)addr(itemF64);
  throw_status(env, mr_group_rank((mr_group *)group, (mr_state *)state, (mr_model *)model, &b, (double *)addr(outScores),
                                  (int32_t *)addr(outOrder)));
}

JNIEXPORT void JNICALL Java_ai_metarank_b200_Native_groupFree(JNIEnv *env, jclass, jlong group) {
  throw_status(env, mr_group_free((mr_group *)group));
}

// OnnxBiEncoder (S/ml/onnx/sbert/OnnxBiEncoder.scala:11-36): weights = the model.safetensors bytes (direct buffer),
// embed = session.run + avgpool over the three [batch x seq] long tensors the tokenizer produced (direct buffers, native order)
JNIEXPORT jlong JNICALL Java_ai_metarank_b200_Native_encoderLoad(JNIEnv *env, jclass, jlong ctx, jobject weights, jlong len, jint nHeads,
                                                                 jdouble layerNormEps) {
  mr_encoder *e = nullptr;
  throw_status(env, mr_encoder_load((mr_ctx *)ctx, (const uint8_t *)env->GetDirectBufferAddress(weights), (size_t)len, nHeads,
                                    layerNormEps, &e));
  return (jlong)e;
}

JNIEXPORT jint JNICALL Java_ai_metarank_b200_Native_encoderDim(JNIEnv *env, jclass, jlong enc) {
  int32_t dim = 0;
  throw_status(env, mr_encoder_info((const mr_encoder *)enc, &dim, nullptr, nullptr, nullptr));
  return dim;
}

JNIEXPORT void JNICALL Java_ai_metarank_b200_Native_encoderEmbed(JNIEnv *env, jclass, jlong enc, jobject inputIds, jobject tokenTypeIds,
                                                                 jobject attentionMask, jint batch, jint seq, jobject out) {
  auto addr = [&](jobject b) { return b ? env->GetDirectBufferAddress(b) : nullptr; };
  throw_status(env, mr_encoder_embed((mr_encoder *)enc, (const int64_t *)addr(inputIds), (const int64_t *)addr(tokenTypeIds),
                                     (const int64_t *)addr(attentionMask), batch, seq, (float *)addr(out)));
}

JNIEXPORT void JNICALL Java_ai_metarank_b200_Native_encoderFree(JNIEnv *env, jclass, jlong enc) {
  throw_status(env, mr_encoder_free((mr_encoder *)enc));
}

}  // extern "C"
#endif  // WITH_JNI
