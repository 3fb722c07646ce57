// Drives the JNI shim (metarank_b200/csrc/jni_shim.cpp, built with -DWITH_JNI against tests/c/jni_stub/jni.h)
// with a fake JNIEnv.  Prints one line per check: "ok <name>" / "FAIL <name> ...".  argv[1] == "gpu": also the
// scoring entry points against the C ABI (needs a B200); otherwise host-only checks (no device present).
#include <jni.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "mr_b200.h"

extern "C" {
jlong Java_ai_metarank_b200_Native_init(JNIEnv *, jclass, jint);
jlong Java_ai_metarank_b200_Native_modelLoad(JNIEnv *, jclass, jlong, jint, jbyteArray, jint);
jdoubleArray Java_ai_metarank_b200_Native_predictMat(JNIEnv *, jclass, jlong, jdoubleArray, jint, jint);
void Java_ai_metarank_b200_Native_predictMatDirect(JNIEnv *, jclass, jlong, jobject, jint, jint, jobject);
void Java_ai_metarank_b200_Native_modelClose(JNIEnv *, jclass, jlong);
jboolean Java_ai_metarank_b200_Native_modelIsClosed(JNIEnv *, jclass, jlong);
jlong Java_ai_metarank_b200_Native_schemaCreate(JNIEnv *, jclass, jlong, jbyteArray);
}

static int failures = 0;
static void check(bool ok, const char *name, const char *detail = "") {
  printf("%s %s %s\n", ok ? "ok" : "FAIL", name, ok ? "" : detail);
  if (!ok) failures++;
}

static jbyteArray bytes(JNIEnv &env, const std::string &s) {
  jbyteArray a = env.NewByteArray((jsize)s.size());
  memcpy(a->data, s.data(), s.size());
  return a;
}

// a two-tree LightGBM model text, enough for the parser and the kernels
static const char *kModel =
    "tree\nversion=v4\nnum_class=1\nnum_tree_per_iteration=1\nlabel_index=0\nmax_feature_idx=1\nobjective=lambdarank\n"
    "feature_names=a b\nfeature_infos=[-1:1] [-1:1]\ntree_sizes=0 0\n\n"
    "Tree=0\nnum_leaves=2\nnum_cat=0\nsplit_feature=0\nsplit_gain=1\nthreshold=0.5\ndecision_type=2\nleft_child=-1\nright_child=-2\n"
    "leaf_value=0.25 -0.75\nleaf_weight=1 1\nleaf_count=1 1\ninternal_value=0\ninternal_weight=0\ninternal_count=2\nis_linear=0\nshrinkage=1\n\n"
    "Tree=1\nnum_leaves=2\nnum_cat=0\nsplit_feature=1\nsplit_gain=1\nthreshold=-0.5\ndecision_type=2\nleft_child=-1\nright_child=-2\n"
    "leaf_value=1 2\nleaf_weight=1 1\nleaf_count=1 1\ninternal_value=0\ninternal_weight=0\ninternal_count=2\nis_linear=0\nshrinkage=1\n\n"
    "end of trees\n";

int main(int argc, char **argv) {
  const bool gpu = argc > 1 && !strcmp(argv[1], "gpu");
  JNIEnv env;
  // schemaCreate is host-only with ctx == 0: a good config yields a handle, a bad one a RuntimeException with the parser's text
  jlong sc = Java_ai_metarank_b200_Native_schemaCreate(&env, nullptr, 0,
      bytes(env, "{\"features\":[{\"name\":\"p\",\"type\":\"number\",\"scope\":\"item\",\"source\":\"metadata.p\"}],\"model_features\":[\"p\"]}"));
  check(sc != 0 && !env.ExceptionCheck(), "schemaCreate host-only");
  Java_ai_metarank_b200_Native_schemaCreate(&env, nullptr, 0, bytes(env, "{\"features\":[{\"name\":\"p\",\"type\":\"nonsense\"}]}"));
  check(env.pending_class == "java/lang/RuntimeException" && !env.pending_message.empty(), "status -> RuntimeException", env.pending_class.c_str());
  env.ExceptionClear();
  check(env.elements_out == 0, "every Get*ArrayElements released");
  if (!gpu) {
    jlong ctx = Java_ai_metarank_b200_Native_init(&env, nullptr, 0);
    check(ctx == 0 && env.pending_class == "java/lang/RuntimeException" && env.pending_message.find("no CPU fallback") != std::string::npos,
          "init without a device throws (no CPU fallback)", env.pending_message.c_str());
    env.ExceptionClear();
  } else {
    jlong ctx = Java_ai_metarank_b200_Native_init(&env, nullptr, 0);
    check(ctx != 0 && !env.ExceptionCheck(), "init", env.pending_message.c_str());
    jlong m = Java_ai_metarank_b200_Native_modelLoad(&env, nullptr, ctx, 0, bytes(env, kModel), 2);
    check(m != 0 && !env.ExceptionCheck(), "modelLoad", env.pending_message.c_str());
    const int rows = 5;
    const double X[rows * 2] = {0.5, -0.5, 0.6, -0.5, 0.0, 0.0, NAN, -1.0, 1.0, NAN};
    // decision_type 2 = default-left bit, missing type None: a NaN input is read as 0.0 (Tree::NumericalDecision)
    const double want[rows] = {0.25 + 1, -0.75 + 1, 0.25 + 2, 0.25 + 1, -0.75 + 2};
    jdoubleArray v = env.NewDoubleArray(rows * 2);
    memcpy(v->data, X, sizeof X);
    jdoubleArray out = Java_ai_metarank_b200_Native_predictMat(&env, nullptr, m, v, rows, 2);
    bool same = out && !env.ExceptionCheck() && out->len == (size_t)rows && !memcmp(out->data, want, sizeof want);
    check(same, "predictMat (array copy-in / copy-back)", env.pending_message.c_str());
    double c_abi[rows];
    check(mr_model_predict_mat((mr_model *)m, X, rows, 2, c_abi) == MR_OK && !memcmp(c_abi, want, sizeof want), "C ABI agrees");
    std::vector<double> xin(X, X + rows * 2), o2(rows, -1.0);
    Java_ai_metarank_b200_Native_predictMatDirect(&env, nullptr, m, env.NewDirectByteBuffer(xin.data(), sizeof X), rows, 2,
                                                  env.NewDirectByteBuffer(o2.data(), rows * 8));
    check(!env.ExceptionCheck() && !memcmp(o2.data(), want, sizeof want), "predictMatDirect (zero copy)", env.pending_message.c_str());
    Java_ai_metarank_b200_Native_predictMatDirect(&env, nullptr, m, env.NewDirectByteBuffer(xin.data(), 8), rows, 2,
                                                  env.NewDirectByteBuffer(o2.data(), rows * 8));
    check(env.pending_class == "java/lang/IllegalArgumentException", "predictMatDirect rejects short buffers");
    env.ExceptionClear();
    check(Java_ai_metarank_b200_Native_modelIsClosed(&env, nullptr, m) == JNI_FALSE, "isClosed false");
    Java_ai_metarank_b200_Native_modelClose(&env, nullptr, m);
    Java_ai_metarank_b200_Native_modelClose(&env, nullptr, m);  // idempotent
    check(!env.ExceptionCheck() && Java_ai_metarank_b200_Native_modelIsClosed(&env, nullptr, m) == JNI_TRUE, "close is idempotent");
    Java_ai_metarank_b200_Native_predictMat(&env, nullptr, m, v, rows, 2);
    check(env.pending_class == "java/lang/RuntimeException" && env.pending_message.find("closed") != std::string::npos, "predict after close throws");
    env.ExceptionClear();
    check(env.elements_out == 0, "every Get*ArrayElements released (gpu)");
  }
  printf("%s\n", failures ? "FAILED" : "ALL OK");
  return failures ? 1 : 0;
}
