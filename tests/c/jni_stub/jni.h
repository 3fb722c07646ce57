/* Test double of <jni.h> — NOT the JDK header.  The build image has no JDK, so metarank_b200/csrc/jni_shim.cpp is
 * compiled against this header-only fake: the handful of JNI types and JNIEnv members the shim uses (names and
 * signatures as in the public JNI specification, chapter 4), implemented over plain heap objects so that a C++
 * driver can call the Java_ai_metarank_b200_Native_* functions and inspect what they did (arrays, direct
 * buffers, the pending exception).  tests/test_jni_shim_cpu.py builds and runs it. */
#ifndef MR_TEST_JNI_STUB_H
#define MR_TEST_JNI_STUB_H

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>

typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef uint8_t jboolean;
typedef double jdouble;
typedef jint jsize;

#define JNI_FALSE 0
#define JNI_TRUE 1
#define JNI_ABORT 2
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL

struct _jobject {
  int kind;        /* 0 class, 1 byte[], 2 double[], 3 direct buffer, 4 long[] */
  size_t len;      /* elements (arrays) / bytes (buffers) */
  void *data;
  std::string name; /* class name */
};
typedef _jobject *jobject;
typedef jobject jclass;
typedef jobject jarray;
typedef jobject jbyteArray;
typedef jobject jdoubleArray;
typedef jobject jlongArray;

struct JNIEnv {
  /* what the test inspects */
  std::string pending_class, pending_message;
  int elements_out = 0; /* Get*ArrayElements without a matching Release */

  jclass FindClass(const char *name) { return new _jobject{0, 0, nullptr, name}; }
  jint ThrowNew(jclass cls, const char *msg) {
    pending_class = cls->name;
    pending_message = msg ? msg : "";
    delete cls;
    return 0;
  }
  bool ExceptionCheck() const { return !pending_class.empty(); }
  void ExceptionClear() { pending_class.clear(); pending_message.clear(); }
  jsize GetArrayLength(jarray a) { return (jsize)a->len; }
  jbyteArray NewByteArray(jsize n) { return new _jobject{1, (size_t)n, calloc((size_t)n + 1, 1), ""}; }
  jdoubleArray NewDoubleArray(jsize n) { return new _jobject{2, (size_t)n, calloc((size_t)n + 1, sizeof(double)), ""}; }
  /* like HotSpot, hand out a COPY: the shim must not rely on pinning */
  jbyte *GetByteArrayElements(jbyteArray a, jboolean *is_copy) {
    if (is_copy) *is_copy = JNI_TRUE;
    elements_out++;
    jbyte *p = (jbyte *)malloc(a->len + 1);
    memcpy(p, a->data, a->len);
    return p;
  }
  void ReleaseByteArrayElements(jbyteArray a, jbyte *p, jint mode) {
    if (mode != JNI_ABORT) memcpy(a->data, p, a->len);
    free(p);
    elements_out--;
  }
  jdouble *GetDoubleArrayElements(jdoubleArray a, jboolean *is_copy) {
    if (is_copy) *is_copy = JNI_TRUE;
    elements_out++;
    jdouble *p = (jdouble *)malloc((a->len + 1) * sizeof(double));
    memcpy(p, a->data, a->len * sizeof(double));
    return p;
  }
  void ReleaseDoubleArrayElements(jdoubleArray a, jdouble *p, jint mode) {
    if (mode != JNI_ABORT) memcpy(a->data, p, a->len * sizeof(double));
    free(p);
    elements_out--;
  }
  jlongArray NewLongArray(jsize n) { return new _jobject{4, (size_t)n, calloc((size_t)n + 1, sizeof(jlong)), ""}; }
  jlong *GetLongArrayElements(jlongArray a, jboolean *is_copy) {
    if (is_copy) *is_copy = JNI_TRUE;
    elements_out++;
    jlong *p = (jlong *)malloc((a->len + 1) * sizeof(jlong));
    memcpy(p, a->data, a->len * sizeof(jlong));
    return p;
  }
  void ReleaseLongArrayElements(jlongArray a, jlong *p, jint mode) {
    if (mode != JNI_ABORT) memcpy(a->data, p, a->len * sizeof(jlong));
    free(p);
    elements_out--;
  }
  jobject NewDirectByteBuffer(void *addr, jlong cap) { return new _jobject{3, (size_t)cap, addr, ""}; }
  void *GetDirectBufferAddress(jobject b) { return b && b->kind == 3 ? b->data : nullptr; }
  jlong GetDirectBufferCapacity(jobject b) { return b && b->kind == 3 ? (jlong)b->len : -1; }
};

#endif /* MR_TEST_JNI_STUB_H */
