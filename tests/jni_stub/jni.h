/* jni.h -- NOT the JDK's header.  This image has no JDK; this stub declares the subset of the JNI API that
 * integration/jni/gs_jni.cpp uses, with the signatures of the Java Native Interface specification (JNI 1.6, "JNI Functions"),
 * so that the shim is at least type-checked by the test suite (tests/test_capi_symbols.py).  integration/build.sh uses the real
 * header wherever a JDK exists. */
#ifndef GS_TEST_JNI_STUB_H
#define GS_TEST_JNI_STUB_H
#include <stdint.h>

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
#define JNI_FALSE 0
#define JNI_TRUE 1

typedef int32_t jint;
typedef int64_t jlong;
typedef int16_t jshort;
typedef uint8_t jboolean;
typedef double jdouble;
typedef jint jsize;

class _jobject {};
class _jclass : public _jobject {};
class _jarray : public _jobject {};
class _jintArray : public _jarray {};
class _jlongArray : public _jarray {};
class _jshortArray : public _jarray {};
class _jobjectArray : public _jarray {};
class _jstring : public _jobject {};
typedef _jobject* jobject;
typedef _jclass* jclass;
typedef _jarray* jarray;
typedef _jintArray* jintArray;
typedef _jlongArray* jlongArray;
typedef _jshortArray* jshortArray;
typedef _jobjectArray* jobjectArray;
typedef _jstring* jstring;
struct _jmethodID;
typedef _jmethodID* jmethodID;

struct JNIEnv {
    jclass FindClass(const char* name);
    jint ThrowNew(jclass clazz, const char* msg);
    jsize GetArrayLength(jarray array);
    jint* GetIntArrayElements(jintArray array, jboolean* isCopy);
    void ReleaseIntArrayElements(jintArray array, jint* elems, jint mode);
    jlong* GetLongArrayElements(jlongArray array, jboolean* isCopy);
    void ReleaseLongArrayElements(jlongArray array, jlong* elems, jint mode);
    void* GetPrimitiveArrayCritical(jarray array, jboolean* isCopy);
    void ReleasePrimitiveArrayCritical(jarray array, void* carray, jint mode);
    jobject NewDirectByteBuffer(void* address, jlong capacity);
    void* GetDirectBufferAddress(jobject buf);
    jobjectArray NewObjectArray(jsize len, jclass clazz, jobject init);
    void SetObjectArrayElement(jobjectArray array, jsize index, jobject val);
    jlongArray NewLongArray(jsize len);
    void SetLongArrayRegion(jlongArray array, jsize start, jsize len, const jlong* buf);
    jmethodID GetStaticMethodID(jclass clazz, const char* name, const char* sig);
    jobject CallStaticObjectMethod(jclass clazz, jmethodID methodID, ...);
    const char* GetStringUTFChars(jstring str, jboolean* isCopy);
    void ReleaseStringUTFChars(jstring str, const char* chars);
};
#endif
