/* fastnn_jni.c - native half of nnet.NativeNN over the C ABI of libfastnn.so (include/fastnn.h).
 * Build where a JDK exists:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../include fastnn_jni.c -L../fastneighbornet_b200 \
 *       -lfastnn -o libfastnn_jni.so
 * This repository only syntax-checks it against a stub jni.h (tests/stubs/jni.h, tests/test_abi.py). */
#include <jni.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "fastnn.h"

static void throw_named(JNIEnv* e, const char* cls, const char* msg) {
    if ((*e)->ExceptionCheck(e)) return;                 /* never stack a second exception on a pending one */
    jclass c = (*e)->FindClass(e, cls);
    if (c) (*e)->ThrowNew(e, c, msg);
}
static void throw_rt(JNIEnv* e, int rc) {
    char msg[600];
    snprintf(msg, sizeof msg, "libfastnn error %d: %s", rc, fnn_last_error());
    throw_named(e, "java/lang/RuntimeException", msg);
}
static void throw_oom(JNIEnv* e, const char* what) { throw_named(e, "java/lang/OutOfMemoryError", what); }
static void throw_arg(JNIEnv* e, const char* what) { throw_named(e, "java/lang/IllegalArgumentException", what); }

/* fnn_order keeps its last context between calls (include/fastnn.h); give the device memory back with the library */
JNIEXPORT void JNICALL JNI_OnUnload(JavaVM* vm, void* reserved) {
    (void)vm; (void)reserved;
    fnn_release_cache();
}

static void set_opts(fnn_opts* o, jint mode, jint mult, jboolean additive, jlong seed) {
    fnn_default_opts(o);
    o->mode = mode;
    o->mult = mult;
    o->additive = additive ? 1 : 0;
    o->seed = seed;
}

static jintArray to_jints(JNIEnv* e, const int32_t* v, jsize len) {
    jintArray out = (*e)->NewIntArray(e, len);
    if (out) (*e)->SetIntArrayRegion(e, out, 0, len, (const jint*)v);
    return out;
}

JNIEXPORT jintArray JNICALL Java_nnet_NativeNN_order(JNIEnv* e, jclass cls, jobjectArray D, jint n, jint mode, jint mult,
                                                     jboolean additive, jlong seed) {
    (void)cls;
    if (n < 1 || D == NULL || (*e)->GetArrayLength(e, D) < n) { throw_arg(e, "NativeNN.order: need nTaxa >= 1 and D with nTaxa rows"); return NULL; }
    double* host = (double*)malloc((size_t)n * (size_t)n * sizeof(double));
    int32_t* ord = (int32_t*)malloc(((size_t)n + 1) * sizeof(int32_t));
    jintArray out = NULL;
    if (!host || !ord) throw_oom(e, "NativeNN.order: host staging buffers");
    else {
        int ok = 1;
        for (jint i = 0; i < n && ok; ++i) {                        /* double[][] rows -> row-major */
            jdoubleArray row = (jdoubleArray)(*e)->GetObjectArrayElement(e, D, i);
            if (row == NULL || (*e)->ExceptionCheck(e)) { ok = 0; if (!(*e)->ExceptionCheck(e)) throw_arg(e, "NativeNN.order: null row in D"); break; }
            if ((*e)->GetArrayLength(e, row) < n) { ok = 0; throw_arg(e, "NativeNN.order: short row in D"); }
            else {
                (*e)->GetDoubleArrayRegion(e, row, 0, n, host + (size_t)i * (size_t)n);
                if ((*e)->ExceptionCheck(e)) ok = 0;                /* no JNI calls other than cleanup with an exception pending */
            }
            (*e)->DeleteLocalRef(e, row);
        }
        if (ok) {
            fnn_opts o;
            set_opts(&o, mode, mult, additive, seed);
            const int rc = fnn_order(&o, host, NULL, n, ord);
            if (rc) throw_rt(e, rc);
            else out = to_jints(e, ord, n + 1);
        }
    }
    free(host);
    free(ord);
    return out;
}

JNIEXPORT jintArray JNICALL Java_nnet_NativeNN_orderFromFile(JNIEnv* e, jclass cls, jstring path, jint n, jint mode, jint mult,
                                                             jboolean additive, jlong seed) {
    (void)cls;
    if (n < 1 || path == NULL) { throw_arg(e, "NativeNN.orderFromFile: need nTaxa >= 1 and a path"); return NULL; }
    const char* p = (*e)->GetStringUTFChars(e, path, NULL);
    int32_t* ord = (int32_t*)malloc(((size_t)n + 1) * sizeof(int32_t));
    jintArray out = NULL;
    if (p && !ord) throw_oom(e, "NativeNN.orderFromFile: ordering buffer");
    if (p && ord) {
        fnn_opts o;
        set_opts(&o, mode, mult, additive, seed);
        const int rc = fnn_order(&o, NULL, p, n, ord);
        if (rc) throw_rt(e, rc);
        else out = to_jints(e, ord, n + 1);
    }
    if (p) (*e)->ReleaseStringUTFChars(e, path, p);
    free(ord);
    return out;
}

JNIEXPORT jdoubleArray JNICALL Java_nnet_NativeNN_splitWeights(JNIEnv* e, jclass cls, jintArray ordering, jdoubleArray dUpper, jint n) {
    (void)cls;
    if (n < 4 || n > 20000) { throw_arg(e, "NativeNN.splitWeights: 4 <= nTaxa <= 20000 (fnn_split_weights)"); return NULL; }   /* before any size arithmetic */
    const size_t np = (size_t)n * ((size_t)n - 1) / 2;
    if (ordering == NULL || dUpper == NULL || (*e)->GetArrayLength(e, ordering) < n + 1 || (size_t)(*e)->GetArrayLength(e, dUpper) < np) {
        throw_arg(e, "NativeNN.splitWeights: ordering needs nTaxa+1 entries, dUpper nTaxa(nTaxa-1)/2");
        return NULL;
    }
    int32_t* ord = (int32_t*)malloc(((size_t)n + 1) * sizeof(int32_t));
    double* d = (double*)malloc(np * sizeof(double));
    double* x = (double*)malloc(np * sizeof(double));
    jdoubleArray out = NULL;
    if (!ord || !d || !x) throw_oom(e, "NativeNN.splitWeights: host buffers");
    else {
        (*e)->GetIntArrayRegion(e, ordering, 0, n + 1, (jint*)ord);
        (*e)->GetDoubleArrayRegion(e, dUpper, 0, (jsize)np, d);
        fnn_opts o;
        fnn_default_opts(&o);
        const int rc = fnn_split_weights(&o, ord, d, n, x, NULL);
        if (rc) throw_rt(e, rc);
        else {
            out = (*e)->NewDoubleArray(e, (jsize)np);
            if (out) (*e)->SetDoubleArrayRegion(e, out, 0, (jsize)np, x);
        }
    }
    free(ord);
    free(d);
    free(x);
    return out;
}

JNIEXPORT jintArray JNICALL Java_nnet_NativeNN_network(JNIEnv* e, jclass cls, jstring phylipPath, jstring nexusPath, jint mode,
                                                       jint mult, jboolean additive, jlong seed, jdouble cutoff,
                                                       jboolean printDistances) {
    (void)cls;
    enum { NAME_LEN = 128 };
    const char* in = (*e)->GetStringUTFChars(e, phylipPath, NULL);
    const char* outp = nexusPath ? (*e)->GetStringUTFChars(e, nexusPath, NULL) : NULL;
    jintArray result = NULL;
    int64_t n = 0;
    int rc = in ? fnn_phylip_taxa(in, &n) : FNN_E_ARG;
    double* D = NULL;
    char* names = NULL;
    int32_t *ord = NULL, *si = NULL, *sj = NULL;
    double* w = NULL;
    /* only ~3.7 n splits survive the cutoff: start with room for 16 n, retry with the exact count */
    int64_t cap = 0;
    if (!rc && n < 1) rc = FNN_E_ARG;
    if (!rc) {
        const int64_t np = n * (n - 1) / 2;
        cap = np < 16 * n ? np : 16 * n;
        D = (double*)malloc((size_t)n * (size_t)n * sizeof(double));
        names = (char*)malloc((size_t)n * NAME_LEN);
        ord = (int32_t*)malloc(((size_t)n + 1) * sizeof(int32_t));
        si = (int32_t*)malloc(((size_t)cap + 1) * sizeof(int32_t));
        sj = (int32_t*)malloc(((size_t)cap + 1) * sizeof(int32_t));
        w = (double*)malloc(((size_t)cap + 1) * sizeof(double));
        if (!D || !names || !ord || !si || !sj || !w) rc = FNN_E_NOMEM;
    }
    if (!rc) rc = fnn_read_phylip(in, n, D, names, NAME_LEN, 0);
    if (!rc) {
        fnn_opts o;
        set_opts(&o, mode, mult, additive, seed);
        int64_t kept = 0;
        if (n < 4) {                                               /* nothing to weigh: ordering only */
            rc = fnn_order(&o, D, NULL, n, ord);
        } else {
            rc = fnn_network(&o, D, n, cutoff, ord, si, sj, w, cap, &kept);
            if (rc == FNN_E_ARG && kept > cap) {                   /* more splits kept than room: *n_out holds the count */
                cap = kept;
                free(si); free(sj); free(w);
                si = (int32_t*)malloc(((size_t)cap + 1) * sizeof(int32_t));
                sj = (int32_t*)malloc(((size_t)cap + 1) * sizeof(int32_t));
                w = (double*)malloc(((size_t)cap + 1) * sizeof(double));
                rc = (si && sj && w) ? fnn_network(&o, D, n, cutoff, ord, si, sj, w, cap, &kept) : FNN_E_NOMEM;
            }
        }
        if (!rc) rc = fnn_write_nexus(outp, n, names, NAME_LEN, printDistances ? D : NULL, ord, si, sj, w, kept, 0);
        if (!rc) result = to_jints(e, ord, (jsize)(n + 1));
    }
    if (rc == FNN_E_NOMEM) throw_oom(e, "NativeNN.network: host buffers");
    else if (rc) throw_rt(e, rc);
    free(D); free(names); free(ord); free(si); free(sj); free(w);
    if (in) (*e)->ReleaseStringUTFChars(e, phylipPath, in);
    if (outp) (*e)->ReleaseStringUTFChars(e, nexusPath, outp);
    return result;
}
