/*
 * jni_shim.c -- libhadoop-4mc.so: the 35 JNI entry points of the reference's native library
 * (SURVEY.md 8b), over the C-ABI of lib4mcgpu.so.  Drop-in: put this library on java.library.path
 * and start the JVM with -Dcom.fing.compression.fourmc.use.libpath=true
 * (FourMcNativeCodeLoader.java:52-53,101-111); the Java classes are unchanged.
 *
 * Block path (18 symbols): Lz4Compressor / Lz4Decompressor natives run on the GPU
 *   <- native/jniCompressor.c:57-194, native/jniDecompressor.c:56-118.
 * ZstdCompressor / ZstdDecompressor natives (4mz writing and reading) run on the GPU
 *   <- native/jniZstdCompressor.c:59-199, native/jniZstdDecompressor.c:58-120.
 * The zstd STREAMING natives (17 symbols) are exported so that class initialisation succeeds, and
 * throw java.lang.InternalError when used (the streaming ZstCodec is out of scope).
 *
 * Same conventions as the reference: field ids cached by initIDs; input = first *DirectBufLen bytes
 * of the direct buffer; on success the length field is reset to 0; on failure InternalError is
 * thrown with "<function> returned: <value>"; NULL buffer addresses return 0 silently.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef FOURMC_REAL_JNI
#include <jni.h>
#else
#include "jni_min.h"
#endif
#include "../../include/fourmc.h"

#define EXC_LEN 256
#define PKG(cls, fn) Java_com_fing_compression_fourmc_##cls##_##fn
#define ZPKG(cls, fn) Java_com_fing_compression_fourmc_zstd_##cls##_##fn

static void throw_ie(JNIEnv *env, const char *msg)
{
    jclass c = (*env)->FindClass(env, "java/lang/InternalError");      /* jnihelper.h:41-48 */
    if (c) { (*env)->ThrowNew(env, c, msg); (*env)->DeleteLocalRef(env, c); }
}

/* one GPU context per calling thread: the Java objects are synchronized per instance, many
 * instances may call from different threads, and a fourmc_ctx is single-threaded */
static __thread fourmc_ctx *t_ctx;
static __thread char *t_bounce;            /* xxhash32: the array's bytes leave the critical region through here */
static __thread size_t t_bounce_cap;
/* a thread that ends gives its context (streams, device workspaces) and bounce buffer back */
static pthread_key_t t_key;
static pthread_once_t t_once = PTHREAD_ONCE_INIT;
static void thread_exit(void *p)
{
    (void)p;
    if (t_ctx) { fourmc_ctx_destroy(t_ctx); t_ctx = NULL; }
    free(t_bounce); t_bounce = NULL; t_bounce_cap = 0;
}
static void make_key(void) { pthread_key_create(&t_key, thread_exit); }
static fourmc_ctx *ctx_get(JNIEnv *env)
{
    if (!t_ctx) {
        if (fourmc_ctx_create(&t_ctx, -1) != FOURMC_OK) {
            t_ctx = NULL;
            throw_ie(env, "lib4mcgpu: no usable CUDA device (there is no CPU fallback)");
            return NULL;
        }
        pthread_once(&t_once, make_key);
        pthread_setspecific(t_key, (void *)1);          /* non-NULL: the destructor runs at thread exit */
    }
    return t_ctx;
}

/* ---- LZ4 compressor: native/jniCompressor.c --------------------------------------------------- */

static jfieldID c_uncompressedDirectBuf, c_uncompressedDirectBufLen, c_compressedDirectBuf, c_directBufferSize;

JNIEXPORT void JNICALL PKG(Lz4Compressor, initIDs)(JNIEnv *env, jclass cls)
{   /* :57-70 (finish / finished are looked up by the reference but never used) */
    c_uncompressedDirectBuf = (*env)->GetFieldID(env, cls, "uncompressedDirectBuf", "Ljava/nio/ByteBuffer;");
    c_uncompressedDirectBufLen = (*env)->GetFieldID(env, cls, "uncompressedDirectBufLen", "I");
    c_compressedDirectBuf = (*env)->GetFieldID(env, cls, "compressedDirectBuf", "Ljava/nio/ByteBuffer;");
    c_directBufferSize = (*env)->GetFieldID(env, cls, "directBufferSize", "I");
}

static jint lz4_compress_common(JNIEnv *env, jobject self, int level, const char *fn)
{
    jobject ub = (*env)->GetObjectField(env, self, c_uncompressedDirectBuf);
    jint ulen = (*env)->GetIntField(env, self, c_uncompressedDirectBufLen);
    jobject cb = (*env)->GetObjectField(env, self, c_compressedDirectBuf);
    const char *src = (const char *)(*env)->GetDirectBufferAddress(env, ub);
    char *dst = (char *)(*env)->GetDirectBufferAddress(env, cb);
    if (src == 0 || dst == 0) return 0;                                        /* :86-88 */
    fourmc_ctx *ctx = ctx_get(env);
    if (!ctx) return 0;
    /* the Java side sizes compressedDirectBuf with compressBound(directBufferSize) (Lz4Compressor.java:136-144) */
    int r = fourmc_lz4_compress(ctx, level, src, ulen, dst, fourmc_lz4_compress_bound(ulen));
    if (r > 0) {
        (*env)->SetIntField(env, self, c_uncompressedDirectBufLen, 0);         /* :93-94 */
    } else {
        char msg[EXC_LEN];
        snprintf(msg, sizeof msg, "%s returned: %d", fn, r);                   /* :96-99 */
        throw_ie(env, msg);
    }
    return r;
}

JNIEXPORT jint JNICALL PKG(Lz4Compressor, compressBytesDirect)(JNIEnv *env, jobject self)
{ return lz4_compress_common(env, self, 1, "LZ4_compress"); }                   /* :72-103 */

JNIEXPORT jint JNICALL PKG(Lz4Compressor, compressBytesDirectMC)(JNIEnv *env, jobject self)
{ return lz4_compress_common(env, self, 2, "LZ4_compressMC"); }                 /* :105-135 */

JNIEXPORT jint JNICALL PKG(Lz4Compressor, compressBytesDirectHC)(JNIEnv *env, jobject self, jint clevel)
{ return lz4_compress_common(env, self, clevel >= 8 ? 4 : 3, "LZ4_compressHC2"); }   /* :138-168 */

JNIEXPORT jint JNICALL PKG(Lz4Compressor, compressBound)(JNIEnv *env, jclass cls, jint n)
{ (void)env; (void)cls; return fourmc_lz4_compress_bound(n); }                  /* :171-174 */

static jint xxhash32_common(JNIEnv *env, jbyteArray buf, jint off, jint len, jint seed)
{   /* :178-194.  The reference hashes inside the critical region; a GPU round trip (or the creation of a CUDA
     * context: seconds) there would hold the collector up, so the context is made first and the region is only as
     * long as one memcpy into this thread's bounce buffer. */
    fourmc_ctx *ctx = ctx_get(env);
    if (!ctx) return 0;
    if (len < 0) len = 0;
    if ((size_t)len > t_bounce_cap) {
        char *nb = (char *)realloc(t_bounce, (size_t)len + 4096);
        if (!nb) { throw_ie(env, "lib4mcgpu: out of memory"); return 0; }
        t_bounce = nb; t_bounce_cap = (size_t)len + 4096;
    }
    char *in = (char *)(*env)->GetPrimitiveArrayCritical(env, buf, 0);
    if (in == NULL) return 0;
    if (len) memcpy(t_bounce, in + off, (size_t)len);
    (*env)->ReleasePrimitiveArrayCritical(env, buf, in, 0);
    int st = FOURMC_E_CUDA;
    const jint h = (jint)fourmc_xxh32(ctx, t_bounce, (size_t)len, (uint32_t)seed, &st);
    if (st != FOURMC_OK) throw_ie(env, "lib4mcgpu: XXH32 failed on the device (there is no CPU fallback)");
    return h;
}

JNIEXPORT jint JNICALL PKG(Lz4Compressor, xxhash32)(JNIEnv *env, jclass cls, jbyteArray b, jint off, jint len, jint seed)
{ (void)cls; return xxhash32_common(env, b, off, len, seed); }

/* ---- LZ4 decompressor: native/jniDecompressor.c ------------------------------------------------ */

static jfieldID d_compressedDirectBuf, d_compressedDirectBufLen, d_uncompressedDirectBuf, d_directBufferSize;

JNIEXPORT void JNICALL PKG(Lz4Decompressor, initIDs)(JNIEnv *env, jclass cls)
{   /* :56-64 -- note the declared types are java/nio/Buffer here, ByteBuffer in the compressor */
    d_compressedDirectBuf = (*env)->GetFieldID(env, cls, "compressedDirectBuf", "Ljava/nio/Buffer;");
    d_compressedDirectBufLen = (*env)->GetFieldID(env, cls, "compressedDirectBufLen", "I");
    d_uncompressedDirectBuf = (*env)->GetFieldID(env, cls, "uncompressedDirectBuf", "Ljava/nio/Buffer;");
    d_directBufferSize = (*env)->GetFieldID(env, cls, "directBufferSize", "I");
}

JNIEXPORT jint JNICALL PKG(Lz4Decompressor, decompressBytesDirect)(JNIEnv *env, jobject self)
{   /* :67-100 */
    jobject cb = (*env)->GetObjectField(env, self, d_compressedDirectBuf);
    jint clen = (*env)->GetIntField(env, self, d_compressedDirectBufLen);
    jobject ub = (*env)->GetObjectField(env, self, d_uncompressedDirectBuf);
    jint cap = (*env)->GetIntField(env, self, d_directBufferSize);
    char *dst = (char *)(*env)->GetDirectBufferAddress(env, ub);
    const char *src = (const char *)(*env)->GetDirectBufferAddress(env, cb);
    if (dst == 0 || src == 0) return 0;
    fourmc_ctx *ctx = ctx_get(env);
    if (!ctx) return 0;
    int r = fourmc_lz4_decompress_safe(ctx, src, clen, dst, cap);
    if (r >= 0) {
        (*env)->SetIntField(env, self, d_compressedDirectBufLen, 0);
    } else {
        char msg[EXC_LEN];
        snprintf(msg, sizeof msg, "LZ4_decompress_safe returned: %d", r);
        throw_ie(env, msg);
    }
    return r;
}

JNIEXPORT jint JNICALL PKG(Lz4Decompressor, xxhash32)(JNIEnv *env, jclass cls, jbyteArray b, jint off, jint len, jint seed)
{ (void)cls; return xxhash32_common(env, b, off, len, seed); }

/* ---- ZSTD block natives (4mz); the streaming zstd natives below are exported only ------------------ */

static jint unsupported(JNIEnv *env, const char *what)
{
    char msg[EXC_LEN];
    snprintf(msg, sizeof msg, "%s: zstd streaming is not implemented by lib4mcgpu (block codecs only)", what);
    throw_ie(env, msg);
    return 0;
}

/* ZstdCompressor: native/jniZstdCompressor.c:59-175 -- implemented (4mz writing) */
static jfieldID zc_uncompressedDirectBuf, zc_uncompressedDirectBufLen, zc_compressedDirectBuf, zc_directBufferSize;

JNIEXPORT void JNICALL PKG(ZstdCompressor, initIDs)(JNIEnv *env, jclass cls)
{   /* :59-70 */
    zc_uncompressedDirectBuf = (*env)->GetFieldID(env, cls, "uncompressedDirectBuf", "Ljava/nio/ByteBuffer;");
    zc_uncompressedDirectBufLen = (*env)->GetFieldID(env, cls, "uncompressedDirectBufLen", "I");
    zc_compressedDirectBuf = (*env)->GetFieldID(env, cls, "compressedDirectBuf", "Ljava/nio/ByteBuffer;");
    zc_directBufferSize = (*env)->GetFieldID(env, cls, "directBufferSize", "I");
}

static jint zstd_compress_common(JNIEnv *env, jobject self, int level)
{
    jobject ub = (*env)->GetObjectField(env, self, zc_uncompressedDirectBuf);
    jint ulen = (*env)->GetIntField(env, self, zc_uncompressedDirectBufLen);
    jobject cb = (*env)->GetObjectField(env, self, zc_compressedDirectBuf);
    const char *src = (const char *)(*env)->GetDirectBufferAddress(env, ub);
    char *dst = (char *)(*env)->GetDirectBufferAddress(env, cb);
    if (src == 0 || dst == 0) return 0;                                        /* :87-89 */
    fourmc_ctx *ctx = ctx_get(env);
    if (!ctx) return 0;
    /* the reference passes a 1 GiB capacity "enforced before in Java code" (:93): the Java side sizes
     * compressedDirectBuf with compressBound(directBufferSize) */
    long long r = fourmc_zstd_compress(ctx, level, src, (size_t)(unsigned)ulen, dst, fourmc_zstd_compress_bound((size_t)(unsigned)ulen));
    if (r >= 0) {
        (*env)->SetIntField(env, self, zc_uncompressedDirectBufLen, 0);        /* :96-97 */
    } else {
        char msg[EXC_LEN];
        snprintf(msg, sizeof msg, "%s returned: %lu", "ZSTD_compress", (unsigned long)r);   /* :99-102 */
        throw_ie(env, msg);
    }
    return (jint)r;
}

JNIEXPORT jint JNICALL PKG(ZstdCompressor, compressBytesDirect)(JNIEnv *env, jobject self) { return zstd_compress_common(env, self, 1); }     /* :72-105, level 1 */
JNIEXPORT jint JNICALL PKG(ZstdCompressor, compressBytesDirectMC)(JNIEnv *env, jobject self) { return zstd_compress_common(env, self, 2); }   /* :107-138, level 3 */
JNIEXPORT jint JNICALL PKG(ZstdCompressor, compressBytesDirectHC)(JNIEnv *env, jobject self, jint l) { return zstd_compress_common(env, self, l >= 12 ? 4 : 3); }   /* :140-172 */
JNIEXPORT jint JNICALL PKG(ZstdCompressor, compressBound)(JNIEnv *env, jclass cls, jint n)
{ (void)env; (void)cls; return n + (n >> 8) + (n < (128 << 10) ? (((128 << 10) - n) >> 11) : 0); }   /* ZSTD_COMPRESSBOUND, zstd.h:204 */
JNIEXPORT jint JNICALL PKG(ZstdCompressor, xxhash32)(JNIEnv *env, jclass cls, jbyteArray b, jint off, jint len, jint seed)
{ (void)cls; return xxhash32_common(env, b, off, len, seed); }
/* ZstdDecompressor: native/jniZstdDecompressor.c:58-101 -- implemented (4mz reading) */
static jfieldID z_compressedDirectBuf, z_compressedDirectBufLen, z_uncompressedDirectBuf, z_directBufferSize;

JNIEXPORT void JNICALL PKG(ZstdDecompressor, initIDs)(JNIEnv *env, jclass cls)
{   /* :58-65 */
    z_compressedDirectBuf = (*env)->GetFieldID(env, cls, "compressedDirectBuf", "Ljava/nio/Buffer;");
    z_compressedDirectBufLen = (*env)->GetFieldID(env, cls, "compressedDirectBufLen", "I");
    z_uncompressedDirectBuf = (*env)->GetFieldID(env, cls, "uncompressedDirectBuf", "Ljava/nio/Buffer;");
    z_directBufferSize = (*env)->GetFieldID(env, cls, "directBufferSize", "I");
}
JNIEXPORT jint JNICALL PKG(ZstdDecompressor, decompressBytesDirect)(JNIEnv *env, jobject self)
{   /* :68-101 */
    jobject cb = (*env)->GetObjectField(env, self, z_compressedDirectBuf);
    jint clen = (*env)->GetIntField(env, self, z_compressedDirectBufLen);
    jobject ub = (*env)->GetObjectField(env, self, z_uncompressedDirectBuf);
    jint cap = (*env)->GetIntField(env, self, z_directBufferSize);
    char *dst = (char *)(*env)->GetDirectBufferAddress(env, ub);
    const char *src = (const char *)(*env)->GetDirectBufferAddress(env, cb);
    if (dst == 0 || src == 0) return 0;
    fourmc_ctx *ctx = ctx_get(env);
    if (!ctx) return 0;
    int r = (int)fourmc_zstd_decompress(ctx, src, (size_t)(unsigned)clen, dst, (size_t)(unsigned)cap);
    if (r >= 0) {
        (*env)->SetIntField(env, self, z_compressedDirectBufLen, 0);
    } else {
        char msg[EXC_LEN];
        snprintf(msg, sizeof msg, "LZ4_decompress_safe returned: %d", r);      /* the reference's own text, :96 */
        throw_ie(env, msg);
    }
    return r;
}
JNIEXPORT jint JNICALL PKG(ZstdDecompressor, xxhash32)(JNIEnv *env, jclass cls, jbyteArray b, jint off, jint len, jint seed)
{ (void)cls; return xxhash32_common(env, b, off, len, seed); }

/* ---- zstd streaming natives (native/jniZstd.c, jniZStreamCompressor.c, jniZStreamDecompressor.c) */

JNIEXPORT jboolean JNICALL ZPKG(Zstd, isError)(JNIEnv *env, jclass cls, jlong code) { (void)env; (void)cls; return code < 0; }
JNIEXPORT jstring JNICALL ZPKG(Zstd, getErrorName)(JNIEnv *env, jclass cls, jlong code)
{ (void)cls; (void)code; return (*env)->NewStringUTF(env, "zstd streaming is not implemented by lib4mcgpu"); }
JNIEXPORT jint JNICALL ZPKG(Zstd, cStreamInSize)(JNIEnv *env, jclass cls) { (void)env; (void)cls; return 128 << 10; }
JNIEXPORT jint JNICALL ZPKG(Zstd, cStreamOutSize)(JNIEnv *env, jclass cls) { (void)env; (void)cls; return (128 << 10) + 512 + 3; }
JNIEXPORT jint JNICALL ZPKG(Zstd, dStreamInSize)(JNIEnv *env, jclass cls) { (void)env; (void)cls; return (128 << 10) + 3; }
JNIEXPORT jint JNICALL ZPKG(Zstd, dStreamOutSize)(JNIEnv *env, jclass cls) { (void)env; (void)cls; return 128 << 10; }

JNIEXPORT void JNICALL ZPKG(ZstdStreamCompressor, initIDs)(JNIEnv *env, jclass cls) { (void)env; (void)cls; }
JNIEXPORT jlong JNICALL ZPKG(ZstdStreamCompressor, createCStream)(JNIEnv *env, jclass cls) { (void)cls; return unsupported(env, "ZSTD_createCStream"); }
JNIEXPORT jint JNICALL ZPKG(ZstdStreamCompressor, freeCStream)(JNIEnv *env, jclass cls, jlong s) { (void)env; (void)cls; (void)s; return 0; }
JNIEXPORT jint JNICALL ZPKG(ZstdStreamCompressor, initCStream)(JNIEnv *env, jclass cls, jlong s, jint l) { (void)cls; (void)s; (void)l; return unsupported(env, "ZSTD_initCStream"); }
JNIEXPORT jint JNICALL ZPKG(ZstdStreamCompressor, compressStream)(JNIEnv *env, jobject self, jlong s, jobject d, jint dl, jobject i, jint il)
{ (void)self; (void)s; (void)d; (void)dl; (void)i; (void)il; return unsupported(env, "ZSTD_compressStream"); }
JNIEXPORT jint JNICALL ZPKG(ZstdStreamCompressor, endStream)(JNIEnv *env, jobject self, jlong s, jobject d, jint o, jint l)
{ (void)self; (void)s; (void)d; (void)o; (void)l; return unsupported(env, "ZSTD_endStream"); }

JNIEXPORT void JNICALL ZPKG(ZstdStreamDecompressor, initIDs)(JNIEnv *env, jclass cls) { (void)env; (void)cls; }
JNIEXPORT jlong JNICALL ZPKG(ZstdStreamDecompressor, createDStream)(JNIEnv *env, jclass cls) { (void)cls; return unsupported(env, "ZSTD_createDStream"); }
JNIEXPORT jint JNICALL ZPKG(ZstdStreamDecompressor, freeDStream)(JNIEnv *env, jclass cls, jlong s) { (void)env; (void)cls; (void)s; return 0; }
JNIEXPORT jint JNICALL ZPKG(ZstdStreamDecompressor, initDStream)(JNIEnv *env, jclass cls, jlong s) { (void)cls; (void)s; return unsupported(env, "ZSTD_initDStream"); }
JNIEXPORT jint JNICALL ZPKG(ZstdStreamDecompressor, decompressStream)(JNIEnv *env, jobject self, jlong s, jobject d, jint dl, jobject i, jint il)
{ (void)self; (void)s; (void)d; (void)dl; (void)i; (void)il; return unsupported(env, "ZSTD_decompressStream"); }
