/* jni_mock.c -- TEST-ONLY fake JVM: a JNIEnv function table with the slots libhadoop-4mc.so uses
 * (4mc_b200/host/jni_min.h) and fake Lz4Compressor / Lz4Decompressor objects, so that the JNI shim
 * can be driven from pytest without a JDK ("No JVM here"). */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../4mc_b200/host/jni_min.h"

typedef struct { void *uncompressedDirectBuf; int uncompressedDirectBufLen; void *compressedDirectBuf; int compressedDirectBufLen;
                 int directBufferSize; } FakeObj;
struct fm_jfieldID_ { int which; };
static struct fm_jfieldID_ F[5] = {{0}, {1}, {2}, {3}, {4}};
static char g_exc[512];
static int g_threw;

static jclass m_FindClass(JNIEnv *e, const char *n) { (void)e; return (jclass)n; }
static jint m_ThrowNew(JNIEnv *e, jclass c, const char *msg) { (void)e; snprintf(g_exc, sizeof g_exc, "%s: %s", (const char *)c, msg); g_threw = 1; return 0; }
static void m_DeleteLocalRef(JNIEnv *e, jobject o) { (void)e; (void)o; }
static jfieldID m_GetFieldID(JNIEnv *e, jclass c, const char *name, const char *sig)
{
    (void)e; (void)c; (void)sig;
    if (!strcmp(name, "uncompressedDirectBuf")) return &F[0];
    if (!strcmp(name, "uncompressedDirectBufLen")) return &F[1];
    if (!strcmp(name, "compressedDirectBuf")) return &F[2];
    if (!strcmp(name, "compressedDirectBufLen")) return &F[3];
    if (!strcmp(name, "directBufferSize")) return &F[4];
    return NULL;
}
static jobject m_GetObjectField(JNIEnv *e, jobject o, jfieldID f) { (void)e; FakeObj *x = (FakeObj *)o; return f->which == 0 ? x->uncompressedDirectBuf : x->compressedDirectBuf; }
static jint m_GetIntField(JNIEnv *e, jobject o, jfieldID f) { (void)e; FakeObj *x = (FakeObj *)o; return f->which == 1 ? x->uncompressedDirectBufLen : f->which == 3 ? x->compressedDirectBufLen : x->directBufferSize; }
static void m_SetIntField(JNIEnv *e, jobject o, jfieldID f, jint v) { (void)e; FakeObj *x = (FakeObj *)o; if (f->which == 1) x->uncompressedDirectBufLen = v; else if (f->which == 3) x->compressedDirectBufLen = v; }
static jstring m_NewStringUTF(JNIEnv *e, const char *s) { (void)e; return (jstring)s; }
static void *m_GetCritical(JNIEnv *e, jarray a, jboolean *c) { (void)e; (void)c; return a; }
static void m_ReleaseCritical(JNIEnv *e, jarray a, void *p, jint m) { (void)e; (void)a; (void)p; (void)m; }
static void *m_GetDirectBufferAddress(JNIEnv *e, jobject b) { (void)e; return b; }     /* a "ByteBuffer" is its address */

static struct JNINativeInterface_ g_tbl;
static JNIEnv g_env = &g_tbl;
static void *g_lib;

static void *sym(const char *name) { void *p = dlsym(g_lib, name); if (!p) { fprintf(stderr, "missing %s\n", name); abort(); } return p; }

int mock_open(const char *libpath)
{
    memset(&g_tbl, 0, sizeof g_tbl);
    g_tbl.FindClass = m_FindClass; g_tbl.ThrowNew = m_ThrowNew; g_tbl.DeleteLocalRef = m_DeleteLocalRef;
    g_tbl.GetFieldID = m_GetFieldID; g_tbl.GetObjectField = m_GetObjectField; g_tbl.GetIntField = m_GetIntField;
    g_tbl.SetIntField = m_SetIntField; g_tbl.NewStringUTF = m_NewStringUTF;
    g_tbl.GetPrimitiveArrayCritical = m_GetCritical; g_tbl.ReleasePrimitiveArrayCritical = m_ReleaseCritical;
    g_tbl.GetDirectBufferAddress = m_GetDirectBufferAddress;
    g_lib = dlopen(libpath, RTLD_NOW);
    if (!g_lib) { fprintf(stderr, "%s\n", dlerror()); return -1; }
    ((void (*)(JNIEnv *, jclass))sym("Java_com_fing_compression_fourmc_Lz4Compressor_initIDs"))(&g_env, NULL);
    ((void (*)(JNIEnv *, jclass))sym("Java_com_fing_compression_fourmc_Lz4Decompressor_initIDs"))(&g_env, NULL);
    ((void (*)(JNIEnv *, jclass))sym("Java_com_fing_compression_fourmc_ZstdDecompressor_initIDs"))(&g_env, NULL);
    ((void (*)(JNIEnv *, jclass))sym("Java_com_fing_compression_fourmc_ZstdCompressor_initIDs"))(&g_env, NULL);
    return 0;
}

/* table layout check: byte offsets of the members the shim calls through */
int mock_slot_offsets(int *out)
{
    const char *b = (const char *)&g_tbl;
    const void *m[] = {&g_tbl.FindClass, &g_tbl.ThrowNew, &g_tbl.DeleteLocalRef, &g_tbl.GetFieldID, &g_tbl.GetObjectField,
                       &g_tbl.GetIntField, &g_tbl.GetLongField, &g_tbl.SetIntField, &g_tbl.SetLongField, &g_tbl.NewStringUTF,
                       &g_tbl.GetPrimitiveArrayCritical, &g_tbl.ReleasePrimitiveArrayCritical, &g_tbl.GetDirectBufferAddress};
    for (int i = 0; i < 13; i++) out[i] = (int)((const char *)m[i] - b);
    return 13;
}

int mock_compress(int which, int hc_level, unsigned char *in, int n, unsigned char *out, int *len_after, int *threw, char *msg)
{
    FakeObj o = {in, n, out, 0, 4 << 20};
    g_threw = 0; g_exc[0] = 0;
    int r;
    if (which == 0) r = ((jint (*)(JNIEnv *, jobject))sym("Java_com_fing_compression_fourmc_Lz4Compressor_compressBytesDirect"))(&g_env, &o);
    else if (which == 1) r = ((jint (*)(JNIEnv *, jobject))sym("Java_com_fing_compression_fourmc_Lz4Compressor_compressBytesDirectMC"))(&g_env, &o);
    else r = ((jint (*)(JNIEnv *, jobject, jint))sym("Java_com_fing_compression_fourmc_Lz4Compressor_compressBytesDirectHC"))(&g_env, &o, hc_level);
    *len_after = o.uncompressedDirectBufLen; *threw = g_threw; strcpy(msg, g_exc);
    return r;
}

int mock_decompress(unsigned char *in, int c, unsigned char *out, int cap, int *len_after, int *threw, char *msg)
{
    FakeObj o = {out, 0, in, c, cap};
    g_threw = 0; g_exc[0] = 0;
    int r = ((jint (*)(JNIEnv *, jobject))sym("Java_com_fing_compression_fourmc_Lz4Decompressor_decompressBytesDirect"))(&g_env, &o);
    *len_after = o.compressedDirectBufLen; *threw = g_threw; strcpy(msg, g_exc);
    return r;
}

int mock_xxh(int cls, unsigned char *buf, int off, int len, int seed)
{
    const char *names[] = {"Java_com_fing_compression_fourmc_Lz4Compressor_xxhash32", "Java_com_fing_compression_fourmc_Lz4Decompressor_xxhash32",
                           "Java_com_fing_compression_fourmc_ZstdCompressor_xxhash32", "Java_com_fing_compression_fourmc_ZstdDecompressor_xxhash32"};
    return ((jint (*)(JNIEnv *, jclass, jbyteArray, jint, jint, jint))sym(names[cls]))(&g_env, NULL, buf, off, len, seed);
}

int mock_bound(int n) { return ((jint (*)(JNIEnv *, jclass, jint))sym("Java_com_fing_compression_fourmc_Lz4Compressor_compressBound"))(&g_env, NULL, n); }

int mock_zstd_decompress(unsigned char *in, int c, unsigned char *out, int cap, int *len_after, int *threw, char *msg)
{
    FakeObj o = {out, 0, in, c, cap};
    g_threw = 0; g_exc[0] = 0;
    int r = ((jint (*)(JNIEnv *, jobject))sym("Java_com_fing_compression_fourmc_ZstdDecompressor_decompressBytesDirect"))(&g_env, &o);
    *len_after = o.compressedDirectBufLen; *threw = g_threw; strcpy(msg, g_exc);
    return r;
}

int mock_zstd_compress(int which, int hc_level, unsigned char *in, int n, unsigned char *out, int *len_after, int *threw, char *msg)
{
    FakeObj o = {in, n, out, 0, 4 << 20};
    g_threw = 0; g_exc[0] = 0;
    int r;
    if (which == 0) r = ((jint (*)(JNIEnv *, jobject))sym("Java_com_fing_compression_fourmc_ZstdCompressor_compressBytesDirect"))(&g_env, &o);
    else if (which == 1) r = ((jint (*)(JNIEnv *, jobject))sym("Java_com_fing_compression_fourmc_ZstdCompressor_compressBytesDirectMC"))(&g_env, &o);
    else r = ((jint (*)(JNIEnv *, jobject, jint))sym("Java_com_fing_compression_fourmc_ZstdCompressor_compressBytesDirectHC"))(&g_env, &o, hc_level);
    *len_after = o.uncompressedDirectBufLen; *threw = g_threw; strcpy(msg, g_exc);
    return r;
}

int mock_zstd_bound(int n) { return ((jint (*)(JNIEnv *, jclass, jint))sym("Java_com_fing_compression_fourmc_ZstdCompressor_compressBound"))(&g_env, NULL, n); }

/* the streaming zstd natives resolve (class initialisation) but throw when used */
int mock_zstd_stream_throws(char *msg)
{
    g_threw = 0;
    ((jlong (*)(JNIEnv *, jclass))sym("Java_com_fing_compression_fourmc_zstd_ZstdStreamCompressor_createCStream"))(&g_env, NULL);
    strcpy(msg, g_exc);
    return g_threw;
}
