/*
 * fourmc_cli.c -- `4mc` command line tool over lib4mcgpu.so.
 *
 * Flag-, message-level- and exit-code-compatible with the reference CLI (native/4mccli.c:170-361,
 * native/4mc.c:135-161,220-386,896-934): -1..-4 level, -d decode, -t test, -c stdout, -f overwrite,
 * -v / -q verbosity, -V version, -h help, -z zstd (4mz),
 * "stdin" / "stdout" / "null" file names, automatic .4mc output names when stdout is a terminal.
 * Exit codes: 1 generic, 2 input, 3 output, 4 content.  All compression, checksum and index work is
 * done by the GPU through the C-ABI; this file only moves bytes between files and host memory.
 * The whole input is held in memory (round 1: no streaming of files larger than RAM): regular files are
 * mapped -- the input read-only, the output as the destination buffer itself -- so that no byte is copied
 * twice on the host.  FOURMC_CLI_TIMING=1 prints the wall-clock phases to stderr.
 */
#define _GNU_SOURCE
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include "../../include/fourmc.h"

#define WELCOME "*** 4mc CLI (lib4mcgpu, B200), 4mc format v1 ***\n"
#define EXT_4MC ".4mc"
#define EXT_4MZ ".4mz"

static int display = 2;      /* 0 none, 1 errors, 2 results, 3 progress, 4 information (4mccli.c:174-177) */
static const char *prog = "4mc";

#define SAY(l, ...) do { if (display >= (l)) fprintf(stderr, __VA_ARGS__); } while (0)
#define DIE(code, ...) do { SAY(1, __VA_ARGS__); SAY(1, "\n"); exit(code); } while (0)

static int usage(void)
{
    fprintf(stderr, "Usage :\n      %s [arg] [input] [output]\n\n", prog);
    fprintf(stderr, "input   : a filename\n          with no FILE, or when FILE is - or stdin, read standard input\n");
    fprintf(stderr, "Arguments :\n -z     : zstd compression (4mz) \n -1     : Fast compression (default) \n"
                    " -2     : Medium compression \n -3     : High compression \n -4     : Ultra compression \n"
                    " -d     : decompression (default for %s and %s exts)\n -f     : overwrite output without prompting \n"
                    " -V     : display Version number and exit\n -v     : verbose mode\n -q     : quiet mode\n"
                    " -h     : display help and exit\n", EXT_4MC, EXT_4MZ);
    return 0;
}

static void badusage(void)
{
    SAY(1, "Incorrect command line arguments\n");
    if (display >= 1) usage();
    exit(1);
}

static unsigned char *read_all(const char *name, size_t *n)
{
    FILE *f = strcmp(name, "stdin") ? fopen(name, "rb") : stdin;
    if (!f) DIE(2, "Cannot open input file: %s", name);                       /* 4mc.c:206 */
    size_t cap = 1 << 20, len = 0;
    if (f != stdin && fseek(f, 0, SEEK_END) == 0) { long sz = ftell(f); if (sz > 0) cap = (size_t)sz + 1; fseek(f, 0, SEEK_SET); }
    unsigned char *buf = (unsigned char *)malloc(cap);
    if (!buf) DIE(1, "Allocation error : not enough memory");
    for (;;) {
        if (len == cap) { cap *= 2; buf = (unsigned char *)realloc(buf, cap); if (!buf) DIE(1, "Allocation error : not enough memory"); }
        size_t r = fread(buf + len, 1, cap - len, f);
        if (r == 0) break;
        len += r;
    }
    if (f != stdin) fclose(f);
    *n = len;
    return buf;
}

static double now_s(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + (double)t.tv_nsec * 1e-9;
}

/* input: a regular file is mapped (no copy), anything else is read */
static unsigned char *map_or_read(const char *name, size_t *n, int *mapped)
{
    *mapped = 0;
    if (strcmp(name, "stdin")) {
        int fd = open(name, O_RDONLY);
        if (fd < 0) DIE(2, "Cannot open input file: %s", name);                /* 4mc.c:206 */
        struct stat st;
        if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {
            void *p = mmap(NULL, (size_t)st.st_size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
            close(fd);
            if (p != MAP_FAILED) { *n = (size_t)st.st_size; *mapped = 1; return (unsigned char *)p; }
        } else {
            close(fd);
        }
    }
    return read_all(name, n);
}

/* destination buffer of `cap` bytes: the output file itself when it is a regular file (mapped, cut to the final
 * size by finish_out), plain memory otherwise (stdout, /dev/null, pipes) */
static unsigned char *out_buffer(FILE *fo, size_t cap, int populate, int *mapped)
{
    *mapped = 0;
    struct stat st;
    const int fd = fileno(fo);
    if (fo != stdout && cap && fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && ftruncate(fd, (off_t)cap) == 0) {
        void *p = mmap(NULL, cap, PROT_READ | PROT_WRITE, MAP_SHARED | (populate ? MAP_POPULATE : 0), fd, 0);
        if (p != MAP_FAILED) { *mapped = 1; return (unsigned char *)p; }
        if (ftruncate(fd, 0)) { /* fall through to memory + fwrite */ }
    }
    unsigned char *out = (unsigned char *)malloc(cap ? cap : 1);
    if (!out) DIE(1, "Allocation error : not enough memory");
    return out;
}

/* used == (size_t)-1: the call failed, leave an empty file behind like a reader that wrote nothing */
static int finish_out(FILE *fo, unsigned char *out, size_t cap, size_t used, int mapped)
{
    if (mapped) {
        munmap(out, cap);
        return ftruncate(fileno(fo), used == (size_t)-1 ? 0 : (off_t)used) == 0;
    }
    const int ok = used == (size_t)-1 || fwrite(out, 1, used, fo) == used;
    free(out);
    return ok;
}

static FILE *open_out(const char *name, int overwrite)
{
    if (!strcmp(name, "stdout")) return stdout;
    if (strcmp(name, "/dev/null")) {
        FILE *t = fopen(name, "rb");
        if (t) {                                                              /* 4mc.c:186-201 */
            fclose(t);
            if (!overwrite) {
                SAY(2, "Warning : %s already exists\n", name);
                SAY(2, "Overwrite ? (Y/N) : ");
                if (display <= 1) DIE(3, "Operation aborted : %s already exists", name);
                int ch = getchar();
                if (ch != 'Y' && ch != 'y') DIE(3, "Operation aborted : %s already exists", name);
            }
        }
    }
    FILE *f = fopen(name, "w+b");                     /* read-write: the file may become the mapped destination */
    if (!f) f = fopen(name, "wb");
    if (!f) DIE(3, "Cannot open output file: %s", name);
    return f;
}

int main(int argc, char **argv)
{
    int level = 0, decode = 0, force_stdout = 0, force_compress = 0, overwrite = 0, zstd = 0;
    const char *in_name = NULL, *out_name = NULL;
    char *dyn = NULL;
    prog = argv[0];

    for (int i = 1; i < argc; i++) {
        char *a = argv[i];
        if (!a) continue;
        if (a[0] == '-') {
            if (a[1] == 0) { if (!in_name) in_name = "stdin"; else out_name = "stdout"; }
            while (a[1] != 0) {
                a++;
                if (*a >= '0' && *a <= '9') {
                    level = 0;
                    while (*a >= '0' && *a <= '9') { level = level * 10 + (*a - '0'); a++; }
                    a--;
                    continue;
                }
                switch (*a) {
                case 'V': fprintf(stderr, WELCOME); return 0;
                case 'h': case 'H': usage(); return 0;
                case 'z': zstd = 1; force_compress = 1; break;
                case 'l': break;                                              /* parsed, unused (4mccli.c:234) */
                case 'd': decode = 1; break;
                case 'c': force_stdout = 1; out_name = "stdout"; display = 1; break;
                case 't': decode = 1; out_name = "/dev/null"; break;
                case 'f': overwrite = 1; break;
                case 'v': display = 4; break;
                case 'q': display--; break;
                default: badusage();
                }
            }
            continue;
        }
        if (!in_name) { in_name = a; continue; }
        if (!out_name) { out_name = strcmp(a, "null") ? a : "/dev/null"; continue; }
    }
    SAY(3, WELCOME);
    if (!in_name) in_name = "stdin";
    if (!strcmp(in_name, "stdin") && isatty(0)) badusage();

    while (!out_name) {                                                       /* 4mccli.c:283-333 */
        if (!isatty(1)) { out_name = "stdout"; break; }
        size_t l = strlen(in_name);
        if (!decode && !force_compress && l > 4 && (!strcmp(in_name + l - 4, EXT_4MC) || !strcmp(in_name + l - 4, EXT_4MZ))) decode = 1;
        if (!decode) {
            dyn = (char *)calloc(1, l + 5);
            strcpy(dyn, in_name); strcpy(dyn + l, zstd ? EXT_4MZ : EXT_4MC);
            out_name = dyn;
            SAY(2, "Compressed filename will be : %s \n", out_name);
            break;
        }
        if (l > 4 && !strcmp(in_name + l - 4, EXT_4MZ)) zstd = 1;
        else if (!(l > 4 && !strcmp(in_name + l - 4, EXT_4MC))) { SAY(1, "Cannot determine an output filename\n"); badusage(); }
        dyn = (char *)calloc(1, l + 1);
        memcpy(dyn, in_name, l - 4);
        out_name = dyn;
        SAY(2, "Decoding file %s \n", out_name);
    }
    if (!strcmp(in_name, "stdin") && !strcmp(out_name, "stdout") && display == 2) display = 1;
    if (!strcmp(out_name, "stdout") && isatty(1) && !force_stdout) badusage();

    clock_t t0 = clock();
    const int timing = getenv("FOURMC_CLI_TIMING") != NULL;
    const double w0 = now_s();
    size_t n = 0;
    int in_mapped = 0, out_mapped = 0;
    unsigned char *in = map_or_read(in_name, &n, &in_mapped);
    FILE *fo = open_out(out_name, overwrite);
    const double w1 = now_s();
    fourmc_ctx *ctx = NULL;
    if (fourmc_ctx_create(&ctx, -1) != FOURMC_OK) DIE(1, "lib4mcgpu: no usable CUDA device (there is no CPU fallback)");
    const double w2 = now_s();
    double w3 = w2, w4 = w2;

    if (!decode) {
        SAY(2, zstd ? "Compression: ZSTD\n" : "Compression: LZ4\n");                /* 4mc.c:241, :410 */
        if ((display == 2) && (level > 1)) display = 3;
        size_t cap = fourmc_4mc_bound(n);
        unsigned char *out = out_buffer(fo, cap, 0, &out_mapped);
        w3 = now_s();
        long long c = zstd ? fourmc_4mz_compress_host(ctx, level < 1 ? 1 : level, in, n, out, cap)
                           : fourmc_4mc_compress_host(ctx, level < 1 ? 1 : level, in, n, out, cap);
        w4 = now_s();
        if (c < 0) {
            finish_out(fo, out, cap, (size_t)-1, out_mapped);
            DIE(c == FOURMC_E_OUTPUT ? 3 : 1, "Compression failed: %s", fourmc_last_error(ctx));
        }
        if (!finish_out(fo, out, cap, (size_t)c, out_mapped)) DIE(3, "Write error : cannot write compressed block");
        SAY(2, "Compressed (%s) %llu bytes into %llu bytes ==> %.2f%% (Ratio=%.3f)\n",
            level <= 1 ? "fast" : level == 2 ? "medium" : level == 3 ? "high" : "ultra", (unsigned long long)n,
            (unsigned long long)c, n ? (double)c / n * 100 : 0.0, c ? (double)n / c : 0.0);
    } else {
        SAY(3, zstd ? "Compression: ZSTD\n" : "Compression: LZ4\n");
        long long sz = zstd ? fourmc_4mz_decoded_size_host(in, n) : fourmc_4mc_decoded_size_host(in, n);
        size_t cap = sz > 0 ? (size_t)sz : 0;
        /* on a malformed container still decode what precedes the damage, like the serial reader */
        if (sz < 0) cap = n * 4 + (64 << 20);
        unsigned char *out = out_buffer(fo, cap, sz > 0, &out_mapped);
        w3 = now_s();
        long long d = zstd ? fourmc_4mz_decompress_host(ctx, in, n, out, cap) : fourmc_4mc_decompress_host(ctx, in, n, out, cap);
        w4 = now_s();
        if (d < 0) {
            int code = d == FOURMC_E_INPUT ? 2 : d == FOURMC_E_OUTPUT ? 3 : d == FOURMC_E_CONTENT ? 4 : 1;
            finish_out(fo, out, cap, (size_t)-1, out_mapped);
            DIE(code, "%s", code == 4 ? "Decoding Failed ! Corrupted input detected !" :
                            code == 2 ? "Read error : cannot read next block" : fourmc_last_error(ctx));
        }
        if (!finish_out(fo, out, cap, (size_t)d, out_mapped)) DIE(3, "Write error : cannot write decoded block");
        SAY(2, "Successfully decoded %llu bytes \n", (unsigned long long)d);
    }
    {
        double s = (double)(clock() - t0) / CLOCKS_PER_SEC;
        SAY(4, "Done in %.2f s ==> %.2f MB/s\n", s, s > 0 ? (double)n / s / 1024 / 1024 : 0.0);
    }
    if (fo != stdout) fclose(fo);
    const double w5 = now_s();
    fourmc_ctx_destroy(ctx);
    if (in_mapped) munmap(in, n); else free(in);
    if (timing)
        fprintf(stderr, "timing: open+map input %.3f s, context %.3f s, output buffer %.3f s, codec call %.3f s, "
                        "write+close %.3f s, teardown %.3f s (input %s, output %s)\n",
                w1 - w0, w2 - w1, w3 - w2, w4 - w3, w5 - w4, now_s() - w5, in_mapped ? "mapped" : "read", out_mapped ? "mapped" : "memory");
    free(dyn);
    return 0;
}
