/*
 * fourmc_cli.c -- `4mc` command line tool over lib4mcgpu.so.
 *
 * Flag-, message-level- and exit-code-compatible with the reference CLI (native/4mccli.c:170-361,
 * native/4mc.c:135-161,220-386,896-934): -1..-4 level, -d decode, -t test, -c stdout, -f overwrite,
 * -v / -q verbosity, -V version, -h help, -z zstd (4mz),
 * "stdin" / "stdout" / "null" file names, automatic .4mc output names when stdout is a terminal.
 * Exit codes: 1 generic, 2 input, 3 output, 4 content.  This file is the argument parser only: the work is
 * done by the four entry points the reference's own CLI calls (native/4mc.h:36-41), which lib4mcgpu.so
 * exports and which stream files of any size through the GPU (4mc_b200/csrc/fileio.h).
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
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

    /* native/4mccli.c:342-356: the library entry points do the work (and exit(1..4) on a fatal error) */
    if (decode) {
        SAY(3, zstd ? "Compression: ZSTD\n" : "Compression: LZ4\n");
        if (zstd) fourMZDecompressFileName(display, overwrite, (char *)in_name, (char *)out_name);
        else fourMcDecompressFileName(display, overwrite, (char *)in_name, (char *)out_name);
    } else {
        SAY(2, zstd ? "Compression: ZSTD\n" : "Compression: LZ4\n");                    /* 4mc.c:241, :410 */
        if (zstd) fourMZcompressFilename(display, overwrite, (char *)in_name, (char *)out_name, level);
        else fourMCcompressFilename(display, overwrite, (char *)in_name, (char *)out_name, level);
    }
    free(dyn);
    return 0;
}
