/*
 * ref_harness.c -- thin driver that links the UNMODIFIED reference objects (compiled in place
 * from /root/reference/src/core/*.c) and the reference's prebuilt lib/libz.a (zlib 1.2.8).
 * TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile into oracle/_ref/ref_harness; nothing of
 * the reference's source is copied into this repo -- only the prototypes of the exported but
 * header-less helpers are declared here.
 *
 *   ref_harness split  IN BITS OUTPREFIX      chunk loop of run_compress (workers.c:779-855) without the
 *                                             codec: writes OUTPREFIX.p0..p3 (planes) and OUTPREFIX.masked
 *   ref_harness merge  P0 P1 P2 P3 OUT        merge_byte_to_float_stream (workers.c:423)
 *   ref_harness inflate PAYLOAD N OUT         mzlib_inf (zip.c:262) on one COMPRESSED payload that inflates to N bytes
 *   ref_harness deflate PLANE OUT             mzlib_def (zip.c:164) per CHUNK_SIZE piece; writes hdr+payload records
 *   ref_harness version                       zlib version the reference binary carries
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.h"
#include "constant.h"
#include "mrczip.h"

/* exported by workers.c, declared in no reference header */
void apply_mask(float *buffer, int startIndex, int num, const int bitsToErase, int isFirstChk);
void split_float_to_byte_stream(float *buf, int num, char *zins[], int bitsToMask, int isFirstChk);
void merge_byte_to_float_stream(float *buf, int num, char *zouts[]);

static FILE *xopen(const char *p, const char *m)
{
    FILE *f = fopen(p, m);
    if (!f) { fprintf(stderr, "ref_harness: cannot open %s\n", p); exit(2); }
    return f;
}

static int cmd_split(const char *in, int bits, const char *prefix)
{
    FILE *fin = xopen(in, "rb");
    char name[1024];
    FILE *fp[4], *fm;
    for (int j = 0; j < 4; j++) { snprintf(name, sizeof name, "%s.p%d", prefix, j); fp[j] = xopen(name, "wb"); }
    snprintf(name, sizeof name, "%s.masked", prefix);
    fm = xopen(name, "wb");
    float *buf = malloc(sizeof(float) * CHUNK_SIZE);
    char *zins[4];
    for (int j = 0; j < 4; j++) zins[j] = malloc(CHUNK_SIZE);
    int isFirstChk = 1;
    int num = fread(buf, sizeof(float), CHUNK_SIZE, fin);
    while (num > 0) {
        split_float_to_byte_stream(buf, num, zins, bits, isFirstChk);
        isFirstChk = 0;
        fwrite(buf, sizeof(float), num, fm);
        for (int j = 0; j < 4; j++) fwrite(zins[j], 1, num, fp[j]);
        num = fread(buf, sizeof(float), CHUNK_SIZE, fin);
    }
    for (int j = 0; j < 4; j++) { fclose(fp[j]); free(zins[j]); }
    fclose(fm); fclose(fin); free(buf);
    return 0;
}

static long fsize(FILE *f) { fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET); return n; }

static int cmd_merge(char **p, const char *out)
{
    FILE *f[4]; char *z[4]; long n = 0;
    for (int j = 0; j < 4; j++) {
        f[j] = xopen(p[j], "rb"); n = fsize(f[j]);
        z[j] = malloc(n ? n : 1);
        if (fread(z[j], 1, n, f[j]) != (size_t)n) return 3;
        fclose(f[j]);
    }
    float *buf = malloc(n ? 4 * n : 4);
    merge_byte_to_float_stream(buf, (int)n, z);
    FILE *fo = xopen(out, "wb");
    fwrite(buf, 4, n, fo);
    fclose(fo);
    return 0;
}

static int cmd_inflate(const char *payload, int n, const char *out)
{
    FILE *f = xopen(payload, "rb");
    long len = fsize(f);
    mzip_t z;
    uint32_t chk = (uint32_t)(n > len ? n : len);
    if (init_mrc_zip_stream(&z, chk, ZLIB_INF, 0) != 0) return 3;
    if (fread(z.in, 1, len, f) != (size_t)len) return 3;
    fclose(f);
    z.inlen = (uint32_t)len;
    memset(z.out, 0xA5, chk);
    char *p = NULL;
    z.unzipfun(&z, COMPRESSED, n, &p);
    z_stream *s = (z_stream *)z.zipper;
    long produced = (long)chk - (long)s->avail_out;
    FILE *fo = xopen(out, "wb");
    fwrite(p, 1, produced, fo);
    fclose(fo);
    printf("inflated %ld of %d bytes, consumed %ld of %ld\n", produced, n, len - (long)s->avail_in, len);
    mzip_term(&z);
    return produced == n ? 0 : 4;
}

static int cmd_deflate(const char *plane, const char *out)
{
    FILE *f = xopen(plane, "rb");
    FILE *fo = xopen(out, "wb");
    mzip_t z;
    if (init_mrc_zip_stream(&z, CHUNK_SIZE, ZLIB_DEF, ZIP_FAST) != 0) return 3;
    int num;
    while ((num = fread(z.zin, 1, CHUNK_SIZE, f)) > 0) {
        char *p; int len;
        z.inlen = num;
        z.zipfun(&z, &p, &len);
        fwrite(p, 1, len, fo);
    }
    fclose(f); fclose(fo);
    mzip_term(&z);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc >= 2 && !strcmp(argv[1], "version")) { printf("%s\n", zlib_version); return 0; }
    if (argc == 5 && !strcmp(argv[1], "split")) return cmd_split(argv[2], atoi(argv[3]), argv[4]);
    if (argc == 7 && !strcmp(argv[1], "merge")) return cmd_merge(argv + 2, argv[6]);
    if (argc == 5 && !strcmp(argv[1], "inflate")) return cmd_inflate(argv[2], atoi(argv[3]), argv[4]);
    if (argc == 4 && !strcmp(argv[1], "deflate")) return cmd_deflate(argv[2], argv[3]);
    fprintf(stderr, "usage: see the header of oracle/ref_harness.c\n");
    return 1;
}
