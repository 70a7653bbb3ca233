/*
 * mrczip_host.c -- the host side of the drop-in, plain C like the reference's: file and MRC I/O,
 * chunking, the container reader / writer and accounting.  It replaces the reference's serial chunk
 * loop (src/core/workers.c:690-881, 568-688), its per-file adapter (src/core/adapt.c:28-90) and the
 * container primitives (src/core/common.c:26-149, src/core/zip.c:381-399), and calls CUDA only
 * through the mzb_* C ABI (fz_api.cu).  Nothing here compresses or inflates on the CPU.
 */
#define _FILE_OFFSET_BITS 64
#define _GNU_SOURCE
#include <fcntl.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <sys/types.h>
#include <unistd.h>

#include "../../include/mrczip_b200.h"

int isTestThroughput = 0; /* reference workers.c:39 */

/* chunks handed to the GPU per call from the FILE* API (bounds pinned + device memory per thread) */
#define HOST_BATCH_CHUNKS 16
/* file I/O of the FILE* API: batches are read / written by IO_THREADS threads with pread / pwrite (a single
 * thread copies page cache <-> pinned memory at a few GB/s, an order of magnitude below what the GPU path takes),
 * and the read of batch b+1 and the write of batch b-1 overlap the GPU work on batch b (two buffers each way).
 * Streams that cannot seek (pipes) take the plain fread / fwrite loop. */
#define IO_THREADS 16 /* upper bound; MRCZIP_IO_THREADS (default 4) says how many are used */
#define IO_SLICE_MIN ((size_t)8 << 20)

/* ------------------------------------------------------------------ common.c equivalents */

uint64_t get_file_size(FILE *fp)
{
    /* common.c:26-39 */
    if (fp == NULL) return (uint64_t)-1;
    long cur = ftell(fp);
    fseek(fp, 0L, SEEK_END);
    uint64_t sz = (uint64_t)ftell(fp);
    fseek(fp, cur, SEEK_SET);
    return sz;
}

double now_sec(void)
{
    /* common.c:41-46 */
    struct timeval tm;
    gettimeofday(&tm, NULL);
    return (double)tm.tv_sec + (double)tm.tv_usec / 1000000.0;
}

void init_context(ctx_t *ctx)
{
    ctx->fileCount = 0;
    ctx->allFileSize = 0;
    ctx->allZipFileSize = 0;
    ctx->zipTime = 0.0;
    ctx->unzipTime = 0.0;
}

void reset_context(ctx_t *ctx) { init_context(ctx); }

void update_context(ctx_t *dst, ctx_t *src)
{
    /* common.c:93-100 */
    dst->fileCount += src->fileCount;
    dst->allFileSize += src->allFileSize;
    dst->allZipFileSize += src->allZipFileSize;
    dst->zipTime += src->zipTime;
    dst->unzipTime += src->unzipTime;
}

void print_context_info(ctx_t *ctx, const char *hintMsg)
{
    /* common.c:66-90: same columns; speed in MB/s with MB = 2^20 bytes */
    const double t = ctx->zipTime > 0.001 ? ctx->zipTime : ctx->unzipTime;
    const char *op = ctx->zipTime > 0.001 ? "(zip)" : "(unzip)";
    printf("-------------------%s--------------------\n", hintMsg);
    printf("[Original File Size(Bytes)]    [Compressed File Size(Bytes)]    [Zip/Unzip Time(s)]    [Speed(MB/s)]    \n");
    printf("%-31lu%-33lu%.4f%-17s%.4f\n", (unsigned long)ctx->allFileSize, (unsigned long)ctx->allZipFileSize, t, op,
           t > 0 ? (double)ctx->allFileSize / (t * 1024.0 * 1024.0) : 0.0);
}

void init_mrczip_header(mrczip_header_t *hd, char type)
{
    /* common.c:102-108 */
    hd->type = type;
    hd->fsz = 0;
    hd->chk = 0;
    memset(hd->ztypes, 0, MZB_PLANES);
}

void print_mrczip_header(mrczip_header_t *hd, const char *hintMsg)
{
    printf("[%s]: Original file size = %lu, chunk size = %u, compresstion type = %d\n", hintMsg,
           (unsigned long)hd->fsz, hd->chk, hd->type);
}

int read_mrczip_header(FILE *fin, mrczip_header_t *hd)
{
    /* common.c:117-135: field-wise, native little-endian, 17 bytes */
    if (fread(&hd->fsz, sizeof(uint64_t), 1, fin) < 1) {
        fprintf(stderr, "[ERROR]:Failed to read file\n");
        return -1;
    }
    if (fread(&hd->chk, sizeof(uint32_t), 1, fin) < 1) return -1;
    if (fread(&hd->type, 1, 1, fin) < 1) return -1;
    for (int i = 0; i < MZB_PLANES; i++)
        if (fread(&hd->ztypes[i], 1, 1, fin) < 1) return -1;
    return 0;
}

int write_mrczip_header(FILE *fout, mrczip_header_t *hd)
{
    /* common.c:137-149 */
    fwrite(&hd->fsz, sizeof(uint64_t), 1, fout);
    fwrite(&hd->chk, sizeof(uint32_t), 1, fout);
    fwrite(&hd->type, 1, 1, fout);
    for (int i = 0; i < MZB_PLANES; i++) fwrite(&hd->ztypes[i], 1, 1, fout);
    return 0;
}

/* ------------------------------------------------------------------ zip.c:381-399 */

void pack_header(char *nbuf, btype_t btype, uint32_t len)
{
    unsigned char *buf = (unsigned char *)nbuf;
    buf[0] = len & 0xFF;
    buf[1] = (len >> 8) & 0xFF;
    buf[2] = (len >> 16) & 0xFF;
    buf[3] = (unsigned char)(((len >> 24) & 0x7F) | ((unsigned)btype << 7));
}

void unpack_header(const char *nbuf, btype_t *btype, uint32_t *len)
{
    const unsigned char *buf = (const unsigned char *)nbuf;
    *btype = (btype_t)((buf[3] & 0x80) >> 7);
    *len = (uint32_t)buf[0] | ((uint32_t)buf[1] << 8) | ((uint32_t)buf[2] << 16) | ((uint32_t)(buf[3] & 0x7f) << 24);
}

/* ------------------------------------------------------------------ per-thread GPU context */

static pthread_key_t g_key;
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

typedef struct {
    mzb_ctx *ctx;
    void *pin_in;        /* [0] of the input double buffer (the only one the fread / fwrite loop uses) */
    size_t pin_in_cap;
    void *pin_out;
    size_t pin_out_cap;
    void *pin_in2;       /* [1]: allocated on the first overlapped call */
    size_t pin_in2_cap;
    void *pin_out2;
    size_t pin_out2_cap;
} thread_state_t;

static void thread_state_free(void *p)
{
    thread_state_t *ts = (thread_state_t *)p;
    if (!ts) return;
    mzb_host_free(ts->pin_in);
    mzb_host_free(ts->pin_out);
    mzb_host_free(ts->pin_in2);
    mzb_host_free(ts->pin_out2);
    mzb_destroy(ts->ctx);
    free(ts);
}

static void make_key(void) { pthread_key_create(&g_key, thread_state_free); }

static thread_state_t *thread_state(void)
{
    pthread_once(&g_once, make_key);
    thread_state_t *ts = (thread_state_t *)pthread_getspecific(g_key);
    if (ts) return ts;
    ts = (thread_state_t *)calloc(1, sizeof(*ts));
    if (!ts) return NULL;
    const char *dev = getenv("MRCZIP_DEVICE");
    if (mzb_create(&ts->ctx, dev ? atoi(dev) : 0, NULL) != MZB_OK) {
        fprintf(stderr, "[%s:%d] ERROR: no usable CUDA device (this build has no CPU fallback)\n", __FILE__, __LINE__);
        free(ts);
        return NULL;
    }
    mzb_set_batch_chunks(ts->ctx, HOST_BATCH_CHUNKS);
    pthread_setspecific(g_key, ts);
    return ts;
}

static int pin_reserve(void **p, size_t *cap, size_t need)
{
    if (*cap >= need) return 0;
    mzb_host_free(*p);
    *p = mzb_host_alloc(need);
    *cap = *p ? need : 0;
    return *p ? 0 : -1;
}

/* per-plane accounting of a run of chunk records, for the print_result-style summary (zip.c:401-466) */
typedef struct {
    uint64_t fsz[MZB_PLANES], zfsz[MZB_PLANES];
} plane_acct_t;

static void account_records(const unsigned char *rec, size_t n, uint32_t chk, uint64_t words, plane_acct_t *a)
{
    size_t off = 0;
    for (uint64_t w0 = 0; w0 < words && off + 16 <= n; w0 += chk) {
        const uint32_t num = (uint32_t)((words - w0) < chk ? (words - w0) : chk);
        const unsigned char *h = rec + off;
        off += 16;
        for (int j = 0; j < MZB_PLANES; j++) {
            btype_t bt;
            uint32_t len;
            unpack_header((const char *)h + 4 * j, &bt, &len);
            a->fsz[j] += num;
            a->zfsz[j] += (uint64_t)len + 4; /* zip.c:180,189: the reference counts the 4-byte header */
            off += len;
        }
    }
}

static void print_result_like(const plane_acct_t *a, double seconds, int is_zip, const char *hintMsg)
{
    /* zip.c:401-466: ratio = compressed / original; throughput in MB/s, MB = 2^20 */
    uint64_t fsz = 0, zfsz = 0;
    printf("-------------------%s Information--------------\n", hintMsg);
    printf("[ByteStreamIndex]   [Before Compress(Bytes)]   [After Compress(Bytes)]   [Compress Ratio]   \n");
    for (int j = 0; j < MZB_PLANES; j++) {
        fsz += a->fsz[j];
        zfsz += a->zfsz[j];
        printf("%-20d%-27lu%-26lu%-19.4f\n", j, (unsigned long)a->fsz[j], (unsigned long)a->zfsz[j],
               a->fsz[j] ? (double)a->zfsz[j] / (double)a->fsz[j] : 0.0);
    }
    printf("%-20s%-27lu%-26lu%-19.4f\n", "Whole File", (unsigned long)fsz, (unsigned long)zfsz,
           fsz ? (double)zfsz / (double)fsz : 0.0);
    if (seconds > 0) {
        printf("---------------------------------------\n");
        printf("%s Throughput: %f MB/s (GPU path, host I/O staging included)\n", is_zip ? "Compression" : "Decompression",
               (double)fsz / (1024.0 * 1024.0 * seconds));
        printf("---------------------------------------\n");
    }
}

/* ------------------------------------------------------------------ overlapped, multi-threaded file I/O */

typedef struct {
    int fd, wr;
    unsigned char *buf;
    size_t n, done;
    off_t off;
} io_slice_t;

static void *io_slice_run(void *p)
{
    io_slice_t *s = (io_slice_t *)p;
    while (s->done < s->n) {
        const ssize_t r = s->wr ? pwrite(s->fd, s->buf + s->done, s->n - s->done, s->off + (off_t)s->done)
                                : pread(s->fd, s->buf + s->done, s->n - s->done, s->off + (off_t)s->done);
        if (r <= 0) break; /* end of file (read) or an error: the caller sees a short count */
        s->done += (size_t)r;
    }
    return NULL;
}

static int g_io_threads = 0; /* 0 = not set: MRCZIP_IO_THREADS, else 4 */
static pthread_once_t g_io_once = PTHREAD_ONCE_INIT;
static int g_io_env = 4;
static void io_threads_init(void)
{
    const char *e = getenv("MRCZIP_IO_THREADS");
    const int w = e ? atoi(e) : 4;
    g_io_env = w < 1 ? 1 : (w > IO_THREADS ? IO_THREADS : w);
}
static int io_threads(void)
{
    const int set = __atomic_load_n(&g_io_threads, __ATOMIC_RELAXED);
    if (set > 0) return set;
    pthread_once(&g_io_once, io_threads_init);
    return g_io_env;
}

int mzb_set_io_threads(int n)
{
    if (n < 0 || n > IO_THREADS) return MZB_E_ARG;
    __atomic_store_n(&g_io_threads, n, __ATOMIC_RELAXED);
    return MZB_OK;
}

/* n bytes at file offset off, split over up to IO_THREADS threads; returns the contiguous byte count done */
static size_t io_parallel(int fd, int wr, void *buf, size_t n, off_t off)
{
    io_slice_t sl[IO_THREADS];
    pthread_t th[IO_THREADS];
    const int want = io_threads();
    int k = (int)(n / IO_SLICE_MIN);
    if (k < 1) k = 1;
    if (k > want) k = want;
    const size_t per = ((n + (size_t)k - 1) / (size_t)k + 4095) & ~(size_t)4095;
    int started = 0;
    for (int i = 0; i < k; i++) {
        const size_t b = (size_t)i * per;
        sl[i].fd = fd; sl[i].wr = wr; sl[i].done = 0;
        sl[i].buf = (unsigned char *)buf + (b < n ? b : n);
        sl[i].n = b < n ? (n - b < per ? n - b : per) : 0;
        sl[i].off = off + (off_t)b;
        if (i + 1 < k && pthread_create(&th[i], NULL, io_slice_run, &sl[i]) == 0) started |= 1 << i;
        else io_slice_run(&sl[i]);
    }
    size_t total = 0;
    int whole = 1;
    for (int i = 0; i < k; i++) {
        if (started & (1 << i)) pthread_join(th[i], NULL);
        if (whole) total += sl[i].done;
        if (sl[i].done < sl[i].n) whole = 0;
    }
    return total;
}

/* one background transfer (a batch read ahead, or a batch written behind) */
typedef struct {
    pthread_t th;
    int active;
    int fd, wr;
    void *buf;
    size_t n, done;
    off_t off;
} io_job_t;

static void *io_job_run(void *p)
{
    io_job_t *j = (io_job_t *)p;
    j->done = io_parallel(j->fd, j->wr, j->buf, j->n, j->off);
    return NULL;
}

static void io_job_start(io_job_t *j, int fd, int wr, void *buf, size_t n, off_t off)
{
    j->fd = fd; j->wr = wr; j->buf = buf; j->n = n; j->off = off; j->done = 0;
    j->active = pthread_create(&j->th, NULL, io_job_run, j) == 0;
    if (!j->active) io_job_run(j);
}

static size_t io_job_wait(io_job_t *j)
{
    if (j->active) { pthread_join(j->th, NULL); j->active = 0; }
    return j->done;
}

/* a regular file we may address by offset: its descriptor and current position; -1 otherwise */
static int io_seekable(FILE *fp, int for_write, off_t *pos)
{
    if (!fp || getenv("MRCZIP_SERIAL_IO")) return -1;
    const int fd = fileno(fp);
    struct stat st;
    if (fd < 0 || fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) return -1;
    if (for_write) {
        /* pwrite ignores its offset on a descriptor opened with O_APPEND: such a FILE* takes the plain fwrite loop */
        const int fl = fcntl(fd, F_GETFL);
        if (fl < 0 || (fl & O_APPEND)) return -1;
        if (fflush(fp) != 0) return -1;
    }
    *pos = ftello(fp);
    return *pos < 0 ? -1 : fd;
}

/* pinned staging, sized by what the file needs (a 64 MiB stack does not pin 1.5 GiB); the second pair only when
 * there is more than one batch to overlap */
static int pin_pair(thread_state_t *ts, size_t in_need, size_t out_need, int two)
{
    if (pin_reserve(&ts->pin_in, &ts->pin_in_cap, in_need) || pin_reserve(&ts->pin_out, &ts->pin_out_cap, out_need)) return -1;
    if (!two) return 0;
    return pin_reserve(&ts->pin_in2, &ts->pin_in2_cap, in_need) || pin_reserve(&ts->pin_out2, &ts->pin_out2_cap, out_need);
}

/* run_compress over seekable files: read b+1 | GPU b | write b-1 */
static int compress_overlapped(thread_state_t *ts, FILE *fin, int fdin, off_t pos_in, FILE *fout, mrczip_header_t *hd,
                               int bitsToMask, plane_acct_t *acct, uint64_t *zbytes)
{
    const uint32_t chk = hd->chk;
    const size_t batch_words = (size_t)HOST_BATCH_CHUNKS * chk;
    const uint64_t avail = hd->fsz > (uint64_t)pos_in ? hd->fsz - (uint64_t)pos_in : 0;
    const uint64_t words_total = avail / 4; /* a ragged tail of 1..3 bytes is dropped (workers.c:744) */
    if (words_total == 0) return MZB_OK;     /* workers.c:757-764: nothing read, nothing written */
    const uint64_t nb = (words_total + batch_words - 1) / batch_words;
    const size_t max_words = nb > 1 ? batch_words : (size_t)words_total;
    const size_t out_cap = mzb_compress_bound(max_words, chk);
    if (pin_pair(ts, max_words * 4, out_cap, nb > 1)) {
        fprintf(stderr, "[%s:%d] ERROR: fail to alloc mem\n", __FILE__, __LINE__);
        return MZB_E_NOMEM;
    }
    int fdout = -1;
    off_t pos_out = 0;
    if (isTestThroughput != 1) {
        write_mrczip_header(fout, hd);
        fdout = io_seekable(fout, 1, &pos_out);
        if (fdout < 0) return MZB_E_IO;
    }
    void *in[2] = {ts->pin_in, ts->pin_in2}, *out[2] = {ts->pin_out, ts->pin_out2};
    io_job_t jr, jw;
    memset(&jr, 0, sizeof(jr));
    memset(&jw, 0, sizeof(jw));
    int rc = MZB_OK, writing = 0;
    {
        const uint64_t w = words_total < batch_words ? words_total : batch_words;
        io_job_start(&jr, fdin, 0, in[0], (size_t)w * 4, pos_in);
    }
    for (uint64_t b = 0; b < nb && rc == MZB_OK; b++) {
        const uint64_t w0 = b * batch_words;
        const uint64_t want = (words_total - w0) < batch_words ? (words_total - w0) : batch_words;
        const uint64_t num = io_job_wait(&jr) / 4;
        if (num < want) { rc = MZB_E_IO; break; }   /* the file shrank under us */
        if (b + 1 < nb) {
            const uint64_t w1 = w0 + batch_words;
            const uint64_t nxt = (words_total - w1) < batch_words ? (words_total - w1) : batch_words;
            io_job_start(&jr, fdin, 0, in[(b + 1) & 1], (size_t)nxt * 4, pos_in + (off_t)(w1 * 4));
        }
        uint64_t sz = 0;
        rc = mzb_compress_host(ts->ctx, in[b & 1], num, bitsToMask, b == 0 ? MZB_MRC_HEADER_WORDS : 0, chk, hd->fsz, 0,
                               out[b & 1], out_cap, &sz);
        if (rc != MZB_OK) {
            fprintf(stderr, "[%s:%d] ERROR: GPU compress failed: %s\n", __FILE__, __LINE__, mzb_strerror(rc));
            break;
        }
        account_records((const unsigned char *)out[b & 1], sz, chk, num, acct);
        *zbytes += sz;
        if (writing && io_job_wait(&jw) != jw.n) { rc = MZB_E_IO; break; }
        writing = 0;
        if (isTestThroughput != 1) {
            io_job_start(&jw, fdout, 1, out[b & 1], (size_t)sz, pos_out);
            writing = 1;
            pos_out += (off_t)sz;
        }
    }
    io_job_wait(&jr);
    if (writing && io_job_wait(&jw) != jw.n && rc == MZB_OK) rc = MZB_E_IO;
    fseeko(fin, 0, SEEK_END);                       /* where the reference's fread loop leaves it */
    if (fdout >= 0) fseeko(fout, pos_out, SEEK_SET);
    return rc;
}

/* extent of the next batch of a container: up to bchunks chunk records starting at file offset off */
/* A plane header of a chunk of num words: RAW payloads are the num plane bytes (zip.c:186-190), COMPRESSED ones are
 * only written when they save more than their 4-byte header (zip.c:177: inlen > len + 4).  Anything else cannot come
 * from the reference or from this library, and would overrun the staging buffers sized for chk bytes per plane. */
static int plane_header_ok(btype_t bt, uint32_t len, uint64_t num)
{
    return bt == RAW ? (uint64_t)len == num : (uint64_t)len + 4u < num;
}

static int walk_batch(int fd, off_t off, uint64_t bchunks, uint64_t words_left, uint32_t chk, size_t cap, size_t *bytes, uint64_t *bw)
{
    size_t fill = 0;
    uint64_t w = 0;
    for (uint64_t c = 0; c < bchunks && w < words_left; c++) {
        unsigned char h[16];
        if (pread(fd, h, 16, off + (off_t)fill) != 16) return MZB_E_FORMAT;
        const uint64_t num = (words_left - w) < chk ? (words_left - w) : chk;
        size_t payload = 0;
        for (int j = 0; j < MZB_PLANES; j++) {
            btype_t bt;
            uint32_t len;
            unpack_header((const char *)h + 4 * j, &bt, &len);
            if (!plane_header_ok(bt, len, num)) return MZB_E_FORMAT;
            payload += len;
        }
        if (fill + 16 + payload > cap) return MZB_E_FORMAT;
        fill += 16 + payload;
        w += num;
    }
    *bytes = fill;
    *bw = w;
    return MZB_OK;
}

/* run_uncompress over seekable files: read b+1 | GPU b | write b-1 */
static int uncompress_overlapped(thread_state_t *ts, FILE *fin, int fdin, off_t pos_in, FILE *fout, uint32_t chk, uint64_t words,
                                 uint64_t bchunks, plane_acct_t *acct, uint64_t *zbytes)
{
    const uint64_t nchunks = (words + chk - 1) / chk;
    if (bchunks > nchunks && nchunks > 0) bchunks = nchunks;
    const size_t in_cap = (size_t)(bchunks * (16 + 4ull * (chk + 4ull))) + 64, out_cap = (size_t)bchunks * chk * 4 + 64;
    if (pin_pair(ts, in_cap, out_cap, nchunks > bchunks)) {
        fprintf(stderr, "[%s:%d] ERROR: fail to alloc mem\n", __FILE__, __LINE__);
        return MZB_E_NOMEM;
    }
    int fdout = -1;
    off_t pos_out = 0;
    if (isTestThroughput != 1) {
        fdout = io_seekable(fout, 1, &pos_out);
        if (fdout < 0) return MZB_E_IO;
    }
    void *in[2] = {ts->pin_in, ts->pin_in2}, *out[2] = {ts->pin_out, ts->pin_out2};
    io_job_t jr, jw;
    memset(&jr, 0, sizeof(jr));
    memset(&jw, 0, sizeof(jw));
    int rc = MZB_OK, writing = 0;
    size_t bytes = 0, nbytes = 0;
    uint64_t bw = 0, nbw = 0;
    off_t off = pos_in;
    if (words > 0 && (rc = walk_batch(fdin, off, bchunks, words, chk, in_cap, &bytes, &bw)) == MZB_OK)
        io_job_start(&jr, fdin, 0, in[0], bytes, off);
    uint64_t b = 0;
    for (uint64_t w0 = 0; w0 < words && rc == MZB_OK; b++) {
        if (io_job_wait(&jr) != bytes) { rc = MZB_E_FORMAT; break; }   /* truncated container */
        off += (off_t)bytes;
        if (w0 + bw < words) {
            if ((rc = walk_batch(fdin, off, bchunks, words - w0 - bw, chk, in_cap, &nbytes, &nbw)) != MZB_OK) break;
            io_job_start(&jr, fdin, 0, in[(b + 1) & 1], nbytes, off);
        }
        uint64_t nw = 0;
        rc = mzb_decompress_host(ts->ctx, in[b & 1], bytes, 0, chk, bw, out[b & 1], out_cap / 4, &nw);
        if (rc != MZB_OK) break;
        account_records((const unsigned char *)in[b & 1], bytes, chk, bw, acct);
        *zbytes += bytes;
        if (writing && io_job_wait(&jw) != jw.n) { rc = MZB_E_IO; break; }
        writing = 0;
        if (isTestThroughput != 1) {
            io_job_start(&jw, fdout, 1, out[b & 1], (size_t)nw * 4, pos_out);
            writing = 1;
            pos_out += (off_t)(nw * 4);
        }
        w0 += bw;
        bytes = nbytes;
        bw = nbw;
    }
    io_job_wait(&jr);
    if (writing && io_job_wait(&jw) != jw.n && rc == MZB_OK) rc = MZB_E_IO;
    fseeko(fin, off, SEEK_SET);
    if (fdout >= 0) fseeko(fout, pos_out, SEEK_SET);
    return rc;
}

/* ------------------------------------------------------------------ run_compress / run_uncompress */

int run_compress(FILE *fin, ctx_t *ctx, FILE *fout, const int bitsToMask, const char *dataConvertedType)
{
    if (!fin || !ctx || (!fout && isTestThroughput != 1)) return MZB_E_ARG;
    if (bitsToMask < 0 || bitsToMask > 32) {
        fprintf(stderr, "[%s:%d] ERROR: bits to erase must be in 0..32 (got %d)\n", __FILE__, __LINE__, bitsToMask);
        return MZB_E_ARG;
    }
    if (dataConvertedType && strcmp(dataConvertedType, "float") != 0) {
        fprintf(stderr, "[%s:%d] ERROR: only the \"float\" path is implemented (got \"%s\")\n", __FILE__, __LINE__, dataConvertedType);
        return MZB_E_ARG;
    }
    thread_state_t *ts = thread_state();
    if (!ts) return MZB_E_CUDA;
    const double begin = now_sec();
    const uint32_t chk = MZB_CHUNK_WORDS;
    const size_t batch_words = (size_t)HOST_BATCH_CHUNKS * chk;
    {   /* regular files: multi-threaded, overlapped I/O (same bytes, same accounting) */
        off_t pos_in = 0, pos_chk = 0;
        const int fdin = io_seekable(fin, 0, &pos_in);
        if (fdin >= 0 && (isTestThroughput == 1 || io_seekable(fout, 1, &pos_chk) >= 0)) {
            mrczip_header_t hd2;
            init_mrczip_header(&hd2, 0);
            hd2.chk = chk;
            hd2.fsz = get_file_size(fin); /* workers.c:743 */
            plane_acct_t acct2;
            memset(&acct2, 0, sizeof(acct2));
            uint64_t zb = 0;
            const int rc2 = compress_overlapped(ts, fin, fdin, pos_in, fout, &hd2, bitsToMask, &acct2, &zb);
            const double dt2 = now_sec() - begin;
            ctx->zipTime += dt2;
            ctx->allZipFileSize += zb; /* workers.c:869-872 */
            print_result_like(&acct2, dt2, 1, "Compression Summary Result");
            return rc2;
        }
    }
    if (pin_reserve(&ts->pin_in, &ts->pin_in_cap, batch_words * 4) ||
        pin_reserve(&ts->pin_out, &ts->pin_out_cap, mzb_compress_bound(batch_words, chk))) {
        fprintf(stderr, "[%s:%d] ERROR: fail to alloc mem\n", __FILE__, __LINE__);
        return MZB_E_NOMEM;
    }
    mrczip_header_t hd;
    init_mrczip_header(&hd, 0);
    hd.chk = chk;
    hd.fsz = get_file_size(fin); /* workers.c:743 */
    plane_acct_t acct;
    memset(&acct, 0, sizeof(acct));
    int first = 1, rc = MZB_OK;
    uint64_t zbytes = 0;
    /* fread of 4-byte items: a ragged tail of 1..3 bytes is dropped exactly like workers.c:744,854 */
    size_t num = fread(ts->pin_in, sizeof(uint32_t), batch_words, fin);
    if (num > 0 && isTestThroughput != 1) write_mrczip_header(fout, &hd); /* workers.c:757-764 */
    while (num > 0) {
        uint64_t sz = 0;
        rc = mzb_compress_host(ts->ctx, ts->pin_in, num, bitsToMask, first ? MZB_MRC_HEADER_WORDS : 0, chk, hd.fsz, 0,
                               ts->pin_out, ts->pin_out_cap, &sz);
        if (rc != MZB_OK) {
            fprintf(stderr, "[%s:%d] ERROR: GPU compress failed: %s\n", __FILE__, __LINE__, mzb_strerror(rc));
            break;
        }
        first = 0;
        account_records((const unsigned char *)ts->pin_out, sz, chk, num, &acct);
        zbytes += sz;
        if (isTestThroughput != 1 && fwrite(ts->pin_out, 1, sz, fout) != sz) { rc = MZB_E_IO; break; }
        num = fread(ts->pin_in, sizeof(uint32_t), batch_words, fin);
    }
    const double dt = now_sec() - begin;
    ctx->zipTime += dt;
    ctx->allZipFileSize += zbytes; /* workers.c:869-872: sum of zfsz == container bytes - 17 */
    print_result_like(&acct, dt, 1, "Compression Summary Result");
    return rc;
}

int run_uncompress(FILE *fin, ctx_t *ctx, mrczip_header_t *hd, FILE *fout, const char *dataConvertedType)
{
    if (!fin || !ctx || !hd || (!fout && isTestThroughput != 1)) return MZB_E_ARG;
    if (dataConvertedType && strcmp(dataConvertedType, "float") != 0) {
        fprintf(stderr, "[%s:%d] ERROR: only the \"float\" path is implemented (got \"%s\")\n", __FILE__, __LINE__, dataConvertedType);
        return MZB_E_ARG;
    }
    if (hd->chk == 0 || hd->chk >= 0x80000000u) {
        fprintf(stderr, "too large chunk size\n"); /* zip.c:325-329 */
        return MZB_E_FORMAT;
    }
    for (int j = 0; j < MZB_PLANES; j++)
        if (hd->ztypes[j] != 0) {
            fprintf(stderr, "[%s:%d] ERROR: unsupported ztype %d (only zlib streams are ever written)\n", __FILE__, __LINE__, hd->ztypes[j]);
            return MZB_E_FORMAT;
        }
    thread_state_t *ts = thread_state();
    if (!ts) return MZB_E_CUDA;
    const double begin = now_sec();
    const uint32_t chk = hd->chk;
    const uint64_t words = hd->fsz / MZB_PLANES; /* workers.c:577 */
    /* batch: as many whole chunks as fit HOST_BATCH_CHUNKS reference-sized chunks */
    uint64_t bchunks = ((uint64_t)HOST_BATCH_CHUNKS * MZB_CHUNK_WORDS) / chk;
    if (bchunks == 0) bchunks = 1;
    if (bchunks > 4096) bchunks = 4096;
    {   /* regular files: multi-threaded, overlapped I/O */
        off_t pos_in = 0, pos_chk = 0;
        const int fdin = io_seekable(fin, 0, &pos_in);
        if (fdin >= 0 && (isTestThroughput == 1 || io_seekable(fout, 1, &pos_chk) >= 0)) {
            plane_acct_t acct2;
            memset(&acct2, 0, sizeof(acct2));
            uint64_t zb = 0;
            const int rc2 = uncompress_overlapped(ts, fin, fdin, pos_in, fout, chk, words, bchunks, &acct2, &zb);
            if (rc2 != MZB_OK) fprintf(stderr, "[%s:%d] ERROR: GPU decompress failed: %s\n", __FILE__, __LINE__, mzb_strerror(rc2));
            const double dt2 = now_sec() - begin;
            ctx->allFileSize += words * MZB_PLANES; /* workers.c:679-684 */
            ctx->allZipFileSize += zb;
            ctx->unzipTime += dt2;
            print_result_like(&acct2, dt2, 0, "Decompress Result Info");
            return rc2;
        }
    }
    const size_t in_cap = (size_t)(bchunks * (16 + 4ull * (chk + 4ull))) + 64;
    if (pin_reserve(&ts->pin_in, &ts->pin_in_cap, in_cap) || pin_reserve(&ts->pin_out, &ts->pin_out_cap, (size_t)bchunks * chk * 4 + 64)) {
        fprintf(stderr, "[%s:%d] ERROR: fail to alloc mem\n", __FILE__, __LINE__);
        return MZB_E_NOMEM;
    }
    plane_acct_t acct;
    memset(&acct, 0, sizeof(acct));
    int rc = MZB_OK;
    uint64_t zbytes = 0;
    unsigned char *in = (unsigned char *)ts->pin_in;
    for (uint64_t w0 = 0; w0 < words && rc == MZB_OK;) {
        /* gather up to bchunks chunk records: 16-byte header, then the four payloads (workers.c:61-69) */
        size_t fill = 0;
        uint64_t bw = 0;
        for (uint64_t c = 0; c < bchunks && w0 + bw < words; c++) {
            const uint64_t num = (words - w0 - bw) < chk ? (words - w0 - bw) : chk;
            if (fread(in + fill, 1, 16, fin) != 16) { rc = MZB_E_FORMAT; break; }
            size_t payload = 0;
            for (int j = 0; j < MZB_PLANES; j++) {
                btype_t bt;
                uint32_t len;
                unpack_header((const char *)in + fill + 4 * j, &bt, &len);
                if (!plane_header_ok(bt, len, num)) { rc = MZB_E_FORMAT; break; }
                payload += len;
            }
            if (rc != MZB_OK) break;
            if (fill + 16 + payload > in_cap) { rc = MZB_E_FORMAT; break; }
            if (fread(in + fill + 16, 1, payload, fin) != payload) { rc = MZB_E_FORMAT; break; }
            fill += 16 + payload;
            bw += num;
        }
        if (rc != MZB_OK) break;
        uint64_t nw = 0;
        rc = mzb_decompress_host(ts->ctx, in, fill, 0, chk, bw, ts->pin_out, ts->pin_out_cap / 4, &nw);
        if (rc != MZB_OK) break;
        account_records(in, fill, chk, bw, &acct);
        zbytes += fill;
        if (isTestThroughput != 1 && fwrite(ts->pin_out, sizeof(uint32_t), nw, fout) != nw) { rc = MZB_E_IO; break; }
        w0 += bw;
    }
    if (rc != MZB_OK) fprintf(stderr, "[%s:%d] ERROR: GPU decompress failed: %s\n", __FILE__, __LINE__, mzb_strerror(rc));
    const double dt = now_sec() - begin;
    /* workers.c:679-684 */
    ctx->allFileSize += words * MZB_PLANES;
    ctx->allZipFileSize += zbytes;
    ctx->unzipTime += dt;
    print_result_like(&acct, dt, 0, "Decompress Result Info");
    return rc;
}

/* ------------------------------------------------------------------ adapt.c:28-90 */

int zip_compress(ctx_t *ctx, const char *src, const char *dst, int bitsToLoss)
{
    FILE *fin = fopen(src, "rb");
    FILE *fout = fopen(dst, "wb");
    if (fin == NULL) { fprintf(stderr, "[%s:%d] ERROR: fail open:%s\n", __FILE__, __LINE__, src); exit(-1); }
    if (fout == NULL) { fprintf(stderr, "[%s:%d] ERROR: fail open:%s\n", __FILE__, __LINE__, dst); exit(-1); }
    ctx->fileCount += 1;
    ctx->allFileSize += get_file_size(fin);
    const int rc = run_compress(fin, ctx, fout, bitsToLoss, "float");
    fclose(fout);
    fclose(fin);
    return rc;
}

int zip_uncompress(ctx_t *ctx, const char *src, const char *dst)
{
    FILE *fin = fopen(src, "rb");
    FILE *fout = fopen(dst, "wb");
    if (fin == NULL) { fprintf(stderr, "[%s:%d] ERROR: fail open:%s\n", __FILE__, __LINE__, src); exit(-1); }
    if (fout == NULL) { fprintf(stderr, "[%s:%d] ERROR: fail open:%s\n", __FILE__, __LINE__, dst); exit(-1); }
    mrczip_header_t hd;
    ctx->fileCount += 1;
    init_mrczip_header(&hd, 0);
    if (read_mrczip_header(fin, &hd) != 0) {
        fclose(fout);
        fclose(fin);
        return -1;
    }
    const int rc = run_uncompress(fin, ctx, &hd, fout, "float");
    fclose(fout);
    fclose(fin);
    return rc;
}
