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
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <sys/types.h>
#include <unistd.h>

#include "../../include/mrczip_b200.h"

int isTestThroughput = 0; /* reference workers.c:39 */

/* chunks handed to the GPU per call from the FILE* API (bounds pinned + device memory per thread) */
#define HOST_BATCH_CHUNKS 16
/* small files share one pass of the kernels (zip_compress_many): at most this many chunks per pass, files of at most
 * MANY_FILE_CHUNKS chunks each (larger ones fill the GPU on their own) */
#define MANY_BATCH_CHUNKS 32
#define MANY_FILE_CHUNKS 8
/* file I/O of the FILE* API: batches are read / written by IO_THREADS threads with pread / pwrite (a single
 * thread copies page cache <-> pinned memory at a few GB/s, an order of magnitude below what the GPU path takes),
 * and the read of batch b+1 and the write of batch b-1 overlap the GPU work on batch b (two buffers each way).
 * Streams that cannot seek (pipes) take the plain fread / fwrite loop. */
#define IO_THREADS 16 /* upper bound; MRCZIP_IO_THREADS (default 8) says how many are used */
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

/* ---- the devices the FILE* entry points use.  Default: every visible CUDA device; MRCZIP_DEVICES=0,2,3 or
 * mzb_set_devices() name a subset, MRCZIP_DEVICE=<n> (older knob) a single one.  Calling threads are dealt devices
 * round robin (mrc_tarx's N workers land on N GPUs); a file of several batches is cut over all of them (below). */
#define MAX_DEVS 16
static int g_devs[MAX_DEVS], g_ndev = 0;
static pthread_mutex_t g_dev_mu = PTHREAD_MUTEX_INITIALIZER;
static unsigned g_next_thread_dev = 0;

static void devices_default_locked(void)
{
    const char *list = getenv("MRCZIP_DEVICES"), *one = getenv("MRCZIP_DEVICE");
    g_ndev = 0;
    if (list && *list) {
        const char *p = list;
        while (*p && g_ndev < MAX_DEVS) {
            char *end;
            const long v = strtol(p, &end, 10);
            if (end == p) break;
            int dup = v < 0 || v >= MAX_DEVS;
            for (int k = 0; k < g_ndev; k++) dup |= g_devs[k] == (int)v;
            if (!dup) g_devs[g_ndev++] = (int)v;
            p = *end == ',' ? end + 1 : end;
        }
    } else if (one && *one) {
        g_devs[g_ndev++] = atoi(one);
    } else {
        int n = mzb_device_count();
        if (n > MAX_DEVS) n = MAX_DEVS;
        for (int i = 0; i < n; i++) g_devs[g_ndev++] = i;
    }
    if (g_ndev == 0) g_devs[g_ndev++] = 0; /* mzb_create reports the missing device */
}

int mzb_set_devices(const int *devices, int n)
{
    if (n < 0 || n > MAX_DEVS || (n > 0 && !devices)) return MZB_E_ARG;
    for (int i = 0; i < n; i++)
        for (int k = 0; k < i; k++)
            if (devices[i] == devices[k] || devices[i] < 0 || devices[i] >= MAX_DEVS) return MZB_E_ARG; /* one worker per device */
    if (n == 1 && (devices[0] < 0 || devices[0] >= MAX_DEVS)) return MZB_E_ARG;
    pthread_mutex_lock(&g_dev_mu);
    if (n == 0) devices_default_locked();
    else {
        for (int i = 0; i < n; i++) g_devs[i] = devices[i];
        g_ndev = n;
    }
    pthread_mutex_unlock(&g_dev_mu);
    return MZB_OK;
}

static int devices_get(int *out)
{
    pthread_mutex_lock(&g_dev_mu);
    if (g_ndev == 0) devices_default_locked();
    const int n = g_ndev;
    for (int i = 0; i < n; i++) out[i] = g_devs[i];
    pthread_mutex_unlock(&g_dev_mu);
    return n;
}

static thread_state_t *thread_state(void)
{
    pthread_once(&g_once, make_key);
    thread_state_t *ts = (thread_state_t *)pthread_getspecific(g_key);
    if (ts) return ts;
    ts = (thread_state_t *)calloc(1, sizeof(*ts));
    if (!ts) return NULL;
    int devs[MAX_DEVS];
    const int nd = devices_get(devs);
    const int dev = devs[__atomic_fetch_add(&g_next_thread_dev, 1u, __ATOMIC_RELAXED) % (unsigned)nd];
    if (mzb_create(&ts->ctx, dev, NULL) != MZB_OK) {
        fprintf(stderr, "[%s:%d] ERROR: no usable CUDA device (this build has no CPU fallback)\n", __FILE__, __LINE__);
        free(ts);
        return NULL;
    }
    mzb_set_batch_chunks(ts->ctx, MANY_BATCH_CHUNKS);
    pthread_setspecific(g_key, ts);
    return ts;
}

/* one persistent context + staging per device for the multi-device file path (made on first use, kept for the life
 * of the process; a mutex per device: concurrent callers take turns on a GPU instead of piling contexts on it) */
typedef struct {
    pthread_mutex_t mu;
    thread_state_t ts;
    int ready;
} dev_slot_t;
static dev_slot_t g_slots[MAX_DEVS];
static pthread_once_t g_slots_once = PTHREAD_ONCE_INIT;
static void slots_init(void)
{
    for (int i = 0; i < MAX_DEVS; i++) pthread_mutex_init(&g_slots[i].mu, NULL);
}

/* locked on return (NULL: no context) */
static thread_state_t *dev_slot_acquire(int dev)
{
    pthread_once(&g_slots_once, slots_init);
    if (dev < 0 || dev >= MAX_DEVS) return NULL;
    dev_slot_t *sl = &g_slots[dev];
    pthread_mutex_lock(&sl->mu);
    if (!sl->ready) {
        memset(&sl->ts, 0, sizeof(sl->ts));
        if (mzb_create(&sl->ts.ctx, dev, NULL) != MZB_OK) { pthread_mutex_unlock(&sl->mu); return NULL; }
        mzb_set_batch_chunks(sl->ts.ctx, HOST_BATCH_CHUNKS);
        sl->ready = 1;
    }
    return &sl->ts;
}
static void dev_slot_release(int dev) { pthread_mutex_unlock(&g_slots[dev].mu); }

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

/* ------------------------------------------------------------------ MRC awareness */

static int g_mrc_aware = -1; /* -1: not set, ask the environment */

int mzb_set_mrc_aware(int on)
{
    __atomic_store_n(&g_mrc_aware, on ? 1 : 0, __ATOMIC_RELAXED);
    return MZB_OK;
}

static int mrc_aware(void)
{
    const int v = __atomic_load_n(&g_mrc_aware, __ATOMIC_RELAXED);
    if (v >= 0) return v;
    const char *e = getenv("MRCZIP_MRC_AWARE");
    return e && atoi(e) != 0;
}

/* Words at the head of the file that keep all their bits, and the bits to erase behind them.  The reference: always
 * 256 words = the 1024-byte MRC header (workers.c:90-94).  MRC-aware: 1024 + next bytes (the extended header is not
 * image data), and nothing is erased when the mode says the file is not float32 (mrcviewer.c:20-71). */
static uint64_t head_exempt_words(const void *first, size_t avail_bytes, uint64_t fsz, int *bits)
{
    mzb_mrc_info mi;
    if (!mrc_aware() || avail_bytes < 1024) return MZB_MRC_HEADER_WORDS;
    if (mzb_mrc_parse(first, 1024, &mi) != MZB_OK || mi.data_offset > fsz) return MZB_MRC_HEADER_WORDS;
    if (!mi.is_float32 && *bits != 0) {
        fprintf(stderr, "[mrczip_b200] MRC mode %d is not float32: no bits are erased (lossless)\n", mi.mode);
        *bits = 0;
    }
    return (mi.data_offset + 3) / 4;
}

static uint32_t batch_exempt(uint64_t exempt_total, uint64_t w0)
{
    const uint64_t e = exempt_total > w0 ? exempt_total - w0 : 0;
    return e > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)e;
}

/* ------------------------------------------------------------------ overlapped, multi-threaded file I/O */

typedef struct {
    int fd, wr;
    unsigned char *buf;
    size_t n, done;
    off_t off;
    unsigned char *map; /* != NULL: copy into a mapping of the file instead of pwrite */
} io_slice_t;

static void *io_slice_run(void *p)
{
    io_slice_t *s = (io_slice_t *)p;
    if (s->map) {
        memcpy(s->map, s->buf, s->n);
        s->done = s->n;
        return NULL;
    }
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
static int g_io_env = 8;
static void io_threads_init(void)
{
    const char *e = getenv("MRCZIP_IO_THREADS");
    const int w = e ? atoi(e) : 8;
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
static size_t io_parallel_map(int fd, int wr, void *buf, size_t n, off_t off, unsigned char *map);
static size_t io_parallel(int fd, int wr, void *buf, size_t n, off_t off) { return io_parallel_map(fd, wr, buf, n, off, NULL); }

static size_t io_parallel_map(int fd, int wr, void *buf, size_t n, off_t off, unsigned char *map)
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
        sl[i].map = map ? map + (b < n ? b : n) : NULL;
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

/* Writes of one batch into a regular file.  Buffered pwrite()s to ONE file take the inode lock one after the other
 * however many threads issue them (measured on tmpfs: 6 GB/s with 2, 4, 8 or 16 threads); page faults on a shared
 * mapping do not.  So the output file is sized up front (io_presize), a batch is copied into a mapping of its byte
 * range by the I/O threads, and the file is cut to its final length at the end.  A descriptor that cannot be mapped
 * for writing (opened write-only and not reachable through /proc/self/fd, or a file system without mmap) keeps pwrite. */
typedef struct {
    int fd;      /* the caller's descriptor (pwrite fallback) */
    int mfd;     /* read-write descriptor of the same file for mapping, -1: none */
    int use_map;
} io_out_t;

static void io_out_open(io_out_t *o, int fd)
{
    o->fd = fd;
    o->mfd = -1;
    o->use_map = 0;
    const char *e = getenv("MRCZIP_MMAP_WRITE");
    if (fd < 0 || (e && atoi(e) == 0)) return;
    const int fl = fcntl(fd, F_GETFL);
    if (fl >= 0 && (fl & O_ACCMODE) == O_RDWR) o->mfd = dup(fd);
    else {
        char path[64];
        snprintf(path, sizeof path, "/proc/self/fd/%d", fd);
        o->mfd = open(path, O_RDWR);
    }
    o->use_map = o->mfd >= 0;
}

static void io_out_close(io_out_t *o)
{
    if (o->mfd >= 0) close(o->mfd);
    o->mfd = -1;
    o->use_map = 0;
}

/* make the file at least `size` bytes long (sparse), once, before any batch is written */
static void io_presize(io_out_t *o, off_t size)
{
    struct stat st;
    if (!o->use_map) return;
    if (fstat(o->mfd, &st) != 0 || (st.st_size < size && ftruncate(o->mfd, size) != 0)) o->use_map = 0;
}

static size_t io_write(io_out_t *o, void *buf, size_t n, off_t off)
{
    if (o->use_map && n > 0) {
        const long page = sysconf(_SC_PAGESIZE);
        const off_t off0 = off & ~((off_t)page - 1);
        const size_t len = (size_t)(off - off0) + n;
        unsigned char *m = (unsigned char *)mmap(NULL, len, PROT_READ | PROT_WRITE, MAP_SHARED, o->mfd, off0);
        if (m != MAP_FAILED) {
            const size_t done = io_parallel_map(o->fd, 1, buf, n, off, m + (off - off0));
            munmap(m, len);
            return done;
        }
        o->use_map = 0; /* this file system does not map: pwrite from here on */
    }
    return io_parallel(o->fd, 1, buf, n, off);
}

static void io_finish(io_out_t *o, off_t end)
{
    if (o->mfd >= 0 && o->use_map) (void)!ftruncate(o->mfd, end);
    io_out_close(o);
}

/* one background transfer (a batch read ahead, or a batch written behind) */
typedef struct {
    pthread_t th;
    int active;
    int fd, wr;
    void *buf;
    size_t n, done;
    off_t off;
    io_out_t *out; /* writes go through io_write when set */
} io_job_t;

static void *io_job_run(void *p)
{
    io_job_t *j = (io_job_t *)p;
    j->done = (j->wr && j->out) ? io_write(j->out, j->buf, j->n, j->off) : io_parallel(j->fd, j->wr, j->buf, j->n, j->off);
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
    io_out_t ow;
    io_out_open(&ow, -1);
    if (isTestThroughput != 1) {
        write_mrczip_header(fout, hd);
        fdout = io_seekable(fout, 1, &pos_out);
        if (fdout < 0) return MZB_E_IO;
        io_out_open(&ow, fdout);
        io_presize(&ow, pos_out + (off_t)mzb_compress_bound(words_total, chk));
    }
    void *in[2] = {ts->pin_in, ts->pin_in2}, *out[2] = {ts->pin_out, ts->pin_out2};
    io_job_t jr, jw;
    memset(&jr, 0, sizeof(jr));
    memset(&jw, 0, sizeof(jw));
    jw.out = &ow;
    int rc = MZB_OK, writing = 0;
    uint64_t exempt_total = MZB_MRC_HEADER_WORDS;
    {
        const uint64_t w = words_total < batch_words ? words_total : batch_words;
        io_job_start(&jr, fdin, 0, in[0], (size_t)w * 4, pos_in);
    }
    for (uint64_t b = 0; b < nb && rc == MZB_OK; b++) {
        const uint64_t w0 = b * batch_words;
        const uint64_t want = (words_total - w0) < batch_words ? (words_total - w0) : batch_words;
        const uint64_t num = io_job_wait(&jr) / 4;
        if (num < want) { rc = MZB_E_IO; break; }   /* the file shrank under us */
        if (b == 0) exempt_total = head_exempt_words(in[0], (size_t)num * 4, hd->fsz, &bitsToMask);
        if (b + 1 < nb) {
            const uint64_t w1 = w0 + batch_words;
            const uint64_t nxt = (words_total - w1) < batch_words ? (words_total - w1) : batch_words;
            io_job_start(&jr, fdin, 0, in[(b + 1) & 1], (size_t)nxt * 4, pos_in + (off_t)(w1 * 4));
        }
        uint64_t sz = 0;
        rc = mzb_compress_host(ts->ctx, in[b & 1], num, bitsToMask, batch_exempt(exempt_total, w0), chk, hd->fsz, 0,
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
    io_finish(&ow, pos_out);
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
    io_out_t ow;
    io_out_open(&ow, -1);
    if (isTestThroughput != 1) {
        fdout = io_seekable(fout, 1, &pos_out);
        if (fdout < 0) return MZB_E_IO;
        io_out_open(&ow, fdout);
        io_presize(&ow, pos_out + (off_t)(words * 4));
    }
    void *in[2] = {ts->pin_in, ts->pin_in2}, *out[2] = {ts->pin_out, ts->pin_out2};
    io_job_t jr, jw;
    memset(&jr, 0, sizeof(jr));
    memset(&jw, 0, sizeof(jw));
    jw.out = &ow;
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
    io_finish(&ow, pos_out);
    if (fdout >= 0) fseeko(fout, pos_out, SEEK_SET);
    return rc;
}

/* ------------------------------------------------------------------ one file over several devices
 * Batches of HOST_BATCH_CHUNKS chunks are independent (every chunk record is): device g takes batches g, g + G, ...,
 * each worker reads its batch, runs it through its own GPU and writes the result at its place.  For compress the
 * place of batch b is known once the sizes of batches 0..b-1 are: the sizes are the only thing the workers exchange
 * ("only per-GPU chunk-size arrays are exchanged"), the container is byte-identical to the one-device one. */
typedef struct {
    int fdin, fdout; /* fdout < 0: isTestThroughput == 1, nothing is written */
    io_out_t ow;
    off_t pos_in, pos_out;
    uint32_t chk;
    uint64_t fsz, words_total, nb;
    size_t batch_words, in_cap, out_cap;
    int bits, ndev;
    const int *devs;
    uint64_t exempt_total;
    /* compress: sizes / offsets per batch; decompress: extents per batch (filled before the workers start) */
    uint64_t *size, *w0, *bw;
    off_t *offset;
    unsigned char *known;
    pthread_mutex_t mu;
    pthread_cond_t cv;
    int rc;
    plane_acct_t acct;
    uint64_t zbytes;
} md_job_t;

typedef struct {
    md_job_t *job;
    int g;
} md_arg_t;

static void md_fail(md_job_t *j, int rc)
{
    pthread_mutex_lock(&j->mu);
    if (j->rc == MZB_OK) j->rc = rc;
    pthread_cond_broadcast(&j->cv);
    pthread_mutex_unlock(&j->mu);
}

static int md_failed(md_job_t *j)
{
    pthread_mutex_lock(&j->mu);
    const int rc = j->rc;
    pthread_mutex_unlock(&j->mu);
    return rc != MZB_OK;
}

static void *md_zip_worker(void *p)
{
    md_arg_t *a = (md_arg_t *)p;
    md_job_t *j = a->job;
    const int dev = j->devs[a->g];
    thread_state_t *ts = dev_slot_acquire(dev);
    if (!ts) { md_fail(j, MZB_E_CUDA); return NULL; }
    if (pin_pair(ts, j->in_cap, j->out_cap, 0)) { dev_slot_release(dev); md_fail(j, MZB_E_NOMEM); return NULL; }
    for (uint64_t b = (uint64_t)a->g; b < j->nb && !md_failed(j); b += (uint64_t)j->ndev) {
        const uint64_t w0 = b * j->batch_words;
        const uint64_t want = (j->words_total - w0) < j->batch_words ? (j->words_total - w0) : j->batch_words;
        if (io_parallel(j->fdin, 0, ts->pin_in, (size_t)want * 4, j->pos_in + (off_t)(w0 * 4)) != (size_t)want * 4) { md_fail(j, MZB_E_IO); break; }
        uint64_t sz = 0;
        const int rc = mzb_compress_host(ts->ctx, ts->pin_in, want, j->bits, batch_exempt(j->exempt_total, w0), j->chk, j->fsz, 0,
                                         ts->pin_out, j->out_cap, &sz);
        if (rc != MZB_OK) { md_fail(j, rc); break; }
        pthread_mutex_lock(&j->mu);
        j->size[b] = sz;
        while (b > 0 && !j->known[b - 1] && j->rc == MZB_OK) pthread_cond_wait(&j->cv, &j->mu);
        const int ok = j->rc == MZB_OK;
        if (ok) {
            j->offset[b] = b ? j->offset[b - 1] + (off_t)j->size[b - 1] : j->pos_out;
            j->known[b] = 1;
            account_records((const unsigned char *)ts->pin_out, sz, j->chk, want, &j->acct);
            j->zbytes += sz;
            pthread_cond_broadcast(&j->cv);
        }
        const off_t at = j->offset[b];
        pthread_mutex_unlock(&j->mu);
        if (!ok) break;
        if (j->fdout >= 0 && io_write(&j->ow, ts->pin_out, (size_t)sz, at) != (size_t)sz) { md_fail(j, MZB_E_IO); break; }
    }
    dev_slot_release(dev);
    return NULL;
}

static void *md_unzip_worker(void *p)
{
    md_arg_t *a = (md_arg_t *)p;
    md_job_t *j = a->job;
    const int dev = j->devs[a->g];
    thread_state_t *ts = dev_slot_acquire(dev);
    if (!ts) { md_fail(j, MZB_E_CUDA); return NULL; }
    if (pin_pair(ts, j->in_cap, j->out_cap, 0)) { dev_slot_release(dev); md_fail(j, MZB_E_NOMEM); return NULL; }
    for (uint64_t b = (uint64_t)a->g; b < j->nb && !md_failed(j); b += (uint64_t)j->ndev) {
        const size_t bytes = (size_t)j->size[b];
        if (io_parallel(j->fdin, 0, ts->pin_in, bytes, j->offset[b]) != bytes) { md_fail(j, MZB_E_FORMAT); break; }
        uint64_t nw = 0;
        const int rc = mzb_decompress_host(ts->ctx, ts->pin_in, bytes, 0, j->chk, j->bw[b], ts->pin_out, j->out_cap / 4, &nw);
        if (rc != MZB_OK) { md_fail(j, rc); break; }
        pthread_mutex_lock(&j->mu);
        account_records((const unsigned char *)ts->pin_in, bytes, j->chk, j->bw[b], &j->acct);
        j->zbytes += bytes;
        pthread_mutex_unlock(&j->mu);
        if (j->fdout >= 0 && io_write(&j->ow, ts->pin_out, (size_t)nw * 4, j->pos_out + (off_t)(j->w0[b] * 4)) != (size_t)nw * 4) {
            md_fail(j, MZB_E_IO);
            break;
        }
    }
    dev_slot_release(dev);
    return NULL;
}

static int md_run(md_job_t *j, void *(*fn)(void *))
{
    pthread_t th[MAX_DEVS];
    md_arg_t args[MAX_DEVS];
    int started = 0;
    pthread_mutex_init(&j->mu, NULL);
    pthread_cond_init(&j->cv, NULL);
    for (int g = 0; g < j->ndev; g++) {
        args[g].job = j;
        args[g].g = g;
        if (pthread_create(&th[g], NULL, fn, &args[g]) == 0) started |= 1 << g;
        else fn(&args[g]);
    }
    for (int g = 0; g < j->ndev; g++)
        if (started & (1 << g)) pthread_join(th[g], NULL);
    pthread_cond_destroy(&j->cv);
    pthread_mutex_destroy(&j->mu);
    return j->rc;
}

/* run_compress over a seekable file of nb >= 2 batches on ndev >= 2 devices */
static int compress_multi_device(FILE *fin, int fdin, off_t pos_in, FILE *fout, mrczip_header_t *hd, int bitsToMask, const int *devs,
                                 int ndev, plane_acct_t *acct, uint64_t *zbytes)
{
    md_job_t j;
    memset(&j, 0, sizeof(j));
    io_out_open(&j.ow, -1);
    j.chk = hd->chk;
    j.batch_words = (size_t)HOST_BATCH_CHUNKS * j.chk;
    const uint64_t avail = hd->fsz > (uint64_t)pos_in ? hd->fsz - (uint64_t)pos_in : 0;
    j.words_total = avail / 4;
    j.nb = (j.words_total + j.batch_words - 1) / j.batch_words;
    j.fsz = hd->fsz;
    j.fdin = fdin;
    j.pos_in = pos_in;
    j.fdout = -1;
    j.bits = bitsToMask;
    j.devs = devs;
    j.ndev = (uint64_t)ndev < j.nb ? ndev : (int)j.nb;
    j.in_cap = j.batch_words * 4;
    j.out_cap = mzb_compress_bound(j.batch_words, j.chk);
    {   /* the head of the file decides how many words keep their bits (and, MRC-aware, whether any are erased) */
        unsigned char head[1024];
        const ssize_t got = pread(fdin, head, sizeof head, pos_in);
        j.exempt_total = head_exempt_words(head, got > 0 ? (size_t)got : 0, hd->fsz, &j.bits);
    }
    if (isTestThroughput != 1) {
        write_mrczip_header(fout, hd);
        j.fdout = io_seekable(fout, 1, &j.pos_out);
        if (j.fdout < 0) return MZB_E_IO;
        io_out_open(&j.ow, j.fdout);
        io_presize(&j.ow, j.pos_out + (off_t)mzb_compress_bound(j.words_total, j.chk));
    }
    j.size = (uint64_t *)calloc(j.nb, sizeof(uint64_t));
    j.offset = (off_t *)calloc(j.nb, sizeof(off_t));
    j.known = (unsigned char *)calloc(j.nb, 1);
    int rc = (j.size && j.offset && j.known) ? md_run(&j, md_zip_worker) : MZB_E_NOMEM;
    off_t end = j.pos_out;
    if (rc == MZB_OK) end = j.offset[j.nb - 1] + (off_t)j.size[j.nb - 1];
    free(j.size); free(j.offset); free(j.known);
    *acct = j.acct;
    *zbytes = j.zbytes;
    fseeko(fin, 0, SEEK_END);
    io_finish(&j.ow, end);
    if (j.fdout >= 0) fseeko(fout, end, SEEK_SET);
    return rc;
}

static int uncompress_multi_device(FILE *fin, int fdin, off_t pos_in, FILE *fout, uint32_t chk, uint64_t words, uint64_t bchunks,
                                   const int *devs, int ndev, plane_acct_t *acct, uint64_t *zbytes)
{
    md_job_t j;
    memset(&j, 0, sizeof(j));
    io_out_open(&j.ow, -1);
    const uint64_t nchunks = (words + chk - 1) / chk;
    j.chk = chk;
    j.nb = (nchunks + bchunks - 1) / bchunks;
    j.fdin = fdin;
    j.fdout = -1;
    j.devs = devs;
    j.ndev = (uint64_t)ndev < j.nb ? ndev : (int)j.nb;
    j.in_cap = (size_t)(bchunks * (16 + 4ull * (chk + 4ull))) + 64;
    j.out_cap = (size_t)bchunks * chk * 4 + 64;
    if (isTestThroughput != 1) {
        j.fdout = io_seekable(fout, 1, &j.pos_out);
        if (j.fdout < 0) return MZB_E_IO;
        io_out_open(&j.ow, j.fdout);
        io_presize(&j.ow, j.pos_out + (off_t)(words * 4));
    }
    j.size = (uint64_t *)calloc(j.nb, sizeof(uint64_t));
    j.offset = (off_t *)calloc(j.nb, sizeof(off_t));
    j.w0 = (uint64_t *)calloc(j.nb, sizeof(uint64_t));
    j.bw = (uint64_t *)calloc(j.nb, sizeof(uint64_t));
    int rc = (j.size && j.offset && j.w0 && j.bw) ? MZB_OK : MZB_E_NOMEM;
    off_t off = pos_in;
    uint64_t w = 0;
    for (uint64_t b = 0; b < j.nb && rc == MZB_OK; b++) {   /* the chunk-header chain (workers.c:61-69), 16 bytes per hop */
        size_t bytes = 0;
        uint64_t bw = 0;
        rc = walk_batch(fdin, off, bchunks, words - w, chk, j.in_cap, &bytes, &bw);
        j.size[b] = bytes; j.offset[b] = off; j.w0[b] = w; j.bw[b] = bw;
        off += (off_t)bytes;
        w += bw;
    }
    if (rc == MZB_OK) rc = md_run(&j, md_unzip_worker);
    free(j.size); free(j.offset); free(j.w0); free(j.bw);
    *acct = j.acct;
    *zbytes = j.zbytes;
    fseeko(fin, off, SEEK_SET);
    io_finish(&j.ow, j.pos_out + (off_t)(words * 4));
    if (j.fdout >= 0) fseeko(fout, j.pos_out + (off_t)(words * 4), SEEK_SET);
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
            int devs[MAX_DEVS];
            const int nd = devices_get(devs);
            const uint64_t words_avail = (hd2.fsz > (uint64_t)pos_in ? hd2.fsz - (uint64_t)pos_in : 0) / 4;
            const int rc2 = (nd > 1 && words_avail > batch_words)
                                ? compress_multi_device(fin, fdin, pos_in, fout, &hd2, bitsToMask, devs, nd, &acct2, &zb)
                                : compress_overlapped(ts, fin, fdin, pos_in, fout, &hd2, bitsToMask, &acct2, &zb);
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
    int rc = MZB_OK, bits = bitsToMask;
    uint64_t zbytes = 0, w0 = 0, exempt_total = MZB_MRC_HEADER_WORDS;
    /* fread of 4-byte items: a ragged tail of 1..3 bytes is dropped exactly like workers.c:744,854 */
    size_t num = fread(ts->pin_in, sizeof(uint32_t), batch_words, fin);
    if (num > 0 && isTestThroughput != 1) write_mrczip_header(fout, &hd); /* workers.c:757-764 */
    if (num > 0) exempt_total = head_exempt_words(ts->pin_in, num * 4, hd.fsz, &bits);
    while (num > 0) {
        uint64_t sz = 0;
        rc = mzb_compress_host(ts->ctx, ts->pin_in, num, bits, batch_exempt(exempt_total, w0), chk, hd.fsz, 0,
                               ts->pin_out, ts->pin_out_cap, &sz);
        if (rc != MZB_OK) {
            fprintf(stderr, "[%s:%d] ERROR: GPU compress failed: %s\n", __FILE__, __LINE__, mzb_strerror(rc));
            break;
        }
        w0 += num;
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
            int devs[MAX_DEVS];
            const int nd = devices_get(devs);
            const int rc2 = (nd > 1 && (words + chk - 1) / chk > bchunks)
                                ? uncompress_multi_device(fin, fdin, pos_in, fout, chk, words, bchunks, devs, nd, &acct2, &zb)
                                : uncompress_overlapped(ts, fin, fdin, pos_in, fout, chk, words, bchunks, &acct2, &zb);
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

/* ------------------------------------------------------------------ many small files, one pass of the kernels each
 * (the reference's deployment shape is thousands of small MRC stacks, mrc_tarx.c:134-176; one at a
 * time they under-fill a GPU).  Files of at most MANY_FILE_CHUNKS chunks are gathered, in list order, into groups of
 * at most MANY_BATCH_CHUNKS chunks that go through mzb_compress_host_many / mzb_decompress_host_many; every output
 * file is byte-identical to what zip_compress / zip_uncompress write for it.  Larger files (and, MRC-aware, files
 * that are not float32) take the one-file path. */
typedef struct {
    FILE *fin;
    uint64_t fsz, words;
    uint32_t chunks;
    int idx;
} many_file_t;

static int many_flush_zip(thread_state_t *ts, ctx_t *ctx, many_file_t *f, int nf, const char *const *dsts, int bits)
{
    if (nf <= 0 || nf > MANY_BATCH_CHUNKS) return MZB_OK;
    const double begin = now_sec();
    size_t in_need = 0, out_need = 0;
    for (int i = 0; i < nf; i++) {
        in_need += (size_t)f[i].words * 4 + 64;
        out_need += mzb_compress_bound(f[i].words, MZB_CHUNK_WORDS) + 64;
    }
    int rc = MZB_OK;
    if (pin_reserve(&ts->pin_in, &ts->pin_in_cap, in_need) || pin_reserve(&ts->pin_out, &ts->pin_out_cap, out_need)) rc = MZB_E_NOMEM;
    mzb_zip_item *it = (mzb_zip_item *)calloc((size_t)nf, sizeof(*it));
    if (!it) rc = MZB_E_NOMEM;
    size_t ia = 0, oa = 0;
    for (int i = 0; i < nf && rc == MZB_OK; i++) {
        unsigned char *in = (unsigned char *)ts->pin_in + ia;
        if (fread(in, 4, (size_t)f[i].words, f[i].fin) != (size_t)f[i].words) { rc = MZB_E_IO; break; }
        int b2 = bits;
        it[i].h_words = in;
        it[i].nwords = f[i].words;
        it[i].exempt_words = (uint32_t)head_exempt_words(in, (size_t)f[i].words * 4, f[i].fsz, &b2);
        it[i].fsz = f[i].fsz;
        it[i].h_out = (unsigned char *)ts->pin_out + oa;
        it[i].out_cap = mzb_compress_bound(f[i].words, MZB_CHUNK_WORDS);
        ia += (size_t)f[i].words * 4 + 64;
        ia &= ~(size_t)15;
        oa += it[i].out_cap + 64;
        oa &= ~(size_t)15;
    }
    if (rc == MZB_OK) rc = mzb_compress_host_many(ts->ctx, it, (uint32_t)nf, bits, MZB_CHUNK_WORDS, 1);
    if (rc != MZB_OK) fprintf(stderr, "[%s:%d] ERROR: GPU compress of a group of %d files failed: %s\n", __FILE__, __LINE__, nf, mzb_strerror(rc));
    for (int i = 0; i < nf; i++) {
        {   /* adapt.c:28-52 opens (and so creates) the destination whatever isTestThroughput says */
            FILE *fo = fopen(dsts[f[i].idx], "wb");
            if (!fo) { fprintf(stderr, "[%s:%d] ERROR: fail open:%s\n", __FILE__, __LINE__, dsts[f[i].idx]); exit(-1); }
            if (rc == MZB_OK && isTestThroughput != 1 && it && it[i].out_size &&
                fwrite(it[i].h_out, 1, (size_t)it[i].out_size, fo) != (size_t)it[i].out_size)
                rc = MZB_E_IO;
            fclose(fo);
        }
        if (rc == MZB_OK && it) {
            ctx->fileCount += 1;
            ctx->allFileSize += f[i].fsz;
            ctx->allZipFileSize += it[i].out_size > MZB_FILE_HEADER_BYTES ? it[i].out_size - MZB_FILE_HEADER_BYTES : 0;
        }
        fclose(f[i].fin);
    }
    ctx->zipTime += now_sec() - begin;
    free(it);
    return rc;
}

int zip_compress_many(ctx_t *ctx, int n, const char *const *srcs, const char *const *dsts, int bitsToLoss)
{
    if (!ctx || n < 0 || (n && (!srcs || !dsts)) || bitsToLoss < 0 || bitsToLoss > 32) return MZB_E_ARG;
    thread_state_t *ts = thread_state();
    if (!ts) return MZB_E_CUDA;
    many_file_t grp[MANY_BATCH_CHUNKS];
    int ng = 0, rc = MZB_OK;
    uint32_t chunks = 0;
    for (int i = 0; i < n; i++) {
        FILE *fin = fopen(srcs[i], "rb");
        if (!fin) { fprintf(stderr, "[%s:%d] ERROR: fail open:%s\n", __FILE__, __LINE__, srcs[i]); exit(-1); }
        const uint64_t fsz = get_file_size(fin), words = fsz / 4;
        const uint64_t ch = (words + MZB_CHUNK_WORDS - 1) / MZB_CHUNK_WORDS;
        int single = ch == 0 || ch > MANY_FILE_CHUNKS;
        if (!single && mrc_aware() && bitsToLoss) {   /* a file that is not float32 is stored with no bits erased: its own pass */
            unsigned char head[1024];
            int b2 = bitsToLoss;
            const size_t got = fread(head, 1, sizeof head, fin);
            rewind(fin);
            head_exempt_words(head, got, fsz, &b2);
            single = b2 != bitsToLoss;
        }
        if (single) {
            fclose(fin);
            const int r = zip_compress(ctx, srcs[i], dsts[i], bitsToLoss);
            if (r != MZB_OK && rc == MZB_OK) rc = r;
            continue;
        }
        if (chunks + ch > MANY_BATCH_CHUNKS || ng == MANY_BATCH_CHUNKS) {
            const int r = many_flush_zip(ts, ctx, grp, ng, dsts, bitsToLoss);
            if (r != MZB_OK && rc == MZB_OK) rc = r;
            ng = 0; chunks = 0;
        }
        grp[ng].fin = fin; grp[ng].fsz = fsz; grp[ng].words = words; grp[ng].chunks = (uint32_t)ch; grp[ng].idx = i;
        ng++;
        chunks += (uint32_t)ch;
    }
    const int r = many_flush_zip(ts, ctx, grp, ng, dsts, bitsToLoss);
    return rc != MZB_OK ? rc : r;
}

static int many_flush_unzip(thread_state_t *ts, ctx_t *ctx, many_file_t *f, int nf, const char *const *dsts)
{
    if (nf <= 0 || nf > MANY_BATCH_CHUNKS) return MZB_OK;
    const double begin = now_sec();
    size_t in_need = 0, out_need = 0;
    for (int i = 0; i < nf; i++) {
        in_need += (size_t)f[i].fsz + 64;          /* fsz: bytes of chunk records here */
        out_need += (size_t)f[i].words * 4 + 64;
    }
    int rc = MZB_OK;
    if (pin_reserve(&ts->pin_in, &ts->pin_in_cap, in_need) || pin_reserve(&ts->pin_out, &ts->pin_out_cap, out_need)) rc = MZB_E_NOMEM;
    mzb_unzip_item *it = (mzb_unzip_item *)calloc((size_t)nf, sizeof(*it));
    if (!it) rc = MZB_E_NOMEM;
    size_t ia = 0, oa = 0;
    for (int i = 0; i < nf && rc == MZB_OK; i++) {
        unsigned char *in = (unsigned char *)ts->pin_in + ia;
        if (fread(in, 1, (size_t)f[i].fsz, f[i].fin) != (size_t)f[i].fsz) { rc = MZB_E_FORMAT; break; }
        it[i].h_in = in;
        it[i].in_size = (size_t)f[i].fsz;
        it[i].nwords = f[i].words;
        it[i].h_words_out = (unsigned char *)ts->pin_out + oa;
        it[i].out_cap_words = f[i].words;
        ia = (ia + (size_t)f[i].fsz + 64) & ~(size_t)15;
        oa = (oa + (size_t)f[i].words * 4 + 64) & ~(size_t)15;
    }
    if (rc == MZB_OK) rc = mzb_decompress_host_many(ts->ctx, it, (uint32_t)nf, MZB_CHUNK_WORDS);
    if (rc != MZB_OK) fprintf(stderr, "[%s:%d] ERROR: GPU decompress of a group of %d files failed: %s\n", __FILE__, __LINE__, nf, mzb_strerror(rc));
    for (int i = 0; i < nf; i++) {
        {
            FILE *fo = fopen(dsts[f[i].idx], "wb");
            if (!fo) { fprintf(stderr, "[%s:%d] ERROR: fail open:%s\n", __FILE__, __LINE__, dsts[f[i].idx]); exit(-1); }
            if (rc == MZB_OK && isTestThroughput != 1 && it && fwrite(it[i].h_words_out, 4, (size_t)f[i].words, fo) != (size_t)f[i].words)
                rc = MZB_E_IO;
            fclose(fo);
        }
        if (rc == MZB_OK) {
            ctx->fileCount += 1;
            ctx->allFileSize += f[i].words * 4;
            ctx->allZipFileSize += f[i].fsz;
        }
        fclose(f[i].fin);
    }
    ctx->unzipTime += now_sec() - begin;
    free(it);
    return rc;
}

int zip_uncompress_many(ctx_t *ctx, int n, const char *const *srcs, const char *const *dsts)
{
    if (!ctx || n < 0 || (n && (!srcs || !dsts))) return MZB_E_ARG;
    thread_state_t *ts = thread_state();
    if (!ts) return MZB_E_CUDA;
    many_file_t grp[MANY_BATCH_CHUNKS];
    int ng = 0, rc = MZB_OK;
    uint32_t chunks = 0;
    for (int i = 0; i < n; i++) {
        FILE *fin = fopen(srcs[i], "rb");
        if (!fin) { fprintf(stderr, "[%s:%d] ERROR: fail open:%s\n", __FILE__, __LINE__, srcs[i]); exit(-1); }
        const uint64_t zsz = get_file_size(fin);
        mrczip_header_t hd;
        init_mrczip_header(&hd, 0);
        int single = zsz < MZB_FILE_HEADER_BYTES || read_mrczip_header(fin, &hd) != 0 || hd.chk != MZB_CHUNK_WORDS;
        const uint64_t words = hd.fsz / MZB_PLANES, ch = (words + MZB_CHUNK_WORDS - 1) / MZB_CHUNK_WORDS;
        for (int j = 0; j < MZB_PLANES; j++) single |= hd.ztypes[j] != 0;
        single |= ch == 0 || ch > MANY_FILE_CHUNKS;
        if (single) {   /* large, empty, odd chunk size or malformed: the one-file path (and its error handling) */
            fclose(fin);
            const int r = zip_uncompress(ctx, srcs[i], dsts[i]);
            if (r != MZB_OK && rc == MZB_OK) rc = r;
            continue;
        }
        if (chunks + ch > MANY_BATCH_CHUNKS || ng == MANY_BATCH_CHUNKS) {
            const int r = many_flush_unzip(ts, ctx, grp, ng, dsts);
            if (r != MZB_OK && rc == MZB_OK) rc = r;
            ng = 0; chunks = 0;
        }
        grp[ng].fin = fin; grp[ng].fsz = zsz - MZB_FILE_HEADER_BYTES; grp[ng].words = words; grp[ng].chunks = (uint32_t)ch; grp[ng].idx = i;
        ng++;
        chunks += (uint32_t)ch;
    }
    const int r = many_flush_unzip(ts, ctx, grp, ng, dsts);
    return rc != MZB_OK ? rc : r;
}
