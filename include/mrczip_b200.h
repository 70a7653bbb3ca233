/*
 * mrczip_b200.h -- C ABI of libmrczip_b200.so: the B200 (sm_100a) drop-in for the float32
 * compress / decompress hot path of ruanhuabin/DataCompressionFloat.
 *
 * Two groups of entry points:
 *
 *  (1) The reference's own names and signatures, so a reference build can link this library
 *      instead of its src/core objects (boundary B of SURVEY.md):
 *        run_compress / run_uncompress        reference src/include/workers.h:30-31
 *        zip_compress / zip_uncompress        reference src/include/adapt.h:30-31
 *        pack_header / unpack_header          reference src/include/mrczip.h:123-124
 *        init/read/write/print_mrczip_header, init/reset/update_context, print_context_info,
 *        get_file_size, now_sec               reference src/include/common.h:43-81
 *        isTestThroughput                     reference src/core/workers.c:39
 *      Types ctx_t and mrczip_header_t are layout-compatible with common.h:33-56.
 *
 *  (2) mzb_* : the thin CUDA layer underneath (plain pointers and sizes, no torch types), for
 *      callers that already hold the data in host or device memory.
 *
 * Conventions: every function returns 0 on success (the reference's only return value) and a
 * negative MZB_E_* where the reference would exit(-1) or silently produce garbage.  There is NO CPU
 * fallback: without a CUDA device the calls fail with MZB_E_CUDA.
 */
#ifndef MRCZIP_B200_H_
#define MRCZIP_B200_H_

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define MZB_OK 0
#define MZB_E_ARG (-1)
#define MZB_E_CUDA (-2)
#define MZB_E_NOMEM (-3)
#define MZB_E_FORMAT (-4)
#define MZB_E_SPACE (-5)
#define MZB_E_IO (-6)

#define MZB_PLANES 4                     /* constant.h:27 COMPRESSION_PATH_NUM */
#define MZB_CHUNK_WORDS (6u * 1048576u)  /* constant.h:25 CHUNK_SIZE (elements) */
#define MZB_FILE_HEADER_BYTES 17         /* common.c:137-149 */
#define MZB_MRC_HEADER_WORDS 256         /* workers.c:90-94 */

/* ------------------------------------------------------------------ (1) reference-named interface */

/* common.h:33-41 */
typedef struct _context_t {
    uint32_t fileCount;
    uint64_t allFileSize;
    uint64_t allZipFileSize;
    double zipTime;
    double unzipTime;
} ctx_t;

/* common.h:50-56 (COMPRESSION_PATH_NUM == 4) */
typedef struct _mrczip_header_t {
    uint64_t fsz;
    uint32_t chk;
    char type;
    char ztypes[MZB_PLANES];
} mrczip_header_t;

/* mrczip.h:31-34 */
typedef enum { COMPRESSED = 0, RAW } btype_t;

extern int isTestThroughput; /* workers.c:39: 1 = suppress every fwrite */

/* workers.h:30: fin at offset 0; writes the 17-byte file header + chunk records to fout.
 * compressPrecision = bits to erase, 0..32 (values outside are rejected with MZB_E_ARG; the reference
 * indexes its 33-entry table unchecked).  dataConvertedType must be "float" ("int" is out of scope). */
int run_compress(FILE *fin, ctx_t *ctx, FILE *fout, const int compressPrecision, const char *dataConvertedType);
/* workers.h:31: fin positioned just past the 17-byte header the caller already parsed into hd. */
int run_uncompress(FILE *fin, ctx_t *ctx, mrczip_header_t *hd, FILE *fout, const char *dataConvertedType);
/* adapt.h:30-31 */
int zip_compress(ctx_t *ctx, const char *src, const char *dst, int bitsToLoss);
int zip_uncompress(ctx_t *ctx, const char *src, const char *dst);

/* Extension for list front ends (reference mrc_tarx.c:134-176 feeds zip_compress one file at a time): n files at
 * once; files of at most 8 chunks are gathered into groups of at most 32 chunks that take ONE pass of the GPU kernels
 * each (small stacks under-fill a GPU one by one).  Every output file is byte-identical to zip_compress's /
 * zip_uncompress's; ctx is accumulated the same way. */
int zip_compress_many(ctx_t *ctx, int n, const char *const *srcs, const char *const *dsts, int bitsToLoss);
int zip_uncompress_many(ctx_t *ctx, int n, const char *const *srcs, const char *const *dsts);

/* mrczip.h:123-124 / zip.c:381-399 */
void pack_header(char *buf, btype_t btype, uint32_t len);
void unpack_header(const char *buf, btype_t *btype, uint32_t *len);

/* common.h:43-81 / common.c */
void init_context(ctx_t *ctx);
void reset_context(ctx_t *ctx);
void update_context(ctx_t *dst, ctx_t *src);
void print_context_info(ctx_t *ctx, const char *hintMsg);
void init_mrczip_header(mrczip_header_t *hd, char type);
int read_mrczip_header(FILE *fin, mrczip_header_t *hd);
int write_mrczip_header(FILE *fout, mrczip_header_t *hd);
void print_mrczip_header(mrczip_header_t *hd, const char *hintMsg);
uint64_t get_file_size(FILE *fp);
double now_sec(void);

/* ------------------------------------------------------------------ (2) CUDA layer */

typedef struct mzb_ctx mzb_ctx;

/* One context per (device, stream).  cuda_stream is a cudaStream_t (NULL: the context creates its own
 * non-blocking stream).  All device work of a context is ordered on that stream.  A context is not
 * thread-safe; use one per thread (the reference calls run_* from N pthreads, mrc_tarx.c:145-161). */
int mzb_create(mzb_ctx **out, int device, void *cuda_stream);
/* Same, but always uses the given stream handle -- including NULL, the legacy default stream. */
int mzb_create_on_stream(mzb_ctx **out, int device, void *cuda_stream);
void mzb_destroy(mzb_ctx *ctx);
/* Chunks processed per kernel batch (bounds scratch memory: about 8 bytes per word of a batch). Default 192 (4.5 GiB of input). */
int mzb_set_batch_chunks(mzb_ctx *ctx, uint32_t chunks);
/* kernel variant selectors used by the benchmarks (0 = default) */
int mzb_set_variant(mzb_ctx *ctx, int split_variant, int merge_variant);
/* inflater of our own streams: 0 = the lean table-loop kernel with the full group kernel behind it (default), 1 = the
 * full group kernel alone (what decodes a code group the lean kernel gives up on; the tests run both) */
int mzb_set_inflate_variant(mzb_ctx *ctx, int variant);

/* CUDA devices visible to the process (0 without a driver / device). */
int mzb_device_count(void);
/* The devices the FILE* entry points use (process wide).  Default (also n = 0): every visible device, or the
 * environment variables MRCZIP_DEVICES=0,2,3 / MRCZIP_DEVICE=<n>.  Calling threads are dealt devices round robin
 * (mrc_tarx's N workers, reference mrc_tarx.c:145-161, land on N GPUs); one file of more than one 16-chunk batch is
 * cut over all of them, batch by batch, and its container is byte-identical to the one-device container. */
int mzb_set_devices(const int *devices, int n);

/* Threads that pread / pwrite one batch of the FILE* entry points (1..16; 0 = back to the default: the environment
 * variable MRCZIP_IO_THREADS, else 8).  Process wide. */
int mzb_set_io_threads(int n);

/* Upper bound of the container for nwords words cut in chk-word chunks (header included). */
size_t mzb_compress_bound(uint64_t nwords, uint32_t chk);

/* Device-resident compress.  d_words: nwords uint32 (16-byte aligned device pointer).
 *   bits          0..32 low mantissa bits to erase (workers.c:29-37)
 *   exempt_words  leading words left unmasked: 256 for the first chunk range of an MRC file, 0 for a
 *                 later chunk range (multi-GPU sharding) -- workers.c:90-94
 *   chk           words per chunk written to the file header; MZB_CHUNK_WORDS is the reference value
 *   fsz           original file size in bytes for the 17-byte header; write_file_header = 0 emits chunk
 *                 records only (a shard of a larger container)
 * Writes to d_out (device, capacity out_cap); *out_size = bytes produced. Synchronises the stream. */
int mzb_compress_device(mzb_ctx *ctx, const void *d_words, uint64_t nwords, int bits, uint32_t exempt_words,
                        uint32_t chk, uint64_t fsz, int write_file_header, void *d_out, size_t out_cap,
                        uint64_t *out_size);

/* Device-resident decompress of chunk records.  d_in/in_size: container bytes on the device;
 * has_file_header != 0: d_in starts with the 17-byte header (chk and nwords are then read from it and the
 * arguments ignored); otherwise d_in starts at a chunk record and (chk, nwords) describe the shard.
 * d_words_out: device buffer for nwords uint32 (capacity out_cap_words). Synchronises the stream. */
int mzb_decompress_device(mzb_ctx *ctx, const void *d_in, size_t in_size, int has_file_header, uint32_t chk,
                          uint64_t nwords, void *d_words_out, uint64_t out_cap_words, uint64_t *nwords_out);

/* Intermediates, for parity checks against the reference's split/merge (workers.c:180-203, 423-442):
 * d_planes holds 4 planes of nwords bytes, plane j at d_planes + j * plane_stride (stride % 16 == 0).
 * Asynchronous on a caller-provided stream (mzb_create_on_stream); a context with its own stream synchronises. */
int mzb_mask_split_device(mzb_ctx *ctx, const void *d_words, uint64_t nwords, int bits, uint32_t exempt_words,
                          void *d_planes, uint64_t plane_stride);
int mzb_merge_device(mzb_ctx *ctx, const void *d_planes, uint64_t plane_stride, uint64_t nwords, void *d_words_out);

/* Host-buffer versions (end to end: H2D, kernels, D2H inside the call).  Pinned host memory gives
 * full PCIe rate; pageable memory works. */
int mzb_compress_host(mzb_ctx *ctx, const void *h_words, uint64_t nwords, int bits, uint32_t exempt_words,
                      uint32_t chk, uint64_t fsz, int write_file_header, void *h_out, size_t out_cap,
                      uint64_t *out_size);
int mzb_decompress_host(mzb_ctx *ctx, const void *h_in, size_t in_size, int has_file_header, uint32_t chk,
                        uint64_t nwords, void *h_words_out, uint64_t out_cap_words, uint64_t *nwords_out);

/* ---- MRC awareness and the error report ------------------------------------------------------------
 * The reference treats the first 1024 bytes of every file as the MRC header and masks everything behind it
 * (workers.c:90-94), whatever the header says.  mzb_mrc_parse reads the fields that matter (reference
 * src/tool/mrcviewer.c:20-71: nx, ny, nz, mod at words 0..3, next = bytes of extended header at word 23). */
typedef struct {
    int32_t nx, ny, nz;
    int32_t mode;          /* 2 = float32, the only mode whose low mantissa bits may be erased */
    int32_t next;          /* bytes of extended header between the 1024-byte header and the data */
    int32_t is_float32;
    uint64_t data_offset;  /* 1024 + next */
} mzb_mrc_info;
/* header: at least 1024 bytes.  MZB_E_FORMAT when the fields are not those of an MRC header (non-positive
 * dimensions, unknown mode, negative next). */
int mzb_mrc_parse(const void *header, size_t len, mzb_mrc_info *out);
/* Process wide, default 0 (byte-compatible with the reference).  1: run_compress / zip_compress read the header and
 * (a) leave 1024 + next bytes unmasked instead of 1024, (b) erase no bits at all when the mode is not float32
 * (the file is then stored losslessly; a note goes to stderr), (c) fall back to the reference's behaviour when the
 * first 1024 bytes are not an MRC header.  The environment variable MRCZIP_MRC_AWARE=1 does the same. */
int mzb_set_mrc_aware(int on);

/* What the reference's erroranalysis tool reports after a lossy round trip (src/tool/erroranalysis.c:188-220):
 * err = |n2 - n1|, relative error err / |n1| where |n1| > 10E-4 and 0 elsewhere; here the two maxima, where they
 * are, and the sum, from one pass on the GPU. */
typedef struct {
    uint64_t count;          /* pairs compared */
    uint64_t nan_count;      /* pairs whose error is not a number (NaN on either side, Inf - Inf): not ranked */
    float max_abs_err, max_abs_n1, max_abs_n2, max_rel_err, max_rel_n1, max_rel_n2;
    uint64_t max_abs_index, max_rel_index;   /* word index (lowest on ties); ~0 when nothing was ranked */
    double sum_abs_err;      /* mean absolute error = sum_abs_err / (count - nan_count) */
} mzb_error_report_t;
/* d_orig / d_other: nwords float32 each on the device.  d_other == NULL: n2 = n1 with `bits` low bits erased behind
 * `exempt_words` words -- the error of the mask itself, no round trip needed (bits is ignored otherwise). */
int mzb_error_report_device(mzb_ctx *ctx, const void *d_orig, const void *d_other, uint64_t nwords, int bits,
                            uint32_t exempt_words, mzb_error_report_t *out);
/* the same on host buffers (staged through the device in batches) */
int mzb_error_report_host(mzb_ctx *ctx, const void *h_orig, const void *h_other, uint64_t nwords, int bits,
                          uint32_t exempt_words, mzb_error_report_t *out);

/* ---- several small inputs in ONE pass of the kernels (thousands of small MRC stacks under-fill a GPU
 * one at a time).  Every item is its own file: own header exemption, own ragged last chunk, own container.  All items
 * share chk (a multiple of 16) and bits; their chunk counts must add up to at most the context's batch size
 * (mzb_set_batch_chunks, default 192).  Results are byte-identical to one mzb_compress_host / mzb_decompress_host
 * call per item. */
typedef struct {
    const void *h_words;    /* in:  nwords uint32 */
    uint64_t nwords;
    uint32_t exempt_words;  /* in:  leading words that keep their bits (256 for an MRC file) */
    uint32_t reserved;
    uint64_t fsz;           /* in:  original file size for the 17-byte header */
    void *h_out;            /* out: container (17-byte header when write_file_header, then the chunk records) */
    size_t out_cap;
    uint64_t out_size;      /* out */
} mzb_zip_item;
int mzb_compress_host_many(mzb_ctx *ctx, mzb_zip_item *items, uint32_t n, int bits, uint32_t chk, int write_file_header);

typedef struct {
    const void *h_in;        /* in:  chunk records of the item (no 17-byte header) */
    size_t in_size;
    uint64_t nwords;         /* in:  words the records inflate to */
    void *h_words_out;       /* out: nwords uint32 */
    uint64_t out_cap_words;
} mzb_unzip_item;
int mzb_decompress_host_many(mzb_ctx *ctx, mzb_unzip_item *items, uint32_t n, uint32_t chk);

/* pinned (page-locked) host memory for callers written in C without the CUDA headers */
void *mzb_host_alloc(size_t bytes);
void mzb_host_free(void *p);

/* passes of the kernel pipelines since the process started (one per batch of chunks handed to the GPU): what a list
 * front end looks at to see that small files really share passes */
void mzb_pass_counts(uint64_t *compress_passes, uint64_t *decompress_passes);

/* counters of the last compress / decompress call */
typedef struct {
    uint64_t bytes_in, bytes_out;
    uint32_t chunks, streams;
    uint32_t raw_streams;      /* streams written RAW (zip.c:186-190) */
    uint32_t stored_subblocks; /* sub-blocks emitted as stored deflate blocks */
    uint32_t general_streams;  /* decode: streams not in this library's sub-block framing (e.g. written by the reference's zlib) */
    uint32_t fast_failed;      /* decode: streams whose sub-block decode failed validation (fell back) */
    uint32_t kernel_launches;  /* kernels launched by the call */
    uint32_t blockpar_streams; /* decode: general streams inflated block-parallel (the rest: one thread per stream) */
    uint32_t zero_subblocks;   /* encode: 16 KiB sub-blocks that are all zero bytes (coded once per group of 32) */
    uint32_t reserved;
} mzb_stats;
int mzb_last_stats(mzb_ctx *ctx, mzb_stats *out);

/* Per-stage device times of the last compress / decompress call, measured with CUDA events on the
 * context's stream (off by default; the events cost a few microseconds per batch). */
int mzb_set_profiling(mzb_ctx *ctx, int on);
int mzb_stage_count(void);
const char *mzb_stage_name(int stage);
int mzb_stage_ms(mzb_ctx *ctx, float *out_ms, int n);

const char *mzb_version(void);
const char *mzb_strerror(int code);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif
#endif /* MRCZIP_B200_H_ */
