/*
 * mrczip_oracle.h -- CPU restatement of the reference hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is the checker, not the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (datacompressionfloat_b200/) never links or imports anything under oracle/.
 *
 * Every function cites the reference file:line (relative to /root/reference) it restates.
 * Parity status: mask / split / merge / container are pinned byte-for-byte against the
 * reference itself built into oracle/_ref (tests/golden/*, tests/test_oracle.py).  The
 * deflate bytes come from zlib (the reference links a prebuilt zlib 1.2.8 binary,
 * lib/libz.a, source not vendored); this restatement drives the system zlib with the
 * reference's parameters and is pinned against containers produced by the reference
 * binary (tests/golden/manifest.json).
 */
#ifndef MRCZIP_ORACLE_H_
#define MRCZIP_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_PLANES 4                 /* constant.h:27 COMPRESSION_PATH_NUM */
#define ORC_CHUNK_WORDS (6 * 1048576) /* constant.h:25 CHUNK_SIZE (elements) */
#define ORC_HDR_SIZE 4               /* mrczip.h:119 HDR_SIZE */
#define ORC_FILE_HEADER_BYTES 17     /* common.c:137-149 field-wise write */
#define ORC_MRC_HEADER_WORDS 256     /* workers.c:90-94 (1024-byte MRC header) */

/* workers.c:29-37 bitsMaskTable[b] == 0xFFFFFFFF << b, table[32] == 0. Returns 0 for b outside 0..32. */
uint32_t orc_mask_for_bits(int bits);

/* workers.c:82-101 apply_mask(buffer, 0, num, bits, isFirstChk) */
void orc_apply_mask(uint32_t *words, int64_t num, int bits, int is_first_chunk);

/* workers.c:180-203 split_float_to_byte_stream: masks `words` IN PLACE, then planes[j][i] = byte j of words[i] */
void orc_split(uint32_t *words, int64_t num, uint8_t *planes[ORC_PLANES], int bits, int is_first_chunk);

/* workers.c:423-442 merge_byte_to_float_stream */
void orc_merge(uint32_t *words, int64_t num, uint8_t *const planes[ORC_PLANES]);

/* zip.c:381-399 pack_header / unpack_header */
void orc_pack_header(uint8_t buf[4], int btype, uint32_t len);
void orc_unpack_header(const uint8_t buf[4], int *btype, uint32_t *len);

/* erasebytes.c:109-134: copy 1024 bytes, AND every later whole word with the mask, trailing bytes dropped
 * exactly like the fread(4-byte items) loop does.  Returns bytes written to out. */
int64_t orc_erasebytes(const uint8_t *file, uint64_t fsz, int bits, uint8_t *out);

/* Upper bound for the container of a file of fsz bytes cut in chk-word chunks. */
size_t orc_compress_bound(uint64_t fsz, uint32_t chk);

/* workers.c:690-881 run_compress + zip.c:164-196 mzlib_def + common.c:137-149:
 * whole-file, in-memory.  One persistent raw-deflate stream per plane
 * (deflateInit2(6, Z_DEFLATED, -15, 9, Z_RLE), zip.c:112-114; constant.h:22-24), one
 * deflate(Z_FULL_FLUSH) per chunk with avail_out = chk (zip.c:170-174), RAW rule
 * inlen > len + 4 (zip.c:177).  reset_per_chunk != 0 re-inits the deflater for every chunk
 * (avoids the state-leak of SURVEY 7.7; output is identical on well-behaved data).
 * Returns container bytes, or <0 on error. */
int64_t orc_compress(const uint8_t *file, uint64_t fsz, int bits, uint32_t chk,
                     uint8_t *out, size_t out_cap, int reset_per_chunk);

/* workers.c:568-688 run_uncompress + workers.c:52-80 + zip.c:262-284 + common.c:117-135.
 * Persistent inflateInit2(-15) stream per plane, inflate(Z_FINISH) per chunk.
 * Returns bytes written ((fsz/4)*4), or <0 on error. */
int64_t orc_decompress(const uint8_t *container, size_t n, uint8_t *out, size_t out_cap);

/* Same container walk, but every COMPRESSED payload is inflated independently by
 * orc_inflate_raw (no zlib involved).  Used to show payloads are self-contained. */
int64_t orc_decompress_noz(const uint8_t *container, size_t n, uint8_t *out, size_t out_cap);

/* Independent plain-C raw-inflate (RFC 1951: stored / fixed / dynamic blocks, 32 KiB window,
 * BFINAL honoured) used to cross-check streams without zlib.  Stops at BFINAL, at end of
 * input on a block boundary, or when out_cap bytes were produced.
 * Returns 0 ok, <0 malformed.  *out_n = bytes produced, *consumed = input bytes consumed. */
int orc_inflate_raw(const uint8_t *in, size_t in_n, uint8_t *out, size_t out_cap,
                    size_t *out_n, size_t *consumed);

/* Parse the container: fills (up to max_streams) per-(chunk,plane) records.  Returns stream count or <0. */
typedef struct {
    uint64_t offset;  /* payload offset in the container */
    uint32_t len;     /* payload bytes */
    uint32_t raw;     /* 1 = RAW */
    uint32_t n;       /* plane bytes this stream inflates to */
} orc_stream_t;
int64_t orc_parse_container(const uint8_t *container, size_t n, uint64_t *fsz, uint32_t *chk,
                            orc_stream_t *streams, size_t max_streams);

const char *orc_zlib_version(void);

#ifdef __cplusplus
}
#endif
#endif
