/*
 * mrczip_oracle.c -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY
 * (see mrczip_oracle.h).  Citations are file:line under /root/reference.
 */
#include "mrczip_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <zlib.h>

/* ---------------------------------------------------------------- mask / split / merge */

uint32_t orc_mask_for_bits(int bits)
{
    /* workers.c:29-37: a 33-entry table of 0xFFFFFFFF << b with table[32] = 0 */
    if (bits < 0 || bits > 32) return 0;
    return bits == 32 ? 0u : (0xFFFFFFFFu << bits);
}

void orc_apply_mask(uint32_t *words, int64_t num, int bits, int is_first_chunk)
{
    /* workers.c:82-101: startIndex += 256 on the first chunk, then AND up to num */
    const uint32_t m = orc_mask_for_bits(bits);
    int64_t i = is_first_chunk ? ORC_MRC_HEADER_WORDS : 0;
    for (; i < num; i++) words[i] &= m;
}

void orc_split(uint32_t *words, int64_t num, uint8_t *planes[ORC_PLANES], int bits, int is_first_chunk)
{
    /* workers.c:180-203: mask first (header words exempt), then scatter ALL words' bytes */
    orc_apply_mask(words, num, bits, is_first_chunk);
    const uint8_t *p = (const uint8_t *)words;
    for (int64_t i = 0; i < num; i++)
        for (int j = 0; j < ORC_PLANES; j++) planes[j][i] = p[4 * i + j];
}

void orc_merge(uint32_t *words, int64_t num, uint8_t *const planes[ORC_PLANES])
{
    /* workers.c:423-442 */
    uint8_t *p = (uint8_t *)words;
    for (int64_t i = 0; i < num; i++)
        for (int j = 0; j < ORC_PLANES; j++) p[4 * i + j] = planes[j][i];
}

/* ---------------------------------------------------------------- headers */

void orc_pack_header(uint8_t buf[4], int btype, uint32_t len)
{
    /* zip.c:381-391 */
    buf[0] = len & 0xFF;
    buf[1] = (len >> 8) & 0xFF;
    buf[2] = (len >> 16) & 0xFF;
    buf[3] = (len >> 24) & 0xFF;
    buf[3] |= (uint8_t)(btype << 7);
}

void orc_unpack_header(const uint8_t buf[4], int *btype, uint32_t *len)
{
    /* zip.c:394-399 */
    *btype = (buf[3] & 0x80) >> 7;
    *len = (uint32_t)buf[0] | ((uint32_t)buf[1] << 8) | ((uint32_t)buf[2] << 16) | ((uint32_t)(buf[3] & 0x7f) << 24);
}

int64_t orc_erasebytes(const uint8_t *file, uint64_t fsz, int bits, uint8_t *out)
{
    /* erasebytes.c:109-134 */
    const uint32_t m = orc_mask_for_bits(bits);
    uint64_t hdr = fsz < 1024 ? fsz : 1024;
    memcpy(out, file, hdr);
    if (fsz <= 1024) return (int64_t)hdr;
    uint64_t nw = (fsz - 1024) / 4;
    for (uint64_t i = 0; i < nw; i++) {
        uint32_t w;
        memcpy(&w, file + 1024 + 4 * i, 4);
        w &= m;
        memcpy(out + 1024 + 4 * i, &w, 4);
    }
    return (int64_t)(1024 + 4 * nw);
}

size_t orc_compress_bound(uint64_t fsz, uint32_t chk)
{
    uint64_t words = fsz / 4;
    uint64_t chunks = chk ? (words + chk - 1) / chk : 0;
    return ORC_FILE_HEADER_BYTES + chunks * 16 + words * 4 + 64;
}

/* ---------------------------------------------------------------- compress */

static int def_init(z_stream *s)
{
    /* zip.c:89-123 (_mzlib_init + zlib_def_init): level 6, raw deflate (-15), memLevel 9, Z_RLE */
    memset(s, 0, sizeof(*s));
    s->data_type = Z_BINARY;
    return deflateInit2(s, 6, Z_DEFLATED, -15, 9, Z_RLE);
}

int64_t orc_compress(const uint8_t *file, uint64_t fsz, int bits, uint32_t chk,
                     uint8_t *out, size_t out_cap, int reset_per_chunk)
{
    if (bits < 0 || bits > 32 || chk == 0 || chk >= 0x80000000u) return -2; /* zip.c:325-329 */
    const uint64_t words = fsz / 4; /* fread of 4-byte items drops a ragged tail, workers.c:744 */
    if (out_cap < orc_compress_bound(fsz, chk)) return -3;
    if (words == 0) return 0; /* workers.c:757-764: header only written when the first fread > 0 */

    uint32_t *buf = (uint32_t *)malloc((size_t)chk * 4);
    uint8_t *pl[ORC_PLANES], *zo[ORC_PLANES];
    z_stream zs[ORC_PLANES];
    for (int j = 0; j < ORC_PLANES; j++) {
        pl[j] = (uint8_t *)malloc((size_t)chk + 4);
        zo[j] = (uint8_t *)malloc((size_t)chk + 4);
        if (def_init(&zs[j]) != Z_OK) return -4;
    }

    uint8_t *o = out;
    /* common.c:137-149: u64 fsz, u32 chk, u8 type (0), u8 ztypes[4] (ZLIB_DEF = 0) */
    memcpy(o, &fsz, 8); o += 8;
    memcpy(o, &chk, 4); o += 4;
    memset(o, 0, 5); o += 5;

    int is_first = 1;
    for (uint64_t w0 = 0; w0 < words; w0 += chk) {
        uint32_t num = (uint32_t)((words - w0) < chk ? (words - w0) : chk);
        memcpy(buf, file + 4 * w0, (size_t)num * 4);
        orc_split(buf, num, pl, bits, is_first); /* workers.c:791 */
        is_first = 0;                            /* workers.c:804 */
        uint8_t *hdr = o;
        o += 16;
        for (int j = 0; j < ORC_PLANES; j++) {
            if (reset_per_chunk) { deflateEnd(&zs[j]); def_init(&zs[j]); }
            /* zip.c:164-176 */
            zs[j].next_in = pl[j];
            zs[j].avail_in = num;
            zs[j].next_out = zo[j];
            zs[j].avail_out = chk;
            deflate(&zs[j], Z_FULL_FLUSH);
            uint32_t len = chk - zs[j].avail_out;
            if (num > len + ORC_HDR_SIZE) { /* zip.c:177 compressible */
                orc_pack_header(hdr + 4 * j, 0, len);
                memcpy(o, zo[j], len);
                o += len;
            } else {
                orc_pack_header(hdr + 4 * j, 1, num);
                memcpy(o, pl[j], num);
                o += num;
            }
        }
    }
    for (int j = 0; j < ORC_PLANES; j++) { deflateEnd(&zs[j]); free(pl[j]); free(zo[j]); }
    free(buf);
    return (int64_t)(o - out);
}

/* ---------------------------------------------------------------- container walk */

int64_t orc_parse_container(const uint8_t *c, size_t n, uint64_t *fsz, uint32_t *chk,
                            orc_stream_t *streams, size_t max_streams)
{
    if (n < ORC_FILE_HEADER_BYTES) return -1; /* common.c:119-123 short read */
    memcpy(fsz, c, 8);
    memcpy(chk, c + 8, 4);
    if (*chk == 0 || *chk >= 0x80000000u) return -2;
    uint64_t words = *fsz / 4; /* workers.c:577 */
    uint64_t off = ORC_FILE_HEADER_BYTES;
    int64_t ns = 0;
    for (uint64_t w0 = 0; w0 < words; w0 += *chk) {
        uint32_t num = (uint32_t)((words - w0) < *chk ? (words - w0) : *chk);
        if (off + 16 > n) return -3;
        const uint8_t *hdr = c + off;
        off += 16;
        for (int j = 0; j < ORC_PLANES; j++) {
            int bt; uint32_t len;
            orc_unpack_header(hdr + 4 * j, &bt, &len); /* workers.c:66 */
            if (off + len > n) return -3;
            if ((size_t)ns < max_streams) {
                streams[ns].offset = off; streams[ns].len = len; streams[ns].raw = (uint32_t)bt; streams[ns].n = num;
            }
            ns++;
            off += len;
        }
    }
    return ns;
}

static int64_t decompress_impl(const uint8_t *c, size_t n, uint8_t *out, size_t out_cap, int use_zlib)
{
    uint64_t fsz; uint32_t chk;
    int64_t ns = orc_parse_container(c, n, &fsz, &chk, NULL, 0);
    if (ns < 0) return ns;
    uint64_t words = fsz / 4;
    if (out_cap < words * 4) return -5;
    if (ns == 0) return 0;
    orc_stream_t *st = (orc_stream_t *)malloc(sizeof(orc_stream_t) * (size_t)ns);
    orc_parse_container(c, n, &fsz, &chk, st, (size_t)ns);

    z_stream zs[ORC_PLANES];
    uint8_t *pl[ORC_PLANES];
    for (int j = 0; j < ORC_PLANES; j++) {
        memset(&zs[j], 0, sizeof(z_stream));
        if (use_zlib && inflateInit2(&zs[j], -15) != Z_OK) return -4; /* zip.c:140-160 */
        pl[j] = (uint8_t *)malloc((size_t)chk + 4);
    }
    int64_t rc = (int64_t)(words * 4);
    uint64_t w0 = 0;
    for (int64_t s = 0; s < ns; s += ORC_PLANES, w0 += chk) {
        uint32_t num = st[s].n;
        uint8_t *src[ORC_PLANES];
        for (int j = 0; j < ORC_PLANES; j++) {
            const orc_stream_t *r = &st[s + j];
            if (r->raw) { src[j] = (uint8_t *)(c + r->offset); continue; } /* zip.c:264-267 */
            if (use_zlib) {
                /* zip.c:269-277: persistent stream, inflate(Z_FINISH), return code ignored */
                zs[j].next_in = (Bytef *)(c + r->offset);
                zs[j].avail_in = r->len;
                zs[j].next_out = pl[j];
                zs[j].avail_out = chk;
                inflate(&zs[j], Z_FINISH);
                if (chk - zs[j].avail_out != num) rc = -6;
            } else {
                size_t on = 0, used = 0;
                if (orc_inflate_raw(c + r->offset, r->len, pl[j], num, &on, &used) != 0 || on != num) rc = -6;
            }
            src[j] = pl[j];
        }
        orc_merge((uint32_t *)(out + 4 * w0), num, src); /* workers.c:618 */
    }
    for (int j = 0; j < ORC_PLANES; j++) { if (use_zlib) inflateEnd(&zs[j]); free(pl[j]); }
    free(st);
    return rc;
}

int64_t orc_decompress(const uint8_t *c, size_t n, uint8_t *out, size_t out_cap)
{
    return decompress_impl(c, n, out, out_cap, 1);
}

int64_t orc_decompress_noz(const uint8_t *c, size_t n, uint8_t *out, size_t out_cap)
{
    return decompress_impl(c, n, out, out_cap, 0);
}

const char *orc_zlib_version(void) { return zlibVersion(); }

/* ---------------------------------------------------------------- independent raw inflate (RFC 1951) */

typedef struct {
    const uint8_t *in; size_t in_n, pos;
    uint64_t acc; int nacc;
} bitrd_t;

static int need(bitrd_t *b, int k)
{
    while (b->nacc < k) {
        if (b->pos >= b->in_n) return -1;
        b->acc |= (uint64_t)b->in[b->pos++] << b->nacc;
        b->nacc += 8;
    }
    return 0;
}
static int getbits(bitrd_t *b, int k, uint32_t *v)
{
    if (k == 0) { *v = 0; return 0; }
    if (need(b, k)) return -1;
    *v = (uint32_t)(b->acc & ((1ull << k) - 1));
    b->acc >>= k; b->nacc -= k;
    return 0;
}

typedef struct { uint16_t count[16]; uint16_t sym[288]; } canon_t;

/* canonical code from lengths; returns 0 complete, >0 incomplete, <0 over-subscribed */
static int canon_build(canon_t *h, const uint8_t *len, int n)
{
    uint16_t offs[16];
    memset(h->count, 0, sizeof(h->count));
    for (int i = 0; i < n; i++) h->count[len[i]]++;
    if (h->count[0] == n) return 0; /* no codes */
    int left = 1;
    for (int l = 1; l <= 15; l++) { left <<= 1; left -= h->count[l]; if (left < 0) return left; }
    offs[1] = 0;
    for (int l = 1; l < 15; l++) offs[l + 1] = offs[l] + h->count[l];
    for (int i = 0; i < n; i++) if (len[i]) h->sym[offs[len[i]]++] = (uint16_t)i;
    return left;
}

static int canon_decode(bitrd_t *b, const canon_t *h)
{
    int code = 0, first = 0, index = 0;
    for (int l = 1; l <= 15; l++) {
        uint32_t bit;
        if (getbits(b, 1, &bit)) return -1;
        code |= (int)bit;
        int cnt = h->count[l];
        if (code - cnt < first) return h->sym[index + (code - first)];
        index += cnt; first += cnt; first <<= 1; code <<= 1;
    }
    return -2;
}

static const uint16_t kLenBase[29] = {3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,115,131,163,195,227,258};
static const uint8_t kLenExtra[29] = {0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0};
static const uint16_t kDistBase[30] = {1,2,3,4,5,7,9,13,17,25,33,49,65,97,129,193,257,385,513,769,1025,1537,2049,3073,4097,6145,8193,12289,16385,24577};
static const uint8_t kDistExtra[30] = {0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13};
static const uint8_t kClOrder[19] = {16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15};

int orc_inflate_raw(const uint8_t *in, size_t in_n, uint8_t *out, size_t out_cap,
                    size_t *out_n, size_t *consumed)
{
    bitrd_t b = { in, in_n, 0, 0, 0 };
    size_t op = 0;
    int rc = 0, last = 0;
    canon_t *ll = (canon_t *)malloc(sizeof(canon_t)), *dd = (canon_t *)malloc(sizeof(canon_t));
    while (!last) {
        /* stop cleanly at end of input on a block boundary (the reference never sets BFINAL) */
        if (b.pos >= b.in_n && b.nacc < 3) break;
        uint32_t v, type;
        if (getbits(&b, 1, &v)) { rc = -1; break; }
        last = (int)v;
        if (getbits(&b, 2, &type)) { rc = -1; break; }
        if (type == 0) {
            b.acc >>= (b.nacc & 7); b.nacc -= (b.nacc & 7);
            uint32_t len, nlen;
            if (getbits(&b, 16, &len) || getbits(&b, 16, &nlen)) { rc = -1; break; }
            if ((len ^ 0xFFFF) != nlen) { rc = -2; break; }
            for (uint32_t i = 0; i < len; i++) {
                uint32_t c;
                if (getbits(&b, 8, &c)) { rc = -1; break; }
                if (op >= out_cap) { rc = -3; break; }
                out[op++] = (uint8_t)c;
            }
            if (rc) break;
            continue;
        }
        if (type == 3) { rc = -2; break; }
        uint8_t lens[320];
        int nlen_ = 288, ndist = 30;
        if (type == 1) {
            int i = 0;
            for (; i < 144; i++) lens[i] = 8;
            for (; i < 256; i++) lens[i] = 9;
            for (; i < 280; i++) lens[i] = 7;
            for (; i < 288; i++) lens[i] = 8;
            canon_build(ll, lens, 288);
            for (i = 0; i < 30; i++) lens[i] = 5;
            canon_build(dd, lens, 30);
        } else {
            uint32_t hlit, hdist, hclen;
            if (getbits(&b, 5, &hlit) || getbits(&b, 5, &hdist) || getbits(&b, 4, &hclen)) { rc = -1; break; }
            nlen_ = (int)hlit + 257; ndist = (int)hdist + 1; int ncode = (int)hclen + 4;
            if (nlen_ > 286 || ndist > 30) { rc = -2; break; }
            uint8_t cl[19];
            memset(cl, 0, sizeof(cl));
            for (int i = 0; i < ncode; i++) { uint32_t x; if (getbits(&b, 3, &x)) { rc = -1; break; } cl[kClOrder[i]] = (uint8_t)x; }
            if (rc) break;
            canon_t clh;
            if (canon_build(&clh, cl, 19) != 0) { rc = -2; break; }
            int idx = 0;
            while (idx < nlen_ + ndist) {
                int sym = canon_decode(&b, &clh);
                if (sym < 0) { rc = -1; break; }
                if (sym < 16) { lens[idx++] = (uint8_t)sym; continue; }
                uint32_t rep; uint8_t val = 0;
                if (sym == 16) { if (idx == 0) { rc = -2; break; } val = lens[idx - 1]; if (getbits(&b, 2, &rep)) { rc = -1; break; } rep += 3; }
                else if (sym == 17) { if (getbits(&b, 3, &rep)) { rc = -1; break; } rep += 3; }
                else { if (getbits(&b, 7, &rep)) { rc = -1; break; } rep += 11; }
                if (idx + (int)rep > nlen_ + ndist) { rc = -2; break; }
                while (rep--) lens[idx++] = val;
            }
            if (rc) break;
            if (lens[256] == 0) { rc = -2; break; }
            int e = canon_build(ll, lens, nlen_);
            if (e < 0 || (e > 0 && nlen_ - ll->count[0] != 1)) { rc = -2; break; }
            e = canon_build(dd, lens + nlen_, ndist);
            if (e < 0 || (e > 0 && ndist - dd->count[0] != 1)) { rc = -2; break; }
        }
        for (;;) {
            int sym = canon_decode(&b, ll);
            if (sym < 0) { rc = -1; break; }
            if (sym < 256) { if (op >= out_cap) { rc = -3; break; } out[op++] = (uint8_t)sym; continue; }
            if (sym == 256) break;
            sym -= 257;
            if (sym >= 29) { rc = -2; break; }
            uint32_t eb;
            if (getbits(&b, kLenExtra[sym], &eb)) { rc = -1; break; }
            uint32_t len = kLenBase[sym] + eb;
            int ds = canon_decode(&b, dd);
            if (ds < 0 || ds >= 30) { rc = -1; break; }
            if (getbits(&b, kDistExtra[ds], &eb)) { rc = -1; break; }
            size_t dist = kDistBase[ds] + eb;
            if (dist > op) { rc = -4; break; }
            if (op + len > out_cap) { rc = -3; break; }
            for (uint32_t i = 0; i < len; i++) { out[op] = out[op - dist]; op++; }
        }
        if (rc) break;
    }
    free(ll); free(dd);
    *out_n = op;
    *consumed = b.pos - (size_t)(b.nacc / 8);
    return rc;
}
