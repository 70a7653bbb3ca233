// fz_common.cuh -- shared constants and small helpers of the B200 float-zip codec.
//
// Everything in the *.cuh codec headers is `__host__ __device__`: the CUDA kernels are the product,
// and tests/hostmodel compiles the very same source for the CPU to check the bit-level logic
// against zlib without a GPU.  No reference code is used here; the container and stream
// format follow /root/reference (citations in include/mrczip_b200.h and DESIGN.md).
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define FZ_HD __host__ __device__ __forceinline__
#define FZ_D __device__ __forceinline__
#else
#define FZ_HD inline
#endif

// ---- container constants (reference: constant.h:22-27, mrczip.h:116-121, common.c:137-149)
#define FZ_SM_COUNT 148        // B200: persistent-style grids are sized in multiples of this
#define FZ_PLANES 4
#define FZ_REF_CHUNK_WORDS (6u * 1048576u)
#define FZ_FILE_HEADER_BYTES 17
#define FZ_CHUNK_HEADER_BYTES 16
#define FZ_MRC_HEADER_WORDS 256
#define FZ_RAW_FLAG 0x80000000u

// ---- sub-block geometry of the GPU deflate encoder
// Every (chunk, plane) stream is cut in independent sub-blocks of FZ_SUB bytes; each sub-block is one
// deflate block (dynamic Huffman, distance-1 matches only -- the reference's Z_RLE strategy) or one
// stored block, followed by an empty stored block (`00 00 FF FF` after byte alignment, exactly zlib's
// sync-flush marker).  The marker makes sub-blocks concatenate at byte granularity and is what the
// GPU inflater searches for to decode a stream with one thread per sub-block.
#ifndef FZ_SUB_LOG2
#define FZ_SUB_LOG2 14
#endif
#define FZ_SUB (1u << FZ_SUB_LOG2)        // 16 KiB (8 KiB: +5 % speed at 4 GiB, 2x decode speed on small inputs, +0.5..1.5 % size)
#define FZ_SLOT_STRIDE (FZ_SUB + 32u)     // scratch bytes reserved per encoded sub-block
#define FZ_GROUP_SUBS 32u                 // sub-blocks one warp of the inflater decodes (and the unit in which all-zero sub-blocks share a fragment)
#define FZ_CODE_SUBS 128u                 // sub-blocks that share one Huffman code: 4 groups = one CTA of the inflater, one lookup table
// A stored sub-block is written as TWO stored blocks (5-byte headers) + the empty stored block (5 bytes): the
// split point is chosen so that the raw bytes can never show the sync marker 00 00 FF FF to the inflater's
// marker scan (see fz_stored_split).  One-byte sub-blocks use a single stored block.
#define FZ_STORED_OVERHEAD 15u
#define FZ_MARKER_LE 0xFFFF0000u          // bytes 00 00 FF FF read as a little-endian uint32
// A sub-block (and a group) is coded only if that saves at least 1 / 2^FZ_MIN_GAIN_SHIFT of its bytes (3 %): nearly
// incompressible bytes decode at one symbol per table lookup, a third of the speed of ordinary planes and a tenth of a
// stored block's copy, for a gain nobody would miss.
#ifndef FZ_MIN_GAIN_SHIFT
#define FZ_MIN_GAIN_SHIFT 5
#endif
#define FZ_SIZE_STORED_FLAG 0x80000000u   // in the per-sub-block size word: emit as stored block
#define FZ_SIZE_ZERO_FLAG 0x40000000u     // the sub-block is 16 KiB of zero bytes (found by the histogram kernel)
#define FZ_SIZE_COPY_FLAG 0x20000000u     // its fragment is byte-identical to the one of sub-block (bits 24..28) of its group
#define FZ_SIZE_MASK 0x00FFFFFFu          // the size itself

// Longest literal/length code the encoder gives a symbol its histogram sample SAW.  Deflate allows 15; 12 is the index
// width of the group inflater's lookup table (FZ_GLUT_BITS), so every such symbol is one table hit -- the canonical
// search for longer codes ran with 4 of 32 lanes active and was 15 % of the inflater's instructions on planes with 30+
// symbols.  Symbols the sample did not see keep the 15-bit limit (fz_ph_lengths, two tiers).
#ifndef FZ_MAX_CODE_BITS
#define FZ_MAX_CODE_BITS 12
#endif
#define FZ_MAX_MATCH 258
#define FZ_MIN_MATCH 3
#define FZ_NUM_LL 286
#define FZ_NUM_D 30
#define FZ_NUM_CL 19
#define FZ_EOB 256

// error codes of the C ABI (0 = ok, like the reference's run_* return value)
#define FZ_OK 0
#define FZ_E_ARG (-1)
#define FZ_E_CUDA (-2)
#define FZ_E_NOMEM (-3)
#define FZ_E_FORMAT (-4)   // malformed container / stream
#define FZ_E_SPACE (-5)    // output buffer too small
#define FZ_E_IO (-6)

FZ_HD uint32_t fz_mask_for_bits(int bits)
{
    // reference workers.c:29-37: 0xFFFFFFFF << b for b in 0..31, 0 for b == 32
    return bits >= 32 ? 0u : (0xFFFFFFFFu << bits);
}

FZ_HD uint32_t fz_bitrev(uint32_t v, int nbits)
{
#if defined(__CUDA_ARCH__)
    return __brev(v) >> (32 - nbits);
#else
    uint32_t r = 0;
    for (int i = 0; i < nbits; i++) { r = (r << 1) | (v & 1); v >>= 1; }
    return r;
#endif
}

FZ_HD int fz_ilog2(uint32_t v)  // floor(log2(v)), v > 0
{
#if defined(__CUDA_ARCH__)
    return 31 - __clz(v);
#else
    return 31 - __builtin_clz(v);
#endif
}

// length 3..258 -> (length code index 0..28, number of extra bits, extra value)   (RFC 1951 3.2.5)
FZ_HD void fz_len_code(uint32_t len, uint32_t &lc, uint32_t &eb, uint32_t &ev)
{
    uint32_t x = len - 3;
    if (x < 8) { lc = x; eb = 0; ev = 0; return; }
    if (x == 255) { lc = 28; eb = 0; ev = 0; return; }
    eb = (uint32_t)fz_ilog2(x) - 2;
    lc = 4 * eb + 4 + ((x >> eb) & 3);
    ev = x & ((1u << eb) - 1);
}

FZ_HD uint32_t fz_len_extra_bits(uint32_t lc)  // extra bits of length code index lc (0..28)
{
    return (lc < 8 || lc == 28) ? 0u : ((lc - 4) >> 2);
}

FZ_HD uint32_t fz_len_base(uint32_t lc)  // base length of length code index lc
{
    if (lc < 8) return lc + 3;
    if (lc == 28) return 258;
    uint32_t eb = (lc - 4) >> 2;
    return 3 + ((4 + ((lc - 4) & 3)) << eb);
}

FZ_HD uint32_t fz_dist_extra_bits(uint32_t dc)  // distance code 0..29
{
    return dc < 4 ? 0u : ((dc - 2) >> 1);
}

FZ_HD uint32_t fz_dist_base(uint32_t dc)
{
    if (dc < 4) return dc + 1;
    uint32_t eb = (dc - 2) >> 1;
    return 1 + ((2 + (dc & 1)) << eb);
}

FZ_HD uint32_t fz_stored_size(uint32_t n) { return n >= 2 ? n + FZ_STORED_OVERHEAD : n + 10u; }

// Split point s (1 <= s <= n-1) of a stored sub-block of n >= 2 bytes.  `first` = offset of the first
// 00 00 FF FF inside the data, or >= n if there is none.  Splitting inside the four bytes breaks the
// pattern (a 5-byte block header lands in the middle); LEN values of 255 are avoided because
// `.. 00 | FF 00 | 00 FF | FF ..` (LEN = 0x00FF followed by a data byte FF) would spell the marker.
FZ_HD uint32_t fz_stored_split(uint32_t n, uint32_t first)
{
    if (first < n) {
        for (uint32_t k = 0; k < 3; k++) {
            const uint32_t s = first + (k == 0 ? 2u : (k == 1 ? 1u : 3u));
            if (s >= 1 && s <= n - 1 && s != 255 && n - s != 255) return s;
        }
    }
    uint32_t s = n / 2;
    while (s == 255 || n - s == 255) s++;  // at most two steps; n >= 2 keeps 1 <= s <= n - 1 (n - s == 255 needs n >= 510)
    return s;
}
