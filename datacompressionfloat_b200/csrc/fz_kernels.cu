// fz_kernels.cu -- hand-written sm_100a kernels of the float-zip hot path.
//
//   split / merge      reference workers.c:82-101,180-203 / 423-442   (HBM-bound, 8 B per word)
//   encode             reference zip.c:164-196 (mzlib_def -> zlib deflate, Z_RLE)   one warp per 16 KiB sub-block
//   layout + gather    reference workers.c:837-850 + zip.c:177-190,381-391          device scan of payload sizes
//   walk               reference workers.c:52-69 (chunk header chain)
//   marker scan, inflate (fast: one thread per sub-block; general: one thread per stream)
//                      reference zip.c:262-284 (mzlib_inf -> zlib inflate)
//
// No tensor cores: nothing here is a contraction.  Integer / byte work only.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "fz_deflate_enc.cuh"
#include "fz_enc2.cuh"
#include "fz_inflate.cuh"
#include "fz_kernels.h"

#define FZ_WARP 32

// words of chunk c of the batch: uniform chunks with a ragged last one, or -- several small files in one batch -- a table
__device__ __forceinline__ uint32_t fz_chunk_n(const FzBatchGeom &g, uint32_t c)
{
    return g.chunk_n ? g.chunk_n[c] : (c == g.nchunks - 1 ? g.last_n : g.chk);
}

// =================================================================================================
// helpers
// =================================================================================================
__device__ __forceinline__ uint4 fz_ld_stream(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t fz_ld_stream32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// 4x4 byte transpose: words (AoS) <-> plane bytes (SoA); the network is its own inverse
__device__ __forceinline__ void fz_transpose4(uint32_t a, uint32_t b, uint32_t c, uint32_t d,
                                              uint32_t &o0, uint32_t &o1, uint32_t &o2, uint32_t &o3)
{
    const uint32_t t0 = __byte_perm(a, b, 0x5140), t1 = __byte_perm(c, d, 0x5140);
    const uint32_t t2 = __byte_perm(a, b, 0x7362), t3 = __byte_perm(c, d, 0x7362);
    o0 = __byte_perm(t0, t1, 0x5410);
    o1 = __byte_perm(t0, t1, 0x7632);
    o2 = __byte_perm(t2, t3, 0x5410);
    o3 = __byte_perm(t2, t3, 0x7632);
}

// bytes [sh, sh+16) of the 32-byte pair A||B  (sh in 0..15, warp-uniform)
__device__ __forceinline__ uint4 fz_funnel16(uint4 A, uint4 B, uint32_t sh)
{
    const uint32_t r = (sh & 3) * 8;
    uint32_t w0, w1, w2, w3, w4;
    switch (sh >> 2) {
        case 0: w0 = A.x; w1 = A.y; w2 = A.z; w3 = A.w; w4 = B.x; break;
        case 1: w0 = A.y; w1 = A.z; w2 = A.w; w3 = B.x; w4 = B.y; break;
        case 2: w0 = A.z; w1 = A.w; w2 = B.x; w3 = B.y; w4 = B.z; break;
        default: w0 = A.w; w1 = B.x; w2 = B.y; w3 = B.z; w4 = B.w; break;
    }
    uint4 o;
    o.x = __funnelshift_r(w0, w1, r);
    o.y = __funnelshift_r(w1, w2, r);
    o.z = __funnelshift_r(w2, w3, r);
    o.w = __funnelshift_r(w3, w4, r);
    return o;
}

// Warp-cooperative copy of n bytes between arbitrarily aligned global addresses: 16-byte aligned stores,
// source re-aligned with a funnel shift.  May read up to 31 bytes past src + n inside the same
// allocation (all our buffers carry that slack); `src_end` clamps reads for foreign buffers.
__device__ __forceinline__ void fz_warp_copy(uint8_t *dst, const uint8_t *src, uint32_t n, const uint8_t *src_end, int lane)
{
    uint32_t head = (16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u;
    if (head > n) head = n;
    if ((uint32_t)lane < head) dst[lane] = src[lane];
    dst += head; src += head; n -= head;
    const uint32_t nvec = n >> 4;
    const uint32_t sh = (uint32_t)((uintptr_t)src & 15u);
    const uint4 *s0 = (const uint4 *)(src - sh);
    const uint4 *send = (const uint4 *)(((uintptr_t)src_end + 15u) & ~(uintptr_t)15u);
    for (uint32_t v = lane; v < nvec; v += FZ_WARP) {
        const uint4 A = s0[v];
        uint4 B = A;
        if (sh && (s0 + v + 1) < send) B = s0[v + 1];
        ((uint4 *)dst)[v] = sh ? fz_funnel16(A, B, sh) : A;
    }
    const uint32_t tail = n & 15u;
    if ((uint32_t)lane < tail) dst[(nvec << 4) + lane] = src[(nvec << 4) + lane];
}

// =================================================================================================
// mask + byte-plane split      words[i] -> planes[j][i] = byte j of (words[i] & mask)   (i >= exempt)
// =================================================================================================
#define FZ_SPLIT_THREADS 256
#define FZ_SPLIT_UNROLL 4

// variant 0: warp-coalesced 128-bit loads, four 32-bit stores (one per plane) per load; every warp
// store instruction covers one full 128-byte line.
__global__ void __launch_bounds__(FZ_SPLIT_THREADS)
fz_split_kernel_v0(const uint4 *__restrict__ words4, uint64_t nvec, uint32_t mask, uint64_t exempt,
                   uint8_t *__restrict__ planes, uint64_t plane_stride, uint32_t skip_planes, uint64_t skip_lo, uint64_t skip_hi)
{
    const uint64_t base = (uint64_t)blockIdx.x * (FZ_SPLIT_THREADS * FZ_SPLIT_UNROLL) + threadIdx.x;
    uint4 v[FZ_SPLIT_UNROLL];
#pragma unroll
    for (int k = 0; k < FZ_SPLIT_UNROLL; k++) {
        const uint64_t i = base + (uint64_t)k * FZ_SPLIT_THREADS;
        if (i < nvec) v[k] = fz_ld_stream(words4 + i);
    }
    uint32_t *p0 = (uint32_t *)planes, *p1 = (uint32_t *)(planes + plane_stride);
    uint32_t *p2 = (uint32_t *)(planes + 2 * plane_stride), *p3 = (uint32_t *)(planes + 3 * plane_stride);
#pragma unroll
    for (int k = 0; k < FZ_SPLIT_UNROLL; k++) {
        const uint64_t i = base + (uint64_t)k * FZ_SPLIT_THREADS;
        if (i < nvec) {
            uint4 w = v[k];
            const uint64_t wi = i * 4;
            if (wi >= exempt) { w.x &= mask; w.y &= mask; w.z &= mask; w.w &= mask; }
            else {
                if (wi + 0 >= exempt) w.x &= mask;
                if (wi + 1 >= exempt) w.y &= mask;
                if (wi + 2 >= exempt) w.z &= mask;
                if (wi + 3 >= exempt) w.w &= mask;
            }
            uint32_t a, b, c, d;
            fz_transpose4(w.x, w.y, w.z, w.w, a, b, c, d);
            // planes the mask erases entirely are not written where the encoder will not read them: whole sub-blocks
            // behind the exempt header words (fz_hist2_kernel flags those all-zero without a look)
            const uint32_t sk = (wi >= skip_lo && wi < skip_hi) ? skip_planes : 0u;
            if (!(sk & 1u)) p0[i] = a;
            if (!(sk & 2u)) p1[i] = b;
            if (!(sk & 4u)) p2[i] = c;
            if (!(sk & 8u)) p3[i] = d;
        }
    }
}

// variant 1: each thread owns 16 consecutive words (four 128-bit loads), one 128-bit store per plane
__global__ void __launch_bounds__(FZ_SPLIT_THREADS)
fz_split_kernel_v1(const uint4 *__restrict__ words4, uint64_t nvec, uint32_t mask, uint64_t exempt,
                   uint8_t *__restrict__ planes, uint64_t plane_stride)
{
    const uint64_t t = (uint64_t)blockIdx.x * FZ_SPLIT_THREADS + threadIdx.x;  // group of 4 vectors
    const uint64_t i0 = t * 4;
    if (i0 >= nvec) return;
    if (i0 + 4 <= nvec) {
        uint4 v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = words4[i0 + k];
        uint32_t o[4][4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint4 w = v[k];
            const uint64_t wi = (i0 + k) * 4;
            if (wi + 0 >= exempt) w.x &= mask;
            if (wi + 1 >= exempt) w.y &= mask;
            if (wi + 2 >= exempt) w.z &= mask;
            if (wi + 3 >= exempt) w.w &= mask;
            fz_transpose4(w.x, w.y, w.z, w.w, o[0][k], o[1][k], o[2][k], o[3][k]);
        }
#pragma unroll
        for (int j = 0; j < 4; j++)
            ((uint4 *)(planes + j * plane_stride))[t] = make_uint4(o[j][0], o[j][1], o[j][2], o[j][3]);
    } else {
        for (uint64_t i = i0; i < nvec; i++) {
            uint4 w = words4[i];
            const uint64_t wi = i * 4;
            if (wi + 0 >= exempt) w.x &= mask;
            if (wi + 1 >= exempt) w.y &= mask;
            if (wi + 2 >= exempt) w.z &= mask;
            if (wi + 3 >= exempt) w.w &= mask;
            uint32_t a, b, c, d;
            fz_transpose4(w.x, w.y, w.z, w.w, a, b, c, d);
            ((uint32_t *)planes)[i] = a;
            ((uint32_t *)(planes + plane_stride))[i] = b;
            ((uint32_t *)(planes + 2 * plane_stride))[i] = c;
            ((uint32_t *)(planes + 3 * plane_stride))[i] = d;
        }
    }
}

// several small files in one batch: chunk c occupies the slot [c * chk, (c + 1) * chk) of the word buffer and of every
// plane (chk % 4 == 0), its first chunk_exempt[c] words keep their bits; words behind chunk_n[c] are never looked at by
// the later stages, so the slots are simply processed whole
__global__ void __launch_bounds__(FZ_SPLIT_THREADS)
fz_split_slots_kernel(const uint4 *__restrict__ words4, uint64_t nvec, uint32_t mask, uint32_t chk, const uint32_t *__restrict__ chunk_exempt,
                      uint8_t *__restrict__ planes, uint64_t plane_stride)
{
    const uint64_t base = (uint64_t)blockIdx.x * (FZ_SPLIT_THREADS * FZ_SPLIT_UNROLL) + threadIdx.x;
    uint32_t *p0 = (uint32_t *)planes, *p1 = (uint32_t *)(planes + plane_stride);
    uint32_t *p2 = (uint32_t *)(planes + 2 * plane_stride), *p3 = (uint32_t *)(planes + 3 * plane_stride);
#pragma unroll
    for (int k = 0; k < FZ_SPLIT_UNROLL; k++) {
        const uint64_t i = base + (uint64_t)k * FZ_SPLIT_THREADS;
        if (i < nvec) {
            uint4 w = fz_ld_stream(words4 + i);
            const uint64_t wi = i * 4;
            const uint32_t c = (uint32_t)(wi / chk);
            const uint64_t local = wi - (uint64_t)c * chk, ex = chunk_exempt[c];
            if (local + 0 >= ex) w.x &= mask;
            if (local + 1 >= ex) w.y &= mask;
            if (local + 2 >= ex) w.z &= mask;
            if (local + 3 >= ex) w.w &= mask;
            uint32_t a, b, cc, d;
            fz_transpose4(w.x, w.y, w.z, w.w, a, b, cc, d);
            p0[i] = a; p1[i] = b; p2[i] = cc; p3[i] = d;
        }
    }
}

void fz_launch_split_slots(const uint32_t *words, uint32_t nchunks, uint32_t chk, uint32_t mask, const uint32_t *chunk_exempt, uint8_t *planes,
                           uint64_t plane_stride, cudaStream_t st)
{
    const uint64_t nvec = (uint64_t)nchunks * chk / 4;
    const uint64_t per = FZ_SPLIT_THREADS * FZ_SPLIT_UNROLL;
    fz_split_slots_kernel<<<(unsigned)((nvec + per - 1) / per), FZ_SPLIT_THREADS, 0, st>>>((const uint4 *)words, nvec, mask, chk, chunk_exempt, planes, plane_stride);
}

// ragged tail: the last nwords % 4 words
__global__ void fz_split_tail_kernel(const uint32_t *__restrict__ words, uint64_t first, uint64_t nwords, uint32_t mask,
                                     uint64_t exempt, uint8_t *__restrict__ planes, uint64_t plane_stride)
{
    const uint64_t i = first + threadIdx.x;
    if (i >= nwords) return;
    uint32_t w = words[i];
    if (i >= exempt) w &= mask;
    for (int j = 0; j < 4; j++) planes[j * plane_stride + i] = (uint8_t)(w >> (8 * j));
}

void fz_launch_split(const uint32_t *words, uint64_t nwords, uint32_t mask, uint64_t exempt_words, uint8_t *planes,
                     uint64_t plane_stride, int variant, cudaStream_t st, uint32_t skip_planes, uint64_t skip_lo, uint64_t skip_hi)
{
    const uint64_t nvec = nwords / 4;
    if (nvec) {
        if (variant == 1) {
            const uint64_t groups = (nvec + 3) / 4;
            const unsigned grid = (unsigned)((groups + FZ_SPLIT_THREADS - 1) / FZ_SPLIT_THREADS);
            fz_split_kernel_v1<<<grid, FZ_SPLIT_THREADS, 0, st>>>((const uint4 *)words, nvec, mask, exempt_words, planes, plane_stride);
        } else {
            const uint64_t per = FZ_SPLIT_THREADS * FZ_SPLIT_UNROLL;
            const unsigned grid = (unsigned)((nvec + per - 1) / per);
            fz_split_kernel_v0<<<grid, FZ_SPLIT_THREADS, 0, st>>>((const uint4 *)words, nvec, mask, exempt_words, planes, plane_stride,
                                                                  skip_planes, skip_lo, skip_hi);
        }
    }
    if (nwords & 3) fz_split_tail_kernel<<<1, 4, 0, st>>>(words, nvec * 4, nwords, mask, exempt_words, planes, plane_stride);
}

// =================================================================================================
// byte-plane merge       words[i] = planes[0][i] | planes[1][i] << 8 | planes[2][i] << 16 | planes[3][i] << 24
// =================================================================================================
__global__ void __launch_bounds__(FZ_SPLIT_THREADS)
fz_merge_kernel_v0(const uint8_t *__restrict__ planes, uint64_t plane_stride, uint64_t nvec, uint4 *__restrict__ words4)
{
    const uint64_t base = (uint64_t)blockIdx.x * (FZ_SPLIT_THREADS * FZ_SPLIT_UNROLL) + threadIdx.x;
    const uint32_t *p0 = (const uint32_t *)planes, *p1 = (const uint32_t *)(planes + plane_stride);
    const uint32_t *p2 = (const uint32_t *)(planes + 2 * plane_stride), *p3 = (const uint32_t *)(planes + 3 * plane_stride);
    uint32_t a[FZ_SPLIT_UNROLL], b[FZ_SPLIT_UNROLL], c[FZ_SPLIT_UNROLL], d[FZ_SPLIT_UNROLL];
#pragma unroll
    for (int k = 0; k < FZ_SPLIT_UNROLL; k++) {
        const uint64_t i = base + (uint64_t)k * FZ_SPLIT_THREADS;
        if (i < nvec) { a[k] = fz_ld_stream32(p0 + i); b[k] = fz_ld_stream32(p1 + i); c[k] = fz_ld_stream32(p2 + i); d[k] = fz_ld_stream32(p3 + i); }
    }
#pragma unroll
    for (int k = 0; k < FZ_SPLIT_UNROLL; k++) {
        const uint64_t i = base + (uint64_t)k * FZ_SPLIT_THREADS;
        if (i < nvec) {
            uint4 w;
            fz_transpose4(a[k], b[k], c[k], d[k], w.x, w.y, w.z, w.w);
            words4[i] = w;
        }
    }
}

// variant 1: one 128-bit load per plane (16 words per thread), four 128-bit stores
__global__ void __launch_bounds__(FZ_SPLIT_THREADS)
fz_merge_kernel_v1(const uint8_t *__restrict__ planes, uint64_t plane_stride, uint64_t nvec, uint4 *__restrict__ words4)
{
    const uint64_t t = (uint64_t)blockIdx.x * FZ_SPLIT_THREADS + threadIdx.x;
    const uint64_t i0 = t * 4;
    if (i0 >= nvec) return;
    if (i0 + 4 <= nvec) {
        uint4 p[4];
#pragma unroll
        for (int j = 0; j < 4; j++) p[j] = fz_ld_stream((const uint4 *)(planes + j * plane_stride) + t);
        uint4 w;
        fz_transpose4(p[0].x, p[1].x, p[2].x, p[3].x, w.x, w.y, w.z, w.w); words4[i0 + 0] = w;
        fz_transpose4(p[0].y, p[1].y, p[2].y, p[3].y, w.x, w.y, w.z, w.w); words4[i0 + 1] = w;
        fz_transpose4(p[0].z, p[1].z, p[2].z, p[3].z, w.x, w.y, w.z, w.w); words4[i0 + 2] = w;
        fz_transpose4(p[0].w, p[1].w, p[2].w, p[3].w, w.x, w.y, w.z, w.w); words4[i0 + 3] = w;
    } else {
        for (uint64_t i = i0; i < nvec; i++) {
            uint4 w;
            fz_transpose4(((const uint32_t *)planes)[i], ((const uint32_t *)(planes + plane_stride))[i],
                          ((const uint32_t *)(planes + 2 * plane_stride))[i], ((const uint32_t *)(planes + 3 * plane_stride))[i],
                          w.x, w.y, w.z, w.w);
            words4[i] = w;
        }
    }
}

__global__ void fz_merge_tail_kernel(const uint8_t *__restrict__ planes, uint64_t plane_stride, uint64_t first, uint64_t nwords,
                                     uint32_t *__restrict__ words)
{
    const uint64_t i = first + threadIdx.x;
    if (i >= nwords) return;
    uint32_t w = 0;
    for (int j = 0; j < 4; j++) w |= (uint32_t)planes[j * plane_stride + i] << (8 * j);
    words[i] = w;
}

void fz_launch_merge(const uint8_t *planes, uint64_t plane_stride, uint64_t nwords, uint32_t *words, int variant, cudaStream_t st)
{
    const uint64_t nvec = nwords / 4;
    if (nvec) {
        if (variant == 1) {
            const uint64_t groups = (nvec + 3) / 4;
            const unsigned grid = (unsigned)((groups + FZ_SPLIT_THREADS - 1) / FZ_SPLIT_THREADS);
            fz_merge_kernel_v1<<<grid, FZ_SPLIT_THREADS, 0, st>>>(planes, plane_stride, nvec, (uint4 *)words);
        } else {
            const uint64_t per = FZ_SPLIT_THREADS * FZ_SPLIT_UNROLL;
            const unsigned grid = (unsigned)((nvec + per - 1) / per);
            fz_merge_kernel_v0<<<grid, FZ_SPLIT_THREADS, 0, st>>>(planes, plane_stride, nvec, (uint4 *)words);
        }
    }
    if (nwords & 3) fz_merge_tail_kernel<<<1, 4, 0, st>>>(planes, plane_stride, nvec * 4, nwords, words);
}

// merge for the decompress pipeline: plane j of chunk c comes from the plane buffer (inflated) or, for a RAW
// stream, straight from its payload inside the container (arbitrary byte alignment: two aligned loads + funnel shift)
__device__ __forceinline__ uint32_t fz_ld_u32_unaligned(const uint8_t *src, uint64_t byte_off, const uint8_t *end)
{
    const uint8_t *p = src + byte_off;
    const uint32_t sk = (uint32_t)((uintptr_t)p & 3u);
    const uint32_t *a = (const uint32_t *)(p - sk);
    const uint32_t w0 = fz_ld_stream32(a);
    if (sk == 0) return w0;
    const uint32_t w1 = ((const uint8_t *)(a + 1) < end) ? fz_ld_stream32(a + 1) : 0u;
    return __funnelshift_r(w0, w1, sk * 8);
}

// one 256-bit store (32-byte aligned address)
__device__ __forceinline__ void fz_st256(void *p, uint4 a, uint4 b)
{
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y),
                 "r"(b.z), "r"(b.w)
                 : "memory");
}

// which chunks take the 16-words-per-thread path: those with at least wide_min RAW planes (0: all of them, the default)
__device__ __forceinline__ bool fz_merge_wide(const uint32_t *__restrict__ stream_hdr, uint32_t c, uint32_t wide_min)
{
    const uint4 h = __ldg((const uint4 *)stream_hdr + c);
    return ((h.x >> 31) + (h.y >> 31) + (h.z >> 31) + (h.w >> 31)) >= wide_min;
}

// 16 words per thread: one 128-bit load per plane -- two and a funnel shift for a RAW payload, which sits at whatever
// address the container gives it -- and two 256-bit stores.  (With four 128-bit stores this path only paid for chunks
// with two or more RAW planes; with 256-bit stores it beats the 4-bytes-per-plane path below on every input measured:
// G b=8 1.42 -> 1.20 ms, S b=12 1.41 -> 1.13, G b=16 1.27 -> 1.02, P b=0 1.03 -> 0.94 per 4 GiB.)  CTA shape: 4096 words,
// a quarter of one sub-block of every plane.
#define FZ_MERGE16_THREADS FZ_SPLIT_THREADS
__device__ __forceinline__ void
fz_merge_streams16(const uint8_t *__restrict__ planes, const uint8_t *__restrict__ container, const uint8_t *container_end,
                          const uint32_t *__restrict__ stream_hdr, const unsigned long long *__restrict__ stream_off,
                          const uint32_t *__restrict__ zero_flags, FzBatchGeom g, uint32_t *__restrict__ words)
{
    const uint32_t c = blockIdx.y;
    const uint32_t n_c = fz_chunk_n(g, c);
    const uint32_t w0 = (blockIdx.x * FZ_MERGE16_THREADS + threadIdx.x) * 16u;   // first word of this thread
    if (blockIdx.x * FZ_MERGE16_THREADS * 16u >= n_c) return;
    const uint32_t sub = (blockIdx.x * FZ_MERGE16_THREADS * 16u) >> FZ_SUB_LOG2;
    uint32_t h[4], zf[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        h[j] = __ldg(stream_hdr + c * 4 + j);
        zf[j] = zero_flags ? __ldg(zero_flags + (size_t)(c * 4 + j) * g.nsub_full + sub) : 0u;
    }
    const uint8_t *src[4];
    bool zero[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const bool raw = (h[j] & FZ_RAW_FLAG) != 0;
        src[j] = raw ? container + stream_off[c * 4 + j] : planes + (uint64_t)j * g.plane_stride + (uint64_t)c * g.chk;
        zero[j] = zf[j] != 0 && !raw;
    }
    uint32_t *out = words + (uint64_t)c * g.chk;
    if (w0 + 16u <= n_c) {
        uint4 p[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (zero[j]) { p[j] = make_uint4(0, 0, 0, 0); continue; }
            const uint8_t *a = src[j] + w0;
            const uint32_t sh = (uint32_t)((uintptr_t)a & 15u);
            if (sh == 0) p[j] = fz_ld_stream((const uint4 *)a);
            else {
                const uint4 *a0 = (const uint4 *)(a - sh);
                const uint4 A = __ldg(a0);
                const uint4 B = ((const uint8_t *)(a0 + 1) < container_end) ? __ldg(a0 + 1) : A;
                p[j] = fz_funnel16(A, B, sh);
            }
        }
        uint4 w, x;
        uint4 *o4 = (uint4 *)(out + w0);
        if (((uintptr_t)o4 & 31u) == 0) {
            // two 256-bit stores (sm_100: STG.256): every store instruction writes whole 32-byte sectors.  With four
            // 128-bit stores per thread each sector was written in two halves by two instructions: twice the write
            // requests on the way to L2 (ncu: lg_throttle was the second stall reason of this kernel).
            fz_transpose4(p[0].x, p[1].x, p[2].x, p[3].x, w.x, w.y, w.z, w.w);
            fz_transpose4(p[0].y, p[1].y, p[2].y, p[3].y, x.x, x.y, x.z, x.w);
            fz_st256(o4, w, x);
            fz_transpose4(p[0].z, p[1].z, p[2].z, p[3].z, w.x, w.y, w.z, w.w);
            fz_transpose4(p[0].w, p[1].w, p[2].w, p[3].w, x.x, x.y, x.z, x.w);
            fz_st256(o4 + 2, w, x);
        } else {
            fz_transpose4(p[0].x, p[1].x, p[2].x, p[3].x, w.x, w.y, w.z, w.w); o4[0] = w;
            fz_transpose4(p[0].y, p[1].y, p[2].y, p[3].y, w.x, w.y, w.z, w.w); o4[1] = w;
            fz_transpose4(p[0].z, p[1].z, p[2].z, p[3].z, w.x, w.y, w.z, w.w); o4[2] = w;
            fz_transpose4(p[0].w, p[1].w, p[2].w, p[3].w, w.x, w.y, w.z, w.w); o4[3] = w;
        }
    } else {
        for (uint32_t i = w0; i < n_c; i++) {   // the ragged end of a chunk
            uint32_t w = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) w |= (zero[j] ? 0u : (uint32_t)src[j][i]) << (8 * j);
            out[i] = w;
        }
    }
}

__global__ void __launch_bounds__(FZ_SPLIT_THREADS, 8)
fz_merge_streams_kernel(const uint8_t *__restrict__ planes, const uint8_t *__restrict__ container, const uint32_t *__restrict__ stream_hdr,
                        const unsigned long long *__restrict__ stream_off, const uint32_t *__restrict__ zero_flags, FzBatchGeom g,
                        uint32_t *__restrict__ words, uint32_t wide_min, const uint8_t *container_end)
{
    const uint32_t c = blockIdx.y;
    if (fz_merge_wide(stream_hdr, c, wide_min)) {   // (wide_min 0: every chunk, 5 or more: none)
        fz_merge_streams16(planes, container, container_end, stream_hdr, stream_off, zero_flags, g, words);
        return;
    }
    const uint32_t n_c = fz_chunk_n(g, c);
    const uint32_t nvec = n_c / 4;
    const uint8_t *src[4];
    const uint8_t *end[4];
    bool zero[4];
    // a CTA covers FZ_SPLIT_THREADS * FZ_SPLIT_UNROLL * 4 = 4096 plane bytes: inside one 16 KiB sub-block
    const uint32_t sub = (blockIdx.x * (FZ_SPLIT_THREADS * FZ_SPLIT_UNROLL * 4)) >> FZ_SUB_LOG2;
    // sub-blocks the inflater found to be all zero were never written to the plane buffer (mask bits >= 8 zero
    // whole byte planes: 1 GiB of stores and 1 GiB of loads per 4 GiB volume that nobody needs).
    // All eight table loads are issued before anything depends on them: these CTAs live for a microsecond.
    uint32_t h[4], zf[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        h[j] = __ldg(stream_hdr + c * 4 + j);
        zf[j] = zero_flags ? __ldg(zero_flags + (size_t)(c * 4 + j) * g.nsub_full + sub) : 0u;
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        src[j] = (h[j] & FZ_RAW_FLAG) ? container + stream_off[c * 4 + j] : planes + (uint64_t)j * g.plane_stride + (uint64_t)c * g.chk;
        end[j] = src[j] + n_c;
        zero[j] = zf[j] != 0 && !(h[j] & FZ_RAW_FLAG);
    }
    uint4 *out4 = (uint4 *)(words + (uint64_t)c * g.chk);
    const uint32_t base = blockIdx.x * (FZ_SPLIT_THREADS * FZ_SPLIT_UNROLL) + threadIdx.x;
    uint32_t a[FZ_SPLIT_UNROLL], b[FZ_SPLIT_UNROLL], cc[FZ_SPLIT_UNROLL], d[FZ_SPLIT_UNROLL];
#pragma unroll
    for (int k = 0; k < FZ_SPLIT_UNROLL; k++) {
        const uint32_t i = base + k * FZ_SPLIT_THREADS;
        if (i < nvec) {
            a[k] = zero[0] ? 0u : fz_ld_u32_unaligned(src[0], (uint64_t)i * 4, end[0]);
            b[k] = zero[1] ? 0u : fz_ld_u32_unaligned(src[1], (uint64_t)i * 4, end[1]);
            cc[k] = zero[2] ? 0u : fz_ld_u32_unaligned(src[2], (uint64_t)i * 4, end[2]);
            d[k] = zero[3] ? 0u : fz_ld_u32_unaligned(src[3], (uint64_t)i * 4, end[3]);
        }
    }
#pragma unroll
    for (int k = 0; k < FZ_SPLIT_UNROLL; k++) {
        const uint32_t i = base + k * FZ_SPLIT_THREADS;
        if (i < nvec) {
            uint4 w;
            fz_transpose4(a[k], b[k], cc[k], d[k], w.x, w.y, w.z, w.w);
            out4[i] = w;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (n_c & 3u)) {  // ragged tail of the last chunk
        const uint32_t i = nvec * 4 + threadIdx.x;
        uint32_t w = 0;
        for (int j = 0; j < 4; j++) {
            const bool zt = zero_flags && !(stream_hdr[c * 4 + j] & FZ_RAW_FLAG) &&
                            zero_flags[(size_t)(c * 4 + j) * g.nsub_full + (i >> FZ_SUB_LOG2)] != 0;
            w |= (zt ? 0u : (uint32_t)src[j][i]) << (8 * j);
        }
        words[(uint64_t)c * g.chk + i] = w;
    }
}

void fz_launch_merge_streams(const uint8_t *planes, const uint8_t *container, uint64_t container_size, const uint32_t *stream_hdr,
                             const unsigned long long *stream_off, const uint32_t *zero_flags, FzBatchGeom g, uint32_t *words,
                             cudaStream_t st)
{
    static int v = -1, wide_min = 0;
    if (v < 0) {
        const char *e = getenv("MRCZIP_MERGE"), *m = getenv("MRCZIP_MERGE_WIDE_MIN");   // (A/B runs)
        if (m && atoi(m) >= 0 && atoi(m) <= 5) wide_min = atoi(m);   // RAW planes a chunk needs for the 16-words-per-thread path (5: never)
        v = (e && atoi(e) == 1) ? 1 : 2;   // 1: the 4-bytes-per-plane kernel only
    }
    if (v == 1) {
        const uint32_t per = FZ_SPLIT_THREADS * FZ_SPLIT_UNROLL;
        dim3 grid((g.chk / 4 + per - 1) / per, g.nchunks);
        fz_merge_streams_kernel<<<grid, FZ_SPLIT_THREADS, 0, st>>>(planes, container, stream_hdr, stream_off, zero_flags, g, words, 9u, container + container_size);
        return;
    }
    // one launch; every chunk takes the 16-words-per-thread path unless MRCZIP_MERGE_WIDE_MIN asks for the old split
    const uint32_t per4 = FZ_SPLIT_THREADS * FZ_SPLIT_UNROLL;
    dim3 grid4((g.chk / 4 + per4 - 1) / per4, g.nchunks);
    fz_merge_streams_kernel<<<grid4, FZ_SPLIT_THREADS, 0, st>>>(planes, container, stream_hdr, stream_off, zero_flags, g, words,
                                                                (uint32_t)wide_min, container + container_size);
}

// =================================================================================================
// deflate: three kernels
//   fz_hist2_kernel       one warp per 16 KiB sub-block: 2 KiB sample (stored / all-zero decisions), token histogram of
//                         every fourth sub-block, accumulated per code group of FZ_CODE_SUBS = 128 sub-blocks
//   fz_group_code_kernel  one warp per code group: Huffman code + block header (once per 2 MiB of plane)
//   fz_emit2_kernel       one warp per sub-block: one pass, lane-parallel bit emission, stored-vs-dynamic decision
// =================================================================================================
#ifndef FZ_ENC_WARPS
#define FZ_ENC_WARPS 4
#endif
__device__ __forceinline__ void fz_cp_async16(uint32_t smem_addr, const void *gptr)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void fz_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void fz_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ragged (last sub-block of a file) or unaligned (chunk sizes that are not a multiple of 16) sub-blocks: straight
// from global memory, no staging -- rare, small, and not worth shared memory that would cost the fast path occupancy
struct GlobLoad16 {
    const uint8_t *src;
    __device__ __forceinline__ FzVec16 operator()(uint32_t i) const
    {
        FzVec16 r;
        if (((uintptr_t)(src + i) & 15u) == 0) {
            const uint4 v = *(const uint4 *)(src + i);
            r.w[0] = v.x; r.w[1] = v.y; r.w[2] = v.z; r.w[3] = v.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++)
                r.w[j] = (uint32_t)src[i + 4 * j] | ((uint32_t)src[i + 4 * j + 1] << 8) | ((uint32_t)src[i + 4 * j + 2] << 16) |
                         ((uint32_t)src[i + 4 * j + 3] << 24);
        }
        return r;
    }
};
struct ZeroLoad16 {
    __device__ __forceinline__ FzVec16 operator()(uint32_t) const { FzVec16 r; r.w[0] = r.w[1] = r.w[2] = r.w[3] = 0; return r; }
};

__device__ __forceinline__ const uint8_t *fz_sub_src(const uint8_t *planes, const FzBatchGeom &g, uint32_t s, uint32_t k)
{
    const uint32_t c = s >> 2, j = s & 3;
    return planes + (uint64_t)j * g.plane_stride + (uint64_t)c * g.chk + (uint64_t)k * FZ_SUB;
}

// sub-block slot t -> (stream, k, n); false if the slot is empty (ragged last chunk)
__device__ __forceinline__ bool fz_slot(const FzBatchGeom &g, uint32_t t, uint32_t &s, uint32_t &k, uint32_t &n)
{
    s = t / g.nsub_full;
    k = t - s * g.nsub_full;
    const uint32_t n_s = fz_chunk_n(g, s / FZ_PLANES);
    const uint32_t off = k * FZ_SUB;
    if (off >= n_s) return false;
    n = min((uint32_t)FZ_SUB, n_s - off);
    return true;
}

// code groups (FZ_CODE_SUBS sub-blocks that share one Huffman code) per stream
__device__ __forceinline__ uint32_t fz_groups_per_stream(const FzBatchGeom &g)
{
    return (g.nsub_full + FZ_CODE_SUBS - 1) / FZ_CODE_SUBS;
}

// Sub-blocks whose byte distribution is (nearly) flat cannot be entropy coded: a 2 KiB sample (the first 64
// bytes of every lane piece) decides that before the 16 KiB are even read.  The plug-in entropy of 2048
// samples of uniform bytes is about 7.91 bits (bias -255 / (2 N ln 2)); anything above FZ_SAMPLE_BITS is
// emitted as a stored block (and, if the whole stream ends up like that, the stream becomes RAW).
#ifndef FZ_SAMPLE_BITS
#define FZ_SAMPLE_BITS 7.85f
#endif

#ifndef FZ_HIST_SAMPLE
#define FZ_HIST_SAMPLE 4u   // 1 = every sub-block is tokenised for the histogram
#endif
__global__ void __launch_bounds__(FZ_ENC_WARPS * FZ_WARP)
fz_group_code_kernel(const uint32_t *__restrict__ ghist, FzBatchGeom g, FzGroupCode *__restrict__ gcodes)
{
    extern __shared__ __align__(16) uint8_t fz_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t gps = fz_groups_per_stream(g);
    const uint32_t gi = blockIdx.x * FZ_ENC_WARPS + warp;
    if (gi >= g.nchunks * FZ_PLANES * gps) return;
    const uint32_t s = gi / gps, gk = gi - s * gps;
    const uint32_t n_s = fz_chunk_n(g, s / FZ_PLANES);
    const uint64_t gbytes = (uint64_t)FZ_SUB * FZ_CODE_SUBS;
    if ((uint64_t)gk * gbytes >= n_s) return;
    const uint32_t gn = (uint32_t)min((uint64_t)n_s - (uint64_t)gk * gbytes, gbytes);
    const uint32_t nsub = (gn + FZ_SUB - 1) / FZ_SUB;
    FzEncState *st = (FzEncState *)fz_smem + warp;
    uint32_t any = 0;
    for (int i = lane; i < 288; i += 32) { const uint32_t v = ghist[(uint64_t)gi * 288 + i]; st->hist[i] = v; any |= v; }
    if (!__any_sync(0xffffffffu, any != 0)) {
        // no tokens at all: every sub-block of the group was ruled incompressible by its sample (flat mantissa
        // planes) and is already marked stored -- nobody will read this group's code
        if (lane == 0) gcodes[gi].stored = 1u;
        return;
    }
    __syncwarp();
    const bool standins = st->hist[287] != 0;
    if (standins) {
        // some sub-blocks of the group were not tokenised (fz_hist2_kernel): every literal and every run length gets a code
        for (int i = lane; i < 286; i += 32)
            if (i != FZ_EOB && st->hist[i] == 0) st->hist[i] = 1;
    }
    __syncwarp();
    if (lane == 0) { st->hist[FZ_EOB] = nsub; st->hist[286] = 0; st->hist[287] = 0; }  // one end-of-block per sub-block
    __syncwarp();
    fz_build_group_code(st, gn, nsub, gcodes + gi, lane, standins ? FZ_HIST_SAMPLE : 0u);
}

// =================================================================================================
// histogram and emission (fz_enc2.cuh): warp-interleaved steps, one pass, coalesced loads and stores
// =================================================================================================
struct GlobVec16 {   // 16-byte aligned source: one streaming 128-bit load per lane, 512 contiguous bytes per warp
    const uint8_t *src;
    __device__ __forceinline__ FzVec16 operator()(uint32_t i) const
    {
        const uint4 v = fz_ld_stream((const uint4 *)(src + i));
        FzVec16 r;
        r.w[0] = v.x; r.w[1] = v.y; r.w[2] = v.z; r.w[3] = v.w;
        return r;
    }
};

// most frequent byte of a 256-bin histogram, ignoring `skip`: count << 8 | byte (warp-wide)
__device__ __forceinline__ uint32_t fz_warp_hist_mode(const uint32_t *hist, uint32_t skip, int lane)
{
    uint32_t best = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t b = (uint32_t)lane * 8u + (uint32_t)i;
        const uint32_t key = b == skip ? 0u : ((hist[b] << 8) | b);
        best = key > best ? key : best;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const uint32_t o = __shfl_xor_sync(0xffffffffu, best, d);
        best = o > best ? o : best;
    }
    return best;
}

#define FZ_HIST2_SKIP_MIN 160u   // of 2048 sample bytes: below ~8 % a private counter costs more than the conflicts it saves

__global__ void __launch_bounds__(FZ_ENC_WARPS * FZ_WARP)
fz_hist2_kernel(const uint8_t *__restrict__ planes, FzBatchGeom g, uint32_t *__restrict__ ghist, uint32_t *__restrict__ sizes,
                const uint32_t *__restrict__ zero_hist, uint32_t zero_planes, uint64_t zero_from, FzStatus *status)
{
    __shared__ uint32_t hist_sh[FZ_ENC_WARPS][288];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t t = blockIdx.x * FZ_ENC_WARPS + warp;
    if (t >= g.nchunks * FZ_PLANES * g.nsub_full) return;
    uint32_t s, k, n;
    if (!fz_slot(g, t, s, k, n)) return;
    uint32_t *hist = hist_sh[warp];
    const uint8_t *src = fz_sub_src(planes, g, s, k);
    uint32_t *gh = ghist + ((uint64_t)s * fz_groups_per_stream(g) + k / FZ_CODE_SUBS) * 288;
    const uint64_t masked_from = g.chunk_exempt ? (uint64_t)(s >> 2) * g.chk + g.chunk_exempt[s >> 2] : zero_from;
    if (n == FZ_SUB && ((zero_planes >> (s & 3u)) & 1u) && (uint64_t)(s >> 2) * g.chk + (uint64_t)k * FZ_SUB >= masked_from) {
        // the mask erases this whole byte plane (8 or more bits erased) and the sub-block lies behind the exempt
        // header words: 16 KiB of zeros, known without reading them
        if (lane == 0) sizes[t] = FZ_SIZE_ZERO_FLAG;
        for (int i = lane; i < 288; i += 32) {
            const uint32_t v = zero_hist[i];
            if (v) atomicAdd(gh + i, v);
        }
        return;
    }
    for (int i = lane; i < 288; i += 32) hist[i] = 0;
    __syncwarp();
    uint32_t skip1 = 0x100u, skip2 = 0x100u;
    const bool aligned = n == FZ_SUB && ((uintptr_t)src & 15u) == 0;
    if (aligned) {
        // ---- sample pass: 64 bytes out of every 512 (2 KiB in all)
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint4 v = *(const uint4 *)(src + lane * (FZ_SUB / 32) + 16 * q);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int b = 0; b < 16; b++) atomicAdd(&hist[(w[b >> 2] >> ((b & 3) * 8)) & 0xffu], 1u);
        }
        __syncwarp();
        float acc = 0.f;
        for (int i = lane; i < 256; i += 32) {
            const float f = (float)hist[i];
            if (f > 0.f) acc += f * __log2f(f);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
        const float bits = 11.f - acc * (1.f / 2048.f);  // log2(2048) - sum f log2 f / N
        if (bits > FZ_SAMPLE_BITS) {
            if (lane == 0) { sizes[t] = fz_stored_size(n) | FZ_SIZE_STORED_FLAG; atomicAdd(&status->n_stored_sub, 1u); }
            return;
        }
        if (hist[0] == 2048u) {
            // The sample is all zeros: one OR over the 16 KiB settles whether the sub-block is; then its tokens are known
            // without scanning (zero_hist, made once by fz_zero_hist_kernel with this very tokeniser), and the emit kernel
            // encodes only the first zero sub-block of each group -- the others are copies of its fragment.
            uint32_t any = 0;
#pragma unroll 4
            for (uint32_t i = lane * 16; i < FZ_SUB; i += FZ_WARP * 16) {
                const uint4 v = *(const uint4 *)(src + i);
                any |= v.x | v.y | v.z | v.w;
            }
            if (!__any_sync(0xffffffffu, any != 0)) {
                if (lane == 0) sizes[t] = FZ_SIZE_ZERO_FLAG;
                for (int i = lane; i < 288; i += 32) {
                    const uint32_t v = zero_hist[i];
                    if (v) atomicAdd(gh + i, v);
                }
                return;
            }
        }
        // the two most frequent bytes of the sample are counted in registers (same-address shared atomics serialise)
        const uint32_t m1 = fz_warp_hist_mode(hist, 0x100u, lane);
        if ((m1 >> 8) >= FZ_HIST2_SKIP_MIN) {
            skip1 = m1 & 0xffu;
            const uint32_t m2 = fz_warp_hist_mode(hist, skip1, lane);
            if ((m2 >> 8) >= FZ_HIST2_SKIP_MIN) skip2 = m2 & 0xffu;
        }
        __syncwarp();
        for (int i = lane; i < 288; i += 32) hist[i] = 0;
        __syncwarp();
    }
    if (lane == 0) sizes[t] = 0;  // to be decided by the emit kernel
    if (aligned && (k % FZ_HIST_SAMPLE) != 0) {
        // The group's code is built from every FZ_HIST_SAMPLE-th sub-block (a group is 512 KiB of one plane: a quarter of
        // it is 8 K..128 K tokens).  Slot 287 of the group histogram counts the sub-blocks left out: the code builder then
        // gives every literal and every run length a code, so whatever they hold can be coded.
        if (lane == 0) atomicAdd(gh + 287, 1u);
        return;
    }
    const FzWarp w{lane};
    if (aligned) fz_hist2_subblock(w, hist, GlobVec16{src}, n, skip1, skip2);
    else fz_hist2_subblock(w, hist, GlobLoad16{src}, n, skip1, skip2);
    const uint32_t weight = aligned ? FZ_HIST_SAMPLE : 1u;
    for (int i = lane; i < 286; i += 32) {
        const uint32_t v = hist[i];
        if (v) atomicAdd(gh + i, v * weight);
    }
}

__global__ void __launch_bounds__(FZ_WARP)
fz_zero_hist2_kernel(uint32_t *__restrict__ zero_hist)
{
    __shared__ uint32_t hist[288];
    const int lane = threadIdx.x;
    for (int i = lane; i < 288; i += 32) hist[i] = 0;
    __syncwarp();
    const FzWarp w{lane};
    fz_hist2_subblock(w, hist, ZeroLoad16{}, FZ_SUB, 0x100u, 0x100u);
    for (int i = lane; i < 288; i += 32) zero_hist[i] = hist[i];
}

struct __align__(16) FzEmit2Smem {
    uint32_t ring[FZ_E2_RING_WORDS];
    uint32_t gc_hot[FZ_GROUP_CODE_HOT_BYTES / 4];
    uint32_t tt[256];   // the group's match tokens
};

#ifndef FZ_EMIT_MINBLOCKS
#define FZ_EMIT_MINBLOCKS 7
#endif
__global__ void __launch_bounds__(FZ_ENC_WARPS * FZ_WARP, FZ_EMIT_MINBLOCKS)
fz_emit2_kernel(const uint8_t *__restrict__ planes, FzBatchGeom g, const FzGroupCode *__restrict__ gcodes,
                uint8_t *__restrict__ scratch, uint32_t *__restrict__ sizes, FzStatus *status)
{
    __shared__ FzEmit2Smem smem[FZ_ENC_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t t = blockIdx.x * FZ_ENC_WARPS + warp;
    if (t >= g.nchunks * FZ_PLANES * g.nsub_full) return;
    uint32_t s, k, n;
    if (!fz_slot(g, t, s, k, n)) return;
    const uint32_t sz0 = sizes[t];
    if (sz0 & FZ_SIZE_STORED_FLAG) return;  // the histogram kernel already ruled this sub-block incompressible
    const FzGroupCode *ggc = gcodes + ((uint64_t)s * fz_groups_per_stream(g) + k / FZ_CODE_SUBS);
    if (ggc->stored) {  // the whole group cannot beat stored blocks: nothing to emit
        if (lane == 0) { sizes[t] = fz_stored_size(n) | FZ_SIZE_STORED_FLAG; atomicAdd(&status->n_stored_sub, 1u); }
        return;
    }
    if (sz0 & FZ_SIZE_ZERO_FLAG) {
        // all-zero sub-blocks of a group are the same bytes coded with the same code: only the first one is encoded,
        // the layout kernel gives the others its size and the gather kernel copies its fragment
        const uint32_t kb = k & ~(uint32_t)(FZ_GROUP_SUBS - 1);
        const uint32_t other = (kb + lane < g.nsub_full) ? sizes[t - (k - kb) + lane] : 0u;
        const uint32_t zmask = __ballot_sync(0xffffffffu, (other & FZ_SIZE_ZERO_FLAG) != 0);
        if ((uint32_t)(__ffs((int)zmask) - 1) != k - kb) return;
    }
    FzEmit2Smem *sm = &smem[warp];
    {
        const uint32_t *src = (const uint32_t *)ggc;
        for (uint32_t i = lane; i < FZ_GROUP_CODE_HOT_BYTES / 4; i += 32) sm->gc_hot[i] = src[i];
    }
    __syncwarp();
    const uint8_t *src = fz_sub_src(planes, g, s, k);
    uint32_t *out = (uint32_t *)(scratch + (uint64_t)t * FZ_SLOT_STRIDE);
    const FzGroupCode *hot = (const FzGroupCode *)sm->gc_hot;
    const FzWarp w{lane};
    uint32_t r;
    if (sz0 & FZ_SIZE_ZERO_FLAG) r = fz_emit2_subblock(w, hot->cl, ggc->hdr, hot->hdr_nbits, sm->ring, sm->tt, ZeroLoad16{}, n, out);
    else if (n == FZ_SUB && ((uintptr_t)src & 15u) == 0) r = fz_emit2_subblock(w, hot->cl, ggc->hdr, hot->hdr_nbits, sm->ring, sm->tt, GlobVec16{src}, n, out);
    else r = fz_emit2_subblock(w, hot->cl, ggc->hdr, hot->hdr_nbits, sm->ring, sm->tt, GlobLoad16{src}, n, out);
    if (lane == 0) {
        sizes[t] = r | (sz0 & FZ_SIZE_ZERO_FLAG);
        if (r & FZ_SIZE_STORED_FLAG) atomicAdd(&status->n_stored_sub, 1u);
    }
}

void fz_launch_zero_hist(uint32_t *zero_hist, cudaStream_t st)
{
    fz_zero_hist2_kernel<<<1, FZ_WARP, 0, st>>>(zero_hist);
}

void fz_launch_encode(const uint8_t *planes, FzBatchGeom g, uint32_t *ghist, void *gcodes, uint8_t *scratch, uint32_t *sizes,
                      const uint32_t *zero_hist, uint32_t zero_planes, uint64_t zero_from, FzStatus *status, cudaStream_t st)
{
    const uint32_t nstreams = g.nchunks * FZ_PLANES;
    const uint32_t total = nstreams * g.nsub_full;
    const uint32_t gps = (g.nsub_full + FZ_CODE_SUBS - 1) / FZ_CODE_SUBS;
    const uint32_t ngroups = nstreams * gps;
    const unsigned grid = (total + FZ_ENC_WARPS - 1) / FZ_ENC_WARPS;
    cudaFuncSetAttribute(fz_group_code_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(FzEncState) * FZ_ENC_WARPS));
    cudaMemsetAsync(ghist, 0, (size_t)ngroups * 288 * sizeof(uint32_t), st);
    fz_hist2_kernel<<<grid, FZ_ENC_WARPS * FZ_WARP, 0, st>>>(planes, g, ghist, sizes, zero_hist, zero_planes, zero_from, status);
    fz_group_code_kernel<<<(ngroups + FZ_ENC_WARPS - 1) / FZ_ENC_WARPS, FZ_ENC_WARPS * FZ_WARP, sizeof(FzEncState) * FZ_ENC_WARPS, st>>>(
        ghist, g, (FzGroupCode *)gcodes);
    fz_emit2_kernel<<<grid, FZ_ENC_WARPS * FZ_WARP, 0, st>>>(planes, g, (const FzGroupCode *)gcodes, scratch, sizes, status);
}

size_t fz_group_code_bytes() { return sizeof(FzGroupCode); }

// =================================================================================================
// layout: per-stream sums + RAW rule, scan over chunk records, chunk headers
// =================================================================================================
__device__ __forceinline__ uint32_t fz_warp_incl_scan(uint32_t v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

// one warp per stream: sub_off[t] = exclusive prefix of the sub-block sizes inside the stream;
// stream_hdr[s] = payload length | RAW flag, with the reference's rule "compressed iff n > len + 4" (zip.c:177)
__global__ void __launch_bounds__(128)
fz_layout_streams_kernel(uint32_t *__restrict__ sizes, FzBatchGeom g, uint32_t *__restrict__ sub_off,
                         uint32_t *__restrict__ stream_hdr, FzStatus *status)
{
    const int lane = threadIdx.x & 31;
    const uint32_t s = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (s >= g.nchunks * FZ_PLANES) return;
    const uint32_t n_s = fz_chunk_n(g, s / FZ_PLANES);
    const uint32_t nsub = (n_s + FZ_SUB - 1) / FZ_SUB;
    uint32_t carry = 0;
    for (uint32_t k0 = 0; k0 < nsub; k0 += 32) {
        const uint32_t k = k0 + lane;
        uint32_t raw = k < nsub ? sizes[s * g.nsub_full + k] : 0u;
        // one iteration = one group: all-zero sub-blocks take size (and fragment) of the group's first one
        const uint32_t zmask = __ballot_sync(0xffffffffu, (raw & FZ_SIZE_ZERO_FLAG) != 0);
        if (zmask) {
            if (lane == 0) atomicAdd(&status->n_zero_sub, (unsigned int)__popc(zmask));
            const int leader = __ffs((int)zmask) - 1;
            const uint32_t lv = __shfl_sync(0xffffffffu, raw, leader);
            if ((raw & FZ_SIZE_ZERO_FLAG) && lane != leader) {
                raw = (lv & FZ_SIZE_STORED_FLAG) ? (lv & (FZ_SIZE_MASK | FZ_SIZE_STORED_FLAG))
                                                 : ((lv & FZ_SIZE_MASK) | FZ_SIZE_COPY_FLAG | ((uint32_t)leader << 24));
                sizes[s * g.nsub_full + k] = raw;
                if (lv & FZ_SIZE_STORED_FLAG) atomicAdd(&status->n_stored_sub, 1u);
            }
        }
        const uint32_t v = raw & FZ_SIZE_MASK;
        const uint32_t inc = fz_warp_incl_scan(v, lane);
        if (k < nsub) sub_off[s * g.nsub_full + k] = carry + inc - v;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) {
        const bool compressed = n_s > carry + 4u;
        stream_hdr[s] = compressed ? carry : (n_s | FZ_RAW_FLAG);
        if (!compressed) atomicAdd(&status->n_raw_streams, 1u);
    }
}

// single block: exclusive scan of the chunk record sizes (16 + four payloads), starting at status->out_end;
// writes the 16-byte chunk headers (reference workers.c:837-842) and the payload offset of every stream.
#define FZ_LAYOUT_THREADS 1024
__global__ void __launch_bounds__(FZ_LAYOUT_THREADS)
fz_layout_chunks_kernel(const uint32_t *__restrict__ stream_hdr, FzBatchGeom g, unsigned long long *__restrict__ stream_off,
                        uint8_t *__restrict__ container, uint64_t container_cap, FzStatus *status)
{
    __shared__ unsigned long long warp_sums[32];
    __shared__ unsigned long long base_sh;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) base_sh = status->out_end;
    __syncthreads();
    for (uint32_t c0 = 0; c0 < g.nchunks; c0 += FZ_LAYOUT_THREADS) {
        const uint32_t c = c0 + threadIdx.x;
        uint32_t h[4] = {0, 0, 0, 0};
        unsigned long long rec = 0;
        if (c < g.nchunks) {
            rec = FZ_CHUNK_HEADER_BYTES;
            for (int j = 0; j < 4; j++) { h[j] = stream_hdr[c * 4 + j]; rec += h[j] & ~FZ_RAW_FLAG; }
        }
        // block exclusive scan of rec
        unsigned long long inc = rec;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        if (lane == 31) warp_sums[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            unsigned long long ws = warp_sums[lane], wi = ws;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long o = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += o;
            }
            warp_sums[lane] = wi - ws;  // exclusive
        }
        __syncthreads();
        const unsigned long long base = base_sh;
        const unsigned long long off = base + warp_sums[warp] + inc - rec;
        if (c < g.nchunks) {
            if (off + rec > container_cap) atomicCAS(&status->error, 0, FZ_E_SPACE);
            else {
                unsigned long long p = off + FZ_CHUNK_HEADER_BYTES;
                for (int j = 0; j < 4; j++) {
                    // pack_header (reference zip.c:381-391): little-endian length, bit 31 = RAW
                    container[off + 4 * j + 0] = (uint8_t)(h[j]);
                    container[off + 4 * j + 1] = (uint8_t)(h[j] >> 8);
                    container[off + 4 * j + 2] = (uint8_t)(h[j] >> 16);
                    container[off + 4 * j + 3] = (uint8_t)(h[j] >> 24);
                    stream_off[c * 4 + j] = p;
                    p += h[j] & ~FZ_RAW_FLAG;
                }
            }
        }
        __syncthreads();
        // total of this tile = exclusive offset of the last thread + its value
        if (threadIdx.x == FZ_LAYOUT_THREADS - 1) base_sh = off + rec;
        __syncthreads();
    }
    if (threadIdx.x == 0) status->out_end = base_sh;
}

void fz_launch_layout(uint32_t *sizes, FzBatchGeom g, uint32_t *sub_off, uint32_t *stream_hdr,
                      unsigned long long *stream_off, uint8_t *container, uint64_t container_cap, FzStatus *status, cudaStream_t st)
{
    const uint32_t nstreams = g.nchunks * FZ_PLANES;
    fz_layout_streams_kernel<<<(nstreams + 3) / 4, 128, 0, st>>>(sizes, g, sub_off, stream_hdr, status);
    fz_layout_chunks_kernel<<<1, FZ_LAYOUT_THREADS, 0, st>>>(stream_hdr, g, stream_off, container, container_cap, status);
}

// one warp per sub-block: move the encoded fragment (or the raw / stored plane bytes) to its place in the container
__global__ void __launch_bounds__(128, 16)
fz_gather_kernel(const uint8_t *__restrict__ planes, const uint8_t *__restrict__ scratch, const uint32_t *__restrict__ sizes,
                 const uint32_t *__restrict__ sub_off, const uint32_t *__restrict__ stream_hdr,
                 const unsigned long long *__restrict__ stream_off, FzBatchGeom g, uint8_t *__restrict__ container,
                 const FzStatus *status)
{
    const int lane = threadIdx.x & 31;
    const uint32_t t = blockIdx.x * 4 + (threadIdx.x >> 5);
    const uint32_t total = g.nchunks * FZ_PLANES * g.nsub_full;
    if (t >= total || status->error) return;
    const uint32_t s = t / g.nsub_full, k = t - s * g.nsub_full;
    const uint32_t n_s = fz_chunk_n(g, s / FZ_PLANES);
    const uint32_t off = k * FZ_SUB;
    if (off >= n_s) return;
    const uint32_t n = min((uint32_t)FZ_SUB, n_s - off);
    const uint8_t *psrc = fz_sub_src(planes, g, s, k);
    uint8_t *dst = container + stream_off[s];
    if (stream_hdr[s] & FZ_RAW_FLAG) {
        fz_warp_copy(dst + off, psrc, n, psrc + n + 32, lane);
        return;
    }
    dst += sub_off[t];
    const uint32_t sz = sizes[t];
    if (sz & FZ_SIZE_STORED_FLAG) {
        // two stored blocks (BFINAL=0, BTYPE=00, LEN, NLEN, data) + empty stored block; the split point breaks a
        // sync-marker pattern in the raw bytes, if there is one (fz_stored_split)
        if (n < 2) {
            if (lane == 0) {
                dst[0] = 0x00; dst[1] = (uint8_t)n; dst[2] = 0; dst[3] = (uint8_t)~n; dst[4] = 0xFF;
                for (uint32_t i = 0; i < n; i++) dst[5 + i] = psrc[i];
                uint8_t *t5 = dst + 5 + n;
                t5[0] = 0x00; t5[1] = 0x00; t5[2] = 0x00; t5[3] = 0xFF; t5[4] = 0xFF;
            }
            return;
        }
        uint32_t first = 0xFFFFFFFFu;
        for (uint32_t i0 = 0; i0 < n; i0 += 32 * 16) {   // warp-uniform trip count
            const uint32_t i = i0 + lane * 16;
            uint32_t m = 0;
            if (i < n) {
                // 16 candidate positions i .. i+15 need bytes [i, i+19); psrc is 16-byte aligned or handled bytewise
                uint32_t W[5];
                if (((uintptr_t)psrc & 3u) == 0) {
#pragma unroll
                    for (int k = 0; k < 5; k++) W[k] = (i + 4 * k < n) ? *(const uint32_t *)(psrc + i + 4 * k) : 0u;
                } else {
#pragma unroll
                    for (int k = 0; k < 5; k++) {
                        uint32_t v = 0;
                        for (int b = 0; b < 4; b++) if (i + 4 * k + b < n) v |= (uint32_t)psrc[i + 4 * k + b] << (8 * b);
                        W[k] = v;
                    }
                }
#pragma unroll
                for (int b = 0; b < 16; b++) {
                    const uint32_t v = __funnelshift_r(W[b >> 2], W[(b >> 2) + 1], (b & 3) * 8);
                    if (v == FZ_MARKER_LE && i + b + 4 <= n) m |= 1u << b;
                }
            }
            const uint32_t any = __ballot_sync(0xffffffffu, m != 0);
            if (any && first == 0xFFFFFFFFu) {
                const int src_lane = __ffs((int)any) - 1;
                const uint32_t mm = __shfl_sync(0xffffffffu, m, src_lane);
                first = i0 + (uint32_t)src_lane * 16 + (uint32_t)(__ffs((int)mm) - 1);
            }
        }
        const uint32_t s = fz_stored_split(n, first);
        const uint32_t r = n - s;
        if (lane == 0) {
            dst[0] = 0x00; dst[1] = (uint8_t)s; dst[2] = (uint8_t)(s >> 8); dst[3] = (uint8_t)~s; dst[4] = (uint8_t)(~s >> 8);
            uint8_t *h2 = dst + 5 + s;
            h2[0] = 0x00; h2[1] = (uint8_t)r; h2[2] = (uint8_t)(r >> 8); h2[3] = (uint8_t)~r; h2[4] = (uint8_t)(~r >> 8);
            uint8_t *t5 = dst + 10 + n;
            t5[0] = 0x00; t5[1] = 0x00; t5[2] = 0x00; t5[3] = 0xFF; t5[4] = 0xFF;
        }
        fz_warp_copy(dst + 5, psrc, s, psrc + n + 32, lane);
        fz_warp_copy(dst + 10 + s, psrc + s, r, psrc + n + 32, lane);
    } else {
        // (an all-zero sub-block that is not the first of its group: the first one's fragment, byte for byte)
        const uint32_t tsrc = (sz & FZ_SIZE_COPY_FLAG) ? t - (k & (FZ_GROUP_SUBS - 1)) + ((sz >> 24) & 31u) : t;
        const uint8_t *ssrc = scratch + (uint64_t)tsrc * FZ_SLOT_STRIDE;
        fz_warp_copy(dst, ssrc, sz & FZ_SIZE_MASK, ssrc + FZ_SLOT_STRIDE, lane);
    }
}

void fz_launch_gather(const uint8_t *planes, const uint8_t *scratch, const uint32_t *sizes, const uint32_t *sub_off,
                      const uint32_t *stream_hdr, const unsigned long long *stream_off, FzBatchGeom g, uint8_t *container,
                      const FzStatus *status, cudaStream_t st)
{
    const uint32_t total = g.nchunks * FZ_PLANES * g.nsub_full;
    fz_gather_kernel<<<(total + 3) / 4, 128, 0, st>>>(planes, scratch, sizes, sub_off, stream_hdr, stream_off, g, container, status);
}

// =================================================================================================
// inflate side
// =================================================================================================

// chunk-header chain (reference workers.c:61-69): serial by nature, 16 bytes per hop
// One thread: the chain is serial by format (16 bytes per hop, one round trip to memory each: 0.12 ms for the 171 chunk
// records of a 4 GiB volume).  Measured and not kept: pulling the 32 KiB around the guessed position of the next header
// into L2 while this one is on its way (records of one file are nearly equally long) changed nothing.
__global__ void fz_walk_kernel(const uint8_t *__restrict__ container, uint64_t container_size, FzBatchGeom g,
                               uint32_t *__restrict__ stream_hdr, unsigned long long *__restrict__ stream_off, FzStatus *status)
{
    if (threadIdx.x != 0 || status->error) return;
    unsigned long long off = status->out_end;
    for (uint32_t c = 0; c < g.nchunks; c++) {
        if (off + FZ_CHUNK_HEADER_BYTES > container_size) { status->error = FZ_E_FORMAT; return; }
        uint8_t b[16];
#pragma unroll
        for (int i = 0; i < 16; i++) b[i] = container[off + i];
        off += FZ_CHUNK_HEADER_BYTES;
        const uint32_t n_s = fz_chunk_n(g, c);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            // unpack_header (reference zip.c:394-399)
            const uint32_t h = (uint32_t)b[4 * j] | ((uint32_t)b[4 * j + 1] << 8) | ((uint32_t)b[4 * j + 2] << 16) | ((uint32_t)b[4 * j + 3] << 24);
            const uint32_t len = h & ~FZ_RAW_FLAG;
            if ((h & FZ_RAW_FLAG) && len != n_s) { status->error = FZ_E_FORMAT; return; }
            if (off + len > container_size) { status->error = FZ_E_FORMAT; return; }
            stream_hdr[c * 4 + j] = h;
            stream_off[c * 4 + j] = off;
            off += len;
        }
    }
    status->out_end = off;
}

void fz_launch_walk(const uint8_t *container, uint64_t container_size, FzBatchGeom g, uint32_t *stream_hdr,
                    unsigned long long *stream_off, FzStatus *status, cudaStream_t st)
{
    fz_walk_kernel<<<1, 32, 0, st>>>(container, container_size, g, stream_hdr, stream_off, status);
}

// ---- sync-marker scan: positions p (stream relative) with bytes p..p+3 == 00 00 FF FF
#define FZ_TILE_BYTES 65536   // payload bytes per marker-scan block (one count per tile)
#define FZ_SLICE_BYTES 4096   // bytes the block covers per iteration (256 threads x 16 positions)
#define FZ_SCAN_THREADS 256

__device__ __forceinline__ uint32_t fz_marker_mask(const uint8_t *base, uint32_t len, uint32_t p0)
{
    // 16 candidate positions p0 .. p0+15; needs bytes [p0, p0+19)
    if (p0 + 4 > len) return 0;
    const uint8_t *addr = base + p0;
    const uint32_t sk = (uint32_t)((uintptr_t)addr & 3u);
    const uint32_t *a0 = (const uint32_t *)(addr - sk);
    const uint32_t *aend = (const uint32_t *)(((uintptr_t)(base + len) + 3u) & ~(uintptr_t)3u);
    uint32_t W[6];
#pragma unroll
    for (int k = 0; k < 6; k++) W[k] = (a0 + k) < aend ? a0[k] : 0u;
    uint32_t V[5];
#pragma unroll
    for (int k = 0; k < 5; k++) V[k] = __funnelshift_r(W[k], W[k + 1], sk * 8);
    // bit i of F: byte i of the 20-byte window is 0xFF (zero-byte trick on ~V, four flags gathered by one multiply).
    // A marker at position b needs FF at b + 2 and b + 3: in compressed data that is one window in four thousand, so
    // the zero bytes are only looked at then.
    uint32_t F = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const uint32_t x = ~V[k];
        const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);   // 0x80 in every zero byte of x
        F |= ((z * 0x00204081u) >> 28) << (4 * k);
    }
    uint32_t m = (F >> 2) & (F >> 3) & 0xffffu;
    if (m) {
        uint32_t Z = 0;
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const uint32_t x = V[k];
            const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);
            Z |= ((z * 0x00204081u) >> 28) << (4 * k);
        }
        m &= Z & (Z >> 1);
        const uint32_t nvalid = len - p0 - 3u;   // positions b with p0 + b + 4 <= len (>= 1 here)
        if (nvalid < 16u) m &= (1u << nvalid) - 1u;
    }
    return m;
}

// One pass.  Block b scans tile b % tiles_per_stream of stream b / tiles_per_stream: the 16 slices' hit masks stay in
// registers, the tile's count is published, the counts of the tiles before it IN THE SAME STREAM are collected by a
// decoupled look-back (tile_state: flag << 30 | value; 1 = this tile's count, 2 = inclusive prefix), and the hits go
// straight to their place
// hits[s * hits_per_stream + rank].  stream_cnt[s] = markers of the whole stream.
#define FZ_TILE_SLICES (FZ_TILE_BYTES / FZ_SLICE_BYTES)
__global__ void __launch_bounds__(FZ_SCAN_THREADS)
fz_marker_kernel(const uint8_t *__restrict__ container, const uint32_t *__restrict__ stream_hdr,
                 const unsigned long long *__restrict__ stream_off, uint32_t tiles_per_stream, uint32_t ntickets,
                 uint32_t *__restrict__ tile_state, uint32_t *__restrict__ stream_cnt, uint32_t *__restrict__ hits, uint32_t hits_per_stream,
                 const FzStatus *status)
{
    __shared__ uint32_t wsum[FZ_SCAN_THREADS / 32];
    __shared__ uint32_t excl_sh, tick_sh;
    if (status->error) {  // a broken chunk chain leaves the stream table undefined: touch nothing else
        for (uint32_t s = blockIdx.x * FZ_SCAN_THREADS + threadIdx.x; s * tiles_per_stream < ntickets; s += gridDim.x * FZ_SCAN_THREADS) stream_cnt[s] = 0;
        return;
    }
    // A persistent grid: tiles are handed out by a ticket, in order, to whichever block is free -- a tile then only ever
    // waits (look-back) for tiles whose blocks are already running.  Most tickets are not worth a block: RAW streams are
    // not scanned and a stream's payload ends long before its last tile slot (one block per slot spent half of this
    // kernel's time starting and ending 61,000 blocks that had nothing to do); thread 0 passes over those on its own.
    for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t;
        for (;;) {
            t = atomicAdd(tile_state + (size_t)ntickets, 1u);
            if (t >= ntickets) { t = ~0u; break; }
            const uint32_t s1 = t / tiles_per_stream, tile1 = t - s1 * tiles_per_stream;
            const uint32_t h1 = stream_hdr[s1];
            if (!((h1 & FZ_RAW_FLAG) || (uint64_t)tile1 * FZ_TILE_BYTES + 4 > (h1 & ~FZ_RAW_FLAG))) break;
            if (tile1 == 0) stream_cnt[s1] = 0;   // (tiles behind a skipped one are skipped too: nobody looks back at it)
        }
        tick_sh = t;
    }
    __syncthreads();
    const uint32_t b = tick_sh;
    if (b == ~0u) return;
    const uint32_t s = b / tiles_per_stream, tile = b - s * tiles_per_stream;
    const uint32_t h = stream_hdr[s];
    const uint32_t len = h & ~FZ_RAW_FLAG;
    const bool last_tile = tile + 1 == tiles_per_stream || (uint64_t)(tile + 1) * FZ_TILE_BYTES + 4 > len;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint8_t *base = container + stream_off[s];
    uint32_t mk[FZ_TILE_SLICES / 2], rk[FZ_TILE_SLICES / 2];   // per slice: hit mask, rank inside the tile (16 bits each)
#pragma unroll
    for (int i = 0; i < FZ_TILE_SLICES / 2; i++) { mk[i] = 0; rk[i] = 0; }
    uint32_t acc = 0;  // markers found in the earlier slices of this tile
#pragma unroll
    for (uint32_t slice = 0; slice < FZ_TILE_SLICES; slice++) {
        const uint32_t q0 = tile * FZ_TILE_BYTES + slice * FZ_SLICE_BYTES;
        if ((uint64_t)q0 + 4 <= len) {   // block-uniform
            const uint32_t m = fz_marker_mask(base, len, q0 + threadIdx.x * 16);
            const uint32_t cnt = __popc(m);
            if (__syncthreads_or(cnt != 0)) {   // (a slice of compressed bytes holds a marker about once in three)
                const uint32_t inc = fz_warp_incl_scan(cnt, lane);
                if (lane == 31) wsum[warp] = inc;
                __syncthreads();
                uint32_t wbase = 0, total = 0;
#pragma unroll
                for (int w = 0; w < FZ_SCAN_THREADS / 32; w++) { if (w < warp) wbase += wsum[w]; total += wsum[w]; }
                mk[slice >> 1] |= m << (16 * (slice & 1));
                rk[slice >> 1] |= ((acc + wbase + inc - cnt) & 0xffffu) << (16 * (slice & 1));
                acc += total;
                __syncthreads();
            }
        }
    }
    if (threadIdx.x == 0) {
        volatile uint32_t *st = tile_state + (size_t)s * tiles_per_stream;
        uint32_t excl = 0;
        if (tile > 0) {
            st[tile] = (1u << 30) | acc;
            __threadfence();
            for (int t = (int)tile - 1; t >= 0; t--) {
                uint32_t v;
                do { v = st[t]; } while ((v >> 30) == 0);
                excl += v & 0x3fffffffu;
                if ((v >> 30) == 2u) break;
            }
        }
        st[tile] = (2u << 30) | (excl + acc);
        __threadfence();
        excl_sh = excl;
        if (last_tile) stream_cnt[s] = excl + acc;
    }
    __syncthreads();
    const uint32_t excl = excl_sh;
    uint32_t *hout = hits + (size_t)s * hits_per_stream;
#pragma unroll
    for (uint32_t slice = 0; slice < FZ_TILE_SLICES; slice++) {
        uint32_t mm = (mk[slice >> 1] >> (16 * (slice & 1))) & 0xffffu;
        if (mm) {
            uint32_t o = excl + ((rk[slice >> 1] >> (16 * (slice & 1))) & 0xffffu);
            const uint32_t p0 = tile * FZ_TILE_BYTES + slice * FZ_SLICE_BYTES + threadIdx.x * 16;
            while (mm) {
                const int bit = __ffs((int)mm) - 1;
                mm &= mm - 1;
                if (o < hits_per_stream) hout[o] = p0 + (uint32_t)bit;
                o++;
            }
        }
    }
    }   // next ticket
}

// ---- classify streams: RAW, fast (our sub-block framing: one marker per sub-block, last one ends the payload), general
__global__ void fz_classify_kernel(const uint32_t *__restrict__ stream_hdr, FzBatchGeom g, const uint32_t *__restrict__ stream_cnt,
                                   const uint32_t *__restrict__ hits, uint32_t hits_per_stream,
                                   uint32_t *__restrict__ stream_mode, uint32_t *__restrict__ stream_fail, FzBlockParBufs bp,
                                   const FzStatus *status)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= g.nchunks * FZ_PLANES) return;
    stream_fail[s] = 0;
    bp.cand_cnt[s] = 0;
    bp.par_ok[s] = 0;
    if (status->error) { stream_mode[s] = 0; return; }
    const uint32_t h = stream_hdr[s];
    if (h & FZ_RAW_FLAG) { stream_mode[s] = 0; return; }
    const uint32_t n_s = fz_chunk_n(g, s / FZ_PLANES);
    const uint32_t m = stream_cnt[s];
    uint32_t mode = 2;
    if (m >= 1 && m <= hits_per_stream && hits[(size_t)s * hits_per_stream + m - 1] + 4 == h) {
        // our framing: one marker per FZ_SUB-byte sub-block
        if (((n_s + FZ_SUB - 1) >> FZ_SUB_LOG2) == m) mode = 1u | ((uint32_t)FZ_SUB_LOG2 << 8);
    }
    stream_mode[s] = mode;
    if (mode == 2) bp.gen_list[atomicAdd(&bp.ctl[0], 1u)] = s;   // zlib-made stream: block-parallel path first
}

// ---- fast path: one WARP per group of 32 sub-blocks, one THREAD per sub-block fragment.
// The 32 fragments of a group carry the same Huffman code (our encoder builds one per group), so the warp
// parses the header once, builds ONE first-level lookup table in shared memory and every lane decodes its
// own fragment with it.  Lanes verify that their header bits equal the leader's; any mismatch, parse error
// or size mismatch flags the stream, which is then re-decoded by the general inflater.
#ifndef FZ_INF_WARPS
#define FZ_INF_WARPS 4
#endif
#ifndef FZ_INF_CARVEOUT_PCT
#define FZ_INF_CARVEOUT_PCT 100    // all of the 228 KB as shared memory: 4 CTAs x 54 KB
#endif
#ifndef FZ_INF_MINBLOCKS
#define FZ_INF_MINBLOCKS 4            // measured: 4 CTAs (102 registers) beat 5 and 6 on the exponent planes; 12-bit table on the 4-bit planes
#endif
#define FZ_ZERO_PROBE_BYTES 96u   // 16 KiB of zeros is ~70 bytes of run codes
#ifndef FZ_RING_CHUNKS
#define FZ_RING_CHUNKS 8u                        // 16-byte chunks of input per lane
#endif
#ifndef FZ_RING_TOPUPS
#define FZ_RING_TOPUPS 3                         // chunks a round can use up (16 iterations x 17 bits + one symbol)
#endif
#define FZ_RING_ROW_WORDS (FZ_RING_CHUNKS * 4u + 4u)  // row pitch in words (16 bytes of padding)
#ifndef FZ_FAST_ITERS
#define FZ_FAST_ITERS 16
#endif
#ifndef FZ_GLUT_BITS
#define FZ_GLUT_BITS 12                          // the CTA's table: 4096 entries, three literals of 4 bits in one lookup
#endif
#define FZ_GLUT_SIZE (1u << FZ_GLUT_BITS)
#define FZ_CODE_WARPS (FZ_CODE_SUBS / FZ_GROUP_SUBS)
static_assert(FZ_INF_WARPS == FZ_CODE_WARPS, "one CTA of the group inflater decodes one code group");
// what fz_inflate_prep_kernel leaves per code group (see the lean inflater below)
struct FzGroupDesc {
    uint32_t state;      // 0: nothing for the group kernels here, 1: the lean kernel decodes it, 2: the full kernel does
    uint32_t hdr_bits;   // bits of the dynamic block header every coded sub-block of the group starts with
    uint32_t run_bit;    // value of the one distance bit that means "distance 1" (fz_dd1_run_bit), 2: no such code
    uint32_t leader;     // sub-block (stream relative) whose header was parsed, ~0: the group has no coded sub-block
    FzCode LL;
    uint32_t pad;
    uint16_t sym[288];   // literal/length symbols sorted by (code length, value)
};
static_assert(sizeof(FzGroupDesc) == 656, "descriptor layout (copied to shared memory word by word)");
#define FZ_DESC_CODE_WORD0 4u                                     // first word of LL in the descriptor
#define FZ_DESC_CODE_WORDS ((sizeof(FzGroupDesc) / 4u) - FZ_DESC_CODE_WORD0)
struct FzGroupSmem {
    // shared by the CTA: the code of the group (its sub-blocks carry the same block header)
    alignas(16) uint32_t lut[FZ_GLUT_SIZE];
    uint16_t tab[FZ_INF_TAB_U16];  // sorted symbols + counters of the leader's parse
    FzCode LL, DD;                 // long codes and distances; the table covers the rest
    uint32_t wmask[FZ_INF_WARPS];  // coded lanes of every warp
    uint32_t hdr_bits, leader_ok, dd1, pad;
    unsigned long long lfrag;      // the leader's fragment
    // per warp: every lane's window on its fragment (cp.async)
    alignas(16) uint32_t ring[FZ_INF_WARPS][FZ_WARP * FZ_RING_ROW_WORDS];
};

__global__ void __launch_bounds__(FZ_INF_WARPS * FZ_WARP, FZ_INF_MINBLOCKS)
fz_inflate_group_kernel(const uint8_t *__restrict__ container, FzBatchGeom g, const uint32_t *__restrict__ stream_hdr,
                        const unsigned long long *__restrict__ stream_off, const uint32_t *__restrict__ stream_cnt,
                        uint32_t hits_per_stream, const uint32_t *__restrict__ hits, const uint32_t *__restrict__ stream_mode,
                        uint32_t *__restrict__ stream_fail, uint32_t *__restrict__ zero_flags, uint8_t *__restrict__ planes,
                        const FzGroupDesc *__restrict__ desc, const FzStatus *status)
{
    extern __shared__ __align__(16) uint8_t fz_smem[];
    FzGroupSmem *sm = (FzGroupSmem *)fz_smem;
    if (status->error) return;
    if (desc && desc[blockIdx.x].state != 2u) return;   // CTA-uniform: the lean kernel decoded this group (or there is nothing to decode)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cps = fz_groups_per_stream(g);      // code groups per stream
    const uint32_t s = blockIdx.x / cps, ck = blockIdx.x - s * cps;
    if ((stream_mode[s] & 0xffu) != 1u) return;      // CTA-uniform
    const size_t h0 = (size_t)s * hits_per_stream;
    const uint32_t m = stream_cnt[s];                // sub-blocks in the stream
    if (ck * FZ_CODE_SUBS >= m) return;              // CTA-uniform
    const uint32_t gk = ck * FZ_CODE_WARPS + (uint32_t)warp;
    const uint32_t k = gk * FZ_GROUP_SUBS + lane;
    const bool valid = k < m;                        // (a warp past the end of the stream idles through the barriers below)
    const uint32_t n_s = fz_chunk_n(g, s / FZ_PLANES);
    uint32_t start = 0, end = 0, expect = 0;
    uint8_t *out = nullptr;
    if (valid) {
        start = k ? hits[h0 + k - 1] + 4 : 0u;
        end = hits[h0 + k] + 4;
        const uint32_t obeg = k << FZ_SUB_LOG2;
        expect = min((uint32_t)FZ_SUB, n_s - obeg);
        out = planes + (uint64_t)(s & 3) * g.plane_stride + (uint64_t)(s >> 2) * g.chk + obeg;
    }
    const uint8_t *frag = container + stream_off[s] + start;
    const uint32_t flen = end - start;

    typedef FzInfTab<1> Tab;
    Tab tab{sm->tab, sm->tab + 288, sm->tab + 320};
    FzInflater<Tab> inf;
    bool live = false, bad = false;
    if (valid) { inf.start(frag, flen, out, expect, tab); inf.shared_tab = true; live = true; }
    inf.bind_codes(&sm->LL, &sm->DD);
    inf.lut_bits = FZ_GLUT_BITS;

    // block type of every lane's first block: 2 (dynamic) lanes share the leader's code
    uint32_t first3 = 7;
    if (valid && flen >= 1) first3 = frag[0] & 7u;   // BFINAL (must be 0) | BTYPE << 1
    const bool coded = valid && first3 == 4u;        // BFINAL = 0, BTYPE = 10
    const uint32_t coded_mask = __ballot_sync(0xffffffffu, coded);
    if (lane == 0) sm->wmask[warp] = coded_mask;
    __syncthreads();
    int lw = -1;
#pragma unroll
    for (int w2 = FZ_INF_WARPS - 1; w2 >= 0; w2--) if (sm->wmask[w2]) lw = w2;
    const bool any_coded = lw >= 0;                  // CTA-uniform
    uint32_t hdr_bits = 0;
    if (any_coded) {
        const int leader = __ffs((int)sm->wmask[lw]) - 1;
        const bool is_leader = warp == lw && lane == leader;
        if (is_leader) {
            inf.shared_tab = false;
            const bool okh = inf.block_header();     // parses the header, fills sm->tab, LL / DD
            inf.shared_tab = true;
            if (!okh || !inf.in_body) { bad = true; live = false; }
            sm->hdr_bits = (uint32_t)((int64_t)flen * 8 - inf.br.bits_left());
            sm->leader_ok = bad ? 0u : 1u;
            sm->dd1 = inf.dd1;
            sm->lfrag = (unsigned long long)(uintptr_t)frag;
        }
        __syncthreads();
        hdr_bits = sm->hdr_bits;
        inf.dd1 = sm->dd1;
        if (!sm->leader_ok) {
            if (coded) { bad = true; live = false; }
        } else {
            // every other coded lane: its first hdr_bits must equal the leader's, then skip them
            if (coded && !is_leader) {
                const uint8_t *lf = (const uint8_t *)(uintptr_t)sm->lfrag;
                const uint32_t nby = hdr_bits >> 3, rem = hdr_bits & 7u;
                bool same = (uint64_t)flen * 8 > hdr_bits;
                if (same) {
                    for (uint32_t i = 0; i < nby; i++) same &= frag[i] == lf[i];
                    if (rem) same &= ((frag[nby] ^ lf[nby]) & ((1u << rem) - 1u)) == 0;
                }
                if (!same) { bad = true; live = false; }
                else {
                    inf.br.init(frag + nby, flen - nby);
                    inf.br.refill();
                    inf.br.drop((int)rem);
                    inf.in_body = true;
                    inf.last = false;
                }
            }
            // first-level table: entry e = the first symbol (and up to two more literals) coded by the bit pattern e
            for (uint32_t e = threadIdx.x; e < FZ_GLUT_SIZE; e += FZ_INF_WARPS * FZ_WARP) sm->lut[e] = fz_lut_entry_bits<FZ_GLUT_BITS>(sm->LL, tab, e);
        }
        __syncthreads();
    }
    if (gk * FZ_GROUP_SUBS >= m) return;             // warp-uniform: nothing of the stream left for this warp
    // Stored sub-blocks (incompressible bytes: two stored blocks and the empty one) are copied by the whole warp, 16
    // bytes per lane and step; left to their lane they would be a word-by-word loop that the 31 others wait for.
    // Anything but exactly that layout stays with its lane and the general code.
    bool done_stored = false;
    {
        uint32_t stored_mask = __ballot_sync(0xffffffffu, valid && first3 == 0u);
        while (stored_mask) {
            const int j = __ffs((int)stored_mask) - 1;
            stored_mask &= stored_mask - 1;
            const uint8_t *f = (const uint8_t *)(uintptr_t)__shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)frag, j);
            uint8_t *o = (uint8_t *)(uintptr_t)__shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)out, j);
            const uint32_t fl = __shfl_sync(0xffffffffu, flen, j), ex = __shfl_sync(0xffffffffu, expect, j);
            uint32_t pos = 0, prod = 0;
            bool ok = false;
            for (int blk = 0; blk < 4; blk++) {   // warp-uniform: every lane reads the same header bytes
                if (pos + 5 > fl) break;
                const uint32_t b0 = f[pos], len = (uint32_t)f[pos + 1] | ((uint32_t)f[pos + 2] << 8);
                const uint32_t nlen = (uint32_t)f[pos + 3] | ((uint32_t)f[pos + 4] << 8);
                if ((b0 & 7u) != 0u || (len ^ 0xffffu) != nlen) break;      // BFINAL = 0, BTYPE = 00, LEN = ~NLEN
                if (len == 0) { ok = pos + 5 == fl && prod == ex; break; }   // the empty block must close the fragment
                if (prod + len > ex || pos + 5 + len > fl) break;
                fz_warp_copy(o + prod, f + pos + 5, len, f + fl, lane);
                pos += 5 + len;
                prod += len;
            }
            if (lane == j && ok) { live = false; done_stored = true; }
        }
    }
    // lock-step drive: lanes reconverge after every symbol
    const uint32_t *lut = any_coded ? sm->lut : nullptr;
    const uint32_t run_bit = fz_dd1_run_bit(inf.dd1);
    // Tiny fragments are almost always sub-blocks of zero bytes (mask bits >= 8 zero whole byte planes).  Count them
    // first without storing: if the fragment is valid and all zero it is only flagged -- the merge supplies the
    // zeros -- which saves 16 KiB of stores here and 16 KiB of loads there.
    bool all_zero = false;
    if (zero_flags) {
        bool probe = live && coded && flen <= FZ_ZERO_PROBE_BYTES;
        if (__any_sync(0xffffffffu, probe)) {
            if (probe) inf.bw.dry = true;
            bool pl = probe;
            while (__any_sync(0xffffffffu, pl)) {
                if (pl) pl = inf.step_lut(lut);
            }
            if (probe) {
                all_zero = inf.rc == FZ_INF_OK && inf.bw.produced() == expect && inf.bw.orv == 0 && !inf.bw.non_rle &&
                           inf.br.bits_left() == 0;
                if (all_zero) live = false;
                else {   // decode it for real: same state as after the header
                    const uint32_t nby = hdr_bits >> 3, rem = hdr_bits & 7u;
                    inf.bw.init(out, expect);
                    inf.br.init(frag + nby, flen - nby);
                    inf.br.refill();
                    inf.br.drop((int)rem);
                    inf.rc = FZ_INF_OK; inf.in_body = true; inf.last = false; inf.saw_eob = false;
                }
            }
        }
        if (valid) zero_flags[(size_t)s * g.nsub_full + k] = all_zero ? 1u : 0u;
    }
    // ---- fast rounds.  The lane reads its fragment through a ring of FZ_RING_CHUNKS 16-byte chunks in shared memory
    // that cp.async tops up once per round (<= 3 chunks; a round takes at most 16 x 17 bits), two rounds ahead of
    // the decoder: `cp.async.wait_group 1` leaves only the newest top-up in flight, and what a round can reach was
    // asked for at least two top-ups ago.  The bit buffer is refilled from the ring by predicated code -- no branch,
    // no global load in the loop.  A lane leaves the fast rounds at the first thing that is not a table hit, a long
    // literal or a run (end of block, errors, end of the output) and hands its position to the general inflater.
    {
        bool fast = live && coded && lut != nullptr && inf.in_body && inf.bw.op0 == 0 && ((uintptr_t)inf.bw.out & 15u) == 0;
        // completed output words wait in pw0..pw2 until the fourth of their 16-byte group arrives and leave as one
        // 128-bit store (4-byte stores cost a 32-byte sector each on the way to L2: 8x the bytes written).  Outside
        // the literal path nothing is pending: FZ_FLUSH_PW stores what is (word index & 3 words) before anything else
        // touches the writer.
        uint32_t pw0 = 0, pw1 = 0, pw2 = 0, lastw = 0;   // lastw: the last completed word (a run needs the byte before it)
#define FZ_FLUSH_PW()                                                                           \
        do {                                                                                    \
            const uint32_t k_ = (inf.bw.op >> 2) & 3u;                                          \
            uint32_t *q_ = (uint32_t *)(inf.bw.out + (inf.bw.op & ~15u));                       \
            if (k_ > 0) q_[0] = pw0;                                                            \
            if (k_ > 1) q_[1] = pw1;                                                            \
            if (k_ > 2) q_[2] = pw2;                                                            \
        } while (0)
        // after the general inflater wrote (it stores every completed word itself): what precedes the pending word
#define FZ_RELOAD_PW()                                                                          \
        do {                                                                                    \
            const uint32_t k_ = (inf.bw.op >> 2) & 3u;                                          \
            const uint32_t *q_ = (const uint32_t *)(inf.bw.out + (inf.bw.op & ~15u));           \
            if (k_ > 0) pw0 = q_[0];                                                            \
            if (k_ > 1) pw1 = q_[1];                                                            \
            if (k_ > 2) pw2 = q_[2];                                                            \
            lastw = inf.bw.op >= 4u ? *(const uint32_t *)(inf.bw.out + (inf.bw.op & ~3u) - 4u) : 0u; \
        } while (0)
        // word w_ completes place kq_ (0..3) of the 16-byte group at gaddr_
#define FZ_WORD_DONE(w_, kq_, gaddr_)                                                           \
        do {                                                                                    \
            if ((kq_) == 3u) *(uint4 *)(gaddr_) = make_uint4(pw0, pw1, pw2, (w_));              \
            pw0 = (kq_) == 0u ? (w_) : pw0;                                                     \
            pw1 = (kq_) == 1u ? (w_) : pw1;                                                     \
            pw2 = (kq_) == 2u ? (w_) : pw2;                                                     \
            lastw = (w_);                                                                       \
        } while (0)
        const uint32_t mis = (uint32_t)((uintptr_t)frag & 15u);
        const uint8_t *gbase = frag - mis;                               // chunk 0
        const uint32_t nchunks = (mis + flen + 15u) >> 4;                // chunks that hold bytes of the fragment
        uint32_t *row = sm->ring[warp] + lane * FZ_RING_ROW_WORDS;
        const uint32_t row_s = (uint32_t)__cvta_generic_to_shared(row);
        uint64_t acc = 0;
        uint32_t nxt = 0, rp = 0, fetched = 0;
        int nacc = 0;
        if (fast) {
            const uint32_t abs_bit = mis * 8u + (uint32_t)((int64_t)flen * 8 - inf.br.bits_left());
            rp = abs_bit >> 5;
            fetched = rp >> 2;
#pragma unroll
            for (uint32_t q = 0; q < FZ_RING_CHUNKS; q++) {
                if (fetched < nchunks) fz_cp_async16(row_s + (fetched & (FZ_RING_CHUNKS - 1)) * 16u, gbase + (size_t)fetched * 16u);
                fetched++;
            }
        }
        fz_cp_async_commit();
        fz_cp_async_wait<0>();
        if (fast) {
            const uint32_t abs_bit = mis * 8u + (uint32_t)((int64_t)flen * 8 - inf.br.bits_left());
            acc = (uint64_t)(row[rp & (FZ_RING_CHUNKS * 4 - 1)] >> (abs_bit & 31u));
            nacc = 32 - (int)(abs_bit & 31u);
            rp++;
            nxt = row[rp & (FZ_RING_CHUNKS * 4 - 1)];
        }
        while (__any_sync(0xffffffffu, live)) {
            if (fast) {
                // chunks below this one are used up -- keeping the two words the bit buffer may still hold bits of:
                // after a symbol taken by the general inflater the lane re-reads its buffer from the ring
                const uint32_t cons = (rp >= 2u ? rp - 2u : 0u) >> 2;
#pragma unroll
                for (int q = 0; q < FZ_RING_TOPUPS; q++) {
                    if (fetched < cons + FZ_RING_CHUNKS) {
                        if (fetched < nchunks) fz_cp_async16(row_s + (fetched & (FZ_RING_CHUNKS - 1)) * 16u, gbase + (size_t)fetched * 16u);
                        fetched++;
                    }
                }
            }
            fz_cp_async_commit();
            fz_cp_async_wait<1>();
            if (live && !fast) live = inf.step_lut(lut);   // end of block, the closing stored block, odd layouts
            else if (live) {
                // literals never overrun the output inside a round: it is cut to what fits (3 bytes per iteration)
                const uint32_t room = inf.bw.cap - inf.bw.op;
                const int iters = room >= 3u * FZ_FAST_ITERS ? FZ_FAST_ITERS : (int)(room / 3u);
                bool slow = iters == 0;
#pragma unroll 1
                for (int it = 0; it < iters; ++it) {
                    // refill without a branch: ORing the next word in early is harmless (its bits land where they
                    // belong and are ORed there again once they count), and `nxt` is re-read every iteration
                    acc |= (uint64_t)nxt << nacc;
                    const bool need = nacc < 32;
                    nacc += need ? 32 : 0;
                    rp += need ? 1u : 0u;
                    nxt = row[rp & (FZ_RING_CHUNKS * 4 - 1)];
                    const uint32_t e = lut[(uint32_t)acc & (FZ_GLUT_SIZE - 1)];
                    if (e & 0x100u) {
                        // a run (distance-1 match) whole: length code, its extra bits, the one distance bit
                        if (!(e & FZ_LUT_MATCH) || run_bit > 1u) { slow = true; break; }
                        const uint32_t cl = (e >> 25) & 15u, xb = (e >> 18) & 7u;
                        const uint32_t a = (uint32_t)(acc >> cl);
                        const uint32_t len = ((e >> 9) & 511u) + (a & ((1u << xb) - 1u));
                        if (((a >> xb) & 1u) != run_bit || inf.bw.op + len > inf.bw.cap || inf.bw.produced() == 0) { slow = true; break; }
                        acc >>= (cl + xb + 1u);
                        nacc -= (int)(cl + xb + 1u);
                        {   // len copies of the byte before, through the pending-word logic (no read of what was written)
                            const uint32_t r = inf.bw.op & 3u;
                            const uint32_t c = r ? (inf.bw.ow >> ((r - 1u) * 8u)) & 0xffu : lastw >> 24;
                            const uint32_t cw = c * 0x01010101u;
                            uint32_t left = len;
                            if (r) {   // complete the pending word first
                                const uint32_t take = left < 4u - r ? left : 4u - r;
                                inf.bw.ow |= (cw & (0xffffffffu >> (32u - 8u * take))) << (8u * r);
                                left -= take;
                                if (r + take == 4u) {
                                    FZ_WORD_DONE(inf.bw.ow, (inf.bw.op >> 2) & 3u, inf.bw.out + (inf.bw.op & ~15u));
                                    inf.bw.ow = 0;
                                }
                                inf.bw.op += take;
                            }
                            while (left >= 4u) {
                                const uint32_t kq = (inf.bw.op >> 2) & 3u;
                                if (kq == 0u && left >= 16u) {
                                    *(uint4 *)(inf.bw.out + inf.bw.op) = make_uint4(cw, cw, cw, cw);
                                    lastw = cw;
                                    inf.bw.op += 16u;
                                    left -= 16u;
                                } else {
                                    FZ_WORD_DONE(cw, kq, inf.bw.out + (inf.bw.op & ~15u));
                                    inf.bw.op += 4u;
                                    left -= 4u;
                                }
                            }
                            if (left) { inf.bw.ow = cw & (0xffffffffu >> (32u - 8u * left)); inf.bw.op += left; }
                        }
                        if (inf.bw.cap - inf.bw.op < 3u * (uint32_t)(iters - it)) break;   // the round's budget is gone
                        continue;
                    }
                    if (e == 0) {
                        // a code longer than the table's index: canonical search; literals stay in the loop
                        uint32_t idx;
                        const int l = fz_decode_idx(sm->LL, (uint32_t)acc & 0x7fffu, idx);
                        if (l == 0 || idx >= 288u) { slow = true; break; }
                        const uint32_t sym = tab.L((int)idx);
                        if (sym >= 256u) { slow = true; break; }
                        acc >>= l;
                        nacc -= l;
                        {   // one byte through the same pending-word logic as below
                            const uint32_t sh1 = (inf.bw.op & 3u) * 8u;
                            const uint32_t t1 = inf.bw.ow | (sym << sh1);
                            if (sh1 == 24u) {
                                FZ_WORD_DONE(t1, (inf.bw.op >> 2) & 3u, inf.bw.out + (inf.bw.op & ~15u));
                                inf.bw.ow = 0;
                            } else inf.bw.ow = t1;
                            inf.bw.op += 1;
                        }
                        continue;
                    }
                    // 1..3 literals: sym1 | sym2 << 8 | sym3 << 16 (unused slots are zero and lie above the bytes that
                    // count) appended to the pending word; a completed word is stored (op0 == 0 on this path)
                    const uint32_t cnt = e >> 29, tl = (e >> 25) & 15u;
                    const uint32_t v = (e & 255u) | ((e >> 1) & 0xffff00u);
                    const uint32_t sh = (inf.bw.op & 3u) * 8u;
                    const uint64_t t = (uint64_t)inf.bw.ow | ((uint64_t)v << sh);
                    const uint32_t kq = (inf.bw.op >> 2) & 3u;     // place of the pending word in its 16-byte group
                    inf.bw.op += cnt;
                    const bool full = sh + cnt * 8u >= 32u;
                    if (full && kq == 3u) *(uint4 *)(inf.bw.out + ((inf.bw.op & ~3u) - 16u)) = make_uint4(pw0, pw1, pw2, (uint32_t)t);
                    pw0 = (full && kq == 0u) ? (uint32_t)t : pw0;
                    pw1 = (full && kq == 1u) ? (uint32_t)t : pw1;
                    pw2 = (full && kq == 2u) ? (uint32_t)t : pw2;
                    lastw = full ? (uint32_t)t : lastw;
                    inf.bw.ow = full ? (uint32_t)(t >> 32) : (uint32_t)t;
                    acc >>= tl;
                    nacc -= (int)tl;
                }
                if (slow) {
                    FZ_FLUSH_PW();
                    // something else (a long length code, end of block, end of the output, an error): the general
                    // inflater takes this one symbol at the lane's bit position, then the ring reader resumes behind it
                    const int64_t rel = (int64_t)rp * 32 - nacc - (int64_t)mis * 8;
                    if (rel < 0 || rel > (int64_t)flen * 8) { bad = true; live = false; fast = false; }
                    else {
                        const uint32_t nby = (uint32_t)(rel >> 3);
                        inf.br.init(frag + nby, flen - nby);
                        inf.br.refill();
                        inf.br.drop((int)(rel & 7));
                        live = inf.step_lut(lut);
                        fast = false;
                        if (live && inf.in_body) {
                            const int64_t left = inf.br.bits_left();
                            const uint32_t abs_bit = mis * 8u + (uint32_t)((int64_t)flen * 8 - left);
                            const uint32_t r0 = abs_bit >> 5;
                            // (one symbol is at most 48 bits: still inside what the ring holds)
                            if (left >= 0 && ((r0 + 2) >> 2) < fetched && (r0 >> 2) + FZ_RING_CHUNKS >= fetched) {
                                fast = true;
                                FZ_RELOAD_PW();
                                rp = r0;
                                acc = (uint64_t)(row[rp & (FZ_RING_CHUNKS * 4 - 1)] >> (abs_bit & 31u));
                                nacc = 32 - (int)(abs_bit & 31u);
                                rp++;
                                nxt = row[rp & (FZ_RING_CHUNKS * 4 - 1)];
                            }
                        }
                    }
                }
            }
        }
    }
    if (valid && !all_zero && !done_stored) {
        uint32_t out_n = 0;
        size_t used = 0;
        const int rc = inf.finish(&out_n, &used);
        // `used` is relative to the (possibly re-based) reader: recompute the absolute end
        const int64_t left = inf.br.bits_left();
        if (bad || rc != FZ_INF_OK || out_n != expect || left != 0) atomicExch(&stream_fail[s], 1u);
    }
}

// =================================================================================================
// The lean group inflater: what decodes our own streams.  The kernel above carries the whole resumable inflater in
// registers (102 per thread: 4 CTAs = 16 warps per SM) although all but a few thousand of a sub-block's symbols are table
// hits.  Here the work is cut in two:
//   fz_inflate_prep_kernel  one warp per code group: finds the group's first coded sub-block, parses its block header
//                           with the general inflater and leaves the literal/length code (15 packed limits + sorted
//                           symbols), the header's bit count and the run bit in a 656-byte descriptor;
//   fz_inflate_lean_kernel  one CTA per code group at a time: builds the 12-bit table from the descriptor, every lane checks
//                           that its own header bits equal the leader's and then runs nothing but the table loop: literals
//                           (1..3 per hit), distance-1 runs, and the end of block followed by the closing empty stored
//                           block, all validated in place.  No general state machine; the encoder gives every symbol its
//                           sample saw a code of at most FZ_MAX_CODE_BITS = FZ_GLUT_BITS bits, the rare longer ones are
//                           turned into a table entry on the spot (fz_lean_long_entry).
// Anything else -- a second coded block, a distance code that is not the 1-bit run code, a parse or size error -- marks
// the GROUP (desc.state = 2) and the kernel above decodes it again from scratch; what that one refuses goes to the general
// inflater, as before.
// =================================================================================================

__global__ void __launch_bounds__(FZ_WARP)
fz_inflate_prep_kernel(const uint8_t *__restrict__ container, FzBatchGeom g, const unsigned long long *__restrict__ stream_off,
                       const uint32_t *__restrict__ stream_cnt, uint32_t hits_per_stream, const uint32_t *__restrict__ hits,
                       const uint32_t *__restrict__ stream_mode, FzGroupDesc *__restrict__ desc, uint32_t *__restrict__ qctr,
                       const uint8_t *planes, uint32_t full_only, const FzStatus *status)
{
    __shared__ uint16_t tabs[FZ_INF_TAB_U16];
    __shared__ FzCode codes[2];
    __shared__ uint32_t res[4];
    const int lane = threadIdx.x;
    FzGroupDesc *d = desc + blockIdx.x;
    if (blockIdx.x == 0) for (uint32_t i = lane; i < FZ_SM_COUNT; i += FZ_WARP) qctr[i] = 0;   // the lean kernel's work queues
    const uint32_t cps = fz_groups_per_stream(g);
    const uint32_t s = blockIdx.x / cps, ck = blockIdx.x - s * cps;
    const uint32_t m = stream_cnt[s];
    if (status->error || (stream_mode[s] & 0xffu) != 1u || ck * FZ_CODE_SUBS >= m) {
        if (lane == 0) d->state = 0;
        return;
    }
    const size_t h0 = (size_t)s * hits_per_stream;
    const uint8_t *base = container + stream_off[s];
    uint32_t leader = ~0u;
#pragma unroll
    for (uint32_t j = 0; j < FZ_CODE_SUBS / FZ_WARP; j++) {
        const uint32_t k = ck * FZ_CODE_SUBS + j * FZ_WARP + (uint32_t)lane;
        bool coded = false;
        if (k < m) {
            const uint32_t start = k ? hits[h0 + k - 1] + 4 : 0u, end = hits[h0 + k] + 4;
            coded = end > start && (base[start] & 7u) == 4u;      // BFINAL = 0, BTYPE = 10
        }
        const uint32_t mask = __ballot_sync(0xffffffffu, coded);
        if (mask && leader == ~0u) leader = ck * FZ_CODE_SUBS + j * FZ_WARP + (uint32_t)(__ffs((int)mask) - 1);
    }
    if (lane == 0) {
        uint32_t state = full_only ? 2u : 1u, hdr_bits = 0, run_bit = 2;
        const uintptr_t o = (uintptr_t)(planes + (uint64_t)(s & 3) * g.plane_stride + (uint64_t)(s >> 2) * g.chk);
        if (o & 15u) state = 2;                                   // the lean writer stores 16 bytes at a time
        if (leader != ~0u && state == 1u) {
            const uint32_t start = leader ? hits[h0 + leader - 1] + 4 : 0u, end = hits[h0 + leader] + 4;
            typedef FzInfTab<1> Tab;
            Tab tab{tabs, tabs + 288, tabs + 320};
            FzInflater<Tab> inf;
            inf.start(base + start, end - start, (uint8_t *)nullptr, FZ_SUB, tab);
            inf.bind_codes(&codes[0], &codes[1]);
            const bool okh = inf.block_header() && inf.in_body && inf.rc == FZ_INF_OK;
            if (!okh || inf.ll_left != 0) state = 2;               // (an incomplete code: not ours)
            else {
                hdr_bits = (uint32_t)((int64_t)(end - start) * 8 - inf.br.bits_left());
                run_bit = fz_dd1_run_bit(inf.dd1);
            }
        }
        res[0] = state; res[1] = hdr_bits; res[2] = run_bit; res[3] = leader;
    }
    __syncwarp();
    if (lane < 4) ((uint32_t *)d)[lane] = res[lane];
    if (res[0] == 1u && leader != ~0u) {
        const uint32_t *ll = (const uint32_t *)&codes[0];
        uint32_t *dw = (uint32_t *)d + FZ_DESC_CODE_WORD0;
        for (uint32_t i = lane; i < 15u; i += FZ_WARP) dw[i] = ll[i];
        const uint32_t *sy = (const uint32_t *)tabs;              // 288 sorted literal/length symbols = 144 words
        for (uint32_t i = lane; i < 144u; i += FZ_WARP) dw[16u + i] = sy[i];
    }
}

#ifndef FZ_LEAN_MINBLOCKS
#define FZ_LEAN_MINBLOCKS 7          // 7 CTAs x 4 warps per SM (<= 72 registers, <= 31.4 KB of shared memory each): the two coded
                                     // planes of a 4 GiB volume (1026 code groups) are resident at once on 148 SMs
#endif
#define FZ_LEAN_CHUNKS 7u            // 16-byte chunks of input per lane (the ring is indexed modulo 7 by wrap-around counters)
#define FZ_LEAN_ROW_WORDS (FZ_LEAN_CHUNKS * 4u)
#define FZ_LEAN_TOPUPS 3             // chunks a round can use up
#define FZ_LEAN_ITERS 12             // table hits per round: 12 x 21 bits (15-bit code, 5 extra bits, distance bit) < 8 words, so
                                     // a round reads at most chunk cons + 3 while chunks < cons + 4 have landed (see the loop)
struct FzSymTab {
    const uint16_t *ll;
    __device__ __forceinline__ uint16_t L(int i) const { return ll[i]; }
};
struct FzLeanSmem {
    alignas(16) uint32_t lut[FZ_GLUT_SIZE];
    alignas(16) uint32_t ring[FZ_INF_WARPS * FZ_WARP * FZ_LEAN_ROW_WORDS];   // every lane's window on its fragment (cp.async)
    alignas(16) uint32_t code[FZ_DESC_CODE_WORDS];                           // FzCode LL, pad, 288 sorted symbols
    uint32_t item, pad[3];                                                   // the code group the CTA works on
};
static_assert(sizeof(FzLeanSmem) + 1024 <= (227 * 1024) / FZ_LEAN_MINBLOCKS, "shared memory of FZ_LEAN_MINBLOCKS CTAs per SM");

// The lean kernel's table entry (the ALU pipe is what bounds the kernel: every field is where one instruction finds it):
//   bits 28..31  code bits the entry consumes (<= 15)
//   bits 24..25  cnt: 1..3 literals, their bytes in bits 0..23 (unused slots zero); 0: not a literal --
//   bit  26      ... a length symbol: base length in bits 0..8, number of extra bits in bits 9..11
//   bit  27      ... the end-of-block symbol
//   0            no code (of <= FZ_GLUT_BITS bits) for this bit pattern
#define FZ_LE_MATCH (1u << 26)
#define FZ_LE_EOB (1u << 27)
__device__ __forceinline__ uint32_t fz_lean_entry_from(uint32_t x)   // from the entry format of fz_lut_entry_bits
{
    if (x == 0u) return 0u;
    const uint32_t s1 = x & 511u, tl = (x >> 25) & 15u;
    if (s1 < 256u) return ((x & 255u) | ((x >> 1) & 0xffff00u)) | ((x >> 29) << 24) | (tl << 28);
    if (x & FZ_LUT_MATCH) return ((x >> 9) & 511u) | (((x >> 18) & 7u) << 9) | FZ_LE_MATCH | (tl << 28);
    return s1 == FZ_EOB ? (FZ_LE_EOB | (tl << 28)) : 0u;
}

// Table entry of a code longer than the table's index, decoded the canonical way: one symbol, its real length (<= 15)
// in the length field.  Only symbols the encoder's sample never saw have such codes.  0 = no such code.
__device__ __noinline__ uint32_t fz_lean_long_entry(const uint32_t *code, uint32_t bits15)
{
    uint32_t idx;
    const int l = fz_decode_idx(*(const FzCode *)code, bits15, idx);
    if (l == 0 || idx >= 288u) return 0u;
    const uint32_t s1 = ((const uint16_t *)(code + 16))[idx];
    uint32_t ent = FZ_LUT_ENTRY(s1, 0, 0, l, 1);
    if (s1 >= 257u && s1 <= 285u) ent |= FZ_LUT_MATCH | (fz_len_base(s1 - 257u) << 9) | (fz_len_extra_bits(s1 - 257u) << 18);
    return fz_lean_entry_from(ent);
}

__device__ __forceinline__ uint32_t fz_lds32(uint32_t saddr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}

// 32 stream bits from bit position `pos` of the aligned words b32[0 .. nw): zero bits past the end
__device__ __forceinline__ uint32_t fz_peek32(const uint32_t *b32, uint32_t nw, uint32_t pos)
{
    const uint32_t i = pos >> 5;
    const uint32_t w0 = i < nw ? b32[i] : 0u, w1 = i + 1 < nw ? b32[i + 1] : 0u;
    return __funnelshift_r(w0, w1, pos & 31u);
}

// one code group (stream s, group ck of the stream), all threads of the CTA
__device__ __forceinline__ void
fz_lean_group(FzLeanSmem *sm, const uint32_t s, const uint32_t ck, FzGroupDesc *d, const uint8_t *__restrict__ container,
              const FzBatchGeom &g, const unsigned long long *__restrict__ stream_off, const uint32_t *__restrict__ stream_cnt,
              uint32_t hits_per_stream, const uint32_t *__restrict__ hits, uint32_t *__restrict__ zero_flags,
              uint8_t *__restrict__ planes)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t m = stream_cnt[s];
    const size_t h0 = (size_t)s * hits_per_stream;
    const uint32_t hdr_bits = d->hdr_bits, run_bit = d->run_bit, leader = d->leader;
    const bool any_coded = leader != ~0u;                 // CTA-uniform
    if (any_coded) {
        const uint32_t *dw = (const uint32_t *)d + FZ_DESC_CODE_WORD0;
        for (uint32_t i = threadIdx.x; i < FZ_DESC_CODE_WORDS; i += FZ_INF_WARPS * FZ_WARP) sm->code[i] = dw[i];
        __syncthreads();
        const FzCode &LL = *(const FzCode *)sm->code;
        const FzSymTab tab{(const uint16_t *)(sm->code + 16)};
        for (uint32_t e = threadIdx.x; e < FZ_GLUT_SIZE; e += FZ_INF_WARPS * FZ_WARP) sm->lut[e] = fz_lean_entry_from(fz_lut_entry_bits<FZ_GLUT_BITS>(LL, tab, e));
        __syncthreads();
    }
    const uint32_t gk = ck * FZ_CODE_WARPS + (uint32_t)warp;
    if (gk * FZ_GROUP_SUBS >= m) return;                  // warp-uniform: nothing of the stream left for this warp
    const uint32_t k = gk * FZ_GROUP_SUBS + lane;
    const bool valid = k < m;
    const uint32_t n_s = fz_chunk_n(g, s / FZ_PLANES);
    uint32_t start = 0, end = 0, cap = 0;
    uint8_t *out = nullptr;
    if (valid) {
        start = k ? hits[h0 + k - 1] + 4 : 0u;
        end = hits[h0 + k] + 4;
        const uint32_t obeg = k << FZ_SUB_LOG2;
        cap = min((uint32_t)FZ_SUB, n_s - obeg);
        out = planes + (uint64_t)(s & 3) * g.plane_stride + (uint64_t)(s >> 2) * g.chk + obeg;
    }
    const uint8_t *sbase = container + stream_off[s];
    const uint8_t *frag = sbase + start;
    const uint32_t flen = end - start;
    uint32_t first3 = 7;
    if (valid && flen >= 1) first3 = frag[0] & 7u;        // BFINAL (must be 0) | BTYPE << 1
    const bool coded = valid && first3 == 4u;
    bool failed = valid && first3 != 4u && first3 != 0u;  // fixed-code or final blocks: not ours
    if (coded && k != leader) {
        // the first hdr_bits of the fragment must equal the leader's: then the group's table is this sub-block's
        const uint8_t *lf = sbase + (leader ? hits[h0 + leader - 1] + 4 : 0u);
        const uint32_t nby = hdr_bits >> 3, rem = hdr_bits & 7u;
        bool same = (uint64_t)flen * 8 > hdr_bits;
        if (same) {
            for (uint32_t i = 0; i < nby; i++) same &= frag[i] == lf[i];
            if (rem) same &= ((frag[nby] ^ lf[nby]) & ((1u << rem) - 1u)) == 0;
        }
        failed |= !same;
    }
    // stored sub-blocks (two stored blocks and the empty one): copied by the whole warp, 16 bytes per lane and step
    {
        uint32_t stored_mask = __ballot_sync(0xffffffffu, valid && first3 == 0u);
        while (stored_mask) {
            const int j = __ffs((int)stored_mask) - 1;
            stored_mask &= stored_mask - 1;
            const uint8_t *f = (const uint8_t *)(uintptr_t)__shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)frag, j);
            uint8_t *o = (uint8_t *)(uintptr_t)__shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)out, j);
            const uint32_t fl = __shfl_sync(0xffffffffu, flen, j), ex = __shfl_sync(0xffffffffu, cap, j);
            uint32_t pos = 0, prod = 0;
            bool ok = false;
            for (int blk = 0; blk < 4; blk++) {   // warp-uniform: every lane reads the same header bytes
                if (pos + 5 > fl) break;
                const uint32_t b0 = f[pos], len = (uint32_t)f[pos + 1] | ((uint32_t)f[pos + 2] << 8);
                const uint32_t nlen = (uint32_t)f[pos + 3] | ((uint32_t)f[pos + 4] << 8);
                if ((b0 & 7u) != 0u || (len ^ 0xffffu) != nlen) break;      // BFINAL = 0, BTYPE = 00, LEN = ~NLEN
                if (len == 0) { ok = pos + 5 == fl && prod == ex; break; }   // the empty block must close the fragment
                if (prod + len > ex || pos + 5 + len > fl) break;
                fz_warp_copy(o + prod, f + pos + 5, len, f + fl, lane);
                pos += 5 + len;
                prod += len;
            }
            if (lane == j && !ok) failed = true;
        }
    }
    const uint32_t *lut = sm->lut;
    const uint32_t mis = (uint32_t)((uintptr_t)frag & 15u);
    const uint32_t endbit = (mis + flen) * 8u;            // ring coordinates: bit 0 = first bit of the 16-byte chunk of frag[0]
    // Tiny fragments are almost always sub-blocks of zero bytes (mask bits >= 8 zero whole byte planes): walk the tokens
    // without storing; if the fragment is valid and all zero it is only flagged -- the merge supplies the zeros.
    bool all_zero = false;
    if (zero_flags && coded && !failed && flen <= FZ_ZERO_PROBE_BYTES) {
        const uint32_t sk = (uint32_t)((uintptr_t)frag & 3u);
        const uint32_t *b32 = (const uint32_t *)(frag - sk);
        const uint32_t nw = (sk + flen + 3u) >> 2, zend = (sk + flen) * 8u;
        uint32_t pos = sk * 8u + hdr_bits, prod = 0;
        bool ok = false;
        while (pos < zend) {
            const uint32_t w = fz_peek32(b32, nw, pos);
            uint32_t e = lut[w & (FZ_GLUT_SIZE - 1)];
            if (e == 0) e = fz_lean_long_entry(sm->code, w & 0x7fffu);
            if (e == 0) break;
            const uint32_t cnt = (e >> 24) & 3u, tl = e >> 28;
            if (cnt == 0u) {
                if (e & FZ_LE_EOB) {                      // the end of block and the closing empty stored block
                    const uint32_t after = pos + tl + 3u;
                    ok = ((w >> tl) & 7u) == 0u && ((after + 7u) & ~7u) + 32u == zend && prod == cap;
                    break;
                }
                const uint32_t xb = (e >> 9) & 7u, a = w >> tl;
                const uint32_t len = (e & 511u) + (a & ((1u << xb) - 1u));
                if (((a >> xb) & 1u) != run_bit || prod == 0 || prod + len > cap) break;
                pos += tl + xb + 1u;
                prod += len;
            } else {
                if ((e & 0xffffffu) != 0u || prod + cnt > cap) break;   // a byte that is not zero
                pos += tl;
                prod += cnt;
            }
        }
        all_zero = ok;
    }
    if (zero_flags && valid) zero_flags[(size_t)s * g.nsub_full + k] = all_zero ? 1u : 0u;

    // ---- the table loop.  The lane reads its fragment through a ring of FZ_LEAN_CHUNKS 16-byte chunks in shared memory
    // that cp.async tops up once per round, two rounds ahead of the decoder: `cp.async.wait_group 1` leaves only the
    // newest top-up in flight.  Why a round never reads what has not landed: with cons = (rp - 2) / 4 the first chunk still
    // in use at the start of a round, rp <= 4 cons + 5; a round advances rp by at most 8 words (FZ_LEAN_ITERS), so it reads
    // words <= 4 cons + 13: chunk cons + 3.  The previous round left fetched = cons' + 7 >= cons + 4 (cons moves by at
    // most 3 chunks per round, which FZ_LEAN_TOPUPS = 3 makes up), and all of that has landed.  Completed output words
    // wait in pw0..pw2 until the fourth of their 16-byte group arrives and leave as one 128-bit store.
    uint32_t live = (coded && !failed && !all_zero) ? 1u : 0u;   // (words, not bools: the compiler packed bools into one register
    uint32_t done_ok = 0;                                        //  and re-packed it in every iteration)
    // Output: `ow` collects the bytes of the word being written; completed words wait in pw0, pw1, pw2 (oldest first: a
    // new word shifts them along -- three predicated moves, no selects) until the fourth of their 16-byte group arrives
    // and the group leaves as one 128-bit store.  pw2 is always the last completed word: a run repeats its top byte.
    uint32_t op = 0, ow = 0, pw0 = 0, pw1 = 0, pw2 = 0;
    {
        const uint8_t *gbase = frag - mis;                               // chunk 0
        const uint32_t nchunks = (mis + flen + 15u) >> 4;                // chunks that hold bytes of the fragment
        uint32_t *row = sm->ring + (warp * FZ_WARP + lane) * FZ_LEAN_ROW_WORDS;
        const uint32_t row_s = (uint32_t)__cvta_generic_to_shared(row);
        const uint32_t lut_s = (uint32_t)__cvta_generic_to_shared(sm->lut);
        // Bit reader: lo and hi are two consecutive words of the fragment, bo (0..31) the bits of lo already used -- one
        // funnel shift gives the next 32 stream bits, which is all a table hit, a run (<= 21 bits) or the end of block
        // (<= 18) needs.  nxt is the word behind hi, read from the ring one iteration before it can become hi; rp its
        // word index (ri = rp modulo the ring's words).  Selects are written as x += flag * (y - x) with flag 0 / 1: the
        // multiply-add pipe is idle in this kernel, the ALU pipe (shifts, logic, compares, selects) is what bounds it.
        const uint32_t abs_bit = mis * 8u + hdr_bits;
        uint32_t lo = 0, hi = 0, nxt = 0, bo = abs_bit & 31u;
        uint32_t rp = abs_bit >> 5, fetched = rp >> 2;
        uint32_t ri = rp % FZ_LEAN_ROW_WORDS;                            // rp modulo the ring's words
        uint32_t fs = fetched % FZ_LEAN_CHUNKS;                          // fetched modulo the ring's chunks
        int room = (int)cap;                                             // bytes the sub-block still takes
        uint32_t pend_n = 0, pend_op = 0, pend_cw = 0;                   // 16-byte groups of a long run left to the warp
        if (live) {
#pragma unroll
            for (uint32_t c = 0; c < FZ_LEAN_CHUNKS; c++) {
                if (fetched < nchunks) fz_cp_async16(row_s + fs * 16u, gbase + (size_t)fetched * 16u);
                fetched++;
                fs = fs + 1u == FZ_LEAN_CHUNKS ? 0u : fs + 1u;
            }
        }
        fz_cp_async_commit();
        fz_cp_async_wait<0>();
        if (live) {
            lo = row[ri];
            ri = ri + 1u == FZ_LEAN_ROW_WORDS ? 0u : ri + 1u;
            hi = row[ri];
            ri = ri + 1u == FZ_LEAN_ROW_WORDS ? 0u : ri + 1u;
            nxt = row[ri];
            rp += 2u;
        }
        // word w_ completes the place (op >> 2) & 3 of its 16-byte group
#define FZ_LEAN_WORD_DONE(w_)                                                                   \
        do {                                                                                    \
            if ((op & 12u) == 12u) *(uint4 *)(out + (op & ~15u)) = make_uint4(pw0, pw1, pw2, (w_)); \
            pw0 = pw1; pw1 = pw2; pw2 = (w_);                                                   \
        } while (0)
        // tl_ bits are used up: move on by a word when lo is spent (flag arithmetic, see above)
#define FZ_LEAN_CONSUME(tl_)                                                                    \
        do {                                                                                    \
            bo += (tl_);                                                                        \
            const uint32_t ni_ = bo >> 5;                                                       \
            bo &= 31u;                                                                          \
            lo += ni_ * (hi - lo);                                                              \
            hi += ni_ * (nxt - hi);                                                             \
            rp += ni_;                                                                          \
            ri += ni_;                                                                          \
            ri = ri == FZ_LEAN_ROW_WORDS ? 0u : ri;                                             \
            nxt = row[ri];                                                                      \
        } while (0)
        while (__any_sync(0xffffffffu, live != 0u)) {
            if (live) {
                // chunks below this one are used up (lo is word rp - 2)
                const uint32_t cons = (rp - 2u) >> 2;
#pragma unroll
                for (int c = 0; c < FZ_LEAN_TOPUPS; c++) {
                    if (fetched < cons + FZ_LEAN_CHUNKS) {
                        if (fetched < nchunks) fz_cp_async16(row_s + fs * 16u, gbase + (size_t)fetched * 16u);
                        fetched++;
                        fs = fs + 1u == FZ_LEAN_CHUNKS ? 0u : fs + 1u;
                    }
                }
            }
            fz_cp_async_commit();
            fz_cp_async_wait<1>();
            if (live) {
#pragma unroll 1
                for (int it = 0; it < FZ_LEAN_ITERS; ++it) {
                    const uint32_t win = __funnelshift_r(lo, hi, bo);      // the next 32 stream bits
                    uint32_t e = fz_lds32(lut_s + ((win & (FZ_GLUT_SIZE - 1)) << 2));   // (a shared-space load: through the generic
                    uint32_t cnt = (e >> 24) & 3u;                                      //  pointer the window base was rebuilt per iteration)
                    if (cnt == 0u) {
                        if (e == 0u) {   // a symbol the encoder's sample never saw (a code longer than the table's index), or garbage
                            e = fz_lean_long_entry(sm->code, win & 0x7fffu);
                            cnt = (e >> 24) & 3u;
                            if (e == 0u) { live = 0u; break; }
                        }
                        if (cnt == 0u) {
                            const uint32_t tl = e >> 28;
                            if (e & FZ_LE_EOB) {
                                // end of block: what follows must be the empty stored block (000, pad to the byte, 00 00 FF FF
                                // -- the marker the scan found) that closes the fragment, and the sub-block must be complete
                                const uint32_t after = (rp - 2u) * 32u + bo + tl + 3u;
                                done_ok = (((win >> tl) & 7u) == 0u && ((after + 7u) & ~7u) + 32u == endbit && room == 0) ? 1u : 0u;
                                live = 0u;
                                break;
                            }
                            // a run (distance-1 match) whole: length code, its extra bits, the one distance bit
                            const uint32_t xb = (e >> 9) & 7u;
                            const uint32_t a = win >> tl;
                            const uint32_t len = (e & 511u) + (a & ((1u << xb) - 1u));
                            room -= (int)len;
                            if (((a >> xb) & 1u) != run_bit || room < 0 || op == 0u) { live = 0u; break; }
                            FZ_LEAN_CONSUME(tl + xb + 1u);
                            {   // len copies of the byte before, through the pending-word logic (no read of what was written)
                                const uint32_t r = op & 3u;
                                const uint32_t c = r ? (ow >> ((r - 1u) * 8u)) & 0xffu : pw2 >> 24;
                                const uint32_t cw = c * 0x01010101u;
                                uint32_t left = len;
                                if (r) {   // complete the pending word first
                                    const uint32_t take = left < 4u - r ? left : 4u - r;
                                    ow |= (cw & (0xffffffffu >> (32u - 8u * take))) << (8u * r);
                                    left -= take;
                                    if (r + take == 4u) {
                                        FZ_LEAN_WORD_DONE(ow);
                                        ow = 0;
                                    }
                                    op += take;
                                }
                                while (left >= 4u && (op & 12u) != 0u) {   // words up to the next 16-byte boundary
                                    FZ_LEAN_WORD_DONE(cw);
                                    op += 4u;
                                    left -= 4u;
                                }
                                if (left >= 16u) {
                                    // whole 16-byte groups, one store each: by this lane -- or, when there are several
                                    // and the lane's slot is free, by the whole warp at the end of the round (a lane
                                    // on its own walks them with 1.3 lanes of 32 active; planes that keep one exponent
                                    // for a while have runs of 30..258 bytes)
                                    const uint32_t nv = left >> 4;
                                    if (nv >= 3u && pend_n == 0u) { pend_n = nv; pend_op = op; pend_cw = cw; }
                                    else
                                        for (uint32_t q = 0; q < nv; q++) *(uint4 *)(out + op + 16u * q) = make_uint4(cw, cw, cw, cw);
                                    pw2 = cw;
                                    op += nv << 4;
                                    left &= 15u;
                                }
                                while (left >= 4u) {
                                    FZ_LEAN_WORD_DONE(cw);
                                    op += 4u;
                                    left -= 4u;
                                }
                                if (left) { ow = cw & (0xffffffffu >> (32u - 8u * left)); op += left; }
                            }
                            continue;
                        }
                    }
                    // 1..3 literals: sym1 | sym2 << 8 | sym3 << 16 (unused slots are zero and lie above the bytes that
                    // count) appended to the pending word
                    room -= (int)cnt;
                    if (room < 0) { live = 0u; break; }
                    const uint32_t v = e & 0xffffffu, tl = e >> 28;
                    const uint32_t sh = (op & 3u) * 8u;
                    const uint64_t t = (uint64_t)ow | ((uint64_t)v << sh);
                    const uint32_t t0 = (uint32_t)t, t1 = (uint32_t)(t >> 32);
                    const uint32_t fi = ((op & 3u) + cnt) >> 2;             // 1: the pending word is complete
                    if (fi != 0u && (op & 12u) == 12u) *(uint4 *)(out + (op & ~15u)) = make_uint4(pw0, pw1, pw2, t0);
                    pw0 += fi * (pw1 - pw0);
                    pw1 += fi * (pw2 - pw1);
                    pw2 += fi * (t0 - pw2);
                    ow = t0 + fi * (t1 - t0);
                    op += cnt;
                    FZ_LEAN_CONSUME(tl);
                }
            }
            // the long runs of this round: every lane stores one 16-byte group of each
            uint32_t pm = __ballot_sync(0xffffffffu, pend_n != 0u);
            while (pm) {
                const int j = __ffs((int)pm) - 1;
                pm &= pm - 1u;
                const uint32_t n_j = __shfl_sync(0xffffffffu, pend_n, j), cw_j = __shfl_sync(0xffffffffu, pend_cw, j);
                const uint32_t op_j = __shfl_sync(0xffffffffu, pend_op, j);
                uint8_t *o_j = (uint8_t *)(uintptr_t)__shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)out, j);
                if ((uint32_t)lane < n_j) *(uint4 *)(o_j + op_j + 16u * (uint32_t)lane) = make_uint4(cw_j, cw_j, cw_j, cw_j);
            }
            pend_n = 0u;
        }
#undef FZ_LEAN_CONSUME
#undef FZ_LEAN_WORD_DONE
        fz_cp_async_wait<0>();   // nothing of this group may still be landing in the ring when the CTA's next group fills it
    }
    if (coded && !failed && !all_zero) {
        if (done_ok) {
            // what is still pending: the completed words of the last 16-byte group (the newest in pw2), then the bytes of
            // the last word
            const uint32_t kq = (op >> 2) & 3u;
            uint32_t *q = (uint32_t *)(out + (op & ~15u));
            if (kq == 1u) q[0] = pw2;
            if (kq == 2u) { q[0] = pw1; q[1] = pw2; }
            if (kq == 3u) { q[0] = pw0; q[1] = pw1; q[2] = pw2; }
            const uint32_t r = op & 3u, w0 = op - r;
            for (uint32_t i = 0; i < r; i++) out[w0 + i] = (uint8_t)(ow >> (8u * i));
        } else failed = true;
    }
    if (failed) d->state = 2u;    // the full group kernel decodes this group again
}

// Work distribution.  A code group of a compressible plane keeps its CTA busy for about as long as the whole kernel
// runs (128 threads x 16 KiB, one symbol chain each), groups of RAW or all-zero planes cost nothing, and a 4 GiB volume
// has only 3.5 expensive groups per SM and coded plane: where they land decides the kernel's time.  Left to the block
// scheduler (one CTA per group) some SMs held 6 or 7 of them and others none.  So the grid is FZ_LEAN_MINBLOCKS resident
// CTAs per SM and the groups are dealt by hand: in PLANE-major order (groups of one plane cost the same) item i belongs
// to SM i mod 148, whose CTAs take them one after the other from the SM's counter; a CTA whose SM has nothing left helps
// itself from the other SMs' queues, so any placement of the CTAs works.
__device__ __forceinline__ uint32_t fz_lean_pull(uint32_t *qctr, uint32_t nitems, const FzGroupDesc *desc, const FzBatchGeom &g,
                                                 uint32_t cps)
{
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    smid %= FZ_SM_COUNT;
    for (uint32_t v = 0; v < FZ_SM_COUNT; v++) {
        const uint32_t j = smid + v >= FZ_SM_COUNT ? smid + v - FZ_SM_COUNT : smid + v;
        for (;;) {
            if (*(volatile uint32_t *)(qctr + j) * FZ_SM_COUNT + j >= nitems) break;   // (a look before the atomic: nothing left there)
            const uint32_t i = atomicAdd(qctr + j, 1u) * FZ_SM_COUNT + j;
            if (i >= nitems) break;
            const uint32_t q = i / cps, ck = i - q * cps;
            const uint32_t s = (q % g.nchunks) * FZ_PLANES + q / g.nchunks;
            if (desc[(size_t)s * cps + ck].state == 1u) return i;                       // everything else is not the lean kernel's
        }
    }
    return ~0u;
}

__global__ void __launch_bounds__(FZ_INF_WARPS * FZ_WARP, FZ_LEAN_MINBLOCKS)
fz_inflate_lean_kernel(const uint8_t *__restrict__ container, FzBatchGeom g, const unsigned long long *__restrict__ stream_off,
                       const uint32_t *__restrict__ stream_cnt, uint32_t hits_per_stream, const uint32_t *__restrict__ hits,
                       FzGroupDesc *__restrict__ desc, uint32_t *__restrict__ qctr, uint32_t nitems,
                       uint32_t *__restrict__ zero_flags, uint8_t *__restrict__ planes)
{
    extern __shared__ __align__(16) uint8_t fz_smem[];
    FzLeanSmem *sm = (FzLeanSmem *)fz_smem;
    const uint32_t cps = fz_groups_per_stream(g);
    for (;;) {
        __syncthreads();                                  // the previous group's table and ring are done with
        if (threadIdx.x == 0) sm->item = fz_lean_pull(qctr, nitems, desc, g, cps);
        __syncthreads();
        const uint32_t i = sm->item;
        if (i == ~0u) break;
        const uint32_t q = i / cps, ck = i - q * cps;
        const uint32_t s = (q % g.nchunks) * FZ_PLANES + q / g.nchunks;
        fz_lean_group(sm, s, ck, desc + (size_t)s * cps + ck, container, g, stream_off, stream_cnt, hits_per_stream, hits, zero_flags, planes);
    }
}

// ---- general path: one thread per stream (reference-made streams: back-to-back blocks, no byte alignment between them)
__global__ void __launch_bounds__(32)
fz_inflate_general_kernel(const uint8_t *__restrict__ container, FzBatchGeom g, const uint32_t *__restrict__ stream_hdr,
                          const unsigned long long *__restrict__ stream_off, const uint32_t *__restrict__ stream_mode,
                          const uint32_t *__restrict__ stream_fail, const uint32_t *__restrict__ par_ok,
                          uint32_t *__restrict__ zero_flags, uint8_t *__restrict__ planes, FzStatus *status)
{
    __shared__ uint16_t tabs[FZ_INF_TAB_U16];
    __shared__ uint32_t lut[FZ_LUT_SIZE];
    __shared__ FzCode codes[2];
    const uint32_t s = blockIdx.x;
    if (threadIdx.x != 0 || status->error) return;
    const uint32_t mode = stream_mode[s] & 0xffu;
    const bool failed = mode == 1u && stream_fail[s];
    if (!(mode == 2u || failed)) return;
    if (failed) atomicAdd(&status->n_fast_failed, 1u);
    atomicAdd(&status->n_general, 1u);
    if (mode == 2u && par_ok[s]) { atomicAdd(&status->n_blockpar, 1u); return; }   // decoded block-parallel
    if (zero_flags) for (uint32_t k = 0; k < g.nsub_full; k++) zero_flags[(size_t)s * g.nsub_full + k] = 0;   // all of it gets written
    const uint32_t len = stream_hdr[s] & ~FZ_RAW_FLAG;
    const uint32_t n_s = fz_chunk_n(g, s / FZ_PLANES);
    uint8_t *out = planes + (uint64_t)(s & 3) * g.plane_stride + (uint64_t)(s >> 2) * g.chk;
    FzInfTab<1> tab{tabs, tabs + 288, tabs + 320};
    uint32_t out_n = 0;
    size_t used = 0;
    const int rc = fz_inflate(container + stream_off[s], (size_t)len, out, n_s, tab, &out_n, &used, lut, codes);
    if (rc != FZ_INF_OK || out_n != n_s) atomicCAS(&status->error, 0, FZ_E_FORMAT);
}

// =================================================================================================
// block-parallel inflate of zlib-made streams (fz_blockpar.cuh): find -> sort -> measure -> chain -> write -> stored
// Every kernel is a fixed-size grid looping over a work list whose length lives on the device (ctl[0] =
// general streams of this batch), so a batch of our own streams costs six empty launches and no host sync.
// =================================================================================================
#define FZ_BP_FIND_THREADS 64
#define FZ_BP_SEG_WORDS 2048            // 32-bit words of a stream searched by one block per work item
#define FZ_BP_QCAP 96                   // per-warp queue of positions that passed the quick test

__device__ __forceinline__ uint32_t fz_bp_stream_n(const FzBatchGeom &g, uint32_t s)
{
    return fz_chunk_n(g, s / FZ_PLANES);
}

// candidates: every bit position p of the payload with a plausible dynamic-block header.
// Three stages: (1) all 32 positions of a word against the fixed header fields and the Kraft sum of the
// code-length code (registers + a 512-byte table); (2) survivors (~0.1 % of the positions) are queued per warp and,
// 32 at a time, walk the code-length data with a 128-byte per-thread table (fz_block_precheck); (3) what is left
// (real headers, practically) gets the inflater's own header parse, one lane at a time.
__global__ void __launch_bounds__(FZ_BP_FIND_THREADS)
fz_bp_find_kernel(const uint8_t *__restrict__ container, uint64_t container_size, const uint32_t *__restrict__ stream_hdr,
                  const unsigned long long *__restrict__ stream_off, FzBlockParBufs bp, uint32_t segs_per_stream,
                  const FzStatus *status)
{
    __shared__ uint16_t tabs[FZ_BP_FIND_THREADS / 32][FZ_INF_TAB_U16];   // one full-parse table per warp
    __shared__ uint32_t cl_luts[32 * FZ_BP_FIND_THREADS];                 // 128 bytes per thread, word-interleaved
    __shared__ uint32_t queue_p[FZ_BP_FIND_THREADS / 32][FZ_BP_QCAP];     // survivors of stage 1: bit position ...
    __shared__ uint32_t queue_s[FZ_BP_FIND_THREADS / 32][FZ_BP_QCAP];     // ... and stream (the queue outlives a segment)
    __shared__ uint32_t qn[FZ_BP_FIND_THREADS / 32];
    __shared__ uint8_t kraft_lut[512];
    const uint32_t ngen = bp.ctl[0];
    if (ngen == 0 || status->error) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const FzInfTab<1> tab{tabs[warp], tabs[warp] + 288, tabs[warp] + 320};
    const FzClLut<FZ_BP_FIND_THREADS> cl_lut{(uint8_t *)(cl_luts + threadIdx.x)};
    const uint32_t *alloc_end = (const uint32_t *)(((uintptr_t)container + container_size + 3u) & ~(uintptr_t)3u);
    for (uint32_t i = threadIdx.x; i < 512; i += FZ_BP_FIND_THREADS) kraft_lut[i] = (uint8_t)fz_kraft3(i);
    if (lane == 0) qn[warp] = 0;
    __syncthreads();
    // stages 2 and 3 for queue entries [i0, i0 + 32) of this warp (warp-wide call)
    auto validate = [&](uint32_t i0, uint32_t nq) {
        const bool have = i0 + lane < nq;
        const uint32_t p = have ? queue_p[warp][i0 + lane] : 0u;
        const uint32_t s = have ? queue_s[warp][i0 + lane] : 0u;
        const uint8_t *in = container + stream_off[s];
        const uint32_t len = stream_hdr[s] & ~FZ_RAW_FLAG;
        const bool pre = have && fz_block_precheck(in, (size_t)len, (uint64_t)p, cl_lut);
        uint32_t m = __ballot_sync(0xffffffffu, pre);
        while (m) {
            const int l = __ffs((int)m) - 1;
            m &= m - 1;
            if (lane == l && fz_block_candidate(in, (size_t)len, (uint64_t)p, tab)) {
                const uint32_t idx = atomicAdd(&bp.cand_cnt[s], 1u);
                if (idx < FZ_BP_CAP) bp.cand_pos[(size_t)s * FZ_BP_CAP + idx] = p;
            }
            __syncwarp();
        }
    };
    // drain the queue in batches of 32 (all of it when `all`), keep the remainder for later
    auto drain = [&](bool all) {
        __syncwarp();
        uint32_t n = qn[warp];
        if (n > FZ_BP_QCAP) n = FZ_BP_QCAP;
        uint32_t done = 0;
        while (done + 32 <= n || (all && done < n)) { validate(done, n); done += 32; }
        if (done) {
            __syncwarp();
            const uint32_t rest = done < n ? n - done : 0u;   // < 32 entries: move them to the front
            uint32_t mp = 0, ms = 0;
            if ((uint32_t)lane < rest) { mp = queue_p[warp][done + lane]; ms = queue_s[warp][done + lane]; }
            __syncwarp();
            if ((uint32_t)lane < rest) { queue_p[warp][lane] = mp; queue_s[warp][lane] = ms; }
            if (lane == 0) qn[warp] = rest;
        }
        __syncwarp();
    };
    for (uint32_t item = blockIdx.x; item < ngen * segs_per_stream; item += gridDim.x) {
        const uint32_t gi = item / segs_per_stream, seg = item - gi * segs_per_stream;
        const uint32_t s = bp.gen_list[gi];
        const uint32_t len = stream_hdr[s] & ~FZ_RAW_FLAG;
        const uint8_t *in = container + stream_off[s];
        const uint32_t skew = (uint32_t)((uintptr_t)in & 3u);
        const uint32_t *abase = (const uint32_t *)(in - skew);
        const uint32_t nwords = (len + skew + 3) >> 2;
        if (seg * FZ_BP_SEG_WORDS >= nwords) continue;   // block-uniform
        const int64_t total_bits = (int64_t)len * 8;
        for (uint32_t it = 0; it < FZ_BP_SEG_WORDS / FZ_BP_FIND_THREADS; it++) {
            const uint32_t wi = seg * FZ_BP_SEG_WORDS + it * FZ_BP_FIND_THREADS + threadIdx.x;
            if (wi < nwords) {
                uint32_t x[5];
#pragma unroll
                for (int k = 0; k < 5; k++) x[k] = (abase + wi + k < alloc_end) ? __ldg(abase + wi + k) : 0u;
                const uint64_t v = (uint64_t)x[0] | ((uint64_t)x[1] << 32);
                uint32_t m = (uint32_t)(~v & ~(v >> 1) & (v >> 2));   // BFINAL = 0, BTYPE = 10 at positions j = 0..31 of this word
                while (m) {
                    const int j = __ffs((int)m) - 1;
                    m &= m - 1;
                    const int64_t p = (int64_t)wi * 32 + j - (int64_t)skew * 8;
                    if (p < 0 || p + 17 > total_bits) continue;
                    const uint64_t lo = (uint64_t)__funnelshift_r(x[0], x[1], j) | ((uint64_t)__funnelshift_r(x[1], x[2], j) << 32);
                    const uint64_t hi = (uint64_t)__funnelshift_r(x[2], x[3], j) | ((uint64_t)__funnelshift_r(x[3], x[4], j) << 32);
                    if (!fz_block_quick_test(lo, hi, kraft_lut)) continue;
                    const uint32_t slot = atomicAdd(&qn[warp], 1u);
                    if (slot < FZ_BP_QCAP) { queue_p[warp][slot] = (uint32_t)p; queue_s[warp][slot] = s; }
                    // more survivors than the queue holds (pathological data): give the stream to the serial inflater
                    else atomicAdd(&bp.cand_cnt[s], FZ_BP_CAP + 1u);
                }
            }
            drain(false);
        }
    }
    drain(true);
}

// ascending candidate positions per stream (bitonic sort of FZ_BP_CAP slots, unused ones padded with ~0)
__global__ void __launch_bounds__(FZ_BP_CAP / 2)
fz_bp_sort_kernel(FzBlockParBufs bp, const FzStatus *status)
{
    __shared__ uint32_t v[FZ_BP_CAP];
    const uint32_t ngen = bp.ctl[0];
    if (ngen == 0 || status->error) return;
    for (uint32_t gi = blockIdx.x; gi < ngen; gi += gridDim.x) {
        const uint32_t s = bp.gen_list[gi];
        const uint32_t n = min(bp.cand_cnt[s], (uint32_t)FZ_BP_CAP);
        uint32_t *pos = bp.cand_pos + (size_t)s * FZ_BP_CAP;
        for (uint32_t i = threadIdx.x; i < FZ_BP_CAP; i += blockDim.x) v[i] = i < n ? pos[i] : 0xFFFFFFFFu;
        __syncthreads();
        for (uint32_t k = 2; k <= FZ_BP_CAP; k <<= 1)
            for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                const uint32_t t = threadIdx.x;
                const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
                const bool up = (i & k) == 0;
                const uint32_t a = v[i], b = v[l];
                if ((a > b) == up) { v[i] = b; v[l] = a; }
                __syncthreads();
            }
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) pos[i] = v[i];
        // work list of the measure pass: one entry per candidate (cand_cnt > FZ_BP_CAP: the stream goes serial)
        __shared__ uint32_t base;
        if (threadIdx.x == 0) base = bp.cand_cnt[s] <= FZ_BP_CAP ? atomicAdd(&bp.ctl[FZ_BP_CTL_NMEASURE], n) : 0xFFFFFFFFu;
        __syncthreads();
        if (base != 0xFFFFFFFFu)
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) bp.items[base + i] = s * FZ_BP_CAP + i;
        __syncthreads();
    }
}

// measure (WRITE = false): decode every candidate block without storing; write (WRITE = true): decode the blocks
// on the chain into the planes.  One WARP per block: the self-synchronising decoder of fz_blockpar.cuh.
#define FZ_BP_SY_WARPS 4
template <bool WRITE>
__global__ void __launch_bounds__(FZ_BP_SY_WARPS * 32, 5)
fz_bp_sync_kernel(const uint8_t *__restrict__ container, FzBatchGeom g, const uint32_t *__restrict__ stream_hdr,
                  const unsigned long long *__restrict__ stream_off, FzBlockParBufs bp, uint8_t *__restrict__ planes,
                  const FzStatus *status)
{
    __shared__ FzSyncState sm[FZ_BP_SY_WARPS];
    const uint32_t ngen = bp.ctl[0];
    if (ngen == 0 || status->error) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    FzSyncState *st = &sm[warp];
    // work list filled by the sort kernel (measure) / the chain kernel (write); warps take entries as they get free
    const uint32_t *list = WRITE ? bp.items + (size_t)bp.nstreams * FZ_BP_CAP : bp.items;
    const uint32_t total = bp.ctl[WRITE ? FZ_BP_CTL_NWRITE : FZ_BP_CTL_NMEASURE];
    uint32_t *cursor = &bp.ctl[WRITE ? FZ_BP_CTL_CUR_WRITE : FZ_BP_CTL_CUR_MEASURE];
    for (;;) {
        uint32_t wi = 0;
        if (lane == 0) wi = atomicAdd(cursor, 1u);
        wi = __shfl_sync(0xffffffffu, wi, 0);
        if (wi >= total) break;
        const size_t ci = list[wi];
        const uint32_t s = (uint32_t)(ci / FZ_BP_CAP);
        if (WRITE && !bp.par_ok[s]) continue;   // a block of this stream already failed: the serial inflater redoes it
        uint32_t off = 0;
        if (WRITE) off = bp.blk_off[ci];
        const uint32_t len = stream_hdr[s] & ~FZ_RAW_FLAG;
        const uint8_t *in = container + stream_off[s];
        const uint32_t bit = bp.cand_pos[ci];
        const FzTilePool pool{bp.tiles, bp.tiles_cap, &bp.ctl[FZ_BP_CTL_TILES]};
        if (WRITE) {
            const FzBlockInfo bi = bp.info[ci];
            uint8_t *out = planes + (uint64_t)(s & 3) * g.plane_stride + (uint64_t)(s >> 2) * g.chk + off;
            const uint32_t rec = bp.first_rec[ci];
            bool ok = false;
            if (rec != FZ_TILE_NONE)   // the measure pass kept the settled parse: decode every sub-range once
                fz_sy_block_from_table(st, in, len, bit, out, bi.out_len, bp.blk_prev[ci], bi.end_bit, pool, rec, &ok, lane);
            else
                fz_sy_block<true>(st, in, len, bit, out, bi.out_len, bp.blk_prev[ci], bi.end_bit, nullptr, &ok, lane);
            if (!ok && lane == 0) atomicExch(&bp.par_ok[s], 0u);
        } else {
            FzBlockInfo bi;
            uint32_t rec = FZ_TILE_NONE;
            // bits to the next candidate header of the stream (the candidates are sorted): how long this block probably is
            const uint32_t idx = (uint32_t)(ci - (size_t)s * FZ_BP_CAP);
            const uint32_t hint = (idx + 1u < bp.cand_cnt[s] ? bp.cand_pos[ci + 1] : len * 8u) - bit;
            fz_sy_block<false>(st, in, len, bit, nullptr, 0, -1, 0, &bi, nullptr, lane, &pool, &rec, hint);
            if (lane == 0) { bp.info[ci] = bi; bp.first_rec[ci] = rec; }
        }
        __syncwarp();
    }
}

// chain: one thread per general stream walks block end -> next block start
__global__ void __launch_bounds__(32)
fz_bp_chain_kernel(const uint8_t *__restrict__ container, FzBatchGeom g, const uint32_t *__restrict__ stream_hdr,
                   const unsigned long long *__restrict__ stream_off, FzBlockParBufs bp, const FzStatus *status)
{
    __shared__ uint16_t tabs[32 * FZ_INF_TAB_U16];   // only touched when a fixed-Huffman block is met
    const uint32_t ngen = bp.ctl[0];
    if (ngen == 0 || status->error) return;
    uint16_t *mine = tabs + threadIdx.x;
    const FzInfTab<32> tab{mine, mine + 288 * 32, mine + 320 * 32};
    for (uint32_t gi = blockIdx.x * blockDim.x + threadIdx.x; gi < ngen; gi += gridDim.x * blockDim.x) {
        const uint32_t s = bp.gen_list[gi];
        const uint32_t ncand = bp.cand_cnt[s];
        uint32_t ok = 0, nst = 0, nblocks = ncand;
        if (ncand <= FZ_BP_CAP) {
            const size_t c0 = (size_t)s * FZ_BP_CAP;
            const int rc = fz_chain_resolve(container + stream_off[s], stream_hdr[s] & ~FZ_RAW_FLAG, fz_bp_stream_n(g, s),
                                            bp.cand_pos + c0, bp.info + c0, ncand, (uint32_t)FZ_BP_CAP, &nblocks, bp.blk_off + c0,
                                            bp.blk_prev + c0, bp.stored + (size_t)s * FZ_BP_STORED_CAP, FZ_BP_STORED_CAP, &nst, tab);
            ok = rc == 0;
        }
        bp.cand_cnt[s] = ok ? nblocks : ncand;
        bp.nstored[s] = ok ? nst : 0u;
        bp.par_ok[s] = ok;
        if (ok) {   // work list of the write pass: the blocks on the chain
            const size_t c0 = (size_t)s * FZ_BP_CAP;
            uint32_t non = 0;
            for (uint32_t i = 0; i < nblocks; i++) non += bp.blk_off[c0 + i] != 0xFFFFFFFFu;
            uint32_t at = atomicAdd(&bp.ctl[FZ_BP_CTL_NWRITE], non);
            uint32_t *list = bp.items + (size_t)bp.nstreams * FZ_BP_CAP;
            for (uint32_t i = 0; i < nblocks; i++)
                if (bp.blk_off[c0 + i] != 0xFFFFFFFFu) list[at++] = (uint32_t)(c0 + i);
            for (uint32_t i = ncand; i < nblocks; i++) bp.first_rec[c0 + i] = FZ_TILE_NONE;   // fixed-Huffman blocks met on the way
        }
    }
}

// stored blocks met on the chain: plain copies, one warp each
__global__ void __launch_bounds__(128)
fz_bp_stored_kernel(const uint8_t *__restrict__ container, uint64_t container_size, FzBatchGeom g,
                    const unsigned long long *__restrict__ stream_off, FzBlockParBufs bp, uint8_t *__restrict__ planes,
                    const FzStatus *status)
{
    const uint32_t ngen = bp.ctl[0];
    if (ngen == 0 || status->error) return;
    const int lane = threadIdx.x & 31;
    const uint32_t nwarps = gridDim.x * 4;
    for (uint32_t wi = blockIdx.x * 4 + (threadIdx.x >> 5); wi < ngen * FZ_BP_STORED_CAP; wi += nwarps) {
        const uint32_t gi = wi / FZ_BP_STORED_CAP, k = wi - gi * FZ_BP_STORED_CAP;
        const uint32_t s = bp.gen_list[gi];
        if (!bp.par_ok[s] || k >= bp.nstored[s]) continue;
        const FzStoredItem it = bp.stored[(size_t)s * FZ_BP_STORED_CAP + k];
        uint8_t *dst = planes + (uint64_t)(s & 3) * g.plane_stride + (uint64_t)(s >> 2) * g.chk + it.out_off;
        fz_warp_copy(dst, container + stream_off[s] + it.src_byte, it.len, container + container_size, lane);
    }
}

// tile records: a stream of <= chk + 4 compressed bytes is (chk * 8) / (32 * FZ_BP_SUB_BITS) tiles plus one partial
// tile per block; false candidates may burn a few more.  When the pool runs dry the write pass searches again.
static uint32_t fz_bp_tiles_cap(uint32_t nstreams, uint32_t chk)
{
    // short blocks take one (smaller) tile each: room for one per candidate -- a block per 512 bytes of payload at the most
    // (a pool that runs out only makes the write pass search again)
    const uint64_t blocks = (uint64_t)chk / 512 + 1 < FZ_BP_CAP ? (uint64_t)chk / 512 + 1 : FZ_BP_CAP;
    const uint64_t per_stream = (uint64_t)chk * 8 / (32u * FZ_BP_SUB_BITS) + blocks + 64;
    const uint64_t cap = per_stream * nstreams;
    return (uint32_t)(cap < 0x7FFFFFFFu ? cap : 0x7FFFFFFFu);
}

size_t fz_blockpar_bytes(uint32_t nstreams, uint32_t chk)
{
    const size_t per_stream = 4 /*gen_list*/ + 4 /*cand_cnt*/ + 4 /*nstored*/ + 4 /*par_ok*/ +
                              (size_t)FZ_BP_CAP * (4 + sizeof(FzBlockInfo) + 4 + 4 + 8 + 4) + (size_t)FZ_BP_STORED_CAP * sizeof(FzStoredItem);
    return 256 + per_stream * nstreams + (size_t)fz_bp_tiles_cap(nstreams, chk) * sizeof(FzTileRec) + 16 * 64;
}

FzBlockParBufs fz_blockpar_carve(void *blob, uint32_t nstreams, uint32_t chk)
{
    FzBlockParBufs b;
    uint8_t *p = (uint8_t *)blob;
    auto take = [&](size_t bytes) { uint8_t *r = p; p += (bytes + 63) & ~(size_t)63; return r; };
    b.ctl = (uint32_t *)take(256);
    b.gen_list = (uint32_t *)take((size_t)nstreams * 4);
    b.cand_cnt = (uint32_t *)take((size_t)nstreams * 4);
    b.nstored = (uint32_t *)take((size_t)nstreams * 4);
    b.par_ok = (uint32_t *)take((size_t)nstreams * 4);
    b.cand_pos = (uint32_t *)take((size_t)nstreams * FZ_BP_CAP * 4);
    b.info = (FzBlockInfo *)take((size_t)nstreams * FZ_BP_CAP * sizeof(FzBlockInfo));
    b.blk_off = (uint32_t *)take((size_t)nstreams * FZ_BP_CAP * 4);
    b.blk_prev = (int *)take((size_t)nstreams * FZ_BP_CAP * 4);
    b.stored = (FzStoredItem *)take((size_t)nstreams * FZ_BP_STORED_CAP * sizeof(FzStoredItem));
    b.items = (uint32_t *)take((size_t)nstreams * FZ_BP_CAP * 4 * 2);
    b.first_rec = (uint32_t *)take((size_t)nstreams * FZ_BP_CAP * 4);
    b.tiles_cap = fz_bp_tiles_cap(nstreams, chk);
    b.tiles = (FzTileRec *)take((size_t)b.tiles_cap * sizeof(FzTileRec));
    b.nstreams = nstreams;
    return b;
}

// ---- RAW payloads: verbatim plane bytes (reference zip.c:264-267)
__global__ void __launch_bounds__(128)
fz_rawcopy_kernel(const uint8_t *__restrict__ container, uint64_t container_size, FzBatchGeom g,
                  const uint32_t *__restrict__ stream_hdr, const unsigned long long *__restrict__ stream_off,
                  uint8_t *__restrict__ planes, const FzStatus *status)
{
    const int lane = threadIdx.x & 31;
    const uint32_t t = blockIdx.x * 4 + (threadIdx.x >> 5);
    const uint32_t total = g.nchunks * FZ_PLANES * g.nsub_full;
    if (t >= total || status->error) return;
    const uint32_t s = t / g.nsub_full, k = t - s * g.nsub_full;
    if (!(stream_hdr[s] & FZ_RAW_FLAG)) return;
    const uint32_t n_s = fz_chunk_n(g, s / FZ_PLANES);
    const uint32_t off = k * FZ_SUB;
    if (off >= n_s) return;
    const uint32_t n = min((uint32_t)FZ_SUB, n_s - off);
    uint8_t *dst = planes + (uint64_t)(s & 3) * g.plane_stride + (uint64_t)(s >> 2) * g.chk + off;
    fz_warp_copy(dst, container + stream_off[s] + off, n, container + container_size, lane);
}

size_t fz_group_desc_bytes(uint32_t nstreams, uint32_t nsub_full)
{
    return (size_t)nstreams * ((nsub_full + FZ_CODE_SUBS - 1) / FZ_CODE_SUBS) * sizeof(FzGroupDesc) + FZ_SM_COUNT * 4u + 16u;
}

void fz_launch_inflate(const uint8_t *container, uint64_t container_size, FzBatchGeom g, const uint32_t *stream_hdr,
                       const unsigned long long *stream_off, FzInflateBufs b, uint8_t *planes, FzStatus *status, cudaStream_t st,
                       fz_mark_fn mark, void *mark_user, bool copy_raw)
{
    const uint32_t nstreams = g.nchunks * FZ_PLANES;
    const uint32_t ntiles = nstreams * b.tiles_per_stream;
    cudaMemsetAsync(b.tile_cnt, 0, ((size_t)ntiles + 1) * 4, st);   // look-back state of the marker scan + its ticket counter
    const uint32_t nscan = ntiles < FZ_SM_COUNT * 6u ? ntiles : FZ_SM_COUNT * 6u;   // 38 registers: 6 blocks of 256 threads per SM
    fz_marker_kernel<<<nscan, FZ_SCAN_THREADS, 0, st>>>(container, stream_hdr, stream_off, b.tiles_per_stream, ntiles, b.tile_cnt, b.stream_cnt, b.hits,
                                                        b.hits_per_stream, status);
    if (mark) mark(mark_user, FZ_ST_MARKERS);
    cudaMemsetAsync(b.bp.ctl, 0, 64, st);
    // zero-sub-block flags only when the merge that follows reads them (copy_raw: the plain merge reads the plane buffer)
    uint32_t *zf = copy_raw ? nullptr : b.zero_flags;
    if (zf) cudaMemsetAsync(zf, 0, (size_t)nstreams * g.nsub_full * 4, st);
    fz_classify_kernel<<<(nstreams + 127) / 128, 128, 0, st>>>(stream_hdr, g, b.stream_cnt, b.hits, b.hits_per_stream, b.stream_mode, b.stream_fail, b.bp, status);
    if (mark) mark(mark_user, FZ_ST_CLASSIFY);
    const uint32_t ncode = nstreams * ((g.nsub_full + FZ_CODE_SUBS - 1) / FZ_CODE_SUBS);   // one CTA per code group
    // Every lane streams its own fragment: what little L1 the shared-memory carve-out leaves decides how often an input
    // word is an L2 round trip.  164 KB of shared memory (4 CTAs of 37 KB) and 92 KB of L1 beat 228 KB / 6 CTAs by
    // 20-25 % on every input measured (sweep in profiles/README.md).
    FzGroupDesc *desc = (FzGroupDesc *)b.group_desc;
    uint32_t *qctr = (uint32_t *)(desc + ncode);             // FZ_SM_COUNT work-queue counters behind the descriptors
    fz_inflate_prep_kernel<<<ncode, FZ_WARP, 0, st>>>(container, g, stream_off, b.stream_cnt, b.hits_per_stream, b.hits, b.stream_mode, desc, qctr, planes, b.full_only ? 1u : 0u, status);
    cudaFuncSetAttribute(fz_inflate_lean_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(fz_inflate_lean_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FzLeanSmem));
    const uint32_t nlean = ncode < FZ_SM_COUNT * FZ_LEAN_MINBLOCKS ? ncode : FZ_SM_COUNT * FZ_LEAN_MINBLOCKS;
    fz_inflate_lean_kernel<<<nlean, FZ_INF_WARPS * FZ_WARP, sizeof(FzLeanSmem), st>>>(container, g, stream_off, b.stream_cnt, b.hits_per_stream, b.hits,
                                                                                     desc, qctr, ncode, zf, planes);
    // groups the lean kernel does not take or gave up on (desc.state == 2): the full inflater, one warp per 32 sub-blocks
    cudaFuncSetAttribute(fz_inflate_group_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, FZ_INF_CARVEOUT_PCT);
    cudaFuncSetAttribute(fz_inflate_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FzGroupSmem));
    fz_inflate_group_kernel<<<ncode, FZ_INF_WARPS * FZ_WARP, sizeof(FzGroupSmem), st>>>(
        container, g, stream_hdr, stream_off, b.stream_cnt, b.hits_per_stream, b.hits, b.stream_mode, b.stream_fail, zf, planes, desc, status);
    if (mark) mark(mark_user, FZ_ST_INFLATE_FAST);
    // zlib-made streams (the reference's own containers): block-parallel; whatever that refuses goes to the serial inflater
    const uint32_t segs = (g.chk + 16 + 4 * FZ_BP_SEG_WORDS - 1) / (4 * FZ_BP_SEG_WORDS) + 1;
    fz_bp_find_kernel<<<FZ_SM_COUNT * 16, FZ_BP_FIND_THREADS, 0, st>>>(container, container_size, stream_hdr, stream_off, b.bp, segs, status);
    fz_bp_sort_kernel<<<FZ_SM_COUNT, FZ_BP_CAP / 2, 0, st>>>(b.bp, status);
    fz_bp_sync_kernel<false><<<FZ_SM_COUNT * 5, FZ_BP_SY_WARPS * 32, 0, st>>>(container, g, stream_hdr, stream_off, b.bp, planes, status);
    fz_bp_chain_kernel<<<FZ_SM_COUNT, 32, 0, st>>>(container, g, stream_hdr, stream_off, b.bp, status);
    fz_bp_sync_kernel<true><<<FZ_SM_COUNT * 5, FZ_BP_SY_WARPS * 32, 0, st>>>(container, g, stream_hdr, stream_off, b.bp, planes, status);
    fz_bp_stored_kernel<<<FZ_SM_COUNT * 4, 128, 0, st>>>(container, container_size, g, stream_off, b.bp, planes, status);
    if (mark) mark(mark_user, FZ_ST_INFLATE_BLOCKPAR);
    fz_inflate_general_kernel<<<nstreams, 32, 0, st>>>(container, g, stream_hdr, stream_off, b.stream_mode, b.stream_fail, b.bp.par_ok, zf, planes, status);
    if (mark) mark(mark_user, FZ_ST_INFLATE_GENERAL);
    if (copy_raw) {  // only when the merge cannot read RAW payloads in place (chunk size not a multiple of 16)
        const uint32_t total = nstreams * g.nsub_full;
        fz_rawcopy_kernel<<<(total + 3) / 4, 128, 0, st>>>(container, container_size, g, stream_hdr, stream_off, planes, status);
        if (mark) mark(mark_user, FZ_ST_RAWCOPY);
    }
}

// =================================================================================================
// error report: what the reference's erroranalysis tool prints after a lossy round trip (src/tool/erroranalysis.c:188-220,
// calculateDiff): err = |n2 - n1|, relative error err / |n1| where |n1| > 10E-4 (0 elsewhere); here the maxima, where they
// are, and the sum, in one HBM-bound pass.  `other` == nullptr compares the words with their own masked form (n2 = n1 &
// mask behind the exempt header words): the error of apply_mask (workers.c:82-101) without a round trip.
// =================================================================================================
#define FZ_ERR_THREADS 256
#define FZ_ERR_BLOCKS (FZ_SM_COUNT * 8)

__device__ __forceinline__ void fz_err_better(float &best, unsigned long long &bi, float v, unsigned long long i)
{
    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
}

__global__ void __launch_bounds__(FZ_ERR_THREADS)
fz_error_kernel(const uint32_t *__restrict__ orig, const uint32_t *__restrict__ other, uint64_t nwords, uint32_t mask,
                uint64_t exempt, FzErrPartial *__restrict__ partial)
{
    __shared__ FzErrPartial sh[FZ_ERR_THREADS / 32];
    float ma = -1.f, mr = -1.f, sum = 0.f;
    unsigned long long ia = ~0ull, ir = ~0ull, nan = 0;
    const uint64_t stride = (uint64_t)gridDim.x * FZ_ERR_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * FZ_ERR_THREADS + threadIdx.x; i < nwords; i += stride) {
        const uint32_t a = fz_ld_stream32(orig + i);
        const uint32_t b = other ? fz_ld_stream32(other + i) : (i >= exempt ? (a & mask) : a);
        const float n1 = __uint_as_float(a), n2 = __uint_as_float(b);
        const float err = fabsf(n2 - n1);
        if (err != err) { nan++; continue; }             // NaN or Inf - Inf: counted, not ranked
        const float rel = fabsf(n1) > 10E-4f ? err / fabsf(n1) : 0.f;
        sum += err;
        fz_err_better(ma, ia, err, i);
        if (rel == rel) fz_err_better(mr, ir, rel, i);
    }
    double dsum = (double)sum;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const float oa = __shfl_xor_sync(0xffffffffu, ma, d), orl = __shfl_xor_sync(0xffffffffu, mr, d);
        const unsigned long long oia = __shfl_xor_sync(0xffffffffu, ia, d), oir = __shfl_xor_sync(0xffffffffu, ir, d);
        fz_err_better(ma, ia, oa, oia);
        fz_err_better(mr, ir, orl, oir);
        dsum += __shfl_xor_sync(0xffffffffu, dsum, d);
        nan += __shfl_xor_sync(0xffffffffu, nan, d);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sh[warp].max_abs = ma; sh[warp].i_abs = ia; sh[warp].max_rel = mr; sh[warp].i_rel = ir; sh[warp].sum = dsum; sh[warp].nan = nan; }
    __syncthreads();
    if (threadIdx.x == 0) {
        FzErrPartial r = sh[0];
        for (int w = 1; w < FZ_ERR_THREADS / 32; w++) {
            fz_err_better(r.max_abs, r.i_abs, sh[w].max_abs, sh[w].i_abs);
            fz_err_better(r.max_rel, r.i_rel, sh[w].max_rel, sh[w].i_rel);
            r.sum += sh[w].sum;
            r.nan += sh[w].nan;
        }
        partial[blockIdx.x] = r;
    }
}

// one block folds the per-block results into partial[0]
__global__ void __launch_bounds__(FZ_ERR_THREADS) fz_error_fold_kernel(FzErrPartial *__restrict__ partial, uint32_t n)
{
    __shared__ FzErrPartial sh[FZ_ERR_THREADS];
    FzErrPartial r;
    r.max_abs = -1.f; r.max_rel = -1.f; r.i_abs = ~0ull; r.i_rel = ~0ull; r.sum = 0.0; r.nan = 0;
    for (uint32_t i = threadIdx.x; i < n; i += FZ_ERR_THREADS) {
        const FzErrPartial p = partial[i];
        fz_err_better(r.max_abs, r.i_abs, p.max_abs, p.i_abs);
        fz_err_better(r.max_rel, r.i_rel, p.max_rel, p.i_rel);
        r.sum += p.sum;
        r.nan += p.nan;
    }
    sh[threadIdx.x] = r;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int t = 1; t < FZ_ERR_THREADS; t++) {
            fz_err_better(r.max_abs, r.i_abs, sh[t].max_abs, sh[t].i_abs);
            fz_err_better(r.max_rel, r.i_rel, sh[t].max_rel, sh[t].i_rel);
            r.sum += sh[t].sum;
            r.nan += sh[t].nan;
        }
        partial[0] = r;
    }
}

uint32_t fz_error_partials() { return FZ_ERR_BLOCKS; }

void fz_launch_error(const uint32_t *orig, const uint32_t *other, uint64_t nwords, uint32_t mask, uint64_t exempt,
                     FzErrPartial *partial, cudaStream_t st)
{
    fz_error_kernel<<<FZ_ERR_BLOCKS, FZ_ERR_THREADS, 0, st>>>(orig, other, nwords, mask, exempt, partial);
    fz_error_fold_kernel<<<1, FZ_ERR_THREADS, 0, st>>>(partial, FZ_ERR_BLOCKS);
}
