// fz_api.cu -- the mzb_* C ABI (include/mrczip_b200.h, group 2): contexts, batching, and the
// compress / decompress pipelines that string the kernels of fz_kernels.cu together.
//
// compress:   split -> encode -> layout (stream sums, RAW rule, scan, chunk headers) -> gather     per batch
// decompress: walk -> marker scan -> classify -> inflate (fast / general) -> RAW copy -> merge      per batch
//
// The container offset runs on the device (FzStatus.out_end) so that batches need no host sync.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/mrczip_b200.h"
#include "fz_kernels.h"

#define FZ_MAX_HOST_BATCHES 8192

#define FZ_CHECK(call)                                                                       \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            fprintf(stderr, "[mrczip_b200] %s:%d CUDA error: %s\n", __FILE__, __LINE__,      \
                    cudaGetErrorString(e_));                                                 \
            return MZB_E_CUDA;                                                               \
        }                                                                                    \
    } while (0)

// inside the host-buffer pipelines: copies and kernels may still be running against the caller's buffers when a call
// fails -- nothing returns before the three streams are idle
#define FZ_CHECK_PIPE(call)                                                                  \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            fprintf(stderr, "[mrczip_b200] %s:%d CUDA error: %s\n", __FILE__, __LINE__,      \
                    cudaGetErrorString(e_));                                                 \
            return pipe_fail(c, MZB_E_CUDA);                                                 \
        }                                                                                    \
    } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct mzb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    uint32_t batch_chunks = 192;  // 4.5 GiB of input per kernel batch: one batch for a 1024^3 volume
    int split_variant = 0, merge_variant = 0, inflate_variant = 0;
    DevBuf planes, scratch, sizes, sub_off, stream_hdr, stream_off, stream_mode, stream_fail;
    DevBuf tile_cnt, block_sums, hits, io_in, io_out, ghist, gcodes, blockpar, zero_flags, zero_hist, err_partial, chunk_tab, group_desc;
    bool zero_hist_ready = false;
    // host-buffer pipeline: copy streams, events, pinned per-batch end offsets
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    std::vector<cudaEvent_t> ev_h2d, ev_comp, ev_d2h;
    unsigned long long *h_ends = nullptr;  // pinned, FZ_MAX_HOST_BATCHES entries
    FzStatus *d_status = nullptr;
    FzStatus *h_status = nullptr;  // pinned
    mzb_stats stats;
    // per-stage CUDA-event timing (off by default)
    bool prof = false;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<int> ev_stage;  // stage id of event i (-1 = start marker)
    size_t ev_used = 0;
    float stage_ms[FZ_ST_COUNT] = {0};
};

static int pipe_fail(mzb_ctx *c, int rc)
{
    if (c->s_h2d) cudaStreamSynchronize(c->s_h2d);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->s_d2h) cudaStreamSynchronize(c->s_d2h);
    cudaGetLastError();
    return rc;
}

static void prof_mark(void *user, int stage)
{
    mzb_ctx *c = (mzb_ctx *)user;
    if (!c->prof) return;
    if (c->ev_used == c->ev_pool.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        c->ev_pool.push_back(e);
        c->ev_stage.push_back(0);
    }
    c->ev_stage[c->ev_used] = stage;
    cudaEventRecord(c->ev_pool[c->ev_used], c->stream);
    c->ev_used++;
}

static void prof_begin(mzb_ctx *c)
{
    c->ev_used = 0;
    for (int i = 0; i < FZ_ST_COUNT; i++) c->stage_ms[i] = 0.f;
    prof_mark(c, -1);
}

static void prof_collect(mzb_ctx *c)  // after the stream was synchronised
{
    if (!c->prof) return;
    for (size_t i = 1; i < c->ev_used; i++) {
        float ms = 0.f;
        if (c->ev_stage[i] >= 0 && cudaEventElapsedTime(&ms, c->ev_pool[i - 1], c->ev_pool[i]) == cudaSuccess)
            c->stage_ms[c->ev_stage[i]] += ms;
    }
}

static const char *kStageNames[FZ_ST_COUNT] = {"split", "encode", "layout", "gather", "walk", "markers", "classify",
                                                "inflate_fast", "inflate_blockpar", "inflate_general", "rawcopy", "merge"};

extern "C" int mzb_set_profiling(mzb_ctx *c, int on)
{
    if (!c) return MZB_E_ARG;
    c->prof = on != 0;
    return MZB_OK;
}
extern "C" int mzb_stage_count(void) { return FZ_ST_COUNT; }
extern "C" const char *mzb_stage_name(int i) { return (i >= 0 && i < FZ_ST_COUNT) ? kStageNames[i] : ""; }
extern "C" int mzb_stage_ms(mzb_ctx *c, float *out, int n)
{
    if (!c || !out) return MZB_E_ARG;
    for (int i = 0; i < n && i < FZ_ST_COUNT; i++) out[i] = c->stage_ms[i];
    return MZB_OK;
}

static int ensure(DevBuf &b, size_t need)
{
    if (b.cap >= need) return MZB_OK;
    if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
    need = (need + 255) & ~(size_t)255;
    cudaError_t e = cudaMalloc(&b.p, need);
    if (e != cudaSuccess) {
        fprintf(stderr, "[mrczip_b200] cudaMalloc(%zu) failed: %s\n", need, cudaGetErrorString(e));
        cudaGetLastError();
        return MZB_E_NOMEM;
    }
    b.cap = need;
    return MZB_OK;
}

static void release(DevBuf &b)
{
    if (b.p) cudaFree(b.p);
    b.p = nullptr; b.cap = 0;
}

extern "C" int mzb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" const char *mzb_version(void) { return "mrczip_b200 0.1 (sm_100a; sub-block " "16 KiB" ")"; }

extern "C" const char *mzb_strerror(int code)
{
    switch (code) {
        case MZB_OK: return "ok";
        case MZB_E_ARG: return "invalid argument";
        case MZB_E_CUDA: return "CUDA error (no device, or a runtime call failed)";
        case MZB_E_NOMEM: return "out of device memory";
        case MZB_E_FORMAT: return "malformed container or deflate stream";
        case MZB_E_SPACE: return "output buffer too small";
        case MZB_E_IO: return "file I/O error";
        default: return "unknown error";
    }
}

static int create_impl(mzb_ctx **out, int device, void *cuda_stream, bool use_given);

extern "C" int mzb_create(mzb_ctx **out, int device, void *cuda_stream)
{
    return create_impl(out, device, cuda_stream, cuda_stream != nullptr);
}

extern "C" int mzb_create_on_stream(mzb_ctx **out, int device, void *cuda_stream)
{
    return create_impl(out, device, cuda_stream, true);
}

static int create_impl(mzb_ctx **out, int device, void *cuda_stream, bool use_given)
{
    if (!out) return MZB_E_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        fprintf(stderr, "[mrczip_b200] no CUDA device: this library has no CPU fallback\n");
        cudaGetLastError();
        return MZB_E_CUDA;
    }
    if (device < 0 || device >= ndev) return MZB_E_ARG;
    FZ_CHECK(cudaSetDevice(device));
    mzb_ctx *c = new mzb_ctx();
    c->device = device;
    if (use_given) c->stream = (cudaStream_t)cuda_stream;
    else {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return MZB_E_CUDA; }
        c->own_stream = true;
    }
    if (cudaMalloc((void **)&c->d_status, sizeof(FzStatus)) != cudaSuccess ||
        cudaHostAlloc((void **)&c->h_status, sizeof(FzStatus), cudaHostAllocDefault) != cudaSuccess) {
        mzb_destroy(c);
        return MZB_E_NOMEM;
    }
    memset(&c->stats, 0, sizeof(c->stats));
    *out = c;
    return MZB_OK;
}

extern "C" void mzb_destroy(mzb_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    DevBuf *all[] = {&c->planes, &c->scratch, &c->sizes, &c->sub_off, &c->stream_hdr, &c->stream_off, &c->stream_mode,
                     &c->stream_fail, &c->tile_cnt, &c->block_sums, &c->hits, &c->io_in, &c->io_out, &c->ghist, &c->gcodes, &c->blockpar, &c->zero_flags, &c->zero_hist, &c->err_partial, &c->chunk_tab, &c->group_desc};
    for (DevBuf *b : all) release(*b);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    for (auto *v : {&c->ev_h2d, &c->ev_comp, &c->ev_d2h})
        for (cudaEvent_t e : *v) cudaEventDestroy(e);
    if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
    if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
    if (c->h_ends) cudaFreeHost(c->h_ends);
    if (c->d_status) cudaFree(c->d_status);
    if (c->h_status) cudaFreeHost(c->h_status);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int mzb_set_batch_chunks(mzb_ctx *c, uint32_t chunks)
{
    if (!c || chunks == 0 || chunks > 4096) return MZB_E_ARG;
    c->batch_chunks = chunks;
    return MZB_OK;
}

extern "C" int mzb_set_variant(mzb_ctx *c, int split_variant, int merge_variant)
{
    if (!c) return MZB_E_ARG;
    c->split_variant = split_variant;
    c->merge_variant = merge_variant;
    return MZB_OK;
}

extern "C" int mzb_set_inflate_variant(mzb_ctx *c, int variant)
{
    if (!c || variant < 0 || variant > 1) return MZB_E_ARG;
    c->inflate_variant = variant;
    return MZB_OK;
}

extern "C" size_t mzb_compress_bound(uint64_t nwords, uint32_t chk)
{
    if (chk == 0) return 0;
    const uint64_t chunks = (nwords + chk - 1) / chk;
    // every stream is at most n bytes (RAW rule), plus 16 bytes of header per chunk
    return (size_t)(MZB_FILE_HEADER_BYTES + chunks * FZ_CHUNK_HEADER_BYTES + nwords * 4 + 64);
}

extern "C" int mzb_last_stats(mzb_ctx *c, mzb_stats *out)
{
    if (!c || !out) return MZB_E_ARG;
    *out = c->stats;
    return MZB_OK;
}

// ---------------------------------------------------------------------------------------------------
static uint32_t batch_chunks_for(const mzb_ctx *c, uint32_t chk)
{
    // the kernels address stream (c, j) at plane_j + c * chk and need 16-byte aligned sub-blocks:
    // a chunk size that is not a multiple of 16 (never written by the reference) goes one chunk at a time
    return (chk % 16u) ? 1u : c->batch_chunks;
}

static int status_reset(mzb_ctx *c, uint64_t out_end)
{
    memset(c->h_status, 0, sizeof(FzStatus));
    c->h_status->out_end = out_end;
    FZ_CHECK(cudaMemcpyAsync(c->d_status, c->h_status, sizeof(FzStatus), cudaMemcpyHostToDevice, c->stream));
    return MZB_OK;
}

static int status_fetch(mzb_ctx *c)
{
    FZ_CHECK(cudaMemcpyAsync(c->h_status, c->d_status, sizeof(FzStatus), cudaMemcpyDeviceToHost, c->stream));
    FZ_CHECK(cudaStreamSynchronize(c->stream));
    FZ_CHECK(cudaGetLastError());
    return MZB_OK;
}

static FzBatchGeom make_geom(uint32_t nchunks, uint32_t chk, uint64_t nwords_batch, uint64_t plane_stride)
{
    FzBatchGeom g;
    g.nchunks = nchunks;
    g.chk = chk;
    g.last_n = (uint32_t)(nwords_batch - (uint64_t)(nchunks - 1) * chk);
    g.nsub_full = (chk + FZ_SUB - 1) / FZ_SUB;
    g.plane_stride = plane_stride;
    g.chunk_n = nullptr;
    g.chunk_exempt = nullptr;
    return g;
}

static uint64_t plane_stride_for(uint32_t bchunks, uint32_t chk)
{
    return (((uint64_t)bchunks * chk + 64) + 255) & ~(uint64_t)255;
}

extern "C" int mzb_mask_split_device(mzb_ctx *c, const void *d_words, uint64_t nwords, int bits, uint32_t exempt_words,
                                     void *d_planes, uint64_t plane_stride)
{
    if (!c || bits < 0 || bits > 32 || ((uintptr_t)d_words & 15) || ((uintptr_t)d_planes & 15) || (plane_stride & 15) ||
        plane_stride < nwords)
        return MZB_E_ARG;
    FZ_CHECK(cudaSetDevice(c->device));
    fz_launch_split((const uint32_t *)d_words, nwords, fz_mask_for_bits(bits), exempt_words, (uint8_t *)d_planes,
                    plane_stride, c->split_variant, c->stream);
    FZ_CHECK(cudaGetLastError());
    if (c->own_stream) FZ_CHECK(cudaStreamSynchronize(c->stream));  // nobody else can order against a private stream
    return MZB_OK;
}

extern "C" int mzb_merge_device(mzb_ctx *c, const void *d_planes, uint64_t plane_stride, uint64_t nwords, void *d_words_out)
{
    if (!c || ((uintptr_t)d_words_out & 15) || ((uintptr_t)d_planes & 15) || (plane_stride & 15) || plane_stride < nwords)
        return MZB_E_ARG;
    FZ_CHECK(cudaSetDevice(c->device));
    fz_launch_merge((const uint8_t *)d_planes, plane_stride, nwords, (uint32_t *)d_words_out, c->merge_variant, c->stream);
    FZ_CHECK(cudaGetLastError());
    if (c->own_stream) FZ_CHECK(cudaStreamSynchronize(c->stream));
    return MZB_OK;
}

// ---------------------------------------------------------------------------------------------------
// shared by the device-resident and the host-buffer paths
static int compress_reserve(mzb_ctx *c, uint32_t bmax, uint32_t chk, uint64_t *pstride_out)
{
    const uint64_t pstride = plane_stride_for(bmax, chk);
    const uint32_t nsub_full = (chk + FZ_SUB - 1) / FZ_SUB;
    const size_t nslots = (size_t)bmax * FZ_PLANES * nsub_full;
    const size_t ngroups = (size_t)bmax * FZ_PLANES * ((nsub_full + FZ_CODE_SUBS - 1) / FZ_CODE_SUBS);
    int rc;
    if ((rc = ensure(c->planes, pstride * FZ_PLANES + 256)) || (rc = ensure(c->scratch, nslots * FZ_SLOT_STRIDE + 256)) ||
        (rc = ensure(c->sizes, nslots * 4)) || (rc = ensure(c->sub_off, nslots * 4)) ||
        (rc = ensure(c->stream_hdr, (size_t)bmax * FZ_PLANES * 4)) || (rc = ensure(c->stream_off, (size_t)bmax * FZ_PLANES * 8)) ||
        (rc = ensure(c->ghist, ngroups * 288 * 4)) || (rc = ensure(c->gcodes, ngroups * fz_group_code_bytes())) ||
        (rc = ensure(c->zero_hist, 288 * 4)))
        return rc;
    *pstride_out = pstride;
    return MZB_OK;
}

static unsigned long long g_passes[2] = {0, 0};
extern "C" void mzb_pass_counts(uint64_t *zp, uint64_t *up)
{
    if (zp) *zp = __atomic_load_n(&g_passes[0], __ATOMIC_RELAXED);
    if (up) *up = __atomic_load_n(&g_passes[1], __ATOMIC_RELAXED);
}

static void compress_enqueue_batch(mzb_ctx *c, const uint32_t *d_words, uint64_t nw, uint32_t nb, uint32_t chk, uint64_t pstride,
                                   uint32_t mask, uint64_t exempt, uint8_t *d_out, size_t out_cap)
{
    const FzBatchGeom g = make_geom(nb, chk, nw, pstride);
    __atomic_fetch_add(&g_passes[0], 1ull, __ATOMIC_RELAXED);
    uint32_t zero_planes = 0;
    for (int j = 0; j < FZ_PLANES; j++)
        if (((mask >> (8 * j)) & 0xffu) == 0) zero_planes |= 1u << j;
    // Planes the mask erases are not even written where the histogram kernel flags whole sub-blocks all-zero without
    // reading them (full sub-blocks behind the exempt words; sub-block starts are multiples of FZ_SUB in the batch when
    // the chunk size is): 1 GiB of stores less per erased plane and 4 GiB volume.
    uint64_t skip_lo = 0, skip_hi = 0;
    if (zero_planes && chk % FZ_SUB == 0 && c->split_variant == 0) {
        skip_lo = (exempt + FZ_SUB - 1) / FZ_SUB * FZ_SUB;
        skip_hi = nw / FZ_SUB * FZ_SUB;
    }
    fz_launch_split(d_words, nw, mask, exempt, (uint8_t *)c->planes.p, pstride, c->split_variant, c->stream,
                    skip_hi > skip_lo ? zero_planes : 0u, skip_lo, skip_hi);
    prof_mark(c, FZ_ST_SPLIT);
    if (!c->zero_hist_ready) { fz_launch_zero_hist((uint32_t *)c->zero_hist.p, c->stream); c->zero_hist_ready = true; }
    fz_launch_encode((const uint8_t *)c->planes.p, g, (uint32_t *)c->ghist.p, c->gcodes.p, (uint8_t *)c->scratch.p,
                     (uint32_t *)c->sizes.p, (const uint32_t *)c->zero_hist.p, zero_planes, exempt, c->d_status, c->stream);
    prof_mark(c, FZ_ST_ENCODE);
    fz_launch_layout((uint32_t *)c->sizes.p, g, (uint32_t *)c->sub_off.p, (uint32_t *)c->stream_hdr.p,
                     (unsigned long long *)c->stream_off.p, d_out, out_cap, c->d_status, c->stream);
    prof_mark(c, FZ_ST_LAYOUT);
    fz_launch_gather((const uint8_t *)c->planes.p, (const uint8_t *)c->scratch.p, (const uint32_t *)c->sizes.p,
                     (const uint32_t *)c->sub_off.p, (const uint32_t *)c->stream_hdr.p,
                     (const unsigned long long *)c->stream_off.p, g, d_out, c->d_status, c->stream);
    prof_mark(c, FZ_ST_GATHER);
}

static int decompress_reserve(mzb_ctx *c, uint32_t bmax, uint32_t chk, uint64_t *pstride_out, FzInflateBufs *ib)
{
    const uint64_t pstride = plane_stride_for(bmax, chk);
    const uint32_t nsub_full = (chk + FZ_SUB - 1) / FZ_SUB;
    const uint32_t nstreams = bmax * FZ_PLANES;
    ib->tiles_per_stream = chk / 65536 + 2;  // FZ_TILE_BYTES
    const size_t ntiles = (size_t)nstreams * ib->tiles_per_stream;
    ib->hits_per_stream = nsub_full * 2 + 16;   // twice what our own framing produces: more markers than that is a foreign stream
    int rc;
    if ((rc = ensure(c->planes, pstride * FZ_PLANES + 256)) || (rc = ensure(c->stream_hdr, (size_t)nstreams * 4)) ||
        (rc = ensure(c->stream_off, (size_t)nstreams * 8)) || (rc = ensure(c->stream_mode, (size_t)nstreams * 4)) ||
        (rc = ensure(c->stream_fail, (size_t)nstreams * 4)) || (rc = ensure(c->tile_cnt, (ntiles + 1) * 4)) ||
        (rc = ensure(c->block_sums, (size_t)nstreams * 4)) || (rc = ensure(c->hits, (size_t)nstreams * ib->hits_per_stream * 4)) ||
        (rc = ensure(c->blockpar, fz_blockpar_bytes(nstreams, chk))) ||
        (rc = ensure(c->zero_flags, (size_t)nstreams * nsub_full * 4)) ||
        (rc = ensure(c->group_desc, fz_group_desc_bytes(nstreams, nsub_full))))
        return rc;
    ib->group_desc = c->group_desc.p;
    ib->full_only = c->inflate_variant == 1;
    ib->zero_flags = (uint32_t *)c->zero_flags.p;
    ib->bp = fz_blockpar_carve(c->blockpar.p, nstreams, chk);
    ib->tile_cnt = (uint32_t *)c->tile_cnt.p;
    ib->stream_cnt = (uint32_t *)c->block_sums.p;
    ib->hits = (uint32_t *)c->hits.p;
    ib->stream_mode = (uint32_t *)c->stream_mode.p;
    ib->stream_fail = (uint32_t *)c->stream_fail.p;
    *pstride_out = pstride;
    return MZB_OK;
}

// records of the batch start at d_in + status->out_end
static void decompress_enqueue_batch(mzb_ctx *c, const uint8_t *d_in, size_t in_size, FzBatchGeom g, const FzInflateBufs &ib,
                                     uint32_t *d_words_out, bool in_place_raw)
{
    __atomic_fetch_add(&g_passes[1], 1ull, __ATOMIC_RELAXED);
    fz_launch_walk(d_in, in_size, g, (uint32_t *)c->stream_hdr.p, (unsigned long long *)c->stream_off.p, c->d_status, c->stream);
    prof_mark(c, FZ_ST_WALK);
    fz_launch_inflate(d_in, in_size, g, (const uint32_t *)c->stream_hdr.p, (const unsigned long long *)c->stream_off.p, ib,
                      (uint8_t *)c->planes.p, c->d_status, c->stream, prof_mark, c, !in_place_raw);
    if (in_place_raw)
        fz_launch_merge_streams((const uint8_t *)c->planes.p, d_in, in_size, (const uint32_t *)c->stream_hdr.p,
                                (const unsigned long long *)c->stream_off.p, ib.zero_flags, g, d_words_out, c->stream);
    else
        fz_launch_merge((const uint8_t *)c->planes.p, g.plane_stride, (uint64_t)(g.nchunks - 1) * g.chk + g.last_n, d_words_out,
                        c->merge_variant, c->stream);
    prof_mark(c, FZ_ST_MERGE);
}

// ---------------------------------------------------------------------------------------------------
static void fill_compress_stats(mzb_ctx *c, uint64_t nwords, uint64_t nchunks_total, uint32_t launches)
{
    c->stats.bytes_in = nwords * 4;
    c->stats.bytes_out = c->h_status->out_end;
    c->stats.chunks = (uint32_t)nchunks_total;
    c->stats.streams = (uint32_t)nchunks_total * FZ_PLANES;
    c->stats.raw_streams = c->h_status->n_raw_streams;
    c->stats.stored_subblocks = c->h_status->n_stored_sub;
    c->stats.zero_subblocks = c->h_status->n_zero_sub;
    c->stats.kernel_launches = launches;
}

static void fill_decompress_stats(mzb_ctx *c, uint64_t bytes_in, uint64_t nwords, uint64_t nchunks_total, uint32_t launches)
{
    c->stats.bytes_in = bytes_in;
    c->stats.bytes_out = nwords * 4;
    c->stats.chunks = (uint32_t)nchunks_total;
    c->stats.streams = (uint32_t)nchunks_total * FZ_PLANES;
    c->stats.general_streams = c->h_status->n_general;
    c->stats.fast_failed = c->h_status->n_fast_failed;
    c->stats.blockpar_streams = c->h_status->n_blockpar;
    c->stats.kernel_launches = launches;
}

static uint32_t compress_launches(uint64_t nw) { return 7 + ((nw & 3) ? 1 : 0); }
static uint32_t decompress_launches(uint64_t nw, bool in_place_raw)
{
    // walk, marker scan, classify, header pass + lean inflate + full group inflate, six block-parallel kernels, general
    // inflate, then either the in-place merge or RAW copy + merge (+ tail)
    return 1 + 2 + 3 + 6 + 1 + (in_place_raw ? 1 : 2) + ((!in_place_raw && (nw & 3)) ? 1 : 0);
}

extern "C" int mzb_compress_device(mzb_ctx *c, const void *d_words, uint64_t nwords, int bits, uint32_t exempt_words,
                                   uint32_t chk, uint64_t fsz, int write_file_header, void *d_out, size_t out_cap,
                                   uint64_t *out_size)
{
    if (!c || !out_size || bits < 0 || bits > 32 || chk == 0 || chk >= 0x80000000u || ((uintptr_t)d_words & 15))
        return MZB_E_ARG;
    *out_size = 0;
    memset(&c->stats, 0, sizeof(c->stats));
    if (nwords == 0) return MZB_OK;  // reference workers.c:757-764: an empty input writes nothing, not even the header
    FZ_CHECK(cudaSetDevice(c->device));
    const uint64_t nchunks_total = (nwords + chk - 1) / chk;
    const uint32_t bmax = (uint32_t)(nchunks_total < batch_chunks_for(c, chk) ? nchunks_total : batch_chunks_for(c, chk));
    uint64_t pstride;
    int rc;
    if ((rc = compress_reserve(c, bmax, chk, &pstride))) return rc;

    uint64_t start = 0;
    if (write_file_header) {
        if (out_cap < MZB_FILE_HEADER_BYTES) return MZB_E_SPACE;
        // common.c:137-149: u64 fsz, u32 chk, u8 type = 0, u8 ztypes[4] = 0 (ZLIB_DEF)
        uint8_t hdr[MZB_FILE_HEADER_BYTES];
        memset(hdr, 0, sizeof(hdr));
        memcpy(hdr, &fsz, 8);
        memcpy(hdr + 8, &chk, 4);
        FZ_CHECK(cudaMemcpyAsync(d_out, hdr, sizeof(hdr), cudaMemcpyHostToDevice, c->stream));
        FZ_CHECK(cudaStreamSynchronize(c->stream));  // hdr is a stack buffer
        start = MZB_FILE_HEADER_BYTES;
    }
    if ((rc = status_reset(c, start))) return rc;

    const uint32_t mask = fz_mask_for_bits(bits);
    uint32_t launches = 0;
    prof_begin(c);
    for (uint64_t c0 = 0; c0 < nchunks_total; c0 += bmax) {
        const uint32_t nb = (uint32_t)((nchunks_total - c0) < bmax ? (nchunks_total - c0) : bmax);
        const uint64_t w0 = c0 * chk;
        const uint64_t nw = (w0 + (uint64_t)nb * chk <= nwords) ? (uint64_t)nb * chk : nwords - w0;
        const uint64_t exempt = exempt_words > w0 ? exempt_words - w0 : 0;
        const uint32_t *src = (const uint32_t *)d_words + w0;
        if ((uintptr_t)src & 15u) {
            // chunk sizes that are not a multiple of 4 words put later chunks off the 16-byte grid the split kernel's
            // 128-bit loads need (never written by the reference; one chunk per batch here): go through an aligned copy
            if ((rc = ensure(c->io_in, nw * 4 + 16))) return rc;
            FZ_CHECK(cudaMemcpyAsync(c->io_in.p, src, nw * 4, cudaMemcpyDeviceToDevice, c->stream));
            src = (const uint32_t *)c->io_in.p;
        }
        compress_enqueue_batch(c, src, nw, nb, chk, pstride, mask, exempt, (uint8_t *)d_out, out_cap);
        launches += compress_launches(nw);
    }
    if ((rc = status_fetch(c))) return rc;
    prof_collect(c);
    fill_compress_stats(c, nwords, nchunks_total, launches);
    if (c->h_status->error) return c->h_status->error;
    *out_size = c->h_status->out_end;
    return MZB_OK;
}

// ---------------------------------------------------------------------------------------------------
extern "C" int mzb_decompress_device(mzb_ctx *c, const void *d_in, size_t in_size, int has_file_header, uint32_t chk,
                                     uint64_t nwords, void *d_words_out, uint64_t out_cap_words, uint64_t *nwords_out)
{
    if (!c || !nwords_out || ((uintptr_t)d_words_out & 15)) return MZB_E_ARG;
    *nwords_out = 0;
    memset(&c->stats, 0, sizeof(c->stats));
    FZ_CHECK(cudaSetDevice(c->device));
    uint64_t start = 0;
    if (has_file_header) {
        if (in_size == 0) return MZB_OK;  // the reference writes no header for an empty file
        if (in_size < MZB_FILE_HEADER_BYTES) return MZB_E_FORMAT;  // common.c:119-123
        uint8_t hdr[MZB_FILE_HEADER_BYTES];
        FZ_CHECK(cudaMemcpyAsync(hdr, d_in, sizeof(hdr), cudaMemcpyDeviceToHost, c->stream));
        FZ_CHECK(cudaStreamSynchronize(c->stream));
        uint64_t fsz;
        memcpy(&fsz, hdr, 8);
        memcpy(&chk, hdr + 8, 4);
        for (int j = 0; j < MZB_PLANES; j++)
            if (hdr[13 + j] != 0) return MZB_E_FORMAT;  // only ztype 0 (zlib) is ever written (workers.c:719)
        nwords = fsz / 4;  // workers.c:577
        start = MZB_FILE_HEADER_BYTES;
    }
    if (chk == 0 || chk >= 0x80000000u) return MZB_E_FORMAT;  // zip.c:325-329
    if (nwords > out_cap_words) return MZB_E_SPACE;
    if (nwords == 0) return MZB_OK;
    const uint64_t nchunks_total = (nwords + chk - 1) / chk;
    if (in_size < start + nchunks_total * FZ_CHUNK_HEADER_BYTES) return MZB_E_FORMAT;
    const uint32_t bmax = (uint32_t)(nchunks_total < batch_chunks_for(c, chk) ? nchunks_total : batch_chunks_for(c, chk));
    uint64_t pstride;
    FzInflateBufs ib;
    int rc;
    if ((rc = decompress_reserve(c, bmax, chk, &pstride, &ib))) return rc;
    if ((rc = status_reset(c, start))) return rc;
    const bool in_place_raw = (chk % 16u) == 0;  // merge reads RAW payloads straight from the container

    uint32_t launches = 0;
    prof_begin(c);
    for (uint64_t c0 = 0; c0 < nchunks_total; c0 += bmax) {
        const uint32_t nb = (uint32_t)((nchunks_total - c0) < bmax ? (nchunks_total - c0) : bmax);
        const uint64_t w0 = c0 * chk;
        const uint64_t nw = (w0 + (uint64_t)nb * chk <= nwords) ? (uint64_t)nb * chk : nwords - w0;
        uint32_t *dst = (uint32_t *)d_words_out + w0;
        const bool staged = ((uintptr_t)dst & 15u) != 0;   // see mzb_compress_device: odd chunk sizes, one chunk per batch
        if (staged && (rc = ensure(c->io_out, nw * 4 + 16))) return rc;
        decompress_enqueue_batch(c, (const uint8_t *)d_in, in_size, make_geom(nb, chk, nw, pstride), ib,
                                 staged ? (uint32_t *)c->io_out.p : dst, in_place_raw);
        if (staged) FZ_CHECK(cudaMemcpyAsync(dst, c->io_out.p, nw * 4, cudaMemcpyDeviceToDevice, c->stream));
        launches += decompress_launches(nw, in_place_raw);
    }
    if ((rc = status_fetch(c))) return rc;
    prof_collect(c);
    fill_decompress_stats(c, c->h_status->out_end, nwords, nchunks_total, launches);
    if (c->h_status->error) return c->h_status->error;
    *nwords_out = nwords;
    return MZB_OK;
}

// ---------------------------------------------------------------------------------------------------
// host-buffer versions: a three-stage pipeline over batches of chunks -- H2D of batch b+1, kernels of batch b
// and D2H of batch b-1 run concurrently on three streams (double-buffered device staging).
static int host_pipe_init(mzb_ctx *c, size_t nbatches)
{
    if (!c->s_h2d) FZ_CHECK(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
    if (!c->s_d2h) FZ_CHECK(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
    if (!c->h_ends) FZ_CHECK(cudaHostAlloc((void **)&c->h_ends, FZ_MAX_HOST_BATCHES * sizeof(unsigned long long), cudaHostAllocDefault));
    for (auto *v : {&c->ev_h2d, &c->ev_comp, &c->ev_d2h})
        while (v->size() < nbatches) {
            cudaEvent_t e;
            FZ_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            v->push_back(e);
        }
    return MZB_OK;
}

static uint32_t host_batch_chunks(const mzb_ctx *c, uint32_t chk)
{
    const uint32_t b = batch_chunks_for(c, chk);
    const char *e = getenv("MRCZIP_HOST_BATCH_CHUNKS");   // tuning knob; default 12 chunks = 288 MiB per batch
    uint32_t hb = e ? (uint32_t)atoi(e) : 12u;
    if (hb == 0) hb = 12u;
    return b > hb ? hb : b;  // finer batches than the device path: more overlap between copies and kernels
}

extern "C" int mzb_compress_host(mzb_ctx *c, const void *h_words, uint64_t nwords, int bits, uint32_t exempt_words,
                                 uint32_t chk, uint64_t fsz, int write_file_header, void *h_out, size_t out_cap,
                                 uint64_t *out_size)
{
    if (!c || !out_size || bits < 0 || bits > 32 || chk == 0 || chk >= 0x80000000u) return MZB_E_ARG;
    *out_size = 0;
    memset(&c->stats, 0, sizeof(c->stats));
    if (nwords == 0) return MZB_OK;
    FZ_CHECK(cudaSetDevice(c->device));
    const uint64_t nchunks_total = (nwords + chk - 1) / chk;
    const uint32_t hb = host_batch_chunks(c, chk);
    const uint32_t bmax = (uint32_t)(nchunks_total < hb ? nchunks_total : hb);
    const size_t nbatches = (size_t)((nchunks_total + bmax - 1) / bmax);
    if (nbatches > FZ_MAX_HOST_BATCHES) return MZB_E_ARG;
    const size_t bound = mzb_compress_bound(nwords, chk);
    const size_t in_stride = (((size_t)bmax * chk * 4 + 255) & ~(size_t)255) + 256;
    uint64_t pstride;
    int rc;
    if ((rc = compress_reserve(c, bmax, chk, &pstride)) || (rc = ensure(c->io_in, 2 * in_stride)) ||
        (rc = ensure(c->io_out, bound + 256)) || (rc = host_pipe_init(c, nbatches)))
        return rc;
    uint8_t *d_out = (uint8_t *)c->io_out.p;
    uint64_t start = 0;
    if (write_file_header) {
        if (out_cap < MZB_FILE_HEADER_BYTES) return MZB_E_SPACE;
        uint8_t *hd = (uint8_t *)h_out;  // common.c:137-149
        memset(hd, 0, MZB_FILE_HEADER_BYTES);
        memcpy(hd, &fsz, 8);
        memcpy(hd + 8, &chk, 4);
        start = MZB_FILE_HEADER_BYTES;
    }
    if ((rc = status_reset(c, start))) return rc;
    const uint32_t mask = fz_mask_for_bits(bits);
    uint32_t launches = 0;
    uint64_t copied_to = start;  // container bytes already on their way to the host
    int result = MZB_OK;
    prof_begin(c);
    for (size_t b = 0; b <= nbatches; b++) {
        if (b < nbatches) {
            const uint64_t c0 = (uint64_t)b * bmax;
            const uint32_t nb = (uint32_t)((nchunks_total - c0) < bmax ? (nchunks_total - c0) : bmax);
            const uint64_t w0 = c0 * chk;
            const uint64_t nw = (w0 + (uint64_t)nb * chk <= nwords) ? (uint64_t)nb * chk : nwords - w0;
            uint8_t *d_in = (uint8_t *)c->io_in.p + (b & 1) * in_stride;
            if (b >= 2) FZ_CHECK_PIPE(cudaStreamWaitEvent(c->s_h2d, c->ev_comp[b - 2], 0));  // staging buffer free again
            FZ_CHECK_PIPE(cudaMemcpyAsync(d_in, (const uint32_t *)h_words + w0, nw * 4, cudaMemcpyHostToDevice, c->s_h2d));
            FZ_CHECK_PIPE(cudaEventRecord(c->ev_h2d[b], c->s_h2d));
            FZ_CHECK_PIPE(cudaStreamWaitEvent(c->stream, c->ev_h2d[b], 0));
            const uint64_t exempt = exempt_words > w0 ? exempt_words - w0 : 0;
            compress_enqueue_batch(c, (const uint32_t *)d_in, nw, nb, chk, pstride, mask, exempt, d_out, bound);
            FZ_CHECK_PIPE(cudaMemcpyAsync(&c->h_ends[b], &c->d_status->out_end, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
            FZ_CHECK_PIPE(cudaEventRecord(c->ev_comp[b], c->stream));
            launches += compress_launches(nw);
        }
        if (b >= 1) {  // batch b-1 is done (or about to be): ship its part of the container while batch b computes
            FZ_CHECK_PIPE(cudaEventSynchronize(c->ev_comp[b - 1]));
            const uint64_t end = c->h_ends[b - 1];
            if (end > out_cap) { result = MZB_E_SPACE; break; }
            if (end > copied_to)
                FZ_CHECK_PIPE(cudaMemcpyAsync((uint8_t *)h_out + copied_to, d_out + copied_to, end - copied_to, cudaMemcpyDeviceToHost, c->s_d2h));
            copied_to = end;
        }
    }
    FZ_CHECK_PIPE(cudaStreamSynchronize(c->s_h2d));
    FZ_CHECK_PIPE(cudaStreamSynchronize(c->s_d2h));
    if ((rc = status_fetch(c))) return pipe_fail(c, rc);
    prof_collect(c);
    fill_compress_stats(c, nwords, nchunks_total, launches);
    if (result != MZB_OK) return result;
    if (c->h_status->error) return c->h_status->error;
    *out_size = c->h_status->out_end;
    return MZB_OK;
}

extern "C" int mzb_decompress_host(mzb_ctx *c, const void *h_in, size_t in_size, int has_file_header, uint32_t chk,
                                   uint64_t nwords, void *h_words_out, uint64_t out_cap_words, uint64_t *nwords_out)
{
    if (!c || !nwords_out) return MZB_E_ARG;
    *nwords_out = 0;
    memset(&c->stats, 0, sizeof(c->stats));
    if (in_size == 0) return MZB_OK;
    FZ_CHECK(cudaSetDevice(c->device));
    const uint8_t *in = (const uint8_t *)h_in;
    uint64_t off = 0;
    if (has_file_header) {
        if (in_size < MZB_FILE_HEADER_BYTES) return MZB_E_FORMAT;  // common.c:119-123
        uint64_t fsz;
        memcpy(&fsz, in, 8);
        memcpy(&chk, in + 8, 4);
        for (int j = 0; j < MZB_PLANES; j++)
            if (in[13 + j] != 0) return MZB_E_FORMAT;
        nwords = fsz / 4;
        off = MZB_FILE_HEADER_BYTES;
    }
    if (chk == 0 || chk >= 0x80000000u) return MZB_E_FORMAT;
    if (nwords > out_cap_words) return MZB_E_SPACE;
    if (nwords == 0) return MZB_OK;
    const uint64_t nchunks_total = (nwords + chk - 1) / chk;
    const uint32_t hb = host_batch_chunks(c, chk);
    const uint32_t bmax = (uint32_t)(nchunks_total < hb ? nchunks_total : hb);
    const size_t nbatches = (size_t)((nchunks_total + bmax - 1) / bmax);
    if (nbatches > FZ_MAX_HOST_BATCHES) return MZB_E_ARG;
    // host walk of the chunk-header chain (workers.c:61-69): byte range of every batch of chunk records
    std::vector<uint64_t> rec_off(nbatches + 1);
    for (uint64_t ci = 0; ci < nchunks_total; ci++) {
        if (ci % bmax == 0) rec_off[ci / bmax] = off;
        if (off + FZ_CHUNK_HEADER_BYTES > in_size) return MZB_E_FORMAT;
        uint64_t rec = FZ_CHUNK_HEADER_BYTES;
        for (int j = 0; j < MZB_PLANES; j++) {
            uint32_t h;
            memcpy(&h, in + off + 4 * j, 4);
            rec += h & ~FZ_RAW_FLAG;
        }
        if (off + rec > in_size) return MZB_E_FORMAT;
        off += rec;
    }
    rec_off[nbatches] = off;
    size_t max_rec = 0;
    for (size_t b = 0; b < nbatches; b++) max_rec = max_rec > rec_off[b + 1] - rec_off[b] ? max_rec : (size_t)(rec_off[b + 1] - rec_off[b]);
    const size_t in_stride = ((max_rec + 255) & ~(size_t)255) + 256;
    const size_t out_stride = (((size_t)bmax * chk * 4 + 255) & ~(size_t)255) + 256;
    uint64_t pstride;
    FzInflateBufs ib;
    int rc;
    if ((rc = decompress_reserve(c, bmax, chk, &pstride, &ib)) || (rc = ensure(c->io_in, 2 * in_stride)) ||
        (rc = ensure(c->io_out, 2 * out_stride)) || (rc = host_pipe_init(c, nbatches)))
        return rc;
    if ((rc = status_reset(c, 0))) return rc;
    const bool in_place_raw = (chk % 16u) == 0;
    uint32_t launches = 0;
    prof_begin(c);
    for (size_t b = 0; b < nbatches; b++) {
        const uint64_t c0 = (uint64_t)b * bmax;
        const uint32_t nb = (uint32_t)((nchunks_total - c0) < bmax ? (nchunks_total - c0) : bmax);
        const uint64_t w0 = c0 * chk;
        const uint64_t nw = (w0 + (uint64_t)nb * chk <= nwords) ? (uint64_t)nb * chk : nwords - w0;
        const size_t rbytes = (size_t)(rec_off[b + 1] - rec_off[b]);
        uint8_t *d_rec = (uint8_t *)c->io_in.p + (b & 1) * in_stride;
        uint32_t *d_w = (uint32_t *)((uint8_t *)c->io_out.p + (b & 1) * out_stride);
        if (b >= 2) FZ_CHECK_PIPE(cudaStreamWaitEvent(c->s_h2d, c->ev_comp[b - 2], 0));
        FZ_CHECK_PIPE(cudaMemcpyAsync(d_rec, in + rec_off[b], rbytes, cudaMemcpyHostToDevice, c->s_h2d));
        FZ_CHECK_PIPE(cudaEventRecord(c->ev_h2d[b], c->s_h2d));
        FZ_CHECK_PIPE(cudaStreamWaitEvent(c->stream, c->ev_h2d[b], 0));
        if (b >= 2) FZ_CHECK_PIPE(cudaStreamWaitEvent(c->stream, c->ev_d2h[b - 2], 0));  // output staging buffer free again
        FZ_CHECK_PIPE(cudaMemsetAsync(&c->d_status->out_end, 0, sizeof(unsigned long long), c->stream));  // records start at 0
        decompress_enqueue_batch(c, d_rec, rbytes, make_geom(nb, chk, nw, pstride), ib, d_w, in_place_raw);
        FZ_CHECK_PIPE(cudaEventRecord(c->ev_comp[b], c->stream));
        FZ_CHECK_PIPE(cudaStreamWaitEvent(c->s_d2h, c->ev_comp[b], 0));
        FZ_CHECK_PIPE(cudaMemcpyAsync((uint32_t *)h_words_out + w0, d_w, nw * 4, cudaMemcpyDeviceToHost, c->s_d2h));
        FZ_CHECK_PIPE(cudaEventRecord(c->ev_d2h[b], c->s_d2h));
        launches += decompress_launches(nw, in_place_raw);
    }
    FZ_CHECK_PIPE(cudaStreamSynchronize(c->s_d2h));
    if ((rc = status_fetch(c))) return pipe_fail(c, rc);
    prof_collect(c);
    fill_decompress_stats(c, off, nwords, nchunks_total, launches);
    if (c->h_status->error) return c->h_status->error;
    *nwords_out = nwords;
    return MZB_OK;
}

// pinned host memory for the C host layer (mrczip_host.c has no CUDA headers)
extern "C" void *mzb_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void mzb_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

// ---------------------------------------------------------------------------------------------------
// error report (reference src/tool/erroranalysis.c:188-220) and the MRC header fields (src/tool/mrcviewer.c:20-71)
static void err_merge(FzErrPartial &r, const FzErrPartial &p, uint64_t base)
{
    if (p.i_abs != ~0ull && (p.max_abs > r.max_abs || (p.max_abs == r.max_abs && p.i_abs + base < r.i_abs))) { r.max_abs = p.max_abs; r.i_abs = p.i_abs + base; }
    if (p.i_rel != ~0ull && (p.max_rel > r.max_rel || (p.max_rel == r.max_rel && p.i_rel + base < r.i_rel))) { r.max_rel = p.max_rel; r.i_rel = p.i_rel + base; }
    r.sum += p.sum;
    r.nan += p.nan;
}

static void err_finish(const FzErrPartial &r, uint64_t nwords, float a1, float a2, float r1, float r2, mzb_error_report_t *out)
{
    memset(out, 0, sizeof(*out));
    out->count = nwords;
    out->nan_count = r.nan;
    out->max_abs_index = r.i_abs;
    out->max_rel_index = r.i_rel;
    out->max_abs_err = r.i_abs == ~0ull ? 0.f : r.max_abs;
    out->max_rel_err = r.i_rel == ~0ull ? 0.f : r.max_rel;
    out->max_abs_n1 = a1; out->max_abs_n2 = a2; out->max_rel_n1 = r1; out->max_rel_n2 = r2;
    out->sum_abs_err = r.sum;
}

static float err_other_value(uint32_t a, uint64_t i, int bits, uint32_t exempt)
{
    const uint32_t b = i >= exempt ? (a & fz_mask_for_bits(bits)) : a;
    float f;
    memcpy(&f, &b, 4);
    return f;
}

extern "C" int mzb_error_report_device(mzb_ctx *c, const void *d_orig, const void *d_other, uint64_t nwords, int bits,
                                       uint32_t exempt_words, mzb_error_report_t *out)
{
    if (!c || !out || !d_orig || ((uintptr_t)d_orig & 3) || ((uintptr_t)d_other & 3) || (!d_other && (bits < 0 || bits > 32)))
        return MZB_E_ARG;
    FZ_CHECK(cudaSetDevice(c->device));
    FzErrPartial r;
    r.max_abs = -1.f; r.max_rel = -1.f; r.i_abs = ~0ull; r.i_rel = ~0ull; r.sum = 0.0; r.nan = 0;
    float v[4] = {0, 0, 0, 0};
    if (nwords) {
        int rc;
        if ((rc = ensure(c->err_partial, (size_t)fz_error_partials() * sizeof(FzErrPartial)))) return rc;
        fz_launch_error((const uint32_t *)d_orig, (const uint32_t *)d_other, nwords, fz_mask_for_bits(d_other ? 0 : bits), exempt_words,
                        (FzErrPartial *)c->err_partial.p, c->stream);
        FZ_CHECK(cudaGetLastError());
        FzErrPartial p;
        FZ_CHECK(cudaMemcpyAsync(&p, c->err_partial.p, sizeof(p), cudaMemcpyDeviceToHost, c->stream));
        FZ_CHECK(cudaStreamSynchronize(c->stream));
        err_merge(r, p, 0);
        const uint64_t idx[2] = {r.i_abs, r.i_rel};
        for (int k = 0; k < 2; k++) {
            if (idx[k] == ~0ull) continue;
            uint32_t a = 0, b = 0;
            FZ_CHECK(cudaMemcpy(&a, (const uint32_t *)d_orig + idx[k], 4, cudaMemcpyDeviceToHost));
            memcpy(&v[2 * k], &a, 4);
            if (d_other) { FZ_CHECK(cudaMemcpy(&b, (const uint32_t *)d_other + idx[k], 4, cudaMemcpyDeviceToHost)); memcpy(&v[2 * k + 1], &b, 4); }
            else v[2 * k + 1] = err_other_value(a, idx[k], bits, exempt_words);
        }
    }
    err_finish(r, nwords, v[0], v[1], v[2], v[3], out);
    return MZB_OK;
}

extern "C" int mzb_error_report_host(mzb_ctx *c, const void *h_orig, const void *h_other, uint64_t nwords, int bits,
                                     uint32_t exempt_words, mzb_error_report_t *out)
{
    if (!c || !out || !h_orig || (!h_other && (bits < 0 || bits > 32))) return MZB_E_ARG;
    FZ_CHECK(cudaSetDevice(c->device));
    const uint64_t batch = 64ull << 20;   // words per batch (256 MiB per buffer)
    const uint64_t bw = nwords < batch ? nwords : batch;
    int rc;
    if (nwords && ((rc = ensure(c->io_in, bw * 4 + 16)) || (h_other && (rc = ensure(c->io_out, bw * 4 + 16))) ||
                   (rc = ensure(c->err_partial, (size_t)fz_error_partials() * sizeof(FzErrPartial)))))
        return rc;
    FzErrPartial r;
    r.max_abs = -1.f; r.max_rel = -1.f; r.i_abs = ~0ull; r.i_rel = ~0ull; r.sum = 0.0; r.nan = 0;
    const uint32_t *ho = (const uint32_t *)h_orig, *hx = (const uint32_t *)h_other;
    for (uint64_t w0 = 0; w0 < nwords; w0 += bw) {
        const uint64_t n = nwords - w0 < bw ? nwords - w0 : bw;
        FZ_CHECK(cudaMemcpyAsync(c->io_in.p, ho + w0, n * 4, cudaMemcpyHostToDevice, c->stream));
        if (hx) FZ_CHECK(cudaMemcpyAsync(c->io_out.p, hx + w0, n * 4, cudaMemcpyHostToDevice, c->stream));
        const uint64_t ex = exempt_words > w0 ? exempt_words - w0 : 0;
        fz_launch_error((const uint32_t *)c->io_in.p, hx ? (const uint32_t *)c->io_out.p : nullptr, n, fz_mask_for_bits(hx ? 0 : bits), ex,
                        (FzErrPartial *)c->err_partial.p, c->stream);
        FzErrPartial p;
        FZ_CHECK(cudaMemcpyAsync(&p, c->err_partial.p, sizeof(p), cudaMemcpyDeviceToHost, c->stream));
        FZ_CHECK(cudaStreamSynchronize(c->stream));
        FZ_CHECK(cudaGetLastError());
        err_merge(r, p, w0);
    }
    float v[4] = {0, 0, 0, 0};
    const uint64_t idx[2] = {r.i_abs, r.i_rel};
    for (int k = 0; k < 2; k++) {
        if (idx[k] == ~0ull) continue;
        memcpy(&v[2 * k], ho + idx[k], 4);
        if (hx) memcpy(&v[2 * k + 1], hx + idx[k], 4);
        else v[2 * k + 1] = err_other_value(ho[idx[k]], idx[k], bits, exempt_words);
    }
    err_finish(r, nwords, v[0], v[1], v[2], v[3], out);
    return MZB_OK;
}

extern "C" int mzb_mrc_parse(const void *header, size_t len, mzb_mrc_info *out)
{
    if (!header || !out || len < 1024) return MZB_E_ARG;
    int32_t w[24];
    memcpy(w, header, sizeof(w));
    memset(out, 0, sizeof(*out));
    out->nx = w[0]; out->ny = w[1]; out->nz = w[2]; out->mode = w[3]; out->next = w[23];
    out->is_float32 = w[3] == 2;
    out->data_offset = 1024ull + (uint64_t)(w[23] > 0 ? w[23] : 0);
    const bool mode_ok = w[3] == 0 || w[3] == 1 || w[3] == 2 || w[3] == 3 || w[3] == 4 || w[3] == 6 || w[3] == 12 || w[3] == 16 || w[3] == 101;
    if (w[0] <= 0 || w[1] <= 0 || w[2] <= 0 || !mode_ok || w[23] < 0) return MZB_E_FORMAT;
    return MZB_OK;
}

// ---------------------------------------------------------------------------------------------------
// several small inputs in one pass of the kernels: every chunk of the batch gets its own word count and exemption
// (FzBatchGeom::chunk_n / chunk_exempt), item i owns the chunk slots [first[i], first[i + 1])
extern "C" int mzb_compress_host_many(mzb_ctx *c, mzb_zip_item *items, uint32_t n, int bits, uint32_t chk, int write_file_header)
{
    if (!c || (!items && n) || bits < 0 || bits > 32 || chk == 0 || chk >= 0x80000000u || (chk % 16u)) return MZB_E_ARG;
    memset(&c->stats, 0, sizeof(c->stats));
    std::vector<uint32_t> first(n + 1, 0), tab;
    for (uint32_t i = 0; i < n; i++) {
        items[i].out_size = 0;
        if (!items[i].h_out || (items[i].nwords && !items[i].h_words)) return MZB_E_ARG;
        first[i + 1] = first[i] + (uint32_t)((items[i].nwords + chk - 1) / chk);
    }
    const uint32_t nchunks = first[n];
    if (nchunks > c->batch_chunks) return MZB_E_ARG;
    if (nchunks == 0) return MZB_OK;   // empty inputs write nothing, not even the header (workers.c:757-764)
    FZ_CHECK(cudaSetDevice(c->device));
    tab.resize(2 * (size_t)nchunks);
    for (uint32_t i = 0; i < n; i++)
        for (uint32_t k = first[i]; k < first[i + 1]; k++) {
            const uint64_t w0 = (uint64_t)(k - first[i]) * chk;
            tab[k] = (uint32_t)(items[i].nwords - w0 < chk ? items[i].nwords - w0 : chk);
            tab[nchunks + k] = (uint32_t)(items[i].exempt_words > w0 ? (items[i].exempt_words - w0 < chk ? items[i].exempt_words - w0 : chk) : 0);
        }
    const size_t slot_bytes = (size_t)chk * 4;
    const size_t bound = (size_t)nchunks * (FZ_CHUNK_HEADER_BYTES + slot_bytes) + 64;
    uint64_t pstride;
    int rc;
    if ((rc = compress_reserve(c, nchunks, chk, &pstride)) || (rc = ensure(c->io_in, (size_t)nchunks * slot_bytes + 256)) ||
        (rc = ensure(c->io_out, bound + 256)) || (rc = ensure(c->chunk_tab, tab.size() * 4)) ||
        (rc = ensure(c->stream_hdr, (size_t)nchunks * FZ_PLANES * 4)))
        return rc;
    if ((rc = status_reset(c, 0))) return rc;
    prof_begin(c);
    FZ_CHECK_PIPE(cudaMemcpyAsync(c->chunk_tab.p, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, c->stream));
    for (uint32_t i = 0; i < n; i++)
        if (items[i].nwords)
            FZ_CHECK_PIPE(cudaMemcpyAsync((uint8_t *)c->io_in.p + (size_t)first[i] * slot_bytes, items[i].h_words, items[i].nwords * 4,
                                          cudaMemcpyHostToDevice, c->stream));
    FzBatchGeom g = make_geom(nchunks, chk, (uint64_t)nchunks * chk, pstride);
    g.chunk_n = (const uint32_t *)c->chunk_tab.p;
    g.chunk_exempt = g.chunk_n + nchunks;
    const uint32_t mask = fz_mask_for_bits(bits);
    uint8_t *d_out = (uint8_t *)c->io_out.p;
    __atomic_fetch_add(&g_passes[0], 1ull, __ATOMIC_RELAXED);
    fz_launch_split_slots((const uint32_t *)c->io_in.p, nchunks, chk, mask, g.chunk_exempt, (uint8_t *)c->planes.p, pstride, c->stream);
    prof_mark(c, FZ_ST_SPLIT);
    if (!c->zero_hist_ready) { fz_launch_zero_hist((uint32_t *)c->zero_hist.p, c->stream); c->zero_hist_ready = true; }
    uint32_t zero_planes = 0;
    for (int j = 0; j < FZ_PLANES; j++)
        if (((mask >> (8 * j)) & 0xffu) == 0) zero_planes |= 1u << j;
    fz_launch_encode((const uint8_t *)c->planes.p, g, (uint32_t *)c->ghist.p, c->gcodes.p, (uint8_t *)c->scratch.p, (uint32_t *)c->sizes.p,
                     (const uint32_t *)c->zero_hist.p, zero_planes, 0, c->d_status, c->stream);
    prof_mark(c, FZ_ST_ENCODE);
    fz_launch_layout((uint32_t *)c->sizes.p, g, (uint32_t *)c->sub_off.p, (uint32_t *)c->stream_hdr.p, (unsigned long long *)c->stream_off.p,
                     d_out, bound, c->d_status, c->stream);
    prof_mark(c, FZ_ST_LAYOUT);
    fz_launch_gather((const uint8_t *)c->planes.p, (const uint8_t *)c->scratch.p, (const uint32_t *)c->sizes.p, (const uint32_t *)c->sub_off.p,
                     (const uint32_t *)c->stream_hdr.p, (const unsigned long long *)c->stream_off.p, g, d_out, c->d_status, c->stream);
    prof_mark(c, FZ_ST_GATHER);
    // record sizes of every chunk -> where each item's records lie in the batch's container
    std::vector<uint32_t> hdr((size_t)nchunks * FZ_PLANES);
    FZ_CHECK_PIPE(cudaMemcpyAsync(hdr.data(), c->stream_hdr.p, hdr.size() * 4, cudaMemcpyDeviceToHost, c->stream));
    if ((rc = status_fetch(c))) return pipe_fail(c, rc);
    prof_collect(c);
    fill_compress_stats(c, 0, nchunks, 8);
    if (c->h_status->error) return c->h_status->error;
    uint64_t off = 0;
    int result = MZB_OK;
    for (uint32_t i = 0; i < n; i++) {
        uint64_t bytes = 0;
        for (uint32_t k = first[i]; k < first[i + 1]; k++) {
            bytes += FZ_CHUNK_HEADER_BYTES;
            for (int j = 0; j < FZ_PLANES; j++) bytes += hdr[(size_t)k * FZ_PLANES + j] & ~FZ_RAW_FLAG;
        }
        c->stats.bytes_in += items[i].nwords * 4;
        if (items[i].nwords == 0) continue;
        const size_t hb = write_file_header ? MZB_FILE_HEADER_BYTES : 0;
        if (items[i].out_cap < hb + bytes) { result = MZB_E_SPACE; off += bytes; continue; }
        uint8_t *o = (uint8_t *)items[i].h_out;
        if (write_file_header) {   // common.c:137-149
            memset(o, 0, MZB_FILE_HEADER_BYTES);
            memcpy(o, &items[i].fsz, 8);
            memcpy(o + 8, &chk, 4);
        }
        FZ_CHECK_PIPE(cudaMemcpyAsync(o + hb, d_out + off, bytes, cudaMemcpyDeviceToHost, c->stream));
        items[i].out_size = hb + bytes;
        off += bytes;
    }
    FZ_CHECK_PIPE(cudaStreamSynchronize(c->stream));
    return result;
}

extern "C" int mzb_decompress_host_many(mzb_ctx *c, mzb_unzip_item *items, uint32_t n, uint32_t chk)
{
    if (!c || (!items && n) || chk == 0 || chk >= 0x80000000u || (chk % 16u)) return MZB_E_ARG;
    memset(&c->stats, 0, sizeof(c->stats));
    std::vector<uint32_t> first(n + 1, 0), tab;
    size_t in_total = 0;
    for (uint32_t i = 0; i < n; i++) {
        if (items[i].nwords > items[i].out_cap_words) return MZB_E_SPACE;
        if (items[i].nwords && (!items[i].h_in || !items[i].h_words_out)) return MZB_E_ARG;
        first[i + 1] = first[i] + (uint32_t)((items[i].nwords + chk - 1) / chk);
        if (items[i].nwords) in_total += items[i].in_size;
    }
    const uint32_t nchunks = first[n];
    if (nchunks > c->batch_chunks) return MZB_E_ARG;
    if (nchunks == 0) return MZB_OK;
    FZ_CHECK(cudaSetDevice(c->device));
    tab.resize(nchunks);
    for (uint32_t i = 0; i < n; i++)
        for (uint32_t k = first[i]; k < first[i + 1]; k++) {
            const uint64_t w0 = (uint64_t)(k - first[i]) * chk;
            tab[k] = (uint32_t)(items[i].nwords - w0 < chk ? items[i].nwords - w0 : chk);
        }
    const size_t slot_bytes = (size_t)chk * 4;
    uint64_t pstride;
    FzInflateBufs ib;
    int rc;
    if ((rc = decompress_reserve(c, nchunks, chk, &pstride, &ib)) || (rc = ensure(c->io_in, in_total + 256)) ||
        (rc = ensure(c->io_out, (size_t)nchunks * slot_bytes + 256)) || (rc = ensure(c->chunk_tab, tab.size() * 4)))
        return rc;
    if ((rc = status_reset(c, 0))) return rc;
    prof_begin(c);
    FZ_CHECK_PIPE(cudaMemcpyAsync(c->chunk_tab.p, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, c->stream));
    size_t at = 0;
    for (uint32_t i = 0; i < n; i++)
        if (items[i].nwords) {   // the items' records back to back: one chain for the walk
            FZ_CHECK_PIPE(cudaMemcpyAsync((uint8_t *)c->io_in.p + at, items[i].h_in, items[i].in_size, cudaMemcpyHostToDevice, c->stream));
            at += items[i].in_size;
        }
    FzBatchGeom g = make_geom(nchunks, chk, (uint64_t)nchunks * chk, pstride);
    g.chunk_n = (const uint32_t *)c->chunk_tab.p;
    decompress_enqueue_batch(c, (const uint8_t *)c->io_in.p, in_total, g, ib, (uint32_t *)c->io_out.p, true);
    for (uint32_t i = 0; i < n; i++)
        if (items[i].nwords)
            FZ_CHECK_PIPE(cudaMemcpyAsync(items[i].h_words_out, (const uint8_t *)c->io_out.p + (size_t)first[i] * slot_bytes, items[i].nwords * 4,
                                          cudaMemcpyDeviceToHost, c->stream));
    if ((rc = status_fetch(c))) return pipe_fail(c, rc);
    prof_collect(c);
    fill_decompress_stats(c, in_total, 0, nchunks, decompress_launches(0, true));
    for (uint32_t i = 0; i < n; i++) c->stats.bytes_out += items[i].nwords * 4;
    if (c->h_status->error) return c->h_status->error;
    if (c->h_status->out_end != in_total) return MZB_E_FORMAT;   // an item's records are longer or shorter than it said
    return MZB_OK;
}
