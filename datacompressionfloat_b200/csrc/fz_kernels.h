// fz_kernels.h -- launch wrappers of the sm_100a kernels (internal C++ interface between fz_kernels.cu and fz_api.cu)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fz_common.cuh"
#include "fz_blockpar.cuh"

// geometry of one batch of chunks handed to the kernels
struct FzBatchGeom {
    uint32_t nchunks;      // chunks in this batch
    uint32_t chk;          // words per full chunk (== plane bytes per full stream)
    uint32_t last_n;       // words in the last chunk of the batch (== chk unless it is the ragged file tail)
    uint32_t nsub_full;    // ceil(chk / FZ_SUB): sub-block slots reserved per stream
    uint64_t plane_stride; // bytes between plane j and plane j+1 in the plane buffer
    // several small files in one batch (device tables, nullptr otherwise): chunk c has chunk_n[c] <= chk words in the
    // slot [c * chk, (c + 1) * chk) of the word buffer and of every plane; its first chunk_exempt[c] words are not masked
    const uint32_t *chunk_n;
    const uint32_t *chunk_exempt;
};


// device status block (one per context)
struct FzStatus {
    unsigned long long out_end;   // running end offset of the container being written / read
    int error;                    // first FZ_E_* raised by a kernel
    unsigned int n_general;       // streams decoded by the general (one thread per stream) inflater
    unsigned int n_fast_failed;   // streams whose sub-block decode failed validation and fell back
    unsigned int n_stored_sub;    // sub-blocks emitted as stored blocks
    unsigned int n_raw_streams;   // streams written RAW
    unsigned int n_blockpar;      // general streams decoded block-parallel (the rest took the serial inflater)
    unsigned int n_zero_sub;      // sub-blocks found to be all zero bytes (encode)
    unsigned int pad0;
};

// optional per-stage timing hook: called after the launches of a stage were enqueued
typedef void (*fz_mark_fn)(void *user, int stage);
enum {
    FZ_ST_SPLIT = 0, FZ_ST_ENCODE, FZ_ST_LAYOUT, FZ_ST_GATHER,
    FZ_ST_WALK, FZ_ST_MARKERS, FZ_ST_CLASSIFY, FZ_ST_INFLATE_FAST, FZ_ST_INFLATE_BLOCKPAR, FZ_ST_INFLATE_GENERAL, FZ_ST_RAWCOPY, FZ_ST_MERGE,
    FZ_ST_COUNT
};

// ---- mask + byte-plane split / merge (HBM-bound)
// skip_planes (variant 0 only): bit j = byte plane j is not written for words in [skip_lo, skip_hi) -- the caller knows
// the mask erases that plane and that nothing reads those bytes (whole sub-blocks the encoder flags all-zero unseen)
void fz_launch_split(const uint32_t *words, uint64_t nwords, uint32_t mask, uint64_t exempt_words,
                     uint8_t *planes, uint64_t plane_stride, int variant, cudaStream_t st,
                     uint32_t skip_planes = 0, uint64_t skip_lo = 0, uint64_t skip_hi = 0);
void fz_launch_merge(const uint8_t *planes, uint64_t plane_stride, uint64_t nwords, uint32_t *words,
                     int variant, cudaStream_t st);
// the split of a batch of chunk slots with a per-chunk exemption table (several small files in one batch; chk % 4 == 0)
void fz_launch_split_slots(const uint32_t *words, uint32_t nchunks, uint32_t chk, uint32_t mask, const uint32_t *chunk_exempt,
                           uint8_t *planes, uint64_t plane_stride, cudaStream_t st);

// ---- deflate side
// ghist: 288 uint32 per group (zeroed by the call); gcodes: fz_group_code_bytes() per group;
// groups per stream = ceil(nsub_full / FZ_GROUP_SUBS)
// zero_hist: 288 uint32 made once per context by fz_launch_zero_hist (token histogram of an all-zero sub-block)
void fz_launch_zero_hist(uint32_t *zero_hist, cudaStream_t st);
// zero_planes: bit j = the mask erases byte plane j entirely; zero_from: first word of the batch that is masked
// (sub-blocks of such planes behind it are all zero and are not even read)
void fz_launch_encode(const uint8_t *planes, FzBatchGeom g, uint32_t *ghist, void *gcodes, uint8_t *scratch, uint32_t *sizes,
                      const uint32_t *zero_hist, uint32_t zero_planes, uint64_t zero_from, FzStatus *status, cudaStream_t st);
size_t fz_group_code_bytes();
// stream sums + RAW decision + scan over chunk records + chunk headers; container offsets continue from status->out_end
void fz_launch_layout(uint32_t *sizes, FzBatchGeom g, uint32_t *sub_off, uint32_t *stream_hdr,
                      unsigned long long *stream_off, uint8_t *container, uint64_t container_cap, FzStatus *status,
                      cudaStream_t st);
void fz_launch_gather(const uint8_t *planes, const uint8_t *scratch, const uint32_t *sizes, const uint32_t *sub_off,
                      const uint32_t *stream_hdr, const unsigned long long *stream_off, FzBatchGeom g,
                      uint8_t *container, const FzStatus *status, cudaStream_t st);

// ---- inflate side
// walks nchunks chunk records starting at status->out_end, fills the stream table, advances status->out_end
void fz_launch_walk(const uint8_t *container, uint64_t container_size, FzBatchGeom g, uint32_t *stream_hdr,
                    unsigned long long *stream_off, FzStatus *status, cudaStream_t st);
// marker scan (count), scan, marker scan (write), classify, fast inflate, general inflate, RAW copy
// scratch of the block-parallel inflate of zlib-made streams (fz_blockpar.cuh)
#define FZ_BP_CAP 512          // candidate blocks per stream (a 6 MiB plane is ~190 zlib blocks)
#define FZ_BP_STORED_CAP 128   // stored blocks per stream met on the chain
struct FzBlockParBufs {
    uint32_t *ctl;         // counters, zeroed per batch: FZ_BP_CTL_*
    uint32_t *gen_list;    // [nstreams] streams classified "general"
    uint32_t *cand_cnt;    // [nstreams]
    uint32_t *nstored;     // [nstreams]
    uint32_t *par_ok;      // [nstreams] 1 = decoded block-parallel
    uint32_t *cand_pos;    // [nstreams * FZ_BP_CAP] bit positions, ascending after the sort
    FzBlockInfo *info;     // [nstreams * FZ_BP_CAP]
    uint32_t *blk_off;     // [nstreams * FZ_BP_CAP] output offset, ~0 = not on the chain
    int *blk_prev;         // [nstreams * FZ_BP_CAP] byte before the block, -1 = none
    FzStoredItem *stored;  // [nstreams * FZ_BP_STORED_CAP]
    uint32_t *items;       // [2][nstreams * FZ_BP_CAP] work lists (block slot indices) of the measure / write pass
    uint32_t *first_rec;   // [nstreams * FZ_BP_CAP] first tile record of the block, FZ_TILE_NONE = none kept
    FzTileRec *tiles;      // [tiles_cap] tile records of the measure pass (bump allocated through ctl[FZ_BP_CTL_TILES])
    uint32_t tiles_cap;
    uint32_t nstreams;
};
enum { FZ_BP_CTL_NGEN = 0, FZ_BP_CTL_NMEASURE, FZ_BP_CTL_CUR_MEASURE, FZ_BP_CTL_NWRITE, FZ_BP_CTL_CUR_WRITE, FZ_BP_CTL_TILES };
size_t fz_blockpar_bytes(uint32_t nstreams, uint32_t chk);
FzBlockParBufs fz_blockpar_carve(void *blob, uint32_t nstreams, uint32_t chk);

struct FzInflateBufs {
    uint32_t *tile_cnt;      // [nstreams * tiles_per_stream] look-back state of the marker scan (zeroed per batch)
    uint32_t *stream_cnt;    // [nstreams] markers found in every stream
    uint32_t *hits;          // marker positions (stream relative): stream s at hits + s * hits_per_stream, in order
    uint32_t hits_per_stream;
    uint32_t *stream_mode;   // [nstreams] 0 raw, 1 fast | sub_log2 << 8, 2 general
    uint32_t *stream_fail;   // [nstreams]
    uint32_t *zero_flags;    // [nstreams * nsub_full] 1 = the sub-block is all zero bytes and was NOT written to the plane buffer
    uint32_t tiles_per_stream;
    void *group_desc;        // [fz_group_desc_bytes()] one descriptor per code group: what the header pass leaves for the lean inflater
    bool full_only;          // tests: every group goes to the full group inflater (the lean kernel's fallback)
    FzBlockParBufs bp;
};
size_t fz_group_desc_bytes(uint32_t nstreams, uint32_t nsub_full);
void fz_launch_inflate(const uint8_t *container, uint64_t container_size, FzBatchGeom g, const uint32_t *stream_hdr,
                       const unsigned long long *stream_off, FzInflateBufs b, uint8_t *planes, FzStatus *status,
                       cudaStream_t st, fz_mark_fn mark, void *mark_user, bool copy_raw);
// merge that reads RAW streams in place from the container (needs chk % 16 == 0)
void fz_launch_merge_streams(const uint8_t *planes, const uint8_t *container, uint64_t container_size, const uint32_t *stream_hdr,
                             const unsigned long long *stream_off, const uint32_t *zero_flags, FzBatchGeom g, uint32_t *words,
                             cudaStream_t st);


// ---- error report (reference src/tool/erroranalysis.c:188-220)
struct FzErrPartial {
    float max_abs, max_rel;
    unsigned long long i_abs, i_rel;   // word index of the maxima (~0: none)
    double sum;                        // sum of the absolute errors
    unsigned long long nan;            // pairs whose error is not a number
};
uint32_t fz_error_partials();
// other == nullptr: compare the words with their own masked form (mask behind `exempt` words); result in partial[0]
void fz_launch_error(const uint32_t *orig, const uint32_t *other, uint64_t nwords, uint32_t mask, uint64_t exempt,
                     FzErrPartial *partial, cudaStream_t st);
