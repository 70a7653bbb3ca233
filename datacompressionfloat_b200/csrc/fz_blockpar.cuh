// fz_blockpar.cuh -- block-parallel inflate of streams written by zlib (the reference's own containers).
//
// A reference payload (reference zip.c:164-196: deflate(level 6, Z_RLE), Z_FULL_FLUSH per chunk) is ~170
// dynamic-Huffman blocks back to back with no byte alignment between them, so it offers no index.  We
// recover one:
//   1. candidates   every bit position is tested for "a dynamic block header starts here": the 17 fixed
//                   header bits, a complete code-length code, then a full parse of the code lengths with
//                   complete literal/length code and an end-of-block symbol (false positives are rare and harmless)
//   2. measure      each candidate block is decoded without storing anything: end bit, output bytes, whether it
//                   starts with a run that continues from the previous block, its last literal
//   3. chain        one cheap serial walk per stream: start at bit 0, hop from block end to the candidate that
//                   starts there (or over a stored block), accumulating output offsets; any inconsistency makes
//                   the stream fall back to the serial inflater
//   4. write        every block on the chain is decoded again, now storing at its offset
// Only distance-1 matches are supported across and inside blocks during the measure pass (what Z_RLE
// produces); a stream with other distances fails the chain and takes the serial path.
//
// `__host__ __device__`: tests/hostmodel runs the same functions on the CPU against zlib streams.
#pragma once
#include "fz_inflate.cuh"

struct FzBlockInfo {
    uint32_t end_bit;   // bit position (in the stream) just after the block's end-of-block symbol
    uint32_t out_len;   // bytes the block inflates to
    uint32_t flags;
    uint32_t last;      // last literal of the block (valid with FZ_BLK_HAS_LITERAL)
};
#define FZ_BLK_OK 1u
#define FZ_BLK_NON_RLE 2u
#define FZ_BLK_STARTS_WITH_MATCH 4u
#define FZ_BLK_HAS_LITERAL 8u
#define FZ_BLK_FINAL 16u

struct FzStoredItem {
    uint32_t src_byte;  // offset of the data in the stream
    uint32_t len;
    uint32_t out_off;
};

// Kraft weight of three packed 3-bit code lengths: sum of 2^(7-len) over the non-zero ones (<= 192)
FZ_HD uint32_t fz_kraft3(uint32_t three)
{
    uint32_t k = 0;
    for (int i = 0; i < 3; i++) { const uint32_t v = (three >> (3 * i)) & 7u; if (v) k += 128u >> v; }
    return k;
}

// cheap test on the first bits of a would-be dynamic block header; lo = stream bits [p, p+64), hi = [p+64, p+128).
// kraft_lut: 512 bytes, kraft_lut[x] = fz_kraft3(x) (shared memory on the GPU), or nullptr
FZ_HD bool fz_block_quick_test(uint64_t lo, uint64_t hi, const uint8_t *kraft_lut = nullptr)
{
    if ((lo & 7u) != 4u) return false;                        // BFINAL = 0, BTYPE = 10 (a final block: see fz_chain_resolve)
    if (((lo >> 3) & 31u) > 29u) return false;                // HLIT  <= 29 (286 codes)
    if (((lo >> 8) & 31u) > 29u) return false;                // HDIST <= 29
    const uint32_t ncl = (uint32_t)((lo >> 13) & 15u) + 4u;
    // the code-length code must be complete (zlib's inflate rejects anything else): sum of 2^(7-len) == 128.
    // The ncl 3-bit lengths start at bit 17: keep exactly those, nine bits (three lengths) per table lookup.
    uint64_t f = (lo >> 17) | (hi << 47);                     // 57 bits = 19 fields
    f &= (1ull << (3 * ncl)) - 1ull;                          // ncl <= 19: shift <= 57
    uint32_t kraft = 0;
    if (kraft_lut) {
#pragma unroll
        for (int i = 0; i < 7; i++) kraft += kraft_lut[(uint32_t)(f >> (9 * i)) & 511u];
    } else {
        for (int i = 0; i < 7; i++) kraft += fz_kraft3((uint32_t)(f >> (9 * i)) & 511u);
    }
    return kraft == 128u;
}

// Per-thread 128-entry table for the code-length code (7-bit index -> symbol | length << 5), STRIDE threads
// interleaved word by word so that the lanes of a warp never collide on a bank.
template <int STRIDE>
struct FzClLut {
    uint8_t *base;   // this thread's first word
    FZ_HD uint8_t &at(uint32_t e) const { return base[(e >> 2) * (STRIDE * 4) + (e & 3u)]; }
};

// Second-stage test of a position that passed fz_block_quick_test: walk the code-length data of the header with
// registers and the small table only -- no symbol tables are built -- and apply the same acceptance rules as
// fz_block_candidate (a complete literal/length code with an end-of-block symbol, distances not over-subscribed).
// ~30x cheaper than the full parse, and it rejects practically everything that is not a real header.
template <class ClLut>
FZ_HD bool fz_block_precheck(const uint8_t *in, size_t in_len, uint64_t bit, const ClLut &lut)
{
    FzBitReader br;
    br.init(in + (bit >> 3), in_len - (size_t)(bit >> 3));
    br.refill();
    br.drop((int)(bit & 7));
    br.refill();
    br.drop(3);                                  // BFINAL, BTYPE (checked by the quick test)
    const uint32_t hlit = br.get(5) + 257, hdist = br.get(5) + 1, hclen = br.get(4) + 4;
    if (hlit > 286 || hdist > 30) return false;
    const uint64_t order_lo = 16ull | (17ull << 5) | (18ull << 10) | (0ull << 15) | (8ull << 20) | (7ull << 25) | (9ull << 30) |
                              (6ull << 35) | (10ull << 40) | (5ull << 45) | (11ull << 50) | (4ull << 55);
    const uint64_t order_hi = 12ull | (3ull << 5) | (13ull << 10) | (2ull << 15) | (14ull << 20) | (1ull << 25) | (15ull << 30);
    uint64_t clpack = 0;                         // 3 bits per symbol
    for (uint32_t i = 0; i < hclen; i++) {
        br.refill();
        const uint32_t sym = (uint32_t)((i < 12 ? order_lo >> (5 * i) : order_hi >> (5 * (i - 12))) & 31u);
        clpack |= (uint64_t)br.get(3) << (3 * sym);
    }
    // canonical codes of the code-length code -> table (the quick test made sure the code is complete)
    uint64_t ccnt = 0;                           // 8 bits per length 0..7
    for (int sy = 0; sy < 19; sy++) ccnt += 1ull << (8 * ((clpack >> (3 * sy)) & 7u));
    uint64_t next = 0;                           // next code per length, 8 bits each
    {
        uint32_t code = 0;
        for (int l = 1; l <= 7; l++) {
            code = (code + (uint32_t)((ccnt >> (8 * (l - 1))) & 0xffu) * (l > 1 ? 1u : 0u)) << 1;
            next |= (uint64_t)(code & 0xffu) << (8 * l);
        }
    }
    for (uint32_t sy = 0; sy < 19; sy++) {
        const uint32_t l = (uint32_t)((clpack >> (3 * sy)) & 7u);
        if (!l) continue;
        const uint32_t c = (uint32_t)((next >> (8 * l)) & 0xffu);
        next += 1ull << (8 * l);
        uint32_t rev = 0;
        for (uint32_t k = 0; k < l; k++) rev |= ((c >> k) & 1u) << (l - 1 - k);
        for (uint32_t e = rev; e < 128u; e += 1u << l) lut.at(e) = (uint8_t)(sy | (l << 5));
    }
    const uint32_t total = hlit + hdist;
    uint32_t i = 0, prev = 0, kll = 0, kdd = 0, eob = 0;
    while (i < total) {
        br.refill();
        const uint32_t e = lut.at(br.peek(7));
        br.drop((int)(e >> 5));
        const uint32_t sy = e & 31u;
        uint32_t rep = 1, val = sy;
        if (sy >= 16) {
            if (sy == 16 && i == 0) return false;
            const int nb = sy == 16 ? 2 : (sy == 17 ? 3 : 7);
            rep = (sy == 18 ? 11u : 3u) + br.get(nb);
            val = sy == 16 ? prev : 0u;
        }
        if (i + rep > total) return false;
        prev = val;
        if (val) {
            const uint32_t n_ll = i < hlit ? (rep < hlit - i ? rep : hlit - i) : 0u;
            kll += n_ll << (15 - val);
            kdd += (rep - n_ll) << (15 - val);
            if (i <= FZ_EOB && FZ_EOB < i + rep) eob = val;
            if (kll > 32768u || kdd > 32768u) return false;   // over-subscribed
        }
        i += rep;
    }
    return kll == 32768u && eob != 0 && br.bits_left() >= 0;
}

// full validation: parse the header exactly like the inflater would
template <class Tab>
FZ_HD bool fz_block_candidate(const uint8_t *in, size_t in_len, uint64_t bit, const Tab &tab)
{
    FzInflater<Tab> inf;
    FzCode ll, dd;
    inf.start_at_bit(in, in_len, bit, nullptr, 0, tab);
    inf.bind_codes(&ll, &dd);
    inf.bw.dry = true;
    if (!inf.block_header() || !inf.in_body) return false;
    // zlib always emits a complete literal/length code with an end-of-block symbol; distances: complete, or the
    // degenerate single code
    return inf.ll_left == 0 && inf.eob_len != 0 && inf.dd_left >= 0;
}

template <class Tab>
FZ_HD void fz_block_measure(const uint8_t *in, size_t in_len, uint64_t bit, const Tab &tab, uint32_t *lut, FzBlockInfo *bi)
{
    FzInflater<Tab> inf;
    FzCode ll, dd;
    inf.start_at_bit(in, in_len, bit, nullptr, 0xFFFFFFF0u, tab);
    inf.bind_codes(&ll, &dd);
    inf.bw.dry = true;
    inf.bw.prev_byte = 0;       // a distance-1 run may continue from the previous block; its value is resolved later
    inf.own_lut = lut;
    inf.one_block = true;
    while (inf.step()) {}
    bi->end_bit = (uint32_t)((bit & ~7ull) + inf.consumed_bits());
    bi->out_len = inf.bw.produced();
    uint32_t f = 0;
    if (inf.rc == FZ_INF_OK && inf.saw_eob && inf.br.bits_left() >= 0) f |= FZ_BLK_OK;
    if (inf.bw.non_rle) f |= FZ_BLK_NON_RLE;
    if (inf.bw.starts_with_match) f |= FZ_BLK_STARTS_WITH_MATCH;
    if (inf.bw.lastc < 0x100u) f |= FZ_BLK_HAS_LITERAL;
    if (inf.last) f |= FZ_BLK_FINAL;
    bi->flags = f;
    bi->last = inf.bw.lastc & 0xffu;
}

// decode the block at `bit` into out[0, out_len); prev_byte = the byte before out[0] (or -1 at the stream start)
template <class Tab>
FZ_HD bool fz_block_write(const uint8_t *in, size_t in_len, uint64_t bit, const Tab &tab, uint32_t *lut, uint8_t *out,
                          uint32_t out_len, int prev_byte, uint32_t expect_end_bit)
{
    FzInflater<Tab> inf;
    FzCode ll, dd;
    inf.start_at_bit(in, in_len, bit, out, out_len, tab);
    inf.bind_codes(&ll, &dd);
    inf.bw.prev_byte = prev_byte;
    inf.own_lut = lut;
    inf.one_block = true;
    while (inf.step()) {}
    inf.bw.finish();
    const uint32_t end_bit = (uint32_t)((bit & ~7ull) + inf.consumed_bits());
    return inf.rc == FZ_INF_OK && inf.saw_eob && inf.bw.produced() == out_len && end_bit == expect_end_bit;
}

FZ_HD uint32_t fz_stream_bits(const uint8_t *in, uint64_t bit, uint32_t n)  // n <= 24 bits at bit position `bit`
{
    const uint64_t by = bit >> 3;
    const uint32_t v = (uint32_t)in[by] | ((uint32_t)in[by + 1] << 8) | ((uint32_t)in[by + 2] << 16) | ((uint32_t)in[by + 3] << 24);
    return (v >> (bit & 7)) & ((1u << n) - 1u);
}

// Serial walk over one stream.  cand_pos[0, ncand) ascending.  On success blk_off[i] = output offset of block i if
// it lies on the chain (else 0xFFFFFFFF), blk_prev[i] = the byte before it (-1 at the stream start), and the
// stored blocks met on the way are listed.  A fixed-Huffman block (zlib picks one now and then for a short or
// nearly incompressible stretch) has no header to search for, and a dynamic block with BFINAL = 1 (at most the
// last one of a stream; the reference never finishes its streams) is not searched for: the walk measures those on
// the spot and appends them to the block list (slots [ncand, *nblocks)).  Returns 0, or < 0 when the stream must take the serial path.
// Reads up to 3 bytes past `in + in_len` (callers provide that slack, as everywhere in this library).
template <class Tab>
FZ_HD int fz_chain_resolve(const uint8_t *in, uint32_t in_len, uint32_t n_out, uint32_t *cand_pos, FzBlockInfo *info,
                           uint32_t ncand, uint32_t cap, uint32_t *nblocks, uint32_t *blk_off, int *blk_prev,
                           FzStoredItem *stored, uint32_t stored_cap, uint32_t *nstored, const Tab &tab)
{
    for (uint32_t i = 0; i < ncand; i++) blk_off[i] = 0xFFFFFFFFu;
    *nstored = 0;
    *nblocks = ncand;
    uint64_t pos = 0;
    uint32_t out = 0;
    int prev = -1;
    const uint64_t total_bits = (uint64_t)in_len * 8;
    for (uint32_t guard = 0; guard < 100000; guard++) {
        if (total_bits - pos < 3) break;
        if (out == n_out && total_bits - pos < 8) break;
        const uint32_t h = fz_stream_bits(in, pos, 3);
        const uint32_t type = h >> 1;
        if (type == 0) {  // stored block: skip to the byte boundary, LEN, NLEN, data
            const uint64_t by = (pos + 3 + 7) >> 3;
            if (by + 4 > in_len) return -1;
            const uint32_t len = (uint32_t)in[by] | ((uint32_t)in[by + 1] << 8);
            const uint32_t nlen = (uint32_t)in[by + 2] | ((uint32_t)in[by + 3] << 8);
            if ((len ^ 0xFFFFu) != nlen || by + 4 + len > in_len || out + len > n_out) return -2;
            if (len) {
                if (*nstored >= stored_cap) return -3;
                stored[*nstored].src_byte = (uint32_t)by + 4;
                stored[*nstored].len = len;
                stored[*nstored].out_off = out;
                (*nstored)++;
                prev = in[by + 4 + len - 1];
                out += len;
            }
            pos = (by + 4 + len) * 8;
        } else if (type == 1 || type == 2) {
            uint32_t k;
            uint32_t lo = 0, hi = ncand;
            if (type == 2 && !(h & 1u)) {  // the candidate starting exactly here
                while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (cand_pos[mid] < pos) lo = mid + 1; else hi = mid; }
                if (lo >= ncand || cand_pos[lo] != pos) return -4;
                k = lo;
            } else {   // fixed code, or a final dynamic block (the search only looks for BFINAL = 0): measured here
                if (*nblocks >= cap) return -8;
                k = (*nblocks)++;
                cand_pos[k] = (uint32_t)pos;
                fz_block_measure(in, (size_t)in_len, pos, tab, (uint32_t *)nullptr, &info[k]);
            }
            const FzBlockInfo &b = info[k];
            if (!(b.flags & FZ_BLK_OK) || (b.flags & FZ_BLK_NON_RLE)) return -5;
            if ((b.flags & FZ_BLK_STARTS_WITH_MATCH) && prev < 0) return -6;
            if (out + b.out_len > n_out || b.end_bit <= pos) return -7;
            blk_off[k] = out;
            blk_prev[k] = prev;
            out += b.out_len;
            if (b.flags & FZ_BLK_HAS_LITERAL) prev = (int)b.last;
            pos = b.end_bit;
        } else return -9;  // invalid block type
        if (h & 1u) break;  // BFINAL
    }
    return out == n_out ? 0 : -10;
}

// ---------------------------------------------------------------------------------------------------
// One block, one WARP: self-synchronising decode.
//
// A Huffman bit stream decoded from a wrong bit position falls back into step with the true symbol boundaries
// after a few symbols.  The block body is cut into tiles of 32 sub-ranges of FZ_BP_SUB_BITS bits; every lane
// decodes its sub-range speculatively from the grid position, then lanes re-decode from the position where
// their predecessor actually stopped until nothing moves any more (lane 0 of a tile starts at a known symbol
// boundary, so the fixed point is the true parse).  Byte counts are prefix-summed over the lanes; the write
// variant then decodes every sub-range once more, now storing at its offset.  All matches must have
// distance 1 (Z_RLE): a sub-range that starts inside a run only needs the last literal before it.
// SPMD style as in fz_deflate_enc.cuh: phases communicate through FzSyncState (shared memory on the GPU).
// ---------------------------------------------------------------------------------------------------
#include "fz_deflate_enc.cuh"   // FZ_PHASE

#if defined(__CUDA_ARCH__)
#define FZ_WARP_ANY(pred) (__any_sync(0xffffffffu, (pred)) != 0)
#else
#define FZ_WARP_ANY(pred) (pred)
#endif

// the big phase bodies are real functions on the device: one copy each, and their registers do not add up
#if defined(__CUDACC__)
#define FZ_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define FZ_HD_NOINLINE inline
#endif

#define FZ_BP_SUB_BITS 2048u         // bits of a sub-range (one lane's share of a tile) when nothing is known about the block ...
#define FZ_BP_SUB_MIN 768u           // ... and the least it is cut down to when the block is known to be short (fz_sy_block, len_hint)
#define FZ_BP_PROBE_BITS 320u       // first probe length of a block ...
#define FZ_BP_PROBE_MAX 1280u       // ... doubled after every tile that needed a redo round, up to this
#define FZ_SY_EOB 1u        // the sub-range ended with the end-of-block symbol
#define FZ_SY_ERR 2u        // decode error (normal for a speculative start; fatal once the parse is settled)
#define FZ_SY_NONRLE 8u
#define FZ_SY_SWM 16u       // starts with a match that reaches before the sub-range
#define FZ_SY_REDO 32u
#define FZ_SY_WERR 256u

#if !defined(__CUDA_ARCH__) && defined(FZ_SY_STATS)
static uint64_t fz_sy_stat_tiles, fz_sy_stat_rounds, fz_sy_stat_redos, fz_sy_stat_redo_lanes;
#endif

// What the measure pass learned about one tile, kept for the write pass (which then decodes every sub-range
// exactly once): where each sub-range really starts, where its bytes go, the byte before it.
struct FzTileRec {
    uint32_t next;        // index of the block's next tile record, FZ_TILE_NONE at the end
    uint32_t last_lane;   // sub-ranges 0 .. last_lane belong to the block
    uint32_t has_eob;     // the last one ends with the end-of-block symbol
    uint32_t end_bit;     // where the last sub-range stopped
    uint32_t total;       // bytes of the tile
    uint32_t pad[3];
    uint32_t start[32];
    uint32_t off[32];     // byte offset from the start of the block
    uint16_t prev[32];    // byte before the sub-range + 1, 0 = the byte before the block (or nothing)
};
#define FZ_TILE_NONE 0xFFFFFFFFu

struct FzTilePool {
    FzTileRec *recs;
    uint32_t cap;
    uint32_t *cursor;     // bump allocator (atomic on the device)
};

struct FzSyncState {
    uint32_t lut[FZ_LUT_SIZE];
    uint16_t tab[FZ_INF_TAB_U16];
    FzCode LL, DD;
    uint32_t start[32], end[32], n[32], flags[32], lastc[32], want[32], off[32];
    uint32_t rend[32];   // the sub-range of a lane ends at the first symbol boundary >= rend[lane]
    int prev[32];
    // written by lane 0 in dedicated phases, read by everybody afterwards
    uint32_t hdr_ok, hdr_end, is_last, dd1;
    uint32_t any_redo;
    uint32_t tile_total, tile_eob_lane, tile_err, tile_nonrle, tile_werr;
    int tile_carry;
    uint32_t rec_idx, rec_prev;   // tile record being written / the one before it
    uint32_t probe_bits;          // length of the probe decode; grows when a block's code is slow to fall into step
    uint32_t sub_bits;            // bits per sub-range in this block (FZ_BP_SUB_MIN .. FZ_BP_SUB_BITS, a multiple of 64)
};

FZ_HD uint32_t fz_tile_alloc(const FzTilePool &pool)
{
#if defined(__CUDA_ARCH__)
    const uint32_t i = atomicAdd(pool.cursor, 1u);
#else
    const uint32_t i = (*pool.cursor)++;
#endif
    return i < pool.cap ? i : FZ_TILE_NONE;
}

FZ_HD_NOINLINE void fz_sy_ph_header(FzSyncState *st, const uint8_t *in, uint32_t in_len, uint32_t bit, int lane)
{
    if (lane != 0) return;
    typedef FzInfTab<1> Tab;
    const Tab tab{st->tab, st->tab + 288, st->tab + 320};
    FzInflater<Tab> inf;
    inf.start_at_bit(in, (size_t)in_len, (uint64_t)bit, nullptr, 0xFFFFFFF0u, tab);
    inf.bind_codes(&st->LL, &st->DD);   // the header parse leaves the codes where every lane reads them
    inf.bw.dry = true;
    const bool ok = inf.block_header() && inf.in_body && inf.rc == FZ_INF_OK;
    st->hdr_ok = ok ? 1u : 0u;
    st->hdr_end = (uint32_t)(((uint64_t)bit & ~7ull) + inf.consumed_bits());
    st->is_last = inf.last ? 1u : 0u;
    st->dd1 = inf.dd1;
}

// the block's lookup table, in two kinds of phases (fz_inflate.cuh): every lane fills its share of single-symbol entries,
// then the entries are packed with further literals from the top down, 32 at a time
FZ_HD void fz_sy_ph_lut_fill(FzSyncState *st, int lane)
{
    const FzInfTab<1> tab{st->tab, st->tab + 288, st->tab + 320};
    fz_lut_fill_lane<FZ_LUT_BITS>(st->lut, st->LL, tab, lane);
}
FZ_HD void fz_sy_ph_lut_pack_put(FzSyncState *st, uint32_t batch, int lane)   // batch >= 1: reads entries below the batch only
{
    const uint32_t e = batch * 32u + (uint32_t)lane;
    st->lut[e] = fz_lut_pack<FZ_LUT_BITS>(st->lut, e);
}
FZ_HD void fz_sy_ph_lut_pack0(FzSyncState *st, int lane)   // batch 0 reads entries of its own batch: compute, then store
{
    st->want[lane] = fz_lut_pack<FZ_LUT_BITS>(st->lut, (uint32_t)lane);   // (want[] is free here)
}
FZ_HD void fz_sy_ph_lut_put0(FzSyncState *st, int lane)
{
    st->lut[lane] = st->want[lane];
}
#define FZ_SY_BUILD_LUT(st, lane)                                                          \
    do {                                                                                    \
        FZ_PHASE(fz_sy_ph_lut_fill(st, lane));                                              \
        for (uint32_t b_ = FZ_LUT_SIZE / 32u - 1u; b_ >= 1u; b_--) FZ_PHASE(fz_sy_ph_lut_pack_put(st, b_, lane)); \
        FZ_PHASE(fz_sy_ph_lut_pack0(st, lane));                                             \
        FZ_PHASE(fz_sy_ph_lut_put0(st, lane));                                              \
    } while (0)

// decode the sub-range of `lane` in tile `tile_pos` from st->start[lane]; out == nullptr: count only
// probe = true: decode FZ_BP_PROBE_BITS from the grid position of the lane and leave the boundary reached there in
// st->start[lane] -- by then the parse has almost surely fallen into step with the true one, so the counting pass
// that follows starts most lanes on a true symbol boundary and only the unlucky ones have to be redone.
FZ_HD_NOINLINE void fz_sy_decode(FzSyncState *st, const uint8_t *in, uint32_t in_len, uint32_t tile_pos, bool active, uint8_t *out, int lane,
                                 bool probe = false)
{
    typedef FzInfTab<1> Tab;
    const Tab tab{st->tab, st->tab + 288, st->tab + 320};
    const uint32_t grid = tile_pos + (uint32_t)lane * st->sub_bits;
    const uint32_t start = probe ? grid : st->start[lane];
    const uint64_t range_end = probe ? (uint64_t)grid + st->probe_bits : (uint64_t)st->rend[lane];
    const bool write = out != nullptr;
    FzInflater<Tab> inf;
    // counting: the sub-range is [start, first symbol boundary >= range_end); writing: the bytes (and the
    // end-of-block symbol) the counting pass recorded for it
    bool live = active && (uint64_t)start < (uint64_t)in_len * 8 &&
                (write ? (st->n[lane] != 0 || (st->flags[lane] & FZ_SY_EOB) != 0) : (uint64_t)start < range_end);
    const bool ran = live;
    int64_t stop_left = 0;   // bits_left() value at which the sub-range is finished
    if (live) {
        inf.start_at_bit(in, (size_t)in_len, (uint64_t)start, out, write ? st->n[lane] : 0xFFFFFFF0u, tab);
        inf.bw.dry = !write;
        inf.bw.prev_byte = write ? st->prev[lane] : 0;
        inf.shared_tab = true;
        inf.in_body = true;
        inf.one_block = true;
        inf.last = st->is_last != 0;
        inf.bind_codes(&st->LL, &st->DD);
        inf.dd1 = st->dd1;
        stop_left = (int64_t)in_len * 8 - (int64_t)range_end;
    }
    const uint32_t *lut = st->lut;
    const uint32_t run_bit = fz_dd1_run_bit(st->dd1);
    // Lock-step drive as in the sub-block inflater: a register-only run of up to 16 table hits (1-3 literals each),
    // then one general step (long code, match, end of block), then the lanes re-vote.
    if (write) {
        // the counting pass may have taken up to three literals in one step across the end of the sub-range:
        // stop by the byte count it recorded, not by position (plus the end-of-block symbol if it saw one)
        const uint32_t n_goal = st->n[lane];
        const bool want_eob = (st->flags[lane] & FZ_SY_EOB) != 0;
        while (FZ_WARP_ANY(live)) {
            if (live) {
#pragma unroll 1
                for (int it = 0; it < 16; ++it) {
                    inf.br.refill();
                    const uint32_t e = lut[(uint32_t)inf.br.acc & (FZ_LUT_SIZE - 1)];
                    const uint32_t cnt = e >> 29;
                    if ((e & 511u) >= 256u) {
                        // a distance-1 match whole: length code, its extra bits, the one distance bit
                        if (!(e & FZ_LUT_MATCH) || run_bit > 1u) break;
                        const uint32_t cl = (e >> 25) & 15u, xb = (e >> 18) & 7u;
                        const uint32_t a = (uint32_t)(inf.br.acc >> cl);
                        const uint32_t len = ((e >> 9) & 511u) + (a & ((1u << xb) - 1u));
                        if (((a >> xb) & 1u) != run_bit || inf.bw.op + len > inf.bw.cap) break;
                        const int pv = inf.bw.produced() ? (int)inf.bw.back(1) : inf.bw.prev_byte;
                        if (pv < 0) break;   // nothing before the fragment to repeat: the general step reports it
                        inf.br.drop((int)(cl + xb + 1u));
                        inf.bw.fill((uint32_t)pv, len);
                        continue;
                    }
                    if (e == 0 || inf.bw.op + cnt > inf.bw.cap) break;
                    inf.br.drop((int)((e >> 25) & 15u));
                    inf.bw.putn((e & 255u) | ((e >> 1) & 0xffff00u), cnt);
                }
                if (inf.bw.produced() >= n_goal && !want_eob) live = false;
                else {
                    live = inf.step_lut(lut);
                    if (live && inf.bw.produced() >= n_goal && !want_eob) live = false;
                }
            }
        }
    } else {
        // The stop position must not depend on where the parse started: two parses that have fallen into step
        // group literals into multi-literal table hits differently, so within reach of the end of the sub-range
        // symbols are taken one at a time -- the sub-range then ends at the FIRST symbol boundary >= range_end.
        const int64_t near_left = stop_left + (int64_t)FZ_LUT_BITS;
        while (FZ_WARP_ANY(live)) {
            if (live) {
#pragma unroll 1
                for (int it = 0; it < 16; ++it) {
                    inf.br.refill();
                    if (inf.br.bits_left() <= near_left) break;
                    const uint32_t e = lut[(uint32_t)inf.br.acc & (FZ_LUT_SIZE - 1)];
                    if ((e & 511u) >= 256u) {
                        if (!(e & FZ_LUT_MATCH) || run_bit > 1u) break;
                        const uint32_t cl = (e >> 25) & 15u, xb = (e >> 18) & 7u;
                        const uint32_t a = (uint32_t)(inf.br.acc >> cl);
                        if (((a >> xb) & 1u) != run_bit) break;
                        inf.br.drop((int)(cl + xb + 1u));
                        if (inf.bw.op == 0) inf.bw.starts_with_match = true;
                        inf.bw.op += ((e >> 9) & 511u) + (a & ((1u << xb) - 1u));
                        continue;
                    }
                    if (e == 0) break;
                    inf.br.drop((int)((e >> 25) & 15u));
                    inf.bw.putn((e & 255u) | ((e >> 1) & 0xffff00u), e >> 29);
                }
                const bool near_end = inf.br.bits_left() <= near_left;
                live = inf.step_lut(near_end ? nullptr : lut);
                // a probe only looks for the true parse: an end-of-block symbol met on the way is (almost surely) part
                // of the not-yet-synchronised garbage -- keep going
                if (probe && !live && inf.saw_eob && inf.rc == FZ_INF_OK) { inf.saw_eob = false; inf.in_body = true; live = true; }
                if (live && inf.br.bits_left() <= stop_left) live = false;
            }
        }
    }
    if (!active) return;
    if (probe) {
        const bool fine = ran && inf.rc == FZ_INF_OK && !inf.saw_eob;
        st->start[lane] = fine ? (uint32_t)((int64_t)in_len * 8 - inf.br.bits_left()) : grid;
        return;
    }
    uint32_t f = st->flags[lane] & ~(FZ_SY_EOB | FZ_SY_ERR | FZ_SY_NONRLE | FZ_SY_SWM);
    if (!ran) {   // predecessor already reached past this sub-range (or the stream ended): nothing of it is ours
        if (!write) { st->end[lane] = start; st->n[lane] = 0; st->lastc[lane] = 0x100; st->flags[lane] = f; }
        return;
    }
    const uint32_t endpos = (uint32_t)((int64_t)in_len * 8 - inf.br.bits_left());
    if (write) {
        inf.bw.finish();
        if (inf.rc != FZ_INF_OK || inf.bw.produced() != st->n[lane] || endpos != st->end[lane] ||
            inf.saw_eob != ((st->flags[lane] & FZ_SY_EOB) != 0))
            st->flags[lane] = st->flags[lane] | FZ_SY_WERR;
        return;
    }
    if (inf.rc != FZ_INF_OK) {
        f |= FZ_SY_ERR;
        st->end[lane] = (uint32_t)(range_end < (uint64_t)in_len * 8 ? range_end : (uint64_t)in_len * 8);
        st->n[lane] = 0;
        st->lastc[lane] = 0x100;
    } else {
        if (inf.saw_eob) f |= FZ_SY_EOB;
        if (inf.bw.non_rle) f |= FZ_SY_NONRLE;
        if (inf.bw.starts_with_match) f |= FZ_SY_SWM;
        st->end[lane] = endpos;
        st->n[lane] = inf.bw.produced();
        st->lastc[lane] = inf.bw.lastc;
    }
    st->flags[lane] = f;
}

FZ_HD void fz_sy_ph_probe(FzSyncState *st, const uint8_t *in, uint32_t in_len, uint32_t tile_pos, int lane)
{
    st->flags[lane] = 0;
    if (lane == 0) st->start[0] = tile_pos;   // the one position known to be a symbol boundary
    fz_sy_decode(st, in, in_len, tile_pos, lane != 0, nullptr, lane, true);
}
FZ_HD void fz_sy_ph_longer_probe(FzSyncState *st, int lane)
{
    if (lane == 0 && st->probe_bits < FZ_BP_PROBE_MAX && st->probe_bits * 2u < st->sub_bits) st->probe_bits *= 2;
}
FZ_HD void fz_sy_ph_ranges(FzSyncState *st, uint32_t tile_pos, int lane)
{
    st->rend[lane] = lane < 31 ? st->start[lane + 1] : tile_pos + 32u * st->sub_bits;
}
FZ_HD void fz_sy_ph_spec(FzSyncState *st, const uint8_t *in, uint32_t in_len, uint32_t tile_pos, int lane)
{
    fz_sy_decode(st, in, in_len, tile_pos, true, nullptr, lane);
}

// A sub-range is settled when it started where its (settled) predecessor stopped; lane 0 always is.  Lanes that
// are not re-decode from their predecessor's current end -- all of them at once.
FZ_HD void fz_sy_ph_look(FzSyncState *st, uint32_t tile_pos, int lane)
{
    uint32_t f = st->flags[lane] & ~FZ_SY_REDO;
    const uint32_t want = lane == 0 ? tile_pos : st->end[lane - 1];
    st->want[lane] = want;
    if (st->start[lane] != want) f |= FZ_SY_REDO;
    st->flags[lane] = f;
}

// lane 0: length of the settled prefix, and whether the block ends inside it (then the rest of the tile is moot:
// what follows the end-of-block symbol is not this block's data, however the speculative lanes parsed it)
FZ_HD void fz_sy_ph_summary(FzSyncState *st, int lane)
{
    if (lane != 0) return;
    uint32_t settled = 32, eob_lane = 32;
    for (int l = 0; l < 32; l++) {
        if (st->flags[l] & FZ_SY_REDO) { settled = (uint32_t)l; break; }
        if (st->flags[l] & (FZ_SY_EOB | FZ_SY_ERR)) { eob_lane = (uint32_t)l; break; }   // an error ends the tile as well
    }
    st->tile_eob_lane = eob_lane;
    st->any_redo = (eob_lane == 32 && settled < 32) ? 1u : 0u;
}

FZ_HD void fz_sy_ph_apply(FzSyncState *st, int lane)
{
    if (st->flags[lane] & FZ_SY_REDO) st->start[lane] = st->want[lane];
}

FZ_HD void fz_sy_ph_redo(FzSyncState *st, const uint8_t *in, uint32_t in_len, uint32_t tile_pos, int lane)
{
    fz_sy_decode(st, in, in_len, tile_pos, (st->flags[lane] & FZ_SY_REDO) != 0, nullptr, lane);
}

// offsets, the byte before every sub-range, totals of the tile (lanes 0 .. last, last = the end-of-block lane or 31)
FZ_HD void fz_sy_ph_scan(FzSyncState *st, int carry, int lane)
{
    if (lane != 0) return;
    uint32_t total = 0, err = 0, nonrle = 0;
    const int last = st->tile_eob_lane < 32u ? (int)st->tile_eob_lane : 31;
    for (int l = 0; l <= last; l++) {
        const uint32_t f = st->flags[l];
        st->off[l] = total;
        st->prev[l] = carry;
        if (f & FZ_SY_ERR) { err = 1; break; }
        if (f & FZ_SY_NONRLE) nonrle = 1;
        total += st->n[l];
        if (st->lastc[l] < 0x100u) carry = (int)st->lastc[l];
    }
    st->tile_total = total;
    st->tile_err = err;
    st->tile_nonrle = nonrle;
    st->tile_carry = carry;
}

FZ_HD void fz_sy_ph_write(FzSyncState *st, const uint8_t *in, uint32_t in_len, uint32_t tile_pos, uint8_t *out_tile, int lane)
{
    const bool active = (uint32_t)lane <= st->tile_eob_lane;
    fz_sy_decode(st, in, in_len, tile_pos, active, out_tile + (active ? st->off[lane] : 0u), lane);
}

FZ_HD void fz_sy_ph_werr(FzSyncState *st, int lane)
{
    if (lane != 0) return;
    uint32_t w = 0;
    for (int l = 0; l < 32; l++) w |= st->flags[l] & FZ_SY_WERR;
    st->tile_werr = w;
}

// measure pass: keep the settled tile for the write pass (lane 0 allocates and links, every lane stores its entry)
FZ_HD void fz_sy_ph_record_alloc(FzSyncState *st, const FzTilePool &pool, int lane)
{
    if (lane != 0) return;
    st->rec_prev = st->rec_idx;
    st->rec_idx = fz_tile_alloc(pool);
}
FZ_HD void fz_sy_ph_record(FzSyncState *st, const FzTilePool &pool, uint32_t out_base, int lane)
{
    if (st->rec_idx == FZ_TILE_NONE) return;
    FzTileRec *r = &pool.recs[st->rec_idx];
    const uint32_t last = st->tile_eob_lane < 32u ? st->tile_eob_lane : 31u;
    r->start[lane] = st->start[lane];
    r->off[lane] = out_base + st->off[lane];
    r->prev[lane] = (uint16_t)(st->prev[lane] + 1);
    if (lane == 0) {
        r->next = FZ_TILE_NONE;
        r->last_lane = last;
        r->has_eob = st->tile_eob_lane < 32u ? 1u : 0u;
        r->end_bit = st->end[last];
        r->total = st->tile_total;
        if (st->rec_prev != FZ_TILE_NONE) pool.recs[st->rec_prev].next = st->rec_idx;
    }
}

// write pass, table driven: load one tile record into the state the decode phase reads
FZ_HD void fz_sy_ph_load(FzSyncState *st, const FzTileRec *r, uint32_t out_len, int block_prev, int lane)
{
    const uint32_t last = r->last_lane;
    uint32_t f = 0;
    if ((uint32_t)lane <= last) {
        const uint32_t off = r->off[lane];
        const uint32_t next_off = (uint32_t)lane < last ? r->off[lane + 1] : r->off[0] + r->total;
        st->start[lane] = r->start[lane];
        st->end[lane] = (uint32_t)lane < last ? r->start[lane + 1] : r->end_bit;
        st->off[lane] = off - r->off[0];
        st->n[lane] = next_off - off;
        st->prev[lane] = r->prev[lane] ? (int)r->prev[lane] - 1 : block_prev;
        if ((uint32_t)lane == last && r->has_eob) f |= FZ_SY_EOB;
        if (next_off < off || next_off > out_len) f |= FZ_SY_WERR;   // corrupt record: refuse
    }
    st->flags[lane] = f;
    if (lane == 0) st->tile_eob_lane = last;   // fz_sy_ph_write decodes lanes 0 .. tile_eob_lane
}

// Write pass over the records the measure pass left: one decode per sub-range, no searching.
FZ_HD void fz_sy_block_from_table(FzSyncState *st, const uint8_t *in, uint32_t in_len, uint32_t bit, uint8_t *out, uint32_t out_len,
                                  int prev_byte, uint32_t expect_end, const FzTilePool &pool, uint32_t first_rec, bool *ok, int lane)
{
    (void)lane;
    FZ_PHASE(fz_sy_ph_header(st, in, in_len, bit, lane));
#if defined(__CUDA_ARCH__)
    if (lane == 0) st->sub_bits = FZ_BP_SUB_BITS;   // (not used when the records say where every sub-range starts)
    __syncwarp();
#else
    st->sub_bits = FZ_BP_SUB_BITS;
#endif
    bool good = st->hdr_ok != 0;
    uint32_t produced = 0, end_bit = bit, rec = first_rec;
    bool done = false;
    if (good) {
        FZ_SY_BUILD_LUT(st, lane);
        for (uint32_t guard = 0; guard < pool.cap && good && !done; guard++) {
            if (rec == FZ_TILE_NONE || rec >= pool.cap) { good = false; break; }
            const FzTileRec *r = &pool.recs[rec];
            if (r->off[0] != produced || r->last_lane > 31u) { good = false; break; }
            FZ_PHASE(fz_sy_ph_load(st, r, out_len, prev_byte, lane));
            // tile_pos is only used for the sub-range grid in counting mode; the write mode stops by byte count
            FZ_PHASE(fz_sy_ph_write(st, in, in_len, 0u, out + produced, lane));
            FZ_PHASE(fz_sy_ph_werr(st, lane));
            if (st->tile_werr) { good = false; break; }
            produced += r->total;
            end_bit = r->end_bit;
            if (r->has_eob) done = true;
            rec = r->next;
#if defined(__CUDA_ARCH__)
            __syncwarp();
#endif
        }
    }
    *ok = good && done && produced == out_len && end_bit == expect_end;
}

// Decode the block whose header starts at `bit`.
//   WRITE = false: fills *bi (as fz_block_measure does);  WRITE = true: stores out[0, out_len), returns success
// in *ok (end position must equal expect_end).  Called by all 32 lanes on the device, once on the host.
template <bool WRITE>
FZ_HD void fz_sy_block(FzSyncState *st, const uint8_t *in, uint32_t in_len, uint32_t bit, uint8_t *out, uint32_t out_len,
                       int prev_byte, uint32_t expect_end, FzBlockInfo *bi, bool *ok, int lane,
                       const FzTilePool *pool = nullptr, uint32_t *first_rec = nullptr, uint32_t len_hint = 0)
{
    (void)lane;
    FZ_PHASE(fz_sy_ph_header(st, in, in_len, bit, lane));
    // len_hint: bits from this header to the next candidate header of the stream (0 = unknown).  zlib closes a block after
    // a fixed number of symbols (16383 or 32767), whatever bits that takes: at 2048 bits per sub-range the last tile of a
    // block had work for some of its 32 lanes only.  The block is cut into equal sub-ranges instead (never shorter than
    // FZ_BP_SUB_MIN: the probe decode in front of every sub-range does not shrink with it).  A hint that is off costs
    // speed, not correctness.
    {
        uint32_t sub = FZ_BP_SUB_BITS;
        const uint32_t hdr = st->hdr_end > bit ? st->hdr_end - bit : 0u;
        if (len_hint > hdr) {
            // as many tiles as sub-ranges of FZ_BP_SUB_BITS would need, all of them full: a block of 1.35 tiles (an exponent
            // plane at zlib's 32767 symbols per block) took two tile times with a third of the lanes idle in the second
            const uint32_t body = len_hint - hdr;
            const uint32_t tiles = (body + 32u * FZ_BP_SUB_BITS - 1u) / (32u * FZ_BP_SUB_BITS);
            sub = (((body + 32u * tiles - 1u) / (32u * tiles)) + 63u) & ~63u;
            sub = sub < FZ_BP_SUB_MIN ? FZ_BP_SUB_MIN : (sub > FZ_BP_SUB_BITS ? FZ_BP_SUB_BITS : sub);
        }
#if defined(__CUDA_ARCH__)
        __syncwarp();
        if (lane == 0) st->sub_bits = sub;
        __syncwarp();
#else
        st->sub_bits = sub;
#endif
    }
    uint32_t flags = 0, produced = 0, end_bit = bit;
    uint32_t rec0 = FZ_TILE_NONE;
    bool table_ok = !WRITE && pool != nullptr;
#if defined(__CUDA_ARCH__)
    if (lane == 0) { st->rec_idx = FZ_TILE_NONE; st->probe_bits = FZ_BP_PROBE_BITS; }
    __syncwarp();
#else
    st->rec_idx = FZ_TILE_NONE;
    st->probe_bits = FZ_BP_PROBE_BITS;
#endif
    int carry = WRITE ? prev_byte : -1;
    bool good = st->hdr_ok != 0, done = false;
    if (good) {
        FZ_SY_BUILD_LUT(st, lane);
        uint32_t tile_pos = st->hdr_end;
        const uint32_t max_tiles = (uint32_t)(((uint64_t)in_len * 8 - tile_pos) / (32u * st->sub_bits)) + 2u;
        for (uint32_t tile = 0; tile < max_tiles && good && !done; tile++) {
            FZ_PHASE(fz_sy_ph_probe(st, in, in_len, tile_pos, lane));
            FZ_PHASE(fz_sy_ph_ranges(st, tile_pos, lane));
            FZ_PHASE(fz_sy_ph_spec(st, in, in_len, tile_pos, lane));
#if !defined(__CUDA_ARCH__) && defined(FZ_SY_STATS)
            fz_sy_stat_tiles++;
#endif
            bool settled = false, redone = false;
            for (int round = 0; round < 34 && !settled; round++) {
                FZ_PHASE(fz_sy_ph_look(st, tile_pos, lane));
                FZ_PHASE(fz_sy_ph_summary(st, lane));
#if !defined(__CUDA_ARCH__) && defined(FZ_SY_STATS)
                fz_sy_stat_rounds++; if (st->any_redo) { fz_sy_stat_redos++; for (int l = 0; l < 32; l++) fz_sy_stat_redo_lanes += (st->flags[l] & FZ_SY_REDO) != 0; }
#endif
                if (st->any_redo) {
                    FZ_PHASE(fz_sy_ph_apply(st, lane));
                    FZ_PHASE(fz_sy_ph_redo(st, in, in_len, tile_pos, lane));
                    redone = true;
                } else settled = true;
            }
            if (!settled) { good = false; break; }
            if (redone) FZ_PHASE(fz_sy_ph_longer_probe(st, lane));
            FZ_PHASE(fz_sy_ph_scan(st, carry, lane));
            const uint32_t total = st->tile_total;
            if (st->tile_err) { good = false; break; }
            if (st->tile_nonrle) flags |= FZ_BLK_NON_RLE;
            if (tile == 0 && (st->flags[0] & FZ_SY_SWM)) flags |= FZ_BLK_STARTS_WITH_MATCH;
            if (WRITE) {
                if ((flags & FZ_BLK_NON_RLE) || (uint64_t)produced + total > out_len) { good = false; break; }
                FZ_PHASE(fz_sy_ph_write(st, in, in_len, tile_pos, out + produced, lane));
                FZ_PHASE(fz_sy_ph_werr(st, lane));
                if (st->tile_werr) { good = false; break; }
            }
            if (table_ok) {
                FZ_PHASE(fz_sy_ph_record_alloc(st, *pool, lane));
                if (st->rec_idx == FZ_TILE_NONE) table_ok = false;   // pool exhausted: the write pass will search again
                else {
                    if (tile == 0) rec0 = st->rec_idx;
                    FZ_PHASE(fz_sy_ph_record(st, *pool, produced, lane));
                }
            }
            produced += total;
            carry = st->tile_carry;
            if (st->tile_eob_lane < 32u) { done = true; end_bit = st->end[st->tile_eob_lane]; }
            else {
                const uint32_t next = st->end[31];
                if (next <= tile_pos) { good = false; break; }
                tile_pos = next;
            }
#if defined(__CUDA_ARCH__)
            __syncwarp();   // everybody has read the tile's summary before the next tile overwrites it
#endif
        }
    }
    good = good && done;
    if (WRITE) {
        *ok = good && produced == out_len && end_bit == expect_end;
    } else {
        if (good) flags |= FZ_BLK_OK;
        if (carry >= 0) flags |= FZ_BLK_HAS_LITERAL;
        if (st->is_last) flags |= FZ_BLK_FINAL;
        bi->end_bit = end_bit;
        bi->out_len = produced;
        bi->flags = flags;
        bi->last = carry >= 0 ? (uint32_t)carry : 0u;
        if (first_rec) *first_rec = (good && table_ok) ? rec0 : FZ_TILE_NONE;
    }
}
