// fz_deflate_enc.cuh -- warp-cooperative deflate encoder for ONE sub-block (<= FZ_SUB bytes) of a byte plane.
//
// Replaces, for one sub-block, what the reference gets from zlib's deflate(Z_RLE, level 6) behind
// mzlib_def (reference zip.c:164-196, constants constant.h:22-24): distance-1 run matches of
// 3..258 bytes, one dynamic-Huffman block, then an empty stored block (the sync-flush marker) so
// the next sub-block starts byte aligned.  BFINAL is never set (reference decoder requirement,
// SURVEY.md 7.2).  The produced bytes are a valid raw-deflate fragment, not zlib's bytes.
//
// One Huffman code serves a GROUP of FZ_GROUP_SUBS consecutive sub-blocks of a stream (512 KiB): the
// histogram of the group is collected first, the code and the block header are built once per group
// (fz_build_group_code), and every sub-block is then emitted as its own dynamic block carrying that
// same header.  The GPU inflater exploits this: the 32 lanes of a warp decode the 32 sub-blocks of a
// group with ONE shared lookup table (it verifies the headers really are identical).
//
// SPMD style: `lane` is 0..31; a phase may only communicate through FzEncState (shared memory on the
// GPU).  FZ_PHASE(x) runs x for this lane and then __syncwarp() on the device; on the host
// (tests/hostmodel) it loops the 32 lanes sequentially -- so the same source is checked on the CPU.
#pragma once
#include "fz_common.cuh"

#if defined(__CUDA_ARCH__)
#define FZ_PHASE(stmt) do { stmt; __syncwarp(); } while (0)
#define FZ_LANE_DECL
#else
#define FZ_PHASE(stmt) do { for (int lane = 0; lane < 32; ++lane) { stmt; } } while (0)
#endif

// true if `pred` holds on every lane of the warp that is executing this line (the host model has one lane at a time)
#if defined(__CUDA_ARCH__)
#define FZ_WARP_ALL(pred) (__all_sync(__activemask(), (pred)) != 0)
#else
#define FZ_WARP_ALL(pred) (pred)
#endif

struct FzVec16 { uint32_t w[4]; };

// Per-warp encoder state.  ~7.3 KB; lives in shared memory on the GPU.
struct FzEncState {
    uint32_t hist[288];      // literal/length frequencies (286 used)
    uint32_t nmatch;         // number of matches (all use distance code 0)
    uint32_t n_active;       // literal/length symbols with freq > 0
    uint32_t hlit;           // number of literal/length code lengths sent (>= 257)
    uint32_t ncl;            // number of code-length code lengths sent (>= 4)
    uint32_t ntok;           // code-length RLE tokens
    uint32_t hdr_nbits;      // bits in hdr[] (block header incl. the 3 type bits)
    uint32_t dyn_bits;       // exact size of the dynamic block incl. header and EOB
    uint32_t pad0;
    uint32_t keys[512];      // sort workspace; then Moffat-Katajainen array
    uint16_t ssym[320];      // symbols in ascending frequency order
    uint16_t code[288];      // bit-reversed canonical codes
    uint8_t len[288];        // code lengths
    uint8_t seq[320];        // hlit literal/length lengths followed by the distance lengths
    uint16_t cltok[320];     // code-length tokens: sym | extra << 5
    uint32_t hdr[160];       // bit-packed block header
    uint32_t lane_cnt[32];   // scratch: per-lane counts
    uint32_t lane_bits[32];  // bits each lane will emit
    uint32_t num_codes[32];  // symbols per code length (1..15), litlen
    uint32_t next_code[16];
    uint16_t rank[32 * 16];  // per-lane per-length symbol counts -> exclusive prefix
    uint32_t clfreq[19];
    uint8_t cllen[19];
    uint8_t clcode[19];
    uint8_t pad1[2];
    uint32_t nzmask[10];
    // bit-merge of lane outputs
    uint32_t fw_idx[32], fw_bits[32], tw_bits[32], crossed[32];
};

// -------------------------------------------------------------------------------------------------
// Run tokeniser (distance-1 matches only, the reference's Z_RLE strategy).  A byte equal to its
// predecessor is a "repeat".  The first FZ_HOLD_AFTER repeats of a run are emitted as literals right
// away; from the next repeat on bytes are withheld and leave as one match (3..258) when the run ends
// or 258 are pending -- or as 1-2 literals if fewer than 3 were withheld.  (zlib withholds from the
// first repeat.  Short runs of a frequent symbol cost about the same either way; starting late keeps
// "a run starts inside this 16-byte group" rare -- with 32 lanes scanning 32 pieces in lock step that
// event is what serialises the warp: at FZ_HOLD_AFTER = 2 it hit 13 % of the groups of a float
// exponent plane, i.e. nearly every warp iteration; at 6 it is 0.02 %.)
// `prev_init` < 0 means "no byte before `begin`" (sub-block start: sub-blocks never reference earlier data).
// -------------------------------------------------------------------------------------------------
#define FZ_HOLD_AFTER 6

FZ_HD int fz_ctz32(uint32_t v)  // v != 0
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}

FZ_HD uint32_t fz_byte_dyn(const FzVec16 &v, uint32_t k)  // byte k (0..15) of a register-resident group
{
    const uint32_t w = k < 8 ? (k < 4 ? v.w[0] : v.w[1]) : (k < 12 ? v.w[2] : v.w[3]);
    return (w >> ((k & 3) * 8)) & 0xffu;
}

// number of consecutive bits equal to bit 0 of x, starting at bit 0 (x != 0 and x != ~0 is not required: capped by `width`)
FZ_HD uint32_t fz_run_len(uint32_t x, uint32_t width)
{
    const uint32_t y = (x & 1u) ? ~x : x;  // now the run is a run of zeros
    const uint32_t r = y ? (uint32_t)fz_ctz32(y) : 32u;
    return r < width ? r : width;
}

// The tokeniser's state between 16-byte groups (so that a caller can feed a piece window by window).
struct FzScan {
    int prev;        // last byte seen, -1 = none
    uint32_t rep;    // repeats of `prev` seen so far in this run (saturates at FZ_HOLD_AFTER)
    uint32_t m;      // withheld bytes
    FZ_HD void init(int prev_init) { prev = prev_init; rep = 0; m = 0; }

    // Whole 16-byte group at once.  Bit k+H of eq = byte k equals its predecessor (H = FZ_HOLD_AFTER); the
    // H bits below stand for the bytes before the group (from `rep`).  Byte k is withheld iff the
    // H+1 flags ending at k are all set.
    template <class Sink>
    FZ_HD void group16(const FzVec16 &v, Sink &sink)
    {
        // four bytes per step: XOR each word with itself shifted up one byte (the predecessor of byte 0 comes from
        // the word before), find the zero bytes exactly, gather their flags with one multiply
        uint32_t eq = 0;
        uint32_t before = (uint32_t)prev << 24;   // prev = -1 (no predecessor) gives 0xFF...: compared as 0xFF, fixed below
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t w = v.w[j];
            const uint32_t x = w ^ ((w << 8) | (before >> 24));
            const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);   // 0x80 in every zero byte of x
            eq |= ((z * 0x00204081u) >> 28) << (4 * j + FZ_HOLD_AFTER);
            before = w;
        }
        if (prev < 0) eq &= ~(1u << FZ_HOLD_AFTER);   // the first byte of a piece has no predecessor
        const int p = (int)(v.w[3] >> 24);
        const uint32_t hist_bits = ((1u << rep) - 1u) << (FZ_HOLD_AFTER - rep);  // the last `rep` flags before the group
        const uint32_t ext = eq | hist_bits;
        uint32_t held = ext;
#pragma unroll
        for (int j = 1; j <= FZ_HOLD_AFTER; j++) held &= ext << j;
        held >>= FZ_HOLD_AFTER;  // 16 bits: byte k of the group is withheld
        if (!Sink::kOrdered) {
            // token order is irrelevant (histogram, bit count): every byte that is not withheld is a literal,
            // all lanes run the same code whatever their data; runs are book-kept below (rare)
            // (when no lane of the warp withholds anything the unpredicated form is a little cheaper)
            if (FZ_WARP_ALL(held == 0)) sink.literal16(v);
            else if (!FZ_WARP_ALL(held == 0xffffu)) sink.literal_masked(v, ~held & 0xffffu);  // (all-run groups: nothing)
        } else if (held == 0 && m == 0) {
            sink.literal16(v);                 // the common case of the ordered (emitting) pass
        }
        if (held | m) {
            // runs: walk the alternating segments of `held`
            uint32_t pos = 0;
            while (pos < 16) {
                const uint32_t rest = held >> pos;
                const uint32_t r = fz_run_len(rest, 16 - pos);
                if (rest & 1u) {               // withheld bytes
                    m += r;
                    pos += r;
                    if (m >= FZ_MAX_MATCH) { sink.match(FZ_MAX_MATCH); m -= FZ_MAX_MATCH; }
                    if (pos < 16) {            // the run ends inside the group
                        if (m >= FZ_MIN_MATCH) sink.match(m); else if (m) sink.literal(fz_byte_dyn(v, pos - 1), m);
                        m = 0;
                    }
                } else {                       // ordinary bytes
                    if (m) {                   // a run carried over from the previous group ends here (pos == 0)
                        if (m >= FZ_MIN_MATCH) sink.match(m); else sink.literal((uint32_t)prev, m);
                        m = 0;
                    }
                    if (Sink::kOrdered)
                        for (uint32_t k = pos; k < pos + r; k++) sink.literal(fz_byte_dyn(v, k), 1);
                    pos += r;
                }
            }
        }
        // repeats at the end of the group: trailing ones of the flag word, saturated
        const uint32_t inv = ~(ext >> FZ_HOLD_AFTER) & 0xffffu;        // zero flag = run break
        const uint32_t t = inv ? (uint32_t)(15 - fz_ilog2(inv)) : 16u + rep;  // flags set after the last break
        rep = t > FZ_HOLD_AFTER ? FZ_HOLD_AFTER : t;
        prev = p;
    }

    // one byte (ragged tails)
    template <class Sink>
    FZ_HD void byte(int c, Sink &sink)
    {
        const bool e = c == prev;
        if (e && rep >= FZ_HOLD_AFTER) {
            if (++m == FZ_MAX_MATCH) { sink.match(FZ_MAX_MATCH); m = 0; }
        } else {
            if (m) {
                if (m >= FZ_MIN_MATCH) sink.match(m); else sink.literal((uint32_t)prev, m);
                m = 0;
            }
            rep = e ? rep + 1 : 0;
            prev = c;
            sink.literal((uint32_t)c, 1);
        }
    }

    // end of the piece: what is still withheld leaves
    template <class Sink>
    FZ_HD void finish(Sink &sink)
    {
        if (m) {
            if (m >= FZ_MIN_MATCH) sink.match(m); else sink.literal((uint32_t)prev, m);
            m = 0;
        }
    }
};

template <class Load16, class LoadByte, class Sink>
FZ_HD void fz_scan_piece(const Load16 &ld, const LoadByte &lb, uint32_t begin, uint32_t end, int prev_init, Sink &sink)
{
    FzScan sc;
    sc.init(prev_init);
    for (uint32_t i = begin; i < end; i += 16) {
        const uint32_t lim = end - i;
        if (lim >= 16) { sc.group16(ld(i), sink); continue; }
        for (uint32_t k = 0; k < lim; k++) sc.byte((int)lb(i + k), sink);   // ragged tail (< 16 bytes): byte by byte
    }
    sc.finish(sink);
}

// A "piece scan" feeds this lane's piece of a sub-block through the tokeniser into a sink: scan(sink, lane).
// Generic form: lane l owns bytes [l*P, min(n, (l+1)*P)) behind random-access loaders.  (The device's fast path
// streams the pieces through a small shared-memory window instead: FzWindowScan in fz_kernels.cu.)
FZ_HD uint32_t fz_piece_len(uint32_t n);
template <class Load16, class LoadByte>
struct FzPieceScan {
    const Load16 &ld;
    const LoadByte &lb;
    uint32_t n;
    template <class Sink>
    FZ_HD void operator()(Sink &sink, int lane) const
    {
        const uint32_t P = fz_piece_len(n);
        uint32_t b = lane * P, e = b + P;
        if (e > n) e = n;
        if (b < e) fz_scan_piece(ld, lb, b, e, b ? (int)lb(b - 1) : -1, sink);
    }
};

FZ_HD void fz_atomic_add(uint32_t *p, uint32_t v)
{
#if defined(__CUDA_ARCH__)
    atomicAdd(p, v);
#else
    *p += v;
#endif
}

#define FZ_BYTE_OF(v, k) (((v).w[(k) >> 2] >> (((k) & 3) * 8)) & 0xffu)

struct FzHistSink {
    static constexpr bool kOrdered = false;
    uint32_t *hist;  // 288 counters (shared memory on the GPU)
    FZ_HD void literal_masked(const FzVec16 &v, uint32_t mask)
    {
#pragma unroll
        for (int k = 0; k < 16; k++) if ((mask >> k) & 1u) fz_atomic_add(&hist[FZ_BYTE_OF(v, k)], 1);
    }
    FZ_HD void literal(uint32_t c, uint32_t n) { fz_atomic_add(&hist[c], n); }
    FZ_HD void literal16(const FzVec16 &v)
    {
#pragma unroll
        for (int k = 0; k < 16; k++) fz_atomic_add(&hist[FZ_BYTE_OF(v, k)], 1);
    }
    FZ_HD void match(uint32_t len)
    {
        uint32_t lc, eb, ev;
        fz_len_code(len, lc, eb, ev);
        fz_atomic_add(&hist[257 + lc], 1);
    }
};

// cl[sym] = bit-reversed code | code length << 16
struct FzCountSink {
    static constexpr bool kOrdered = false;
    const uint32_t *cl;
    uint32_t bits;
    FZ_HD void literal_masked(const FzVec16 &v, uint32_t mask)
    {
        uint32_t b = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) b += ((mask >> k) & 1u) ? (cl[FZ_BYTE_OF(v, k)] >> 16) : 0u;
        bits += b;
    }
    FZ_HD void literal(uint32_t c, uint32_t n) { bits += n * (cl[c] >> 16); }
    FZ_HD void literal16(const FzVec16 &v)
    {
        uint32_t b = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) b += cl[FZ_BYTE_OF(v, k)] >> 16;
        bits += b;
    }
    FZ_HD void match(uint32_t mlen)
    {
        uint32_t lc, eb, ev;
        fz_len_code(mlen, lc, eb, ev);
        bits += (cl[257 + lc] >> 16) + eb + 1;  // + 1-bit distance code
    }
};

// LSB-first bit writer into 32-bit words.  The first word a lane touches and its last partial
// word are NOT stored: they are returned for the cross-lane merge (several lanes may share a word).
struct FzBitWriter {
    uint32_t *out;       // word-addressed output (4-byte aligned)
    uint64_t acc;
    uint32_t nbits;      // valid bits in acc (< 32 between puts)
    uint32_t widx;       // index of the word acc's low 32 bits go to
    uint32_t first_idx, first_bits;
    bool crossed;
    FZ_HD void init(uint32_t *o, uint32_t bit_off)
    {
        out = o; acc = 0; nbits = bit_off & 31; widx = bit_off >> 5;
        first_idx = widx; first_bits = 0; crossed = false;
    }
    FZ_HD void put(uint32_t v, uint32_t n)  // n <= 32
    {
        acc |= (uint64_t)v << nbits;
        nbits += n;
        if (nbits >= 32) {
            const uint32_t w = (uint32_t)acc;
            if (!crossed) { first_bits = w; crossed = true; } else out[widx] = w;
            widx++; acc >>= 32; nbits -= 32;
        }
    }
    FZ_HD void align_byte() { nbits = (nbits + 7) & ~7u; if (nbits >= 32) put(0, 0); }
    FZ_HD uint32_t bitpos() const { return widx * 32 + nbits; }
};

struct FzEmitSink {
    static constexpr bool kOrdered = true;
    const uint32_t *cl;
    FzBitWriter bw;
    FZ_HD void literal_masked(const FzVec16 &, uint32_t) {}
    FZ_HD void literal(uint32_t c, uint32_t n)
    {
        const uint32_t e = cl[c];
        for (uint32_t i = 0; i < n; i++) bw.put(e & 0xffffu, e >> 16);
    }
    FZ_HD void literal16(const FzVec16 &v)
    {
        // two codes (<= 15 bits each) per append
#pragma unroll
        for (int k = 0; k < 16; k += 2) {
            const uint32_t e0 = cl[FZ_BYTE_OF(v, k)], e1 = cl[FZ_BYTE_OF(v, k + 1)];
            const uint32_t l0 = e0 >> 16;
            bw.put((e0 & 0xffffu) | ((e1 & 0xffffu) << l0), l0 + (e1 >> 16));
        }
    }
    FZ_HD void match(uint32_t mlen)
    {
        uint32_t lc, eb, ev;
        fz_len_code(mlen, lc, eb, ev);
        const uint32_t e = cl[257 + lc];
        bw.put(e & 0xffffu, e >> 16);
        bw.put(ev, eb + 1);  // extra bits, then the 1-bit distance code '0' (distance 1)
    }
};

// -------------------------------------------------------------------------------------------------
// Huffman construction phases (literal/length alphabet).
// -------------------------------------------------------------------------------------------------
#define FZ_SYMS_PER_LANE 9  // 32 * 9 = 288

FZ_HD void fz_ph_zero_len(FzEncState *st, int lane)
{
    for (int i = lane; i < 288; i += 32) st->len[i] = 0;
}

FZ_HD void fz_ph_count_active(FzEncState *st, int lane)
{
    uint32_t c = 0;
    for (int k = 0; k < FZ_SYMS_PER_LANE; k++) {
        const int s = lane * FZ_SYMS_PER_LANE + k;
        c += st->hist[s] != 0 ? 1u : 0u;
    }
    st->lane_cnt[lane] = c;
}

FZ_HD void fz_ph_compact(FzEncState *st, int lane)
{
    uint32_t off = 0, total = 0;
    for (int l = 0; l < 32; l++) { const uint32_t c = st->lane_cnt[l]; if (l < lane) off += c; total += c; }
    for (int k = 0; k < FZ_SYMS_PER_LANE; k++) {
        const int s = lane * FZ_SYMS_PER_LANE + k;
        const uint32_t f = st->hist[s];
        if (f) st->keys[off++] = (f << 9) | (uint32_t)s;
    }
    // pad to the next power of two with +inf keys for the bitonic network
    uint32_t np2 = 32;
    while (np2 < total) np2 <<= 1;
    for (uint32_t i = total + lane; i < np2; i += 32) st->keys[i] = 0xFFFFFFFFu;
    if (lane == 0) st->n_active = total;
}

// one compare-exchange layer (k, j) of the bitonic network over np2 keys
FZ_HD void fz_ph_bitonic(uint32_t *keys, uint32_t np2, uint32_t k, uint32_t j, int lane)
{
    for (uint32_t t = lane; t < np2 / 2; t += 32) {
        const uint32_t i = 2 * t - (t & (j - 1));  // insert a 0 bit at position log2(j)
        const uint32_t p = i | j;
        const uint32_t a = keys[i], b = keys[p];
        const bool up = (i & k) == 0;
        if ((a > b) == up) { keys[i] = b; keys[p] = a; }
    }
}

// serial (lane 0): in-place minimum-redundancy code lengths (Moffat & Katajainen 1995) over the
// ascending frequencies, then the length limit.  keys[] holds freq << 9 | sym on entry.
FZ_HD void fz_ph_lengths(uint32_t *A, uint16_t *ssym, uint32_t *num_codes, int n, int maxbits, int lane)
{
    if (lane != 0) return;
    for (int i = 0; i < n; i++) { const uint32_t k = A[i]; ssym[i] = (uint16_t)(k & 511u); A[i] = k >> 9; }
    for (int l = 0; l < 32; l++) num_codes[l] = 0;
    if (n == 1) { A[0] = 1; num_codes[1] = 1; return; }
    // phase 1: internal node weights + parent pointers
    A[0] += A[1];
    int root = 0, leaf = 2, next;
    for (next = 1; next < n - 1; next++) {
        if (leaf >= n || A[root] < A[leaf]) { A[next] = A[root]; A[root++] = (uint32_t)next; } else A[next] = A[leaf++];
        if (leaf >= n || (root < next && A[root] < A[leaf])) { A[next] += A[root]; A[root++] = (uint32_t)next; } else A[next] += A[leaf++];
    }
    // phase 2: internal node depths
    A[n - 2] = 0;
    for (next = n - 3; next >= 0; next--) A[next] = A[A[next]] + 1;
    // phase 3: leaf depths
    int avbl = 1, used = 0, dpth = 0;
    root = n - 2; next = n - 1;
    while (avbl > 0) {
        while (root >= 0 && (int)A[root] == dpth) { used++; root--; }
        while (avbl > used) { A[next--] = (uint32_t)dpth; avbl--; }
        avbl = 2 * used; dpth++; used = 0;
    }
    // histogram of lengths, clamped at 31
    for (int i = 0; i < n; i++) { uint32_t l = A[i]; if (l > 31) l = 31; num_codes[l]++; }
    // enforce maxbits: fold the overflow into maxbits, then repair the Kraft sum
    uint32_t over = 0;
    for (int l = maxbits + 1; l < 32; l++) { over += num_codes[l]; num_codes[l] = 0; }
    if (over) {
        num_codes[maxbits] += over;
        uint32_t total = 0;
        for (int l = maxbits; l > 0; l--) total += num_codes[l] << (maxbits - l);
        while (total != (1u << maxbits)) {
            num_codes[maxbits]--;
            for (int l = maxbits - 1; l > 0; l--)
                if (num_codes[l]) { num_codes[l]--; num_codes[l + 1] += 2; break; }
            total--;
        }
    }
    // lengths by ascending frequency: longest first
    int i = 0;
    for (int l = maxbits; l > 0; l--)
        for (uint32_t c = 0; c < num_codes[l]; c++) A[i++] = (uint32_t)l;
}

FZ_HD void fz_ph_scatter_len(FzEncState *st, int lane)
{
    const int n = (int)st->n_active;
    for (int i = lane; i < n; i += 32) st->len[st->ssym[i]] = (uint8_t)st->keys[i];
    if (lane == 0) {
        uint32_t code = 0;
        st->next_code[0] = 0;
        for (int l = 1; l <= 15; l++) { code = (code + st->num_codes[l - 1]) << 1; st->next_code[l] = code; }
        // num_codes[0] is 0 here (only active symbols were counted)
    }
}

FZ_HD void fz_ph_rank_count(FzEncState *st, int lane)
{
    uint16_t *r = &st->rank[lane * 16];
    for (int l = 0; l < 16; l++) r[l] = 0;
    for (int k = 0; k < FZ_SYMS_PER_LANE; k++) r[st->len[lane * FZ_SYMS_PER_LANE + k]]++;
}

FZ_HD void fz_ph_rank_scan(FzEncState *st, int lane)
{
    if (lane >= 16) return;  // lane = code length
    uint32_t run = 0;
    for (int l = 0; l < 32; l++) { const uint16_t c = st->rank[l * 16 + lane]; st->rank[l * 16 + lane] = (uint16_t)run; run += c; }
}

FZ_HD void fz_ph_assign_codes(FzEncState *st, int lane)
{
    uint16_t *r = &st->rank[lane * 16];
    uint32_t hi = 0;
    for (int k = 0; k < FZ_SYMS_PER_LANE; k++) {
        const int s = lane * FZ_SYMS_PER_LANE + k;
        const uint32_t l = st->len[s];
        if (l) {
            st->code[s] = (uint16_t)fz_bitrev(st->next_code[l] + r[l], (int)l);
            r[l]++;
            hi = (uint32_t)s + 1;
        } else st->code[s] = 0;
    }
    st->lane_cnt[lane] = hi;  // highest used symbol + 1 in this lane's range
}

// hlit, the code-length sequence and its non-zero bitmap
FZ_HD void fz_ph_seq(FzEncState *st, int lane)
{
    uint32_t hlit = 257;
    for (int l = 0; l < 32; l++) if (st->lane_cnt[l] > hlit) hlit = st->lane_cnt[l];
    const uint32_t total = hlit + 2;  // two distance codes of one bit each (zlib also always sends >= 2)
    for (uint32_t i = lane; i < 320; i += 32) st->seq[i] = i < hlit ? st->len[i] : (i < total ? 1 : 0);
    if (lane == 0) st->hlit = hlit;
}

FZ_HD void fz_ph_nzmask(FzEncState *st, int lane)
{
    if (lane >= 10) return;
    uint32_t m = 0;
    for (int b = 0; b < 32; b++) if (st->seq[lane * 32 + b]) m |= 1u << b;
    st->nzmask[lane] = m;
}

FZ_HD int fz_ctz(uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}

// serial (lane 0): run-length code the length sequence with the 16/17/18 repeat codes (RFC 1951 3.2.7)
FZ_HD void fz_ph_cl_tokens(FzEncState *st, int lane)
{
    if (lane != 0) return;
    for (int i = 0; i < 19; i++) st->clfreq[i] = 0;
    const uint32_t total = st->hlit + 2;
    uint32_t i = 0, nt = 0;
    while (i < total) {
        const uint32_t v = st->seq[i];
        uint32_t j;
        if (v == 0) {
            // next non-zero entry at or after i (the sequence ends with the two distance lengths, non-zero)
            uint32_t w = i >> 5;
            uint32_t m = st->nzmask[w] & (0xFFFFFFFFu << (i & 31));
            while (m == 0) m = st->nzmask[++w];
            j = w * 32 + (uint32_t)fz_ctz(m);
            uint32_t r = j - i;
            while (r >= 11) { const uint32_t t = r < 138 ? r : 138; st->cltok[nt++] = (uint16_t)(18 | ((t - 11) << 5)); st->clfreq[18]++; r -= t; }
            if (r >= 3) { st->cltok[nt++] = (uint16_t)(17 | ((r - 3) << 5)); st->clfreq[17]++; r = 0; }
            while (r--) { st->cltok[nt++] = 0; st->clfreq[0]++; }
        } else {
            j = i + 1;
            while (j < total && st->seq[j] == v) j++;
            uint32_t r = j - i - 1;
            st->cltok[nt++] = (uint16_t)v; st->clfreq[v]++;
            while (r >= 3) { const uint32_t t = r < 6 ? r : 6; st->cltok[nt++] = (uint16_t)(16 | ((t - 3) << 5)); st->clfreq[16]++; r -= t; }
            while (r--) { st->cltok[nt++] = (uint16_t)v; st->clfreq[v]++; }
        }
        i = j;
    }
    st->ntok = nt;
}

// serial (lane 0): Huffman code for the 19 code-length symbols (7-bit limit) and the packed header
FZ_HD void fz_ph_header(FzEncState *st, int lane)
{
    if (lane != 0) return;
    // tiny alphabet: insertion sort of the active symbols by (freq, sym)
    uint32_t *A = st->keys;
    int n = 0;
    for (int s = 0; s < 19; s++) {
        st->cllen[s] = 0;
        if (st->clfreq[s]) {
            const uint32_t key = (st->clfreq[s] << 9) | (uint32_t)s;
            int p = n++;
            while (p > 0 && A[p - 1] > key) { A[p] = A[p - 1]; p--; }
            A[p] = key;
        }
    }
    if (n == 1) {
        // the code-length code must be complete for zlib's inflate: add a second, unused symbol
        const uint32_t s0 = A[0] & 511u;
        const uint32_t dummy = s0 == 0 ? 1u : 0u;
        st->cllen[s0] = 1; st->cllen[dummy] = 1;
    } else {
        uint16_t *ssym = st->ssym;
        uint32_t *nc = st->num_codes;  // litlen counts no longer needed
        fz_ph_lengths(A, ssym, nc, n, 7, 0);
        for (int i = 0; i < n; i++) st->cllen[ssym[i]] = (uint8_t)A[i];
    }
    // canonical codes
    uint32_t blc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, nx[8];
    for (int s = 0; s < 19; s++) blc[st->cllen[s]]++;
    blc[0] = 0;
    uint32_t code = 0;
    nx[0] = 0;
    for (int l = 1; l <= 7; l++) { code = (code + blc[l - 1]) << 1; nx[l] = code; }
    for (int s = 0; s < 19; s++) {
        const uint32_t l = st->cllen[s];
        st->clcode[s] = l ? (uint8_t)fz_bitrev(nx[l]++, (int)l) : 0;
    }
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    uint32_t ncl = 19;
    while (ncl > 4 && st->cllen[order[ncl - 1]] == 0) ncl--;
    st->ncl = ncl;
    // pack: BFINAL=0, BTYPE=10, HLIT, HDIST, HCLEN, 3-bit lengths, RLE tokens
    uint64_t acc = 0;
    uint32_t nb = 0, w = 0;
#define FZ_HPUT(v, n_)                                                          \
    do {                                                                        \
        acc |= (uint64_t)(v) << nb; nb += (n_);                                 \
        if (nb >= 32) { st->hdr[w++] = (uint32_t)acc; acc >>= 32; nb -= 32; }   \
    } while (0)
    FZ_HPUT(0u, 1);
    FZ_HPUT(2u, 2);
    FZ_HPUT(st->hlit - 257, 5);
    FZ_HPUT(1u, 5);  // HDIST = 2 - 1
    FZ_HPUT(ncl - 4, 4);
    for (uint32_t i = 0; i < ncl; i++) FZ_HPUT((uint32_t)st->cllen[order[i]], 3);
    for (uint32_t t = 0; t < st->ntok; t++) {
        const uint32_t tok = st->cltok[t], s = tok & 31u, ex = tok >> 5;
        FZ_HPUT((uint32_t)st->clcode[s], (uint32_t)st->cllen[s]);
        if (s == 16) FZ_HPUT(ex, 2);
        else if (s == 17) FZ_HPUT(ex, 3);
        else if (s == 18) FZ_HPUT(ex, 7);
    }
    st->hdr_nbits = w * 32 + nb;
    st->hdr[w] = (uint32_t)acc;
#undef FZ_HPUT
}

// exact payload bits (without headers) of coding the group's tokens with the code just built
FZ_HD void fz_ph_cost_partial(FzEncState *st, int lane)
{
    uint64_t b = 0;
    for (int k = 0; k < FZ_SYMS_PER_LANE; k++) {
        const int s = lane * FZ_SYMS_PER_LANE + k;
        if (s >= FZ_NUM_LL) break;
        const uint64_t f = st->hist[s];
        b += f * st->len[s];
        if (s > FZ_EOB) b += f * (fz_len_extra_bits((uint32_t)(s - 257)) + 1);
    }
    st->lane_cnt[lane] = (uint32_t)b;          // 64-bit partial sum, low / high halves
    st->lane_bits[lane] = (uint32_t)(b >> 32);
}

// -------------------------------------------------------------------------------------------------
// Piece geometry: lane l owns bytes [l*P, min(n, (l+1)*P)), P a multiple of 16.
// -------------------------------------------------------------------------------------------------
FZ_HD uint32_t fz_piece_len(uint32_t n) { return (((n + 31) / 32) + 15) & ~15u; }

// histogram of one sub-block's tokens into hist[288] (must be zeroed; one warp's shared memory)
template <class Scan>
FZ_HD void fz_ph_hist_sc(uint32_t *hist, const Scan &scan, int lane)
{
    FzHistSink sink{hist};
    scan(sink, lane);
}
template <class Load16, class LoadByte>
FZ_HD void fz_ph_hist(uint32_t *hist, const Load16 &ld, const LoadByte &lb, uint32_t n, int lane)
{
    fz_ph_hist_sc(hist, FzPieceScan<Load16, LoadByte>{ld, lb, n}, lane);
}

// The code of a group as the emit kernel consumes it (global memory, copied to shared memory per warp)
struct FzGroupCode {
    // hot part (copied to shared memory by every emitting warp): FZ_GROUP_CODE_HOT_BYTES
    uint32_t cl[288];     // bit-reversed canonical code | code length << 16
    uint32_t hdr_nbits;
    uint32_t stored;      // 1: coding this group cannot beat stored blocks -- emit every sub-block stored
    uint32_t pad[2];
    // cold part (read by lane 0 only)
    uint32_t hdr[160];    // bit-packed dynamic block header (BFINAL=0, BTYPE=10, lengths)
};
#define FZ_GROUP_CODE_HOT_BYTES (288 * 4 + 16)

// Build the group's Huffman code and block header from its token histogram.
//   st->hist[0..287] = token frequencies of the whole group, st->hist[256] = number of sub-blocks (EOBs)
//   group_bytes = plane bytes in the group, nsub = sub-blocks in the group
template <int DUMMY = 0>
FZ_HD void fz_build_group_code(FzEncState *st, uint32_t group_bytes, uint32_t nsub, FzGroupCode *out, int lane)
{
    (void)lane;
    FZ_PHASE(fz_ph_zero_len(st, lane));
    FZ_PHASE(fz_ph_count_active(st, lane));
    FZ_PHASE(fz_ph_compact(st, lane));
    {
        uint32_t np2 = 32;
        while (np2 < st->n_active) np2 <<= 1;
        for (uint32_t k = 2; k <= np2; k <<= 1)
            for (uint32_t j = k >> 1; j > 0; j >>= 1) FZ_PHASE(fz_ph_bitonic(st->keys, np2, k, j, lane));
    }
    FZ_PHASE(fz_ph_lengths(st->keys, st->ssym, st->num_codes, (int)st->n_active, 15, lane));
    FZ_PHASE(fz_ph_scatter_len(st, lane));
    FZ_PHASE(fz_ph_rank_count(st, lane));
    FZ_PHASE(fz_ph_rank_scan(st, lane));
    FZ_PHASE(fz_ph_assign_codes(st, lane));
    FZ_PHASE(fz_ph_seq(st, lane));
    FZ_PHASE(fz_ph_nzmask(st, lane));
    FZ_PHASE(fz_ph_cl_tokens(st, lane));
    FZ_PHASE(fz_ph_header(st, lane));
    FZ_PHASE(fz_ph_cost_partial(st, lane));
    // group decision: all sub-blocks dynamic vs all stored (the per-sub-block decision is exact, in fz_emit_subblock)
    uint64_t bits = (uint64_t)nsub * (st->hdr_nbits + 3 + 4 + 32);  // header + empty stored block (avg pad 4) per sub-block
    for (int l = 0; l < 32; l++) bits += ((uint64_t)st->lane_bits[l] << 32) | st->lane_cnt[l];
    const uint64_t stored_bits = 8ull * ((uint64_t)group_bytes + (uint64_t)FZ_STORED_OVERHEAD * nsub);
    const uint32_t stored = bits + (stored_bits >> FZ_MIN_GAIN_SHIFT) >= stored_bits ? 1u : 0u;
#if defined(__CUDA_ARCH__)
    for (int i = lane; i < 288; i += 32) out->cl[i] = (uint32_t)st->code[i] | ((uint32_t)st->len[i] << 16);
    for (int i = lane; i < 160; i += 32) out->hdr[i] = st->hdr[i];
    if (lane == 0) { out->hdr_nbits = st->hdr_nbits; out->stored = stored; }
    __syncwarp();
#else
    for (int i = 0; i < 288; i++) out->cl[i] = (uint32_t)st->code[i] | ((uint32_t)st->len[i] << 16);
    for (int i = 0; i < 160; i++) out->hdr[i] = st->hdr[i];
    out->hdr_nbits = st->hdr_nbits; out->stored = stored;
#endif
}

// per-warp scratch of the emit stage
struct FzEmitState {
    uint32_t lane_bits[32];
    uint32_t fw_idx[32], fw_bits[32], tw_bits[32], crossed[32];
    uint32_t total_bits;
    uint32_t false_marker;  // the fragment contains 00 00 FF FF before its final four bytes
    uint32_t pad[2];
};

template <class Scan>
FZ_HD void fz_ph_count(const FzGroupCode *gc, FzEmitState *es, const Scan &scan, int lane)
{
    FzCountSink sink{gc->cl, 0};
    scan(sink, lane);
    if (lane == 0) sink.bits += gc->hdr_nbits;
    es->lane_bits[lane] = sink.bits;
}

// emit this lane's tokens at its bit offset; lane 0 prepends the block header, lane 31 appends
// EOB + the empty stored block (000, pad to byte, 00 00 FF FF)
template <class Scan>
FZ_HD void fz_ph_emit(const FzGroupCode *gc, const uint32_t *hdr, FzEmitState *es, const Scan &scan, uint32_t *out, int lane)
{
    uint32_t off = 0;
    for (int l = 0; l < lane; l++) off += es->lane_bits[l];
    FzEmitSink sink;
    sink.cl = gc->cl;
    sink.bw.init(out, off);
    if (lane == 0) {
        uint32_t nb = gc->hdr_nbits, w = 0;
        while (nb >= 32) { sink.bw.put(hdr[w++], 32); nb -= 32; }
        if (nb) sink.bw.put(hdr[w] & ((1u << nb) - 1), nb);
    }
    scan(sink, lane);
    if (lane == 31) {
        sink.bw.put(gc->cl[FZ_EOB] & 0xffffu, gc->cl[FZ_EOB] >> 16);
        sink.bw.put(0, 3);
        sink.bw.align_byte();
        sink.bw.put(0x0000u, 16);
        sink.bw.put(0xFFFFu, 16);
        es->total_bits = sink.bw.bitpos();  // total bits of the sub-block fragment (byte aligned)
    }
    es->fw_idx[lane] = sink.bw.first_idx;
    es->crossed[lane] = sink.bw.crossed ? 1u : 0u;
    if (sink.bw.crossed) { es->fw_bits[lane] = sink.bw.first_bits; es->tw_bits[lane] = (uint32_t)sink.bw.acc; }
    else { es->fw_bits[lane] = (uint32_t)sink.bw.acc; es->tw_bits[lane] = 0; }
}

// words shared by several lanes: the lane that completes a word ORs in what earlier lanes left there
FZ_HD void fz_ph_merge(FzEmitState *es, uint32_t *out, int lane)
{
    uint32_t carry = 0;
    for (int j = lane - 1; j >= 0; j--) {
        if (es->crossed[j]) { carry |= es->tw_bits[j]; break; }
        carry |= es->fw_bits[j];
    }
    if (es->crossed[lane]) out[es->fw_idx[lane]] = es->fw_bits[lane] | carry;
    if (lane == 31) {
        const uint32_t total_bits = es->total_bits;
        if (total_bits & 31) out[total_bits >> 5] = es->crossed[31] ? es->tw_bits[31] : (es->fw_bits[31] | carry);
    }
}

// Does the emitted fragment show the sync marker anywhere but in its last four bytes?  (About one
// fragment in a million does; the inflater finds sub-blocks by that marker, so such a fragment is
// emitted stored instead.)  out[] was written by this warp; volatile reads keep L1 out of the way.
FZ_HD void fz_ph_check_marker(FzEmitState *es, const uint32_t *out, uint32_t total_bytes, int lane)
{
    const volatile uint32_t *o = (const volatile uint32_t *)out;
    const uint32_t nwords = (total_bytes + 3) / 4;
    bool hit = false;
    for (uint32_t i = lane; i < nwords; i += 32) {
        const uint32_t w0 = o[i];
        const uint32_t w1 = (i + 1 < nwords) ? o[i + 1] : 0u;
        for (uint32_t k = 0; k < 4; k++) {
            const uint32_t v = k ? ((w0 >> (8 * k)) | (w1 << (32 - 8 * k))) : w0;
            if (v == FZ_MARKER_LE && i * 4 + k + 4 < total_bytes) hit = true;  // position + 4 == total_bytes is the real one
        }
    }
    if (hit) es->false_marker = 1;
}

// -------------------------------------------------------------------------------------------------
// Emit one sub-block with its group's code.  Returns the fragment size in bytes, or
// n + FZ_STORED_OVERHEAD with FZ_SIZE_STORED_FLAG set when a stored block is smaller; then nothing is
// written to `out` (the gather kernel synthesises stored blocks from the plane bytes).
// On the device every lane of the warp calls this with its own `lane`; on the host `lane` is unused.
// `out` needs room for FZ_SLOT_STRIDE bytes.
// -------------------------------------------------------------------------------------------------
// `gc` needs only the hot part of FzGroupCode (FZ_GROUP_CODE_HOT_BYTES); `hdr` points at the group's header words.
// `scan` is the piece scan of the sub-block (it runs twice: size, then emission).
template <class Scan>
FZ_HD uint32_t fz_emit_subblock_sc(const FzGroupCode *gc, const uint32_t *hdr, FzEmitState *es, const Scan &scan, uint32_t n,
                                   uint32_t *out, int lane)
{
    (void)lane;
    const uint32_t stored = fz_stored_size(n) | FZ_SIZE_STORED_FLAG;
    if (gc->stored) return stored;
    FZ_PHASE(fz_ph_count(gc, es, scan, lane));
    uint32_t bits = gc->cl[FZ_EOB] >> 16;
    for (int l = 0; l < 32; l++) bits += es->lane_bits[l];
    // dynamic fragment = block bits + 3 (empty stored header) -> byte boundary + 4 marker bytes
    const uint32_t dyn_bytes = (bits + 3 + 7) / 8 + 4;
    if (dyn_bytes + (n >> FZ_MIN_GAIN_SHIFT) >= fz_stored_size(n)) return stored;
    FZ_PHASE(if (lane == 0) es->false_marker = 0; fz_ph_emit(gc, hdr, es, scan, out, lane));
    FZ_PHASE(fz_ph_merge(es, out, lane));
    FZ_PHASE(fz_ph_check_marker(es, out, es->total_bits / 8, lane));
    if (es->false_marker) return stored;
    return es->total_bits / 8;
}

template <class Load16, class LoadByte>
FZ_HD uint32_t fz_emit_subblock(const FzGroupCode *gc, const uint32_t *hdr, FzEmitState *es, const Load16 &ld,
                                const LoadByte &lb, uint32_t n, uint32_t *out, int lane)
{
    return fz_emit_subblock_sc(gc, hdr, es, FzPieceScan<Load16, LoadByte>{ld, lb, n}, n, out, lane);
}

// =================================================================================================
// Window-interleaved piece geometry (lead #1) -- NOT used by the kernels yet; checked on the CPU by
// tests/hostmodel against zlib.
//
// Today lane l owns ONE contiguous 512-byte piece of the sub-block, so the bit offset of lane l + 1 is known only
// after lane l's whole piece has been counted: fz_emit_subblock scans the sub-block twice (count, emit).  Here the
// sub-block is cut into windows of 32 pieces of FZ_IPIECE bytes, piece p = window p / 32, lane p % 32, emitted in
// piece order.  Inside a window the count, the warp prefix and the emission all work on the 2 KiB the warp already
// holds in shared memory, and the only thing carried from window to window is the running bit position and the
// partial word at its end.  The tokeniser restarts at every piece start, as it does today at the 32 piece starts;
// tools/piece_geometry_study.py: no measurable size cost except on all-zero sub-blocks.
// =================================================================================================
#define FZ_IPIECE 64u
#define FZ_IWIN (FZ_IPIECE * 32u)

template <class Load16, class LoadByte>
struct FzWindowPieceScan {
    const Load16 &ld;
    const LoadByte &lb;
    uint32_t n, w;
    template <class Sink>
    FZ_HD void operator()(Sink &sink, int lane) const
    {
        const uint32_t b = w * FZ_IWIN + (uint32_t)lane * FZ_IPIECE;
        uint32_t e = b + FZ_IPIECE;
        if (e > n) e = n;
        if (b < e) fz_scan_piece(ld, lb, b, e, b ? (int)lb(b - 1) : -1, sink);
    }
};

template <class Load16, class LoadByte>
FZ_HD void fz_ph_hist_interleaved(uint32_t *hist, const Load16 &ld, const LoadByte &lb, uint32_t n, int lane)
{
    for (uint32_t w = 0; w * FZ_IWIN < n; w++) fz_ph_hist_sc(hist, FzWindowPieceScan<Load16, LoadByte>{ld, lb, n, w}, lane);
}

struct FzEmitStateI {
    FzEmitState es;        // per-window lane exchange (lane_bits, first / last word of every lane)
    uint32_t base_bits;    // bits emitted by the windows before this one
    uint32_t carry_in;     // what they left in the word that holds bit `base_bits` (its low base_bits & 31 bits)
    uint32_t carry_next;
    uint32_t pad;
};

template <class Scan>
FZ_HD void fz_ph_count_window(const FzGroupCode *gc, FzEmitStateI *st, const Scan &scan, bool first, int lane)
{
    FzCountSink cs{gc->cl, 0};
    scan(cs, lane);
    if (first && lane == 0) cs.bits += gc->hdr_nbits;
    st->es.lane_bits[lane] = cs.bits;
}

template <class Scan>
FZ_HD void fz_ph_emit_window(const FzGroupCode *gc, const uint32_t *hdr, FzEmitStateI *st, const Scan &scan, bool first, bool last,
                             uint32_t *out, int lane)
{
    FzEmitState *es = &st->es;
    uint32_t off = st->base_bits;
    for (int l = 0; l < lane; l++) off += es->lane_bits[l];
    FzEmitSink sink;
    sink.cl = gc->cl;
    sink.bw.init(out, off);
    if (first && lane == 0) {
        uint32_t nb = gc->hdr_nbits, w = 0;
        while (nb >= 32) { sink.bw.put(hdr[w++], 32); nb -= 32; }
        if (nb) sink.bw.put(hdr[w] & ((1u << nb) - 1), nb);
    }
    scan(sink, lane);
    if (last && lane == 31) {
        sink.bw.put(gc->cl[FZ_EOB] & 0xffffu, gc->cl[FZ_EOB] >> 16);
        sink.bw.put(0, 3);
        sink.bw.align_byte();
        sink.bw.put(0x0000u, 16);
        sink.bw.put(0xFFFFu, 16);
        es->total_bits = sink.bw.bitpos();
    }
    es->fw_idx[lane] = sink.bw.first_idx;
    es->crossed[lane] = sink.bw.crossed ? 1u : 0u;
    if (sink.bw.crossed) { es->fw_bits[lane] = sink.bw.first_bits; es->tw_bits[lane] = (uint32_t)sink.bw.acc; }
    else { es->fw_bits[lane] = (uint32_t)sink.bw.acc; es->tw_bits[lane] = 0; }
}

// words shared by several lanes (and by the window before): the lane that completes a word ORs in what the others left
FZ_HD void fz_ph_merge_window(FzEmitStateI *st, bool last, uint32_t *out, int lane)
{
    FzEmitState *es = &st->es;
    uint32_t carry = 0;
    int j = lane - 1;
    for (; j >= 0; j--) {
        if (es->crossed[j]) { carry |= es->tw_bits[j]; break; }
        carry |= es->fw_bits[j];
    }
    if (j < 0) carry |= st->carry_in;   // nobody before this lane completed a word: the previous window's bits are still in it
    if (es->crossed[lane]) out[es->fw_idx[lane]] = es->fw_bits[lane] | carry;
    if (lane == 31) {
        const uint32_t tail = es->crossed[31] ? es->tw_bits[31] : (es->fw_bits[31] | carry);
        if (last) {
            const uint32_t total_bits = es->total_bits;
            if (total_bits & 31) out[total_bits >> 5] = tail;
        } else st->carry_next = tail;
    }
}

FZ_HD void fz_ph_advance_window(FzEmitStateI *st, int lane)
{
    if (lane != 0) return;
    uint32_t b = st->base_bits;
    for (int l = 0; l < 32; l++) b += st->es.lane_bits[l];
    st->base_bits = b;
    st->carry_in = st->carry_next;
}

// Emit one sub-block with its group's code, window by window.  Same contract as fz_emit_subblock_sc; the exact size
// is known only after the emission (nothing is lost: a stored result ignores what was written to `out`).
// `win` supplies the windows: win.enter(w, lane) / win.leave(w, lane) bracket the work on window w (the device stages
// the 2 KiB there), win.scan(w) is the piece scan of that window.
template <class Windows>
FZ_HD uint32_t fz_emit_subblock_iw(const FzGroupCode *gc, const uint32_t *hdr, FzEmitStateI *st, Windows &win, uint32_t n,
                                   uint32_t *out, int lane)
{
    (void)lane;
    const uint32_t stored = fz_stored_size(n) | FZ_SIZE_STORED_FLAG;
    if (gc->stored) return stored;
    FZ_PHASE(if (lane == 0) { st->base_bits = 0; st->carry_in = 0; st->carry_next = 0; st->es.false_marker = 0; st->es.total_bits = 0; });
    const uint32_t nwin = (n + FZ_IWIN - 1) / FZ_IWIN;
    for (uint32_t w = 0; w < nwin; w++) {
        const bool first = w == 0, last = w + 1 == nwin;
        FZ_PHASE(win.enter(w, lane));
        FZ_PHASE(fz_ph_count_window(gc, st, win.scan(w), first, lane));
        {   // Stop as soon as the sub-block cannot beat a stored block any more: nothing beyond the slot is ever
            // written (a stored block is smaller than FZ_SLOT_STRIDE), and hopeless sub-blocks cost one count.
            // Every lane reads the same 32 counts, so the decision is uniform.  + 64 bits: end of block and marker.
            uint32_t tot = st->base_bits;
            for (int l = 0; l < 32; l++) tot += st->es.lane_bits[l];
            if ((tot + 64u) / 8u + (n >> FZ_MIN_GAIN_SHIFT) >= fz_stored_size(n)) return stored;
        }
        FZ_PHASE(fz_ph_emit_window(gc, hdr, st, win.scan(w), first, last, out, lane));
        FZ_PHASE(fz_ph_merge_window(st, last, out, lane));
        FZ_PHASE(fz_ph_advance_window(st, lane));
        FZ_PHASE(win.leave(w, lane));
    }
    const uint32_t total_bytes = st->es.total_bits / 8;
    if (total_bytes + (n >> FZ_MIN_GAIN_SHIFT) >= fz_stored_size(n)) return stored;
    FZ_PHASE(fz_ph_check_marker(&st->es, out, total_bytes, lane));
    if (st->es.false_marker) return stored;
    return total_bytes;
}

// windows behind random-access loaders (the CPU model; ragged or unaligned sub-blocks on the device)
template <class Load16, class LoadByte>
struct FzLoaderWindows {
    const Load16 &ld;
    const LoadByte &lb;
    uint32_t n;
    FZ_HD void enter(uint32_t, int) {}
    FZ_HD void leave(uint32_t, int) {}
    FZ_HD FzWindowPieceScan<Load16, LoadByte> scan(uint32_t w) const { return FzWindowPieceScan<Load16, LoadByte>{ld, lb, n, w}; }
};

template <class Load16, class LoadByte>
FZ_HD uint32_t fz_emit_subblock_interleaved(const FzGroupCode *gc, const uint32_t *hdr, FzEmitStateI *st, const Load16 &ld,
                                            const LoadByte &lb, uint32_t n, uint32_t *out, int lane)
{
    FzLoaderWindows<Load16, LoadByte> win{ld, lb, n};
    return fz_emit_subblock_iw(gc, hdr, st, win, n, out, lane);
}
