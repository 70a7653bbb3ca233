// fz_inflate.cuh -- thread-serial raw-inflate (RFC 1951: stored / fixed / dynamic, any distance) of ONE
// deflate fragment.  Replaces what the reference gets from zlib's inflate() behind mzlib_inf
// (reference zip.c:262-284) for one payload -- or, on streams made by our encoder, for one
// sub-block (one GPU thread per sub-block; see fz_kernels.cu).
//
// `__host__ __device__`: tests/hostmodel runs this very source on the CPU against zlib streams.
//
// Decoding uses left-aligned canonical codes: take the next 15 stream bits, bit-reversed into a
// 15-bit number w; the code length is the first l with w < limit[l]; the symbol index is
// (w >> (15-l)) + delta[l] into the symbols sorted by (length, value).  limit and delta are packed
// into one 32-bit word per length (limit << 16 | delta & 0xffff) held in fifteen SCALAR fields, so
// they live in registers (arrays ended up in local memory and made the kernel 50x slower); only the
// sorted symbol tables and the per-length counters are in (shared) memory.
#pragma once
#include "fz_common.cuh"

// per-thread tables; STRIDE interleaves the threads of a block in shared memory
template <int STRIDE>
struct FzInfTab {
    uint16_t *ll;   // 288 sorted literal/length symbols
    uint16_t *dd;   // 32 sorted distance symbols
    uint16_t *cnt;  // 32 counters: [0..15] literal/length per code length, [16..31] distance
    FZ_HD uint16_t &L(int i) const { return ll[i * STRIDE]; }
    FZ_HD uint16_t &D(int i) const { return dd[i * STRIDE]; }
    FZ_HD uint16_t &C(int i) const { return cnt[i * STRIDE]; }
};
#define FZ_INF_TAB_U16 (288 + 32 + 32)  // uint16 entries per decoding thread

// Bit reader over an arbitrarily aligned byte range: aligned 32-bit loads, the next word fetched one refill ahead.
// (A variant fetching aligned 16-byte chunks one chunk ahead cut the L2 round trips fourfold -- ncu shows a 3.5 % L1
// hit rate here, L1 being carved away by shared memory -- but its longer refill and 16 more registers cost more than
// that bought on every input except exponent planes: measured, not kept.)
struct FzBitReader {
    const uint32_t *w;     // address of the word held in `nxt` (may run past wend: zero bits are fed then)
    const uint32_t *wend;  // one past the last word that holds input bytes
    uint64_t acc;
    uint32_t nxt;          // the next word, loaded one refill ahead so its latency hides behind decoding
    int nacc;              // valid bits in acc
    int tail_pad;          // bits of the last word that lie beyond the end of the input
    FZ_HD void init(const uint8_t *in, size_t in_len)
    {
        const uintptr_t a = (uintptr_t)in;
        const unsigned sk = (unsigned)(a & 3);
        w = (const uint32_t *)(a - sk);
        const size_t nw = (in_len + sk + 3) / 4;
        wend = w + nw;
        tail_pad = (int)(nw * 32 - 8 * sk - in_len * 8);
        acc = 0; nacc = 0; nxt = 0;
        if (nw > 0) { acc = (uint64_t)(*w++ >> (8 * sk)); nacc = 32 - 8 * (int)sk; }
        else tail_pad = 0;
        if (w < wend) nxt = *w;
    }
    // bits of real input not yet consumed (negative after an overrun); derived, not tracked per symbol
    FZ_HD int64_t bits_left() const { return (int64_t)(wend - w) * 32 + nacc - tail_pad; }
    FZ_HD void refill()  // guarantees nacc >= 32 (zero bits past the end of the input)
    {
        if (nacc < 32) {
            acc |= (uint64_t)nxt << nacc;
            nacc += 32;
            w++;
            nxt = (w < wend) ? *w : 0u;
        }
    }
    FZ_HD uint32_t peek(int n) const { return (uint32_t)acc & ((1u << n) - 1); }
    FZ_HD void drop(int n) { acc >>= n; nacc -= n; }
    FZ_HD uint32_t get(int n) { const uint32_t v = peek(n); drop(n); return v; }  // n <= 16, after refill
    FZ_HD void align_byte() { const int k = (int)(bits_left() & 7); drop(k); }
    // byte-aligned reader: address of the next unread input byte (`nxt` holds the word at w, acc the bytes before it)
    FZ_HD const uint8_t *byte_ptr() const { return (const uint8_t *)w - (nacc >> 3); }
};

struct FzByteWriter {
    uint8_t *out;      // 4-byte aligned base; the fragment's bytes are out[op0 .. cap)
    uint32_t op, cap;  // write position / end, both relative to `out`
    uint32_t op0;      // 0..3: leading bytes of the first word that belong to a neighbouring fragment
    uint32_t ow;       // pending bytes of the current word
    int prev_byte;     // the byte just before the fragment (block-parallel decode), or -1: nothing may be referenced there
    bool dry;          // count only, store nothing (measure pass of the block-parallel decode)
    bool non_rle;      // dry: a match with a distance other than 1 was seen (its bytes are unknown without history)
    bool starts_with_match;  // dry: the fragment begins with a distance-1 match reaching the byte before it
    uint32_t lastc;    // dry: value of the last literal, 0x100 = none yet
    uint32_t orv;      // dry: OR of all literal values (0 = the fragment is nothing but zero bytes)
    FZ_HD void init(uint8_t *o, uint32_t c)
    {
        const uint32_t mis = (uint32_t)((uintptr_t)o & 3u);
        out = o - mis; op0 = mis; op = mis; cap = c + mis; ow = 0;
        prev_byte = -1; dry = false; non_rle = false; starts_with_match = false; lastc = 0x100; orv = 0;
    }
    FZ_HD uint32_t produced() const { return op - op0; }
    FZ_HD void store_word(uint32_t widx_bytes, uint32_t w)  // the word starting at byte offset widx_bytes is complete
    {
        if (widx_bytes == 0 && op0) { for (uint32_t i = op0; i < 4; i++) out[i] = (uint8_t)(w >> (8 * i)); }
        else *(uint32_t *)(out + widx_bytes) = w;
    }
    FZ_HD void put(uint32_t c)
    {
        if (dry) { op++; lastc = c; orv |= c; return; }
        ow |= c << ((op & 3) * 8);
        op++;
        if ((op & 3) == 0) { store_word(op - 4, ow); ow = 0; }
    }
    FZ_HD uint32_t back(uint32_t dist) const  // byte written `dist` positions ago (dist <= produced())
    {
        const uint32_t p = op - dist;
        if ((p >> 2) == (op >> 2)) return (ow >> ((p & 3) * 8)) & 0xffu;  // still pending in ow
        return out[p];
    }
    // append cnt (1..3) bytes packed little-endian in v; the caller checked op + cnt <= cap
    FZ_HD void putn(uint32_t v, uint32_t cnt)
    {
        if (dry) { op += cnt; lastc = (v >> (8 * (cnt - 1))) & 0xffu; orv |= v; return; }
        const uint32_t sh = (op & 3) * 8;
        const uint64_t t = (uint64_t)ow | ((uint64_t)v << sh);
        const uint32_t nb = sh + cnt * 8;  // bits now pending
        op += cnt;
        if (nb >= 32) { store_word((op & ~3u) - 4, (uint32_t)t); ow = (uint32_t)(t >> 32); }
        else ow = (uint32_t)t;
    }
    FZ_HD void fill(uint32_t c, uint32_t len)  // len copies of byte c (a distance-1 match), whole words where possible
    {
        if (dry) { op += len; return; }
        while (len && (op & 3)) { put(c); len--; }
        const uint32_t w = c * 0x01010101u;
        // long runs (zero planes are nothing else): 16-byte stores once the address allows it
        while (len >= 4 && ((uintptr_t)(out + op) & 15u)) { *(uint32_t *)(out + op) = w; op += 4; len -= 4; }
        while (len >= 16) {
#if defined(__CUDA_ARCH__)
            *(uint4 *)(out + op) = make_uint4(w, w, w, w);
#else
            uint32_t *q = (uint32_t *)(out + op);
            q[0] = w; q[1] = w; q[2] = w; q[3] = w;
#endif
            op += 16; len -= 16;
        }
        while (len >= 4) { *(uint32_t *)(out + op) = w; op += 4; len -= 4; }
        while (len) { put(c); len--; }
    }
    // append len bytes read from src (any alignment; reads stay inside the aligned words that hold src[0, len))
    FZ_HD void copy_in(const uint8_t *src, uint32_t len)
    {
        if (len == 0) return;
        if (dry) { op += len; lastc = src[len - 1]; orv |= 1u; return; }   // (not examined: counts as "not all zero")
        while (len && (op & 3)) { put(*src++); len--; }
        if (len >= 4) {
            const unsigned sk = (unsigned)((uintptr_t)src & 3);
            const uint32_t *sw = (const uint32_t *)(src - sk);
            const uint32_t nw = len >> 2;
            if (sk == 0) {
                for (uint32_t i = 0; i < nw; i++) { *(uint32_t *)(out + op) = sw[i]; op += 4; }
            } else {
                uint32_t a = sw[0];
                for (uint32_t i = 0; i < nw; i++) {
                    const uint32_t b = sw[i + 1];   // holds at least one byte of src[4i, 4i+4): inside the range
                    *(uint32_t *)(out + op) = (a >> (8 * sk)) | (b << (32 - 8 * sk));
                    op += 4;
                    a = b;
                }
            }
            src += (size_t)nw * 4;
            len &= 3;
        }
        while (len) { put(*src++); len--; }
    }
    FZ_HD void finish()
    {
        if (dry) return;
        const uint32_t r = op & 3, w0 = op - r;
        for (uint32_t i = (w0 == 0 ? op0 : 0u); i < r; i++) out[w0 + i] = (uint8_t)(ow >> (8 * i));
    }
};

// first-level lookup table of the warp-shared fast path: index = next FZ_LUT_BITS stream bits.
//   entry == 0                 : code longer than FZ_LUT_BITS -> canonical search
//   bits  0..8   sym1          : first symbol (literal, end-of-block or length code)
//   bits  9..16  sym2          : second literal, bits 17..24 sym3: third literal (when count says so)
//   bits 25..28  total length  : code bits consumed by the `count` symbols of this entry
//   bits 29..30  count         : 1..3 symbols; entries with count > 1 hold literals only
#ifndef FZ_LUT_BITS
#define FZ_LUT_BITS 11
#endif
#define FZ_LUT_SIZE (1 << FZ_LUT_BITS)
#define FZ_LUT_MATCH (1u << 24)   // entry of a length symbol: bits 9..17 base length, bits 18..20 extra bit count
#define FZ_LUT_ENTRY(s1, s2, s3, total, cnt) \
    ((uint32_t)(s1) | ((uint32_t)(s2) << 9) | ((uint32_t)(s3) << 17) | ((uint32_t)(total) << 25) | ((uint32_t)(cnt) << 29))

#define FZ_INF_OK 0
#define FZ_INF_E_INPUT (-1)     // ran out of input / truncated
#define FZ_INF_E_DATA (-2)      // invalid block type, lengths, code or distance
#define FZ_INF_E_SPACE (-3)     // more output than out_cap
#define FZ_INF_E_HISTORY (-4)   // distance reaches before the start of this fragment

// one canonical code: p<l> = limit_l << 16 | (delta_l & 0xffff) for code length l
struct FzCode {
    uint32_t p1, p2, p3, p4, p5, p6, p7, p8, p9, p10, p11, p12, p13, p14, p15;
};

#define FZ_FOR_LEN_1_15(M) M(1) M(2) M(3) M(4) M(5) M(6) M(7) M(8) M(9) M(10) M(11) M(12) M(13) M(14) M(15)

// Build the code from per-length counts cnt(l), l = 1..15 (read through `rd`), and turn the counts
// into first-index offsets in place (written through `wr`).  Returns the Kraft remainder:
// 0 complete, > 0 incomplete, < 0 over-subscribed.
template <class Rd, class Wr>
FZ_HD int fz_code_build(FzCode &r, const Rd &rd, const Wr &wr)
{
    uint32_t code = 0, idx = 0;
    int left = 1;
#define FZ_BUILD_STEP(L)                                                         \
    {                                                                            \
        code <<= 1; left <<= 1;                                                  \
        const uint32_t c = rd(L);                                                \
        left -= (int)c;                                                          \
        wr(L, idx);                                                              \
        const uint32_t delta = (idx - code) & 0xffffu;                           \
        code += c; idx += c;                                                     \
        r.p##L = ((code << (15 - L)) << 16) | delta;                             \
    }
    FZ_FOR_LEN_1_15(FZ_BUILD_STEP)
#undef FZ_BUILD_STEP
    return left;
}

FZ_HD uint32_t fz_rev15(uint32_t v)  // reverse the low 15 bits
{
#if defined(__CUDA_ARCH__)
    return __brev(v) >> 17;
#else
    uint32_t r = 0;
    for (int i = 0; i < 15; i++) { r = (r << 1) | ((v >> i) & 1); }
    return r;
#endif
}

// decode one symbol index (into the sorted table) from the next 15 bits; returns the code length (0 = invalid).
// Branch-free on purpose: the 32 lanes of a warp decode 32 different streams, so any data-dependent
// branch here would serialise them.
FZ_HD int fz_decode_idx(const FzCode &r, uint32_t bits15, uint32_t &idx)
{
    const uint32_t w = fz_rev15(bits15);
    const uint32_t x = (w << 16) | 0xffffu;  // x < p_l  <=>  w < limit_l
    // limits are non-decreasing in l: the code length is 1 + #{l : x >= p_l}; sel = p of that length
    uint32_t len = 1, sel = r.p15;
#define FZ_DEC_SEL(L) sel = (x < r.p##L) ? r.p##L : sel;
    FZ_DEC_SEL(14) FZ_DEC_SEL(13) FZ_DEC_SEL(12) FZ_DEC_SEL(11) FZ_DEC_SEL(10) FZ_DEC_SEL(9) FZ_DEC_SEL(8)
    FZ_DEC_SEL(7) FZ_DEC_SEL(6) FZ_DEC_SEL(5) FZ_DEC_SEL(4) FZ_DEC_SEL(3) FZ_DEC_SEL(2) FZ_DEC_SEL(1)
#undef FZ_DEC_SEL
#define FZ_DEC_CNT(L) len += (x >= r.p##L) ? 1u : 0u;
    FZ_FOR_LEN_1_15(FZ_DEC_CNT)
#undef FZ_DEC_CNT
    idx = (w >> (15 - (len & 15))) + (uint32_t)((int32_t)(sel << 16) >> 16);
    return len <= 15 ? (int)len : 0;
}

// first-level table entry for the LUT_BITS-bit pattern e: the first symbol and up to two more literals
template <int LUT_BITS, class Tab>
FZ_HD uint32_t fz_lut_entry_bits(const FzCode &LL, const Tab &tab, uint32_t e)
{
    uint32_t idx;
    int l = fz_decode_idx(LL, e, idx);
    if (!(l >= 1 && l <= LUT_BITS && idx < 288)) return 0;
    const uint32_t s1 = tab.L((int)idx);
    uint32_t s2 = 0, s3 = 0, cnt = 1, total = (uint32_t)l;
    if (s1 < 256u && total < (uint32_t)LUT_BITS) {
        l = fz_decode_idx(LL, e >> total, idx);
        if (l >= 1 && total + l <= (uint32_t)LUT_BITS && idx < 288 && tab.L((int)idx) < 256u) {
            s2 = tab.L((int)idx); total += l; cnt = 2;
            if (total < (uint32_t)LUT_BITS) {
                l = fz_decode_idx(LL, e >> total, idx);
                if (l >= 1 && total + l <= (uint32_t)LUT_BITS && idx < 288 && tab.L((int)idx) < 256u) {
                    s3 = tab.L((int)idx); total += l; cnt = 3;
                }
            }
        }
    }
    uint32_t ent = FZ_LUT_ENTRY(s1, s2, s3, total, cnt);
    // length symbols: base length and number of extra bits ride in the (otherwise empty) literal slots, so that the
    // register-only loops can take a whole distance-1 match without leaving for the general step
    if (s1 >= 257u && s1 <= 285u) ent |= FZ_LUT_MATCH | (fz_len_base(s1 - 257u) << 9) | (fz_len_extra_bits(s1 - 257u) << 18);
    return ent;
}
template <class Tab>
FZ_HD uint32_t fz_lut_entry(const FzCode &LL, const Tab &tab, uint32_t e) { return fz_lut_entry_bits<FZ_LUT_BITS>(LL, tab, e); }

// The same table (entry for entry what fz_lut_entry_bits gives), built by the 32 lanes of a warp WITHOUT a canonical search
// per entry.  In MSB-first order the patterns of a canonical code are consecutive intervals, one per symbol in sorted
// order: every lane takes 2^LUT_BITS / 32 consecutive patterns, finds the symbol its first one belongs to (one search) and
// walks on from there; the bit-reversed pattern is the table index.  Two and three literals per entry then come from
// looking the rest of the pattern up in the table itself -- entries are packed from the top down, 32 at a time, because
// pattern e only ever looks at patterns below e / 2, which are still single symbols then.  (Per block of a zlib stream
// the table cost 6144 searches of ~50 instructions: a third of the block-parallel decoder's instructions.)
// Call with all 32 lanes (the host model loops them: `phase` 0 = fill, 1.. = packing batches; see fz_lut_build_phases).
template <int LUT_BITS>
FZ_HD uint32_t fz_lut_single(uint32_t sym, uint32_t l)
{
    uint32_t ent = FZ_LUT_ENTRY(sym, 0, 0, l, 1);
    if (sym >= 257u && sym <= 285u) ent |= FZ_LUT_MATCH | (fz_len_base(sym - 257u) << 9) | (fz_len_extra_bits(sym - 257u) << 18);
    return ent;
}

template <int LUT_BITS, class Tab>
FZ_HD void fz_lut_fill_lane(uint32_t *lut, const FzCode &LL, const Tab &tab, int lane)
{
    const uint32_t per = (1u << LUT_BITS) / 32u;
    const uint32_t *P = (const uint32_t *)&LL;          // P[L - 1] = limit_L << 16 | delta_L
    uint32_t m = (uint32_t)lane * per;                  // first pattern of this lane, MSB-first, LUT_BITS bits
    // patterns below code_end_L << (LUT_BITS - L) start with a code of at most L bits; the interval of length L begins
    // where the one of length L - 1 ended (first_code_L = 2 * code_end_{L-1})
    uint32_t L = 0, code_end = 0;
    for (uint32_t i = 0; i < per; i++, m++) {
        while (L <= (uint32_t)LUT_BITS && m >= (code_end << ((uint32_t)LUT_BITS - L))) {   // (empty lengths are skipped)
            L++;
            if (L <= (uint32_t)LUT_BITS) code_end = (P[L - 1] >> 16) >> (15u - L);
        }
        uint32_t ent = 0;                                // past the last interval: a longer code (or none)
        if (L <= (uint32_t)LUT_BITS) {
            const uint32_t code = m >> ((uint32_t)LUT_BITS - L);                   // the L-bit code this pattern starts with
            const uint32_t idx = (code + (P[L - 1] & 0xffffu)) & 0xffffu;          // index into the sorted symbols
            if (idx < 288u) ent = fz_lut_single<LUT_BITS>(tab.L((int)idx), L);
        }
        lut[fz_bitrev(m, LUT_BITS)] = ent;
    }
}

// entry e with up to two more literals looked up in the (still single-symbol) entries below it
template <int LUT_BITS>
FZ_HD uint32_t fz_lut_pack(const uint32_t *lut, uint32_t e)
{
    const uint32_t a = lut[e];
    const uint32_t s1 = a & 511u;
    uint32_t total = (a >> 25) & 15u;
    if (a == 0 || s1 >= 256u || total >= (uint32_t)LUT_BITS) return a;
    const uint32_t b = lut[e >> total];
    const uint32_t l2 = (b >> 25) & 15u;
    if (b == 0 || (b & 511u) >= 256u || total + l2 > (uint32_t)LUT_BITS) return a;
    uint32_t s2 = b & 511u, s3 = 0, cnt = 2;
    total += l2;
    if (total < (uint32_t)LUT_BITS) {
        const uint32_t c = lut[e >> total];
        const uint32_t l3 = (c >> 25) & 15u;
        if (c != 0 && (c & 511u) < 256u && total + l3 <= (uint32_t)LUT_BITS) { s3 = c & 511u; total += l3; cnt = 3; }
    }
    return FZ_LUT_ENTRY(s1, s2, s3, total, cnt);
}

// Which value of a 1-bit distance code means "distance 1" (what run-length streams use): 0 or 1, or 2 = this block's
// distance code is not of that kind (see FzInflater::dd1).
FZ_HD uint32_t fz_dd1_run_bit(uint32_t dd1)
{
    if (!dd1) return 2u;
    if ((dd1 & 0xffu) == 0u) return 0u;
    if (((dd1 >> 16) & 3u) == 2u && ((dd1 >> 8) & 0xffu) == 0u) return 1u;
    return 2u;
}

// ---------------------------------------------------------------------------------------------------
// The inflater as a resumable state machine: step() does one unit of work (one block header, or one
// literal/length symbol including its match copy) and returns false when the fragment is finished.
// The GPU kernel drives the 32 lanes of a warp in lock step -- `while (__any_sync(~0u, live)) if (live)
// live = inf.step();` -- so that lanes reconverge after every symbol; a plain nested loop left each lane
// running on its own (1.25 active threads per instruction, 50x slower).
// ---------------------------------------------------------------------------------------------------
template <class Tab>
struct FzInflater {
    FzBitReader br;
    FzByteWriter bw;
    FzCode *LL, *DD;         // the codes of the current block, in storage the caller binds (bind_codes) before the first
                             // step: 30 words that a warp decoding with ONE code keeps in shared memory instead of
                             // 30 registers per lane
    Tab tab;
    size_t in_len;
    int rc;
    bool last, in_body;
    bool shared_tab;  // tables are shared with other lanes: this lane must not rebuild them (no further coded block)
    uint32_t *own_lut;  // optional FZ_LUT_SIZE-entry table this thread (re)builds after every block header
    int lut_bits;       // index width of the table handed to step_lut (FZ_LUT_BITS unless the caller built a wider one)
    bool one_block;     // stop after the first coded block (block-parallel decode of zlib-made streams)
    bool saw_eob;       // ... and it ended properly with its end-of-block symbol
    int ll_left, dd_left;  // Kraft remainders of the last dynamic header (0 = complete code)
    uint32_t dd1;          // distance code with 1-bit codes only (what RLE-style streams carry): 1 << 31 | number of codes
                           // << 16 | symbol of code '1' << 8 | symbol of code '0'; 0 = general code
    uint32_t eob_len;      // code length of the end-of-block symbol in the last header

    FZ_HD void start(const uint8_t *in, size_t in_len_, uint8_t *out, uint32_t out_cap, const Tab &t)
    {
        br.init(in, in_len_);
        bw.init(out, out_cap);
        tab = t;
        in_len = in_len_;
        rc = FZ_INF_OK;
        last = false;
        in_body = false;
        shared_tab = false;
        own_lut = nullptr;
        lut_bits = FZ_LUT_BITS;
        one_block = false; saw_eob = false; ll_left = 0; dd_left = 0; eob_len = 1; dd1 = 0;
    }
    // start `bit` bits into the input (block-parallel decode)
    FZ_HD void start_at_bit(const uint8_t *in, size_t in_len_, uint64_t bit, uint8_t *out, uint32_t out_cap, const Tab &t)
    {
        start(in + (bit >> 3), in_len_ - (size_t)(bit >> 3), out, out_cap, t);
        br.refill();
        br.drop((int)(bit & 7));
    }
    FZ_HD void bind_codes(FzCode *ll, FzCode *dd) { LL = ll; DD = dd; }
    FZ_HD uint64_t consumed_bits() const { return (uint64_t)((int64_t)in_len * 8 - br.bits_left()); }

    // returns true while there is more to do
    FZ_HD bool step()
    {
        if (in_body) return body_symbol(own_lut);
        return block_header();
    }
    FZ_HD bool step_lut(const uint32_t *lut)
    {
        if (in_body) return body_symbol(lut);
        return block_header();
    }

    FZ_HD int finish(uint32_t *out_n, size_t *in_used)
    {
        bw.finish();
        if (rc == FZ_INF_OK && br.bits_left() < 0) rc = FZ_INF_E_INPUT;
        *out_n = bw.produced();
        const int64_t used_bits = (int64_t)in_len * 8 - br.bits_left();
        *in_used = (size_t)((used_bits + 7) / 8);
        return rc;
    }

    FZ_HD bool fail(int code) { rc = code; return false; }

    FZ_HD bool body_symbol(const uint32_t *lut)
    {
        br.refill();
        // measure pass over a speculative block: a false candidate may run off the end of the input, where the
        // reader feeds zero bits for ever
        if (bw.dry && br.bits_left() < 0) return fail(FZ_INF_E_INPUT);
        uint32_t idx, sym;
        int l;
        const uint32_t e = lut ? lut[br.peek(lut_bits)] : 0u;
        if (e) {
            const uint32_t cnt = e >> 29;
            sym = e & 511u;
            if (sym < 256u && bw.op + cnt <= bw.cap) {   // 1..3 literals at once
                br.drop((int)((e >> 25) & 15u));
                bw.putn((e & 255u) | ((e >> 1) & 0xffff00u), cnt);
                return true;
            }
            if (cnt > 1) {   // not enough room for all of them: take the first one the slow way
                l = fz_decode_idx(*LL, br.peek(15), idx);
                if (l == 0) return fail(FZ_INF_E_DATA);
                br.drop(l);
                sym = tab.L((int)idx);
            } else br.drop((int)((e >> 25) & 15u));
        } else {
            l = fz_decode_idx(*LL, br.peek(15), idx);
            if (l == 0) return fail(FZ_INF_E_DATA);
            br.drop(l);
            sym = tab.L((int)idx);
        }
        if (sym < 256) {
            if (bw.op >= bw.cap) return fail(FZ_INF_E_SPACE);
            bw.put(sym);
            return true;
        }
        if (sym == FZ_EOB) {
            in_body = false;
            if (br.bits_left() < 0) return fail(FZ_INF_E_INPUT);
            saw_eob = true;
            return !last && !one_block;
        }
        sym -= 257;
        if (sym >= 29) return fail(FZ_INF_E_DATA);
        br.refill();
        const uint32_t len = fz_len_base(sym) + br.get((int)fz_len_extra_bits(sym));
        uint32_t ds;
        if (dd1) {
            const uint32_t b = br.get(1);
            if (b >= ((dd1 >> 16) & 3u)) return fail(FZ_INF_E_DATA);
            ds = (dd1 >> (8 * b)) & 0xffu;
        } else {
            l = fz_decode_idx(*DD, br.peek(15), idx);
            if (l == 0) return fail(FZ_INF_E_DATA);
            br.drop(l);
            ds = tab.D((int)idx);
        }
        if (ds >= 30) return fail(FZ_INF_E_DATA);
        br.refill();
        const uint32_t dist = fz_dist_base(ds) + br.get((int)fz_dist_extra_bits(ds));
        if (bw.op + len > bw.cap) return fail(FZ_INF_E_SPACE);
        if (dist > bw.produced()) {
            // only the byte just before the fragment may be reached, by a run that continues across the block boundary
            if (!(bw.prev_byte >= 0 && dist == 1)) return fail(FZ_INF_E_HISTORY);
            if (bw.dry) bw.starts_with_match = true;
            bw.fill((uint32_t)bw.prev_byte, len);
            return true;
        }
        if (bw.dry) {
            if (dist != 1) bw.non_rle = true;
            bw.op += len;
            return true;
        }
        if (dist == 1) {
            bw.fill(bw.back(1), len);
        } else {
            for (uint32_t i = 0; i < len; i++) bw.put(bw.back(dist));
        }
        return true;
    }

    FZ_HD bool block_header()
    {
        if (br.bits_left() < 3) return false;                       // nothing but padding left
        if (bw.op == bw.cap && br.bits_left() < 8) return false;    // full output, only pad bits left
        br.refill();
        last = br.get(1) != 0;
        const uint32_t type = br.get(2);
        if (type == 0) {
            br.align_byte();
            br.refill();
            const uint32_t len = br.get(16);
            br.refill();
            const uint32_t nlen = br.get(16);
            if ((len ^ 0xFFFFu) != nlen) return fail(FZ_INF_E_DATA);
            if (br.bits_left() < (int64_t)len * 8) return fail(FZ_INF_E_INPUT);
            if (bw.op + len > bw.cap) return fail(FZ_INF_E_SPACE);
            // the payload of a stored block is copied word-wise straight from the input, not through the bit reader
            const int64_t rest = br.bits_left() - (int64_t)len * 8;
            const uint8_t *src = br.byte_ptr();
            bw.copy_in(src, len);
            br.init(src + len, (size_t)(rest >> 3));
            return !last;
        }
        if (type == 3) return fail(FZ_INF_E_DATA);
        if (shared_tab) return fail(FZ_INF_E_DATA);  // a second coded block would overwrite the shared tables

        auto rd_ll = [&](int l) -> uint32_t { return tab.C(l); };
        auto wr_ll = [&](int l, uint32_t v) { tab.C(l) = (uint16_t)v; };
        auto rd_dd = [&](int l) -> uint32_t { return tab.C(16 + l); };
        auto wr_dd = [&](int l, uint32_t v) { tab.C(16 + l) = (uint16_t)v; };

        if (type == 1) {
            // fixed code: litlen lengths 8 (0-143), 9 (144-255), 7 (256-279), 8 (280-287); 32 distance codes of 5 bits
            for (int l = 0; l < 32; l++) tab.C(l) = 0;
            tab.C(7) = 24; tab.C(8) = 152; tab.C(9) = 112;
            tab.C(16 + 5) = 32;
            dd1 = 0;
            fz_code_build(*LL, rd_ll, wr_ll);
            fz_code_build(*DD, rd_dd, wr_dd);
            for (int i = 0; i < 24; i++) tab.L(i) = (uint16_t)(256 + i);
            for (int i = 0; i < 144; i++) tab.L(24 + i) = (uint16_t)i;
            for (int i = 0; i < 8; i++) tab.L(168 + i) = (uint16_t)(280 + i);
            for (int i = 0; i < 112; i++) tab.L(176 + i) = (uint16_t)(144 + i);
            for (int i = 0; i < 32; i++) tab.D(i) = (uint16_t)i;
            if (own_lut) for (uint32_t e = 0; e < FZ_LUT_SIZE; e++) own_lut[e] = fz_lut_entry(*LL, tab, e);
            in_body = true;
            return true;
        }

        // order of the code-length code lengths (RFC 1951 3.2.7), 5 bits each, packed
        const uint64_t order_lo = 16ull | (17ull << 5) | (18ull << 10) | (0ull << 15) | (8ull << 20) | (7ull << 25) | (9ull << 30) |
                                  (6ull << 35) | (10ull << 40) | (5ull << 45) | (11ull << 50) | (4ull << 55);
        const uint64_t order_hi = 12ull | (3ull << 5) | (13ull << 10) | (2ull << 15) | (14ull << 20) | (1ull << 25) | (15ull << 30);
        br.refill();
        const uint32_t hlit = br.get(5) + 257, hdist = br.get(5) + 1, hclen = br.get(4) + 4;
        if (hlit > 286 || hdist > 30) return fail(FZ_INF_E_DATA);
        // code-length code: 19 symbols of <= 7 bits; everything about it is kept packed in registers
        uint64_t clpack = 0;  // 3 bits per symbol
        for (uint32_t i = 0; i < hclen; i++) {
            br.refill();
            const uint32_t sym = (uint32_t)((i < 12 ? order_lo >> (5 * i) : order_hi >> (5 * (i - 12))) & 31u);
            clpack |= (uint64_t)br.get(3) << (3 * sym);
        }
        uint64_t ccnt = 0;  // 8 bits per code length 0..7
        for (int s = 0; s < 19; s++) ccnt += 1ull << (8 * ((clpack >> (3 * s)) & 7u));
        uint64_t coffs = 0;  // first sorted index per code length, 8 bits each
        FzCode CL;
        auto rd_cl = [&](int l) -> uint32_t { return l <= 7 ? (uint32_t)((ccnt >> (8 * l)) & 0xffu) : 0u; };
        auto wr_cl = [&](int l, uint32_t v) { if (l <= 7) coffs |= (uint64_t)v << (8 * l); };
        if (fz_code_build(CL, rd_cl, wr_cl) != 0) return fail(FZ_INF_E_DATA);  // zlib requires a complete code here
        uint64_t clsym_lo = 0, clsym_hi = 0;  // sorted symbols, 5 bits each (12 + 7)
        for (uint32_t s = 0; s < 19; s++) {
            const uint32_t l = (uint32_t)((clpack >> (3 * s)) & 7u);
            if (l) {
                const uint32_t pos = (uint32_t)((coffs >> (8 * l)) & 0xffu);
                coffs += 1ull << (8 * l);
                if (pos < 12) clsym_lo |= (uint64_t)s << (5 * pos); else clsym_hi |= (uint64_t)s << (5 * (pos - 12));
            }
        }
        // two passes over the code-length data: count per length, then place the sorted symbols
        const FzBitReader mark = br;
        for (int pass = 0; pass < 2; pass++) {
            if (pass == 0) { for (int l = 0; l < 32; l++) tab.C(l) = 0; }
            else {
                uint32_t dlong = 0;   // distance codes longer than one bit
                for (int l = 2; l <= 15; l++) dlong += tab.C(16 + l);
                const uint32_t d1 = tab.C(16 + 1);
                dd1 = (dlong == 0 && d1 >= 1 && d1 <= 2) ? (0x80000000u | (d1 << 16)) : 0u;
                const int e1 = fz_code_build(*LL, rd_ll, wr_ll);
                const int e2 = fz_code_build(*DD, rd_dd, wr_dd);
                if (e1 < 0 || e2 < 0) return fail(FZ_INF_E_DATA);  // over-subscribed
                ll_left = e1; dd_left = e2;
                br = mark;
            }
            uint32_t i = 0, prev = 0;
            const uint32_t total = hlit + hdist;
            if (pass == 0) eob_len = 0;
            while (i < total) {
                br.refill();
                uint32_t idx;
                const int l = fz_decode_idx(CL, br.peek(15), idx);
                if (l == 0 || l > 7 || idx >= 19) return fail(FZ_INF_E_DATA);
                br.drop(l);
                const uint32_t s = (uint32_t)((idx < 12 ? clsym_lo >> (5 * idx) : clsym_hi >> (5 * (idx - 12))) & 31u);
                uint32_t rep = 1, val = s;
                if (s == 16) { if (i == 0) return fail(FZ_INF_E_DATA); val = prev; rep = 3 + br.get(2); }
                else if (s == 17) { val = 0; rep = 3 + br.get(3); }
                else if (s == 18) { val = 0; rep = 11 + br.get(7); }
                if (i + rep > total) return fail(FZ_INF_E_DATA);
                prev = val;
                if (val == 0) { i += rep; continue; }
                if (i <= FZ_EOB && FZ_EOB < i + rep) eob_len = val;
                if (pass == 0) {
                    for (uint32_t k = 0; k < rep; k++, i++) tab.C((i < hlit ? 0 : 16) + (int)val)++;
                } else {
                    for (uint32_t k = 0; k < rep; k++, i++) {
                        if (i < hlit) { const int o = tab.C((int)val)++; tab.L(o) = (uint16_t)i; }
                        else { const int o = tab.C(16 + (int)val)++; tab.D(o) = (uint16_t)(i - hlit); }
                    }
                }
            }
            if (br.bits_left() < 0) return fail(FZ_INF_E_INPUT);
        }
        if (dd1) dd1 |= (uint32_t)tab.D(0) | ((dd1 >> 16) & 2u ? (uint32_t)tab.D(1) << 8 : 0u);   // symbols are < 30
        if (own_lut) for (uint32_t e = 0; e < FZ_LUT_SIZE; e++) own_lut[e] = fz_lut_entry(*LL, tab, e);
        in_body = true;
        return true;
    }
};

// Inflate one fragment in one go.  Stops after a BFINAL block, or at the end of the input on a block boundary.
//   *out_n   bytes produced
//   *in_used input bytes consumed (rounded up to whole bytes)
// `out` must be 4-byte aligned.  Reads whole aligned 32-bit words around [in, in+in_len).
template <class Tab>
FZ_HD int fz_inflate(const uint8_t *in, size_t in_len, uint8_t *out, uint32_t out_cap, const Tab &tab,
                     uint32_t *out_n, size_t *in_used, uint32_t *lut = nullptr, FzCode *codes = nullptr)
{
    FzInflater<Tab> inf;
    FzCode ll, dd;
    inf.start(in, in_len, out, out_cap, tab);
    inf.bind_codes(codes ? &codes[0] : &ll, codes ? &codes[1] : &dd);
    inf.own_lut = lut;  // FZ_LUT_SIZE entries, or none
    while (inf.step()) {}
    return inf.finish(out_n, in_used);
}
