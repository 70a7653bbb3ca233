// fz_deflate_enc.cuh -- the Huffman code and dynamic-block header of a CODE GROUP, built by one warp.
//
// The encoder replaces what the reference gets from zlib's deflate(Z_RLE, level 6) behind mzlib_def (reference
// zip.c:164-196, constants constant.h:22-24): distance-1 run matches of 3..258 bytes, dynamic-Huffman blocks, an empty
// stored block (the sync-flush marker) behind every sub-block so that the next one starts byte aligned; BFINAL is
// never set (reference decoder requirement).  The bytes are valid raw deflate, not zlib's bytes.
//
// One Huffman code serves FZ_CODE_SUBS consecutive sub-blocks of a stream (2 MiB of a plane): the token histogram of
// the group is collected first (fz_enc2.cuh), the code and the block header are built once per group here
// (fz_build_group_code), and every sub-block is then emitted as its own dynamic block carrying that same header
// (fz_enc2.cuh).  The GPU inflater exploits this: the four warps of a CTA decode the 128 sub-blocks of a group with
// ONE shared lookup table (it verifies that the headers really are identical).
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
    uint32_t n_low;          // stand-in symbols (frequency 1 in a group with unsampled sub-blocks): outside the tree
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

FZ_HD int fz_ctz32(uint32_t v)  // v != 0
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}

#define FZ_BYTE_OF(v, k) (((v).w[(k) >> 2] >> (((k) & 3) * 8)) & 0xffu)

// -------------------------------------------------------------------------------------------------
// Huffman construction phases (literal/length alphabet).
// -------------------------------------------------------------------------------------------------
#define FZ_SYMS_PER_LANE 9  // 32 * 9 = 288

FZ_HD void fz_ph_zero_len(FzEncState *st, int lane)
{
    for (int i = lane; i < 288; i += 32) st->len[i] = 0;
}

// `low` = the frequency that marks a stand-in (1 when the group has unsampled sub-blocks, 0 = there are none): such
// symbols do not take part in the sort or the tree -- they get the longest code afterwards (fz_ph_standins)
FZ_HD void fz_ph_count_active(FzEncState *st, uint32_t low, int lane)
{
    uint32_t c = 0, cl = 0;
    for (int k = 0; k < FZ_SYMS_PER_LANE; k++) {
        const int s = lane * FZ_SYMS_PER_LANE + k;
        const uint32_t f = st->hist[s];
        c += f > low ? 1u : 0u;
        cl += (f != 0 && f <= low) ? 1u : 0u;
    }
    st->lane_cnt[lane] = c;
    st->lane_bits[lane] = cl;
}

FZ_HD void fz_ph_compact(FzEncState *st, uint32_t low, int lane)
{
    uint32_t off = 0, total = 0, nlow = 0;
    for (int l = 0; l < 32; l++) { const uint32_t c = st->lane_cnt[l]; if (l < lane) off += c; total += c; nlow += st->lane_bits[l]; }
    for (int k = 0; k < FZ_SYMS_PER_LANE; k++) {
        const int s = lane * FZ_SYMS_PER_LANE + k;
        const uint32_t f = st->hist[s];
        if (f > low) st->keys[off++] = (f << 9) | (uint32_t)s;
        else if (f != 0) st->len[s] = 15;   // a stand-in: the longest code deflate allows (fz_ph_standins may shorten a few)
    }
    // pad to the next power of two with +inf keys for the bitonic network
    uint32_t np2 = 32;
    while (np2 < total) np2 <<= 1;
    for (uint32_t i = total + lane; i < np2; i += 32) st->keys[i] = 0xFFFFFFFFu;
    if (lane == 0) { st->n_active = total; st->n_low = nlow; }
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

// serial (lane 0): make room in the code for the stand-ins.  On entry keys[0 .. n_active) are the code lengths of the
// symbols the sample saw (a complete code of at most FZ_MAX_CODE_BITS bits, ascending frequency = descending length),
// len[s] = 15 for every stand-in s, num_codes[] counts the former.  The Kraft sum K (units of 2^-15) is then over by
// one unit per stand-in: lengthen the least frequent seen symbols that may still grow (never past FZ_MAX_CODE_BITS: a
// seen symbol stays one table lookup in the inflater), then hand back what that overshot by shortening whatever fits --
// seen symbols, most frequent first, then stand-ins.  (Stand-ins used to go through the sort and the tree like everybody
// else: 286 symbols instead of 20..60, and a 12-bit limit for all of them cost 1..2 % in size.)
FZ_HD void fz_ph_standins(FzEncState *st, int lane)
{
    if (lane != 0 || st->n_low == 0) return;
    uint32_t *A = st->keys;
    const int n = (int)st->n_active;
    const uint32_t full = 1u << 15, lim = FZ_MAX_CODE_BITS;
    uint32_t K = st->n_low;
    for (int i = 0; i < n; i++) K += 1u << (15u - A[i]);
    while (K > full) {
        bool moved = false;
        for (int i = 0; i < n && K > full; i++)
            if (A[i] < lim) { K -= 1u << (14u - A[i]); A[i]++; moved = true; }
        if (!moved) break;   // (cannot happen: 286 symbols at their limits fill 7 % of the code space)
    }
    uint32_t D = K <= full ? full - K : 0u;
    while (D) {
        bool moved = false;
        for (int i = n - 1; i >= 0 && D; i--) {
            const uint32_t gain = 1u << (15u - A[i]);
            if (A[i] > 1u && gain <= D) { A[i]--; D -= gain; moved = true; }
        }
        for (int s2 = 0; s2 < FZ_NUM_LL && D; s2++) {
            if (st->hist[s2] != 1u || st->len[s2] == 0) continue;   // stand-ins only (len is still 0 for the seen symbols)
            const uint32_t l = st->len[s2], gain = 1u << (15u - l);
            if (l > 1u && gain <= D) { st->len[s2] = (uint8_t)(l - 1u); D -= gain; moved = true; }
        }
        if (!moved) break;   // (cannot happen: D is a multiple of the step of the longest code present)
    }
    for (int l = 0; l < 32; l++) st->num_codes[l] = 0;
    for (int i = 0; i < n; i++) st->num_codes[A[i]]++;
    for (int s2 = 0; s2 < FZ_NUM_LL; s2++)
        if (st->hist[s2] == 1u && st->len[s2] != 0) st->num_codes[st->len[s2]]++;
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
//   sample: 0 = the histogram counts every sub-block; k > 1 = it counts every k-th sub-block k times over, and symbols of
//   frequency 1 are stand-ins for "may occur in the sub-blocks that were not sampled"
template <int DUMMY = 0>
FZ_HD void fz_build_group_code(FzEncState *st, uint32_t group_bytes, uint32_t nsub, FzGroupCode *out, int lane, uint32_t sample = 0)
{
    (void)lane;
    const uint32_t low = sample > 1u ? 1u : 0u;
    FZ_PHASE(fz_ph_zero_len(st, lane));
    FZ_PHASE(fz_ph_count_active(st, low, lane));
    FZ_PHASE(fz_ph_compact(st, low, lane));
    {
        uint32_t np2 = 32;
        while (np2 < st->n_active) np2 <<= 1;
        for (uint32_t k = 2; k <= np2; k <<= 1)
            for (uint32_t j = k >> 1; j > 0; j >>= 1) FZ_PHASE(fz_ph_bitonic(st->keys, np2, k, j, lane));
    }
    FZ_PHASE(fz_ph_lengths(st->keys, st->ssym, st->num_codes, (int)st->n_active, FZ_MAX_CODE_BITS, lane));
    FZ_PHASE(fz_ph_standins(st, lane));
    FZ_PHASE(fz_ph_scatter_len(st, lane));
    FZ_PHASE(fz_ph_rank_count(st, lane));
    FZ_PHASE(fz_ph_rank_scan(st, lane));
    FZ_PHASE(fz_ph_assign_codes(st, lane));
    FZ_PHASE(fz_ph_seq(st, lane));
    FZ_PHASE(fz_ph_nzmask(st, lane));
    FZ_PHASE(fz_ph_cl_tokens(st, lane));
    FZ_PHASE(fz_ph_header(st, lane));
    FZ_PHASE(fz_ph_cost_partial(st, lane));
    // group decision: all sub-blocks dynamic vs all stored (the per-sub-block decision is exact, in fz_emit2_subblock)
    uint64_t bits = (uint64_t)nsub * (st->hdr_nbits + 3 + 4 + 32);  // header + empty stored block (avg pad 4) per sub-block
    uint64_t payload = 0;
    for (int l = 0; l < 32; l++) payload += ((uint64_t)st->lane_bits[l] << 32) | st->lane_cnt[l];
    // a sampled histogram stands for ceil(nsub / sample) * sample sub-blocks: in a ragged group (the end of a stream, small
    // chunks) that is more than the nsub it has, and the estimate would call compressible groups incompressible
    if (sample > 1u) payload = payload * nsub / ((uint64_t)((nsub + sample - 1u) / sample) * sample);
    bits += payload;
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

