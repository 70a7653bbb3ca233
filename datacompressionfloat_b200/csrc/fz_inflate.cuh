// fz_inflate.cuh -- thread-serial raw-inflate (RFC 1951: stored / fixed / dynamic, any distance) of ONE
// deflate fragment.  Replaces what the reference gets from zlib's inflate() behind mzlib_inf
// (reference zip.c:262-284) for one payload -- or, on streams made by our encoder, for one
// sub-block (one GPU thread per sub-block; see fz_kernels_inflate.cu).
//
// `__host__ __device__`: tests/hostmodel runs this very source on the CPU against zlib streams.
//
// Decoding uses left-aligned canonical codes: for the next 15 stream bits, bit-reversed into a
// 15-bit number w, the code length is the first l with w < limit[l]; the symbol index is
// (w >> (15-l)) + delta[l] into the symbols sorted by (length, value).  limit/delta live in
// registers (all loops over l are fully unrolled); only the sorted symbol tables are in memory.
#pragma once
#include "fz_common.cuh"

// sorted-symbol tables of one decoding thread; STRIDE interleaves the threads of a warp in shared memory
template <int STRIDE>
struct FzInfTab {
    uint16_t *ll;  // 288 entries
    uint16_t *dd;  // 32 entries
    FZ_HD uint16_t &L(int i) const { return ll[i * STRIDE]; }
    FZ_HD uint16_t &D(int i) const { return dd[i * STRIDE]; }
};

struct FzBitReader {
    const uint32_t *w;   // aligned word pointer
    uint64_t acc;
    int nacc;            // valid bits in acc
    int64_t nwords;      // words that may still be loaded
    int64_t bits_left;   // bits of real input not yet consumed (goes negative on overrun)
    FZ_HD void init(const uint8_t *in, size_t in_len)
    {
        const uintptr_t a = (uintptr_t)in;
        const unsigned sk = (unsigned)(a & 3);
        w = (const uint32_t *)(a - sk);
        nwords = (int64_t)((in_len + sk + 3) / 4);
        bits_left = (int64_t)in_len * 8;
        acc = 0; nacc = 0;
        if (nwords > 0) { acc = (uint64_t)(*w++ >> (8 * sk)); nacc = 32 - 8 * (int)sk; nwords--; }
    }
    FZ_HD void refill()  // guarantees nacc >= 32 (zero bits past the end of the input)
    {
        if (nacc < 32) {
            uint32_t v = 0;
            if (nwords > 0) { v = *w++; nwords--; }
            acc |= (uint64_t)v << nacc;
            nacc += 32;
        }
    }
    FZ_HD uint32_t peek(int n) const { return (uint32_t)acc & ((1u << n) - 1); }
    FZ_HD void drop(int n) { acc >>= n; nacc -= n; bits_left -= n; }
    FZ_HD uint32_t get(int n) { const uint32_t v = peek(n); drop(n); return v; }  // n <= 16, after refill
    FZ_HD void align_byte() { const int k = (int)(bits_left & 7); drop(k); }
};

struct FzByteWriter {
    uint8_t *out;     // 4-byte aligned
    uint32_t op, cap;
    uint32_t ow;      // pending bytes of the current word
    FZ_HD void init(uint8_t *o, uint32_t c) { out = o; op = 0; cap = c; ow = 0; }
    FZ_HD void put(uint32_t c)
    {
        ow |= c << ((op & 3) * 8);
        op++;
        if ((op & 3) == 0) { *(uint32_t *)(out + op - 4) = ow; ow = 0; }
    }
    FZ_HD uint32_t back(uint32_t dist) const  // byte written `dist` positions ago (dist <= op)
    {
        const uint32_t p = op - dist;
        if ((p >> 2) == (op >> 2)) return (ow >> ((p & 3) * 8)) & 0xffu;  // still pending in ow
        return out[p];
    }
    FZ_HD void finish()
    {
        const uint32_t r = op & 3;
        for (uint32_t i = 0; i < r; i++) out[op - r + i] = (uint8_t)(ow >> (8 * i));
    }
};

#define FZ_INF_OK 0
#define FZ_INF_E_INPUT (-1)     // ran out of input / truncated
#define FZ_INF_E_DATA (-2)      // invalid block type, lengths, code or distance
#define FZ_INF_E_SPACE (-3)     // more output than out_cap
#define FZ_INF_E_HISTORY (-4)   // distance reaches before the start of this fragment

struct FzCodeRegs {
    uint32_t limit[16];  // left-aligned (15-bit) exclusive upper bound per length, [0] unused
    int32_t delta[16];   // sorted-index offset minus first code per length
};

// limit/delta from per-length counts; returns <0 if over-subscribed
FZ_HD int fz_code_regs(FzCodeRegs &r, const uint16_t *cnt /*[16]*/, uint16_t *offs /*[16] out: first index per length*/)
{
    uint32_t code = 0, idx = 0;
    int left = 1;
    r.limit[0] = 0; r.delta[0] = 0;
#pragma unroll
    for (int l = 1; l <= 15; l++) {
        code <<= 1;
        left <<= 1;
        const uint32_t c = cnt[l];
        left -= (int)c;
        offs[l] = (uint16_t)idx;
        r.delta[l] = (int32_t)idx - (int32_t)code;
        code += c;
        idx += c;
        r.limit[l] = code << (15 - l);
    }
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

// decode one symbol index (into the sorted table); returns length used (0 = invalid code)
FZ_HD int fz_decode_idx(const FzCodeRegs &r, uint32_t bits15, uint32_t &idx)
{
    const uint32_t w = fz_rev15(bits15);
#pragma unroll
    for (int l = 1; l <= 15; l++) {
        if (w < r.limit[l]) { idx = (uint32_t)((int32_t)(w >> (15 - l)) + r.delta[l]); return l; }
    }
    return 0;
}

// Inflate one fragment.  Stops after a BFINAL block, or at the end of the input on a block boundary.
//   *out_n   bytes produced
//   *in_used input bytes consumed (rounded up to whole bytes)
// `out` must be 4-byte aligned.  Reads whole aligned 32-bit words around [in, in+in_len).
template <class Tab>
FZ_HD int fz_inflate(const uint8_t *in, size_t in_len, uint8_t *out, uint32_t out_cap, const Tab &tab,
                     uint32_t *out_n, size_t *in_used)
{
    FzBitReader br;
    br.init(in, in_len);
    FzByteWriter bw;
    bw.init(out, out_cap);
    int rc = FZ_INF_OK;
    bool last = false;
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

    while (!last) {
        if (br.bits_left < 3) break;                       // nothing but padding left
        if (bw.op == out_cap && br.bits_left < 8) break;  // full output, only pad bits left
        br.refill();
        last = br.get(1) != 0;
        const uint32_t type = br.get(2);
        if (type == 0) {
            br.align_byte();
            br.refill();
            const uint32_t len = br.get(16);
            br.refill();
            const uint32_t nlen = br.get(16);
            if ((len ^ 0xFFFFu) != nlen) { rc = FZ_INF_E_DATA; break; }
            if (br.bits_left < (int64_t)len * 8) { rc = FZ_INF_E_INPUT; break; }
            if (bw.op + len > out_cap) { rc = FZ_INF_E_SPACE; break; }
            for (uint32_t i = 0; i < len; i++) { br.refill(); bw.put(br.get(8)); }
            continue;
        }
        if (type == 3) { rc = FZ_INF_E_DATA; break; }

        FzCodeRegs LL, DD;
        uint16_t cnt[16], offs[16];
        if (type == 1) {
            // fixed code: litlen lengths 8 (0-143), 9 (144-255), 7 (256-279), 8 (280-287); 30 distance codes of 5 bits
            for (int l = 0; l < 16; l++) cnt[l] = 0;
            cnt[7] = 24; cnt[8] = 152; cnt[9] = 112;
            fz_code_regs(LL, cnt, offs);
            for (int i = 0; i < 24; i++) tab.L(i) = (uint16_t)(256 + i);
            for (int i = 0; i < 144; i++) tab.L(24 + i) = (uint16_t)i;
            for (int i = 0; i < 8; i++) tab.L(168 + i) = (uint16_t)(280 + i);
            for (int i = 0; i < 112; i++) tab.L(176 + i) = (uint16_t)(144 + i);
            for (int l = 0; l < 16; l++) cnt[l] = 0;
            cnt[5] = 32;
            fz_code_regs(DD, cnt, offs);
            for (int i = 0; i < 32; i++) tab.D(i) = (uint16_t)i;
        } else {
            br.refill();
            const uint32_t hlit = br.get(5) + 257, hdist = br.get(5) + 1, hclen = br.get(4) + 4;
            if (hlit > 286 || hdist > 30) { rc = FZ_INF_E_DATA; break; }
            // code-length code: 19 symbols, <= 7 bits
            uint8_t cl[19];
            for (int i = 0; i < 19; i++) cl[i] = 0;
            for (uint32_t i = 0; i < hclen; i++) { br.refill(); cl[order[i]] = (uint8_t)br.get(3); }
            uint16_t ccnt[16], coffs[16];
            for (int l = 0; l < 16; l++) ccnt[l] = 0;
            for (int i = 0; i < 19; i++) ccnt[cl[i]]++;
            ccnt[0] = 0;
            FzCodeRegs CL;
            if (fz_code_regs(CL, ccnt, coffs) != 0) { rc = FZ_INF_E_DATA; break; }  // zlib requires a complete code here
            uint8_t clsym[19];
            for (int i = 0; i < 19; i++) if (cl[i]) clsym[coffs[cl[i]]++] = (uint8_t)i;

            // two passes over the code-length data: count per length, then place the sorted symbols
            const FzBitReader mark = br;
            uint16_t cnt_d[16], offs_d[16];
            for (int pass = 0; pass < 2 && rc == FZ_INF_OK; pass++) {
                if (pass == 0) { for (int l = 0; l < 16; l++) { cnt[l] = 0; cnt_d[l] = 0; } }
                else {
                    cnt[0] = 0; cnt_d[0] = 0;
                    const int e1 = fz_code_regs(LL, cnt, offs);
                    const int e2 = fz_code_regs(DD, cnt_d, offs_d);
                    if (e1 < 0 || e2 < 0) { rc = FZ_INF_E_DATA; break; }  // over-subscribed
                    br = mark;
                }
                uint32_t i = 0, prev = 0;
                const uint32_t total = hlit + hdist;
                while (i < total) {
                    br.refill();
                    uint32_t idx;
                    const int l = fz_decode_idx(CL, br.peek(15), idx);
                    if (l == 0 || l > 7) { rc = FZ_INF_E_DATA; break; }
                    br.drop(l);
                    const uint32_t s = clsym[idx];
                    uint32_t rep = 1, val = s;
                    if (s == 16) { if (i == 0) { rc = FZ_INF_E_DATA; break; } val = prev; rep = 3 + br.get(2); }
                    else if (s == 17) { val = 0; rep = 3 + br.get(3); }
                    else if (s == 18) { val = 0; rep = 11 + br.get(7); }
                    if (i + rep > total) { rc = FZ_INF_E_DATA; break; }
                    prev = val;
                    if (pass == 0) {
                        for (uint32_t k = 0; k < rep; k++, i++) { if (i < hlit) cnt[val]++; else cnt_d[val]++; }
                    } else if (val) {
                        for (uint32_t k = 0; k < rep; k++, i++) {
                            if (i < hlit) tab.L(offs[val]++) = (uint16_t)i; else tab.D(offs_d[val]++) = (uint16_t)(i - hlit);
                        }
                    } else i += rep;
                }
                if (br.bits_left < 0) rc = FZ_INF_E_INPUT;
            }
            if (rc != FZ_INF_OK) break;
        }

        // ---- block body
        for (;;) {
            br.refill();
            if (br.bits_left < 0) { rc = FZ_INF_E_INPUT; break; }
            uint32_t idx;
            int l = fz_decode_idx(LL, br.peek(15), idx);
            if (l == 0) { rc = FZ_INF_E_DATA; break; }
            br.drop(l);
            uint32_t sym = tab.L((int)idx);
            if (sym < 256) {
                if (bw.op >= out_cap) { rc = FZ_INF_E_SPACE; break; }
                bw.put(sym);
                continue;
            }
            if (sym == FZ_EOB) break;
            sym -= 257;
            if (sym >= 29) { rc = FZ_INF_E_DATA; break; }
            br.refill();
            const uint32_t len = fz_len_base(sym) + br.get((int)fz_len_extra_bits(sym));
            l = fz_decode_idx(DD, br.peek(15), idx);
            if (l == 0) { rc = FZ_INF_E_DATA; break; }
            br.drop(l);
            const uint32_t ds = tab.D((int)idx);
            if (ds >= 30) { rc = FZ_INF_E_DATA; break; }
            br.refill();
            const uint32_t dist = fz_dist_base(ds) + br.get((int)fz_dist_extra_bits(ds));
            if (dist > bw.op) { rc = FZ_INF_E_HISTORY; break; }
            if (bw.op + len > out_cap) { rc = FZ_INF_E_SPACE; break; }
            if (dist == 1) {
                const uint32_t c = bw.back(1);
                for (uint32_t i = 0; i < len; i++) bw.put(c);
            } else {
                for (uint32_t i = 0; i < len; i++) bw.put(bw.back(dist));
            }
        }
        if (rc != FZ_INF_OK) break;
        if (br.bits_left < 0) { rc = FZ_INF_E_INPUT; break; }
    }
    bw.finish();
    if (rc == FZ_INF_OK && br.bits_left < 0) rc = FZ_INF_E_INPUT;
    *out_n = bw.op;
    const int64_t used_bits = (int64_t)in_len * 8 - br.bits_left;
    *in_used = (size_t)((used_bits + 7) / 8);
    return rc;
}
