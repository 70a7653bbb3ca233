// host_model.cpp -- compiles the product's __host__ __device__ codec headers for the CPU so the bit-level
// logic of the GPU encoder / inflater can be checked against zlib without a GPU.
// TEST INFRASTRUCTURE ONLY (built by tests/hostmodel/build.py into tests/hostmodel/libhostmodel.so);
// the product never loads it.
#define FZ_SY_STATS 1
#include <cstdio>
#include <cstdlib>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../datacompressionfloat_b200/csrc/fz_deflate_enc.cuh"
#include "../../datacompressionfloat_b200/csrc/fz_enc2.cuh"
#include <functional>
#include <thread>
#include "../../datacompressionfloat_b200/csrc/fz_inflate.cuh"
#include "../../datacompressionfloat_b200/csrc/fz_blockpar.cuh"

namespace {
struct HostLoad16 {
    const uint8_t *p;
    FzVec16 operator()(uint32_t i) const { FzVec16 v; memcpy(v.w, p + i, 16); return v; }
};
struct HostLoadByte {
    const uint8_t *p;
    uint32_t operator()(uint32_t i) const { return p[i]; }
};

// what the gather kernel synthesises for a stored sub-block: two stored blocks + empty stored block
size_t put_stored(uint8_t *o, const uint8_t *in, uint32_t n)
{
    uint8_t *b = o;
    auto hdr = [&](uint32_t len) {
        *o++ = 0x00;
        *o++ = (uint8_t)(len & 0xff); *o++ = (uint8_t)(len >> 8);
        *o++ = (uint8_t)(~len & 0xff); *o++ = (uint8_t)((~len >> 8) & 0xff);
    };
    if (n < 2) { hdr(n); memcpy(o, in, n); o += n; }
    else {
        uint32_t first = 0xFFFFFFFFu;
        for (uint32_t i = 0; i + 4 <= n; i++)
            if (in[i] == 0 && in[i + 1] == 0 && in[i + 2] == 0xFF && in[i + 3] == 0xFF) { first = i; break; }
        const uint32_t s = fz_stored_split(n, first);
        hdr(s); memcpy(o, in, s); o += s;
        hdr(n - s); memcpy(o, in + s, n - s); o += n - s;
    }
    *o++ = 0x00; *o++ = 0x00; *o++ = 0x00; *o++ = 0xFF; *o++ = 0xFF;
    return (size_t)(o - b);
}
}  // namespace

extern "C" {

// ---------------------------------------------------------------------------------------------------------------------
// the encoder (fz_enc2.cuh): the device source run by 32 host threads in lock step, next to a plain sequential
// restatement of the token rule and the bit layout -- the two must produce the same bytes
}  // extern "C"
namespace {
void run_warp(const std::function<void(const FzWarp &)> &fn)
{
    FzWarpShared sh;
    pthread_barrier_init(&sh.bar, nullptr, 32);
    std::vector<std::thread> th;
    for (int l = 0; l < 32; l++) th.emplace_back([&, l] { FzWarp w{l, &sh}; fn(w); });
    for (auto &t : th) t.join();
    pthread_barrier_destroy(&sh.bar);
}

template <class Lit, class Match>
void seq_tokens(const uint8_t *p, uint32_t n, Lit &&lit, Match &&match)
{
    // the quad rule of fz_enc2.cuh, one quad after the other
    uint32_t m = 0, runb = 0;
    bool prevE = false, prevE2 = false;
    auto rest = [&]() {
        if (m >= FZ_MIN_MATCH) match(m); else for (uint32_t i = 0; i < m; i++) lit(runb);
        m = 0;
    };
    for (uint32_t q0 = 0; q0 < n; q0 += 4) {
        const uint32_t nb = n - q0 < 4 ? n - q0 : 4;
        bool E = nb == 4 && q0 > 0;
        for (uint32_t i = 0; E && i < 4; i++) E = p[q0 + i] == p[q0 + i - 1];
        if (E && prevE && (FZ_E2_LEAD_QUADS < 2 || prevE2)) {
            runb = p[q0];
            m += 4;
            if (m >= FZ_MAX_MATCH) { match(FZ_MAX_MATCH); m -= FZ_MAX_MATCH; }
        } else {
            if (m) rest();
            for (uint32_t i = 0; i < nb; i++) lit(p[q0 + i]);
        }
        prevE2 = prevE;
        prevE = E;
    }
    if (m) rest();
}

struct SeqBits {
    std::vector<uint8_t> b;
    uint64_t nbits = 0;
    void put(uint32_t v, uint32_t n)
    {
        for (uint32_t i = 0; i < n; i++, nbits++) {
            if ((nbits & 7) == 0) b.push_back(0);
            b.back() |= (uint8_t)(((v >> i) & 1u) << (nbits & 7));
        }
    }
};

// the fragment of one sub-block, or empty = stored
std::vector<uint8_t> seq_emit(const FzGroupCode *gc, const uint8_t *p, uint32_t n)
{
    SeqBits sb;
    for (uint32_t i = 0; i < gc->hdr_nbits; i++) sb.put((gc->hdr[i >> 5] >> (i & 31)) & 1u, 1);
    seq_tokens(p, n,
               [&](uint32_t c) { sb.put(gc->cl[c] & 0xffffu, gc->cl[c] >> 16); },
               [&](uint32_t len) {
                   uint32_t lc, eb, ev;
                   fz_len_code(len, lc, eb, ev);
                   sb.put(gc->cl[257 + lc] & 0xffffu, gc->cl[257 + lc] >> 16);
                   sb.put(ev, eb);
                   sb.put(0, 1);
               });
    sb.put(gc->cl[FZ_EOB] & 0xffffu, gc->cl[FZ_EOB] >> 16);
    sb.put(0, 3);
    while (sb.nbits & 7) sb.put(0, 1);
    sb.put(0xFFFF0000u, 32);
    const uint32_t limit = fz_stored_size(n) - (n >> FZ_MIN_GAIN_SHIFT);
    if (sb.b.size() >= limit) return {};
    for (size_t i = 0; i + 4 < sb.b.size(); i++)
        if (sb.b[i] == 0 && sb.b[i + 1] == 0 && sb.b[i + 2] == 0xFF && sb.b[i + 3] == 0xFF) return {};
    return sb.b;
}
}  // namespace
extern "C" {

// mode 0: the device source under the thread model; mode 1: the sequential restatement.  skip: pass the sample's two
// most frequent bytes to the histogram (what the kernel does) or not.
int hm_hist_sample = 0;
void hm_set_hist_sample(int k) { hm_hist_sample = k; }
uint64_t hm_encode_stream_v2(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t cap, int mode, int skip, uint64_t *nstored)
{
    uint64_t o = 0, ns = 0;
    std::vector<uint8_t> pad(FZ_SUB + 64);
    std::vector<uint32_t> slot(FZ_SLOT_STRIDE / 4 + 8);
    FzEncState *st = (FzEncState *)malloc(sizeof(FzEncState));
    FzGroupCode *gc = (FzGroupCode *)malloc(sizeof(FzGroupCode));
    const uint64_t gbytes = (uint64_t)FZ_SUB * FZ_CODE_SUBS;
    for (uint64_t g0 = 0; g0 < n; g0 += gbytes) {
        const uint64_t gn = (n - g0) < gbytes ? (n - g0) : gbytes;
        const uint32_t nsub = (uint32_t)((gn + FZ_SUB - 1) / FZ_SUB);
        memset(st, 0xCD, sizeof(FzEncState));
        memset(st->hist, 0, sizeof(st->hist));
        for (uint32_t k = 0; k < nsub; k++) {
            const uint32_t m = (uint32_t)((gn - (uint64_t)k * FZ_SUB) < FZ_SUB ? (gn - (uint64_t)k * FZ_SUB) : FZ_SUB);
            memset(pad.data(), 0x5A, pad.size());   // what lies behind a ragged sub-block is arbitrary
            memcpy(pad.data(), in + g0 + (uint64_t)k * FZ_SUB, m);
            uint32_t h[288];
            memset(h, 0, sizeof(h));
            if (mode == 0) {
                uint32_t s1 = 0x100, s2 = 0x100;
                if (skip) {   // two most frequent bytes of the first 2 KiB
                    uint32_t cnt[256] = {0};
                    for (uint32_t i = 0; i < m && i < 2048; i++) cnt[pad[i]]++;
                    s1 = 0;
                    for (uint32_t c = 1; c < 256; c++) if (cnt[c] > cnt[s1]) s1 = c;
                    s2 = s1 == 0 ? 1 : 0;
                    for (uint32_t c = 0; c < 256; c++) if (c != s1 && cnt[c] > cnt[s2]) s2 = c;
                }
                HostLoad16 ld{pad.data()};
                run_warp([&](const FzWarp &w) { fz_hist2_subblock(w, h, ld, m, s1, s2); });
            } else {
                seq_tokens(pad.data(), m, [&](uint32_t c) { h[c]++; },
                           [&](uint32_t len) { uint32_t lc, eb, ev; fz_len_code(len, lc, eb, ev); h[257 + lc]++; });
            }
            if (hm_hist_sample > 1) {
                if (k % hm_hist_sample != 0) continue;
                for (int i = 0; i < 288; i++) h[i] *= hm_hist_sample;
            }
            for (int i = 0; i < 288; i++) st->hist[i] += h[i];
        }
        if (hm_hist_sample > 1) {
            for (int i = 0; i < 256; i++) st->hist[i] += 1;
            for (int i = 257; i < 286; i++) if (st->hist[i] == 0) st->hist[i] = 1;
        }
        st->hist[FZ_EOB] = nsub;
        memset(gc, 0xEE, sizeof(FzGroupCode));
        fz_build_group_code(st, (uint32_t)gn, nsub, gc, 0, hm_hist_sample > 1 ? (uint32_t)hm_hist_sample : 0u);
        for (uint32_t k = 0; k < nsub; k++) {
            const uint32_t m = (uint32_t)((gn - (uint64_t)k * FZ_SUB) < FZ_SUB ? (gn - (uint64_t)k * FZ_SUB) : FZ_SUB);
            memset(pad.data(), 0x5A, pad.size());
            memcpy(pad.data(), in + g0 + (uint64_t)k * FZ_SUB, m);
            uint32_t r = fz_stored_size(m) | FZ_SIZE_STORED_FLAG;
            if (!gc->stored) {
                if (mode == 0) {
                    for (auto &w : slot) w = 0xDEADBEEFu;
                    std::vector<uint32_t> ring(FZ_E2_RING_WORDS, 0xABABABABu), tt(256, 0xCDCDCDCDu);
                    HostLoad16 ld{pad.data()};
                    uint32_t res[32];
                    run_warp([&](const FzWarp &w) {
                        res[w.lane] = fz_emit2_subblock(w, gc->cl, gc->hdr, gc->hdr_nbits, ring.data(), tt.data(), ld, m, slot.data());
                    });
                    for (int l = 1; l < 32; l++) if (res[l] != res[0]) return (uint64_t)-2;
                    r = res[0];
                } else {
                    const std::vector<uint8_t> f = seq_emit(gc, pad.data(), m);
                    if (!f.empty()) { r = (uint32_t)f.size(); memcpy(slot.data(), f.data(), f.size()); }
                }
            }
            if (r & FZ_SIZE_STORED_FLAG) {
                if (o + fz_stored_size(m) > cap) return (uint64_t)-1;
                o += put_stored(out + o, in + g0 + (uint64_t)k * FZ_SUB, m);
                ns++;
            } else {
                if (o + r > cap) return (uint64_t)-1;
                memcpy(out + o, slot.data(), r);
                o += r;
            }
        }
    }
    free(st); free(gc);
    if (nstored) *nstored = ns;
    return o;
}

int hm_inflate(const uint8_t *in, uint64_t in_len, uint8_t *out, uint32_t out_cap, uint32_t *out_n, uint64_t *in_used)
{
    // word-aligned, padded copies (the device reads whole aligned words; out must be 4-byte aligned)
    std::vector<uint32_t> ibuf((in_len + 16) / 4 + 1, 0);
    std::vector<uint32_t> obuf(out_cap / 4 + 2, 0);
    uint8_t *ip = (uint8_t *)ibuf.data() + 1;  // deliberately misaligned input
    memcpy(ip, in, in_len);
    uint16_t ll[288], dd[32], cnt[32];
    FzInfTab<1> tab{ll, dd, cnt};
    size_t used = 0;
    // both paths of the general inflater must agree: canonical-only decode and table-driven decode
    int rc = fz_inflate(ip, (size_t)in_len, (uint8_t *)obuf.data(), out_cap, tab, out_n, &used);
    std::vector<uint32_t> lut(FZ_LUT_SIZE), obuf2(out_cap / 4 + 2, 0);
    uint32_t out_n2 = 0;
    size_t used2 = 0;
    int rc2 = fz_inflate(ip, (size_t)in_len, (uint8_t *)obuf2.data(), out_cap, tab, &out_n2, &used2, lut.data());
    if (rc2 != rc || out_n2 != *out_n || used2 != used || memcmp(obuf.data(), obuf2.data(), *out_n) != 0) return -99;
    memcpy(out, obuf.data(), *out_n);
    *in_used = used;
    return rc;
}

// test hooks: the false-marker check of the emit stage (what FzRingOut::flush does vector by vector), and the stored form
uint32_t hm_check_marker(const uint8_t *frag, uint32_t total_bytes)
{
    std::vector<uint32_t> w((total_bytes + 15) / 16 * 4 + 4, 0);
    memcpy(w.data(), frag, total_bytes);
    uint32_t pw = 0x55555555u, bad = 0;
    for (uint32_t vi = 0; vi < (total_bytes + 15) / 16; vi++) {
        const uint32_t *x = &w[vi * 4];
        bad |= fz_marker_vec(pw, x[0], x[1], x[2], x[3], vi, total_bytes) ? 1u : 0u;
        pw = x[3];
    }
    return bad;
}
uint32_t hm_put_stored(const uint8_t *in, uint32_t n, uint8_t *out) { return (uint32_t)put_stored(out, in, n); }

// block-parallel inflate of a zlib-made stream, exactly the four stages the GPU runs (candidates, measure,
// chain, write); returns 0, a negative chain code (the GPU would take the serial path), or 100+ on a write error
void hm_sy_stats(uint64_t *o) { o[0] = fz_sy_stat_tiles; o[1] = fz_sy_stat_rounds; o[2] = fz_sy_stat_redos; o[3] = fz_sy_stat_redo_lanes; }
void hm_sy_stats_reset(void) { fz_sy_stat_tiles = fz_sy_stat_rounds = fz_sy_stat_redos = fz_sy_stat_redo_lanes = 0; }
uint32_t hm_tile_pool_cap = 4096, hm_table_blocks = 0;
void hm_set_tile_pool_cap(uint32_t n) { hm_tile_pool_cap = n; }
uint32_t hm_get_table_blocks(void) { return hm_table_blocks; }
int hm_blockpar_sync = 1;      // 1: warp-synchronising block decoder (what the GPU runs), 0: serial per-block functions
void hm_set_blockpar_sync(int on) { hm_blockpar_sync = on; }
uint64_t hm_precheck_passes = 0;
uint64_t hm_last_precheck_passes(void) { return hm_precheck_passes; }
// Builds the lookup table of a literal/length code given by its 288 code lengths both ways -- one canonical search per
// entry (fz_lut_entry_bits) and the warp's interval walk + in-table packing (fz_lut_fill_lane / fz_lut_pack) -- and returns
// the number of entries that differ (-1: the lengths are over-subscribed).
int hm_lut_compare(const uint8_t *lens)
{
    uint16_t ll[288], dd[32], cnt[32];
    FzInfTab<1> tab{ll, dd, cnt};
    for (int l = 0; l < 32; l++) cnt[l] = 0;
    for (int i = 0; i < 288; i++) if (lens[i]) cnt[lens[i]]++;
    FzCode LL;
    auto rd = [&](int l) -> uint32_t { return cnt[l]; };
    auto wr = [&](int l, uint32_t v) { cnt[l] = (uint16_t)v; };
    if (fz_code_build(LL, rd, wr) < 0) return -1;
    for (int i = 0; i < 288; i++) if (lens[i]) ll[cnt[lens[i]]++] = (uint16_t)i;
    std::vector<uint32_t> a(FZ_LUT_SIZE), b(FZ_LUT_SIZE, 0xDEADBEEFu);
    for (uint32_t e = 0; e < FZ_LUT_SIZE; e++) a[e] = fz_lut_entry(LL, tab, e);
    for (int lane = 0; lane < 32; lane++) fz_lut_fill_lane<FZ_LUT_BITS>(b.data(), LL, tab, lane);
    for (uint32_t bt = FZ_LUT_SIZE / 32u - 1u; bt >= 1u; bt--)
        for (uint32_t lane = 0; lane < 32; lane++) b[bt * 32u + lane] = fz_lut_pack<FZ_LUT_BITS>(b.data(), bt * 32u + lane);
    uint32_t first[32];
    for (uint32_t lane = 0; lane < 32; lane++) first[lane] = fz_lut_pack<FZ_LUT_BITS>(b.data(), lane);
    for (uint32_t lane = 0; lane < 32; lane++) b[lane] = first[lane];
    int diff = 0;
    for (uint32_t e = 0; e < FZ_LUT_SIZE; e++) diff += a[e] != b[e];
    return diff;
}

int hm_blockpar_hint = 1;   // pass the distance to the next candidate to the measure pass (what the kernel does)
void hm_set_blockpar_hint(int on) { hm_blockpar_hint = on; }
uint64_t hm_quick_passes = 0;  // positions that passed fz_block_quick_test in the last hm_inflate_blockpar call
uint64_t hm_last_quick_passes(void) { return hm_quick_passes; }

int hm_inflate_blockpar(const uint8_t *in_, uint32_t in_len, uint8_t *out_, uint32_t n_out, uint32_t *ncand_out, uint32_t *nfalse_out)
{
    std::vector<uint32_t> ibuf(in_len / 4 + 8, 0);
    uint8_t *in = (uint8_t *)ibuf.data() + 3;  // misaligned on purpose
    memcpy(in, in_, in_len);
    std::vector<uint32_t> obuf(n_out / 4 + 4, 0xA5A5A5A5u);
    uint8_t *out = (uint8_t *)obuf.data();
    uint16_t ll[288], dd[32], cnt[32];
    FzInfTab<1> tab{ll, dd, cnt};
    std::vector<uint32_t> lut(FZ_LUT_SIZE);
    auto byte_at = [&](uint64_t by) -> uint64_t { return by < in_len ? in[by] : 0; };
    auto bits64 = [&](uint64_t bit) -> uint64_t {  // stream bits [bit, bit + 64), zero past the end
        const uint64_t by = bit >> 3;
        const unsigned sh = (unsigned)(bit & 7);
        uint64_t v = 0;
        for (int k = 0; k < 8; k++) v |= byte_at(by + k) << (8 * k);
        return sh ? (v >> sh) | (byte_at(by + 8) << (64 - sh)) : v;
    };
    std::vector<uint32_t> cand;
    hm_quick_passes = 0;
    hm_precheck_passes = 0;
    hm_table_blocks = 0;
    const uint64_t total_bits = (uint64_t)in_len * 8;
    for (uint64_t bit = 0; bit + 17 <= total_bits; bit++) {
        if (!fz_block_quick_test(bits64(bit), bits64(bit + 64))) continue;
        hm_quick_passes++;
        {   // the cheap second stage must never reject what the full parse accepts (and is run first on the GPU)
            uint8_t cl[128 + 8];
            const FzClLut<1> cl_lut{cl};
            const bool pre = fz_block_precheck(in, in_len, bit, cl_lut);
            const bool full = fz_block_candidate(in, in_len, bit, tab);
            if (pre) hm_precheck_passes++;
            if (full && !pre) return -1000;
            if (!pre) continue;
        }
        if (fz_block_candidate(in, in_len, bit, tab)) cand.push_back((uint32_t)bit);
    }
    const uint32_t ncand = (uint32_t)cand.size();
    const uint32_t cap = ncand + 64;
    cand.resize(cap);
    std::vector<FzBlockInfo> info(cap);
    std::vector<FzSyncState> sy(1);
    // tile records of the measure pass; hm_tile_pool_cap = 0 exercises the "pool exhausted" path (write pass searches again)
    std::vector<FzTileRec> recs(hm_tile_pool_cap ? hm_tile_pool_cap : 1);
    uint32_t pool_cursor = 0;
    FzTilePool pool{recs.data(), hm_tile_pool_cap, &pool_cursor};
    std::vector<uint32_t> first_rec(cap, FZ_TILE_NONE);
    for (uint32_t i = 0; i < ncand; i++) {
        if (hm_blockpar_sync) {
            // the hint the kernel gives: bits to the next candidate header (the candidates are in ascending order here)
            const uint32_t hint = hm_blockpar_hint ? (i + 1 < ncand ? cand[i + 1] - cand[i] : (uint32_t)(total_bits - cand[i])) : 0u;
            fz_sy_block<false>(sy.data(), in, in_len, cand[i], nullptr, 0, -1, 0, &info[i], nullptr, 0, &pool, &first_rec[i], hint);
            FzBlockInfo ref;   // the serial measure must agree exactly
            fz_block_measure(in, in_len, cand[i], tab, lut.data(), &ref);
            const bool ref_usable = (ref.flags & FZ_BLK_OK) && !(ref.flags & FZ_BLK_NON_RLE);
            const bool got_usable = (info[i].flags & FZ_BLK_OK) && !(info[i].flags & FZ_BLK_NON_RLE);
            if (ref_usable != got_usable) return 1000 + (int)i;
            if (ref_usable && (ref.end_bit != info[i].end_bit || ref.out_len != info[i].out_len || ref.flags != info[i].flags ||
                               ((ref.flags & FZ_BLK_HAS_LITERAL) && ref.last != info[i].last)))
                return 2000 + (int)i;
        } else fz_block_measure(in, in_len, cand[i], tab, lut.data(), &info[i]);
    }
    std::vector<uint32_t> off(cap);
    std::vector<int> prev(cap);
    std::vector<FzStoredItem> stored(64);
    uint32_t nst = 0, nblocks = 0;
    const int rc = fz_chain_resolve(in, in_len, n_out, cand.data(), info.data(), ncand, cap, &nblocks, off.data(), prev.data(),
                                    stored.data(), 64, &nst, tab);
    *ncand_out = ncand;
    uint32_t nfalse = 0;
    for (uint32_t i = 0; i < ncand; i++) nfalse += off[i] == 0xFFFFFFFFu;
    *nfalse_out = nfalse;
    if (rc != 0) return rc;
    for (uint32_t i = 0; i < nblocks; i++) {
        if (off[i] == 0xFFFFFFFFu) continue;
        if (hm_blockpar_sync) {
            bool ok = false;
            if (first_rec[i] != FZ_TILE_NONE) {
                fz_sy_block_from_table(sy.data(), in, in_len, cand[i], out + off[i], info[i].out_len, prev[i], info[i].end_bit, pool,
                                       first_rec[i], &ok, 0);
                hm_table_blocks++;
            } else
                fz_sy_block<true>(sy.data(), in, in_len, cand[i], out + off[i], info[i].out_len, prev[i], info[i].end_bit, nullptr, &ok, 0);
            if (!ok) return 100 + (int)i;
        } else if (!fz_block_write(in, in_len, cand[i], tab, lut.data(), out + off[i], info[i].out_len, prev[i], info[i].end_bit)) return 100 + (int)i;
    }
    for (uint32_t k = 0; k < nst; k++) memcpy(out + stored[k].out_off, in + stored[k].src_byte, stored[k].len);
    memcpy(out_, out, n_out);
    return 0;
}

uint32_t hm_sub_bytes(void) { return FZ_SUB; }
uint32_t hm_group_subs(void) { return FZ_GROUP_SUBS; }
uint32_t hm_enc_state_bytes(void) { return (uint32_t)sizeof(FzEncState); }

}  // extern "C"
