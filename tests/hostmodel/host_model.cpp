// host_model.cpp -- compiles the product's __host__ __device__ codec headers for the CPU so the bit-level
// logic of the GPU encoder / inflater can be checked against zlib without a GPU.
// TEST INFRASTRUCTURE ONLY (built by tests/hostmodel/build.py into tests/hostmodel/libhostmodel.so);
// the product never loads it.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../datacompressionfloat_b200/csrc/fz_deflate_enc.cuh"
#include "../../datacompressionfloat_b200/csrc/fz_inflate.cuh"

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

// a whole plane stream (any n): groups of FZ_GROUP_SUBS sub-blocks share one code, exactly like the
// three GPU kernels (histogram, group code, emit) do it
uint64_t hm_encode_stream(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t cap, uint32_t sub, uint64_t *nstored)
{
    uint64_t o = 0, ns = 0;
    if (sub != FZ_SUB) return (uint64_t)-1;
    std::vector<uint8_t> pad(FZ_SUB + 32);
    std::vector<uint32_t> slot(FZ_SLOT_STRIDE / 4 + 8);
    FzEncState *st = (FzEncState *)malloc(sizeof(FzEncState));
    FzGroupCode *gc = (FzGroupCode *)malloc(sizeof(FzGroupCode));
    FzEmitState *es = (FzEmitState *)malloc(sizeof(FzEmitState));
    const uint64_t gbytes = (uint64_t)FZ_SUB * FZ_GROUP_SUBS;
    for (uint64_t g0 = 0; g0 < n; g0 += gbytes) {
        const uint64_t gn = (n - g0) < gbytes ? (n - g0) : gbytes;
        const uint32_t nsub = (uint32_t)((gn + FZ_SUB - 1) / FZ_SUB);
        memset(st, 0xCD, sizeof(FzEncState));
        memset(st->hist, 0, sizeof(st->hist));
        // stage 1: histogram of the group's tokens
        for (uint32_t k = 0; k < nsub; k++) {
            const uint32_t m = (uint32_t)((gn - (uint64_t)k * FZ_SUB) < FZ_SUB ? (gn - (uint64_t)k * FZ_SUB) : FZ_SUB);
            memset(pad.data(), 0, pad.size());
            memcpy(pad.data(), in + g0 + (uint64_t)k * FZ_SUB, m);
            HostLoad16 ld{pad.data()};
            HostLoadByte lb{pad.data()};
            uint32_t h[288];
            memset(h, 0, sizeof(h));
            for (int lane = 0; lane < 32; lane++) fz_ph_hist(h, ld, lb, m, lane);
            for (int i = 0; i < 288; i++) st->hist[i] += h[i];
        }
        st->hist[FZ_EOB] = nsub;
        // stage 2: one code + header for the group
        memset(gc, 0xEE, sizeof(FzGroupCode));
        fz_build_group_code(st, (uint32_t)gn, nsub, gc, 0);
        // stage 3: emit every sub-block with it
        for (uint32_t k = 0; k < nsub; k++) {
            const uint32_t m = (uint32_t)((gn - (uint64_t)k * FZ_SUB) < FZ_SUB ? (gn - (uint64_t)k * FZ_SUB) : FZ_SUB);
            memset(pad.data(), 0, pad.size());
            memcpy(pad.data(), in + g0 + (uint64_t)k * FZ_SUB, m);
            HostLoad16 ld{pad.data()};
            HostLoadByte lb{pad.data()};
            for (auto &w : slot) w = 0xDEADBEEFu;  // garbage: the encoder must write every word it owns
            memset(es, 0xAB, sizeof(FzEmitState));
            const uint32_t r = fz_emit_subblock(gc, gc->hdr, es, ld, lb, m, slot.data(), 0);
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
    free(st); free(gc); free(es);
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

// test hooks: the false-marker check of the emit stage, and the stored form
uint32_t hm_check_marker(const uint8_t *frag, uint32_t total_bytes)
{
    std::vector<uint32_t> w((total_bytes + 7) / 4 + 1, 0);
    memcpy(w.data(), frag, total_bytes);
    FzEmitState es;
    memset(&es, 0, sizeof es);
    for (int lane = 0; lane < 32; lane++) fz_ph_check_marker(&es, w.data(), total_bytes, lane);
    return es.false_marker;
}
uint32_t hm_put_stored(const uint8_t *in, uint32_t n, uint8_t *out) { return (uint32_t)put_stored(out, in, n); }

uint32_t hm_sub_bytes(void) { return FZ_SUB; }
uint32_t hm_group_subs(void) { return FZ_GROUP_SUBS; }
uint32_t hm_enc_state_bytes(void) { return (uint32_t)sizeof(FzEncState); }

}  // extern "C"
