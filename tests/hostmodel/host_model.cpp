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
}  // namespace

extern "C" {

// one sub-block (n <= FZ_SUB) -> complete fragment in out (cap >= n + 64); returns bytes; *stored = 1 if stored form
uint32_t hm_encode_subblock(const uint8_t *in, uint32_t n, uint8_t *out, int *stored)
{
    std::vector<uint8_t> pad(n + 32, 0);
    memcpy(pad.data(), in, n);
    std::vector<uint32_t> slot(FZ_SLOT_STRIDE / 4 + 8, 0xDEADBEEFu);  // garbage: the encoder must write every word it owns
    FzEncState *st = (FzEncState *)malloc(sizeof(FzEncState));
    memset(st, 0xCD, sizeof(FzEncState));
    HostLoad16 ld{pad.data()};
    HostLoadByte lb{pad.data()};
    uint32_t r = fz_encode_subblock(st, ld, lb, n, slot.data(), 0);
    free(st);
    if (r & FZ_SIZE_STORED_FLAG) {
        // what the gather kernel synthesises: stored block + empty stored block
        uint8_t *o = out;
        *o++ = 0x00;
        *o++ = (uint8_t)(n & 0xff); *o++ = (uint8_t)(n >> 8);
        *o++ = (uint8_t)(~n & 0xff); *o++ = (uint8_t)((~n >> 8) & 0xff);
        memcpy(o, in, n); o += n;
        *o++ = 0x00; *o++ = 0x00; *o++ = 0x00; *o++ = 0xFF; *o++ = 0xFF;
        *stored = 1;
        return (uint32_t)(o - out);
    }
    memcpy(out, slot.data(), r);
    *stored = 0;
    return r;
}

// a whole plane stream (any n) as the concatenation of its sub-blocks
uint64_t hm_encode_stream(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t cap, uint32_t sub, uint64_t *nstored)
{
    uint64_t o = 0, ns = 0;
    std::vector<uint8_t> tmp(FZ_SUB + 64);
    for (uint64_t p = 0; p < n; p += sub) {
        uint32_t m = (uint32_t)((n - p) < sub ? (n - p) : sub);
        int stored = 0;
        uint32_t r = hm_encode_subblock(in + p, m, tmp.data(), &stored);
        if (o + r > cap) return (uint64_t)-1;
        memcpy(out + o, tmp.data(), r);
        o += r; ns += (uint64_t)stored;
    }
    if (nstored) *nstored = ns;
    return o;
}

int hm_inflate(const uint8_t *in, uint64_t in_len, uint8_t *out, uint32_t out_cap, uint32_t *out_n, uint64_t *in_used)
{
    // word-aligned, padded copies (the device reads whole aligned words; out must be 4-byte aligned)
    std::vector<uint32_t> ibuf((in_len + 16) / 4 + 1, 0);
    std::vector<uint32_t> obuf(out_cap / 4 + 2, 0);
    for (int mis = 0; mis < 1; mis++) {}
    uint8_t *ip = (uint8_t *)ibuf.data() + 1;  // deliberately misaligned input
    memcpy(ip, in, in_len);
    uint16_t ll[288], dd[32], cnt[32];
    FzInfTab<1> tab{ll, dd, cnt};
    size_t used = 0;
    int rc = fz_inflate(ip, (size_t)in_len, (uint8_t *)obuf.data(), out_cap, tab, out_n, &used);
    memcpy(out, obuf.data(), *out_n);
    *in_used = used;
    return rc;
}

uint32_t hm_sub_bytes(void) { return FZ_SUB; }
uint32_t hm_enc_state_bytes(void) { return (uint32_t)sizeof(FzEncState); }

}  // extern "C"
