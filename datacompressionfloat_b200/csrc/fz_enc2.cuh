// fz_enc2.cuh -- the warp-interleaved, single-pass deflate encoder for ONE sub-block (<= FZ_SUB bytes) of a byte plane.
//
// Same contract as fz_deflate_enc.cuh (what the reference gets from zlib's deflate(Z_RLE, level 6) behind mzlib_def,
// reference zip.c:164-196): distance-1 run matches of 3..258 bytes, one dynamic-Huffman block with the group's code,
// then an empty stored block (the sync-flush marker).  What changed is HOW the 32 lanes of the warp share the work:
//
//   * step geometry: in step s lane l owns the 32 bytes [s*1024 + l*32, +32) of the sub-block.  The warp reads 1 KiB
//     of contiguous bytes per step (two 128-bit loads per lane, no staging), and the token order is simply the byte
//     order -- there are no piece boundaries inside a sub-block any more.
//   * the run tokeniser is data parallel and works on aligned quads of 4 bytes: a quad is "held" (part of a distance-1
//     match) when its bytes and the quad before it repeat one byte; one shuffle brings the neighbour's flag, one
//     ballot the lanes whose four quads are all held, and that is all the cross-lane traffic a run needs.  Runs are
//     therefore multiples of 4 bytes at multiples of 4: a lane's output is eight slots, each either the codes of
//     four literals or one match token -- the same instructions either way.
//   * ONE pass: every lane looks its 16 codes up, a warp prefix sum over the lanes' bit counts gives each lane its
//     bit offset, and the bits are ORed into a 2 KiB ring in shared memory (shared-memory atomics: neighbouring
//     lanes share words).  Completed 16-byte vectors leave the ring as coalesced 128-bit stores, and are checked
//     for a chance occurrence of the sync marker on the way out.  The counting pass of the first encoder (a full
//     second tokenisation of the sub-block) and its cp.async window are gone.
//
// The code is written against a tiny warp interface (shuffle, ballot, shared-memory atomics) so that the very same
// source runs on the CPU in tests/hostmodel (32 host threads in lock step) and is checked there against zlib and
// against a plain sequential restatement of the token rule.
#pragma once
#include "fz_deflate_enc.cuh"

#if !defined(__CUDA_ARCH__)
#include <pthread.h>
#endif

FZ_HD uint32_t fz_popc32(uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popc(v);
#else
    return (uint32_t)__builtin_popcount(v);
#endif
}
FZ_HD uint32_t fz_clz32(uint32_t v)  // v != 0
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__clz((int)v);
#else
    return (uint32_t)__builtin_clz(v);
#endif
}

// ---- the warp interface ------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
struct FzWarp {
    int lane;
    FZ_D uint32_t shfl(uint32_t v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
    FZ_D uint32_t shfl_up(uint32_t v, int d) const { return __shfl_up_sync(0xffffffffu, v, d); }
    FZ_D uint32_t ballot(bool p) const { return __ballot_sync(0xffffffffu, p); }
    FZ_D void sync() const { __syncwarp(); }
    FZ_D void atom_or(uint32_t *p, uint32_t v) const { atomicOr(p, v); }
    FZ_D void atom_add(uint32_t *p, uint32_t v) const { atomicAdd(p, v); }
};
#else
// host model: 32 threads, every collective is a pair of barriers around an exchange array
struct FzWarpShared {
    pthread_barrier_t bar;
    uint32_t x[32];
};
struct FzWarp {
    int lane;
    FzWarpShared *sh;
    void wait() const { pthread_barrier_wait(&sh->bar); }
    uint32_t shfl(uint32_t v, int src) const
    {
        sh->x[lane] = v; wait();
        const uint32_t r = sh->x[src & 31]; wait();
        return r;
    }
    uint32_t shfl_up(uint32_t v, int d) const
    {
        sh->x[lane] = v; wait();
        const uint32_t r = lane >= d ? sh->x[lane - d] : v; wait();
        return r;
    }
    uint32_t ballot(bool p) const
    {
        sh->x[lane] = p ? 1u : 0u; wait();
        uint32_t r = 0;
        for (int i = 0; i < 32; i++) r |= sh->x[i] << i;
        wait();
        return r;
    }
    void sync() const { wait(); }
    void atom_or(uint32_t *p, uint32_t v) const { __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
    void atom_add(uint32_t *p, uint32_t v) const { __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
};
#endif

template <class W>
FZ_HD uint32_t fz_warp_incl_sum(const W &w, uint32_t v)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = w.shfl_up(v, d);
        if (w.lane >= d) v += o;
    }
    return v;
}

// ---- the tokeniser -------------------------------------------------------------------------------------------------
// Token rule (sequential definition; seq_tokens in tests/hostmodel restates it quad by quad):
//   the sub-block is cut into aligned quads of 4 bytes.  E[q] = the four bytes of quad q all equal the byte before
//   them (E = 0 for the first quad, which has no byte before it, and for a ragged last quad of fewer than 4 bytes).
//   Quad q is HELD iff E[q], E[q-1] and E[q-2] (FZ_E2_LEAD_QUADS = 2 quads of lead): its bytes continue a run that is at
//   least 9 bytes long already.  Every byte outside a held quad is a literal.  (One quad of lead -- E[q] and E[q-1] --
//   was the first rule: on exponent and count planes, where a byte repeats its neighbour with p = 0.45, it turned 4 % of
//   the quads into 4..8-byte matches that save no bits over 1..2-bit literals but made a quarter of the inflater's
//   instructions, executed by 1.5 lanes of 32; with two quads of lead such data has no matches at all, long runs --
//   zero planes, flat regions -- keep theirs.)  A maximal run of held quads leaves as distance-1 matches: with m the bytes of
//   the run not yet emitted, every held quad adds 4 to m and a match of 258 leaves at the quad where m reaches 258
//   (m -= 258); at the last quad of the run the rest leaves, as a match for m >= 3, as m literals for m = 1, 2.
//   Tokens leave in byte order.
// (The first encoder withheld bytes one by one; quads make a lane's output uniform slots, cost nothing measurable on
//  float planes, and let an inflater copy runs as whole words.)
#ifndef FZ_E2_LEAD_QUADS
#define FZ_E2_LEAD_QUADS 2     // all-equal quads in front of a held quad
#endif
#define FZ_E2_LQ 8             // quads per lane and step
#define FZ_E2_LB (4u * FZ_E2_LQ)        // bytes per lane and step
#define FZ_E2_STEP (32u * FZ_E2_LB)     // bytes per warp and step
#define FZ_E2_LQ_MASK ((1u << FZ_E2_LQ) - 1u)

struct FzLaneQuads { uint32_t w[FZ_E2_LQ]; };

// what the warp carries from step to step (the same value in every lane)
struct FzTokCarry {
    uint32_t prev;     // last byte of the step before; 0x100 = none (start of the sub-block)
    uint32_t prevE;    // E of its last two quads (bit 1 = the last one)
    uint32_t m;        // bytes of the current run not yet emitted (0..257)
    FZ_HD void init() { prev = 0x100u; prevE = 0; m = 0; }
};

// this lane's quads, tokenised
struct FzTok {
    uint32_t hq;      // bit g: quad g is held
    uint32_t m_in;    // bytes of a run pending in front of quad 0 (0..257)
    uint32_t pb;      // the byte before quad 0 (what such a run repeats)
};

// bit k: byte k of the 16 bytes x[0..3] equals the byte splatted in `splat4` (b * 0x01010101)
FZ_HD uint32_t fz_byte_eq_mask(const uint32_t *x4, uint32_t splat4)
{
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t x = x4[j] ^ splat4;
        const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);   // 0x80 in every zero byte of x
        m |= ((z * 0x00204081u) >> 28) << (4 * j);
    }
    return m;
}

FZ_HD void fz_store_vec16(uint32_t *p, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3)
{
#if defined(__CUDA_ARCH__)
    *(uint4 *)p = make_uint4(x0, x1, x2, x3);
#else
    p[0] = x0; p[1] = x1; p[2] = x2; p[3] = x3;
#endif
}

FZ_HD uint32_t fz_splat_top_byte(uint32_t w)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(w, 0, 0x3333);
#else
    return (w >> 24) * 0x01010101u;
#endif
}

// this lane's bytes of step s: [s * FZ_E2_STEP + lane * FZ_E2_LB, + FZ_E2_LB), as far as they exist
template <class Load16>
FZ_HD FzLaneQuads fz_lane_load(const Load16 &ld, uint32_t pos, uint32_t n)
{
    FzLaneQuads v;
#pragma unroll
    for (int h = 0; h < FZ_E2_LQ / 4; h++) {
        FzVec16 x;
        x.w[0] = x.w[1] = x.w[2] = x.w[3] = 0;
        if (pos + 16u * h < n) x = ld(pos + 16u * h);
        v.w[4 * h] = x.w[0]; v.w[4 * h + 1] = x.w[1]; v.w[4 * h + 2] = x.w[2]; v.w[4 * h + 3] = x.w[3];
    }
    return v;
}

// Tokenise one step.  nv = valid bytes of this lane (0..FZ_E2_LB; fewer only in the ragged last step of a short
// sub-block).  Returns true (in every lane) when the step has runs to book-keep (t.hq / t.m_in say which); otherwise
// every valid byte is a literal.  `c` is advanced.
template <class W>
FZ_HD bool fz_tok_step(const W &w, const FzLaneQuads &v, uint32_t nv, FzTokCarry &c, FzTok &t)
{
    const int lane = w.lane;
    const uint32_t lastb = v.w[FZ_E2_LQ - 1] >> 24;
    uint32_t pb = w.shfl_up(lastb, 1);
    if (lane == 0) pb = c.prev;
    // E: a quad whose four bytes equal the byte before it is that byte splatted
    uint32_t E = (pb <= 0xffu && v.w[0] == pb * 0x01010101u) ? 1u : 0u;
#pragma unroll
    for (int g = 1; g < FZ_E2_LQ; g++) E |= (v.w[g] == fz_splat_top_byte(v.w[g - 1]) ? 1u : 0u) << g;
    E &= (1u << (nv >> 2)) - 1u;                     // whole quads only
    // pE: E of the two quads before quad 0 (bit 1 = the last quad of the lane before, bit 0 = the one before that)
    uint32_t pE = w.shfl_up((E >> (FZ_E2_LQ - 2)) & 3u, 1);
    if (lane == 0) pE = c.prevE;
#if FZ_E2_LEAD_QUADS == 2
    const uint32_t hq = E & ((E << 1) | (pE >> 1)) & ((E << 2) | pE);
#else
    const uint32_t hq = E & ((E << 1) | (pE >> 1));
#endif
    t.hq = hq;
    t.m_in = 0;
    t.pb = pb & 0xffu;
    const uint32_t anyh = w.ballot(hq != 0);
    const uint32_t m0 = c.m;
    c.prev = w.shfl(lastb, 31);
    c.prevE = w.shfl((E >> (FZ_E2_LQ - 2)) & 3u, 31);
    if (anyh == 0 && m0 == 0) return false;
    // ---- runs.  A lane "passes" when all of its quads are held: the run goes through it.
    const bool passes = hq == FZ_E2_LQ_MASK;
    const uint32_t B = w.ballot(passes);
    // held quads at the top of a lane that does not pass: what it hands to the next lane
    const uint32_t top = passes ? 0u : fz_clz32(~(hq << (32 - FZ_E2_LQ)));
    const uint32_t below = ~B & ((1u << lane) - 1u);          // lanes before this one that do not pass
    const int j = below ? 31 - (int)fz_clz32(below) : 0;
    const uint32_t tj = w.shfl(top, j);
    const uint32_t m = (below ? 4u * tj + FZ_E2_LB * (uint32_t)(lane - 1 - j) : m0 + FZ_E2_LB * (uint32_t)lane) % FZ_MAX_MATCH;
    t.m_in = m;
    c.m = w.shfl(passes ? (m + FZ_E2_LB) % FZ_MAX_MATCH : 4u * top, 31);
    return true;
}

// the run tokens of a tokenised lane: tin = bytes of the run that ended with the lane before (its rest leaves in front
// of quad 0; 0 = none); for held quad g: cross[g] = a match of 258 leaves there, tl[g] = bytes of the rest that leaves
// there because the run ends (0 = none; the last quad of the lane never ends a run: the next lane knows)
FZ_HD void fz_tok_lens(const FzTok &t, uint32_t &tin, uint32_t &cross, uint32_t tl[FZ_E2_LQ])
{
    uint32_t cnt = t.m_in;
    tin = (t.hq & 1u) ? 0u : cnt;
    cross = 0;
#pragma unroll
    for (int g = 0; g < FZ_E2_LQ; g++) {
        tl[g] = 0;
        if ((t.hq >> g) & 1u) {
            cnt += 4u;
            if (cnt >= FZ_MAX_MATCH) { cnt -= FZ_MAX_MATCH; cross |= 1u << g; }
            const bool more = g < FZ_E2_LQ - 1 ? ((t.hq >> (g + 1)) & 1u) != 0 : true;
            if (!more) { tl[g] = cnt; cnt = 0; }
        } else cnt = 0;
    }
}

// the bits of what is left of a run of `byte`: a match of n bytes at distance 1 for n >= 3, n literals for n = 1, 2
// (<= 30 bits).  cl[] = code | len << 16.
FZ_HD void fz_run_token(const uint32_t *cl, uint32_t n, uint32_t byte, uint32_t &bits, uint32_t &nbits)
{
    if (n >= FZ_MIN_MATCH) {
        uint32_t lc, eb, ev;
        fz_len_code(n, lc, eb, ev);
        const uint32_t e = cl[257 + lc];
        bits = (e & 0xffffu) | (ev << (e >> 16));   // length code, extra bits, then the 1-bit distance code '0'
        nbits = (e >> 16) + eb + 1u;
    } else {
        const uint32_t e = cl[byte], l = e >> 16;
        bits = (e & 0xffffu) | (n == 2u ? (e & 0xffffu) << l : 0u);
        nbits = l * n;
    }
}

// ... and what it adds to a token histogram
template <class W>
FZ_HD void fz_run_count(const W &w, uint32_t *hist, uint32_t n, uint32_t byte)
{
    if (n >= FZ_MIN_MATCH) {
        uint32_t lc, eb, ev;
        fz_len_code(n, lc, eb, ev);
        w.atom_add(&hist[257 + lc], 1u);
    } else if (n) w.atom_add(&hist[byte], n);
}

// ---- histogram of one sub-block's tokens -------------------------------------------------------------------------
// hist[288] (this warp's, zeroed).  skip1 / skip2: byte values counted in registers instead of shared memory (the two
// most frequent bytes of the sub-block's sample: same-address shared atomics serialise, and on exponent planes two
// values are most of the plane); 0x100 = none.
template <class W, class Load16>
FZ_HD void fz_hist2_subblock(const W &w, uint32_t *hist, const Load16 &ld, uint32_t n, uint32_t skip1, uint32_t skip2)
{
    const int lane = w.lane;
    FzTokCarry c;
    c.init();
    uint32_t n1 = 0, n2 = 0;
    const uint32_t nsteps = (n + FZ_E2_STEP - 1u) / FZ_E2_STEP;
    const uint32_t s1 = (skip1 & 0xffu) * 0x01010101u, s2 = (skip2 & 0xffu) * 0x01010101u;
    FzLaneQuads v = fz_lane_load(ld, (uint32_t)lane * FZ_E2_LB, n);
    for (uint32_t s = 0; s < nsteps; s++) {
        const uint32_t pos = s * FZ_E2_STEP + (uint32_t)lane * FZ_E2_LB;
        const uint32_t nv = n > pos ? (n - pos < FZ_E2_LB ? n - pos : FZ_E2_LB) : 0u;
        const FzLaneQuads vn = fz_lane_load(ld, pos + FZ_E2_STEP, n);
        FzTok t;
        const bool slow = fz_tok_step(w, v, nv, c, t);
#pragma unroll
        for (int h = 0; h < FZ_E2_LQ / 4; h++) {
            const uint32_t nvh = nv > 16u * h ? (nv - 16u * h < 16u ? nv - 16u * h : 16u) : 0u;
            const uint32_t hn = (t.hq >> (4 * h)) & 15u;                                   // held quads of these 16 bytes
            const uint32_t hb = ((hn | (hn << 3) | (hn << 6) | (hn << 9)) & 0x1111u) * 15u;  // ... as byte flags
            uint32_t lit = ((1u << nvh) - 1u) & ~hb;
            if (skip1 < 0x100u) {
                const uint32_t m1 = fz_byte_eq_mask(&v.w[4 * h], s1) & lit;
                n1 += fz_popc32(m1);
                lit &= ~m1;
            }
            if (skip2 < 0x100u) {
                const uint32_t m2 = fz_byte_eq_mask(&v.w[4 * h], s2) & lit;
                n2 += fz_popc32(m2);
                lit &= ~m2;
            }
            if (lit == 0xffffu) {
#pragma unroll
                for (int k = 0; k < 16; k++) w.atom_add(&hist[(v.w[4 * h + (k >> 2)] >> ((k & 3) * 8)) & 0xffu], 1u);
            } else {
#pragma unroll
                for (int k = 0; k < 16; k++)
                    if ((lit >> k) & 1u) w.atom_add(&hist[(v.w[4 * h + (k >> 2)] >> ((k & 3) * 8)) & 0xffu], 1u);
            }
        }
        if (slow) {
            uint32_t tin, cross, tl[FZ_E2_LQ];
            fz_tok_lens(t, tin, cross, tl);
            if (tin) fz_run_count(w, hist, tin, t.pb);
            if (cross) w.atom_add(&hist[285], fz_popc32(cross));     // matches of 258
#pragma unroll
            for (int g = 0; g < FZ_E2_LQ; g++)
                if (tl[g]) fz_run_count(w, hist, tl[g], v.w[g] & 0xffu);
        }
        v = vn;
    }
    if (lane == 0 && c.m) fz_run_count(w, hist, c.m, c.prev & 0xffu);   // the run that reaches the end of the sub-block
    if (n1) w.atom_add(&hist[skip1 & 0xffu], n1);
    if (n2) w.atom_add(&hist[skip2 & 0xffu], n2);
    w.sync();
}

// ---- emission ------------------------------------------------------------------------------------------------------
#define FZ_E2_RING_WORDS 1024u                     // 4 KiB: < 33 vectors waiting + at most 32 x (32 x 15 + 21) bits of a step
#define FZ_E2_RING_MASK (FZ_E2_RING_WORDS - 1u)
#define FZ_E2_FLUSH_VECS 32u                       // vectors leave the ring 32 at a time (one per lane)

// <= 32 bits at bit offset `off` of the ring
template <class W>
FZ_HD void fz_ring_put32(const W &w, uint32_t *ring, uint32_t off, uint32_t bits, uint32_t nbits)
{
    const uint32_t s = off & 31u, wi = (off >> 5) & FZ_E2_RING_MASK;
    w.atom_or(&ring[wi], bits << s);
    if (s + nbits > 32u) w.atom_or(&ring[(wi + 1u) & FZ_E2_RING_MASK], bits >> (32u - s));
}

// <= 64 bits
template <class W>
FZ_HD void fz_ring_put64(const W &w, uint32_t *ring, uint32_t off, uint64_t bits, uint32_t nbits)
{
    const uint32_t s = off & 31u, wi = off >> 5;
    const uint64_t lo = bits << s;
    w.atom_or(&ring[wi & FZ_E2_RING_MASK], (uint32_t)lo);
    if (s + nbits > 32u) w.atom_or(&ring[(wi + 1u) & FZ_E2_RING_MASK], (uint32_t)(lo >> 32));
    if (s + nbits > 64u) w.atom_or(&ring[(wi + 2u) & FZ_E2_RING_MASK], (uint32_t)(bits >> (64u - s)));
}

// Does the 20-byte window prev | x0..x3 show 00 00 FF FF at byte offsets 1..16 (i.e. ending inside x0..x3)?
// bit o - 1 of the result.  The filter in front (two adjacent FF bytes anywhere) lets one vector in three thousand through.
FZ_HD uint32_t fz_marker_in20(uint32_t prev, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3)
{
    const uint32_t a[5] = {prev, x0, x1, x2, x3};
    uint32_t any = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const uint32_t nx = k < 4 ? a[k + 1] : 0u;
        const uint32_t t = a[k] & ((a[k] >> 8) | (nx << 24));   // byte i: bytes i and i + 1 ANDed
        any |= (~t - 0x01010101u) & t & 0x80808080u;            // non-zero iff some byte of t is FF
    }
    if (any == 0) return 0;
    uint32_t F = 0, Z = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        uint32_t x = ~a[k];
        uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);
        F |= ((z * 0x00204081u) >> 28) << (4 * k);
        x = a[k];
        z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);
        Z |= ((z * 0x00204081u) >> 28) << (4 * k);
    }
    return ((Z & (Z >> 1) & (F >> 2) & (F >> 3)) >> 1) & 0xffffu;
}

// vector vi of a fragment (pw = the word before it): does a sync marker END inside it that is not the fragment's own last
// four bytes?  total_bytes = 0: the fragment goes on.
FZ_HD bool fz_marker_vec(uint32_t pw, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t vi, uint32_t total_bytes)
{
    uint32_t m = fz_marker_in20(pw, x0, x1, x2, x3);
    if (m && total_bytes) {
        // window o (1..16) starts at byte 16 vi - 4 + o; only starts below total_bytes - 4 count
        const int64_t lim = (int64_t)total_bytes - 1 - (int64_t)vi * 16;   // number of counted windows
        if (lim <= 0) m = 0;
        else if (lim < 16) m &= (1u << lim) - 1u;
    }
    return m != 0;
}

// the ring's flush state (the same in every lane)
struct FzRingOut {
    uint32_t *ring;
    uint32_t *out;        // the fragment's slot in global memory, 16-byte aligned
    uint32_t vdone;       // 16-byte vectors already written to `out`
    uint32_t tailw;       // the last word written (what a marker window may start in)
    uint32_t bad;         // this lane saw the sync marker somewhere it must not be
    FZ_HD void init(uint32_t *r, uint32_t *o) { ring = r; out = o; vdone = 0; tailw = 0x55555555u; bad = 0; }

    // vectors [vdone, vend) leave the ring.  total_bytes != 0: the fragment ends there (its last four bytes are the one
    // marker that is supposed to be there).
    template <class W>
    FZ_HD void flush(const W &w, uint32_t vend, uint32_t total_bytes)
    {
        for (uint32_t base = vdone; base < vend; base += 32u) {
            const uint32_t vi = base + (uint32_t)w.lane;
            const bool act = vi < vend;
            uint32_t x0 = 0, x1 = 0, x2 = 0, x3 = 0;
            if (act) {
                uint32_t *p = ring + ((vi * 4u) & FZ_E2_RING_MASK);
                x0 = p[0]; x1 = p[1]; x2 = p[2]; x3 = p[3];
                p[0] = 0; p[1] = 0; p[2] = 0; p[3] = 0;
                fz_store_vec16(out + (size_t)vi * 4u, x0, x1, x2, x3);
            }
            uint32_t pw = w.shfl_up(x3, 1);
            if (w.lane == 0) pw = tailw;
            if (act && fz_marker_vec(pw, x0, x1, x2, x3, vi, total_bytes)) bad = 1;
            const uint32_t last = vend - base < 32u ? vend - base - 1u : 31u;
            tailw = w.shfl(x3, (int)last);
        }
        if (vend > vdone) vdone = vend;
        w.sync();
    }
};

// Emit one sub-block with its group's code.  Returns the fragment size in bytes, or fz_stored_size(n) |
// FZ_SIZE_STORED_FLAG when a stored block is the better (or the only safe) choice; then what was written to `out` is
// ignored (the gather kernel synthesises stored blocks from the plane bytes).
//   cl[288]   the group's code table (code | len << 16), in shared memory on the device
//   hdr       the group's block header (hdr_nbits bits)
//   ring      FZ_E2_RING_WORDS words of this warp (any content)
//   tt        256 words of this warp: filled here with the group's match tokens (lengths 3..258)
//   out       FZ_SLOT_STRIDE bytes, 16-byte aligned
template <class W, class Load16>
FZ_HD uint32_t fz_emit2_subblock(const W &w, const uint32_t *cl, const uint32_t *hdr, uint32_t hdr_nbits, uint32_t *ring,
                                 uint32_t *tt, const Load16 &ld, uint32_t n, uint32_t *out)
{
    const int lane = w.lane;
    const uint32_t stored = fz_stored_size(n) | FZ_SIZE_STORED_FLAG;
    const uint32_t limit = fz_stored_size(n) - (n >> FZ_MIN_GAIN_SHIFT);   // a coded fragment must stay below this many bytes
    if ((hdr_nbits >> 3) + 5u >= limit) return stored;
    for (uint32_t i = lane; i < FZ_E2_RING_WORDS; i += 32u) ring[i] = 0;
    for (uint32_t i = lane; i < 256u; i += 32u) {   // the bits of a match of i + 3 bytes: bits | nbits << 24 (<= 21 bits)
        uint32_t b, l;
        fz_run_token(cl, i + 3u, 0u, b, l);
        tt[i] = b | (l << 24);
    }
    w.sync();
    const uint32_t t258 = tt[255];
    {   // the block header (<= 160 words)
        const uint32_t nw = (hdr_nbits + 31u) >> 5;
        for (uint32_t i = lane; i < nw; i += 32u) {
            uint32_t x = hdr[i];
            if (i == nw - 1u && (hdr_nbits & 31u)) x &= (1u << (hdr_nbits & 31u)) - 1u;
            ring[i] = x;
        }
        w.sync();
    }
    FzRingOut ro;
    ro.init(ring, out);
    uint32_t P = hdr_nbits;                          // bits emitted so far
    ro.flush(w, P >> 7, 0);

    FzTokCarry c;
    c.init();
    const uint32_t nsteps = (n + FZ_E2_STEP - 1u) / FZ_E2_STEP;
    FzLaneQuads v = fz_lane_load(ld, (uint32_t)lane * FZ_E2_LB, n);
    for (uint32_t s = 0; s < nsteps; s++) {
        const uint32_t pos = s * FZ_E2_STEP + (uint32_t)lane * FZ_E2_LB;
        const uint32_t nv = n > pos ? (n - pos < FZ_E2_LB ? n - pos : FZ_E2_LB) : 0u;
        const FzLaneQuads vn = fz_lane_load(ld, pos + FZ_E2_STEP, n);
        FzTok t;
        const bool slow = fz_tok_step(w, v, nv, c, t);
        // one slot per quad: the codes of its four literals ...
        uint64_t qv[FZ_E2_LQ];
        uint32_t ql[FZ_E2_LQ];
#pragma unroll
        for (int g = 0; g < FZ_E2_LQ; g++) {
            const uint32_t x = v.w[g];
            uint32_t e0 = cl[x & 0xffu], e1 = cl[(x >> 8) & 0xffu], e2 = cl[(x >> 16) & 0xffu], e3 = cl[x >> 24];
            if (nv < FZ_E2_LB) {   // the ragged last step of a short sub-block
                if (4u * g + 0u >= nv) e0 = 0;
                if (4u * g + 1u >= nv) e1 = 0;
                if (4u * g + 2u >= nv) e2 = 0;
                if (4u * g + 3u >= nv) e3 = 0;
            }
            const uint32_t l0 = e0 >> 16, l2 = e2 >> 16;
            const uint32_t p0 = (e0 & 0xffffu) | ((e1 & 0xffffu) << l0), pl0 = l0 + (e1 >> 16);
            const uint32_t p1 = (e2 & 0xffffu) | ((e3 & 0xffffu) << l2), pl1 = l2 + (e3 >> 16);
            qv[g] = (uint64_t)p0 | ((uint64_t)p1 << pl0);
            ql[g] = pl0 + pl1;
        }
        uint32_t tinb = 0, tinl = 0;
        if (slow) {   // ... or, for a held quad, what leaves there: a match of 258, the rest of a run that ends, or nothing
            uint32_t tin, cross, tl[FZ_E2_LQ];
            fz_tok_lens(t, tin, cross, tl);
            if (tin >= FZ_MIN_MATCH) { const uint32_t x = tt[tin - 3u]; tinb = x & 0xffffffu; tinl = x >> 24; }
            else if (tin) fz_run_token(cl, tin, t.pb, tinb, tinl);
#pragma unroll
            for (int g = 0; g < FZ_E2_LQ; g++)
                if ((t.hq >> g) & 1u) {
                    uint64_t b = 0;
                    uint32_t l = 0;
                    if ((cross >> g) & 1u) { b = t258 & 0xffffffu; l = t258 >> 24; }
                    if (tl[g] >= FZ_MIN_MATCH) {
                        const uint32_t x = tt[tl[g] - 3u];
                        b |= (uint64_t)(x & 0xffffffu) << l;
                        l += x >> 24;
                    } else if (tl[g]) {   // one or two bytes left over behind a match of 258: literals
                        uint32_t eb, el;
                        fz_run_token(cl, tl[g], v.w[g] & 0xffu, eb, el);
                        b |= (uint64_t)eb << l;
                        l += el;
                    }
                    qv[g] = b;
                    ql[g] = l;
                }
        }
        uint32_t tot = tinl;
#pragma unroll
        for (int g = 0; g < FZ_E2_LQ; g++) tot += ql[g];
        const uint32_t inc = fz_warp_incl_sum(w, tot);
        uint32_t off = P + inc - tot;
        if (tinl) { fz_ring_put32(w, ring, off, tinb, tinl); off += tinl; }
#pragma unroll
        for (int g = 0; g < FZ_E2_LQ; g++) {
            if (ql[g]) fz_ring_put64(w, ring, off, qv[g], ql[g]);
            off += ql[g];
        }
        P += w.shfl(inc, 31);
        w.sync();
        if ((P >> 3) + 5u >= limit) return stored;   // cannot beat a stored block any more (and must not outgrow the slot)
        if ((P >> 7) - ro.vdone >= FZ_E2_FLUSH_VECS) ro.flush(w, P >> 7, 0);
        v = vn;
    }
    // ---- the run that reaches the end of the sub-block, end of block, the empty stored block
    if (lane == 0) {
        uint32_t off = P;
        if (c.m) {
            uint32_t b, l;
            fz_run_token(cl, c.m, c.prev & 0xffu, b, l);
            fz_ring_put32(w, ring, off, b, l);
            off += l;
        }
        fz_ring_put32(w, ring, off, cl[FZ_EOB] & 0xffffu, cl[FZ_EOB] >> 16);
        off += cl[FZ_EOB] >> 16;
        off += 3u;                                   // BFINAL = 0, BTYPE = 00
        off = (off + 7u) & ~7u;
        fz_ring_put32(w, ring, off, 0xFFFF0000u, 32u);   // LEN = 0, NLEN = 0xFFFF
        off += 32u;
        P = off;
    }
    P = w.shfl(P, 0);
    w.sync();
    const uint32_t total = P >> 3;
    if (total >= limit) return stored;
    ro.flush(w, (P + 127u) >> 7, total);
    if (w.ballot(ro.bad != 0)) return stored;
    return total;
}
