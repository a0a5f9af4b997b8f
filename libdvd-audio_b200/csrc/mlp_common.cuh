// mlp_common.cuh — device-side pieces shared by the MLP kernel files (mlp_decode.cu: check data,
// header passes, complete decoder, rematrix; mlp_fused.cu: the fused entropy + filter + output
// pass): the shared-memory ring bit reader, the per-access-unit records the header passes leave
// behind, and the prediction-filter step.
#pragma once
#include "common.cuh"
#include "kernels.cuh"
#include <cstddef>
#include <mutex>

// "this kernel's attributes have been set on the current device" (function attributes are per
// device; a process may run engines on several, and from several threads: the set-up runs under
// a lock and a device counts as done only once the call has succeeded)
struct PerDeviceOnce {
    std::mutex mu;
    bool done[64] = {};
    template <typename F> int run(F &&setup)
    {
        int dev = 0;
        const bool known = cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64;
        std::lock_guard<std::mutex> hold(mu);
        if (known && done[dev]) return 0;
        const int rc = setup();
        if (rc == 0 && known) done[dev] = true;
        return rc;
    }
};

// ------------------------------------------------------------- bit reader

// MSB-first reader over the elementary stream (reference src/bitstream.c:1077-1111).
//
// Every lane walks its own stream, so plain loads would miss in a different
// cache line per lane and — worse — any per-lane "refill when low" branch
// diverges: an event that is rare for one lane happens almost every step for
// some lane of the warp.  So:
//   * each lane owns a 256-byte ring in shared memory (four 64-byte chunks),
//     layout [16-byte slot][lane];
//   * it fills the ring itself with cp.async (16 bytes per copy, L2 -> shared),
//     at warp-uniform points (top of every 8-frame iteration): up to two chunks,
//     always two commit groups, then wait_group 2 — everything issued in earlier
//     iterations has landed, nothing ever waits on DRAM in steady state;
//   * the hot loop tops the 64-bit window up without branching (the load from
//     the ring is unconditional, the merge is predicated).
// Headers use the same reader through the checked (cold) entry points.
#ifndef DVDA_UNROLL
#define DVDA_UNROLL 8
#endif
#ifndef RING_SLOTS
#define RING_SLOTS 16                 // 16-byte slots per lane: 256 bytes
#endif
#ifndef CHUNK_WORDS
#define CHUNK_WORDS 16                // 64 bytes per cp.async group
#endif
#define CHUNK_SLOTS (CHUNK_WORDS / 4)
#define RING_WORDS (RING_SLOTS * 4)

struct Rd {
    const uint8_t *es;      // elementary stream (global)
    uint32_t ring;          // shared-memory address of this lane's slot 0
    uint64_t win;           // upcoming bits, MSB aligned
    int32_t avail;          // valid bits in win
    uint32_t next_w;        // absolute word index of the next word to pull
    uint32_t fill_c;        // chunks [.., fill_c) have been issued
    uint32_t safe_w;        // words [.., safe_w) are known to have landed
    uint32_t base_w;        // word index the bit counter is relative to
    uint32_t ahead;         // hot loop only: ring word next_w, fetched one step early
};

__device__ __forceinline__ void cp_async16(uint32_t smem, const void *g)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void rd_issue_chunk(Rd &r, uint32_t c)
{
    const uint8_t *g = r.es + (uint64_t)c * (CHUNK_WORDS * 4);
#pragma unroll
    for (int t = 0; t < CHUNK_SLOTS; t++) cp_async16(r.ring + (((c * CHUNK_SLOTS + t) & (RING_SLOTS - 1)) << 9), g + t * 16);
}
// may chunk fill_c be written?  Its slot held chunk fill_c - 8, which must lie
// entirely behind the read position (the slot held chunk fill_c - RING_CHUNKS).
#define RING_CHUNKS (RING_SLOTS / CHUNK_SLOTS)
__device__ __forceinline__ bool rd_room(const Rd &r) { return (int32_t)((r.fill_c - (RING_CHUNKS - 1)) * CHUNK_WORDS - r.next_w) <= 0; }

__device__ __forceinline__ void rd_init(Rd &r, const uint8_t *es, uint32_t ring)
{
    r.es = es; r.ring = ring; r.win = 0; r.avail = 0; r.next_w = 0; r.fill_c = 0; r.safe_w = 0; r.base_w = 0;
}

// cold: make sure words [next_w, next_w + need) are in the ring
__device__ __forceinline__ void rd_slow_fill(Rd &r, uint32_t need)
{
    while (r.fill_c * CHUNK_WORDS < r.next_w + need + CHUNK_WORDS && rd_room(r)) { rd_issue_chunk(r, r.fill_c); r.fill_c++; }
    cp_commit();
    cp_wait<0>();
    r.safe_w = r.fill_c * CHUNK_WORDS;
}

// warp-uniform prefetch point: keep the ring ~6 chunks ahead of the reader and
// guarantee that the next `need` words have landed (the hot loop reads them
// without checking).  In steady state the guarantee holds by construction;
// right after a long header it may not, then the lane takes the slow path.
__device__ __forceinline__ void rd_prefetch(Rd &r, uint32_t need)
{
    const uint32_t landed = r.fill_c * CHUNK_WORDS;
#pragma unroll
    for (int t = 0; t < 2; t++) {
#pragma unroll
        for (int u = 0; u < 16 / CHUNK_WORDS; u++)
            if (r.fill_c * CHUNK_WORDS < r.next_w + RING_WORDS - CHUNK_WORDS && rd_room(r)) { rd_issue_chunk(r, r.fill_c); r.fill_c++; }
        cp_commit();
    }
    cp_wait<2>();
    r.safe_w = landed;
    if (r.next_w + need > r.safe_w) rd_slow_fill(r, need);
}

// fill the whole ring ahead of the reader without waiting (cold starts: one DRAM
// latency then covers ~450 bytes, about one access unit)
__device__ __forceinline__ void rd_issue_ahead(Rd &r)
{
    while (r.fill_c * CHUNK_WORDS < r.next_w + RING_WORDS - CHUNK_WORDS && rd_room(r)) { rd_issue_chunk(r, r.fill_c); r.fill_c++; }
    cp_commit();
}

// position the reader at an absolute byte offset
__device__ __forceinline__ void rd_seat(Rd &r, uint64_t byte_pos)
{
    const uint32_t w = (uint32_t)(byte_pos >> 2);
    if (w >= r.fill_c * CHUNK_WORDS || w + RING_WORDS - CHUNK_WORDS < r.fill_c * CHUNK_WORDS) {
        cp_wait<0>();                       // nothing may still be landing in slots we reuse
        r.fill_c = w / CHUNK_WORDS;
        r.safe_w = r.fill_c * CHUNK_WORDS;
    }
    r.next_w = w;
    r.base_w = w;
    r.win = 0;
    r.avail = 0;
}

__device__ __forceinline__ uint32_t rd_ring_word(const Rd &r, uint32_t w)
{
    uint32_t word;
    const uint32_t addr = r.ring + (((w >> 2) & (RING_SLOTS - 1)) << 9) + ((w & 3) << 2);
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(word) : "r"(addr) : "memory");
    return __byte_perm(word, 0, 0x0123);
}

// checked pull (headers, generic path): avail must be <= 32
__device__ __forceinline__ void rd_pull(Rd &r)
{
    if (r.next_w >= r.safe_w) rd_slow_fill(r, 48);
    const uint32_t word = rd_ring_word(r, r.next_w);
    r.win |= (uint64_t)word << (32 - r.avail);
    r.avail += 32;
    r.next_w++;
}

// hot pull: no branch; afterwards avail > 32.  The caller guarantees (through
// rd_prefetch) that the next words have landed.  The ring word is fetched one
// step ahead (r.ahead) so that the shared-memory latency is off the critical
// path of the window; rd_hot_begin() primes it.
__device__ __forceinline__ void rd_hot_begin(Rd &r) { r.ahead = rd_ring_word(r, r.next_w); }
__device__ __forceinline__ void rd_top_up(Rd &r)
{
    const bool need = r.avail <= 32;
    const uint64_t add = (uint64_t)r.ahead << ((32 - r.avail) & 63);
    r.win |= need ? add : 0ull;
    r.avail += need ? 32 : 0;
    r.next_w += need ? 1u : 0u;
    r.ahead = rd_ring_word(r, r.next_w);
}

// next n bits (1..32) without consuming them
template <typename RD>
__device__ __forceinline__ uint32_t rd_peek(RD &r, uint32_t n)
{
    if (r.avail < (int32_t)n) rd_pull(r);
    return (uint32_t)(r.win >> (64 - n));
}
template <typename RD>
__device__ __forceinline__ void rd_drop(RD &r, uint32_t n) { r.win <<= n; r.avail -= n; }
template <typename RD>
__device__ __forceinline__ uint32_t rd_get(RD &r, uint32_t n)
{
    if (!n) return 0;
    const uint32_t v = rd_peek(r, n);
    rd_drop(r, n);
    return v;
}
// two's complement, n in 1..32 (src/bitstream.c:1198-1206)
template <typename RD>
__device__ __forceinline__ int32_t rd_get_s(RD &r, uint32_t n)
{
    const uint32_t v = rd_get(r, n);
    return (int32_t)(v << (32 - n)) >> (32 - n);
}
template <typename RD>
__device__ __forceinline__ void rd_skip(RD &r, uint32_t n)
{
    while (n > 32) { rd_get(r, 32); n -= 32; }
    rd_get(r, n);
}
// bits consumed since the reader was seated (counted from the seated word's first bit)
template <typename RD>
__device__ __forceinline__ uint32_t rd_pos(const RD &r) { return (r.next_w - r.base_w) * 32 - r.avail; }

// LSB-bypass bits of one frame: one bit per matrix that has the flag, in matrix order
__device__ __forceinline__ uint32_t bypass_bits(Rd &b, uint32_t want_mask)
{
    uint32_t out = 0;
    if (want_mask) {
        uint32_t t = __popc(want_mask);
        const uint32_t bits = rd_get(b, t);
        uint32_t m = want_mask;
        while (m) {
            const uint32_t k = __ffs(m) - 1;
            m &= m - 1;
            t--;
            out |= ((bits >> t) & 1u) << k;
        }
    }
    return out;
}

// 32 x 32 -> 64-bit multiply-add in one instruction (IMAD.WIDE)
__device__ __forceinline__ long long mad_wide(int32_t a, int32_t b, long long c)
{
    long long d;
    asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
    return d;
}

__device__ __forceinline__ uint32_t noise_step(uint32_t seed)
{
    const uint32_t sh = (seed >> 7) & 0xFFFF;
    return (seed << 16) ^ sh ^ (sh << 5);
}

// RIFF WAVE slot of MLP channel c (table at mlp.c:416-438): identity except for
// assignments 0x12-0x14
__device__ __forceinline__ uint32_t wave_slot(uint32_t assignment, uint32_t c)
{
    if (assignment == 0x12 || assignment == 0x13) return (0x24310u >> (4 * c)) & 15;      // 0,1,3,4,2
    if (assignment == 0x14) return (0x325410u >> (4 * c)) & 15;                          // 0,1,4,5,2,3
    return c;
}

struct ChanSnap { int32_t sho; uint8_t cb, lsb_bits, q, shift; };
struct AuSnap {
    uint64_t bit0;          // absolute bit position (in the ES) of the first residual bit
    uint64_t bit_end;       // absolute bit position of the end of the substream data
    uint16_t block_size;
    uint8_t want, valid;
    uint8_t min_ch, nch, has_matrix, pad1;
    ChanSnap ch[4];
};

// noise generator advanced by n frames
__device__ __forceinline__ uint32_t noise_advance(uint32_t seed, uint32_t n)
{
    for (uint32_t i = 0; i < n; i++) seed = noise_step(seed);
    return seed;
}

struct SegCtx { uint32_t seed; uint8_t min_ch, max_ch, mmc, flags, noise_shift, ok, pad[2]; };

#define CD_PRESENT 1u
#define CD_FIR 2u
#define CD_IIR 4u
#define CD_IIR_STATE 8u
#define CD_OFFSET 16u
#define AD_BLOCK 1u
#define AD_MATRIX 2u
#define AD_SHIFT 4u
#define AD_Q 8u
struct ChanHead {
    int32_t huff_offset;
    uint8_t fir_order, fir_shift, iir_order, iir_shift;
    uint8_t codebook, huff_lsbs, present, pad;
};
struct ChanCoef {
    int32_t ist[8];                  // IIR history as transmitted: [0] pairs with coefficient 0
    int16_t fir_c[8], iir_c[8];
};
// What an access unit transmits, in three tables indexed alike ([2][cap_au]): the head is all the
// resolve pass looks at and is written for every unit with parameters (consecutive units lie
// next to each other: a segment's chain is a dense run of memory); filter coefficients and
// matrix coefficients are written — and read — only where they are transmitted.
struct __align__(16) AuDelta {
    uint16_t block_size;
    uint8_t present, matrix_len;
    uint8_t mat_out[DVDA_MAX_MAT], mat_bypass[DVDA_MAX_MAT];
    uint8_t out_shift[DVDA_MAX_CH], q[DVDA_MAX_CH];
    ChanHead ch[4];
};
struct __align__(16) MatCoef { int16_t c[DVDA_MAX_MAT][DVDA_MAX_CH]; };
static_assert(sizeof(AuDelta) == 80 && sizeof(ChanCoef) == 64 && sizeof(MatCoef) == 96, "AuDelta layout");

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- filter passes: one channel's set-up from the delta of an access unit -----------------
struct FiltSetup { uint32_t fo, io, fsh, ish, q; };      // orders, shifts, quant_step_size in force

// The head fields are in registers already (the output pass loads them one access unit ahead):
// only the 64 bytes of coefficients and histories are fetched here, with four independent loads.
struct DeltaHead { uint32_t fchg, w0, qv, h1, h2, seed, pset; };
__device__ __forceinline__ void filt_take_head(const DeltaHead &H, const ChanCoef &K, FiltSetup &F,
                                               int32_t (&cf)[8], int32_t (&ci)[8], int32_t (&ih)[8])
{
    const uint32_t h1 = H.h1, p = (H.h2 >> 16) & 0xFF;
    if ((H.w0 >> 16) & AD_Q) F.q = H.qv;
    if (!(p & (CD_FIR | CD_IIR))) return;
    const uint4 *kw = reinterpret_cast<const uint4 *>(&K);
    const uint4 s0 = kw[0], s1 = kw[1], fc = kw[2], ic = kw[3];
    if (p & CD_FIR) {
        F.fo = h1 & 0xFF; F.fsh = (h1 >> 8) & 0xFF;
        const uint32_t w[4] = {fc.x, fc.y, fc.z, fc.w};
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int32_t v = (int32_t)(int16_t)(w[j >> 1] >> (16 * (j & 1)));
            cf[j] = (uint32_t)j < F.fo ? v : 0;
        }
    }
    if (p & CD_IIR) {
        F.io = (h1 >> 16) & 0xFF; F.ish = h1 >> 24;
        const uint32_t w[4] = {ic.x, ic.y, ic.z, ic.w};
        const int32_t st[8] = {(int32_t)s0.x, (int32_t)s0.y, (int32_t)s0.z, (int32_t)s0.w,
                               (int32_t)s1.x, (int32_t)s1.y, (int32_t)s1.z, (int32_t)s1.w};
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int32_t v = (int32_t)(int16_t)(w[j >> 1] >> (16 * (j & 1)));
            ci[j] = (uint32_t)j < F.io ? v : 0;
            ih[j] = ((p & CD_IIR_STATE) && (uint32_t)j < F.io) ? st[j] : 0;
        }
    }
}
__device__ __forceinline__ uint32_t filt_shift(const FiltSetup &F)
{
    return (F.fsh > 0 && F.ish > 0) ? F.fsh : F.fo > 0 ? F.fsh : F.ish;
}

template <int NF, int NI>
__device__ __forceinline__ void filt8(const int32_t (&cf)[8], const int32_t (&ci)[8], int32_t (&fh)[8], int32_t (&ih)[8],
                                      int32_t (&r)[8], uint32_t shift, uint32_t qmask)
{
#pragma unroll
    for (int j = 0; j < 8; j++) {
        long long s0 = 0, s1 = 0;
#pragma unroll
        for (int t = NF - 1; t >= 0; t--) s0 = mad_wide(cf[t], fh[(t - j) & 7], s0);   // oldest taps first:
#pragma unroll
        for (int t = NI - 1; t >= 0; t--) s1 = mad_wide(ci[t], ih[(t - j) & 7], s1);   // short dependent chain
        const int32_t ssum = (NF + NI) ? (int32_t)((s0 + s1) >> shift) : 0;
        const int32_t x = (int32_t)(((uint32_t)ssum + (uint32_t)r[j]) & qmask);
        fh[(7 - j) & 7] = x;
        ih[(7 - j) & 7] = (int32_t)((uint32_t)x - (uint32_t)ssum);
        r[j] = x;
    }
}


// Does the output pass of the fast path take this track (one substream of up to four channels,
// or the stereo pair + up to four more in a second substream)?  What it does not take — and the
// segments the header passes gave up on — goes through the tiles and k_rematrix.
__device__ __forceinline__ bool track_fusable(const TrackDev &T)
{
    if (T.nss == 1) return T.channels >= 1 && T.channels <= 4;
    return T.nss == 2 && T.channels >= 3 && T.channels <= 6;
}
__device__ __forceinline__ bool seg_output_done(const MlpTables &m, const TrackDev &T, uint32_t seg)
{
    if (!m.fast || !track_fusable(T)) return false;
    uint32_t fl = m.ss_flags_fast[seg];
    if (T.nss == 2) fl |= m.ss_flags_fast[m.cap_seg + seg];
    return !(fl & SEG_FALLBACK);
}
