// mlp_index.cu — finding the work: major syncs, track boundaries, restart
// segments, access units.
//
// Replaces (reference tree):
//   src/dvd-audio.c:1250-1286 find_major_sync         -> k_sync_scan (all positions at once)
//   src/dvd-audio.c:1318-1365 locate_mlp_parameters   -> k_track_setup (first sync, parameters)
//   src/dvd-audio.c:1367-1421 mlp_data_to_major_sync  -> k_track_setup (end of track)
//   src/dvd-audio.c:952-1082  PCM track length rules  -> k_track_setup
//   src/mlp.c:384-405         read_mlp_frame          -> k_au_chase (12-bit length chain)
//   src/dvd-audio.c:766-775   "a packet that completes no access unit ends the
//                              stream"                -> k_yield_mark / k_yield_find
//
// The reference discovers all of this one packet at a time while decoding.  Here
// the sync pattern is matched at every byte of the elementary stream in parallel,
// tracks are resolved with binary searches over prefix-summed tables, and each
// restart segment walks its own chain of access-unit lengths.
#include "common.cuh"
#include "kernels.cuh"


// A sync position starts a *segment* only if the access unit there is a
// well-formed major sync (1 or 2 substreams) and every substream opens with a
// restart header: params-present bit, restart bit, 13-bit sync 0x18F5
// (reference mlp.c:749-759, 822-835).  Anything else stays inside the previous
// segment (or is a chance match in the payload).
__device__ __noinline__ bool sync_starts_segment(const uint8_t *es, uint64_t p, uint64_t es_total)
{
    if (p + 4 + 28 + 2 > es_total) return false;
    const uint32_t total = ((ld_u8(es + p) & 15u) << 8 | ld_u8(es + p + 1)) * 2;
    const uint32_t ns = ld_u8(es + p + 20) >> 4;
    if (ns != 1 && ns != 2) return false;
    uint64_t d = p + 32;
    uint32_t end[2] = {0, 0};
    for (uint32_t k = 0; k < ns; k++) {
        if (d + 2 > es_total) return false;
        const uint32_t b0 = ld_u8(es + d);
        end[k] = (((b0 & 15u) << 8) | ld_u8(es + d + 1)) * 2;
        d += 2 + ((b0 >> 7) ? 2 : 0);
    }
    for (uint32_t k = 0; k < ns; k++) {
        const uint64_t s = d + (k ? end[0] : 0);
        if (s + 2 > p + total || s + 2 > es_total) return false;
        if (ld_u8(es + s) != 0xF1 || (ld_u8(es + s + 1) & 0xFE) != 0xEA) return false;
    }
    return true;
}

// The sync patterns are found while the stream is gathered (k_es_gather, demux.cu): every
// 512-byte chunk of the stream has a count and two slots (offset in the chunk).  k_sync_validate,
// one thread per chunk with matches, puts a chunk's slots in stream order and judges them
// (slot bit 15: starts a segment); after the counts are scanned, k_sync_emit (one thread per
// chunk) moves the slots to their places in the ordered lists.  A chunk with more matches than
// slots is searched again by its thread (never seen outside of tests with synthetic pattern
// floods).
#define SYNC_SLOTS 2

__global__ void k_sync_validate(const uint8_t *__restrict__ es, const DecCounts *__restrict__ cnt, uint32_t chunks_cap,
                                const uint32_t *__restrict__ cnt_raw, uint32_t *__restrict__ cnt_valid,
                                uint16_t *__restrict__ slots, uint32_t nslots)
{
    const uint32_t ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= chunks_cap) return;
    const uint32_t n = cnt_raw[ch];
    if (!n) return;                                            // (cnt_valid was cleared together with cnt_raw)
    const uint64_t es_total = cnt->es_total;
    const uint64_t p0 = (uint64_t)ch * SYNC_CHUNK;
    uint32_t nv = 0;
    if (n <= nslots) {
        uint32_t e[SYNC_SLOTS];
        for (uint32_t i = 0; i < n; i++) e[i] = slots[(uint64_t)ch * SYNC_SLOTS + i] & 0x7FFFu;
        if (n == 2 && e[0] > e[1]) { const uint32_t t = e[0]; e[0] = e[1]; e[1] = t; }      // (they arrived in any order)
        for (uint32_t i = 0; i < n; i++) {
            const uint32_t v = sync_starts_segment(es, p0 + e[i], es_total) ? 1u : 0u;
            nv += v;
            slots[(uint64_t)ch * SYNC_SLOTS + i] = (uint16_t)(e[i] | (v << 15));
        }
    } else {
        for (uint32_t j = 0; j < SYNC_CHUNK; j++) {
            const uint64_t p = p0 + j;
            if (p + 8 > es_total) break;
            if (ld_be32(es + p + 4) == 0xF8726FBBu && sync_starts_segment(es, p, es_total)) nv++;
        }
    }
    cnt_valid[ch] = nv;
}

__global__ void k_sync_emit(const uint8_t *__restrict__ es, const DecCounts *__restrict__ cnt,
                            const uint32_t *__restrict__ cnt_raw, const uint16_t *__restrict__ slots, uint32_t nslots,
                            const uint32_t *__restrict__ base_raw, const uint32_t *__restrict__ base_valid,
                            uint64_t *__restrict__ raw, uint32_t cap_raw, uint64_t *__restrict__ valid, uint32_t cap_valid)
{
    // (the lists were sized before their lengths were known: nothing is written behind their ends)
    const uint64_t es_total = cnt->es_total;
    const uint32_t chunks = (uint32_t)((es_total + SYNC_CHUNK - 1) / SYNC_CHUNK);
    const uint32_t ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= chunks) return;
    const uint32_t n = cnt_raw[ch];
    if (!n) return;
    uint32_t ir = base_raw[ch], iv = base_valid[ch];
    const uint64_t p0 = (uint64_t)ch * SYNC_CHUNK;
    if (n <= nslots) {
        for (uint32_t i = 0; i < n; i++) {
            const uint32_t e = slots[(uint64_t)ch * SYNC_SLOTS + i];
            if (ir < cap_raw) raw[ir] = p0 + (e & 0x7FFF);
            ir++;
            if (e >> 15) { if (iv < cap_valid) valid[iv] = p0 + (e & 0x7FFF); iv++; }
        }
        return;
    }
    for (uint32_t j = 0; j < SYNC_CHUNK; j++) {
        const uint64_t p = p0 + j;
        if (p + 8 > es_total) break;
        if (ld_be32(es + p + 4) != 0xF8726FBBu) continue;
        if (ir < cap_raw) raw[ir] = p;
        ir++;
        if (sync_starts_segment(es, p, es_total)) { if (iv < cap_valid) valid[iv] = p; iv++; }
    }
}

int launch_sync_validate(const uint8_t *es, const DecCounts *cnt, uint32_t chunks_cap, const uint32_t *cnt_raw, uint32_t *cnt_valid,
                         uint16_t *slots, uint32_t nslots, cudaStream_t s)
{
    LAUNCH(k_sync_validate, div_up_u32(chunks_cap, 256), 256, 0, s, es, cnt, chunks_cap, cnt_raw, cnt_valid, slots, nslots < SYNC_SLOTS ? nslots : SYNC_SLOTS);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
int launch_sync_fill(const uint8_t *es, const DecCounts *cnt, uint32_t chunks_cap, const uint32_t *cnt_raw, const uint16_t *slots, uint32_t nslots,
                     const uint32_t *base_raw, const uint32_t *base_valid, uint64_t *raw, uint32_t cap_raw,
                     uint64_t *valid, uint32_t cap_valid, cudaStream_t s)
{
    LAUNCH(k_sync_emit, div_up_u32(chunks_cap, 128), 128, 0, s, es, cnt, cnt_raw, slots, nslots < SYNC_SLOTS ? nslots : SYNC_SLOTS, base_raw, base_valid, raw, cap_raw, valid, cap_valid);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------- track setup

__device__ __forceinline__ uint32_t unpack_bps(uint32_t f) { return f == 0 ? 16 : f == 1 ? 20 : f == 2 ? 24 : 0; }
__device__ __forceinline__ uint32_t unpack_rate(uint32_t f)
{
    switch (f) {
    case 0: return 48000; case 1: return 96000; case 2: return 192000;
    case 8: return 44100; case 9: return 88200; case 10: return 176400;
    default: return 0;
    }
}
__device__ __forceinline__ uint32_t channel_count(uint32_t a)
{
    const unsigned long long packed =
        (1ull << 0) | (2ull << 3) | (3ull << 6) | (4ull << 9) | (3ull << 12) | (4ull << 15) | (5ull << 18) |
        (3ull << 21) | (4ull << 24) | (5ull << 27) | (4ull << 30) | (5ull << 33) | (6ull << 36) | (4ull << 39) |
        (5ull << 42) | (4ull << 45) | (5ull << 48) | (6ull << 51) | (5ull << 54) | (5ull << 57) | (6ull << 60);
    return a <= 20 ? (uint32_t)((packed >> (3 * a)) & 7) : 0;
}

// The searches of the track set-up, by a whole warp: every lane probes the end of one of 32
// stripes of the range, a ballot picks the stripe, four or five rounds instead of seventeen
// dependent loads.  pred is monotone (false ... false true ... true); all lanes pass the same
// arguments.  Returns the first index in [lo, hi) for which pred holds, hi if there is none.
template <typename P>
__device__ __forceinline__ uint32_t warp_first_true(uint32_t lo, uint32_t hi, P pred)
{
    const uint32_t lane = threadIdx.x & 31;
    if (lo >= hi) return hi;
    while (hi - lo > 32) {
        const uint32_t step = (hi - lo + 31) >> 5;
        const uint64_t e = (uint64_t)lo + (uint64_t)(lane + 1) * step - 1;     // last index of this lane's stripe
        const bool t = e >= hi || pred((uint32_t)e);
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, t);
        if (!m) return hi;
        const uint32_t k = __ffs(m) - 1;
        const uint64_t ek = (uint64_t)lo + (uint64_t)(k + 1) * step - 1;
        lo += k * step;
        if (ek < hi) hi = (uint32_t)ek;               // pred holds at ek: "none before it" means ek itself
    }
    const uint32_t i = lo + lane;
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, i >= hi || pred(i));
    if (!m) return hi;
    return min(lo + (uint32_t)(__ffs(m) - 1), hi);
}
template <typename T>
__device__ __forceinline__ uint32_t warp_lower_bound(const T *a, uint32_t n, T x)      // first a[i] >= x
{
    return warp_first_true(0u, n, [=](uint32_t i) { return a[i] >= x; });
}
template <typename T>
__device__ __forceinline__ uint32_t warp_upper_bound(const T *a, uint32_t n, T x)      // first a[i] > x
{
    return warp_first_true(0u, n, [=](uint32_t i) { return a[i] > x; });
}
// first packet index i >= from whose prefix count differs from prefix[from], or np
__device__ __forceinline__ uint32_t warp_first_flagged(const uint32_t *prefix, uint32_t np, uint32_t from)
{
    if (from >= np) return np;
    const uint32_t base = prefix[from];
    // prefix[i + 1] > base  <=>  some flag in [from, i]
    return warp_first_true(from, np, [=](uint32_t i) { return prefix[i + 1] > base; });
}

// One warp per track (all lanes hold the same values; the searches are shared out, lane 0 stores).
// Everything is a table lookup or a search in a sorted table.
__global__ void __launch_bounds__(64) k_track_setup(TrackSetupArgs a, TrackDev *__restrict__ tracks, uint32_t n_tracks)
{
    const uint32_t ti = blockIdx.x * 2 + (threadIdx.x >> 5);
    if (ti >= n_tracks) return;
    const uint32_t n_raw = (uint32_t)min(a.cnt->n_raw, (uint64_t)a.cap_raw), n_valid = (uint32_t)min(a.cnt->n_valid, (uint64_t)a.cap_valid);
    const uint32_t np = (uint32_t)a.cnt->np;
    TrackDev T = tracks[ti];
    T.status = 1;
    T.error_flags = 0;
    T.codec = -1;
    T.frames = 0;
    T.nseg = 0; T.ngrp = 0; T.err_seg = 0xFFFFFFFFu;
    T.pk_lo = T.pk_hi = T.pk_x = T.pcm_pk_end = 0;
    T.es_start = T.es_end = T.es_cut = 0;
    T.truncated = 0;
    T.stopped = 0;
    T.pk_open = T.pk_check = T.pk_check_end = 0;
    do {
        if (a.cnt->overflow) break;                            // a table was too small: the host comes back
        if (T.first_sector >= a.n_sectors) break;              // aob_reader_seek fails (aob.c:181-199)
        // a broken packet chain ends the stream for good (packet.c:60-116)
        uint32_t dead = a.n_sectors;
        {
            const uint32_t base = a.bad_prefix[T.first_sector];
            const uint32_t *bp = a.bad_prefix;
            dead = warp_first_true(T.first_sector, a.n_sectors, [=](uint32_t i) { return bp[i + 1] > base; });
        }
        T.pk_lo = a.sec_base[T.first_sector];
        T.pk_hi = dead < a.n_sectors ? a.sec_base[dead + 1] : np;
        if (T.pk_lo >= T.pk_hi) break;                         // no audio packet
        const uint32_t codec = a.pt.codec[T.pk_lo];
        uint32_t pk_x = (T.last_sector < a.n_sectors - 1 && T.last_sector + 1 > T.last_sector)
                            ? a.sec_base[T.last_sector + 1] : np;

        if (codec == CODEC_PCM) {
            // open_pcm_track_reader (dvd-audio.c:952-1014)
            const uint32_t prm = a.pt.params[T.pk_lo];
            T.g0_bps = (prm >> 20) & 15; T.g1_bps = (prm >> 16) & 15;
            T.g0_rate = (prm >> 12) & 15; T.g1_rate = (prm >> 8) & 15;
            T.assignment = prm & 0xFF;
            T.bits = unpack_bps(T.g0_bps);
            T.rate = unpack_rate(T.g0_rate);
            T.channels = channel_count(T.assignment);
            if (!T.channels || (T.bits != 16 && T.bits != 24)) break;
            T.codec = 0;
            T.pcm_chunk = (T.bits >> 3) * T.channels * 2;
            // frame budget: lround(pts * rate / 90000)
            // (a window of a longer PCM track: the caller keeps the books and passes the frames still to come)
            const uint64_t total = (T.cont & TRACK_PCM_FRAMES) ? (uint64_t)T.pts_length
                                                               : (uint64_t)llround((double)T.pts_length * (double)T.rate / 90000.0);
            T.pcm_frame0 = a.pk_pf[T.pk_lo];
            // stop in front of the first later packet that is not PCM / differs / is empty
            uint32_t stop = warp_first_flagged(a.pk_pcm_stop, np, T.pk_lo + 1);
            if (stop > T.pk_hi) stop = T.pk_hi;
            // ... and behind the packet in which the budget is used up: first i with
            // frames(pk_lo..i) >= total
            const uint64_t *pf = a.pk_pf;
            const uint64_t frame0 = T.pcm_frame0;
            const uint32_t lo = warp_first_true(T.pk_lo, stop, [=](uint32_t i) { return pf[i + 1] - frame0 >= total; });
            T.pcm_pk_end = lo < stop ? lo + 1 : stop;
            T.truncated = (lo >= stop && stop == np);        // budget left, buffer used up
            T.frames = a.pk_pf[T.pcm_pk_end] - T.pcm_frame0;
            T.status = 0;
        } else if (codec == CODEC_MLP) {
            if (!a.mlp_searched) { atomicOr(&a.cnt->overflow, CAP_SHAPE); break; }   // the sync search was left out: the host comes back
            // open_mlp_track_reader / locate_mlp_parameters (dvd-audio.c:1094-1149, 1318-1365)
            const uint64_t es_avail = a.pk_es[T.pk_hi];
            const uint64_t es_lo = a.pk_es[T.pk_lo];
            const uint32_t ci = warp_lower_bound(a.raw, n_raw, es_lo);
            if (ci >= n_raw || a.raw[ci] + 18 > es_avail) {
                if (T.cont & TRACK_CONT_PREV) { T.status = 0; T.codec = 1; T.truncated = 1; }   // an empty part
                break;                                               // reference asserts
            }
            const uint64_t p = a.raw[ci];
            T.es_start = p;
            const uint32_t b8 = ld_u8(a.es + p + 8), b9 = ld_u8(a.es + p + 9);
            T.g0_bps = b8 >> 4; T.g1_bps = b8 & 15; T.g0_rate = b9 >> 4; T.g1_rate = b9 & 15;
            T.assignment = ld_u8(a.es + p + 11) & 31;
            T.bits = unpack_bps(T.g0_bps);
            T.rate = unpack_rate(T.g0_rate);
            T.channels = channel_count(T.assignment);
            if (!T.channels) break;
            T.codec = 1;
            T.status = 0;
            T.nss = (p + 21 <= es_avail) ? (ld_u8(a.es + p + 20) >> 4) : 0;
            const uint32_t rc = T.g0_rate & 7;
            T.au_nominal = 40u << (rc > 2 ? 0 : rc);
            // the packet holding byte p + 17 is the last one consumed while opening
            T.pk_open = warp_upper_bound(a.pk_es, np + 1, p + 17) - 1;
            // a continued part whose range holds no sync at all is empty: the sync it found
            // lies in (and will be found again by) a later part
            const bool empty_part = (T.cont & TRACK_CONT_PREV) && pk_x < T.pk_hi && p >= a.pk_es[pk_x];
            if (pk_x < T.pk_open + 1) pk_x = T.pk_open + 1;
            // zero-yield rule: a real track start exempts the packets consumed while opening;
            // a continued part checks from its first packet on, unless the previous part's
            // last access unit ended inside that packet (then the packet did yield)
            T.pk_check = T.pk_open + 1;
            if (T.cont & TRACK_CONT_PREV)
                T.pk_check = p > es_lo ? warp_upper_bound(a.pk_es, np + 1, p - 1) : T.pk_lo;
            // end of the track
            uint64_t es_end = es_avail;
            if (pk_x < T.pk_hi) {
                const uint64_t P0 = a.pk_es[pk_x];
                if (a.pt.codec[pk_x] != CODEC_MLP) {
                    es_end = P0;                               // codec mismatch: nothing more
                } else {
                    const uint32_t cj = warp_lower_bound(a.raw, n_raw, P0);
                    if (cj < n_raw && a.raw[cj] + 8 <= es_avail) es_end = a.raw[cj];
                    else if (T.cont & TRACK_CONT_NEXT) { es_end = es_avail; T.truncated = 1; }   // the parts behind are empty
                    else { es_end = (es_avail >= P0 + 8) ? es_avail - 7 : P0; T.truncated = 1; }
                }
            } else {
                T.truncated = 1;                               // no packet behind last_sector in the buffer
            }
            // a non-MLP audio packet met while decoding ends the stream (dvd-audio.c:1203-1208)
            const uint32_t nm = warp_first_flagged(a.pk_nonmlp, np, (T.cont & TRACK_CONT_PREV) ? T.pk_lo : T.pk_open + 1);
            if (nm < pk_x && nm < T.pk_hi && a.pk_es[nm] < es_end) es_end = a.pk_es[nm];
            if (es_end < p) es_end = p;
            T.es_end = es_end;
            // a part that is continued also answers for the packets its last access units end in
            T.pk_check_end = pk_x;
            if ((T.cont & TRACK_CONT_NEXT) && pk_x < T.pk_hi && es_end > a.pk_es[pk_x])
                T.pk_check_end = min(T.pk_hi, warp_upper_bound(a.pk_es, np + 1, es_end - 1));
            T.es_cut = es_end;
            if (empty_part) T.es_end = T.es_cut = p;
            if (T.nss != 1 && T.nss != 2) {
                // the first access unit is not a usable major sync: nothing decodes
                T.error_flags |= ERR_SYNTAX;
                T.es_end = T.es_cut = p;
            }
            // A continued part has to begin where a segment begins: at a major sync whose substreams
            // open with a restart header.  If the first sync behind the cut is not one (a stream may
            // carry major syncs without restart headers), the decoder state of the part before is
            // needed: the caller has to decode the two parts as one.
            if ((T.cont & TRACK_CONT_PREV) && !empty_part) {
                const uint32_t vi = warp_lower_bound(a.valid, n_valid, p);
                if (!(vi < n_valid && a.valid[vi] == p)) { T.stopped = 2; T.es_end = T.es_cut = p; }
            }
            // restart segments: the start plus every segment-starting sync inside (p, es_end)
            T.cand_lo = warp_upper_bound(a.valid, n_valid, p);
            const uint32_t cand_hi = warp_lower_bound(a.valid, n_valid, T.es_end);
            T.nseg = T.es_end > p ? 1 + (cand_hi > T.cand_lo ? cand_hi - T.cand_lo : 0) : 0;
            T.ngrp = (T.nseg + DVDA_LANES - 1) / DVDA_LANES;
        }
        T.pk_x = pk_x;
    } while (0);
    if ((threadIdx.x & 31) == 0) tracks[ti] = T;
}

int launch_track_setup(TrackSetupArgs a, TrackDev *tracks, uint32_t n_tracks, cudaStream_t s)
{
    LAUNCH(k_track_setup, div_up_u32(n_tracks, 2), 64, 0, s, a, tracks, n_tracks);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// ----------------------------------------------------------- the plan
// One block: the tracks' places in the segment and group tables (prefix sums over the tracks,
// in order), the work list of the per-segment passes, totals and flags for the host — what the
// host used to compute between two round trips.
#define PLAN_THREADS 256
__device__ __forceinline__ void plan_check(DecCounts *__restrict__ cnt, const PlanLimits &lim);
__global__ void __launch_bounds__(PLAN_THREADS)
k_track_plan(TrackDev *__restrict__ tracks, uint32_t n_tracks, uint32_t *__restrict__ trk_pk_lo,
             uint32_t *__restrict__ trk_seg_base, uint32_t *__restrict__ trk_grp_base,
             DecWork *__restrict__ work, OutWork *__restrict__ out_work, uint32_t cap_work, uint32_t cap_seg, uint32_t cap_grp,
             uint32_t cap_sync, DecCounts *__restrict__ cnt, uint32_t check_now, PlanLimits lim)
{
    __shared__ uint32_t s_flags[4];
    if (threadIdx.x < 4) s_flags[threadIdx.x] = 0;
    __syncthreads();
    uint64_t seg_carry = 0, grp_carry = 0, pair_carry = 0, row_carry = 0, pcm_carry = 0, ow_carry = 0, orow_carry = 0;
    for (uint32_t base = 0; base < n_tracks; base += PLAN_THREADS) {
        const uint32_t i = base + threadIdx.x;
        uint32_t nseg = 0, ngrp = 0, nss = 0, out_warps = 0, n0 = 0, n1 = 0;
        uint64_t pcm = 0;
        if (i < n_tracks) {
            TrackDev &T = tracks[i];
            const bool mlp = T.status == 0 && T.codec == 1;
            if (!mlp) { T.nseg = 0; T.ngrp = 0; }
            nseg = T.nseg; ngrp = T.ngrp;
            nss = (mlp && nseg) ? T.nss : 0;
            if (T.status == 0 && T.codec == 0) { pcm = T.frames * T.channels; atomicOr(&s_flags[0], 1u); }
            if (mlp && nseg) {
                atomicOr(&s_flags[1], 1u);
                atomicMax(&s_flags[2], T.nss);
                atomicOr(&s_flags[3], 1u << T.channels);
            }
            trk_pk_lo[i] = T.pk_lo;
            // the output pass: lanes = (segment, channel) over both substreams, 32 / channels segments per warp
            if (nss && (T.nss == 1 ? (T.channels >= 1 && T.channels <= 4) : (T.nss == 2 && T.channels >= 3 && T.channels <= 6))) {
                n0 = T.nss == 1 ? T.channels : 2; n1 = T.channels - n0;
                const uint32_t spw = out_segs_per_warp(T.channels);
                out_warps = ngrp * ((32 + spw - 1) / spw);
            }
        }
        // (packed: two sums per scan)
        uint64_t tot_a, tot_b, tot_c;
        const uint64_t ex_a = block_excl_scan<PLAN_THREADS>((uint64_t)nseg | (uint64_t)ngrp << 32, &tot_a);
        const uint64_t ex_b = block_excl_scan<PLAN_THREADS>((uint64_t)(ngrp * nss) | (uint64_t)nss << 32, &tot_b);
        block_excl_scan<PLAN_THREADS>(pcm, &tot_c);
        uint64_t tot_d;
        const uint64_t ex_d = block_excl_scan<PLAN_THREADS>((uint64_t)out_warps | (uint64_t)(out_warps ? 1u : 0u) << 32, &tot_d);
        if (out_warps) {
            const uint32_t orow = (uint32_t)(orow_carry + (ex_d >> 32));
            OutWork w = {(uint32_t)(ow_carry + (ex_d & 0xFFFFFFFFu)), i, n0, n1};
            out_work[orow] = w;                            // (at most one row per track: the list has n_tracks rows)
        }
        ow_carry += tot_d & 0xFFFFFFFFu; orow_carry += tot_d >> 32;
        if (i < n_tracks) {
            TrackDev &T = tracks[i];
            const uint32_t seg_base = (uint32_t)(seg_carry + (ex_a & 0xFFFFFFFFu));
            const uint32_t grp_base = (uint32_t)(grp_carry + (ex_a >> 32));
            T.seg_base = seg_base; T.grp_base = grp_base;
            trk_seg_base[i] = seg_base; trk_grp_base[i] = grp_base;
            uint32_t warp0 = (uint32_t)(pair_carry + (ex_b & 0xFFFFFFFFu));
            uint32_t row = (uint32_t)(row_carry + (ex_b >> 32));
            for (uint32_t k = 0; k < nss; k++, row++, warp0 += ngrp) {
                // substream 0 of a two-substream stream carries the stereo pair (DVD-Audio layout)
                const uint32_t nch = T.nss == 1 ? T.channels : (k == 0 ? 2 : T.channels - 2);
                if (row < cap_work) { DecWork w = {warp0, i, k, nch}; work[row] = w; }
            }
        }
        seg_carry += tot_a & 0xFFFFFFFFu; grp_carry += tot_a >> 32;
        pair_carry += tot_b & 0xFFFFFFFFu; row_carry += tot_b >> 32;
        pcm_carry += tot_c;
    }
    if (threadIdx.x == 0) {
        trk_seg_base[n_tracks] = (uint32_t)seg_carry; trk_grp_base[n_tracks] = (uint32_t)grp_carry;
        uint32_t over = cnt->overflow;
        if (cnt->n_raw > cap_sync || cnt->n_valid > cap_sync) { over |= CAP_SYNC; cnt->need_sync = max(cnt->n_raw, cnt->n_valid); }
        if (seg_carry > cap_seg) { over |= CAP_SEG; cnt->need_seg = seg_carry; }
        if (grp_carry > cap_grp || row_carry > cap_work) { over |= CAP_GRP; cnt->need_grp = grp_carry; }
        cnt->overflow = over;
        // with a table too small nothing of the MLP side runs: the host comes back
        const bool stop = (over & (CAP_SYNC | CAP_SEG | CAP_GRP | CAP_ROWS)) != 0;
        cnt->nseg = stop ? 0u : (uint32_t)seg_carry;
        cnt->ngroups = stop ? 0u : (uint32_t)grp_carry;
        cnt->npairs = stop ? 0u : (uint32_t)pair_carry;
        cnt->nwork = stop ? 0u : (uint32_t)row_carry;
        cnt->nout_work = stop ? 0u : (uint32_t)orow_carry;
        cnt->nout_warps = stop ? 0u : (uint32_t)ow_carry;
        cnt->pcm_fixed = pcm_carry + 4ull * n_tracks + 64;
        cnt->any_pcm = s_flags[0]; cnt->any_mlp = s_flags[1]; cnt->nss_max = s_flags[2] ? s_flags[2] : 1; cnt->chan_mask = s_flags[3];
        cnt->max_au = 0; cnt->max_chunks = 0;
        if (check_now) plan_check(cnt, lim);
    }
}

__device__ __forceinline__ void plan_check(DecCounts *__restrict__ cnt, const PlanLimits &lim)
{
    uint32_t over = cnt->overflow;
    if (cnt->nau > lim.cap_au) { over |= CAP_AU; cnt->need_au = cnt->nau; }
    if (cnt->cells > lim.cap_cells) { over |= CAP_CELLS; cnt->need_cells = cnt->cells; }
    if (cnt->max_au > lim.max_au) over |= CAP_MAX_AU;
    // kernels that were left out, or sized for less, because the previous decode had no use for them
    if ((cnt->any_pcm && !lim.pcm) || (cnt->any_mlp && !lim.mlp) || (cnt->any_mlp && cnt->nss_max > lim.nss) ||
        cnt->nout_warps > lim.out_warps) over |= CAP_SHAPE;
    cnt->overflow = over;
    if (over & (CAP_AU | CAP_CELLS | CAP_MAX_AU | CAP_SHAPE)) {
        // nothing may be written to tables that are too small, nothing read that was not written: the MLP side stops here
        cnt->nau = 0; cnt->nseg = 0; cnt->ngroups = 0; cnt->npairs = 0; cnt->nwork = 0; cnt->nout_work = 0; cnt->nout_warps = 0;
    }
}

__global__ void k_plan_check(DecCounts *__restrict__ cnt, PlanLimits lim) { plan_check(cnt, lim); }

int launch_track_plan(TrackDev *tracks, uint32_t n_tracks, uint32_t *trk_pk_lo, uint32_t *trk_seg_base, uint32_t *trk_grp_base,
                      DecWork *work, OutWork *out_work, uint32_t cap_work, uint32_t cap_seg, uint32_t cap_grp, uint32_t cap_sync,
                      DecCounts *cnt, const PlanLimits *check_now, cudaStream_t s)
{
    const PlanLimits none = {};
    LAUNCH(k_track_plan, 1, PLAN_THREADS, 0, s, tracks, n_tracks, trk_pk_lo, trk_seg_base, trk_grp_base, work, out_work, cap_work, cap_seg, cap_grp, cap_sync, cnt,
           check_now ? 1u : 0u, check_now ? *check_now : none);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
int launch_plan_check(DecCounts *cnt, PlanLimits lim, cudaStream_t s)
{
    LAUNCH(k_plan_check, 1, 1, 0, s, cnt, lim);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------- segments and AUs

// one thread per segment: where it starts and where it must end
__global__ void k_segment_fill(const TrackDev *__restrict__ tracks, uint32_t n_tracks,
                               const uint32_t *__restrict__ trk_seg_base, const uint64_t *__restrict__ valid,
                               SegDev *__restrict__ segs, const DecCounts *__restrict__ cnt)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt->nseg) return;
    const uint32_t t = upper_bound_dev(trk_seg_base, n_tracks, i) - 1;
    const TrackDev &T = tracks[t];
    const uint32_t j = i - trk_seg_base[t];
    SegDev S;
    S.es_pos = j == 0 ? T.es_start : valid[T.cand_lo + j - 1];
    S.es_limit = j + 1 == T.nseg ? T.es_end : valid[T.cand_lo + j];
    S.track = t;
    S.n_au = 0; S.au_base = 0; S.flags = 0; S.frames = 0; S.err = 0; S.err_au = 0xFFFFFFFFu; S.pad = 0; S.frame0 = 0;
    segs[i] = S;
}

int launch_segment_fill(const TrackDev *tracks, uint32_t n_tracks, const uint32_t *trk_seg_base,
                        const uint64_t *valid, SegDev *segs, uint32_t cap_seg, const DecCounts *cnt, cudaStream_t s)
{
    if (!cap_seg) return 0;
    LAUNCH(k_segment_fill, div_up_u32(cap_seg, 128), 128, 0, s, tracks, n_tracks, trk_seg_base, valid, segs, cnt);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// One thread per segment walks the chain of 12-bit access-unit lengths
// (reference mlp.c:392-394).  Pass 1 counts and notes the first AU_NOTED positions of the
// segment (relative to its start; noted[j * nseg + segment]: coalesced).  Pass 2 (fill), once
// the access units are numbered, copies them to their places — independent loads, not a second
// walk down the chain; only a segment with more access units than were noted walks on from the
// last noted one.
#define AU_NOTED 32
__global__ void k_au_chase(const uint8_t *__restrict__ es, SegDev *__restrict__ segs, uint32_t nseg /* rows: stride of `noted` */,
                           const DecCounts *__restrict__ cnt,
                           const TrackDev *__restrict__ tracks, uint32_t *__restrict__ seg_nau,
                           uint64_t *__restrict__ au_pos, uint32_t *__restrict__ au_seg,
                           const uint32_t *__restrict__ seg_au_base, uint32_t *__restrict__ noted, uint32_t n_noted, int fill)
{
    // (n_noted: positions noted per segment, AU_NOTED unless a test wants the walk-on path)
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nseg) return;
    if (i >= cnt->nseg) { if (!fill) seg_nau[i] = 0; return; }         // rows behind the last segment count nothing
    if (fill && cnt->nau == 0) return;                                 // (the access-unit tables were too small: nothing is filled)
    SegDev &S = segs[i];
    const uint64_t limit = S.es_limit;
    const uint64_t pos0 = S.es_pos;
    if (fill) {
        const uint32_t base = seg_au_base[i], n_au = S.n_au;
        S.au_base = base;
        const uint32_t copied = min(n_au, n_noted);
#pragma unroll 8
        for (uint32_t j = 0; j < copied; j++) {
            au_pos[base + j] = pos0 + noted[(uint64_t)j * nseg + i];
            au_seg[base + j] = i;
        }
        if (n_au > n_noted) {
            uint64_t pos = pos0 + noted[(uint64_t)(n_noted - 1) * nseg + i];
            for (uint32_t n = n_noted - 1; n < n_au; n++) {
                if (n >= n_noted) { au_pos[base + n] = pos; au_seg[base + n] = i; }
                pos += (((ld_u8(es + pos) & 15u) << 8) | ld_u8(es + pos + 1)) * 2;
            }
        }
        return;
    }
    uint64_t pos = pos0;
    uint32_t n = 0;
    bool stalled = false;
    while (pos + 4 <= limit) {
        const uint32_t total = (((ld_u8(es + pos) & 15u) << 8) | ld_u8(es + pos + 1)) * 2;
        if (total < 4) { stalled = true; break; }              // the reference's queue never advances again
        if (pos + total > limit) break;                        // incomplete (end of track) or overshoot
        if (n < n_noted) noted[(uint64_t)n * nseg + i] = (uint32_t)(pos - pos0);
        pos += total;
        n++;
    }
    seg_nau[i] = n;
    S.n_au = n;
    const TrackDev &T = tracks[S.track];
    const bool last = (i + 1 - T.seg_base) == T.nseg;
    // every segment but the last must land exactly on the next one; so must the last
    // one of a part that is continued (the cut has to be a real access-unit boundary)
    const bool must_land = !last || ((T.cont & TRACK_CONT_NEXT) && !T.truncated);
    if (stalled || (must_land && pos != limit)) S.flags |= SEG_IRREGULAR;
}

size_t au_noted_bytes(uint32_t nseg) { return (size_t)AU_NOTED * nseg * sizeof(uint32_t); }

int launch_au_chase(const uint8_t *es, SegDev *segs, uint32_t cap_seg, const DecCounts *cnt, const TrackDev *tracks,
                    uint32_t *seg_nau, uint64_t *au_pos, uint32_t *au_seg, const uint32_t *seg_au_base,
                    uint32_t *noted, int fill, cudaStream_t s)
{
    if (!cap_seg) return 0;
    const uint32_t n_noted = getenv("DVDAGPU_SMALL_TABLES") ? 2u : (uint32_t)AU_NOTED;          // (test hook)
    LAUNCH(k_au_chase, div_up_u32(cap_seg, 128), 128, 0, s, es, segs, cap_seg, cnt, tracks, seg_nau, au_pos, au_seg, seg_au_base, noted, n_noted, fill);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// ---- "a packet that completes no access unit ends the stream" ---------------
// dvda_read() stops at the first decode call that returns no frames
// (dvd-audio.c:766-775); decode_mlp_audio is called once per packet, so a packet
// in which no access unit ends terminates the track.  Mark the packets in which
// some access unit ends, then find the first unmarked one per track.

__device__ __forceinline__ bool au_yields(const uint8_t *p, uint32_t total, const TrackDev &T);
__global__ void k_yield_mark(MlpTables m, const uint32_t *__restrict__ seg_au_base, uint8_t *__restrict__ pk_yield)
{
    // The access units of a warp follow each other in the stream and end within a handful of
    // packets: the warp searches the packet of its first access unit together (32 probes and a
    // ballot per round), the lanes walk on from there.
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    const bool have = a < m.cnt->nau;
    const uint32_t np = (uint32_t)m.cnt->np;
    if (!__any_sync(0xFFFFFFFFu, have)) return;
    uint64_t last_byte = 0;
    bool yields = false;
    if (have) {
        const uint64_t pos = m.au_pos[a];
        const uint8_t *p = m.es + pos;
        const uint32_t total = (((ld_u8(p) & 15u) << 8) | ld_u8(p + 1)) * 2;
        last_byte = pos + total - 1;
        // An access unit whose major sync states other stream parameters than the track's is dropped
        // (mlp.c:449-455): it yields no frames, so it does not keep its packet alive either.
        yields = au_yields(p, total, m.tracks[m.segs[m.au_seg[a]].track]);
    }
    const uint64_t x0 = __shfl_sync(0xFFFFFFFFu, last_byte, 0);          // lane 0 has the lowest access unit
    const uint64_t *pk_es = m.pk_es;
    uint32_t pk = warp_first_true(0u, np + 1, [=](uint32_t i) { return pk_es[i] > x0; }) - 1;
    if (have) {
        if (last_byte < x0) pk = upper_bound_dev(m.pk_es, np + 1, last_byte) - 1;   // (not in stream order: on its own)
        else {
            // a few packets on, as a rule; behind a track boundary inside the warp there may be thousands
            // of rows without stream bytes (a PCM track) in between: search then
            uint32_t steps = 0;
            while (pk + 1 <= np && pk_es[pk + 1] <= last_byte) {
                pk++;
                if (++steps == 8) { pk = upper_bound_dev(m.pk_es, np + 1, last_byte) - 1; break; }
            }
        }
        if (yields) pk_yield[pk] = 1;
    }
    (void)seg_au_base;
}

// does the access unit at p (total bytes) yield frames, i.e. is it not one of those dropped for a
// major sync that states other stream parameters than the track's (mlp.c:449-455)?
__device__ __forceinline__ bool au_yields(const uint8_t *p, uint32_t total, const TrackDev &T)
{
    if (total < 32 || ld_be32(p + 4) != 0xF8726FBBu) return true;
    const uint32_t ns = ld_u8(p + 20) >> 4;
    if (ns != 1 && ns != 2) return true;
    const uint32_t b8 = ld_u8(p + 8), b9 = ld_u8(p + 9), asg = ld_u8(p + 11) & 31;
    return !((b8 >> 4) != T.g0_bps || (b8 & 15) != T.g1_bps || (b9 >> 4) != T.g0_rate || (b9 & 15) != T.g1_rate || asg != T.assignment);
}

// A part that is continued shares the packet its cut lies in with the part behind it: access units of
// that part may end in the packet too, and keep it alive.  One thread per track walks on from the cut
// while the access units still end inside that packet.
__global__ void k_yield_tail(MlpTables m, uint8_t *__restrict__ pk_yield)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m.n_tracks) return;
    const TrackDev &T = m.tracks[t];
    if (T.status != 0 || T.codec != 1 || !(T.cont & TRACK_CONT_NEXT) || T.truncated || T.es_end <= T.es_start) return;
    const uint32_t np = (uint32_t)m.cnt->np;
    const uint32_t pk = upper_bound_dev(m.pk_es, np + 1, T.es_end - 1) - 1;      // the packet of the last byte in front of the cut
    if (pk >= np) return;
    const uint64_t pk_end = m.pk_es[pk + 1];
    uint64_t pos = T.es_end;
    const uint64_t avail = m.pk_es[min(T.pk_hi, np)];
    while (pos + 4 <= pk_end && pos + 4 <= avail) {
        const uint32_t total = (((ld_u8(m.es + pos) & 15u) << 8) | ld_u8(m.es + pos + 1)) * 2;
        if (total < 4 || pos + total > pk_end) break;
        if (au_yields(m.es + pos, total, T)) { pk_yield[pk] = 1; break; }
        pos += total;
    }
}

__global__ void k_yield_find(MlpTables m, PacketTable pt, const uint32_t *__restrict__ trk_pk_lo,
                             const uint8_t *__restrict__ pk_yield)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.cnt->np || pk_yield[i] || pt.codec[i] != CODEC_MLP) return;
    // Tracks are sorted by first packet.  Normally the owner is the last track that
    // starts at or before packet i; parts of one long track may also answer for the
    // first packets of the parts behind them, so look a few tracks back as well.
    const uint32_t t = upper_bound_dev(trk_pk_lo, m.n_tracks, i);
    for (uint32_t tt = t, steps = 0; tt > 0 && steps < 64; tt--, steps++) {
        TrackDev &T = m.tracks[tt - 1];
        if (T.status != 0 || T.codec != 1) continue;
        // only packets handed to decode_mlp_audio inside the track's sector range count
        if (i < T.pk_check || i >= T.pk_check_end || i >= T.pk_hi) continue;
        atomicMin((unsigned long long *)&T.es_cut, (unsigned long long)m.pk_es[i]);
    }
}

int launch_yield(MlpTables m, uint32_t rows, const uint32_t *seg_au_base, PacketTable pt, const uint32_t *trk_pk_lo,
                 uint8_t *pk_yield, bool any_parts, cudaStream_t s)
{
    if (!rows) return 0;
    CUDA_TRY(cudaMemsetAsync(pk_yield, 0, rows, s));
    if (m.cap_au) LAUNCH(k_yield_mark, div_up_u32(m.cap_au, 256), 256, 0, s, m, seg_au_base, pk_yield);
    if (any_parts) LAUNCH(k_yield_tail, div_up_u32(m.n_tracks, 128), 128, 0, s, m, pk_yield);
    LAUNCH(k_yield_find, div_up_u32(rows, 256), 256, 0, s, m, pt, trk_pk_lo, pk_yield);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------ groups

// A group is up to 32 consecutive segments of one track: the 32 lanes of the
// decoding warp.  Its tile holds `cap` frames per segment.
__global__ void k_group_setup(const TrackDev *__restrict__ tracks, uint32_t n_tracks,
                              const uint32_t *__restrict__ trk_grp_base, const SegDev *__restrict__ segs,
                              GroupDev *__restrict__ groups, uint32_t cap_grp, DecCounts *__restrict__ cnt,
                              uint32_t *__restrict__ grp_cells, const uint32_t *__restrict__ seg_need)
{
    // one warp per group, lane = segment of the group
    const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (g >= cap_grp) return;
    if (g >= cnt->ngroups) { if (lane == 0) grp_cells[g] = 0; return; }  // rows behind the last group count nothing
    const uint32_t t = upper_bound_dev(trk_grp_base, n_tracks, g) - 1;
    const TrackDev &T = tracks[t];
    const uint32_t j = g - trk_grp_base[t];
    GroupDev G;
    G.track = t;
    G.seg0 = T.seg_base + j * DVDA_LANES;
    G.nseg = min((uint32_t)DVDA_LANES, T.nseg - j * DVDA_LANES);
    uint32_t need = 0, n_au = 0;
    if (lane < G.nseg) {
        const SegDev &S = segs[G.seg0 + lane];
        n_au = S.n_au;
        // (another attempt after a tile overflow: the frame counts the previous one found)
        need = max(n_au * T.au_nominal, seg_need[G.seg0 + lane]);
    }
    const uint32_t cap = __reduce_max_sync(0xFFFFFFFFu, need), most = __reduce_max_sync(0xFFFFFFFFu, n_au);
    if (lane) return;
    G.cap = cap;
    G.tile_off = 0; G.byp_off = 0;
    groups[g] = G;
    grp_cells[g] = cap * T.channels;
    atomicMax(&cnt->max_au, most);
    atomicMax(&cnt->max_chunks, (cap + 31) / 32);        // most 32-frame chunks in a group
}

// ... and, on the way, the start values of the fast path's per-segment tables (two stream
// operations fewer in the chain): zero16 = 16-byte pieces to clear (segment contexts exist only
// where pass A0 goes), flags = words to set to SEG_FALLBACK (every segment starts out flagged:
// substreams with more than 4 channels are not visited by the fast path at all)
__global__ void k_group_offsets(GroupDev *__restrict__ groups, const DecCounts *__restrict__ cnt, const uint64_t *__restrict__ cell_base,
                                uint4 *__restrict__ zero16, uint32_t n_zero16, uint32_t *__restrict__ flags, uint32_t n_flags)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x, step = gridDim.x * blockDim.x;
    for (uint32_t i = g; i < n_zero16; i += step) zero16[i] = make_uint4(0, 0, 0, 0);
    for (uint32_t i = g; i < n_flags; i += step) flags[i] = SEG_FALLBACK;
    if (g >= cnt->ngroups) return;
    groups[g].tile_off = cell_base[g] * DVDA_LANES;
    groups[g].byp_off = cell_base[g] * DVDA_LANES;
}

int launch_group_setup(const TrackDev *tracks, uint32_t n_tracks, const uint32_t *trk_grp_base, const SegDev *segs,
                       GroupDev *groups, uint32_t cap_grp, DecCounts *cnt, uint32_t *grp_cells, const uint32_t *seg_need, cudaStream_t s)
{
    if (!cap_grp) return 0;
    LAUNCH(k_group_setup, div_up_u32((uint64_t)cap_grp * 32, 128), 128, 0, s, tracks, n_tracks, trk_grp_base, segs, groups, cap_grp, cnt, grp_cells, seg_need);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
int launch_group_offsets(GroupDev *groups, uint32_t cap_grp, const DecCounts *cnt, const uint64_t *cell_base,
                         void *zero, size_t zero_bytes, uint32_t *flags, uint32_t n_flags, cudaStream_t s)
{
    if (!cap_grp && !zero_bytes && !n_flags) return 0;
    const uint32_t n16 = (uint32_t)(zero_bytes / 16);          // (the caller's tables are whole 16-byte pieces)
    const uint32_t work = max(max(cap_grp, n16), n_flags);
    LAUNCH(k_group_offsets, min(div_up_u32(work, 128), 148u * 8), 128, 0, s, groups, cnt, cell_base, static_cast<uint4 *>(zero), n16, flags, n_flags);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
