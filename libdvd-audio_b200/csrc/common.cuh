// common.cuh — shared definitions of the sm_100a DVD-Audio decode engine.
//
// Data model (all arrays live in one device arena owned by the context):
//
//   sectors[n_sectors * 2048]      the AOB bytes as uploaded
//   packet table (one row per 0xBD audio packet, in sector order)
//   ES[es_total]                   all MLP payload bytes, concatenated
//   sync lists                     raw (pattern only) and valid (restart segments)
//   segment table                  one row per restart-delimited segment
//   AU table                       one row per MLP access unit
//   tiles                          filtered samples, layout [group][frame][channel][lane]
//   pcm                            final interleaved int32, track after track
//
// Vocabulary follows the reference: sector, pack, packet, elementary stream
// (ES), access unit (AU), major sync, substream, block, restart header.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define DVDA_SECTOR 2048u
#define DVDA_MAX_CH 8
#define DVDA_MAX_MAT 6
#define DVDA_LANES 32          // segments per group = lanes per warp
#define DVDA_ES_PAD 65536      // zero bytes behind the ES so that bit readers may run ahead

#define CODEC_PCM 0xA0
#define CODEC_MLP 0xA1

// error bits (mirror include/dvdagpu.h)
#define ERR_PARITY (1 << 4)
#define ERR_CRC (1 << 5)
#define ERR_SYNTAX (1 << 6)

// segment flags
#define SEG_IRREGULAR 1u       // access-unit chain did not land on the next sync
#define SEG_NEEDS_CARRY 2u     // FIR history of the previous segment is needed
#define SEG_OVERFLOW 4u        // more frames than the tile has room for
#define SEG_FALLBACK 8u        // the three-pass fast path gave up: decode with the complete decoder
#define SEG_WANTS_PREV 16u     // ... because it needs the previous segment's FIR history
#define STATUS_WANTS_REMATRIX 0x100u   // k_seg_finalize -> host: some segment still needs k_rematrix (fast path)
#define STATUS_PCM_SMALL 0x200u        // k_track_out_base -> output pass, host: the samples do not fit the buffer sized in advance

// TrackDev.cont / dvdagpu_track_desc.flags
#define TRACK_CONT_PREV 1u
#define TRACK_CONT_NEXT 2u
#define TRACK_PCM_FRAMES 4u    // PCM: pts_length holds the frame budget itself (DVDAGPU_PCM_BUDGET_IN_FRAMES)

// What the demux and index stages find out about the input, kept in device memory: the launch
// sequence of a decode is fixed before any of it is known to the host.  Every table is sized in
// advance (from the input's size, or from what the previous decode of this context needed), the
// kernels take their bounds from here and their grids are sized for the tables' capacities; the
// host reads this once, together with the results, when the decode is over, and repeats the
// decode with larger tables if one was too small (`overflow`).
struct DecCounts {
    uint64_t np;               // audio packets (total of the per-sector counts)
    uint64_t es_total;         // elementary-stream bytes
    uint64_t n_raw, n_valid;   // sync patterns / those that start a restart segment
    uint64_t nau;              // access units
    uint64_t cells;            // tile cells (frames x channels per group, summed)
    uint64_t pcm_fixed;        // samples of the PCM tracks + alignment slack of the output buffer
    uint32_t nseg, ngroups;
    uint32_t npairs;           // (group, substream) pairs: warps of the per-segment passes
    uint32_t nwork;            // rows of the work list: (track, substream) pairs with segments
    uint32_t nout_work;        // rows of the output pass's work list: tracks it takes
    uint32_t nout_warps;       // its warps
    uint32_t max_au;           // most access units in a segment
    uint32_t max_chunks;       // most 32-frame chunks in a group
    uint32_t any_pcm, any_mlp, nss_max;
    uint32_t chan_mask;        // bit n: some MLP track has n channels
    uint32_t overflow;         // CAP_* bits
    uint32_t pad[3];
    // what a table that overflowed would have needed (the counts above are zeroed then, so that
    // the stages behind do nothing)
    uint64_t need_rows, need_sync, need_seg, need_grp, need_au, need_cells;
};
#define CAP_ROWS 1u            // packet table
#define CAP_SYNC 2u            // sync lists
#define CAP_SEG 4u             // segment table
#define CAP_GRP 8u             // group table, work list
#define CAP_AU 16u             // access-unit tables
#define CAP_CELLS 32u          // tiles
#define CAP_MAX_AU 64u         // grid of the entropy pass (access units per segment)
#define CAP_SHAPE 256u         // a kernel that was left out has work after all (PCM tracks, two substreams, ...)

struct TrackDev {
    // inputs
    uint32_t first_sector, last_sector, pts_length;
    uint32_t cont;             // bit 0: continues a previous part, bit 1: is continued by a next part
    // results of track setup
    int32_t status, error_flags, codec;
    uint32_t g0_bps, g1_bps, g0_rate, g1_rate, assignment;
    uint32_t channels, bits, rate;
    uint32_t pk_lo;            // first audio packet of the track
    uint32_t pk_hi;            // first packet no longer reachable (dead sector / end)
    uint32_t pk_x;             // first audio packet in a sector > last_sector
    // MLP
    uint64_t es_start, es_end; // elementary-stream byte range [start, end)
    uint64_t es_cut;           // zero-yield packet rule: no access unit may end behind this
    uint32_t pk_open;          // last packet consumed while opening the reader
    uint32_t pk_check;         // packets [pk_check, pk_check_end) are subject to the zero-yield rule
    uint32_t pk_check_end;
    uint32_t stopped;          // 1: ended before its natural end, 2: needs the previous part's FIR history
    uint32_t nss;              // substreams
    uint32_t au_nominal;       // expected frames per access unit (tile sizing only)
    uint32_t cand_lo, nseg;    // valid syncs after es_start, segments
    uint32_t seg_base, grp_base, ngrp;
    uint32_t err_seg;          // first segment (track-relative) that hit an error
    uint32_t truncated;        // the sector buffer ended before the track did
    // PCM
    uint32_t pcm_chunk;        // bytes per chunk (two frames)
    uint32_t pcm_pk_end;       // packets [pk_lo, pcm_pk_end) are unpacked
    uint64_t pcm_frame0;       // scanned frame count in front of pk_lo
    // output
    uint64_t frames;
    uint64_t out_base;         // first sample in the pcm buffer
};

struct SegDev {
    uint64_t es_pos;           // first byte of the segment's first access unit
    uint64_t es_limit;         // next segment / end of track
    uint32_t track;
    uint32_t n_au;             // complete access units in [es_pos, es_limit)
    uint32_t au_base;          // first row in the AU table
    uint32_t flags;
    uint32_t frames;           // frames decoded (by substream 0)
    uint32_t err;              // ERR_* bits met in this segment
    uint32_t err_au;           // access unit (segment-relative) in front of which decoding stopped
    uint32_t pad;
    uint64_t frame0;           // frames of the track in front of this segment
};

struct GroupDev {
    uint32_t track;
    uint32_t seg0;             // first segment (global index)
    uint32_t nseg;             // 1..32
    uint32_t cap;              // frames per segment the tile has room for
    uint64_t tile_off;         // offset of the tile in int32 units: [cap][channels][32]
    uint64_t byp_off;          // offset of the bypass-bit tile in bytes: [cap][32]
};

// rematrix parameters in force at the end of an access unit (reference mlp.c:504-525)
struct alignas(16) ParamSet {
    int16_t coeff[DVDA_MAX_MAT][DVDA_MAX_CH];   // a matrix row = one 16-byte load
    uint8_t out_ch[DVDA_MAX_MAT];               // } one 8-byte load
    uint8_t matrix_len, mmc;                    // }
    uint8_t q[DVDA_MAX_CH];
    uint8_t out_shift[DVDA_MAX_CH];
    uint8_t noise_shift, uses_noise;
    uint8_t pad[6];
};
static_assert(sizeof(ParamSet) == 128, "ParamSet: 128 bytes, vector loads of its parts");

// Segments a warp of the output pass takes: lanes = (segment, channel).  (At most 16: the warp
// keeps 8 words per segment about its rows.)
__host__ __device__ __forceinline__ uint32_t out_segs_per_warp(uint32_t channels) { return channels <= 1 ? 16u : 32u / channels; }

struct AuDev {
    uint32_t frame0;           // first frame, segment-relative
    uint32_t nframes;
    uint32_t seed;             // noise generator state at the start of the AU
    uint32_t pset;             // row of the ParamSet table (global AU index that wrote it)
};

#define CUDA_TRY(expr)                                                         \
    do {                                                                       \
        cudaError_t e_ = (expr);                                               \
        if (e_ != cudaSuccess) {                                               \
            dvdagpu_set_error("%s failed: %s (%s:%d)", #expr,                  \
                              cudaGetErrorString(e_), __FILE__, __LINE__);     \
            return -1;                                                         \
        }                                                                      \
    } while (0)

void dvdagpu_set_error(const char *fmt, ...);

// ---- device helpers -------------------------------------------------------

__device__ __forceinline__ uint32_t ld_be32_aligned(const uint32_t *p)
{
    return __byte_perm(__ldg(p), 0, 0x0123);
}

// bytes are fetched one by one where alignment is unknown (headers only)
__device__ __forceinline__ uint32_t ld_u8(const uint8_t *p) { return __ldg(p); }
__device__ __forceinline__ uint32_t ld_be16(const uint8_t *p) { return (ld_u8(p) << 8) | ld_u8(p + 1); }
__device__ __forceinline__ uint32_t ld_be32(const uint8_t *p)
{
    return (ld_u8(p) << 24) | (ld_u8(p + 1) << 16) | (ld_u8(p + 2) << 8) | ld_u8(p + 3);
}

// first index in sorted a[0..n) with a[i] >= x
template <typename T>
__device__ __forceinline__ uint32_t lower_bound_dev(const T *a, uint32_t n, T x)
{
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (a[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}
// first index with a[i] > x
template <typename T>
__device__ __forceinline__ uint32_t upper_bound_dev(const T *a, uint32_t n, T x)
{
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (a[mid] <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// ---- block-wide exclusive scan (one value per thread) -------------------
__device__ __forceinline__ uint64_t warp_incl_scan(uint64_t v)
{
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint64_t o = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= (unsigned)d) v += o;
    }
    return v;
}

// exclusive scan of one value per thread across the block; returns the
// exclusive prefix, *total gets the block sum
template <int THREADS>
__device__ __forceinline__ uint64_t block_excl_scan(uint64_t v, uint64_t *total)
{
    __shared__ uint64_t warp_sums[THREADS / 32];
    __shared__ uint64_t block_total;
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint64_t incl = warp_incl_scan(v);
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint64_t s = lane < THREADS / 32 ? warp_sums[lane] : 0;
        const uint64_t si = warp_incl_scan(s);
        if (lane < THREADS / 32) warp_sums[lane] = si - s;
        if (lane == 31) block_total = si;
    }
    __syncthreads();
    const uint64_t r = incl - v + warp_sums[wid];
    *total = block_total;
    __syncthreads();
    return r;
}

// ---- launch helpers (engine.cu counts launches for the stats) -------------
extern thread_local uint32_t g_launch_count;
// DVDAGPU_TRACE=1: an event behind every launch and the host time of its enqueue, printed as a
// time line at the end of the decode (engine.cu)
extern thread_local bool g_trace_on;
void trace_mark(const char *what, cudaStream_t s);
#define LAUNCH(kernel, grid, block, smem, stream, ...)                         \
    do {                                                                       \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);            \
        g_launch_count++;                                                      \
        if (g_trace_on && !g_capturing) trace_mark(#kernel, (stream));         \
    } while (0)

// timing events: inside a stream capture they are recorded as external events (readable after
// the graph has run), outside as plain ones
// Only with profiling switched on (dvdagpu_set_profiling): a timing event is a small write to host
// memory in the middle of the launch sequence, and while bulk copies run on the link each of
// them holds the stream up for tens of microseconds.
extern thread_local bool g_capturing;
extern thread_local bool g_profiling;
static inline cudaError_t record_timing(cudaEvent_t e, cudaStream_t s)
{
    if (!g_profiling) return cudaSuccess;
    return cudaEventRecordWithFlags(e, s, g_capturing ? cudaEventRecordExternal : cudaEventRecordDefault);
}

static inline uint32_t div_up_u32(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

// exclusive scans (scan.cu): out has n + 1 entries, out[n] = total
int scan_u32_to_u64(const uint32_t *in, uint64_t *out, uint64_t n, void *tmp, size_t tmp_bytes, cudaStream_t s);
int scan_u32_to_u32(const uint32_t *in, uint32_t *out, uint64_t n, void *tmp, size_t tmp_bytes, cudaStream_t s);
// total_copy (may be null): per table, a second place for its total, e.g. mapped host memory
int scan_batch(const uint32_t *const in[], void *const out[], const bool out64[], int njobs, uint64_t n,
               void *tmp, size_t tmp_bytes, cudaStream_t s, uint64_t *const total_copy[] = nullptr);
size_t scan_tmp_bytes(uint64_t n);
