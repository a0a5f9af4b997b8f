/* dvdagpu.h — C ABI of the sm_100a DVD-Audio decode engine (libdvdagpu.so).
 *
 * This is the private seam between the C host library (libdvd-audio.so, public
 * API in dvd-audio.h) and the CUDA kernels.  Plain pointers and sizes only — no
 * C++ or torch types — so the reference's own C code, or any FFI (ctypes, cgo,
 * JNI), can bind it.  Each entry point names the reference interface it replaces
 * (paths relative to the reference tree).
 *
 * What it replaces: the whole per-track decode the reference performs lazily,
 * packet by packet, inside dvda_read():
 *     src/dvd-audio.c:597-657   dvda_open_track_reader  (codec probe, MLP sync search)
 *     src/dvd-audio.c:1016-1082 decode_pcm_audio        (PCM packet loop)
 *     src/dvd-audio.c:1151-1227 decode_mlp_audio        (MLP packet loop, track end)
 *     src/packet.c:60-188       pack header / PES demux
 *     src/pcm.c:98-169          PCM unpack
 *     src/mlp.c:344-1399        MLP access-unit decode
 * Here a track (or a batch of tracks over one sector buffer) is decoded in one
 * call, entirely on the GPU, and the interleaved int PCM is fetched afterwards.
 * There is no CPU fallback: without a usable CUDA device dvdagpu_create() fails.
 */
#ifndef DVDAGPU_H
#define DVDAGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dvdagpu_ctx dvdagpu_ctx;

/* status of one track (dvdagpu_track_result.status) */
enum {
    DVDAGPU_TRACK_OK = 0,
    DVDAGPU_TRACK_NO_AUDIO = 1      /* reference: dvda_open_track_reader() == NULL */
};

/* error_flags bits: the track ended early in front of a damaged access unit */
enum {
    DVDAGPU_ERR_PARITY = 1 << 4,    /* reference prints "parity mismatch" (mlp.c:692-697) */
    DVDAGPU_ERR_CRC    = 1 << 5,    /* reference prints "CRC-8 mismatch" (mlp.c:701-706) */
    DVDAGPU_ERR_SYNTAX = 1 << 6     /* malformed access unit (reference asserts) */
};

/* One track to decode: sector numbers are indices into the sector buffer
 * handed to dvdagpu_decode_*().  Replaces the (first, last, PTS length) triple
 * the reference passes to open_pcm_track_reader / open_mlp_track_reader
 * (dvd-audio.c:637-646). */
typedef struct {
    uint32_t first_sector;      /* first sector of the track */
    uint32_t last_sector;       /* last sector (MLP: decoding runs on to the next major sync) */
    uint32_t pts_length;        /* track length in 90 kHz ticks (PCM frame budget) */
    uint32_t flags;             /* DVDAGPU_PART_* : the "track" is one part of a longer MLP track */
} dvdagpu_track_desc;

/* A long MLP track may be decoded in parts, each part given as its own
 * "track" over consecutive sector ranges (the cut lands on the first major sync
 * behind last_sector, exactly like the disc's own track boundaries).  The flags
 * keep the reference's per-packet rules exact across the cuts. */
enum {
    DVDAGPU_PART_CONTINUES_PREVIOUS = 1,   /* not the real start of the track */
    DVDAGPU_PART_CONTINUED_BY_NEXT = 2,    /* not the real end of the track */
    /* A long PCM track is read in windows of whole sectors too: every packet is independent, what
     * carries over is only the frame budget (dvd-audio.c:1016-1082).  With this flag pts_length holds
     * the frames still to be delivered instead of the track's length in ticks; the window then
     * behaves like the track from there on (whole packets, stop at a packet that is not PCM or
     * changes the parameters). */
    DVDAGPU_PCM_BUDGET_IN_FRAMES = 4
};

/* What the reference exposes through dvda_codec(), dvda_bits_per_sample(),
 * dvda_sample_rate(), dvda_channel_count() plus the decoded length. */
typedef struct {
    int32_t status;             /* DVDAGPU_TRACK_* */
    int32_t error_flags;        /* DVDAGPU_ERR_* */
    int32_t codec;              /* 0 = PCM, 1 = MLP (dvda_codec_t) */
    uint32_t group_0_bps, group_1_bps, group_0_rate, group_1_rate, channel_assignment;
    uint32_t channels, bits_per_sample, sample_rate;
    uint32_t truncated;         /* 1: the sector buffer ended before the track's natural end
                                   (next major sync / PCM frame budget); pass more sectors */
    uint32_t stopped;           /* 0: ran to its natural end; 1: ended early (damage, or the reference's
                                   "packet without a complete access unit" rule): later parts must be
                                   dropped; 2: a continued part needs the previous part's filter
                                   history: decode the parts in one piece instead */
    uint32_t reserved;
    uint64_t frames;            /* PCM frames decoded */
    uint64_t pcm_offset;        /* first sample of the track in the engine's PCM buffer (in int32 units) */
} dvdagpu_track_result;

/* per-stage device times of the last decode, milliseconds (CUDA events).  The times are only
 * taken with profiling switched on (dvdagpu_set_profiling): the events cost a decode a few
 * microseconds each, tens of microseconds while bulk copies run on the PCIe link. */
typedef struct {
    float demux_ms;             /* sector scan + packet tables + elementary-stream gather */
    float index_ms;             /* sync search, access-unit chase, segment table */
    float decode_ms;            /* check data + entropy decode + prediction filters */
    float output_ms;            /* rematrix / shift / interleave, PCM unpack */
    float total_ms;
    uint32_t launches;          /* kernels launched by the last decode */
    uint32_t segments;          /* restart segments decoded */
    uint64_t access_units;
    uint64_t es_bytes;          /* MLP elementary-stream bytes */
    uint64_t samples;           /* total samples produced */
    float kernel_ms[16];        /* device time of the main kernels, see DVDAGPU_K_* */
} dvdagpu_stats;

/* indices into dvdagpu_stats.kernel_ms */
enum {
    DVDAGPU_K_ES_GATHER = 0,    /* payload bytes -> elementary stream */
    DVDAGPU_K_SYNC_SCAN = 1,    /* major sync search (both passes) */
    DVDAGPU_K_AU_CHASE = 2,     /* access-unit chains (both passes) */
    DVDAGPU_K_CHECKDATA = 3,    /* parity / CRC-8 */
    DVDAGPU_K_MLP_DECODE = 4,   /* complete single-pass MLP decoder (fallback of the three-pass path) */
    DVDAGPU_K_CARRY_FIX = 5,    /* segments needing the previous segment's FIR history */
    DVDAGPU_K_REMATRIX = 6,     /* matrices, bypass, shift, interleave */
    DVDAGPU_K_PCM_UNPACK = 7,
    DVDAGPU_K_MLP_SEGCTX = 8,   /* three-pass path, A0: restart header of every segment -> parsing context */
    DVDAGPU_K_MLP_ENTROPY = 9,  /* three-pass path, B: residual entropy decode, one lane per access unit */
    DVDAGPU_K_MLP_FILTER = 10,  /* three-pass path, C: FIR/IIR prediction, one lane per channel (2 substreams) */
    DVDAGPU_K_MLP_FILTER_OUT = 11, /* three-pass path, C (1 substream): prediction + rematrix + interleaved output */
    DVDAGPU_K_MLP_AU_PARSE = 12, /* three-pass path, A1: parameter block of every access unit, as a delta */
    DVDAGPU_K_MLP_RESOLVE = 13,  /* three-pass path, A2: parameter chain of every segment resolved per access unit */
    DVDAGPU_K_MLP_FUSED = 14     /* default path: entropy decode + prediction + rematrix + interleaved output, one lane per channel */
};

/* number of CUDA devices the engine can use (0 = none) */
int dvdagpu_device_count(void);

/* Creates an engine on the given CUDA device.  NULL if there is no usable
 * device (message in dvdagpu_last_error()). */
dvdagpu_ctx *dvdagpu_create(int device);
void dvdagpu_destroy(dvdagpu_ctx *ctx);

/* Run all work of this context on an existing CUDA stream (a cudaStream_t
 * passed as void*; NULL = the context's own stream).  Lets a caller time the
 * engine with its own events or order it after its own copies.  The context's
 * own stream has the highest stream priority (its side work runs on a stream of
 * the lowest, beside the decode chain); a caller's stream keeps whatever
 * priority it was created with. */
int dvdagpu_set_stream(dvdagpu_ctx *ctx, void *cuda_stream);

/* Decode n_tracks tracks whose sectors live in HOST memory (`sectors`,
 * n_sectors * 2048 bytes; pinned memory makes the upload asynchronous).
 * Uploads the sectors, runs every kernel, leaves the PCM in the engine's device
 * buffer and fills results[].  Returns 0, or nonzero on an engine error
 * (dvdagpu_last_error()).  Replaces: dvda_open_track_reader + the decode loop. */
int dvdagpu_decode_host(dvdagpu_ctx *ctx, const uint8_t *sectors, uint64_t n_sectors,
                        uint32_t n_tracks, const dvdagpu_track_desc *tracks,
                        dvdagpu_track_result *results);

/* Same, with the sectors already resident in DEVICE memory. */
int dvdagpu_decode_device(dvdagpu_ctx *ctx, const void *device_sectors, uint64_t n_sectors,
                          uint32_t n_tracks, const dvdagpu_track_desc *tracks,
                          dvdagpu_track_result *results);

/* Decode ONE track from host memory with the copies overlapped: the track is
 * cut into parts of `part_sectors` sectors (0 = default), part i+1 is uploaded
 * and part i-1 downloaded while part i is decoded, and the interleaved samples
 * land in `pcm_host` (pinned memory, room for `pcm_capacity` int32) in order.
 * result->pcm_offset is 0.  Falls back to the one-piece path by itself when the
 * track cannot be cut (PCM, short, filter history crossing a cut).  Returns 0,
 * 3 if pcm_host is too small, another nonzero value on an engine error.
 * Replaces: dvda_open_track_reader + decode loop + dvda_read copies for a whole
 * track (dvd-audio.c:597-795). */
int dvdagpu_decode_track_pipelined(dvdagpu_ctx *ctx, const uint8_t *sectors, uint64_t n_sectors,
                                   const dvdagpu_track_desc *track, uint32_t part_sectors,
                                   int32_t *pcm_host, uint64_t pcm_capacity,
                                   dvdagpu_track_result *result);

/* Copy decoded samples of the last decode to host memory: `count` int32
 * samples starting at sample `offset` of the engine's PCM buffer (use
 * results[t].pcm_offset).  Replaces the interleave copy of dvda_read
 * (dvd-audio.c:781-792).  Synchronous. */
int dvdagpu_fetch(dvdagpu_ctx *ctx, uint64_t offset, uint64_t count, int32_t *dst);

/* device pointer of the engine's PCM buffer (int32 samples) and its length */
const void *dvdagpu_pcm_device(dvdagpu_ctx *ctx, uint64_t *n_samples);

/* pinned host memory helpers for callers that stage sectors / PCM themselves */
void *dvdagpu_host_alloc(size_t bytes);
void dvdagpu_host_free(void *p);
/* bytes of dvdagpu_host_alloc() memory live now, and the most that ever were (reset_peak: start a
 * new measurement).  The host library's track readers keep their pinned memory bounded whatever
 * the track's length; this is how a test sees it. */
void dvdagpu_host_usage(uint64_t *live_bytes, uint64_t *peak_bytes, int reset_peak);

int dvdagpu_get_stats(dvdagpu_ctx *ctx, dvdagpu_stats *out);
/* per-stage and per-kernel CUDA-event timing of the following decodes on / off (default off) */
int dvdagpu_set_profiling(dvdagpu_ctx *ctx, int on);

/* message for the last failure on this thread ("" if none) */
const char *dvdagpu_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
