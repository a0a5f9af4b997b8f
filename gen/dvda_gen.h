/* dvda_gen.h — synthetic DVD-Audio disc generator (test / bench infrastructure).
 *
 * Emits AUDIO_TS.IFO, ATS_01_0.IFO and ATS_01_[1-9].AOB files holding PCM
 * packets and *valid MLP access units* produced in encoder fashion: bounded
 * per-channel target signals are chosen first, random prediction filters,
 * matrices, codebooks and block splits second, and the residuals are derived
 * from them, so every stream decodes without integer overflow.  Layouts follow
 * SURVEY.md Appendix A; the constraints that keep the reference decoder out of
 * undefined behaviour follow Appendix B (G1-G7).  Every fixture the tests use
 * is additionally verified by a round trip through the reference decoder.
 */
#ifndef DVDA_GEN_H
#define DVDA_GEN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* feature bits for dvda_gen_track_t.features */
enum {
    DVDA_GEN_CHECKDATA    = 1 << 0,  /* parity + CRC-8 at the end of each substream */
    DVDA_GEN_BYPASS       = 1 << 1,  /* matrices with LSB bypass bits */
    DVDA_GEN_NOISE        = 1 << 2,  /* non-zero noise coefficients / noise_shift */
    DVDA_GEN_QUANT        = 1 << 3,  /* non-zero quant_step_size on some channels */
    DVDA_GEN_OUTSHIFT     = 1 << 4,  /* non-zero output_shift */
    DVDA_GEN_EXTRAWORD    = 1 << 5,  /* substream directory extra words */
    DVDA_GEN_TERMINATOR   = 1 << 6,  /* 0xD234D234 after the last block of some AUs */
    DVDA_GEN_FLAGS        = 1 << 7,  /* explicit / partially cleared presence flags */
    DVDA_GEN_FIR_CARRY    = 1 << 8,  /* FIR order > 0 in the block right after a restart
                                        (segment then depends on the previous one) */
    DVDA_GEN_SPARSE       = 1 << 9,  /* omit parameter blocks when the old ones still fit */
    DVDA_GEN_MID_RESTART  = 1 << 10, /* restart headers in AUs without a major sync */
    DVDA_GEN_SYNC_NO_RST  = 1 << 11, /* some major-sync AUs carry no restart header */
    DVDA_GEN_SS1_CHK_QUIRK= 1 << 12, /* substream 1 advertises the opposite checkdata bit */
    DVDA_GEN_MIDAU_PARAMS = 1 << 13, /* later blocks of an AU change matrices / shifts */
    DVDA_GEN_RANDOM_PADS  = 1 << 14, /* random pad_1 / pad_2 / stuffing per packet */
    DVDA_GEN_PCM_RAGGED   = 1 << 15, /* PCM packets end in a partial chunk */
    DVDA_GEN_TWO_PACKETS  = 1 << 16, /* two audio packets in some sectors */
    DVDA_GEN_MAX_ORDERS   = 1 << 17, /* alternate FIR4+IIR4 / FIR8 (entropy + filter stress) */
    DVDA_GEN_SYNC_PARAM_DUP = 1 << 18, /* in front of some major-sync access units a copy of the unit whose major
                                        sync states other stream parameters: the decoder drops it (mlp.c:449-455) */
    DVDA_GEN_PCM_PARAM_CHANGE = 1 << 19 /* PCM: from the middle of the track on the packets state other stream
                                        parameters: the track ends there (dvd-audio.c:1049-1055) */
};

typedef struct {
    int32_t codec;            /* 0 = PCM, 1 = MLP */
    int32_t bps_code;         /* 0 = 16 bit, 2 = 24 bit */
    int32_t rate_code;        /* 0,1,2 = 48/96/192 kHz; 8,9,10 = 44.1/88.2/176.4 kHz */
    int32_t assignment;       /* channel assignment 0..20 */
    int64_t frames;           /* PCM frames wanted (rounded up to whole AUs / chunks) */
    uint64_t seed;
    int32_t join_previous;    /* MLP: continue the previous track's elementary stream */
    int32_t features;         /* DVDA_GEN_* bits */
    int32_t substreams;       /* MLP: 1 or 2 */
    int32_t au_frames;        /* MLP: frames per access unit, 0 = 40 * rate multiple */
    int32_t restart_interval; /* MLP: access units per major sync (>= 1) */
    int32_t max_blocks;       /* MLP: max blocks per AU per substream (>= 1) */
    int32_t fir_max;          /* MLP: max FIR order (0..8) */
    int32_t iir_max;          /* MLP: max IIR order (0..8) */
    int32_t codebooks;        /* MLP: bitmask of usable codebooks (bit0 = none .. bit3 = cb3) */
    int32_t matrices;         /* MLP: max matrices per substream (0..6) */
    int32_t noise_bits;       /* MLP: log2 amplitude of the unpredictable signal part */
    int32_t min_lsbs;         /* MLP: lower bound for LSB_bits (stress), 0 = fit */
    int32_t start_shift;      /* sectors added to the first sector the title's tables state for this track
                                 (may be negative): a track that does not start where its audio does
                                 (reference TODO:64-80; the reference probes from the stated sector on) */
} dvda_gen_track_t;

/* what was written for each track (outputs) */
typedef struct {
    uint32_t first_sector;
    uint32_t last_sector;     /* as the title's tables imply it (next first - 1) */
    uint32_t pts_length;
    uint32_t channels;
    int64_t frames;           /* frames the generator encoded for this track */
    int64_t payload_bytes;    /* PCM sample bytes / MLP elementary-stream bytes */
} dvda_gen_info_t;

/* Writes a disc into directory `dir` (must exist).  `tracks_per_title[t]` tracks
 * for each of `n_titles` titles, taken consecutively from `tracks`.
 * max_aob_bytes: split the AOB stream into files of at most this size
 * (0 = 1 GiB; always < 4 GiB, at most 9 files).
 * Returns 0 on success, a negative code on error (message in dvda_gen_error()). */
int dvda_gen_disc(const char *dir,
                  int n_titles,
                  const int32_t *tracks_per_title,
                  const dvda_gen_track_t *tracks,
                  dvda_gen_info_t *info_out,
                  uint64_t max_aob_bytes);

const char *dvda_gen_error(void);

/* 64-bit FNV-1a, continuing from h (the per-track hash oracle/api_dump.c prints) */
uint64_t dvda_gen_fnv1a(const void *data, uint64_t nbytes, uint64_t h);

#ifdef __cplusplus
}
#endif
#endif
