/* dvda_gen.c — synthetic DVD-Audio disc generator.  See dvda_gen.h.
 *
 * Test / bench infrastructure: this is an *encoder* for the container and the
 * two codecs the decode path understands.  It shares no code with the decoder
 * (libdvd-audio_b200/) nor with the checker (oracle/).
 *
 * Bitstream layouts: SURVEY.md Appendix A (A.1-A.14); each emitter below names
 * the reference parser it is the inverse of.
 */
#define _POSIX_C_SOURCE 200809L
#include "dvda_gen.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define SECTOR 2048
#define MAXCH 8
#define MAXMAT 6

static char g_err[512];
const char *dvda_gen_error(void) { return g_err; }
static int fail(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return -1;
}

/* ------------------------------------------------------------------ RNG */
typedef struct { uint64_t s; } rng_t;
static inline uint64_t rng_u64(rng_t *r)
{
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline uint32_t rng_below(rng_t *r, uint32_t n) { return n ? (uint32_t)((rng_u64(r) >> 32) * (uint64_t)n >> 32) : 0; }
static inline int rng_range(rng_t *r, int lo, int hi) { return lo + (int)rng_below(r, (uint32_t)(hi - lo + 1)); }
static inline int rng_pct(rng_t *r, int pct) { return (int)rng_below(r, 100) < pct; }

/* ----------------------------------------------------------- bit writer */
typedef struct {
    uint8_t *buf;
    size_t cap;       /* bytes */
    size_t nbits;
} bitw_t;

static void bw_init(bitw_t *w, size_t cap)
{
    w->buf = calloc(cap, 1);
    w->cap = cap;
    w->nbits = 0;
}
static void bw_reset(bitw_t *w)
{
    size_t used = (w->nbits + 7) / 8 + 8;
    memset(w->buf, 0, used < w->cap ? used : w->cap);
    w->nbits = 0;
}
static void bw_free(bitw_t *w) { free(w->buf); w->buf = NULL; }

/* MSB-first, n <= 32 (reference bit order: src/bitstream.c:1077-1111) */
static inline void bw_put(bitw_t *w, unsigned n, uint32_t v)
{
    if (!n) return;
    if ((w->nbits + n + 7) / 8 + 8 > w->cap) {
        size_t ncap = w->cap * 2 + 64;
        w->buf = realloc(w->buf, ncap);
        memset(w->buf + w->cap, 0, ncap - w->cap);
        w->cap = ncap;
    }
    if (n < 32) v &= (1u << n) - 1u;
    /* place v so that its top bit lands at bit position nbits */
    size_t byte = w->nbits >> 3;
    unsigned off = (unsigned)(w->nbits & 7);
    uint64_t acc = (uint64_t)v << (64 - n - off);   /* n + off <= 39 */
    for (unsigned i = 0; i < 5; i++) {
        w->buf[byte + i] |= (uint8_t)(acc >> (56 - 8 * i));
    }
    w->nbits += n;
}
/* two's complement in n bits (reference src/bitstream.c:1198-1206) */
static inline void bw_put_s(bitw_t *w, unsigned n, int32_t v) { bw_put(w, n, (uint32_t)v); }
static inline void bw_align(bitw_t *w) { w->nbits = (w->nbits + 7) & ~(size_t)7; }
static inline size_t bw_bytes(const bitw_t *w) { return (w->nbits + 7) >> 3; }

/* big-endian helpers for the byte-oriented tables */
static void put_be16(uint8_t *p, unsigned v) { p[0] = (uint8_t)(v >> 8); p[1] = (uint8_t)v; }
static void put_be32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }

/* ------------------------------------------------------- field decoders */
static unsigned rate_hz(int code)
{
    switch (code) {
    case 0: return 48000; case 1: return 96000; case 2: return 192000;
    case 8: return 44100; case 9: return 88200; case 10: return 176400;
    default: return 0;
    }
}
static unsigned rate_mult(int code) { return (code & 7) == 0 ? 1 : (code & 7) == 1 ? 2 : 4; }
static unsigned channels_of(int a)
{
    static const uint8_t n[21] = {1, 2, 3, 4, 3, 4, 5, 3, 4, 5, 4, 5, 6, 4, 5, 4, 5, 6, 5, 5, 6};
    return (a >= 0 && a <= 20) ? n[a] : 0;
}

/* ---------------------------------------------------------------- muxer */
typedef struct {
    char dir[1024];
    FILE *f;
    int aob_index;             /* 1..9 */
    uint64_t aob_bytes;
    uint64_t max_aob_bytes;
    uint32_t sectors;          /* global sector count written so far */
    rng_t rng;
    /* pending MLP elementary-stream bytes not yet placed in a sector */
    uint8_t *pend;
    size_t pend_len, pend_cap;
    int features;
    uint64_t pts;
} mux_t;

static int mux_open_next(mux_t *m)
{
    char path[1200];
    if (m->f) fclose(m->f);
    m->aob_index++;
    if (m->aob_index > 9) return fail("more than 9 AOB files needed");
    snprintf(path, sizeof path, "%s/ATS_01_%d.AOB", m->dir, m->aob_index);
    m->f = fopen(path, "wb");
    if (!m->f) return fail("cannot create %s", path);
    m->aob_bytes = 0;
    return 0;
}

static int mux_write_sector(mux_t *m, const uint8_t *sec)
{
    if (!m->f || m->aob_bytes + SECTOR > m->max_aob_bytes) {
        if (mux_open_next(m)) return -1;
    }
    if (fwrite(sec, 1, SECTOR, m->f) != SECTOR) return fail("short write");
    m->aob_bytes += SECTOR;
    m->sectors++;
    return 0;
}

/* pack header, inverse of read_pack_header (reference src/packet.c:137-188) */
static size_t put_pack_header(uint8_t *sec, mux_t *m, unsigned stuffing)
{
    bitw_t w = {sec, SECTOR, 0};
    const uint64_t pts = m->pts;
    bw_put(&w, 32, 0x000001BA);
    bw_put(&w, 2, 1);
    bw_put(&w, 3, (uint32_t)((pts >> 30) & 7));
    bw_put(&w, 1, 1);
    bw_put(&w, 15, (uint32_t)((pts >> 15) & 0x7FFF));
    bw_put(&w, 1, 1);
    bw_put(&w, 15, (uint32_t)(pts & 0x7FFF));
    bw_put(&w, 1, 1);
    bw_put(&w, 9, 0);
    bw_put(&w, 1, 1);
    bw_put(&w, 22, 0x1869F & 0x3FFFFF);
    bw_put(&w, 2, 3);
    bw_put(&w, 5, 0x1F);
    bw_put(&w, 3, stuffing);
    for (unsigned i = 0; i < stuffing; i++) sec[14 + i] = 0xFF;
    m->pts += 2880;
    return 14 + stuffing;
}

/* One sector carrying audio payload in one (or two) audio packets.
 * codec_id 0xA0 (PCM, `params` = the 9 parameter bytes) or 0xA1 (MLP).
 * Takes up to len bytes from data (PCM: whole `granule`s) and returns the
 * number consumed, (size_t)-1 on error.  Packet syntax: reference
 * src/packet.c:97-108, src/dvd-audio.c:1238-1248. */
static size_t mux_audio_sector(mux_t *m, int codec_id, const uint8_t *params,
                               const uint8_t *data, size_t len, size_t granule)
{
    uint8_t sec[SECTOR];
    memset(sec, 0, sizeof sec);
    const int rnd = (m->features & DVDA_GEN_RANDOM_PADS) != 0;
    const int is_pcm = codec_id == 0xA0;
    const unsigned stuffing = rnd ? (unsigned)rng_below(&m->rng, 8) : 0;
    size_t pos = put_pack_header(sec, m, stuffing);
    size_t consumed = 0;
    const int npk = ((m->features & DVDA_GEN_TWO_PACKETS) && rng_pct(&m->rng, 40)) ? 2 : 1;
    const size_t min_packet = 6 + 7 + 3 + 9 + 5 + granule + 6;

    for (int p = 0; p < npk; p++) {
        const size_t room_total = SECTOR - pos;
        unsigned pad1 = rnd ? (unsigned)rng_below(&m->rng, 4) : 0;
        unsigned pad2 = is_pcm ? 9u : 0u;
        if (rnd) pad2 += (unsigned)rng_below(&m->rng, 6);
        size_t hdr = 6 + 7 + pad1 + pad2;
        if (room_total < hdr + granule) break;
        size_t room = room_total - hdr;                 /* most body bytes possible */
        int last = (p == npk - 1);
        if (!last) {
            size_t cut = room / 4 + rng_below(&m->rng, (uint32_t)(room / 2));
            if (room - cut < min_packet) last = 1; else room = cut;
        }
        const size_t avail = len - consumed;
        size_t take = avail < room ? avail : room;
        if (is_pcm) take = (take / granule) * granule;
        if (take == 0) break;
        size_t junk = 0;
        if (is_pcm && (m->features & DVDA_GEN_PCM_RAGGED) && granule > 1 && rng_pct(&m->rng, 30)) {
            /* a trailing partial chunk the decoder must drop (reference pcm.c:147) */
            junk = 1 + rng_below(&m->rng, (uint32_t)(granule - 1));
            if (take + junk > room) junk = room - take;
        }
        const size_t body = take + junk;
        if (last) {
            const size_t left = room_total - hdr - body;
            if (left >= 1 && left <= 5) { pad2 += (unsigned)left; hdr += left; }
        }
        const size_t pes_len = hdr - 6 + body;
        uint8_t *q = sec + pos;
        q[0] = 0; q[1] = 0; q[2] = 1; q[3] = 0xBD;
        put_be16(q + 4, (unsigned)pes_len);
        q += 6;
        q[0] = 0x80; q[1] = 0x00;                        /* 16 skipped bits */
        q[2] = (uint8_t)pad1;
        q += 3;
        memset(q, 0xFF, pad1);
        q += pad1;
        q[0] = (uint8_t)codec_id;
        q[1] = 0; q[2] = 0;
        q[3] = (uint8_t)pad2;
        q += 4;
        if (is_pcm) {
            memcpy(q, params, 9);
            memset(q + 9, 0, pad2 - 9);
        } else {
            for (unsigned i = 0; i < pad2; i++) q[i] = (uint8_t)rng_u64(&m->rng);
        }
        q += pad2;
        memcpy(q, data + consumed, take);
        for (size_t i = 0; i < junk; i++) q[take + i] = (uint8_t)rng_u64(&m->rng);
        consumed += take;
        pos += hdr + body;
        if (last || consumed == len) break;
    }
    /* tile the rest of the sector with a padding packet (reference packet.c:68) */
    const size_t left = SECTOR - pos;
    if (left >= 6) {
        uint8_t *q = sec + pos;
        q[0] = 0; q[1] = 0; q[2] = 1; q[3] = 0xBE;
        put_be16(q + 4, (unsigned)(left - 6));
        memset(q + 6, 0xFF, left - 6);
    } else if (left != 0) {
        fail("internal: %zu untiled bytes in sector", left);
        return (size_t)-1;
    }
    if (consumed == 0) {
        fail("internal: sector consumed no payload");
        return (size_t)-1;
    }
    if (mux_write_sector(m, sec)) return (size_t)-1;
    return consumed;
}

/* MLP: append elementary-stream bytes; emit full sectors as they fill */
static int mux_mlp_append(mux_t *m, const uint8_t *es, size_t n)
{
    if (m->pend_len + n > m->pend_cap) {
        m->pend_cap = (m->pend_len + n) * 2 + 4096;
        m->pend = realloc(m->pend, m->pend_cap);
    }
    memcpy(m->pend + m->pend_len, es, n);
    m->pend_len += n;
    /* a sector never holds more than ~2021 stream bytes; keep < 2100 pending */
    size_t off = 0;
    while (m->pend_len - off >= 2100) {
        size_t used = mux_audio_sector(m, 0xA1, NULL, m->pend + off, m->pend_len - off, 1);
        if (used == (size_t)-1) return -1;
        off += used;
    }
    if (off) {
        memmove(m->pend, m->pend + off, m->pend_len - off);
        m->pend_len -= off;
    }
    return 0;
}

/* MLP: flush the remaining stream bytes into (short) final sectors */
static int mux_mlp_flush(mux_t *m)
{
    size_t off = 0;
    while (off < m->pend_len) {
        size_t used = mux_audio_sector(m, 0xA1, NULL, m->pend + off, m->pend_len - off, 1);
        if (used == (size_t)-1) return -1;
        off += used;
    }
    m->pend_len = 0;
    return 0;
}

/* -------------------------------------------------------------- PCM side */

/* The AOB sample layout as *group lists* (behaviour of the
 * reference's permutation, src/pcm.c:103-166): a chunk is two frames, samples
 * numbered frame*channels+channel.  A group lists sample numbers in stream
 * order; 16-bit groups carry big-endian samples, 24-bit groups carry all
 * (high, middle) byte pairs first and then all low bytes. */
typedef struct { int n; int s[12]; } pcm_group_t;
typedef struct { int ngroups; pcm_group_t g[2]; } pcm_layout_t;

static void pcm_layout(int bps24, int ch, pcm_layout_t *L)
{
    memset(L, 0, sizeof *L);
    const int n = 2 * ch;
    int split = 0;   /* 1 = two-group layout */
    if (ch == 6) split = 1;
    if (bps24 && ch >= 3) split = 1;
    if (!split) {
        L->ngroups = 1;
        L->g[0].n = n;
        for (int i = 0; i < n; i++) L->g[0].s[i] = i;
        return;
    }
    /* first group: a run of channels starting at channel 2, both frames */
    int a0 = 2, a1;              /* channels [a0, a1) form group A */
    switch (ch) {
    case 3: a1 = 3; break;
    case 4: a1 = 4; break;
    case 5: a1 = 5; break;
    default: a1 = 4; break;      /* 6 channels */
    }
    L->ngroups = 2;
    for (int f = 0; f < 2; f++)
        for (int c = a0; c < a1; c++) L->g[0].s[L->g[0].n++] = f * ch + c;
    for (int f = 0; f < 2; f++)
        for (int c = 0; c < ch; c++)
            if (c < a0 || c >= a1) L->g[1].s[L->g[1].n++] = f * ch + c;
}

/* pack one chunk (2 frames) of ints into AOB byte order */
static size_t pcm_pack_chunk(const pcm_layout_t *L, int bps24, const int32_t *smp, uint8_t *out)
{
    size_t o = 0;
    for (int g = 0; g < L->ngroups; g++) {
        const pcm_group_t *G = &L->g[g];
        if (!bps24) {
            for (int i = 0; i < G->n; i++) {
                uint32_t v = (uint32_t)smp[G->s[i]];
                out[o++] = (uint8_t)(v >> 8);
                out[o++] = (uint8_t)v;
            }
        } else {
            for (int i = 0; i < G->n; i++) {
                uint32_t v = (uint32_t)smp[G->s[i]];
                out[o++] = (uint8_t)(v >> 16);
                out[o++] = (uint8_t)(v >> 8);
            }
            for (int i = 0; i < G->n; i++) out[o++] = (uint8_t)smp[G->s[i]];
        }
    }
    return o;
}

static int gen_pcm_track(mux_t *m, const dvda_gen_track_t *t, dvda_gen_info_t *info)
{
    const int bps24 = t->bps_code != 0;
    const int ch = (int)channels_of(t->assignment);
    if (!ch) return fail("bad channel assignment %d", t->assignment);
    if (t->bps_code != 0 && t->bps_code != 2) return fail("PCM bps code %d not generated", t->bps_code);
    const size_t chunk = (size_t)(bps24 ? 3 : 2) * (size_t)ch * 2;
    pcm_layout_t L;
    pcm_layout(bps24, ch, &L);
    rng_t rng = {t->seed * 0x9E3779B97F4A7C15ULL + 17};

    /* 9 parameter bytes, inverse of dvda_pcmdecoder_decode_params (pcm.c:79-96) */
    uint8_t params[9];
    params[0] = 0; params[1] = 0x10;          /* first_audio_frame */
    params[2] = 0;
    params[3] = (uint8_t)((t->bps_code << 4) | (ch > 2 ? t->bps_code : 0xF));
    params[4] = (uint8_t)((t->rate_code << 4) | (ch > 2 ? t->rate_code : 0xF));
    params[5] = 0;
    params[6] = (uint8_t)t->assignment;
    params[7] = 0;
    params[8] = 0x80;

    int64_t frames = (t->frames + 1) & ~(int64_t)1;
    if (frames < 2) frames = 2;
    info->first_sector = m->sectors;
    info->channels = (uint32_t)ch;

    const size_t buf_chunks = 4096;
    uint8_t *buf = malloc(buf_chunks * chunk + chunk);
    size_t have = 0;               /* bytes in buf */
    int64_t made = 0;
    int32_t smp[12];
    const uint32_t range = bps24 ? (1u << 24) : (1u << 16);
    int64_t payload = 0;
    int64_t emitted_frames = 0;
    while (made < frames || have) {
        while (made < frames && have + chunk <= buf_chunks * chunk) {
            for (int i = 0; i < 2 * ch; i++) {
                uint32_t r = (uint32_t)(rng_u64(&rng) >> 20) & (range - 1);
                smp[i] = (int32_t)r - (int32_t)(range >> 1);
            }
            have += pcm_pack_chunk(&L, bps24, smp, buf + have);
            made += 2;
        }
        /* emit sectors while at least one sector's worth (or the tail) is buffered */
        size_t off = 0;
        while (have - off >= 2100 || (made >= frames && have - off > 0)) {
            /* from the middle of the track on: other stream parameters (the second group's rate code) */
            if ((t->features & DVDA_GEN_PCM_PARAM_CHANGE) && emitted_frames >= frames / 2) params[4] = (uint8_t)((params[4] & 0xF0) | ((params[4] & 0x0F) ^ 0x01));
            size_t used = mux_audio_sector(m, 0xA0, params, buf + off, have - off, chunk);
            if (used == (size_t)-1) { free(buf); return -1; }
            emitted_frames += (int64_t)(used / chunk) * 2;
            off += used;
            payload += (int64_t)used;
        }
        memmove(buf, buf + off, have - off);
        have -= off;
    }
    free(buf);
    info->frames = frames;
    info->payload_bytes = payload;
    const double pts = (double)frames * 90000.0 / (double)rate_hz(t->rate_code);
    info->pts_length = (uint32_t)llround(pts);
    return 0;
}

/* -------------------------------------------------------------- MLP side */

/* Huffman codes, written out from the prefix structure of the three MLP
 * codebooks (reference src/mlp_codebook{1,2,3}.json):
 *   value <= 6      : (8 - value) zeros then a one
 *   "centre" values : '1' + (2 | 1 | 0) literal bits           (cb 1 | 2 | 3)
 *   high values     : '01' + k zeros + '1', k = value - hi_base
 */
static void put_huffman(bitw_t *w, int cb, int v)
{
    static const int centre_bits[4] = {0, 2, 1, 0};
    static const int hi_base[4] = {0, 11, 9, 8};
    if (v <= 6) {
        bw_put(w, (unsigned)(9 - v), 1);
    } else if (v < hi_base[cb]) {
        bw_put(w, 1, 1);
        bw_put(w, (unsigned)centre_bits[cb], (uint32_t)(v - 7));
    } else {
        const int k = v - hi_base[cb];
        bw_put(w, 2, 1);
        bw_put(w, (unsigned)(k + 1), 1);
    }
}
static const int MAX_MSB[4] = {0, 17, 15, 14};

typedef struct {
    int order, shift, coeff_bits, coeff_shift;
    int coeff[8];            /* as the decoder holds them (already << coeff_shift) */
} gfilt_t;

typedef struct {
    gfilt_t fir, iir;
    int fhist[8], flen;      /* most recent first */
    int ihist[8], ilen;
    int iir_state_bits, iir_state_shift;  /* for transmission of the IIR state */
    int iir_sent[8];         /* the state values a fresh IIR block transmits */
    int offset, codebook, lsbs;
    /* target signal oscillator */
    double s1, c1, sw1, cw1, a1;
    double s2, c2, sw2, cw2, a2;
} gchan_t;

typedef struct {
    int out_ch, frac, bypass;
    int raw[8], present[8];
} gmat_t;

typedef struct {
    int min_ch, max_ch, mmc;
    int flags[8];
    int block_size;
    int matrix_len;
    gmat_t mat[MAXMAT];
    int oshift[MAXCH];
    int q[MAXCH];
    gchan_t ch[MAXCH];
    int noise_shift;
    uint32_t seed;
    int segment;             /* restart headers seen since track start */
} gss_t;

typedef struct {
    const dvda_gen_track_t *t;
    rng_t rng;
    int nch, nss, au_frames;
    gss_t ss[2];
    bitw_t bw[2];
    bitw_t au;
    int32_t *target;         /* [MAXCH][au_frames] scratch */
    int32_t *resid;          /* [MAXCH][au_frames] scratch */
} genc_t;

static void osc_init(gchan_t *c, rng_t *r)
{
    const double two_pi = 6.283185307179586;
    double w1 = 0.002 + 0.03 * (double)rng_below(r, 1000) / 1000.0;
    double w2 = 0.0005 + 0.004 * (double)rng_below(r, 1000) / 1000.0;
    double p1 = two_pi * (double)rng_below(r, 1000) / 1000.0;
    double p2 = two_pi * (double)rng_below(r, 1000) / 1000.0;
    c->s1 = sin(p1); c->c1 = cos(p1); c->sw1 = sin(w1); c->cw1 = cos(w1);
    c->s2 = sin(p2); c->c2 = cos(p2); c->sw2 = sin(w2); c->cw2 = cos(w2);
    c->a1 = (double)(1 << 18) * (0.5 + (double)rng_below(r, 1000) / 1000.0);
    c->a2 = (double)(1 << 19) * (0.5 + (double)rng_below(r, 1000) / 1000.0);
}

static inline int32_t osc_next(gchan_t *c, rng_t *r, int noise_bits)
{
    double s = c->s1 * c->cw1 + c->c1 * c->sw1;
    double co = c->c1 * c->cw1 - c->s1 * c->sw1;
    c->s1 = s; c->c1 = co;
    s = c->s2 * c->cw2 + c->c2 * c->sw2;
    co = c->c2 * c->cw2 - c->s2 * c->sw2;
    c->s2 = s; c->c2 = co;
    int32_t v = (int32_t)(c->a1 * c->s1 + c->a2 * c->s2);
    if (noise_bits > 0) {
        uint32_t n = (uint32_t)(rng_u64(r) >> 32);
        v += (int32_t)(n >> (32 - noise_bits - 1)) - (1 << noise_bits);
    }
    return v;
}

static inline int mask_q(int x, int q) { return q ? (int)((unsigned)(x >> q) << q) : x; }

static int filter_shift(const gfilt_t *fir, const gfilt_t *iir)
{
    /* reference src/mlp.c:1262-1270 */
    if (fir->shift > 0 && iir->shift > 0) return fir->shift;
    if (fir->order > 0) return fir->shift;
    return iir->shift;
}

/* fresh coefficients.  kind 0 = FIR, 1 = IIR */
static void new_filter(gfilt_t *f, int kind, int order, int shift, rng_t *r)
{
    memset(f, 0, sizeof *f);
    f->order = order;
    if (!order) return;
    f->shift = shift;
    static const int binom[5][4] = {{0}, {1}, {2, -1}, {3, -3, 1}, {4, -6, 4, -1}};
    double c[8];
    if (kind == 0 && order <= 4 && rng_pct(r, 60)) {
        for (int j = 0; j < order; j++)
            c[j] = (double)binom[order][j] * (0.9 + 0.1 * (double)rng_below(r, 1000) / 1000.0);
    } else {
        /* random taps with a bounded absolute sum (2.0 for FIR, 0.5 for IIR) */
        double budget = kind == 0 ? 2.0 : 0.5, sum = 0;
        for (int j = 0; j < order; j++) {
            c[j] = ((double)rng_below(r, 2001) - 1000.0) / 1000.0 / (double)(j + 1);
            sum += fabs(c[j]);
        }
        if (sum > 0) for (int j = 0; j < order; j++) c[j] *= budget / sum * (0.3 + 0.7 * (double)rng_below(r, 1000) / 1000.0);
    }
    f->coeff_shift = (int)rng_below(r, 3);
    int maxabs = 1;
    int v[8];
    for (int j = 0; j < order; j++) {
        long x = lround(c[j] * (double)(1 << shift) / (double)(1 << f->coeff_shift));
        v[j] = (int)x;
        if (abs(v[j]) + 1 > maxabs) maxabs = abs(v[j]) + 1;
    }
    int bits = 1;
    while ((1 << (bits - 1)) < maxabs) bits++;       /* signed width */
    if (bits < 2) bits = 2;
    while (bits + f->coeff_shift > 16) {              /* clamp into 16 bits total */
        if (f->coeff_shift > 0) { f->coeff_shift--; for (int j = 0; j < order; j++) v[j] *= 2; bits++; }
        else break;
    }
    if (bits + f->coeff_shift > 16) {
        bits = 16 - f->coeff_shift;
        const int lim = (1 << (bits - 1)) - 1;
        for (int j = 0; j < order; j++) { if (v[j] > lim) v[j] = lim; if (v[j] < -lim) v[j] = -lim; }
    }
    if (bits < 16 - f->coeff_shift && rng_pct(r, 30)) bits++;   /* slack width */
    f->coeff_bits = bits;
    for (int j = 0; j < order; j++) f->coeff[j] = (int)((unsigned)v[j] << f->coeff_shift);
}

/* inverse of decode_FIR_parameters / decode_IIR_parameters (mlp.c:1029-1120) */
static void put_filter(bitw_t *w, const gfilt_t *f, int is_iir, const gchan_t *c)
{
    bw_put(w, 4, (uint32_t)f->order);
    if (!f->order) return;
    bw_put(w, 4, (uint32_t)f->shift);
    bw_put(w, 5, (uint32_t)f->coeff_bits);
    bw_put(w, 3, (uint32_t)f->coeff_shift);
    for (int j = 0; j < f->order; j++)
        bw_put_s(w, (unsigned)f->coeff_bits, f->coeff[j] >> f->coeff_shift);
    if (!is_iir) {
        bw_put(w, 1, 0);
    } else {
        /* always send state */
        bw_put(w, 1, 1);
        bw_put(w, 4, (uint32_t)c->iir_state_bits);
        bw_put(w, 4, (uint32_t)c->iir_state_shift);
        for (int j = 0; j < f->order; j++)
            bw_put_s(w, (unsigned)c->iir_state_bits, c->iir_sent[j] >> c->iir_state_shift);
    }
}

static void new_matrices(gss_t *p, const dvda_gen_track_t *t, rng_t *r, int keep_len)
{
    int len = keep_len >= 0 ? keep_len : rng_range(r, rng_pct(r, 85) ? 1 : 0, t->matrices);
    p->matrix_len = len;
    for (int m = 0; m < len; m++) {
        gmat_t *M = &p->mat[m];
        memset(M, 0, sizeof *M);
        M->out_ch = rng_range(r, 0, p->mmc);
        M->frac = rng_range(r, 8, 14);
        M->bypass = (t->features & DVDA_GEN_BYPASS) ? rng_pct(r, 60) : 0;
        const int one = 1 << M->frac;
        const int nin = p->mmc + 1;
        for (int c = 0; c < nin; c++) {
            if (c == M->out_ch) {
                M->present[c] = rng_pct(r, 90);
                M->raw[c] = rng_pct(r, 85) ? one : -one + (int)rng_below(r, (uint32_t)one);
            } else {
                M->present[c] = rng_pct(r, 60);
                const int lim = one / (2 * nin);
                M->raw[c] = rng_range(r, -lim, lim);
            }
            if (!M->present[c]) M->raw[c] = 0;
        }
        for (int c = nin; c < nin + 2; c++) {
            if (t->features & DVDA_GEN_NOISE) {
                M->present[c] = rng_pct(r, 60);
                const int lim = one >> 4;
                M->raw[c] = M->present[c] ? rng_range(r, -lim, lim) : 0;
            } else {
                M->present[c] = rng_pct(r, 10);   /* explicit zero */
                M->raw[c] = 0;
            }
        }
    }
}

/* inverse of decode_matrix_parameters (mlp.c:995-1027) */
static void put_matrices(bitw_t *w, const gss_t *p)
{
    bw_put(w, 4, (uint32_t)p->matrix_len);
    for (int m = 0; m < p->matrix_len; m++) {
        const gmat_t *M = &p->mat[m];
        bw_put(w, 4, (uint32_t)M->out_ch);
        bw_put(w, 4, (uint32_t)M->frac);
        bw_put(w, 1, (uint32_t)M->bypass);
        for (int c = 0; c < p->mmc + 3; c++) {
            bw_put(w, 1, (uint32_t)M->present[c]);
            if (M->present[c]) bw_put_s(w, (unsigned)(M->frac + 2), M->raw[c]);
        }
    }
}

/* residual coding fit: smallest LSB width n (>= min_n) such that all of
 * [rmin, rmax] is representable; offset chosen inside the feasible window.
 * Arithmetic mirrors decode_residual_data (mlp.c:1151-1176). */
static int64_t sho_k(int cb, int n)   /* offset - signed_huffman_offset */
{
    if (cb) {
        const int ss = n + 2 - cb;
        return 7 * ((int64_t)1 << n) + (ss >= 0 ? ((int64_t)1 << ss) : 0);
    }
    return n >= 1 ? ((int64_t)1 << (n - 1)) : 0;
}
static int64_t span_of(int cb, int n) { return cb ? (int64_t)(MAX_MSB[cb] + 1) << n : (int64_t)1 << n; }

static int fit_coding(int64_t rmin, int64_t rmax, int cb, int min_n, int max_n,
                      int fixed_offset, int cur_offset, rng_t *r, int *n_out, int *off_out)
{
    for (int n = min_n; n <= max_n; n++) {
        const int64_t K = sho_k(cb, n), S = span_of(cb, n);
        int64_t lo = rmax - S + 1 + K;    /* offset >= lo */
        int64_t hi = rmin + K;            /* offset <= hi */
        if (lo < -16384) lo = -16384;
        if (hi > 16383) hi = 16383;
        if (lo > hi) continue;
        if (fixed_offset) {
            if (cur_offset < lo || cur_offset > hi) continue;
            *n_out = n; *off_out = cur_offset;
            return 1;
        }
        /* near the middle of the window, with some jitter */
        int64_t mid = (lo + hi) / 2, w = (hi - lo) / 4;
        int64_t o = mid + (w > 0 ? (int64_t)rng_below(r, (uint32_t)(2 * w + 1)) - w : 0);
        *n_out = n; *off_out = (int)o;
        return 1;
    }
    return 0;
}

/* inverse of decode_restart_header (mlp.c:809-854) */
static void put_restart_header(bitw_t *w, gss_t *p, rng_t *r)
{
    bw_put(w, 13, 0x18F5);
    bw_put(w, 1, 0);
    bw_put(w, 16, (uint32_t)rng_u64(r));
    bw_put(w, 4, (uint32_t)p->min_ch);
    bw_put(w, 4, (uint32_t)p->max_ch);
    bw_put(w, 4, (uint32_t)p->mmc);
    bw_put(w, 4, (uint32_t)p->noise_shift);
    bw_put(w, 23, p->seed);
    bw_put(w, 19, (uint32_t)rng_u64(r));
    bw_put(w, 1, (uint32_t)rng_u64(r));
    bw_put(w, 8, (uint32_t)rng_u64(r));
    bw_put(w, 16, (uint32_t)rng_u64(r));
    for (int c = 0; c <= p->mmc; c++)
        bw_put(w, 6, rng_pct(r, 80) ? (uint32_t)c : rng_below(r, (uint32_t)p->mmc + 1));
    bw_put(w, 8, (uint32_t)rng_u64(r));
}

/* One block of one substream: choose parameters, derive residuals, write.
 * Inverse of decode_block (mlp.c:741-807) and everything it calls. */
static int encode_block(genc_t *e, int s, int n, int restart, int first_of_track,
                        int first_in_au, int au_matrix_len)
{
    const dvda_gen_track_t *t = e->t;
    rng_t *r = &e->rng;
    gss_t *p = &e->ss[s];
    bitw_t *w = &e->bw[s];
    const int F = t->features;
    const int nown = p->max_ch - p->min_ch + 1;

    int want = restart || n != p->block_size || !(F & DVDA_GEN_SPARSE) || rng_pct(r, 25);
    /* what this block will (re)send */
    int send_flags = 0, send_bs = 0, send_mat = 0, send_os = 0, send_q = 0;
    int send_ch[MAXCH] = {0}, send_fir[MAXCH] = {0}, send_iir[MAXCH] = {0}, send_off[MAXCH] = {0};
    int flags_explicit = 0;

    if (restart) {
        p->segment++;
        p->noise_shift = (F & DVDA_GEN_NOISE) ? rng_range(r, 0, 6) : rng_range(r, 0, 2);
        p->seed = (uint32_t)rng_u64(r) & 0x7FFFFF;
        /* presence flags */
        if ((F & DVDA_GEN_FLAGS) && rng_pct(r, 50)) {
            flags_explicit = 1;
            p->flags[0] = rng_pct(r, 80);
            for (int k = 1; k < 7; k++) p->flags[k] = rng_pct(r, 80);
            p->flags[7] = 1;
        } else {
            flags_explicit = rng_pct(r, 30) && (F & DVDA_GEN_FLAGS);
            for (int k = 0; k < 8; k++) p->flags[k] = 1;
        }
        /* reset-to-default semantics (mlp.c:904-989) */
        p->block_size = 8;
        p->matrix_len = 0;
        memset(p->oshift, 0, sizeof p->oshift);
        memset(p->q, 0, sizeof p->q);
        for (int c = p->min_ch; c <= p->max_ch; c++) {
            gchan_t *C = &p->ch[c];
            memset(&C->fir, 0, sizeof C->fir);
            memset(&C->iir, 0, sizeof C->iir);
            C->ilen = 0;
            C->offset = 0; C->codebook = 0; C->lsbs = 24;
        }
    } else if (want && p->flags[0] && (F & DVDA_GEN_FLAGS) && rng_pct(r, 10)) {
        send_flags = 1;
        /* only ever widen here, so features in use stay usable */
        for (int k = 1; k < 7; k++) p->flags[k] = p->flags[k] | rng_pct(r, 50);
        p->flags[0] = rng_pct(r, 85);
        p->flags[7] = 1;
    }

    if (want) {
        if (n != p->block_size || rng_pct(r, 15)) { send_bs = 1; p->block_size = n; }
        const int top_ok = first_in_au || (F & DVDA_GEN_MIDAU_PARAMS);
        if (t->matrices > 0 && p->flags[6] && top_ok && (restart ? rng_pct(r, 85) : rng_pct(r, 12))) {
            send_mat = 1;
            new_matrices(p, t, r, first_in_au ? -1 : au_matrix_len);
        }
        if ((F & DVDA_GEN_OUTSHIFT) && p->flags[5] && top_ok && (restart ? rng_pct(r, 70) : rng_pct(r, 8))) {
            send_os = 1;
            for (int c = 0; c <= p->mmc; c++) p->oshift[c] = rng_pct(r, 50) ? 0 : rng_range(r, 0, 2);
        }
        if ((F & DVDA_GEN_QUANT) && p->flags[4] && top_ok && (restart ? rng_pct(r, 70) : rng_pct(r, 8))) {
            send_q = 1;
            for (int c = 0; c <= p->max_ch; c++) p->q[c] = rng_pct(r, 55) ? 0 : rng_range(r, 1, 3);
        }
        for (int c = p->min_ch; c <= p->max_ch; c++) {
            gchan_t *C = &p->ch[c];
            int new_fir = 0, new_iir = 0;
            if (restart) {
                /* G1: FIR order 0 right after a restart unless the carry case is wanted */
                if ((F & DVDA_GEN_FIR_CARRY) && !first_of_track && C->flen >= 8 && p->flags[3] && rng_pct(r, 60)) new_fir = 1;
                if (p->flags[2] && t->iir_max > 0 && rng_pct(r, 50)) new_iir = 1;
            } else {
                if (p->flags[3] && t->fir_max > 0 && rng_pct(r, 45)) new_fir = 1;
                if (p->flags[2] && t->iir_max > 0 && rng_pct(r, 35)) new_iir = 1;
            }
            int fo = C->fir.order, io = C->iir.order;
            if (F & DVDA_GEN_MAX_ORDERS) {
                /* alternate FIR4+IIR4 and FIR8 by segment; first block after a restart stays FIR 0 */
                if (restart) { new_fir = 0; new_iir = 1; io = 4; }
                else if (C->fir.order == 0 && C->flen >= 8) {
                    new_fir = 1; new_iir = 1;
                    if (p->segment & 1) { fo = 4; io = 4; } else { fo = 8; io = 0; }
                } else { new_fir = 0; new_iir = 0; }
            } else {
                if (new_fir) fo = rng_range(r, rng_pct(r, 85) ? 1 : 0, t->fir_max);
                if (new_iir) io = rng_range(r, rng_pct(r, 85) ? 1 : 0, t->iir_max);
                if (fo + io > 8) {
                    if (new_iir) io = 8 - fo; else fo = 8 - io;
                }
                if (fo > C->flen) fo = C->flen;            /* never read missing FIR history */
            }
            if (new_fir || new_iir) {
                int shift = (C->fir.order && !new_fir) ? C->fir.shift
                          : (C->iir.order && !new_iir) ? C->iir.shift
                          : rng_range(r, 8, 12);
                if (new_fir) new_filter(&C->fir, 0, fo, shift, r);
                if (new_iir) {
                    new_filter(&C->iir, 1, io, shift, r);
                    /* transmitted state replaces the history (mlp.c:1098-1108) */
                    C->iir_state_bits = rng_range(r, 1, 12);
                    C->iir_state_shift = rng_range(r, 0, 6);
                    C->ilen = io;
                    for (int k = 0; k < io; k++) {
                        const int lim = 1 << (C->iir_state_bits - 1);
                        int v = rng_range(r, -lim, lim - 1);
                        C->ihist[k] = (int)((unsigned)v << C->iir_state_shift);
                        C->iir_sent[k] = C->ihist[k];
                    }
                }
                send_fir[c] = new_fir; send_iir[c] = new_iir; send_ch[c] = 1;
            }
        }
    }

    /* targets + residuals with the parameters now in force */
    int64_t rmin[MAXCH], rmax[MAXCH];
    for (int c = p->min_ch; c <= p->max_ch; c++) {
        gchan_t *C = &p->ch[c];
        const int q = p->q[c];
        const int shift = filter_shift(&C->fir, &C->iir);
        int32_t *res = e->resid + (size_t)c * (size_t)e->au_frames;
        if (C->fir.order > C->flen || C->iir.order > C->ilen)
            return fail("internal: filter history too short");
        if (C->fir.shift > 0 && C->iir.shift > 0 && C->fir.shift != C->iir.shift)
            return fail("internal: filter shifts differ");
        rmin[c] = INT64_MAX; rmax[c] = INT64_MIN;
        for (int i = 0; i < n; i++) {
            int32_t tv = osc_next(C, r, t->noise_bits);
            tv = mask_q(tv, q);
            int64_t sum = 0;
            for (int j = 0; j < C->fir.order; j++) sum += (int64_t)C->fir.coeff[j] * C->fhist[j];
            for (int k = 0; k < C->iir.order; k++) sum += (int64_t)C->iir.coeff[k] * C->ihist[k];
            const int64_t ps = sum >> shift;
            if (ps > (1 << 29) || ps < -(1 << 29)) return fail("prediction out of range");
            const int pred = (int)ps;
            const int rv = tv - mask_q(pred, q);      /* multiple of 1 << q */
            const int64_t coded = (int64_t)(rv >> q);
            res[i] = (int32_t)coded;
            if (coded < rmin[c]) rmin[c] = coded;
            if (coded > rmax[c]) rmax[c] = coded;
            memmove(C->fhist + 1, C->fhist, 7 * sizeof(int));
            C->fhist[0] = tv;
            if (C->flen < 8) C->flen++;
            memmove(C->ihist + 1, C->ihist, 7 * sizeof(int));
            C->ihist[0] = tv - pred;
            if (C->ilen < 8) C->ilen++;
        }
    }

    /* residual coding parameters */
    for (int c = p->min_ch; c <= p->max_ch; c++) {
        gchan_t *C = &p->ch[c];
        const int q = p->q[c];
        /* do the current ones still fit? */
        int keep = 0;
        if (!send_ch[c] && !restart && (F & DVDA_GEN_SPARSE) && C->lsbs >= q) {
            const int n0 = C->lsbs - q;
            const int64_t sho = C->offset - sho_k(C->codebook, n0);
            if (rmin[c] >= sho && rmax[c] < sho + span_of(C->codebook, n0)) keep = 1;
        }
        if (keep && !rng_pct(r, 10)) continue;
        if (!want && !keep) want = 1;
        /* pick a codebook among the allowed ones */
        int allowed = t->codebooks & 0xF;
        if (!allowed) allowed = 0xF;
        int cb;
        do { cb = (int)rng_below(r, 4); } while (!((allowed >> cb) & 1));
        int nn, off;
        const int can_off = p->flags[1];
        int min_n = t->min_lsbs > 0 ? t->min_lsbs : 0;
        if (min_n > 24 - q) min_n = 24 - q;
        if (!fit_coding(rmin[c], rmax[c], cb, min_n, 24 - q, !can_off, C->offset, r, &nn, &off)) {
            /* try every codebook before giving up */
            int ok = 0;
            for (cb = 3; cb >= 0 && !ok; cb--)
                ok = fit_coding(rmin[c], rmax[c], cb, 0, 24 - q, !can_off, C->offset, r, &nn, &off);
            if (!ok) return fail("residual range [%lld,%lld] not representable", (long long)rmin[c], (long long)rmax[c]);
            cb++;
        }
        if (rng_pct(r, 15) && nn < 24 - q) {
            /* a wider-than-needed LSB field is still valid if it fits */
            int n2, o2;
            if (fit_coding(rmin[c], rmax[c], cb, nn + 1, nn + 1, !can_off, C->offset, r, &n2, &o2)) { nn = n2; off = o2; }
        }
        send_ch[c] = 1;
        send_off[c] = can_off && (off != C->offset || rng_pct(r, 20));
        if (!send_off[c] && off != C->offset) return fail("internal: offset change without flag");
        C->codebook = cb; C->lsbs = nn + q; C->offset = off;
    }
    if (!want) {
        for (int c = p->min_ch; c <= p->max_ch; c++) if (send_ch[c]) want = 1;
    }

    /* ---- write ---- */
    bw_put(w, 1, (uint32_t)want);
    if (want) {
        bw_put(w, 1, (uint32_t)restart);
        if (restart) {
            put_restart_header(w, p, r);
            bw_put(w, 1, (uint32_t)flags_explicit);
            if (flags_explicit) for (int k = 0; k < 8; k++) bw_put(w, 1, (uint32_t)p->flags[k]);
        } else if (p->flags[0] || send_flags) {
            /* note: flags[0] as it was BEFORE this block decides whether the bit exists;
               send_flags is only ever set when it was 1 */
            bw_put(w, 1, (uint32_t)send_flags);
            if (send_flags) for (int k = 0; k < 8; k++) bw_put(w, 1, (uint32_t)p->flags[k]);
        }
        if (p->flags[7]) {
            bw_put(w, 1, (uint32_t)send_bs);
            if (send_bs) bw_put(w, 9, (uint32_t)p->block_size);
        }
        if (p->flags[6]) {
            bw_put(w, 1, (uint32_t)send_mat);
            if (send_mat) put_matrices(w, p);
        }
        if (p->flags[5]) {
            bw_put(w, 1, (uint32_t)send_os);
            if (send_os) for (int c = 0; c <= p->mmc; c++) bw_put_s(w, 4, p->oshift[c]);
        }
        if (p->flags[4]) {
            bw_put(w, 1, (uint32_t)send_q);
            if (send_q) for (int c = 0; c <= p->max_ch; c++) bw_put(w, 4, (uint32_t)p->q[c]);
        }
        for (int c = p->min_ch; c <= p->max_ch; c++) {
            const gchan_t *C = &p->ch[c];
            bw_put(w, 1, (uint32_t)send_ch[c]);
            if (!send_ch[c]) continue;
            if (p->flags[3]) {
                bw_put(w, 1, (uint32_t)send_fir[c]);
                if (send_fir[c]) put_filter(w, &C->fir, 0, C);
            }
            if (p->flags[2]) {
                bw_put(w, 1, (uint32_t)send_iir[c]);
                /* transmits iir_sent[], the state chosen before this block was filtered */
                if (send_iir[c]) put_filter(w, &C->iir, 1, C);
            }
            if (p->flags[1]) {
                bw_put(w, 1, (uint32_t)send_off[c]);
                if (send_off[c]) bw_put_s(w, 15, C->offset);
            }
            bw_put(w, 2, (uint32_t)C->codebook);
            bw_put(w, 5, (uint32_t)C->lsbs);
        }
    }
    (void)nown;

    /* residual data (mlp.c:1194-1238) */
    for (int i = 0; i < n; i++) {
        for (int m = 0; m < p->matrix_len; m++)
            if (p->mat[m].bypass) bw_put(w, 1, (uint32_t)rng_u64(r));
        for (int c = p->min_ch; c <= p->max_ch; c++) {
            const gchan_t *C = &p->ch[c];
            const int nn = C->lsbs - p->q[c];
            const int64_t sho = C->offset - sho_k(C->codebook, nn);
            const int64_t v = (int64_t)e->resid[(size_t)c * (size_t)e->au_frames + i] - sho;
            if (v < 0 || v >= span_of(C->codebook, nn)) return fail("internal: residual does not fit its coding");
            if (C->codebook) put_huffman(w, C->codebook, (int)(v >> nn));
            bw_put(w, (unsigned)nn, (uint32_t)(v & (((int64_t)1 << nn) - 1)));
        }
    }
    return 0;
}

/* CRC-8, polynomial x^8+x^6+x^5+x+1 (0x63), MSB first — the check byte the
 * reference accumulates in checkdata_callback (mlp.c:1360-1399). */
static uint8_t g_crc8[256];
static void crc8_init(void)
{
    for (int i = 0; i < 256; i++) {
        unsigned c = (unsigned)i;
        for (int b = 0; b < 8; b++) c = (c & 0x80) ? ((c << 1) ^ 0x63) & 0xFF : (c << 1) & 0xFF;
        g_crc8[i] = (uint8_t)c;
    }
}

static void enc_init_substreams(genc_t *e)
{
    const int nch = e->nch;
    memset(e->ss, 0, sizeof e->ss);
    if (e->nss == 1) {
        e->ss[0].min_ch = 0; e->ss[0].max_ch = nch - 1; e->ss[0].mmc = nch - 1;
    } else {
        e->ss[0].min_ch = 0; e->ss[0].max_ch = 1; e->ss[0].mmc = 1;
        e->ss[1].min_ch = 2; e->ss[1].max_ch = nch - 1; e->ss[1].mmc = nch - 1;
    }
    for (int s = 0; s < e->nss; s++) {
        e->ss[s].block_size = 8;
        for (int k = 0; k < 8; k++) e->ss[s].flags[k] = 1;
        for (int c = 0; c < MAXCH; c++) { e->ss[s].ch[c].lsbs = 24; osc_init(&e->ss[s].ch[c], &e->rng); }
    }
}

/* a track reader starts with empty filter histories (mlp.c:265-309) */
static void enc_new_decoder(genc_t *e)
{
    for (int s = 0; s < e->nss; s++) {
        e->ss[s].segment = 0;
        for (int c = 0; c < MAXCH; c++) { e->ss[s].ch[c].flen = 0; e->ss[s].ch[c].ilen = 0; }
    }
}

/* One access unit (inverse of read_mlp_frame mlp.c:384,
 * read_major_sync :614, read_substream_info :656, read_substream :670,
 * decode_substream :714).  Appends the AU to e->au. */
static int encode_au(genc_t *e, int is_sync, int with_restart, int first_of_track)
{
    const dvda_gen_track_t *t = e->t;
    rng_t *r = &e->rng;
    const int F = t->features;
    const int checkdata = (F & DVDA_GEN_CHECKDATA) != 0;
    size_t ss_len[2] = {0, 0};

    for (int s = 0; s < e->nss; s++) {
        bitw_t *w = &e->bw[s];
        bw_reset(w);
        /* split the AU into blocks of >= 8 frames */
        int nb = t->max_blocks > 1 ? rng_range(r, 1, t->max_blocks) : 1;
        while (nb > 1 && e->au_frames / nb < 8) nb--;
        int sizes[16];
        if (nb > 16) nb = 16;
        int left = e->au_frames;
        for (int b = 0; b < nb; b++) {
            const int rest = nb - b - 1;
            int sz = (b == nb - 1) ? left : rng_range(r, 8, left - 8 * rest);
            sizes[b] = sz;
            left -= sz;
        }
        int restart0 = with_restart;
        if (!restart0 && !is_sync && (F & DVDA_GEN_MID_RESTART) && rng_pct(r, 6)) restart0 = 1;
        int au_matrix_len = 0;
        for (int b = 0; b < nb; b++) {
            if (encode_block(e, s, sizes[b], b == 0 ? restart0 : 0, first_of_track, b == 0, au_matrix_len))
                return -1;
            if (b == 0) au_matrix_len = e->ss[s].matrix_len;
            bw_put(w, 1, b == nb - 1);
        }
        bw_align(w);
        if ((F & DVDA_GEN_TERMINATOR) && rng_pct(r, 10)) bw_put(w, 32, 0xD234D234);
        else if (rng_pct(r, 5)) bw_put(w, 16, (uint32_t)rng_u64(r));   /* harmless slack */
        size_t nbytes = bw_bytes(w);
        if (nbytes & 1) { bw_put(w, 8, 0); nbytes++; }
        if (checkdata) {
            /* parity ^ 0xA9 and CRC over all bytes but the last two (mlp.c:677-706) */
            uint8_t parity = 0, crc = 0x3C, fin = 0;
            for (size_t i = 0; i < nbytes; i++) {
                parity ^= w->buf[i];
                fin = crc ^ w->buf[i];
                crc = g_crc8[fin];
            }
            bw_put(w, 8, (uint32_t)(parity ^ 0xA9));
            bw_put(w, 8, fin);
            nbytes += 2;
        }
        ss_len[s] = nbytes;
    }

    int extraword[2] = {0, 0};
    size_t dir_len = 0;
    for (int s = 0; s < e->nss; s++) {
        extraword[s] = (F & DVDA_GEN_EXTRAWORD) ? rng_pct(r, 30) : 0;
        dir_len += 2 + (extraword[s] ? 2 : 0);
    }
    const size_t total = 4 + (is_sync ? 28 : 0) + dir_len + ss_len[0] + ss_len[1];
    if (total / 2 > 4095) return fail("access unit too large (%zu bytes)", total);

    bitw_t *a = &e->au;
    bw_put(a, 4, (uint32_t)rng_u64(r));            /* check nibble: not verified (mlp.c:392) */
    bw_put(a, 12, (uint32_t)(total / 2));
    bw_put(a, 16, (uint32_t)rng_u64(r));           /* input timing */
    if (is_sync) {
        const int multi = e->nch > 2;
        bw_put(a, 24, 0xF8726F);
        bw_put(a, 8, 0xBB);
        bw_put(a, 4, (uint32_t)t->bps_code);
        bw_put(a, 4, multi ? (uint32_t)t->bps_code : 0xF);
        bw_put(a, 4, (uint32_t)t->rate_code);
        bw_put(a, 4, multi ? (uint32_t)t->rate_code : 0xF);
        bw_put(a, 11, 0);
        bw_put(a, 5, (uint32_t)t->assignment);
        bw_put(a, 24, 0xB75200); bw_put(a, 24, 0x000000);   /* 48 skipped bits */
        bw_put(a, 1, 1);
        bw_put(a, 15, 0x1234);
        bw_put(a, 4, (uint32_t)e->nss);
        bw_put(a, 32, (uint32_t)rng_u64(r)); bw_put(a, 32, (uint32_t)rng_u64(r));
        bw_put(a, 28, (uint32_t)rng_u64(r));                /* 92 skipped bits */
    }
    size_t end = 0;
    for (int s = 0; s < e->nss; s++) {
        end += ss_len[s];
        int chk_bit = checkdata;
        if (s == 1 && (F & DVDA_GEN_SS1_CHK_QUIRK)) chk_bit = !checkdata;  /* ignored: mlp.c:543-545 */
        bw_put(a, 1, (uint32_t)extraword[s]);
        bw_put(a, 1, (uint32_t)!with_restart);
        bw_put(a, 1, (uint32_t)chk_bit);
        bw_put(a, 1, 0);
        bw_put(a, 12, (uint32_t)(end / 2));
        if (extraword[s]) bw_put(a, 16, (uint32_t)rng_u64(r));
    }
    for (int s = 0; s < e->nss; s++)
        for (size_t i = 0; i < ss_len[s]; i++) bw_put(a, 8, e->bw[s].buf[i]);
    return 0;
}

/* true if bytes p+4..p+7 look like a major sync (dvd-audio.c:1250-1286) */
static int looks_like_sync(const uint8_t *p) { return p[4] == 0xF8 && p[5] == 0x72 && p[6] == 0x6F && p[7] == 0xBB; }

static int gen_mlp_track(mux_t *m, genc_t *e, const dvda_gen_track_t *t, dvda_gen_info_t *info, int joined)
{
    const int nch = (int)channels_of(t->assignment);
    if (!nch) return fail("bad channel assignment %d", t->assignment);
    if (t->substreams == 2 && nch < 3) return fail("two substreams need at least 3 channels");
    if (!joined) {
        /* fresh elementary stream */
        memset(e, 0, sizeof *e);
        e->t = t;
        e->rng.s = t->seed * 0xD1342543DE82EF95ULL + 99;
        e->nch = nch;
        e->nss = t->substreams == 2 ? 2 : 1;
        e->au_frames = t->au_frames > 0 ? t->au_frames : 40 * (int)rate_mult(t->rate_code);
        if (e->au_frames < 8 || e->au_frames > 16 * 511) return fail("bad au_frames");
        bw_init(&e->bw[0], 1 << 16);
        bw_init(&e->bw[1], 1 << 16);
        bw_init(&e->au, 1 << 20);
        e->target = NULL;
        e->resid = malloc(sizeof(int32_t) * MAXCH * (size_t)e->au_frames);
        enc_init_substreams(e);
    } else {
        if (e->nch != nch || e->nss != (t->substreams == 2 ? 2 : 1)) return fail("joined track changes the stream layout");
        e->t = t;
    }
    enc_new_decoder(e);

    const int interval = t->restart_interval > 0 ? t->restart_interval : 16;
    int64_t n_au = (t->frames + e->au_frames - 1) / e->au_frames;
    if (n_au < 1) n_au = 1;
    n_au = (n_au + interval - 1) / interval * interval;

    /* the sector that will hold this track's first byte */
    info->first_sector = m->sectors;
    info->channels = (uint32_t)nch;
    int64_t es_bytes = 0;

    for (int64_t a = 0; a < n_au; a++) {
        const int is_sync = (a % interval) == 0;
        int with_restart = is_sync;
        if (is_sync && a != 0 && (t->features & DVDA_GEN_SYNC_NO_RST) && rng_pct(&e->rng, 30)) with_restart = 0;
        const size_t before = bw_bytes(&e->au);
        if (encode_au(e, is_sync, with_restart, a == 0)) return -1;
        if (is_sync && a != 0 && (t->features & DVDA_GEN_SYNC_PARAM_DUP) && rng_pct(&e->rng, 35)) {
            /* the same access unit twice; the first copy's major sync states another channel
               assignment, so the decoder drops it without looking inside */
            const size_t n = bw_bytes(&e->au) - before;
            for (size_t i = 0; i < n; i++) bw_put(&e->au, 8, e->au.buf[before + i]);
            uint8_t *dup = e->au.buf + before;
            dup[11] = (uint8_t)((dup[11] & 0xE0) | (((dup[11] & 0x1F) + 1) % 21));
        }
        /* G6: a non-sync AU must not look like a sync to the byte scanner.  The
           scanner tests every byte position, so check the window around the AU. */
        (void)before;
        if (bw_bytes(&e->au) >= (1 << 19) || a == n_au - 1) {
            const size_t n = bw_bytes(&e->au);
            if (mux_mlp_append(m, e->au.buf, n)) return -1;
            es_bytes += (int64_t)n;
            bw_reset(&e->au);
        }
    }
    info->frames = n_au * e->au_frames;
    info->payload_bytes = es_bytes;
    info->pts_length = (uint32_t)llround((double)info->frames * 90000.0 / (double)rate_hz(t->rate_code));
    (void)looks_like_sync;
    return 0;
}

static void enc_free(genc_t *e)
{
    bw_free(&e->bw[0]);
    bw_free(&e->bw[1]);
    bw_free(&e->au);
    free(e->resid);
    memset(e, 0, sizeof *e);
}

/* ------------------------------------------------------------- IFO files */

/* AUDIO_TS.IFO (parsed by get_titleset_count dvd-audio.c:824) */
static int write_amg(const char *dir)
{
    char path[1200];
    uint8_t buf[SECTOR];
    memset(buf, 0, sizeof buf);
    memcpy(buf, "DVDAUDIO-AMG", 12);
    buf[63] = 1;                                  /* one title set */
    snprintf(path, sizeof path, "%s/AUDIO_TS.IFO", dir);
    FILE *f = fopen(path, "wb");
    if (!f) return fail("cannot create %s", path);
    fwrite(buf, 1, sizeof buf, f);
    fclose(f);
    return 0;
}

/* ATS_01_0.IFO (parsed by parse_ats_XX_0_ifo dvd-audio.c:860-950) */
static int write_ats(const char *dir, int n_titles, const int32_t *tpt,
                     const dvda_gen_info_t *info, const uint32_t *own_last)
{
    size_t size = SECTOR + 8 + 8 * (size_t)n_titles;
    size_t *tab = malloc(sizeof(size_t) * (size_t)n_titles);
    for (int t = 0; t < n_titles; t++) {
        tab[t] = size - SECTOR;
        size += 16 + 20 * (size_t)tpt[t] + 12 * (size_t)tpt[t];
    }
    size = (size + SECTOR - 1) / SECTOR * SECTOR;
    uint8_t *buf = calloc(size, 1);
    memcpy(buf, "DVDAUDIO-ATS", 12);
    uint8_t *p = buf + SECTOR;
    put_be16(p, (unsigned)n_titles);
    p += 8;
    int k = 0;
    for (int t = 0; t < n_titles; t++) {
        p[0] = (uint8_t)(0x80 | (t + 1));
        put_be32(p + 4, (uint32_t)tab[t]);
        p += 8;
    }
    for (int t = 0; t < n_titles; t++) {
        uint8_t *q = buf + SECTOR + tab[t];
        uint32_t title_pts = 0;
        for (int i = 0; i < tpt[t]; i++) title_pts += info[k + i].pts_length;
        q[2] = (uint8_t)tpt[t];                   /* tracks */
        q[3] = (uint8_t)tpt[t];                   /* indexes */
        put_be32(q + 4, title_pts);
        put_be16(q + 12, (unsigned)(16 + 20 * tpt[t]));  /* sector pointer table offset */
        uint8_t *tr = q + 16;
        uint8_t *ix = q + 16 + 20 * tpt[t];
        uint32_t pts_index = 0;
        for (int i = 0; i < tpt[t]; i++) {
            tr[4] = (uint8_t)(i + 1);             /* index number */
            put_be32(tr + 6, pts_index);
            put_be32(tr + 10, info[k + i].pts_length);
            pts_index += info[k + i].pts_length;
            tr += 20;
            put_be32(ix, 0x01000000);
            put_be32(ix + 4, info[k + i].first_sector);
            put_be32(ix + 8, own_last[k + i]);
            ix += 12;
        }
        k += tpt[t];
    }
    char path[1200];
    snprintf(path, sizeof path, "%s/ATS_01_0.IFO", dir);
    FILE *f = fopen(path, "wb");
    if (!f) { free(buf); free(tab); return fail("cannot create %s", path); }
    fwrite(buf, 1, size, f);
    fclose(f);
    free(buf);
    free(tab);
    return 0;
}

/* ------------------------------------------------------------ entry point */

int dvda_gen_disc(const char *dir, int n_titles, const int32_t *tracks_per_title,
                  const dvda_gen_track_t *tracks, dvda_gen_info_t *info,
                  uint64_t max_aob_bytes)
{
    g_err[0] = 0;
    crc8_init();
    if (n_titles < 1) return fail("need at least one title");
    int total = 0;
    for (int t = 0; t < n_titles; t++) {
        if (tracks_per_title[t] < 1 || tracks_per_title[t] > 99) return fail("1..99 tracks per title");
        total += tracks_per_title[t];
    }
    mux_t m;
    memset(&m, 0, sizeof m);
    snprintf(m.dir, sizeof m.dir, "%s", dir);
    if (!max_aob_bytes) max_aob_bytes = 1ULL << 30;
    if (max_aob_bytes >= (4ULL << 30)) max_aob_bytes = (4ULL << 30) - SECTOR;
    m.max_aob_bytes = max_aob_bytes / SECTOR * SECTOR;
    m.rng.s = tracks[0].seed ^ 0xA5A5A5A5DEADBEEFULL;

    genc_t enc;
    memset(&enc, 0, sizeof enc);
    int enc_live = 0;
    uint32_t *own_last = calloc((size_t)total, sizeof(uint32_t));
    int rc = 0;

    for (int k = 0; k < total && !rc; k++) {
        const dvda_gen_track_t *t = &tracks[k];
        m.features = t->features;
        memset(&info[k], 0, sizeof info[k]);
        const int joined = t->codec == 1 && t->join_previous && enc_live && k > 0;
        if (!joined && enc_live) {
            /* previous elementary stream ends here */
            if (mux_mlp_flush(&m)) { rc = -1; break; }
            enc_free(&enc);
            enc_live = 0;
        }
        if (k > 0 && !joined) own_last[k - 1] = m.sectors - 1;
        if (t->codec == 0) {
            rc = gen_pcm_track(&m, t, &info[k]);
        } else {
            if (joined) {
                /* previous track's own last sector: the one holding its last byte.
                   Pending bytes go into sector m.sectors together with our first. */
                own_last[k - 1] = m.pend_len ? m.sectors : m.sectors - 1;
            }
            rc = gen_mlp_track(&m, &enc, t, &info[k], joined);
            enc_live = (rc == 0);
        }
    }
    if (!rc && enc_live) { rc = mux_mlp_flush(&m); }
    if (enc_live) enc_free(&enc);
    if (!rc) own_last[total - 1] = m.sectors - 1;
    if (m.f) fclose(m.f);
    free(m.pend);

    if (!rc) {
        /* tracks whose tables point somewhere else than where their audio starts */
        for (int k = 0; k < total; k++) {
            if (!tracks[k].start_shift) continue;
            int64_t s = (int64_t)info[k].first_sector + tracks[k].start_shift;
            const int64_t lo = k > 0 ? (int64_t)info[k - 1].first_sector + 1 : 0;
            if (s < lo) s = lo;
            if (s > (int64_t)m.sectors - 1) s = (int64_t)m.sectors - 1;
            info[k].first_sector = (uint32_t)s;
        }
        /* last sectors as the reference derives them (dvd-audio.c:459-499) */
        int k = 0;
        for (int t = 0; t < n_titles; t++) {
            for (int i = 0; i < tracks_per_title[t]; i++, k++) {
                const int last_track = (i == tracks_per_title[t] - 1);
                if (!last_track) info[k].last_sector = info[k + 1].first_sector - 1;
                else if (t == n_titles - 1) info[k].last_sector = own_last[k];
                else {
                    uint32_t a = info[k + 1].first_sector - 1;
                    info[k].last_sector = a > own_last[k] ? a : own_last[k];
                }
            }
        }
        rc = write_amg(dir);
        if (!rc) rc = write_ats(dir, n_titles, tracks_per_title, info, own_last);
    }
    free(own_last);
    return rc;
}

/* 64-bit FNV-1a over a byte range, continuing from h (start with 0xCBF29CE484222325): the hash
 * oracle/api_dump.c prints per track.  Lets bench.py and the tests hash hundreds of megabytes of
 * decoded samples in about a second. */
uint64_t dvda_gen_fnv1a(const void *data, uint64_t nbytes, uint64_t h)
{
    const unsigned char *p = (const unsigned char *)data;
    for (uint64_t i = 0; i < nbytes; i++) {
        h ^= p[i];
        h *= 0x100000001B3ULL;
    }
    return h;
}
