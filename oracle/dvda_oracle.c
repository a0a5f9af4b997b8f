/* dvda_oracle.c — plain-C restatement of the reference's decode path.
 *
 * TEST INFRASTRUCTURE ONLY (see dvda_oracle.h).  Parity status: PINNED against
 * the unmodified reference build in oracle/_ref (tests/test_oracle_vs_reference.py)
 * and the golden hashes in tests/golden/.
 *
 * It restates WHAT the reference computes on the path
 *   AOB sectors -> pack/PES demux -> PCM unpack | MLP decode -> interleaved int
 * as one sequential, single-threaded routine over an in-memory sector array.
 * Every function cites the reference code it follows (paths relative to
 * /root/reference).  It is written independently of both the reference (no
 * table-driven bit reader, no growable arrays, no setjmp) and of the CUDA
 * engine (no segment cutting, no parallel structure).
 *
 * Deliberate difference: where the stock reference build assert()-aborts or runs
 * into undefined behaviour on a damaged stream,
 * this restatement ends the track in front of the offending access unit and
 * records an error flag; the CUDA engine does the same.
 */
#include "dvda_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define SECTOR 2048u
#define MAX_CH 8
#define MAX_MAT 6

/* ------------------------------------------------------------------ bits */

/* MSB-first bit reader (reference src/bitstream.c:1077-1111 semantics) */
typedef struct {
    const uint8_t *p;
    size_t nbits;      /* total bits available */
    size_t pos;        /* next bit */
    int err;           /* set when a read ran past the end (the reference longjmps) */
} br_t;

static void br_open(br_t *b, const uint8_t *p, size_t nbytes)
{
    b->p = p; b->nbits = nbytes * 8; b->pos = 0; b->err = 0;
}

static uint32_t br_u(br_t *b, unsigned n)
{
    if (n == 0) return 0;
    if (b->pos + n > b->nbits) { b->err = 1; b->pos = b->nbits; return 0; }
    uint64_t acc = 0;
    size_t byte = b->pos >> 3;
    const unsigned off = (unsigned)(b->pos & 7);
    const unsigned need = (off + n + 7) >> 3;          /* <= 5 */
    for (unsigned i = 0; i < need; i++) acc = (acc << 8) | b->p[byte + i];
    acc >>= (need * 8 - off - n);
    b->pos += n;
    return (uint32_t)(acc & (n == 32 ? 0xFFFFFFFFull : ((1ull << n) - 1)));
}

/* sign bit then magnitude bits: two's complement (src/bitstream.c:1198-1206) */
static int32_t br_s(br_t *b, unsigned n)
{
    if (n == 0) { b->err = 1; return 0; }              /* reference underflows: G2 */
    const uint32_t v = br_u(b, n);
    if (n < 32 && (v >> (n - 1))) return (int32_t)v - (int32_t)(1u << n);
    return (int32_t)v;
}

static void br_skip(br_t *b, size_t n)
{
    if (b->pos + n > b->nbits) { b->err = 1; b->pos = b->nbits; return; }
    b->pos += n;
}

static size_t br_left_bytes(const br_t *b) { return (b->nbits - b->pos) >> 3; }

uint32_t dvda_oracle_read_bits(const uint8_t *buf, size_t len, size_t bitpos, unsigned n)
{
    br_t b; br_open(&b, buf, len); b.pos = bitpos; return br_u(&b, n);
}
int32_t dvda_oracle_read_signed(const uint8_t *buf, size_t len, size_t bitpos, unsigned n)
{
    br_t b; br_open(&b, buf, len); b.pos = bitpos; return br_s(&b, n);
}

/* The three MLP codebooks (src/mlp_codebook{1,2,3}.json) walked bit by bit, the
 * way the reference's jump tables do (src/bitstream.c:1806-1833), using the
 * prefix structure of the codes:
 *   0^z 1      (2 <= z <= 8)         -> 8 - z
 *   0^9                              -> invalid
 *   1 + {2,1,0} literal bits         -> 7 + literal            (cb 1,2,3)
 *   01 0^k 1   (0 <= k <= 6)         -> {11,9,8} + k           (cb 1,2,3)
 *   01 0^7                           -> invalid
 */
static int huffman(br_t *b, int cb)
{
    static const unsigned lit_bits[4] = {0, 2, 1, 0};
    static const int hi_base[4] = {0, 11, 9, 8};
    if (br_u(b, 1)) return 7 + (int)br_u(b, lit_bits[cb]);
    if (br_u(b, 1)) {                                   /* "01" */
        for (int k = 0; k < 7; k++) if (br_u(b, 1)) return hi_base[cb] + k;
        return -1;
    }
    for (int z = 2; z <= 8; z++) if (br_u(b, 1)) return 8 - z;   /* "00..." */
    return -1;
}
int dvda_oracle_huffman(const uint8_t *buf, size_t len, size_t bitpos, int cb, unsigned *code_len)
{
    br_t b; br_open(&b, buf, len); b.pos = bitpos;
    const int v = huffman(&b, cb);
    if (code_len) *code_len = (unsigned)(b.pos - bitpos);
    return v;
}

/* CRC-8, polynomial 0x63 MSB first: same values as the table at mlp.c:1363-1395 */
static uint8_t crc8_tab[256];
static int crc8_ready;
static void crc8_build(void)
{
    for (unsigned i = 0; i < 256; i++) {
        unsigned c = i;
        for (int k = 0; k < 8; k++) c = (c & 0x80) ? ((c << 1) ^ 0x63) & 0xFF : (c << 1) & 0xFF;
        crc8_tab[i] = (uint8_t)c;
    }
    crc8_ready = 1;
}
uint8_t dvda_oracle_crc8_table(unsigned i)
{
    if (!crc8_ready) crc8_build();
    return crc8_tab[i & 0xFF];
}

/* ------------------------------------------------------ field unpacking */

/* src/dvd-audio.c:1423-1496 */
static unsigned unpack_bps(unsigned f) { return f == 0 ? 16 : f == 1 ? 20 : f == 2 ? 24 : 0; }
static unsigned unpack_rate(unsigned f)
{
    switch (f) {
    case 0: return 48000; case 1: return 96000; case 2: return 192000;
    case 8: return 44100; case 9: return 88200; case 10: return 176400;
    default: return 0;
    }
}
static unsigned unpack_channels(unsigned a)
{
    static const uint8_t n[21] = {1, 2, 3, 4, 3, 4, 5, 3, 4, 5, 4, 5, 6, 4, 5, 4, 5, 6, 5, 5, 6};
    return a <= 20 ? n[a] : 0;
}

/* --------------------------------------------------------- sector demux */

/* Sequential packet source over the sector array: src/aob.c:157-213 (sector
 * order), src/packet.c:60-188 (pack header, PES chain, 0xBD filter). */
typedef struct {
    const uint8_t *base;
    uint64_t n_sectors;
    uint64_t next;          /* next sector to fetch */
    const uint8_t *sec;     /* current sector or NULL */
    unsigned off;           /* read offset inside it */
    int dead;               /* a NULL was returned once: nothing more comes */
} demux_t;

typedef struct {
    const uint8_t *data;    /* PES payload */
    unsigned len;
    uint32_t sector;        /* global sector number */
} packet_t;

/* src/packet.c:137-188; returns header size incl. stuffing, 0 if invalid */
static unsigned pack_header_size(const uint8_t *s)
{
    if (!(s[0] == 0 && s[1] == 0 && s[2] == 1 && s[3] == 0xBA)) return 0;
    /* marker bits: 01 at the top of byte 4, then 1s after each field */
    if ((s[4] >> 6) != 1) return 0;
    if (!((s[4] >> 2) & 1)) return 0;
    if (!((s[6] >> 2) & 1)) return 0;
    if (!((s[8] >> 2) & 1)) return 0;
    if (!(s[9] & 1)) return 0;
    if ((s[12] & 3) != 3) return 0;
    return 14u + (s[13] & 7u);
}

static int demux_next_audio(demux_t *d, packet_t *pk)
{
    for (;;) {
        if (d->dead) return 0;
        if (!d->sec || d->off == SECTOR) {
            if (d->next >= d->n_sectors) { d->dead = 1; return 0; }
            d->sec = d->base + d->next * (uint64_t)SECTOR;
            d->next++;
            d->off = pack_header_size(d->sec);
            if (!d->off) { d->dead = 1; return 0; }
        }
        /* 24u start code, 8u stream id, 16u length (packet.c:97-108) */
        if (d->off + 6 > SECTOR) { d->dead = 1; return 0; }
        const uint8_t *h = d->sec + d->off;
        if (!(h[0] == 0 && h[1] == 0 && h[2] == 1)) { d->dead = 1; return 0; }
        const unsigned id = h[3], len = ((unsigned)h[4] << 8) | h[5];
        if (d->off + 6 + len > SECTOR) { d->dead = 1; return 0; }
        pk->data = h + 6;
        pk->len = len;
        pk->sector = (uint32_t)(d->next - 1);          /* packet.c:88 */
        d->off += 6 + len;
        if (id == 0xBD) return 1;                      /* packet.c:129-134 */
    }
}

/* audio packet header: "16p 8u" pad1 skip "8u 8p 8p 8u" (dvd-audio.c:1238-1248).
 * Returns offset of the first byte after pad_2_size, or -1 if the packet is
 * too short. */
static int audio_header(const packet_t *pk, unsigned *codec, unsigned *pad2)
{
    if (pk->len < 3) return -1;
    const unsigned pad1 = pk->data[2];
    if (pk->len < 3 + pad1 + 4) return -1;
    *codec = pk->data[3 + pad1];
    *pad2 = pk->data[3 + pad1 + 3];
    return (int)(3 + pad1 + 4);
}

/* ------------------------------------------------------------ output */

typedef struct {
    int32_t *v;
    uint64_t frames, cap;
    unsigned ch;
} out_t;

static int32_t *out_reserve(out_t *o, uint64_t add)
{
    if (o->frames + add > o->cap) {
        uint64_t nc = (o->frames + add) * 2 + 4096;
        o->v = realloc(o->v, sizeof(int32_t) * nc * o->ch);
        o->cap = nc;
    }
    return o->v + o->frames * o->ch;
}

/* ---------------------------------------------------------------- PCM */

/* The reference permutes every chunk (two frames) with a byte table and then
 * reads little-endian samples (src/pcm.c:103-166).  We rebuild that table from
 * the layout it encodes: a chunk is one or two groups of samples; a 16-bit group
 * is big-endian samples, a 24-bit group is all (high, middle) byte pairs
 * followed by all low bytes. */
void dvda_oracle_pcm_permutation(int bits, int channels, uint8_t *table)
{
    const int bytes = bits / 8, n = 2 * channels;
    int order[12], cnt = 0, first_group;
    int two_groups = (channels == 6) || (bits == 24 && channels >= 3);
    if (!two_groups) {
        for (int i = 0; i < n; i++) order[cnt++] = i;
        first_group = n;
    } else {
        const int hi = (channels == 6) ? 4 : channels;   /* group A = channels [2, hi) */
        for (int f = 0; f < 2; f++) for (int c = 2; c < hi; c++) order[cnt++] = f * channels + c;
        first_group = cnt;
        for (int f = 0; f < 2; f++) for (int c = 0; c < channels; c++)
            if (c < 2 || c >= hi) order[cnt++] = f * channels + c;
    }
    int i = 0;   /* input byte index */
    for (int g = 0; g < 2; g++) {
        const int a = g == 0 ? 0 : first_group, b = g == 0 ? first_group : n;
        if (a == b) continue;
        if (bytes == 2) {
            for (int k = a; k < b; k++) { table[i++] = (uint8_t)(order[k] * 2 + 1); table[i++] = (uint8_t)(order[k] * 2); }
        } else {
            for (int k = a; k < b; k++) { table[i++] = (uint8_t)(order[k] * 3 + 2); table[i++] = (uint8_t)(order[k] * 3 + 1); }
            for (int k = a; k < b; k++) table[i++] = (uint8_t)(order[k] * 3);
        }
    }
}

typedef struct {
    unsigned bps_index;       /* 0: 16-bit converter, 1: 24-bit converter (pcm.c:55-62) */
    unsigned channels, bytes_per_sample, chunk;
    uint8_t table[36];
} pcm_dec_t;

/* src/pcm.c:98-169 */
static unsigned pcm_decode_packet(const pcm_dec_t *d, const uint8_t *p, size_t len, out_t *o)
{
    unsigned frames = 0;
    uint8_t un[36];
    while (len >= d->chunk && d->chunk) {
        for (unsigned i = 0; i < d->chunk; i++) un[d->table[i]] = p[i];
        int32_t *dst = out_reserve(o, 2);
        const uint8_t *s = un;
        for (unsigned i = 0; i < 2 * d->channels; i++) {
            int32_t v;
            if (d->bps_index == 0) v = (int16_t)(s[0] | (s[1] << 8));
            else { v = s[0] | (s[1] << 8) | (s[2] << 16); if (v & 0x800000) v -= 0x1000000; }
            dst[i] = v;                    /* sample i -> frame i / ch, channel i % ch */
            s += d->bytes_per_sample;
        }
        o->frames += 2;
        frames += 2;
        p += d->chunk;
        len -= d->chunk;
    }
    return frames;
}

/* 9 parameter bytes (src/pcm.c:79-96) */
static int pcm_params(const uint8_t *p, size_t len, unsigned f[5])
{
    if (len < 9) return -1;
    f[0] = p[3] >> 4; f[1] = p[3] & 15; f[2] = p[4] >> 4; f[3] = p[4] & 15; f[4] = p[6];
    return 0;
}

/* src/dvd-audio.c:952-1082 */
static int decode_pcm_track(demux_t *dx, const packet_t *first, int hdr_off, unsigned pad2,
                            uint32_t pts_length, dvda_oracle_result *r)
{
    unsigned f[5];
    if (pcm_params(first->data + hdr_off, first->len - (size_t)hdr_off, f)) return 1;
    r->codec = 0;
    r->group_0_bps = f[0]; r->group_1_bps = f[1]; r->group_0_rate = f[2]; r->group_1_rate = f[3];
    r->channel_assignment = f[4];
    r->bits_per_sample = unpack_bps(f[0]);
    r->sample_rate = unpack_rate(f[2]);
    r->channels = unpack_channels(f[4]);
    if (!r->channels || pad2 < 9) return 1;       /* the reference would crash / underflow */

    pcm_dec_t d;
    memset(&d, 0, sizeof d);
    d.bps_index = r->bits_per_sample == 16 ? 0 : 1;
    d.channels = r->channels;
    d.bytes_per_sample = r->bits_per_sample / 8;
    d.chunk = d.bytes_per_sample * d.channels * 2;
    if (r->bits_per_sample != 16 && r->bits_per_sample != 24) return 1;   /* 20-bit: reference mis-indexes its table */
    dvda_oracle_pcm_permutation((int)r->bits_per_sample, (int)r->channels, d.table);

    const double total_d = (double)pts_length * (double)r->sample_rate / 90000.0;
    const uint64_t total = (uint64_t)lround(total_d);
    uint64_t remaining = total;
    out_t o = {NULL, 0, 0, r->channels};

    /* first packet: decoded at open, result not inspected (dvd-audio.c:998-1007) */
    {
        const size_t off = (size_t)hdr_off + pad2;
        unsigned got = 0;
        if (off <= first->len) got = pcm_decode_packet(&d, first->data + off, first->len - off, &o);
        remaining -= got < total ? got : total;
    }
    for (;;) {
        if (!remaining) break;                               /* dvd-audio.c:1022 */
        packet_t pk;
        if (!demux_next_audio(dx, &pk)) break;
        unsigned codec, p2;
        const int ho = audio_header(&pk, &codec, &p2);
        if (ho < 0 || codec != 0xA0) break;
        unsigned g[5];
        if (pcm_params(pk.data + ho, pk.len - (size_t)ho, g)) break;
        if (memcmp(f, g, sizeof f)) break;                   /* stream_parameters.h:31 */
        if (p2 < 9 || (size_t)ho + p2 > pk.len) break;
        const unsigned got = pcm_decode_packet(&d, pk.data + ho + p2, pk.len - (size_t)ho - p2, &o);
        remaining -= got < remaining ? got : remaining;
        if (!got) break;                                     /* dvd-audio.c:770-774 */
    }
    r->frames = o.frames;
    r->pcm = o.v;
    return 0;
}

/* ---------------------------------------------------------------- MLP */

typedef struct {
    unsigned order, shift;
    int coeff[8];
} filt_t;

typedef struct {
    filt_t fir, iir;
    int fstate[8]; unsigned flen;     /* most recent first */
    int istate[8]; unsigned ilen;
    int huff_offset;
    unsigned codebook, huff_lsbs;
} chan_t;

typedef struct {
    unsigned out_ch, lsb_bypass;
    int coeff[MAX_CH];
} matrix_t;

typedef struct {
    /* directory entry */
    unsigned extraword, nonrestart, checkdata, end;
    /* restart header */
    unsigned min_ch, max_ch, mmc, noise_shift, seed;
    int have_header;
    /* decoding parameters */
    unsigned flags[8], block_size, matrix_len;
    matrix_t mat[MAX_MAT];
    unsigned out_shift[MAX_CH], q[MAX_CH];
    chan_t ch[MAX_CH];
} ss_t;

typedef struct {
    unsigned params[5];
    int sync_seen;
    unsigned substreams;
    ss_t ss[2];
    /* per-AU work area */
    int32_t *chan[MAX_CH];   /* filtered samples of the AU per MLP channel */
    unsigned chan_len[MAX_CH];
    uint8_t *bypass[MAX_MAT];
    unsigned bypass_len[MAX_MAT];
    unsigned cap;
    int error;
} mlp_t;

static inline int mask_q(int x, unsigned q) { return q ? (int)((unsigned)(x >> q) << q) : x; }

static void au_buffers(mlp_t *m, unsigned need)
{
    if (need <= m->cap) return;
    m->cap = need * 2 + 256;
    for (int c = 0; c < MAX_CH; c++) m->chan[c] = realloc(m->chan[c], sizeof(int32_t) * m->cap);
    for (int k = 0; k < MAX_MAT; k++) m->bypass[k] = realloc(m->bypass[k], m->cap);
}

/* src/mlp.c:809-854 */
static int restart_header(br_t *b, ss_t *s)
{
    const unsigned sync = br_u(b, 13), noise_type = br_u(b, 1);
    br_skip(b, 16);
    s->min_ch = br_u(b, 4); s->max_ch = br_u(b, 4); s->mmc = br_u(b, 4);
    s->noise_shift = br_u(b, 4);
    s->seed = br_u(b, 23);
    br_skip(b, 19 + 1 + 8 + 16);
    if (b->err || sync != 0x18F5 || noise_type != 0) return 0;
    if (s->max_ch < s->min_ch || s->mmc < s->max_ch) return 0;
    if (s->mmc >= MAX_CH) return 0;                /* reference arrays hold 8 (and coeff[] overflows beyond 5) */
    for (unsigned c = 0; c <= s->mmc; c++) if (br_u(b, 6) > s->mmc) return 0;
    br_skip(b, 8);
    s->have_header = 1;
    return !b->err;
}

/* src/mlp.c:1029-1120 */
static int filter_params(br_t *b, filt_t *f, chan_t *c, int is_iir)
{
    const unsigned order = br_u(b, 4);
    if (order > 8) return 0;
    if (order == 0) {
        f->order = 0; f->shift = 0;
        if (is_iir) c->ilen = 0;
        return 1;
    }
    f->shift = br_u(b, 4);
    const unsigned bits = br_u(b, 5);
    if (bits < 1 || bits > 16) return 0;
    const unsigned cshift = br_u(b, 3);
    if (bits + cshift > 16) return 0;
    f->order = order;
    for (unsigned i = 0; i < order; i++) f->coeff[i] = (int)((unsigned)br_s(b, bits) << cshift);
    if (!is_iir) {
        if (br_u(b, 1)) return 0;
    } else {
        c->ilen = 0;
        if (br_u(b, 1)) {
            const unsigned sbits = br_u(b, 4), sshift = br_u(b, 4);
            /* sent oldest-last: the first value pairs with coeff[0] (mlp.c:1103-1107) */
            for (unsigned i = 0; i < order; i++) c->istate[i] = (int)((unsigned)br_s(b, sbits) << sshift);
            c->ilen = order;
        }
    }
    return !b->err;
}

/* src/mlp.c:856-1027 */
static int decoding_params(br_t *b, ss_t *s, int restart)
{
    if (restart) {
        if (br_u(b, 1)) for (int k = 0; k < 8; k++) s->flags[k] = br_u(b, 1);
        else for (int k = 0; k < 8; k++) s->flags[k] = 1;
    } else if (s->flags[0] && br_u(b, 1)) {
        for (int k = 0; k < 8; k++) s->flags[k] = br_u(b, 1);
    }
    if (s->flags[7] && br_u(b, 1)) {
        if ((s->block_size = br_u(b, 9)) < 8) return 0;
    } else if (restart) s->block_size = 8;

    if (s->flags[6] && br_u(b, 1)) {
        s->matrix_len = br_u(b, 4);
        if (s->matrix_len > MAX_MAT) return 0;       /* reference overruns its array */
        for (unsigned m = 0; m < s->matrix_len; m++) {
            matrix_t *M = &s->mat[m];
            if ((M->out_ch = br_u(b, 4)) > s->mmc) return 0;
            const unsigned frac = br_u(b, 4);
            if (frac > 14) return 0;
            M->lsb_bypass = br_u(b, 1);
            if (s->mmc + 3 > MAX_CH) return 0;       /* reference coeff[8] */
            for (unsigned c = 0; c < s->mmc + 3; c++)
                M->coeff[c] = br_u(b, 1) ? (int)((unsigned)br_s(b, frac + 2) << (14 - frac)) : 0;
        }
    } else if (restart) s->matrix_len = 0;

    if (s->flags[5] && br_u(b, 1)) {
        for (unsigned c = 0; c <= s->mmc; c++) s->out_shift[c] = (unsigned)br_s(b, 4);
    } else if (restart) memset(s->out_shift, 0, sizeof s->out_shift);

    if (s->flags[4] && br_u(b, 1)) {
        for (unsigned c = 0; c <= s->max_ch; c++) s->q[c] = br_u(b, 4);
    } else if (restart) memset(s->q, 0, sizeof s->q);

    for (unsigned c = s->min_ch; c <= s->max_ch; c++) {
        chan_t *C = &s->ch[c];
        if (br_u(b, 1)) {
            if (s->flags[3] && br_u(b, 1)) { if (!filter_params(b, &C->fir, C, 0)) return 0; }
            else if (restart) { C->fir.order = 0; C->fir.shift = 0; }
            if (s->flags[2] && br_u(b, 1)) { if (!filter_params(b, &C->iir, C, 1)) return 0; }
            else if (restart) { C->iir.order = 0; C->iir.shift = 0; C->ilen = 0; }
            if (s->flags[1] && br_u(b, 1)) C->huff_offset = br_s(b, 15);
            else if (restart) C->huff_offset = 0;
            C->codebook = br_u(b, 2);
            if ((C->huff_lsbs = br_u(b, 5)) > 24) return 0;
        } else if (restart) {
            C->fir.order = 0; C->fir.shift = 0;
            C->iir.order = 0; C->iir.shift = 0; C->ilen = 0;
            C->huff_offset = 0; C->codebook = 0; C->huff_lsbs = 24;
        }
    }
    return !b->err;
}

/* One block: src/mlp.c:741-807 (block), :1122-1241 (residuals), :1243-1306
 * (filter).  Appends block_size filtered samples per channel to m->chan[]. */
static unsigned decode_block(mlp_t *m, ss_t *s, br_t *b)
{
    if (br_u(b, 1)) {
        const int restart = (int)br_u(b, 1);
        if (restart && !restart_header(b, s)) return 0;
        if (!s->have_header) return 0;              /* reference would use garbage */
        if (!decoding_params(b, s, restart)) return 0;
    }
    if (b->err || !s->have_header) return 0;

    const unsigned n = s->block_size;
    int sho[MAX_CH];
    unsigned lsb_bits[MAX_CH];
    for (unsigned c = s->min_ch; c <= s->max_ch; c++) {
        const chan_t *C = &s->ch[c];
        if (C->huff_lsbs < s->q[c]) return 0;       /* unsigned underflow in the reference: G3 */
        const unsigned nb = C->huff_lsbs - s->q[c];
        lsb_bits[c] = nb;
        if (C->codebook) {
            const int ss = (int)nb + 2 - (int)C->codebook;
            sho[c] = C->huff_offset - 7 * (1 << nb) - (ss >= 0 ? (1 << ss) : 0);
        } else {
            sho[c] = C->huff_offset - (nb >= 1 ? (1 << (nb - 1)) : 0);
        }
        /* filter sanity (mlp.c:1260-1270) and history availability */
        if (C->fir.order + C->iir.order > 8) return 0;
        if (C->fir.shift > 0 && C->iir.shift > 0 && C->fir.shift != C->iir.shift) return 0;
        if (C->fir.order > C->flen || C->iir.order > C->ilen) return 0;   /* reference reads out of bounds: G1, G2 */
    }
    unsigned base = m->chan_len[s->min_ch];
    au_buffers(m, base + n);

    /* entropy decode: per frame bypass bits, then one residual per channel */
    for (unsigned i = 0; i < n; i++) {
        for (unsigned k = 0; k < s->matrix_len; k++) {
            au_buffers(m, m->bypass_len[k] + 1);
            m->bypass[k][m->bypass_len[k]++] = s->mat[k].lsb_bypass ? (uint8_t)br_u(b, 1) : 0;
        }
        for (unsigned c = s->min_ch; c <= s->max_ch; c++) {
            int msb = 0;
            if (s->ch[c].codebook) { msb = huffman(b, (int)s->ch[c].codebook); if (msb < 0) return 0; }
            const int lsb = (int)br_u(b, lsb_bits[c]);
            m->chan[c][base + i] = (int)((unsigned)((msb << lsb_bits[c]) + lsb + sho[c]) << s->q[c]);
        }
        if (b->err) return 0;
    }
    /* prediction filters, in place */
    for (unsigned c = s->min_ch; c <= s->max_ch; c++) {
        chan_t *C = &s->ch[c];
        const unsigned shift = (C->fir.shift > 0 && C->iir.shift > 0) ? C->fir.shift
                             : C->fir.order > 0 ? C->fir.shift : C->iir.shift;
        int32_t *x = m->chan[c] + base;
        for (unsigned i = 0; i < n; i++) {
            int64_t sum = 0;
            for (unsigned j = 0; j < C->fir.order; j++) sum += (int64_t)C->fir.coeff[j] * C->fstate[j];
            for (unsigned k = 0; k < C->iir.order; k++) sum += (int64_t)C->iir.coeff[k] * C->istate[k];
            const int ssum = (int)(sum >> shift);
            const int v = mask_q((int)((unsigned)ssum + (unsigned)x[i]), s->q[c]);
            x[i] = v;
            memmove(C->fstate + 1, C->fstate, 7 * sizeof(int));
            C->fstate[0] = v;
            if (C->flen < 8) C->flen++;
            memmove(C->istate + 1, C->istate, 7 * sizeof(int));
            C->istate[0] = (int)((unsigned)v - (unsigned)ssum);
            if (C->ilen < 8) C->ilen++;
        }
        m->chan_len[c] = base + n;
    }
    return n;
}

/* src/mlp.c:714-739 */
static unsigned decode_substream(mlp_t *m, ss_t *s, const uint8_t *p, size_t len)
{
    br_t b;
    br_open(&b, p, len);
    unsigned frames = 0;
    do {
        const unsigned n = decode_block(m, s, &b);
        if (!n) return 0;
        frames += n;
    } while (br_u(&b, 1) == 0 && !b.err);
    if (b.err) return 0;
    /* byte-align + optional 0xD234D234: nothing observable */
    return frames;
}

/* parity / CRC-8 (src/mlp.c:670-712, 1360-1399).  Returns 0 ok, else error bit. */
static int check_substream(const uint8_t *p, size_t len)
{
    if (!crc8_ready) crc8_build();
    uint8_t parity = 0, crc = 0x3C, fin = 0;
    for (size_t i = 0; i + 2 < len; i++) {
        parity ^= p[i];
        fin = crc ^ p[i];
        crc = crc8_tab[fin];
    }
    if ((uint8_t)(p[len - 2] ^ parity) != 0xA9) return DVDA_ORACLE_ERR_PARITY;
    if (p[len - 1] != fin) return DVDA_ORACLE_ERR_CRC;
    return 0;
}

/* noise + matrices + LSB bypass for one AU (src/mlp.c:1308-1358) */
static void rematrix(mlp_t *m, ss_t *s, unsigned n)
{
    for (unsigned i = 0; i < n; i++) {
        const unsigned sh = (s->seed >> 7) & 0xFFFF;
        const int n0 = (int)((unsigned)(int)(int8_t)(s->seed >> 15) << s->noise_shift);
        const int n1 = (int)((unsigned)(int)(int8_t)sh << s->noise_shift);
        s->seed = ((s->seed << 16) & 0xFFFFFFFFu) ^ sh ^ (sh << 5);
        for (unsigned k = 0; k < s->matrix_len; k++) {
            const matrix_t *M = &s->mat[k];
            int64_t sum = 0;
            for (unsigned c = 0; c <= s->mmc; c++) sum += (int64_t)m->chan[c][i] * M->coeff[c];
            sum += (int64_t)n0 * M->coeff[s->mmc + 1];
            sum += (int64_t)n1 * M->coeff[s->mmc + 2];
            m->chan[M->out_ch][i] = mask_q((int)(sum >> 14), s->q[M->out_ch]) + m->bypass[k][i];
        }
    }
}

/* channel placement (table at src/mlp.c:416-438): RIFF WAVE slot of MLP channel c */
static int wave_slot(unsigned assignment, unsigned c)
{
    if (assignment == 0x12 || assignment == 0x13) { static const int t[5] = {0, 1, 3, 4, 2}; return t[c]; }
    if (assignment == 0x14) { static const int t[6] = {0, 1, 4, 5, 2, 3}; return t[c]; }
    return (int)c;
}

/* One access unit without its 4-byte header: src/mlp.c:407-654.
 * Returns frames appended; 0 = dropped; sets m->error on damage. */
static unsigned decode_au(mlp_t *m, const uint8_t *p, size_t len, out_t *o)
{
    size_t pos = 0;
    /* major sync: 28 bytes (mlp.c:614-654) */
    if (len >= 28 && p[0] == 0xF8 && p[1] == 0x72 && p[2] == 0x6F && p[3] == 0xBB) {
        const unsigned ns = p[16] >> 4;
        if (ns == 1 || ns == 2) {
            unsigned f[5] = {p[4] >> 4, p[4] & 15u, p[5] >> 4, p[5] & 15u, p[7] & 31u};
            pos = 28;
            if (m->sync_seen) {
                if (memcmp(f, m->params, sizeof f)) return 0;        /* mlp.c:452-455 */
            } else {
                /* the decoder's own copy; the track parameters were fixed at open */
                m->sync_seen = 1;
                m->substreams = ns;
            }
        }
    }
    if (!m->sync_seen) { m->error |= DVDA_ORACLE_ERR_SYNTAX; return 0; }

    for (unsigned k = 0; k < m->substreams; k++) {
        if (pos + 2 > len) { m->error |= DVDA_ORACLE_ERR_SYNTAX; return 0; }
        ss_t *s = &m->ss[k];
        s->extraword = p[pos] >> 7; s->nonrestart = (p[pos] >> 6) & 1; s->checkdata = (p[pos] >> 5) & 1;
        s->end = ((((unsigned)p[pos] & 15u) << 8) | p[pos + 1]) * 2;
        pos += 2;
        if (s->extraword) pos += 2;
    }
    for (int c = 0; c < MAX_CH; c++) m->chan_len[c] = 0;

    unsigned frames[2] = {0, 0};
    size_t start = 0;
    for (unsigned k = 0; k < m->substreams; k++) {
        ss_t *s = &m->ss[k];
        if (s->end < start) { m->error |= DVDA_ORACLE_ERR_SYNTAX; return 0; }
        size_t sl = s->end - start;
        const uint8_t *sp = p + pos + start;
        if (pos + s->end > len) { m->error |= DVDA_ORACLE_ERR_SYNTAX; return 0; }
        /* substream 1 is checked iff substream 0 says so (mlp.c:543-545) */
        if (m->ss[0].checkdata) {
            if (sl < 2) { m->error |= DVDA_ORACLE_ERR_SYNTAX; return 0; }
            const int e = check_substream(sp, sl);
            if (e) { m->error |= e; return 0; }
            sl -= 2;
        }
        for (int q = 0; q < MAX_MAT; q++) m->bypass_len[q] = 0;      /* mlp.c:481-482, 552-553 */
        frames[k] = decode_substream(m, s, sp, sl);
        if (!frames[k]) { m->error |= DVDA_ORACLE_ERR_SYNTAX; return 0; }
        start = s->end;
    }
    if (m->substreams == 2 && frames[1] != frames[0]) { m->error |= DVDA_ORACLE_ERR_SYNTAX; return 0; }

    /* the LAST substream's end-of-AU parameters govern every channel (mlp.c:504-538, 575-608) */
    ss_t *g = &m->ss[m->substreams - 1];
    const unsigned n = frames[0];
    for (unsigned c = 0; c <= g->mmc; c++) {
        if (m->chan_len[c] != n) {                   /* a channel no substream produced */
            au_buffers(m, n);
            memset(m->chan[c], 0, sizeof(int32_t) * n);
        }
    }
    /* bypass arrays shorter than the AU: see DESIGN.md (reference reads stale memory) */
    for (unsigned k = 0; k < g->matrix_len; k++)
        while (m->bypass_len[k] < n) { au_buffers(m, n); m->bypass[k][m->bypass_len[k]++] = 0; }
    rematrix(m, g, n);
    for (unsigned c = 0; c <= g->mmc; c++)
        if (g->out_shift[c])
            for (unsigned i = 0; i < n; i++) m->chan[c][i] = (int)((unsigned)m->chan[c][i] << (g->out_shift[c] & 31));

    int32_t *dst = out_reserve(o, n);
    for (unsigned c = 0; c < o->ch; c++) {
        const int slot = wave_slot(m->params[4], c);
        for (unsigned i = 0; i < n; i++) dst[(size_t)i * o->ch + (unsigned)slot] = m->chan[c][i];
    }
    o->frames += n;
    return n;
}

/* elementary-stream queue */
typedef struct { uint8_t *v; size_t len, cap, rd; } esq_t;
static void esq_push(esq_t *q, const uint8_t *p, size_t n)
{
    if (q->len + n > q->cap) { q->cap = (q->len + n) * 2 + 4096; q->v = realloc(q->v, q->cap); }
    memcpy(q->v + q->len, p, n);
    q->len += n;
}

/* src/mlp.c:360-405: decode every complete AU in the queue */
static unsigned drain(mlp_t *m, esq_t *q, out_t *o, dvda_oracle_result *r)
{
    unsigned frames = 0;
    while (!m->error) {
        if (q->len - q->rd < 4) break;
        const uint8_t *h = q->v + q->rd;
        const size_t total = (size_t)((((unsigned)h[0] & 15u) << 8) | h[1]) * 2;
        if (total < 4) break;                        /* size underflows in the reference: stalls for good */
        if (q->len - q->rd < total) break;
        const unsigned n = decode_au(m, h + 4, total - 4, o);
        if (m->error) break;                         /* track ends in front of a damaged AU */
        q->rd += total;
        frames += n;
        r->access_units++;
        r->es_bytes += total;
    }
    return frames;
}

static int has_sync_at(const uint8_t *p) { return p[4] == 0xF8 && p[5] == 0x72 && p[6] == 0x6F && p[7] == 0xBB; }

/* next MLP packet's stream bytes into q (src/dvd-audio.c:1288-1316) */
static int enqueue_mlp_packet(demux_t *dx, esq_t *q)
{
    for (;;) {
        packet_t pk;
        if (!demux_next_audio(dx, &pk)) return 0;
        unsigned codec, pad2;
        const int ho = audio_header(&pk, &codec, &pad2);
        if (ho < 0) return 0;
        if (codec != 0xA1) continue;
        if ((size_t)ho + pad2 > pk.len) return 0;
        esq_push(q, pk.data + ho + pad2, pk.len - (size_t)ho - pad2);
        return 1;
    }
}

/* src/dvd-audio.c:1094-1227, 1250-1421 */
static int decode_mlp_track(demux_t *dx, const packet_t *first, int hdr_off, unsigned pad2,
                            uint32_t last_sector, dvda_oracle_result *r)
{
    esq_t q = {0};
    if ((size_t)hdr_off + pad2 > first->len) return 1;
    esq_push(&q, first->data + hdr_off + pad2, first->len - (size_t)hdr_off - pad2);

    /* locate_mlp_parameters: slide to the first position whose bytes 4..7 are the sync */
    size_t at = 0;
    for (;;) {
        while (q.len - at >= 8 && !has_sync_at(q.v + at)) at++;
        if (q.len - at >= 8) break;
        if (!enqueue_mlp_packet(dx, &q)) { free(q.v); return 1; }   /* reference asserts */
    }
    while (q.len - at < 18) if (!enqueue_mlp_packet(dx, &q)) { free(q.v); return 1; }
    const uint8_t *s = q.v + at;
    r->codec = 1;
    r->group_0_bps = s[8] >> 4; r->group_1_bps = s[8] & 15; r->group_0_rate = s[9] >> 4; r->group_1_rate = s[9] & 15;
    r->channel_assignment = s[11] & 31;
    r->bits_per_sample = unpack_bps(r->group_0_bps);
    r->sample_rate = unpack_rate(r->group_0_rate);
    r->channels = unpack_channels(r->channel_assignment);
    if (!r->channels) { free(q.v); return 1; }
    q.rd = at;

    mlp_t *m = calloc(1, sizeof *m);
    m->params[0] = r->group_0_bps; m->params[1] = r->group_1_bps; m->params[2] = r->group_0_rate;
    m->params[3] = r->group_1_rate; m->params[4] = r->channel_assignment;
    out_t o = {NULL, 0, 0, r->channels};

    drain(m, &q, &o, r);                              /* at open; result not inspected */
    while (!m->error) {
        packet_t pk;
        if (!demux_next_audio(dx, &pk)) break;
        unsigned codec, p2;
        if (pk.sector > last_sector) {
            /* mlp_data_to_major_sync: only the bytes in front of the next sync */
            const int ho = audio_header(&pk, &codec, &p2);
            if (ho < 0 || codec != 0xA1 || (size_t)ho + p2 > pk.len) break;
            esq_t t = {0};
            esq_push(&t, pk.data + ho + p2, pk.len - (size_t)ho - p2);
            size_t k = 0;
            for (;;) {
                while (t.len - k >= 8 && !has_sync_at(t.v + k)) k++;
                if (t.len - k >= 8) break;
                if (!enqueue_mlp_packet(dx, &t)) break;              /* reference asserts */
            }
            if (k) { esq_push(&q, t.v, k); drain(m, &q, &o, r); }
            free(t.v);
            break;
        }
        const int ho = audio_header(&pk, &codec, &p2);
        if (ho < 0 || codec != 0xA1 || (size_t)ho + p2 > pk.len) break;
        esq_push(&q, pk.data + ho + p2, pk.len - (size_t)ho - p2);
        if (!drain(m, &q, &o, r)) break;              /* dvd-audio.c:770-774 */
        /* compact the queue now and then */
        if (q.rd > (1u << 20)) { memmove(q.v, q.v + q.rd, q.len - q.rd); q.len -= q.rd; q.rd = 0; }
    }
    r->error_flags = m->error;
    r->frames = o.frames;
    r->pcm = o.v;
    for (int c = 0; c < MAX_CH; c++) free(m->chan[c]);
    for (int k = 0; k < MAX_MAT; k++) free(m->bypass[k]);
    free(m);
    free(q.v);
    return 0;
}

/* ------------------------------------------------------------- entry */

int dvda_oracle_decode_track(const uint8_t *sectors, uint64_t n_sectors,
                             uint32_t first_sector, uint32_t last_sector,
                             uint32_t pts_length, dvda_oracle_result *out)
{
    memset(out, 0, sizeof *out);
    out->status = DVDA_ORACLE_NO_AUDIO;
    if (first_sector >= n_sectors) return 1;            /* aob_reader_seek fails (aob.c:181-199) */
    demux_t dx = {sectors, n_sectors, first_sector, NULL, 0, 0};
    packet_t pk;
    if (!demux_next_audio(&dx, &pk)) return 1;
    unsigned codec, pad2;
    const int ho = audio_header(&pk, &codec, &pad2);
    if (ho < 0) return 1;
    int rc;
    if (codec == 0xA0) rc = decode_pcm_track(&dx, &pk, ho, pad2, pts_length, out);
    else if (codec == 0xA1) rc = decode_mlp_track(&dx, &pk, ho, pad2, last_sector, out);
    else return 1;
    if (rc) { free(out->pcm); out->pcm = NULL; return 1; }
    out->status = DVDA_ORACLE_OK;
    return 0;
}

void dvda_oracle_free(dvda_oracle_result *r)
{
    free(r->pcm);
    r->pcm = NULL;
}
