/* dvd-audio.c — host side of the B200-native libdvd-audio: the public API of
 * include/dvd-audio.h in plain C.
 *
 * The disc model (AUDIO_TS.IFO / ATS_xx_0.IFO tables, titles, tracks, sector
 * ranges) is ordinary host code and mirrors the behaviour of the reference's
 * src/dvd-audio.c:324-595, 824-950 and src/audio_ts.c:38-73; it is written from
 * the table layouts, not from the reference's bit-reader
 * calls.  Everything from dvda_open_track_reader() on is different: the track's
 * sectors are read into pinned memory and handed to the CUDA engine
 * (include/dvdagpu.h), which decodes the whole track on the GPU; dvda_read()
 * then copies slices of the result.  There is no CPU decode path: if the engine
 * cannot be created, dvda_open_track_reader() fails.
 *
 * Environment: DVDA_B200_DEVICE selects the CUDA device (default 0).
 */
#define _POSIX_C_SOURCE 200809L
#include "dvd-audio.h"
#include "dvdagpu.h"

#include <ctype.h>
#include <dirent.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#define SECTOR_SIZE 2048u
#define MAX_AOBS 9

/* ------------------------------------------------------------ disc model */

struct disc_path {
    char *audio_ts;
    char *device;
};

struct ifo_track { unsigned index_number, pts_index, pts_length; };
struct ifo_index { unsigned first_sector, last_sector; };
struct ifo_title {
    unsigned track_count, index_count, pts_length;
    struct ifo_track track[256];
    struct ifo_index index[256];
};

struct DVDA_s { struct disc_path disc; unsigned titleset_count; };
struct DVDA_Titleset_s {
    struct disc_path disc;
    unsigned number, title_count;
    struct ifo_title *title;
};
struct title_track { unsigned pts_index, pts_length, first_sector, last_sector; };
struct DVDA_Title_s {
    struct disc_path disc;
    unsigned titleset, number, track_count, pts_length;
    struct title_track tracks[256];
};
struct DVDA_Track_s {
    struct disc_path disc;
    unsigned titleset, title, number;
    struct title_track t;
};
struct DVDA_Track_Reader_s {
    dvda_codec_t codec;
    unsigned bits, rate, channels, assignment;
    unsigned long long frames, cursor;
    int *pcm;                   /* frames * channels, pinned */
    size_t pcm_bytes;
};

static char *dup_str(const char *s)
{
    if (!s) return NULL;
    size_t n = strlen(s) + 1;
    char *d = malloc(n);
    if (d) memcpy(d, s, n);
    return d;
}
static void path_set(struct disc_path *p, const char *audio_ts, const char *device)
{
    p->audio_ts = dup_str(audio_ts);
    p->device = dup_str(device);
}
static void path_free(struct disc_path *p) { free(p->audio_ts); free(p->device); }

/* case-insensitive lookup of `name` inside the AUDIO_TS directory
   (behaviour of src/audio_ts.c:38-73) */
static char *find_file(const char *dir, const char *name)
{
    DIR *d = opendir(dir);
    if (!d) return NULL;
    char *found = NULL;
    struct dirent *e;
    while (!found && (e = readdir(d)) != NULL) {
        const char *a = name, *b = e->d_name;
        while (*a && *b && toupper((unsigned char)*a) == toupper((unsigned char)*b)) { a++; b++; }
        if (*a == 0 && *b == 0) {
            size_t n = strlen(dir) + 1 + strlen(e->d_name) + 1;
            found = malloc(n);
            if (found) snprintf(found, n, "%s/%s", dir, e->d_name);
        }
    }
    closedir(d);
    return found;
}

static unsigned be16(const unsigned char *p) { return ((unsigned)p[0] << 8) | p[1]; }
static unsigned be32(const unsigned char *p)
{
    return ((unsigned)p[0] << 24) | ((unsigned)p[1] << 16) | ((unsigned)p[2] << 8) | p[3];
}

static unsigned char *slurp(const char *path, size_t *len)
{
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    unsigned char *buf = NULL;
    size_t cap = 0, n = 0;
    for (;;) {
        if (n == cap) {
            cap = cap ? cap * 2 : 8192;
            unsigned char *nb = realloc(buf, cap);
            if (!nb) { free(buf); fclose(f); return NULL; }
            buf = nb;
        }
        size_t got = fread(buf + n, 1, cap - n, f);
        if (!got) break;
        n += got;
    }
    fclose(f);
    *len = n;
    return buf;
}

DVDA *dvda_open(const char *audio_ts_path, const char *device)
{
    if (!audio_ts_path) return NULL;
    char *ifo = find_file(audio_ts_path, "AUDIO_TS.IFO");
    if (!ifo) return NULL;
    size_t len = 0;
    unsigned char *b = slurp(ifo, &len);
    free(ifo);
    /* identifier, then the title set count in byte 63 of a >= 104 byte table
       (the reference's parse consumes 104 bytes) */
    unsigned count = 0;
    if (b && len >= 104 && !memcmp(b, "DVDAUDIO-AMG", 12)) count = b[63];
    free(b);
    if (!count) return NULL;
    DVDA *d = malloc(sizeof *d);
    if (!d) return NULL;
    path_set(&d->disc, audio_ts_path, device);
    d->titleset_count = count;
    return d;
}

void dvda_close(DVDA *dvda)
{
    if (!dvda) return;
    path_free(&dvda->disc);
    free(dvda);
}

unsigned dvda_titleset_count(const DVDA *dvda) { return dvda->titleset_count; }

/* ATS_xx_0.IFO.  Returns 0 if any table runs off the file —
   the reference reports "I/O error" on stderr and fails the open. */
static int parse_ats(const unsigned char *b, size_t len, DVDA_Titleset *ts)
{
    if (len < 12 || memcmp(b, "DVDAUDIO-ATS", 12)) return 0;
    size_t p = SECTOR_SIZE;
    if (p + 8 > len) return 0;
    ts->title_count = be16(b + p);
    p += 8;
    ts->title = calloc(ts->title_count ? ts->title_count : 1, sizeof *ts->title);
    if (!ts->title) return 0;
    for (unsigned i = 0; i < ts->title_count; i++, p += 8) {
        if (p + 8 > len) return 0;
        const size_t tab = (size_t)SECTOR_SIZE + be32(b + p + 4);
        struct ifo_title *t = &ts->title[i];
        if (tab + 16 > len) return 0;
        t->track_count = b[tab + 2];
        t->index_count = b[tab + 3];
        t->pts_length = be32(b + tab + 4);
        const size_t ptrs = tab + be16(b + tab + 12);
        size_t q = tab + 16;
        for (unsigned k = 0; k < t->track_count; k++, q += 20) {
            if (q + 20 > len) return 0;
            t->track[k].index_number = b[q + 4];
            t->track[k].pts_index = be32(b + q + 6);
            t->track[k].pts_length = be32(b + q + 10);
        }
        q = ptrs;
        for (unsigned k = 0; k < t->index_count; k++, q += 12) {
            if (q + 12 > len) return 0;
            t->index[k].first_sector = be32(b + q + 4);
            t->index[k].last_sector = be32(b + q + 8);
        }
    }
    return 1;
}

DVDA_Titleset *dvda_open_titleset(DVDA *dvda, unsigned titleset)
{
    char name[16];
    snprintf(name, sizeof name, "ATS_%02u_0.IFO", titleset > 99 ? 99 : titleset);
    char *path = find_file(dvda->disc.audio_ts, name);
    if (!path) return NULL;
    size_t len = 0;
    unsigned char *b = slurp(path, &len);
    free(path);
    if (!b) return NULL;
    DVDA_Titleset *ts = calloc(1, sizeof *ts);
    if (!ts) { free(b); return NULL; }
    ts->number = titleset;
    if (!parse_ats(b, len, ts)) {
        if (len >= 12 && !memcmp(b, "DVDAUDIO-ATS", 12)) fprintf(stderr, "I/O error\n");
        else fprintf(stderr, "I/O error\n");
        free(ts->title);
        free(ts);
        free(b);
        return NULL;
    }
    free(b);
    path_set(&ts->disc, dvda->disc.audio_ts, dvda->disc.device);
    return ts;
}

void dvda_close_titleset(DVDA_Titleset *ts)
{
    if (!ts) return;
    path_free(&ts->disc);
    free(ts->title);
    free(ts);
}

unsigned dvda_titleset_number(const DVDA_Titleset *ts) { return ts->number; }
unsigned dvda_title_count(const DVDA_Titleset *ts) { return ts->title_count; }

DVDA_Title *dvda_open_title(DVDA_Titleset *ts, unsigned title_num)
{
    if (title_num == 0 || title_num > ts->title_count) return NULL;
    const struct ifo_title *it = &ts->title[title_num - 1];
    DVDA_Title *t = calloc(1, sizeof *t);
    if (!t) return NULL;
    path_set(&t->disc, ts->disc.audio_ts, ts->disc.device);
    t->titleset = ts->number;
    t->number = title_num;
    t->track_count = it->track_count;
    t->pts_length = it->pts_length;
    /* A track ends one sector before the next track starts; the last track of
       a title runs to the next title's first track (or its own index's end,
       whichever is later); the very last track ends with its own index
       (behaviour of src/dvd-audio.c:459-499). */
    for (unsigned i = 0; i < it->track_count; i++) {
        const struct ifo_index *ix = &it->index[(it->track[i].index_number - 1) & 255];
        struct title_track *o = &t->tracks[i];
        o->pts_index = it->track[i].pts_index;
        o->pts_length = it->track[i].pts_length;
        o->first_sector = ix->first_sector;
        if (i + 1 < it->track_count) {
            o->last_sector = it->index[(it->track[i + 1].index_number - 1) & 255].first_sector - 1;
        } else if (title_num == ts->title_count || ts->title[title_num].track_count == 0) {
            o->last_sector = ix->last_sector;
        } else {
            const struct ifo_title *nt = &ts->title[title_num];
            const unsigned nf = nt->index[(nt->track[0].index_number - 1) & 255].first_sector - 1;
            o->last_sector = nf > ix->last_sector ? nf : ix->last_sector;
        }
    }
    return t;
}

void dvda_close_title(DVDA_Title *t)
{
    if (!t) return;
    path_free(&t->disc);
    free(t);
}

unsigned dvda_title_number(const DVDA_Title *t) { return t->number; }
unsigned dvda_track_count(const DVDA_Title *t) { return t->track_count; }
unsigned dvda_title_pts_length(const DVDA_Title *t) { return t->pts_length; }

DVDA_Track *dvda_open_track(DVDA_Title *title, unsigned track_num)
{
    if (track_num == 0 || track_num > title->track_count) return NULL;
    DVDA_Track *k = calloc(1, sizeof *k);
    if (!k) return NULL;
    path_set(&k->disc, title->disc.audio_ts, title->disc.device);
    k->titleset = title->titleset;
    k->title = title->number;
    k->number = track_num;
    k->t = title->tracks[track_num - 1];
    return k;
}

void dvda_close_track(DVDA_Track *k)
{
    if (!k) return;
    path_free(&k->disc);
    free(k);
}

unsigned dvda_track_number(const DVDA_Track *k) { return k->number; }
unsigned dvda_track_pts_index(const DVDA_Track *k) { return k->t.pts_index; }
unsigned dvda_track_pts_length(const DVDA_Track *k) { return k->t.pts_length; }
unsigned dvda_track_first_sector(const DVDA_Track *k) { return k->t.first_sector; }
unsigned dvda_track_last_sector(const DVDA_Track *k) { return k->t.last_sector; }

/* ------------------------------------------------------------ AOB access */

/* the title set's ATS_tt_1.AOB .. ATS_tt_9.AOB as one run of sectors
   (behaviour of src/aob.c:90-123, 181-213) */
struct aob_set {
    FILE *file[MAX_AOBS];
    unsigned long long sectors[MAX_AOBS];
    unsigned count;
    unsigned long long total;
};

static void aobs_close(struct aob_set *a)
{
    for (unsigned i = 0; i < a->count; i++) fclose(a->file[i]);
    a->count = 0;
}

static void aobs_open(struct aob_set *a, const char *dir, unsigned titleset)
{
    memset(a, 0, sizeof *a);
    for (unsigned n = 1; n <= MAX_AOBS; n++) {
        char name[16];
        snprintf(name, sizeof name, "ATS_%02u_%u.AOB", titleset % 100, n);
        char *path = find_file(dir, name);
        if (!path) break;
        struct stat st;
        FILE *f = NULL;
        if (stat(path, &st) == 0) f = fopen(path, "rb");
        free(path);
        if (!f) break;
        a->file[a->count] = f;
        a->sectors[a->count] = (unsigned long long)st.st_size / SECTOR_SIZE;
        a->total += a->sectors[a->count];
        a->count++;
    }
}

/* reads sectors [first, first + n) into dst; returns sectors read */
static unsigned long long aobs_read(struct aob_set *a, unsigned long long first, unsigned long long n, unsigned char *dst)
{
    unsigned long long done = 0, base = 0;
    for (unsigned i = 0; i < a->count && done < n; i++) {
        const unsigned long long end = base + a->sectors[i];
        const unsigned long long want = first + done;
        if (want < end) {
            unsigned long long take = end - want;
            if (take > n - done) take = n - done;
            if (fseeko(a->file[i], (off_t)((want - base) * SECTOR_SIZE), SEEK_SET)) break;
            const size_t got = fread(dst + done * SECTOR_SIZE, SECTOR_SIZE, (size_t)take, a->file[i]);
            done += got;
            if (got != take) break;
        }
        base = end;
    }
    return done;
}

/* ------------------------------------------------------- engine + buffers */

static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static dvdagpu_ctx *g_engine;
static int g_engine_failed;

/* small cache of pinned buffers: page-locking is slow, tracks come in a row */
#define POOL_SLOTS 4
static struct { void *p; size_t bytes; } g_pool[POOL_SLOTS];

static void *pool_take(size_t bytes)
{
    int best = -1;
    for (int i = 0; i < POOL_SLOTS; i++)
        if (g_pool[i].p && g_pool[i].bytes >= bytes && (best < 0 || g_pool[i].bytes < g_pool[best].bytes)) best = i;
    if (best >= 0) {
        void *p = g_pool[best].p;
        g_pool[best].p = NULL;
        return p;
    }
    return dvdagpu_host_alloc(bytes);
}

static void pool_give(void *p, size_t bytes)
{
    if (!p) return;
    int slot = -1;
    for (int i = 0; i < POOL_SLOTS; i++) if (!g_pool[i].p) { slot = i; break; }
    if (slot < 0) {
        /* evict the smallest */
        slot = 0;
        for (int i = 1; i < POOL_SLOTS; i++) if (g_pool[i].bytes < g_pool[slot].bytes) slot = i;
        if (g_pool[slot].bytes >= bytes) { dvdagpu_host_free(p); return; }
        dvdagpu_host_free(g_pool[slot].p);
    }
    g_pool[slot].p = p;
    g_pool[slot].bytes = bytes;
}

static dvdagpu_ctx *engine(void)
{
    if (!g_engine && !g_engine_failed) {
        const char *dev = getenv("DVDA_B200_DEVICE");
        g_engine = dvdagpu_create(dev ? atoi(dev) : 0);
        if (!g_engine) {
            g_engine_failed = 1;
            fprintf(stderr, "libdvd-audio (B200): %s\n", dvdagpu_last_error());
        }
    }
    return g_engine;
}

/* ------------------------------------------------------------ track reader */

static size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

DVDA_Track_Reader *dvda_open_track_reader(const DVDA_Track *track)
{
    struct aob_set aobs;
    aobs_open(&aobs, track->disc.audio_ts, track->titleset);
    const unsigned long long first = track->t.first_sector;
    if (!aobs.count || first >= aobs.total) { aobs_close(&aobs); return NULL; }

    DVDA_Track_Reader *r = NULL;
    pthread_mutex_lock(&g_lock);
    dvdagpu_ctx *eng = engine();
    if (!eng) goto out;

    /* Read the track's sectors plus a margin for the run to the next major
       sync into pinned memory; widen the window while the engine reports that
       it ran out.  A probe of the first sectors tells the sample format, which
       sizes the PCM buffer (from the track's PTS length) for the pipelined
       decode; if that estimate is too small the one-piece path is used. */
    unsigned long long last = track->t.last_sector < first ? first : track->t.last_sector;
    unsigned long long margin = 64;
    for (;;) {
        unsigned long long stop = last + 1 + margin;
        if (stop > aobs.total) stop = aobs.total;
        const unsigned long long n = stop - first;
        const size_t sec_bytes = round_up((size_t)n * SECTOR_SIZE, 1 << 20);
        unsigned char *sec = pool_take(sec_bytes);
        if (!sec) break;
        const unsigned long long got = aobs_read(&aobs, first, n, sec);
        dvdagpu_track_desc desc = {0, 0, track->t.pts_length, 0};
        desc.last_sector = (uint32_t)(track->t.last_sector >= first ? track->t.last_sector - first : 0);
        dvdagpu_track_result res;
        memset(&res, 0, sizeof res);
        int rc = got ? 0 : -1;
        int *pcm = NULL;
        size_t pcm_bytes = 0;
        int have_pcm = 0;
        if (!rc && got > 4096) {
            /* long track: probe the format on a short window, then decode with overlapped copies */
            dvdagpu_track_desc pd = {0, 63, track->t.pts_length, 0};
            dvdagpu_track_result pr;
            if (!dvdagpu_decode_host(eng, sec, 256, 1, &pd, &pr) && pr.status == DVDAGPU_TRACK_OK &&
                pr.codec == 1 && pr.sample_rate && pr.channels) {
                const double seconds = (double)track->t.pts_length / PTS_PER_SECOND + 2.0;
                const unsigned long long cap = (unsigned long long)(seconds * pr.sample_rate) * pr.channels;
                pcm_bytes = round_up((size_t)cap * sizeof(int) + 1, 1 << 20);
                pcm = pool_take(pcm_bytes);
                if (pcm) {
                    rc = dvdagpu_decode_track_pipelined(eng, sec, got, &desc, 0, pcm, pcm_bytes / sizeof(int), &res);
                    if (rc == 0) have_pcm = 1;
                    else { pool_give(pcm, pcm_bytes); pcm = NULL; rc = 0; }   /* too small: one piece */
                }
            }
        }
        if (!rc && !have_pcm) rc = dvdagpu_decode_host(eng, sec, got, 1, &desc, &res);
        pool_give(sec, sec_bytes);
        if (rc) {
            if (got) fprintf(stderr, "libdvd-audio (B200): %s\n", dvdagpu_last_error());
            if (pcm) pool_give(pcm, pcm_bytes);
            break;
        }
        if (res.status != DVDAGPU_TRACK_OK) { if (pcm) pool_give(pcm, pcm_bytes); break; }
        if (res.truncated && stop < aobs.total) { if (pcm) pool_give(pcm, pcm_bytes); margin *= 8; continue; }

        if (res.error_flags & DVDAGPU_ERR_PARITY) fprintf(stderr, "parity mismatch\n");
        if (res.error_flags & DVDAGPU_ERR_CRC) fprintf(stderr, "CRC-8 mismatch\n");
        r = calloc(1, sizeof *r);
        if (!r) { if (pcm) pool_give(pcm, pcm_bytes); break; }
        r->codec = res.codec ? DVDA_MLP : DVDA_PCM;
        r->bits = res.bits_per_sample;
        r->rate = res.sample_rate;
        r->channels = res.channels;
        r->assignment = res.channel_assignment;
        r->frames = res.frames;
        if (have_pcm) {
            r->pcm = pcm;
            r->pcm_bytes = pcm_bytes;
        } else {
            r->pcm_bytes = round_up((size_t)res.frames * res.channels * sizeof(int) + 1, 1 << 20);
            r->pcm = pool_take(r->pcm_bytes);
            if (!r->pcm || dvdagpu_fetch(eng, res.pcm_offset, res.frames * res.channels, r->pcm)) {
                pool_give(r->pcm, r->pcm_bytes);
                free(r);
                r = NULL;
            }
        }
        break;
    }
out:
    pthread_mutex_unlock(&g_lock);
    aobs_close(&aobs);
    return r;
}

void dvda_close_track_reader(DVDA_Track_Reader *r)
{
    if (!r) return;
    pthread_mutex_lock(&g_lock);
    pool_give(r->pcm, r->pcm_bytes);
    pthread_mutex_unlock(&g_lock);
    free(r);
}

dvda_codec_t dvda_codec(const DVDA_Track_Reader *r) { return r->codec; }
unsigned dvda_bits_per_sample(const DVDA_Track_Reader *r) { return r->bits; }
unsigned dvda_sample_rate(const DVDA_Track_Reader *r) { return r->rate; }
unsigned dvda_channel_count(const DVDA_Track_Reader *r) { return r->channels; }

unsigned dvda_riff_wave_channel_mask(const DVDA_Track_Reader *r)
{
    /* speaker bits: FL FR FC LFE BL BR .. BC (0x100); one mask per channel
       assignment 0..20 (behaviour of src/dvd-audio.c:689-749) */
    enum { FL = 1, FR = 2, FC = 4, LFE = 8, BL = 0x10, BR = 0x20, BC = 0x100 };
    static const unsigned mask[21] = {
        FC,
        FL | FR,
        FL | FR | BC,
        FL | FR | BL | BR,
        FL | FR | LFE,
        FL | FR | LFE | BC,
        FL | FR | LFE | BL | BR,
        FL | FR | FC,
        FL | FR | FC | BC,
        FL | FR | FC | BL | BR,
        FL | FR | FC | LFE,
        FL | FR | FC | LFE | BC,
        FL | FR | FC | LFE | BL | BR,
        FL | FR | FC | BC,
        FL | FR | FC | BL | BR,
        FL | FR | FC | LFE,
        FL | FR | FC | LFE | BC,
        FL | FR | FC | LFE | BL | BR,
        FL | FR | BL | BR | LFE,
        FL | FR | BL | BR | FC,
        FL | FR | BL | BR | FC | LFE
    };
    return r->assignment <= 20 ? mask[r->assignment] : 0;
}

unsigned dvda_read(DVDA_Track_Reader *r, unsigned pcm_frames, int buffer[])
{
    if (!pcm_frames) return 0;
    unsigned long long left = r->frames - r->cursor;
    unsigned n = left < pcm_frames ? (unsigned)left : pcm_frames;
    if (n) {
        memcpy(buffer, r->pcm + r->cursor * r->channels, (size_t)n * r->channels * sizeof(int));
        r->cursor += n;
    }
    return n;
}
