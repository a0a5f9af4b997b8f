/* dvd-audio.c — host side of the B200-native libdvd-audio: the public API of
 * include/dvd-audio.h in plain C.
 *
 * The disc model (AUDIO_TS.IFO / ATS_xx_0.IFO tables, titles, tracks, sector
 * ranges) is ordinary host code and mirrors the behaviour of the reference's
 * src/dvd-audio.c:324-595, 824-950 and src/audio_ts.c:38-73; it is written from
 * the table layouts, not from the reference's bit-reader
 * calls.  Everything from dvda_open_track_reader() on is different: a reader
 * cuts its track into parts of a few thousand sectors, reads each part's
 * sectors into pinned memory and hands it to a pool of CUDA engine contexts
 * (include/dvdagpu.h; one worker thread per context, on one or several GPUs),
 * which decode a few parts ahead of dvda_read(); dvda_read() copies from the
 * part at the cursor.  Memory per reader is bounded by the parts in flight,
 * whatever the track's length.  There is no CPU decode path: if no engine can
 * be created, dvda_open_track_reader() fails.
 *
 * Environment: DVDA_B200_DEVICES ("all" or a list; default DVDA_B200_DEVICE or
 * 0), DVDA_B200_CONTEXTS (contexts per device, default 2),
 * DVDA_B200_PART_SECTORS (default 8192).
 */
#define _POSIX_C_SOURCE 200809L
#include "dvd-audio.h"
#include "dvdagpu.h"

#include <ctype.h>
#include <dirent.h>
#include <fcntl.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

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
        fprintf(stderr, "I/O error\n");
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
   (behaviour of src/aob.c:90-123, 181-213).  Positional reads on file
   descriptors: several engine workers read their parts at the same time. */
struct aob_set {
    int fd[MAX_AOBS];
    unsigned long long sectors[MAX_AOBS];
    unsigned count;
    unsigned long long total;
};

static void aobs_close(struct aob_set *a)
{
    for (unsigned i = 0; i < a->count; i++) close(a->fd[i]);
    a->count = 0;
}

static void aobs_open(struct aob_set *a, const char *dir, unsigned titleset)
{
    memset(a, 0, sizeof *a);
    for (unsigned n = 1; n <= MAX_AOBS; n++) {
        char name[16];
        snprintf(name, sizeof name, "ATS_%02u_%u.AOB", titleset > 99 ? 99 : titleset, n);
        char *path = find_file(dir, name);
        if (!path) break;
        struct stat st;
        int fd = -1;
        if (stat(path, &st) == 0) fd = open(path, O_RDONLY);
        free(path);
        if (fd < 0) break;
        a->fd[a->count] = fd;
        a->sectors[a->count] = (unsigned long long)st.st_size / SECTOR_SIZE;
        a->total += a->sectors[a->count];
        a->count++;
    }
}

/* reads sectors [first, first + n) into dst; returns sectors read */
static unsigned long long aobs_read(const struct aob_set *a, unsigned long long first, unsigned long long n, unsigned char *dst)
{
    unsigned long long done = 0, base = 0;
    for (unsigned i = 0; i < a->count && done < n; i++) {
        const unsigned long long end = base + a->sectors[i];
        const unsigned long long want = first + done;
        if (want < end) {
            unsigned long long take = end - want;
            if (take > n - done) take = n - done;
            size_t bytes = (size_t)take * SECTOR_SIZE, got = 0;
            off_t at = (off_t)((want - base) * SECTOR_SIZE);
            while (got < bytes) {
                const ssize_t r = pread(a->fd[i], dst + done * SECTOR_SIZE + got, bytes - got, at + (off_t)got);
                if (r <= 0) break;
                got += (size_t)r;
            }
            done += got / SECTOR_SIZE;
            if (got != bytes) break;
        }
        base = end;
    }
    return done;
}

/* ------------------------------------------------------- the engine pool
 *
 * One worker thread per engine context.  DVDA_B200_DEVICES picks the CUDA
 * devices ("all", or a comma-separated list; default: DVDA_B200_DEVICE or 0),
 * DVDA_B200_CONTEXTS the contexts per device (default 2: while one part's
 * samples travel to the host the next part is decoded).  A track reader cuts
 * its track into parts of DVDA_B200_PART_SECTORS sectors (default 8192 = 16 MB)
 * and deals them to the workers in turn, a few parts ahead of dvda_read(): the
 * parts of a long track are decoded on all devices at once and gathered in
 * order in the reader — a host-side gather, no other exchange.  Pinned memory
 * per reader is bounded by the parts in flight, whatever the track's length. */

struct part_job;
struct worker {
    pthread_t thread;
    dvdagpu_ctx *ctx;
    int device;
    pthread_mutex_t lock;
    pthread_cond_t wake;
    struct part_job *head, *tail;       /* queued jobs */
    int quit;
};

/* one part of a track: a window of sectors in, interleaved samples out */
struct part_job {
    struct part_job *next;
    const struct aob_set *aobs;
    unsigned long long first, n_sectors;        /* window in the title set's sector run */
    dvdagpu_track_desc desc;                    /* relative to the window */
    unsigned char *sec;  size_t sec_cap;        /* pinned */
    int *pcm;            size_t pcm_cap;        /* pinned, in ints */
    unsigned long long got;                     /* sectors read */
    dvdagpu_track_result res;
    int rc;                                     /* 0 ok, else the engine failed (message printed) */
    int done;
    pthread_mutex_t lock;
    pthread_cond_t cond;
};

static pthread_mutex_t g_pool_lock = PTHREAD_MUTEX_INITIALIZER;
static struct worker *g_workers;
static unsigned g_nworkers, g_next_worker;
static int g_pool_failed;

static size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static void run_job(struct worker *w, struct part_job *j)
{
    j->rc = -1;
    memset(&j->res, 0, sizeof j->res);
    j->got = aobs_read(j->aobs, j->first, j->n_sectors, j->sec);
    if (j->got) {
        /* the window may be shorter than asked for (end of the title set) */
        if (j->desc.last_sector >= j->got) j->desc.last_sector = (uint32_t)(j->got - 1);
        j->rc = dvdagpu_decode_host(w->ctx, j->sec, j->got, 1, &j->desc, &j->res);
        if (j->rc) fprintf(stderr, "libdvd-audio (B200): %s\n", dvdagpu_last_error());
        if (!j->rc && j->res.status == DVDAGPU_TRACK_OK) {
            const size_t n = (size_t)(j->res.frames * j->res.channels);
            if (n > j->pcm_cap) {
                dvdagpu_host_free(j->pcm);
                j->pcm_cap = round_up(n + n / 8 + 1, 1 << 18);
                j->pcm = dvdagpu_host_alloc(j->pcm_cap * sizeof(int));
                if (!j->pcm) { j->pcm_cap = 0; j->rc = -1; fprintf(stderr, "libdvd-audio (B200): %s\n", dvdagpu_last_error()); }
            }
            if (!j->rc && n && dvdagpu_fetch(w->ctx, j->res.pcm_offset, n, j->pcm)) {
                j->rc = -1;
                fprintf(stderr, "libdvd-audio (B200): %s\n", dvdagpu_last_error());
            }
        }
    }
    pthread_mutex_lock(&j->lock);
    j->done = 1;
    pthread_cond_signal(&j->cond);
    pthread_mutex_unlock(&j->lock);
}

static void *worker_main(void *arg)
{
    struct worker *w = arg;
    for (;;) {
        pthread_mutex_lock(&w->lock);
        while (!w->head && !w->quit) pthread_cond_wait(&w->wake, &w->lock);
        struct part_job *j = w->head;
        if (j) { w->head = j->next; if (!w->head) w->tail = NULL; }
        const int quit = w->quit && !j;
        pthread_mutex_unlock(&w->lock);
        if (quit) break;
        if (j) run_job(w, j);
    }
    return NULL;
}

static void pool_shutdown(void)
{
    pthread_mutex_lock(&g_pool_lock);
    for (unsigned i = 0; i < g_nworkers; i++) {
        struct worker *w = &g_workers[i];
        pthread_mutex_lock(&w->lock);
        w->quit = 1;
        pthread_cond_signal(&w->wake);
        pthread_mutex_unlock(&w->lock);
        pthread_join(w->thread, NULL);
        dvdagpu_destroy(w->ctx);
    }
    free(g_workers);
    g_workers = NULL;
    g_nworkers = 0;
    pthread_mutex_unlock(&g_pool_lock);
}

/* the workers, created on first use; 0 if there is no usable engine (no CPU path) */
static unsigned pool_workers(void)
{
    pthread_mutex_lock(&g_pool_lock);
    if (!g_workers && !g_pool_failed) {
        int devices[64], nd = 0;
        const int present = dvdagpu_device_count();
        const char *list = getenv("DVDA_B200_DEVICES");
        if (list && !strcmp(list, "all")) {
            for (int d = 0; d < present && nd < 64; d++) devices[nd++] = d;
        } else if (list && *list) {
            for (const char *p = list; *p && nd < 64;) {
                char *end;
                const long d = strtol(p, &end, 10);
                if (end == p) break;
                devices[nd++] = (int)d;
                p = *end == ',' ? end + 1 : end;
            }
        } else {
            const char *one = getenv("DVDA_B200_DEVICE");
            devices[nd++] = one ? atoi(one) : 0;
        }
        const char *pc = getenv("DVDA_B200_CONTEXTS");
        int per = pc ? atoi(pc) : 2;
        if (per < 1) per = 1;
        if (per > 4) per = 4;
        g_workers = calloc((size_t)nd * (size_t)per, sizeof *g_workers);
        /* contexts of one device are not neighbours in the list: consecutive parts go to different devices */
        for (int k = 0; k < per && g_workers; k++) {
            for (int i = 0; i < nd; i++) {
                struct worker *w = &g_workers[g_nworkers];
                w->ctx = dvdagpu_create(devices[i]);
                if (!w->ctx) {
                    if (k == 0) fprintf(stderr, "libdvd-audio (B200): %s\n", dvdagpu_last_error());
                    continue;
                }
                w->device = devices[i];
                pthread_mutex_init(&w->lock, NULL);
                pthread_cond_init(&w->wake, NULL);
                if (pthread_create(&w->thread, NULL, worker_main, w)) { dvdagpu_destroy(w->ctx); continue; }
                g_nworkers++;
            }
        }
        if (!g_nworkers) { free(g_workers); g_workers = NULL; g_pool_failed = 1; }
        else atexit(pool_shutdown);
    }
    const unsigned n = g_nworkers;
    pthread_mutex_unlock(&g_pool_lock);
    return n;
}

static void pool_submit(struct part_job *j)
{
    pthread_mutex_lock(&g_pool_lock);
    struct worker *w = &g_workers[g_next_worker++ % g_nworkers];
    pthread_mutex_unlock(&g_pool_lock);
    j->done = 0;
    j->next = NULL;
    pthread_mutex_lock(&w->lock);
    if (w->tail) w->tail->next = j; else w->head = j;
    w->tail = j;
    pthread_cond_signal(&w->wake);
    pthread_mutex_unlock(&w->lock);
}

static void job_wait(struct part_job *j)
{
    pthread_mutex_lock(&j->lock);
    while (!j->done) pthread_cond_wait(&j->cond, &j->lock);
    pthread_mutex_unlock(&j->lock);
}

/* ------------------------------------------------------------ track reader */

#define MAX_AHEAD 6
#define PART_MARGIN 64          /* sectors behind a part: the run to the next major sync */

struct DVDA_Track_Reader_s {
    dvda_codec_t codec;
    unsigned bits, rate, channels, assignment;
    struct aob_set aobs;
    unsigned long long first, last;             /* the track's sectors */
    unsigned pts_length;
    unsigned part_sectors;
    /* part i lives in ring[i % slots] from its submission until the part behind it is taken */
    struct part_job ring[MAX_AHEAD];
    unsigned slots;
    unsigned long long next_sector;             /* where the next part begins */
    unsigned long long parts_submitted, parts_taken, parts_released;
    int no_more;                                /* no further part will be submitted */
    /* the part being read */
    struct part_job *cur;
    unsigned long long cur_frames, cur_pos;     /* frames in it, frames already handed out */
    unsigned long long cur_skip;                /* samples at its front that belong to the part before (see take_part) */
    unsigned long long pcm_budget;              /* PCM: frames still to come */
    unsigned long long prev_first, prev_frames; /* MLP: the previous part, in case this one needs its filter history */
    unsigned prev_flags;
    int ended;
};

static void reader_free(DVDA_Track_Reader *r)
{
    /* nothing may be in flight when the buffers go */
    for (unsigned long long i = r->parts_taken; i < r->parts_submitted; i++) job_wait(&r->ring[i % r->slots]);
    for (unsigned i = 0; i < MAX_AHEAD; i++) {
        if (r->ring[i].sec) dvdagpu_host_free(r->ring[i].sec);
        if (r->ring[i].pcm) dvdagpu_host_free(r->ring[i].pcm);
        pthread_mutex_destroy(&r->ring[i].lock);
        pthread_cond_destroy(&r->ring[i].cond);
    }
    aobs_close(&r->aobs);
    free(r);
}

/* fills slot `j` with the window [first, first + sectors + margin) and hands it to a worker */
static int submit_window(DVDA_Track_Reader *r, struct part_job *j, unsigned long long first, unsigned long long sectors,
                         unsigned long long margin, unsigned pts, unsigned flags)
{
    unsigned long long n = sectors + margin;
    if (first + n > r->aobs.total) n = r->aobs.total - first;
    const size_t bytes = round_up((size_t)n * SECTOR_SIZE, 1 << 16);
    if (bytes > j->sec_cap) {
        dvdagpu_host_free(j->sec);
        j->sec = dvdagpu_host_alloc(bytes);
        j->sec_cap = j->sec ? bytes : 0;
        if (!j->sec) { fprintf(stderr, "libdvd-audio (B200): %s\n", dvdagpu_last_error()); return -1; }
    }
    j->aobs = &r->aobs;
    j->first = first;
    j->n_sectors = n;
    j->desc.first_sector = 0;
    j->desc.last_sector = (uint32_t)(sectors ? sectors - 1 : 0);
    j->desc.pts_length = pts;
    j->desc.flags = flags;
    pool_submit(j);
    return 0;
}

/* submits further parts of the track while there are any and slots are free */
static void submit_next(DVDA_Track_Reader *r)
{
    while (!r->no_more && r->parts_submitted - r->parts_released < r->slots) {
        struct part_job *j = &r->ring[r->parts_submitted % r->slots];
        unsigned long long sectors, margin;
        unsigned flags, pts = r->pts_length;
        if (r->parts_submitted > 0 && r->codec == DVDA_PCM) {
            /* PCM parts follow one another (each needs the frames still to come) and run on behind
               the track's last sector while the budget lasts, like the reference (dvd-audio.c:1016-1082) */
            if (r->parts_submitted > r->parts_taken) break;
            if (!r->pcm_budget || r->next_sector >= r->aobs.total) { r->no_more = 1; break; }
            sectors = r->aobs.total - r->next_sector;
            if (sectors > r->part_sectors) sectors = r->part_sectors;
            margin = 0;
            flags = DVDAGPU_PCM_BUDGET_IN_FRAMES;
            pts = r->pcm_budget > 0xFFFFFFFFull ? 0xFFFFFFFFu : (unsigned)r->pcm_budget;
        } else {
            if (r->next_sector > r->last || r->next_sector >= r->aobs.total) { r->no_more = 1; break; }
            sectors = r->last - r->next_sector + 1;
            const int final = sectors <= r->part_sectors + r->part_sectors / 4;     /* (a short rest goes with the last part) */
            if (!final) sectors = r->part_sectors;
            margin = PART_MARGIN;
            flags = (r->parts_submitted ? DVDAGPU_PART_CONTINUES_PREVIOUS : 0u) | (final ? 0u : DVDAGPU_PART_CONTINUED_BY_NEXT);
            if (final) r->no_more = 1;
        }
        if (submit_window(r, j, r->next_sector, sectors, margin, pts, flags)) { r->no_more = 1; break; }
        r->parts_submitted++;
        r->next_sector += sectors;
    }
}

/* makes the next finished part the current one; 0 when the track is over */
static int take_part(DVDA_Track_Reader *r)
{
    for (;;) {
        r->cur = NULL;
        r->parts_released = r->parts_taken;         /* the part that was being read is done with */
        if (r->ended) return 0;
        submit_next(r);
        if (r->parts_taken == r->parts_submitted) { r->ended = 1; return 0; }
        const unsigned long long index = r->parts_taken;
        struct part_job *j = &r->ring[index % r->slots];
        job_wait(j);
        r->parts_taken++;
        if (getenv("DVDA_B200_DEBUG"))
            fprintf(stderr, "[dvda] part %llu: sectors [%llu, +%llu) last %u flags %u -> rc %d status %d codec %d frames %llu stopped %u truncated %u err %x\n",
                    index, j->first, j->got, j->desc.last_sector, j->desc.flags, j->rc, j->res.status, j->res.codec,
                    (unsigned long long)j->res.frames, j->res.stopped, j->res.truncated, (unsigned)j->res.error_flags);
        if (j->rc || j->res.status != DVDAGPU_TRACK_OK) { r->ended = 1; return 0; }
        unsigned long long skip = 0;
        const int is_pcm = j->res.codec == 0;
        const int other_params = index > 0 && (j->res.channels != r->channels || j->res.sample_rate != r->rate ||
                                               j->res.bits_per_sample != r->bits || j->res.channel_assignment != r->assignment ||
                                               (is_pcm ? DVDA_PCM : DVDA_MLP) != r->codec);
        if (!is_pcm && index > 0 && r->codec == DVDA_MLP && (j->res.stopped == 2 || other_params)) {
            /* The part cannot stand alone: it needs the decoder state of the one before it (FIR
               history survives a restart header, mlp.c:948-952; a major sync without restart header),
               or it begins at a major sync that states other stream parameters than the track's (such
               an access unit is dropped, mlp.c:449-455 — but only the track's first sync says what
               the track's parameters are).  Both parts are decoded as one window, the samples of the
               part before are dropped. */
            const unsigned long long sectors = j->first + j->desc.last_sector + 1 - r->prev_first;
            const unsigned flags = (r->prev_flags & DVDAGPU_PART_CONTINUES_PREVIOUS) | (j->desc.flags & DVDAGPU_PART_CONTINUED_BY_NEXT);
            if (submit_window(r, j, r->prev_first, sectors, PART_MARGIN, r->pts_length, flags)) { r->ended = 1; return 0; }
            job_wait(j);
            if (j->rc || j->res.status != DVDAGPU_TRACK_OK || j->res.stopped == 2 || j->res.frames < r->prev_frames ||
                j->res.channels != r->channels || j->res.sample_rate != r->rate || j->res.channel_assignment != r->assignment) { r->ended = 1; return 0; }
            skip = r->prev_frames;
            /* (the parts already submitted behind it were cut at the same places: they stay valid) */
        } else if (!is_pcm && j->res.truncated && !(j->desc.flags & DVDAGPU_PART_CONTINUED_BY_NEXT) &&
                   j->first + j->n_sectors < r->aobs.total && j->n_sectors < (1ull << 22)) {
            /* the last part ran out of sectors before the next major sync: once more with a wider margin */
            const unsigned long long sectors = j->desc.last_sector + 1ull;
            if (submit_window(r, j, j->first, sectors, (j->n_sectors - sectors) * 8 + PART_MARGIN, r->pts_length, j->desc.flags)) { r->ended = 1; return 0; }
            r->parts_taken--;           /* (the same part again) */
            continue;
        }
        if (index == 0) {
            r->codec = is_pcm ? DVDA_PCM : DVDA_MLP;
            r->bits = j->res.bits_per_sample; r->rate = j->res.sample_rate;
            r->channels = j->res.channels; r->assignment = j->res.channel_assignment;
            if (is_pcm) {
                const double total = (double)r->pts_length * (double)r->rate / (double)PTS_PER_SECOND;
                r->pcm_budget = (unsigned long long)(total + 0.5);
                /* the window was the part and its margin: all of it was delivered */
                r->next_sector = j->first + j->got;
                r->no_more = 0;
            }
        } else if (skip == 0 && other_params) {
            r->ended = 1;               /* the stream changed its parameters: the track is over (dvd-audio.c:1049-1055) */
            return 0;
        }
        if (j->res.error_flags & DVDAGPU_ERR_PARITY) fprintf(stderr, "parity mismatch\n");
        if (j->res.error_flags & DVDAGPU_ERR_CRC) fprintf(stderr, "CRC-8 mismatch\n");
        const unsigned long long frames = j->res.frames - skip;
        if (is_pcm) {
            r->pcm_budget = frames >= r->pcm_budget ? 0 : r->pcm_budget - frames;
            /* over: the budget is used up, or the window ended early (a packet that is not PCM, changed parameters) */
            if (!r->pcm_budget || !j->res.truncated) r->no_more = 1;
            else if (index > 0) r->next_sector = j->first + j->got;
        } else {
            r->prev_first = j->first; r->prev_frames = j->res.frames; r->prev_flags = j->desc.flags;
            if (j->res.stopped == 1) { r->no_more = 1; r->ended = 1; }      /* the track ended inside this part: nothing behind it counts */
        }
        r->cur = j;
        r->cur_frames = frames;
        r->cur_skip = skip * r->channels;
        r->cur_pos = 0;
        if (!r->ended) submit_next(r);  /* keep the workers busy while this part is read */
        if (!frames) { if (r->ended) return 0; continue; }      /* (a part without a major sync is empty) */
        return 1;
    }
}

DVDA_Track_Reader *dvda_open_track_reader(const DVDA_Track *track)
{
    DVDA_Track_Reader *r = calloc(1, sizeof *r);
    if (!r) return NULL;
    for (unsigned i = 0; i < MAX_AHEAD; i++) {
        pthread_mutex_init(&r->ring[i].lock, NULL);
        pthread_cond_init(&r->ring[i].cond, NULL);
    }
    aobs_open(&r->aobs, track->disc.audio_ts, track->titleset);
    r->first = track->t.first_sector;
    r->last = track->t.last_sector < r->first ? r->first : track->t.last_sector;
    r->pts_length = track->t.pts_length;
    const unsigned workers = (r->aobs.count && r->first < r->aobs.total) ? pool_workers() : 0;
    if (!workers) { reader_free(r); return NULL; }
    const char *ps = getenv("DVDA_B200_PART_SECTORS");
    r->part_sectors = ps ? (unsigned)atoi(ps) : 8192u;
    if (r->part_sectors < 16) r->part_sectors = 16;
    r->next_sector = r->first;
    /* the first part tells the format (and whether the track can be opened at all): it goes alone,
       the parts behind it are cut according to the codec */
    r->slots = 1;
    const int ok = take_part(r);
    if (!ok && !r->channels) { reader_free(r); return NULL; }
    r->slots = workers + 1 > MAX_AHEAD ? MAX_AHEAD : workers + 1;
    if (!r->ended) submit_next(r);
    return r;
}

void dvda_close_track_reader(DVDA_Track_Reader *r)
{
    if (r) reader_free(r);
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
    /* as much as was asked for, unless the track ends first (behaviour of src/dvd-audio.c:751-795);
       the parts behind the one being read are decoded meanwhile */
    unsigned done = 0;
    while (done < pcm_frames) {
        if (!r->cur || r->cur_pos == r->cur_frames) {
            if (!take_part(r)) break;
        }
        unsigned long long n = r->cur_frames - r->cur_pos;
        if (n > pcm_frames - done) n = pcm_frames - done;
        memcpy(buffer + (size_t)done * r->channels, r->cur->pcm + r->cur_skip + r->cur_pos * r->channels,
               (size_t)n * r->channels * sizeof(int));
        r->cur_pos += n;
        done += (unsigned)n;
    }
    return done;
}
