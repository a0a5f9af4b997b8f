/* api_dump — decode tracks through the PUBLIC dvd-audio.h API and dump the raw
 * ints dvda_read() returns.  TEST / BENCH INFRASTRUCTURE.
 *
 * The program touches nothing but the public header, so one source serves two
 * binaries: oracle/_ref/ref_dump (linked against the unmodified reference
 * library — the parity oracle and the CPU baseline) and build/b200_dump (linked
 * against this repo's GPU-backed libdvd-audio) — which is also the drop-in
 * demonstration.  We dump at dvda_read() level instead of comparing .wav files
 * because the reference's dvda2wav writes sign + low bits rather than two's
 * complement (reference src/bitstream.c:2831-2843), mangling out-of-range
 * samples.
 *
 * usage: api_dump AUDIO_TS [-s titleset] [-T title] [-t track] [-c frames]
 *                 [-o out.raw] [-n] [-r repeat]
 *   -T/-t 0 (default) = all titles / all tracks
 *   -c   frames per dvda_read() call (default 4096, what dvda2wav uses,
 *        reference utils/dvda2wav.c:326)
 *   -o   append every selected track's interleaved int32 samples to this file
 *   -n   decode and discard (timing mode)
 *   -r   decode the selection this many times (timing mode)
 * One line per decoded track goes to stdout:
 *   track <title> <track> codec=<PCM|MLP> bps=.. rate=.. ch=.. mask=.. frames=..
 *         first=.. last=.. pts=.. fnv=<64-bit FNV-1a of the int32 LE samples>
 * and a final "elapsed <seconds> samples <n>" line.
 */
#define _POSIX_C_SOURCE 200809L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <time.h>
#include "dvd-audio.h"

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static uint64_t fnv1a(uint64_t h, const int *v, size_t n)
{
    for (size_t i = 0; i < n; i++) {
        uint32_t x = (uint32_t)v[i];
        for (int b = 0; b < 4; b++) {
            h ^= (x >> (8 * b)) & 0xFF;
            h *= 0x100000001B3ULL;
        }
    }
    return h;
}

static int dump_track(DVDA_Title *title, unsigned track_num, unsigned chunk,
                      FILE *out, int quiet, uint64_t *total_samples)
{
    DVDA_Track *track = dvda_open_track(title, track_num);
    if (!track) {
        fprintf(stderr, "cannot open track %u\n", track_num);
        return 1;
    }
    const unsigned first = dvda_track_first_sector(track);
    const unsigned last = dvda_track_last_sector(track);
    const unsigned pts = dvda_track_pts_length(track);
    DVDA_Track_Reader *r = dvda_open_track_reader(track);
    dvda_close_track(track);   /* children outlive parents */
    if (!r) {
        fprintf(stderr, "cannot open reader for track %u\n", track_num);
        return 1;
    }
    const unsigned ch = dvda_channel_count(r);
    int *buf = malloc(sizeof(int) * (size_t)chunk * (ch ? ch : 1));
    uint64_t frames = 0;
    uint64_t h = 0xCBF29CE484222325ULL;
    unsigned got;
    while ((got = dvda_read(r, chunk, buf)) > 0) {
        if (!quiet)
            h = fnv1a(h, buf, (size_t)got * ch);
        if (out) {
            /* int is 32-bit little-endian on every platform we build for */
            fwrite(buf, sizeof(int), (size_t)got * ch, out);
        }
        frames += got;
    }
    if (!quiet) {
        printf("track %u %u codec=%s bps=%u rate=%u ch=%u mask=%u frames=%llu "
               "first=%u last=%u pts=%u fnv=%016llx\n",
               dvda_title_number(title), track_num,
               dvda_codec(r) == DVDA_MLP ? "MLP" : "PCM",
               dvda_bits_per_sample(r), dvda_sample_rate(r), ch,
               dvda_riff_wave_channel_mask(r),
               (unsigned long long)frames, first, last, pts,
               (unsigned long long)h);
    }
    *total_samples += frames * ch;
    free(buf);
    dvda_close_track_reader(r);
    return 0;
}

int main(int argc, char **argv)
{
    const char *path = NULL, *out_path = NULL;
    unsigned titleset = 1, title_sel = 0, track_sel = 0, chunk = 4096, repeat = 1;
    int quiet = 0;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "-s") && i + 1 < argc) titleset = (unsigned)atoi(argv[++i]);
        else if (!strcmp(argv[i], "-T") && i + 1 < argc) title_sel = (unsigned)atoi(argv[++i]);
        else if (!strcmp(argv[i], "-t") && i + 1 < argc) track_sel = (unsigned)atoi(argv[++i]);
        else if (!strcmp(argv[i], "-c") && i + 1 < argc) chunk = (unsigned)atoi(argv[++i]);
        else if (!strcmp(argv[i], "-o") && i + 1 < argc) out_path = argv[++i];
        else if (!strcmp(argv[i], "-r") && i + 1 < argc) repeat = (unsigned)atoi(argv[++i]);
        else if (!strcmp(argv[i], "-n")) quiet = 1;
        else if (argv[i][0] != '-') path = argv[i];
        else { fprintf(stderr, "bad argument %s\n", argv[i]); return 2; }
    }
    if (!path || !chunk) {
        fprintf(stderr, "usage: %s AUDIO_TS [-s ts] [-T title] [-t track] "
                "[-c frames] [-o out.raw] [-n] [-r repeat]\n", argv[0]);
        return 2;
    }

    DVDA *dvda = dvda_open(path, NULL);
    if (!dvda) { fprintf(stderr, "cannot open %s\n", path); return 1; }
    DVDA_Titleset *ts = dvda_open_titleset(dvda, titleset);
    if (!ts) { fprintf(stderr, "cannot open titleset %u\n", titleset); dvda_close(dvda); return 1; }

    FILE *out = NULL;
    if (out_path && !(out = fopen(out_path, "wb"))) {
        fprintf(stderr, "cannot write %s\n", out_path);
        return 1;
    }

    int rc = 0;
    uint64_t total_samples = 0;
    const double t0 = now_s();
    for (unsigned rep = 0; rep < repeat; rep++) {
        for (unsigned t = 1; t <= dvda_title_count(ts); t++) {
            if (title_sel && t != title_sel) continue;
            DVDA_Title *title = dvda_open_title(ts, t);
            if (!title) { rc = 1; continue; }
            for (unsigned k = 1; k <= dvda_track_count(title); k++) {
                if (track_sel && k != track_sel) continue;
                rc |= dump_track(title, k, chunk, out, quiet, &total_samples);
            }
            dvda_close_title(title);
        }
    }
    const double t1 = now_s();
    printf("elapsed %.6f samples %llu\n", t1 - t0, (unsigned long long)total_samples);

    if (out) fclose(out);
    dvda_close_titleset(ts);
    dvda_close(dvda);
    return rc;
}
