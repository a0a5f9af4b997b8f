/* dvda2wav — extract every track of a DVD-Audio title set to .wav files through
 * the public dvd-audio.h API (GPU engine behind it).
 *
 * Equivalent of the reference's tool (utils/dvda2wav.c): same options, same file
 * names (track-TT-KK.wav), same WAVE_FORMAT_EXTENSIBLE header (68 bytes: format
 * tag 0xFFFE, cbSize 22, PCM sub-format GUID, channel mask), same progress lines.
 * Samples are written as little-endian two's complement; the reference writes
 * sign bit + low bits (src/bitstream.c:2831-2843), which is the same thing for
 * every sample that fits its bits-per-sample — i.e. for every real disc.
 *
 * usage: dvda2wav -A AUDIO_TS [-c cdrom] [-T title] [-t track] [-d output dir]
 */
#define _POSIX_C_SOURCE 200809L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <getopt.h>
#include "dvd-audio.h"

#define FRAMES_PER_READ 4096

static void put_le(FILE *f, unsigned bytes, uint32_t v)
{
    for (unsigned i = 0; i < bytes; i++) fputc((int)((v >> (8 * i)) & 0xFF), f);
}

static void wave_header(FILE *f, unsigned rate, unsigned channels, unsigned mask, unsigned bits, uint32_t frames)
{
    static const uint8_t pcm_guid[16] = {1, 0, 0, 0, 0, 0, 16, 0, 128, 0, 0, 170, 0, 56, 155, 113};
    const unsigned bytes = bits / 8;
    const uint32_t data = bytes * channels * frames;
    fwrite("RIFF", 1, 4, f);
    put_le(f, 4, 4 + (8 + 40) + 8 + data + (data & 1));
    fwrite("WAVE", 1, 4, f);
    fwrite("fmt ", 1, 4, f);
    put_le(f, 4, 40);
    put_le(f, 2, 0xFFFE);
    put_le(f, 2, channels);
    put_le(f, 4, rate);
    put_le(f, 4, rate * channels * bytes);
    put_le(f, 2, channels * bytes);
    put_le(f, 2, bits);
    put_le(f, 2, 22);
    put_le(f, 2, bits);
    put_le(f, 4, mask);
    fwrite(pcm_guid, 1, 16, f);
    fwrite("data", 1, 4, f);
    put_le(f, 4, data);
}

static int extract_track(DVDA_Title *title, unsigned track_num, const char *dir)
{
    DVDA_Track *track = dvda_open_track(title, track_num);
    if (!track) {
        fprintf(stderr, "*** Error: unable to open track %u\n", track_num);
        return 1;
    }
    DVDA_Track_Reader *r = dvda_open_track_reader(track);
    if (!r) {
        fprintf(stderr, "*** Error: unable to open track %u for reading\n", track_num);
        dvda_close_track(track);
        return 1;
    }
    char path[4096];
    snprintf(path, sizeof path, "%s/track-%2.2u-%2.2u.wav", dir, dvda_title_number(title), dvda_track_number(track));
    dvda_close_track(track);

    FILE *f = fopen(path, "wb");
    if (!f) {
        fprintf(stderr, "*** Error: unable to open \"%s\" for writing\n", path);
        dvda_close_track_reader(r);
        return 1;
    }
    const unsigned ch = dvda_channel_count(r), bits = dvda_bits_per_sample(r), bytes = bits / 8;
    printf("* Extracting %s track  %u channels  %u Hz  %u bps\n",
           dvda_codec(r) == DVDA_MLP ? "MLP" : "PCM", ch, dvda_sample_rate(r), bits);
    wave_header(f, dvda_sample_rate(r), ch, dvda_riff_wave_channel_mask(r), bits, 0);

    int *buf = malloc(sizeof(int) * FRAMES_PER_READ * (ch ? ch : 1));
    uint8_t *out = malloc((size_t)FRAMES_PER_READ * (ch ? ch : 1) * 4);
    uint32_t total = 0;
    unsigned got;
    while ((got = dvda_read(r, FRAMES_PER_READ, buf)) > 0) {
        size_t o = 0;
        for (unsigned i = 0; i < got * ch; i++) {
            const uint32_t v = (uint32_t)buf[i];
            for (unsigned b = 0; b < bytes; b++) out[o++] = (uint8_t)(v >> (8 * b));
        }
        fwrite(out, 1, o, f);
        total += got;
    }
    rewind(f);
    wave_header(f, dvda_sample_rate(r), ch, dvda_riff_wave_channel_mask(r), bits, total);
    fclose(f);
    free(buf);
    free(out);
    dvda_close_track_reader(r);
    printf("* Wrote: \"%s\"\n", path);
    return 0;
}

int main(int argc, char **argv)
{
    const char *audio_ts = NULL, *cdrom = NULL, *dir = ".";
    unsigned title_sel = 0, track_sel = 0;
    static const struct option opts[] = {
        {"audio_ts", required_argument, NULL, 'A'}, {"cdrom", required_argument, NULL, 'c'},
        {"title", required_argument, NULL, 'T'}, {"track", required_argument, NULL, 't'},
        {"dir", required_argument, NULL, 'd'}, {"help", no_argument, NULL, 'h'}, {NULL, 0, NULL, 0}};
    int c;
    while ((c = getopt_long(argc, argv, "A:c:T:t:d:h", opts, NULL)) != -1) {
        switch (c) {
        case 'A': audio_ts = optarg; break;
        case 'c': cdrom = optarg; break;
        case 'T': title_sel = (unsigned)strtoul(optarg, NULL, 10); break;
        case 't': track_sel = (unsigned)strtoul(optarg, NULL, 10); break;
        case 'd': dir = optarg; break;
        default:
            printf("*** Usage: dvda2wav -A [AUDIO_TS] -c [cdrom] -T [title] -t [track] -d [output dir]\n");
            return c == 'h' ? 0 : 1;
        }
    }
    if (!audio_ts) {
        fprintf(stderr, "*** Error: AUDIO_TS directory is required\n");
        return 1;
    }
    DVDA *dvda = dvda_open(audio_ts, cdrom);
    if (!dvda) {
        fprintf(stderr, "*** Error: unable to open DVD-A\n");
        return 1;
    }
    /* like the reference's tool: title set 1 only (utils/dvda2wav.c:83) */
    DVDA_Titleset *ts = dvda_open_titleset(dvda, 1);
    if (!ts) {
        fprintf(stderr, "*** Error: unable to open titleset 1\n");
        dvda_close(dvda);
        return 1;
    }
    int rc = 0;
    for (unsigned t = 1; t <= dvda_title_count(ts); t++) {
        if (title_sel && t != title_sel) continue;
        DVDA_Title *title = dvda_open_title(ts, t);
        if (!title) { fprintf(stderr, "*** Error: unable to open title %u\n", t); rc = 1; continue; }
        for (unsigned k = 1; k <= dvda_track_count(title); k++) {
            if (track_sel && k != track_sel) continue;
            rc |= extract_track(title, k, dir);
        }
        dvda_close_title(title);
    }
    dvda_close_titleset(ts);
    dvda_close(dvda);
    return rc;
}
