/* dvd-audio.h — public C API of the B200-native DVD-Audio decoding engine.
 *
 * This is the drop-in boundary: the declarations below are call-compatible,
 * symbol for symbol, with the reference library's public header
 * (reference include/dvd-audio.h:36-201), so a program written against
 * libdvd-audio (e.g. its dvda2wav tool) relinks against this library
 * unchanged.  Everything behind dvda_open_track_reader()/dvda_read() is
 * different: the AOB sectors of a track are handed to the CUDA shim
 * (include/dvdagpu.h) and decoded on an sm_100a device; there is no CPU
 * decode path in this library.
 *
 * Object model (all numbers are 1-based):
 *
 *   DVDA  --dvda_open_titleset-->  DVDA_Titleset  --dvda_open_title-->
 *   DVDA_Title  --dvda_open_track-->  DVDA_Track  --dvda_open_track_reader-->
 *   DVDA_Track_Reader  --dvda_read-->  interleaved int PCM
 *
 * Every object returned by an "open" call is owned by the caller and released
 * with the matching "close".  A child keeps no pointer into its parent, so a
 * parent may be closed while its children are still in use.  An "open" call
 * returns NULL on failure.  Distinct readers may be driven from distinct
 * threads; one reader is not re-entrant.
 */
#ifndef DVD_AUDIO_B200_PUBLIC_H
#define DVD_AUDIO_B200_PUBLIC_H

#include <inttypes.h>

#ifdef __cplusplus
extern "C" {
#endif

/* library version — kept equal to the reference release this API mirrors */
#define LIBDVDAUDIO_MAJOR_VERSION 1
#define LIBDVDAUDIO_MINOR_VERSION 0
#define LIBDVDAUDIO_RELEASE_VERSION 1

#define DVDA_B200_STR_(x) #x
#define DVDA_B200_STR(x) DVDA_B200_STR_(x)
#define LIBDVDAUDIO_VERSION_STRING                 \
    DVDA_B200_STR(LIBDVDAUDIO_MAJOR_VERSION) "."   \
    DVDA_B200_STR(LIBDVDAUDIO_MINOR_VERSION) "."   \
    DVDA_B200_STR(LIBDVDAUDIO_RELEASE_VERSION)

/* presentation-time-stamp ticks per second used by every *_pts_* accessor */
#define PTS_PER_SECOND 90000

/* opaque handles */
typedef struct DVDA_s DVDA;
typedef struct DVDA_Titleset_s DVDA_Titleset;
typedef struct DVDA_Title_s DVDA_Title;
typedef struct DVDA_Track_s DVDA_Track;
typedef struct DVDA_Index_s DVDA_Index;
typedef struct DVDA_Track_Reader_s DVDA_Track_Reader;

/* audio coding of a track */
typedef enum {DVDA_PCM, DVDA_MLP} dvda_codec_t;

/* ---- disc ------------------------------------------------------------- */

/* audio_ts_path: the disc's AUDIO_TS directory.  device: optional drive
 * node (may be NULL; CPPM decryption is out of scope and the argument is
 * only remembered).  NULL if AUDIO_TS.IFO is missing or not a DVD-Audio
 * manager file. */
DVDA *dvda_open(const char *audio_ts_path, const char *device);
void dvda_close(DVDA *dvda);
unsigned dvda_titleset_count(const DVDA *dvda);

/* ---- title set ---------------------------------------------------------- */

/* NULL if ATS_<titleset>_0.IFO is missing or fails to parse */
DVDA_Titleset *dvda_open_titleset(DVDA *dvda, unsigned titleset);
void dvda_close_titleset(DVDA_Titleset *titleset);
unsigned dvda_titleset_number(const DVDA_Titleset *titleset);
unsigned dvda_title_count(const DVDA_Titleset *titleset);

/* ---- title -------------------------------------------------------------- */

/* NULL if title is 0 or beyond dvda_title_count() */
DVDA_Title *dvda_open_title(DVDA_Titleset *titleset, unsigned title);
void dvda_close_title(DVDA_Title *title);
unsigned dvda_title_number(const DVDA_Title *title);
unsigned dvda_track_count(const DVDA_Title *title);
/* whole-title length in PTS ticks */
unsigned dvda_title_pts_length(const DVDA_Title *title);

/* ---- track -------------------------------------------------------------- */

/* NULL if track is 0 or beyond dvda_track_count() */
DVDA_Track *dvda_open_track(DVDA_Title *title, unsigned track);
void dvda_close_track(DVDA_Track *track);
unsigned dvda_track_number(const DVDA_Track *track);
/* start / length of the track in PTS ticks */
unsigned dvda_track_pts_index(const DVDA_Track *track);
unsigned dvda_track_pts_length(const DVDA_Track *track);
/* sector range inside the title set's concatenated AOB files; audio may
 * begin after the start of the first and end before the end of the last */
unsigned dvda_track_first_sector(const DVDA_Track *track);
unsigned dvda_track_last_sector(const DVDA_Track *track);

/* ---- track reader ------------------------------------------------------- */

/* Opens the track's audio for reading.  In this engine the call uploads the
 * track's sectors to the GPU and decodes the whole track there; dvda_read()
 * then serves slices of the result.  NULL on any error (no AOB data, no
 * audio packet, unknown codec, no usable CUDA device). */
DVDA_Track_Reader *dvda_open_track_reader(const DVDA_Track *track);
void dvda_close_track_reader(DVDA_Track_Reader *reader);

dvda_codec_t dvda_codec(const DVDA_Track_Reader *reader);
/* 16 or 24 */
unsigned dvda_bits_per_sample(const DVDA_Track_Reader *reader);
/* in Hz */
unsigned dvda_sample_rate(const DVDA_Track_Reader *reader);
unsigned dvda_channel_count(const DVDA_Track_Reader *reader);
/* RIFF WAVE dwChannelMask for the track's channel assignment */
unsigned dvda_riff_wave_channel_mask(const DVDA_Track_Reader *reader);

/* Fills buffer (room for dvda_channel_count() * pcm_frames ints) with up to
 * pcm_frames frames, samples interleaved frame by frame in RIFF WAVE channel
 * order.  Returns the number of frames delivered: fewer than asked only at
 * the end of the track, 0 once the track is exhausted (or pcm_frames is 0). */
unsigned dvda_read(DVDA_Track_Reader *reader,
                   unsigned pcm_frames,
                   int buffer[]);

#ifdef __cplusplus
}
#endif

#endif /* DVD_AUDIO_B200_PUBLIC_H */
