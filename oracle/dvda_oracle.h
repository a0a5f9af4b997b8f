/* dvda_oracle.h — CPU restatement of the reference decode path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under libdvd-audio_b200/ may include,
 * link or call this; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg use it, and only as the checker.
 *
 * Parity status: PINNED.  The restatement is checked bit-for-bit against the
 * unmodified reference (oracle/_ref, built by oracle/Makefile from
 * /root/reference) on generated discs in tests/test_oracle_vs_reference.py,
 * and against the committed golden hashes in tests/golden/ that were produced
 * by that reference build (tests/golden/make_golden.py).  The reference's own
 * test-suite holds no vectors for this path; the bit-reader
 * known answers it does hold (src/bitstream.c:4864-4868, 4940-4944) are
 * checked in tests/test_oracle_units.py.
 */
#ifndef DVDA_ORACLE_H
#define DVDA_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    DVDA_ORACLE_OK = 0,
    DVDA_ORACLE_NO_AUDIO = 1,       /* no audio packet / unknown codec / no sync: open fails */
    DVDA_ORACLE_ERR_PARITY = 1 << 4,
    DVDA_ORACLE_ERR_CRC = 1 << 5,
    DVDA_ORACLE_ERR_SYNTAX = 1 << 6 /* malformed access unit: track ends before it */
};

typedef struct {
    int status;                 /* DVDA_ORACLE_OK or DVDA_ORACLE_NO_AUDIO */
    int error_flags;            /* DVDA_ORACLE_ERR_* met while decoding */
    int codec;                  /* 0 = PCM, 1 = MLP */
    unsigned group_0_bps, group_1_bps, group_0_rate, group_1_rate, channel_assignment;
    unsigned channels, bits_per_sample, sample_rate;
    uint64_t frames;
    int32_t *pcm;               /* frames * channels, interleaved, RIFF WAVE order */
    /* introspection for kernel debugging */
    uint64_t access_units;      /* MLP access units decoded */
    uint64_t es_bytes;          /* MLP elementary-stream bytes consumed by them */
} dvda_oracle_result;

/* Decodes one track the way dvda_open_track_reader() + dvda_read()-until-0 do
 * (reference src/dvd-audio.c:597-795).  `sectors` is the title set's AOB data
 * from global sector 0 (the concatenation of ATS_tt_1..9.AOB), n_sectors long.
 * Returns 0 and fills *out (release with dvda_oracle_free) or nonzero if the
 * reference would have returned NULL from dvda_open_track_reader. */
int dvda_oracle_decode_track(const uint8_t *sectors, uint64_t n_sectors,
                             uint32_t first_sector, uint32_t last_sector,
                             uint32_t pts_length, dvda_oracle_result *out);

void dvda_oracle_free(dvda_oracle_result *r);

/* unit-level entry points used by tests */
uint32_t dvda_oracle_read_bits(const uint8_t *buf, size_t len, size_t bitpos, unsigned n);
int32_t dvda_oracle_read_signed(const uint8_t *buf, size_t len, size_t bitpos, unsigned n);
/* decodes one Huffman symbol of codebook cb (1..3) at bitpos; returns the
 * value (-1 = invalid code) and stores the code length */
int dvda_oracle_huffman(const uint8_t *buf, size_t len, size_t bitpos, int cb, unsigned *code_len);
uint8_t dvda_oracle_crc8_table(unsigned i);
/* fills table[chunk_size] with the PCM byte permutation for (bits 16|24, channels) */
void dvda_oracle_pcm_permutation(int bits, int channels, uint8_t *table);

#ifdef __cplusplus
}
#endif
#endif
