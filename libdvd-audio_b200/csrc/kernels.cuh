// kernels.cuh — tables shared between the kernel files and the host-side
// launchers the engine (engine.cu) calls.
#pragma once
#include "common.cuh"

// one row per audio (0xBD) packet, in sector order
struct PacketTable {
    uint32_t *sector;      // sector index inside the uploaded buffer
    uint16_t *off;         // offset in the sector of the byte behind pad_2_size
    uint16_t *len;         // bytes from there to the end of the packet
    uint8_t *codec;        // 0xA0 PCM, 0xA1 MLP, 0xFF anything else / malformed
    uint8_t *pad2;         // pad_2_size
    uint32_t *params;      // PCM: packed stream parameters of this packet
    uint32_t *mlp_len;     // MLP: elementary-stream bytes in this packet
    uint32_t *pcm_frames;  // PCM: whole frames in this packet
};

struct AuSnap;      // mlp_decode.cu: what pass B needs to entropy-decode one access unit
struct SegCtx;      // mlp_decode.cu: what a segment's restart header fixes for its parameter blocks
struct AuDelta;     // mlp_decode.cu: the parameters one access unit transmits

// everything the MLP kernels need to find their data
struct MlpTables {
    const uint8_t *es;             // elementary stream (padded with DVDA_ES_PAD zero bytes)
    uint64_t es_total;
    const uint64_t *pk_es;         // [np + 1] ES offset of every packet
    uint32_t np;
    TrackDev *tracks;
    uint32_t n_tracks;
    SegDev *segs;
    uint32_t nseg;
    GroupDev *groups;
    uint32_t ngroups;
    uint64_t *au_pos;              // [nau] ES offset of every access unit
    uint8_t *au_err;               // [nau] 0 ok, 1 drop (parameter change), else ERR_* bits
    AuDev *au;                     // [nau]
    ParamSet *psets;               // [nau]
    uint32_t *au_frames_ss;        // [2][nau] frames each substream decoded
    uint32_t nau;
    uint32_t *ss_flags;            // [2][nseg] SEG_* per substream
    uint32_t *ss_flags_prev;       // snapshot taken before the carry fix-up
    uint32_t *ss_flags_fast;       // snapshot taken after passes A and B of the fast path
    int32_t *fir_tail;             // [2][nseg][8 ch][8] last outputs per channel
    int32_t *tiles;
    uint8_t *bypass;
    int32_t *pcm;
    AuSnap *au_snap;               // [2][nau]
    uint32_t *au_seg;              // [nau]: segment of every access unit
    uint32_t nss_max;              // most substreams in a track of the batch
    uint8_t *au_fchg;              // [2][nau]: bit cc = the filter set-up of channel cc changes with this access unit
    SegCtx *seg_ctx;               // [2][nseg]
    AuDelta *au_delta;             // [2][nau], written where the AU brings parameters
    uint32_t fast;                 // 1, 2: the complete decoder only takes segments flagged SEG_FALLBACK
                                   // (1: three-pass path, 2: header passes + fused entropy/filter/output pass)
    uint32_t max_au;               // largest access-unit count of a segment
    const uint32_t *status;        // the batch's status word (SEG_OVERFLOW, STATUS_*)
    uint32_t *any_fallback;        // set by the flag kernels of the fast path when the complete decoder has work at all
    uint32_t *ss_sticky;           // [2][nseg] SEG_FALLBACK the fused pass asked for (it found out too late: the decode is repeated)
    uint32_t *status_rw;           // the batch's status word, for the kernels that set bits in it
    const uint16_t *huff_lut;      // [4][512] Huffman look-up table in device memory (built once per device)
};

// demux.cu
int upload_pcm_tables(const uint8_t *tables);
int launch_sector_count(const uint8_t *sectors, uint32_t n_sectors, uint32_t *sec_cnt, uint32_t *sec_bad, cudaStream_t s);
// rows: rows the packet table has room for (the packet count is still on the device: sec_base[n_sectors])
int launch_packet_fill(const uint8_t *sectors, uint32_t n_sectors, const uint32_t *sec_base, PacketTable pt, uint32_t rows,
                       uint32_t *nonmlp, uint32_t *pcm_stop, cudaStream_t s);
int launch_es_gather(const uint8_t *sectors, PacketTable pt, uint32_t np, const uint64_t *pk_es, uint8_t *es, cudaStream_t s);
int launch_pcm_unpack(const uint8_t *sectors, PacketTable pt, uint32_t np, const uint64_t *pk_pf,
                      const TrackDev *tracks, const uint32_t *trk_pk_lo, uint32_t n_tracks, int32_t *pcm, cudaStream_t s);

// mlp_index.cu
#define SYNC_CHUNK 512u           // ES bytes per warp step of the sync search
#define SYNC_SLOT_BYTES 4u        // per chunk: 2 slots of 16 bits
int launch_sync_count(const uint8_t *es, uint64_t es_total, uint32_t *cnt_raw, uint32_t *cnt_valid, uint16_t *slots, uint32_t nslots, cudaStream_t s);
int launch_sync_fill(const uint8_t *es, uint64_t es_total, const uint32_t *cnt_raw, const uint16_t *slots, uint32_t nslots,
                     const uint32_t *base_raw, const uint32_t *base_valid, uint64_t *raw, uint32_t cap_raw,
                     uint64_t *valid, uint32_t cap_valid, cudaStream_t s);
struct TrackSetupArgs {
    const uint8_t *es;
    uint64_t es_total;
    uint32_t n_sectors;
    const uint32_t *sec_base;      // [n_sectors + 1]
    const uint32_t *bad_prefix;    // [n_sectors + 1]
    PacketTable pt;
    uint32_t np;
    const uint64_t *pk_es;         // [np + 1]
    const uint64_t *pk_pf;         // [np + 1] PCM frames
    const uint32_t *pk_nonmlp;     // [np + 1]
    const uint32_t *pk_pcm_stop;   // [np + 1]
    // the sync lists: their lengths are still on the device when the kernel runs (the lists were
    // sized in advance: cap_*; entries beyond were not written and the host comes back)
    const uint64_t *raw;
    const uint32_t *n_raw;
    uint32_t cap_raw;
    const uint64_t *valid;
    const uint32_t *n_valid;
    uint32_t cap_valid;
};
int launch_track_setup(TrackSetupArgs a, TrackDev *tracks, uint32_t n_tracks, cudaStream_t s);
int launch_segment_fill(const TrackDev *tracks, uint32_t n_tracks, const uint32_t *trk_seg_base,
                        const uint64_t *valid, SegDev *segs, uint32_t nseg, cudaStream_t s);
// noted: scratch of au_noted_bytes(nseg) bytes shared by the two passes (count, then fill)
size_t au_noted_bytes(uint32_t nseg);
int launch_au_chase(const uint8_t *es, SegDev *segs, uint32_t nseg, const TrackDev *tracks,
                    uint32_t *seg_nau, uint64_t *au_pos, uint32_t *au_seg, const uint32_t *seg_au_base,
                    uint32_t *noted, int fill, cudaStream_t s);
int launch_yield(MlpTables m, const uint32_t *seg_au_base, PacketTable pt, const uint32_t *trk_pk_lo,
                 uint8_t *pk_yield, cudaStream_t s);
int launch_group_setup(const TrackDev *tracks, uint32_t n_tracks, const uint32_t *trk_grp_base, const SegDev *segs,
                       GroupDev *groups, uint32_t ngroups, uint32_t *grp_cells, uint32_t *grp_chunks, uint32_t *max_au, cudaStream_t s);
int launch_group_offsets(GroupDev *groups, uint32_t ngroups, const uint64_t *cell_base, cudaStream_t s);

// mlp_decode.cu
int launch_checkdata(MlpTables m, const uint32_t *seg_au_base, cudaStream_t s);
bool checkdata_windowed(const MlpTables &m);     // which of the two kernels launch_checkdata picks
// a run of decode warps: the groups of substream k of one track
struct DecWork { uint32_t warp0, track, k, pad; };
int launch_mlp_decode(MlpTables m, const DecWork *const work[5], const uint32_t n_work[5], const uint32_t n_warps[5], cudaStream_t s);
int launch_carry_fix(MlpTables m, cudaStream_t s);
int launch_mlp_filter_out(MlpTables m, const DecWork *const work[5], const uint32_t n_work[5], const uint32_t n_warps[5], cudaStream_t s);
size_t au_snap_bytes();
size_t seg_ctx_bytes();
size_t au_delta_bytes();
// fast path: pass A (headers), B (entropy, one lane per access unit), C (filters, one lane per channel)
// headers_only: passes A0 .. A2 alone (the fused pass does the rest)
int launch_mlp_fast(MlpTables m, const DecWork *const work[5], const uint32_t n_work[5], const uint32_t n_warps[5],
                    cudaEvent_t (*kev)[2], bool *kev_used, cudaEvent_t checked, bool headers_only, cudaStream_t s);
const uint16_t *huff_lut_device();      // address of the table on the current device

// mlp_fused.cu: entropy decode + prediction + rematrix + interleaved output in one pass.
// A run of warps: the groups of one track, SUB warps each (lanes = (segment, channel) over both substreams).
struct FusedWork { uint32_t warp0, track, n0, n1; };     // n0, n1: channels of substream 0 / 1 (n1 = 0: one substream)
// work[c], n_work[c], n_warps[c] for class c: 0 = at most two channels per substream, 1 = up to four
int launch_mlp_fused(MlpTables m, const FusedWork *const work[2], const uint32_t n_work[2], const uint32_t n_warps[2], cudaStream_t s);
uint32_t fused_warps_per_group(uint32_t n0, uint32_t n1);
int launch_seg_finalize(MlpTables m, uint32_t *seg_frames, uint32_t *status, cudaStream_t s);
int launch_track_finalize(MlpTables m, const uint64_t *seg_frame_scan, const uint32_t *status, cudaStream_t s);
int launch_rematrix(MlpTables m, uint32_t max_chunks, uint32_t channel_mask, cudaStream_t s);
