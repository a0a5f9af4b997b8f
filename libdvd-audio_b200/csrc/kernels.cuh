// kernels.cuh — tables shared between the kernel files and the host-side
// launchers the engine (engine.cu) calls.
//
// The launch sequence of a decode is static: nothing the kernels find out about the input
// (packet, sync, segment, access-unit counts ...) goes back to the host before the decode is
// over.  Every table is sized in advance (cap_* / rows), the kernels take their bounds from the
// DecCounts record in device memory and their grids cover the tables' capacities.
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

struct AuSnap;      // mlp_common.cuh: what pass B needs to entropy-decode one access unit
struct SegCtx;      // mlp_common.cuh: what a segment's restart header fixes for its parameter blocks
struct AuDelta;     // mlp_common.cuh: the parameters one access unit transmits (head)
struct ChanCoef;    //   ... a channel's filter coefficients and histories
struct MatCoef;     //   ... the matrix coefficients

// everything the MLP kernels need to find their data.  Counts live in device memory (cnt);
// the cap_* members are what the tables were sized for (strides of the per-substream tables).
struct MlpTables {
    const uint8_t *es;             // elementary stream (padded with DVDA_ES_PAD zero bytes)
    const uint64_t *pk_es;         // [np + 1] ES offset of every packet
    const DecCounts *cnt;          // np, es_total, nseg, ngroups, nau, ...
    uint32_t cap_seg;              // rows of the segment table; stride of the [2][cap_seg] tables
    uint32_t cap_au;               // rows of the access-unit tables (+ 1 spare); stride of the [2][cap_au] tables
    uint32_t cap_grp;
    TrackDev *tracks;
    uint32_t n_tracks;
    SegDev *segs;
    GroupDev *groups;
    uint64_t *au_pos;              // [nau] ES offset of every access unit
    uint8_t *au_err;               // [nau] 0 ok, 1 drop (parameter change), else ERR_* bits
    AuDev *au;                     // [nau]
    ParamSet *psets;               // [nau]
    uint32_t *au_frames_ss;        // [2][cap_au] frames each substream decoded
    uint32_t *ss_flags;            // [2][cap_seg] SEG_* per substream
    uint32_t *ss_flags_prev;       // snapshot taken before the carry fix-up
    uint32_t *ss_flags_fast;       // snapshot taken after passes A and B of the fast path
    int32_t *fir_tail;             // [2][cap_seg][8 ch][8] last outputs per channel
    int32_t *tiles;
    uint8_t *bypass;
    int32_t *pcm;
    AuSnap *au_snap;               // [2][cap_au]
    uint32_t *au_seg;              // [nau]: segment of every access unit
    uint8_t *au_fchg;              // [2][cap_au]: bit cc = the filter set-up of channel cc changes with this access unit
    SegCtx *seg_ctx;               // [2][cap_seg]
    AuDelta *au_delta;             // [2][cap_au], written where the AU brings parameters
    ChanCoef *au_cf;               // [2][cap_au][4], where a channel's filters are transmitted
    MatCoef *au_mcoef;             // [2][cap_au], where matrices are transmitted
    uint32_t fast;                 // 1: the complete decoder only takes segments flagged SEG_FALLBACK
    const uint32_t *status;        // the batch's status word (SEG_OVERFLOW, STATUS_*)
    uint32_t *any_fallback;        // set by the flag kernels of the fast path when the complete decoder has work at all
    uint32_t *seg_need;            // [cap_seg] frames a segment turned out to need (tile overflow: the decode is repeated)
    uint32_t *status_rw;           // the batch's status word, for the kernels that set bits in it
    const uint16_t *huff_lut;      // [4][512] Huffman look-up table in device memory (built once per device)
};

// Track descriptors of a decode as a kernel argument (up to TRACKS_BY_ARG tracks): no load from
// host memory stands in the decode chain.
#define TRACKS_BY_ARG 192
struct TrackArgs { uint32_t t[TRACKS_BY_ARG][4]; };   // first sector, last sector, PTS length, DVDAGPU_PART_* / budget flags

// demux.cu
int upload_pcm_tables(const uint8_t *tables);
int launch_sector_count(const uint8_t *sectors, uint32_t n_sectors, uint32_t *sec_cnt, uint32_t *sec_bad, cudaStream_t s);
// rows: rows the packet table has room for (the packet count is still on the device: sec_base[n_sectors])
int launch_packet_fill(const uint8_t *sectors, uint32_t n_sectors, const uint32_t *sec_base, PacketTable pt, uint32_t rows,
                       uint32_t *nonmlp, uint32_t *pcm_stop, cudaStream_t s);
// one warp per row of the packet table (the rows behind the last packet are empty); also zeroes the
// pad behind the stream and checks the table's capacity against the packet count
// ... and notes the major-sync patterns it comes across in the slots of their 512-byte chunks
// (cnt_raw must be zero before)
int launch_es_gather(const uint8_t *sectors, PacketTable pt, uint32_t rows, const uint64_t *pk_es, uint8_t *es,
                     DecCounts *cnt, uint32_t *cnt_raw, uint16_t *slots, uint32_t nslots, bool any_mlp,
                     TrackDev *tracks, uint32_t n_tracks, const TrackArgs *targs, cudaStream_t s);
int launch_pcm_unpack(const uint8_t *sectors, PacketTable pt, uint32_t rows, const DecCounts *cnt, const uint32_t *status,
                      const uint64_t *pk_pf, const TrackDev *tracks, const uint32_t *trk_pk_lo, uint32_t n_tracks, int32_t *pcm, cudaStream_t s);

// mlp_index.cu
#define SYNC_CHUNK 512u           // ES bytes per warp step of the sync search
#define SYNC_SLOT_BYTES 4u        // per chunk: 2 slots of 16 bits
// chunks_cap: chunks the stream buffer has room for (the stream's size is still on the device)
// the chunks' matches (found by the gather) put in order and judged: which of them start a restart segment
int launch_sync_validate(const uint8_t *es, const DecCounts *cnt, uint32_t chunks_cap, const uint32_t *cnt_raw, uint32_t *cnt_valid,
                         uint16_t *slots, uint32_t nslots, cudaStream_t s);
int launch_sync_fill(const uint8_t *es, const DecCounts *cnt, uint32_t chunks_cap, const uint32_t *cnt_raw, const uint16_t *slots, uint32_t nslots,
                     const uint32_t *base_raw, const uint32_t *base_valid, uint64_t *raw, uint32_t cap_raw,
                     uint64_t *valid, uint32_t cap_valid, cudaStream_t s);
struct TrackSetupArgs {
    const uint8_t *es;
    DecCounts *cnt;                // np, n_raw, n_valid (still on the device when the kernel runs)
    uint32_t mlp_searched;         // 0: the sync search was left out (no MLP track expected): an MLP track raises CAP_SHAPE
    uint32_t n_sectors;
    const uint32_t *sec_base;      // [n_sectors + 1]
    const uint32_t *bad_prefix;    // [n_sectors + 1]
    PacketTable pt;
    const uint64_t *pk_es;         // [np + 1]
    const uint64_t *pk_pf;         // [np + 1] PCM frames
    const uint32_t *pk_nonmlp;     // [np + 1]
    const uint32_t *pk_pcm_stop;   // [np + 1]
    // the sync lists were sized in advance (cap_*): entries beyond were not written and the host comes back
    const uint64_t *raw;
    uint32_t cap_raw;
    const uint64_t *valid;
    uint32_t cap_valid;
};
int launch_track_setup(TrackSetupArgs a, TrackDev *tracks, uint32_t n_tracks, cudaStream_t s);
// What the host used to do between the track set-up and the segment table: the tracks' places in
// the segment and group tables, the work lists, totals and flags into cnt, capacities checked.
struct DecWork { uint32_t warp0, track, k, nch; };   // a run of warps: the groups of substream k (nch channels) of one track
// the output pass of the fast path: a run of warps per track (n0, n1: channels of substream 0 / 1, n1 = 0: one substream)
struct OutWork { uint32_t warp0, track, n0, n1; };
// after the access units are counted and the groups set up: capacities of the access-unit tables, the tiles
// and the grids that were sized from the previous decode (lim_*: what the launches of this decode cover)
struct PlanLimits { uint32_t cap_au; uint64_t cap_cells; uint32_t max_au, nss, out_warps, pcm, mlp; };
int launch_track_plan(TrackDev *tracks, uint32_t n_tracks, uint32_t *trk_pk_lo, uint32_t *trk_seg_base, uint32_t *trk_grp_base,
                      DecWork *work, OutWork *out_work, uint32_t cap_work, uint32_t cap_seg, uint32_t cap_grp, uint32_t cap_sync,
                      DecCounts *cnt, const PlanLimits *check_now, cudaStream_t s);
// (check_now: no MLP side follows — the plan's block does k_plan_check's work itself)
int launch_plan_check(DecCounts *cnt, PlanLimits lim, cudaStream_t s);
int launch_segment_fill(const TrackDev *tracks, uint32_t n_tracks, const uint32_t *trk_seg_base,
                        const uint64_t *valid, SegDev *segs, uint32_t cap_seg, const DecCounts *cnt, cudaStream_t s);
// noted: scratch of au_noted_bytes(cap_seg) bytes shared by the two passes (count, then fill)
size_t au_noted_bytes(uint32_t nseg);
int launch_au_chase(const uint8_t *es, SegDev *segs, uint32_t cap_seg, const DecCounts *cnt, const TrackDev *tracks,
                    uint32_t *seg_nau, uint64_t *au_pos, uint32_t *au_seg, const uint32_t *seg_au_base,
                    uint32_t *noted, int fill, cudaStream_t s);
// any_parts: some track of the batch is a part that is continued by another one
int launch_yield(MlpTables m, uint32_t rows, const uint32_t *seg_au_base, PacketTable pt, const uint32_t *trk_pk_lo,
                 uint8_t *pk_yield, bool any_parts, cudaStream_t s);
int launch_group_setup(const TrackDev *tracks, uint32_t n_tracks, const uint32_t *trk_grp_base, const SegDev *segs,
                       GroupDev *groups, uint32_t cap_grp, DecCounts *cnt, uint32_t *grp_cells, const uint32_t *seg_need, cudaStream_t s);
int launch_group_offsets(GroupDev *groups, uint32_t cap_grp, const DecCounts *cnt, const uint64_t *cell_base,
                         void *zero, size_t zero_bytes, uint32_t *flags, uint32_t n_flags, cudaStream_t s);

// mlp_decode.cu
// windowed: which of the two check-data kernels (small access units: shared-memory windows)
int launch_checkdata(MlpTables m, const uint32_t *seg_au_base, bool windowed, cudaStream_t s);
#define CHK_WINDOWED_MAX_AU_BYTES 512
// the complete decoder over the work list (cap_pairs: (group, substream) pairs the grid covers)
int launch_mlp_decode(MlpTables m, const DecWork *work, uint32_t cap_pairs, cudaStream_t s);
int launch_carry_fix(MlpTables m, cudaStream_t s);
int launch_mlp_filter_out(MlpTables m, const OutWork *work, uint32_t cap_warps, cudaStream_t s);
size_t au_snap_bytes();
size_t seg_ctx_bytes();
size_t au_delta_bytes();                 // per access unit and substream, all three tables
void au_delta_split(void *base, size_t entries, MlpTables &m);   // places the three tables in one buffer of entries * au_delta_bytes()
// fast path: pass A (headers), B (entropy, one lane per access unit), then the flags
int launch_mlp_fast(MlpTables m, const DecWork *work, uint32_t cap_pairs, uint32_t lim_max_au, uint32_t lim_nss,
                    cudaEvent_t (*kev)[2], bool *kev_used, cudaEvent_t checked, cudaStream_t s);
const uint16_t *huff_lut_device();      // address of the table on the current device
int launch_seg_finalize(MlpTables m, uint32_t *seg_frames, uint32_t *status, cudaStream_t s);
int launch_track_finalize(MlpTables m, const uint64_t *seg_frame_scan, const uint32_t *status, cudaStream_t s);
int launch_rematrix(MlpTables m, cudaStream_t s);
