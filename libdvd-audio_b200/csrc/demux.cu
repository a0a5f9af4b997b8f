// demux.cu — AOB sector scan, audio packet table, elementary-stream gather and
// PCM unpack kernels.
//
// Replaces (reference tree):
//   src/packet.c:137-188  read_pack_header          -> pack_header_size()
//   src/packet.c:60-135   packet_reader_next_packet / _next_audio_packet
//                                                   -> k_sector_count + k_packet_fill
//   src/dvd-audio.c:1238-1248 read_audio_packet_header -> k_packet_fill
//   src/pcm.c:79-96       dvda_pcmdecoder_decode_params -> k_packet_fill
//   src/bitstream.c:2259,2442 queue enqueue / push copies -> k_es_gather
//   src/pcm.c:98-169      dvda_pcmdecoder_decode_packet + src/dvd-audio.c:781-792
//                         interleave                 -> k_pcm_unpack
//
// The reference pulls one sector at a time through a byte queue; here every
// sector is looked at by its own thread, the audio packets found are numbered by
// a prefix sum, and payload bytes are moved once, coalesced, into one contiguous
// elementary stream.
#include "common.cuh"
#include "kernels.cuh"

// ---------------------------------------------------------------- sector scan

// size of the pack header incl. stuffing, 0 if the sector does not start with a
// valid one (sync 0x000001BA, marker bits 01,1,1,1,1,11)
__device__ __forceinline__ uint32_t pack_header_size(const uint8_t *s)
{
    // the first 16 bytes of a sector are 16-byte aligned: one vector load
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(s));
    if (v.x != 0xBA010000u) return 0;                 // bytes 00 00 01 BA
    const uint32_t b4 = v.y & 0xFF, b6 = (v.y >> 16) & 0xFF;
    const uint32_t b8 = v.z & 0xFF, b9 = (v.z >> 8) & 0xFF;
    const uint32_t b12 = v.w & 0xFF, b13 = (v.w >> 8) & 0xFF;
    if ((b4 >> 6) != 1 || !((b4 >> 2) & 1) || !((b6 >> 2) & 1) || !((b8 >> 2) & 1) ||
        !(b9 & 1) || (b12 & 3) != 3)
        return 0;
    return 14u + (b13 & 7u);
}

struct PacketWalk {
    const uint8_t *sec;
    uint32_t off;
    bool bad;
};

// advances to the next packet of the sector; false when the sector is used up
// (or the chain is broken: w.bad)
__device__ __forceinline__ bool next_packet(PacketWalk &w, uint32_t &id, uint32_t &payload, uint32_t &len)
{
    if (w.off == DVDA_SECTOR) return false;
    if (w.off + 6 > DVDA_SECTOR) { w.bad = true; return false; }
    const uint8_t *h = w.sec + w.off;
    if (ld_u8(h) != 0 || ld_u8(h + 1) != 0 || ld_u8(h + 2) != 1) { w.bad = true; return false; }
    id = ld_u8(h + 3);
    len = ld_be16(h + 4);
    if (w.off + 6 + len > DVDA_SECTOR) { w.bad = true; return false; }
    payload = w.off + 6;
    w.off += 6 + len;
    return true;
}

// one thread per sector: how many audio packets, and is the packet chain intact
__global__ void k_sector_count(const uint8_t *__restrict__ sectors, uint32_t n_sectors,
                               uint32_t *__restrict__ sec_cnt, uint32_t *__restrict__ sec_bad)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_sectors) return;
    PacketWalk w = {sectors + (uint64_t)s * DVDA_SECTOR, 0, false};
    w.off = pack_header_size(w.sec);
    uint32_t cnt = 0;
    if (!w.off) {
        w.bad = true;
    } else {
        uint32_t id, payload, len;
        while (next_packet(w, id, payload, len)) cnt += (id == 0xBD);
    }
    sec_cnt[s] = cnt;
    sec_bad[s] = w.bad ? 1u : 0u;
}

__device__ __forceinline__ uint32_t pcm_bytes_per_sample(uint32_t g0_bps) { return g0_bps == 0 ? 2u : g0_bps == 2 ? 3u : 0u; }
__device__ __forceinline__ uint32_t channels_of(uint32_t a)
{
    // channel count per assignment (reference dvd-audio.c:1459-1496), 3 bits each
    const unsigned long long packed =
        (1ull << 0) | (2ull << 3) | (3ull << 6) | (4ull << 9) | (3ull << 12) | (4ull << 15) | (5ull << 18) |
        (3ull << 21) | (4ull << 24) | (5ull << 27) | (4ull << 30) | (5ull << 33) | (6ull << 36) | (4ull << 39) |
        (5ull << 42) | (4ull << 45) | (5ull << 48) | (6ull << 51) | (5ull << 54) | (5ull << 57) | (6ull << 60);
    return a <= 20 ? (uint32_t)((packed >> (3 * a)) & 7) : 0;
}

// one thread per sector: rows of the packet table
// (`rows`: rows the table has room for — it is sized before the packet count is known to the host)
__global__ void k_packet_fill(const uint8_t *__restrict__ sectors, uint32_t n_sectors,
                              const uint32_t *__restrict__ sec_base, PacketTable pt, uint32_t rows)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_sectors) return;
    PacketWalk w = {sectors + (uint64_t)s * DVDA_SECTOR, 0, false};
    w.off = pack_header_size(w.sec);
    if (!w.off) return;
    uint32_t i = sec_base[s];
    uint32_t id, payload, len;
    while (next_packet(w, id, payload, len)) {
        if (id != 0xBD) continue;
        // "16p 8u" pad_1 skip "8u 8p 8p 8u"
        uint32_t codec = 0xFF, pad2 = 0, hdr = 0;
        const uint8_t *p = w.sec + payload;
        if (len >= 3) {
            const uint32_t pad1 = ld_u8(p + 2);
            if (len >= 3 + pad1 + 4) {
                codec = ld_u8(p + 3 + pad1);
                pad2 = ld_u8(p + 6 + pad1);
                hdr = 7 + pad1;
            }
        }
        uint32_t rest = len - hdr;                   // bytes behind the pad_2_size byte
        uint32_t mlp_len = 0, pcm_frames = 0, params = 0xFFFFFFFFu;
        if (codec == CODEC_MLP) {
            if (pad2 <= rest) mlp_len = rest - pad2; else codec = 0xFF;
        } else if (codec == CODEC_PCM) {
            if (pad2 >= 9 && pad2 <= rest) {
                const uint8_t *q = p + hdr;           // 9 parameter bytes
                const uint32_t b3 = ld_u8(q + 3), b4 = ld_u8(q + 4), asg = ld_u8(q + 6);
                params = (b3 << 16) | (b4 << 8) | asg;
                const uint32_t chunk = pcm_bytes_per_sample(b3 >> 4) * channels_of(asg) * 2;
                if (chunk) pcm_frames = ((rest - pad2) / chunk) * 2;
            } else codec = 0xFF;
        } else {
            codec = 0xFF;
        }
        if (i >= rows) break;                        // the host grows the table and comes back
        pt.sector[i] = s;
        pt.off[i] = (uint16_t)(payload + hdr);
        pt.len[i] = (uint16_t)rest;
        pt.codec[i] = (uint8_t)codec;
        pt.pad2[i] = (uint8_t)pad2;
        pt.params[i] = params;
        pt.mlp_len[i] = mlp_len;
        pt.pcm_frames[i] = pcm_frames;
        i++;
    }
}

// per-packet flags that feed the prefix counts used by track setup
// (runs over all rows of the table; the ones behind the last packet are cleared so that the prefix
// sums over the whole table end in the right totals)
__global__ void k_packet_flags(PacketTable pt, uint32_t rows, const uint32_t *__restrict__ np_dev,
                               uint32_t *__restrict__ nonmlp, uint32_t *__restrict__ pcm_stop)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    if (i >= *np_dev) { pt.mlp_len[i] = 0; pt.pcm_frames[i] = 0; nonmlp[i] = 0; pcm_stop[i] = 0; return; }
    nonmlp[i] = pt.codec[i] != CODEC_MLP;
    // a PCM track stops in front of a packet that is not PCM, changes the stream
    // parameters or holds no whole chunk (reference dvd-audio.c:1042-1056, 770-774)
    uint32_t stop = pt.codec[i] != CODEC_PCM || pt.pcm_frames[i] == 0;
    if (!stop && i > 0) stop = pt.codec[i - 1] != CODEC_PCM || pt.params[i - 1] != pt.params[i];
    pcm_stop[i] = stop;
}

// ------------------------------------------------------- elementary stream

// One warp per MLP packet: payload bytes -> ES[es_off ...].  Byte-granular on
// both sides (neither side is aligned); lanes take consecutive bytes so the
// accesses coalesce into 32-byte segments.
//
// While it has the bytes in hand the warp also looks for the major-sync pattern F8 72 6F BB
// (dvd-audio.c:1250-1286 find_major_sync tests exactly these four bytes, at offset 4 of an access
// unit): every position whose first byte lies in this packet is this warp's, the few positions at
// the packet's edges — the pattern may run on into the next packet — are checked byte by byte
// through the packet table.  A match goes to the slots of its 512-byte chunk of the stream
// (k_sync_validate / k_sync_emit order and judge them later): the elementary stream is not read
// again for the search.
__device__ __forceinline__ uint32_t stream_byte(const uint8_t *__restrict__ sectors, const PacketTable &pt, uint32_t np,
                                                uint32_t row, uint32_t q, bool &have)
{
    // byte q of the stream counted from the first payload byte of packet `row` (q may lie in a later packet)
    while (row < np) {
        const uint32_t n = pt.mlp_len[row];
        if (q < n) { have = true; return ld_u8(sectors + (uint64_t)pt.sector[row] * DVDA_SECTOR + pt.off[row] + pt.pad2[row] + q); }
        q -= n;
        row++;
    }
    have = false;
    return 0;
}
__device__ __forceinline__ void sync_note(uint64_t p, uint32_t *__restrict__ cnt_raw, uint16_t *__restrict__ slots, uint32_t nslots)
{
    const uint32_t chunk = (uint32_t)(p / SYNC_CHUNK);
    const uint32_t i = atomicAdd(&cnt_raw[chunk], 1u);
    if (i < nslots) slots[(uint64_t)chunk * 2 + i] = (uint16_t)(p % SYNC_CHUNK);
}
__global__ void k_es_gather(const uint8_t *__restrict__ sectors, PacketTable pt, uint32_t rows, const DecCounts *__restrict__ cnt,
                            const uint64_t *__restrict__ pk_es, uint8_t *__restrict__ es,
                            uint32_t *__restrict__ cnt_raw, uint16_t *__restrict__ slots, uint32_t nslots)
{
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (warp >= rows) return;
    // (a warp lives for a few memory latencies: everything it will need from the tables is asked
    // for up front, the next packet's row included)
    const uint32_t n = pt.mlp_len[warp];
    const uint32_t np = min((uint32_t)cnt->np, rows);          // (more packets than rows: the table is incomplete, the decode will be repeated)
    uint32_t nrow = warp + 1 < np ? warp + 1 : warp;
    const uint32_t nlen0 = pt.mlp_len[nrow];
    const uint8_t *src = sectors + (uint64_t)pt.sector[warp] * DVDA_SECTOR + pt.off[warp] + pt.pad2[warp];
    const uint8_t *nsrc = sectors + (uint64_t)pt.sector[nrow] * DVDA_SECTOR + pt.off[nrow] + pt.pad2[nrow];
    const uint64_t es_off = pk_es[warp];
    if (!n) return;
    uint32_t nlen = warp + 1 < np ? nlen0 : 0u;
    if (!nlen && warp + 1 < np) {
        // (the next row carries no stream bytes: look further, 32 rows a round — a PCM track
        // in between is thousands of them)
        nrow = np;
        for (uint32_t base = warp + 2; base < np; base += 32) {
            const uint32_t r = base + lane;
            const uint32_t hit = __ballot_sync(0xFFFFFFFFu, r < np && pt.mlp_len[r] != 0);
            if (hit) { nrow = base + __ffs(hit) - 1; break; }
        }
        if (nrow < np) { nlen = pt.mlp_len[nrow]; nsrc = sectors + (uint64_t)pt.sector[nrow] * DVDA_SECTOR + pt.off[nrow] + pt.pad2[nrow]; }
    }
    uint8_t *dst = es + es_off;
    // head bytes up to a 4-byte boundary of dst, then whole words, then the tail
    const uint32_t head = min(n, (uint32_t)((4 - ((uintptr_t)dst & 3)) & 3));
    const uint32_t words = (n - head) >> 2;
    const uint8_t *s = src + head;
    uint32_t *d = reinterpret_cast<uint32_t *>(dst + head);
    const uint32_t mis = (uint32_t)((uintptr_t)s & 3);
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(s - mis);
    // payload offset of the first byte of source word i: lo_w + 4 i
    const int32_t lo_w = (int32_t)head - (int32_t)mis;

    // The positions the words below do not cover: in front of the first word, behind the last one,
    // and the last three of the packet, whose pattern would run on into the next packet.  A lane
    // each (a dozen at most), loaded now, looked at when the copy is done.  (A pattern that would
    // need more than the next packet with bytes — packets of under three bytes — goes the long way
    // through the table.)
    const int32_t hi_w = lo_w + 4 * (int32_t)words;        // first position behind the words' range
    const uint32_t n_lo = lo_w > 0 ? (uint32_t)lo_w : 0u;
    uint32_t first_hi = hi_w > 0 ? (uint32_t)hi_w : 0u;
    if (n >= 3 && first_hi > n - 3) first_hi = n - 3;      // (positions behind n - 4 are nobody's in the loop below)
    if (first_hi < n_lo) first_hi = n_lo;
    const uint32_t n_edge = n_lo + (n > first_hi ? n - first_hi : 0u);
    uint32_t edge_q = 0, edge_v = 0;
    bool edge_ok = false;
    if (lane < n_edge) {
        edge_q = lane < n_lo ? lane : first_hi + (lane - n_lo);
        edge_ok = true;
        if (edge_q + 4 <= n) {
            for (uint32_t k = 0; k < 4; k++) edge_v |= ld_u8(src + edge_q + k) << (8 * k);
        } else if (edge_q + 4 - n <= nlen) {
            for (uint32_t k = 0; k < 4; k++) edge_v |= (edge_q + k < n ? ld_u8(src + edge_q + k) : ld_u8(nsrc + (edge_q + k - n))) << (8 * k);
        } else {
            for (uint32_t k = 0; k < 4 && edge_ok; k++) {
                bool have;
                edge_v |= stream_byte(sectors, pt, np, warp, edge_q + k, have) << (8 * k);
                edge_ok = have;
            }
        }
    }

    if (lane < head) dst[lane] = ld_u8(src + lane);
    // The copy loop stays free of branches (its loads run ahead of one another); it only notes, a bit
    // per round, in which of this lane's source words a byte F8 — the pattern's first — occurred
    // (three instructions, true for one word in sixty).  Those words are looked at again below.
    uint32_t seen = 0;
    {
        uint32_t j = 0;
        for (uint32_t i = lane; i < words; i += 32, j++) {
            const uint32_t a = __ldg(sw + i);
            const uint32_t b = mis ? __ldg(sw + i + 1) : 0u;      // still inside the payload
            d[i] = __funnelshift_r(a, b, mis * 8);
            const uint32_t t = a ^ 0xF8F8F8F8u;
            seen |= ((t - 0x01010101u) & ~t & 0x80808080u) ? 1u << (j & 31) : 0u;
        }
    }
    while (seen) {
        const uint32_t j = __ffs(seen) - 1;
        seen &= seen - 1;
        // (payloads are at most 2 KiB: 16 rounds; a bit stands for rounds j, j + 32, ... all the same)
        for (uint32_t i = lane + 32 * j; i < words; i += 32 * 32) {
            const uint32_t a = __ldg(sw + i);
            const uint32_t b = lo_w + 4 * (int32_t)(i + 1) < (int32_t)n ? __ldg(sw + i + 1) : 0u;   // (only where it still holds payload bytes)
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int32_t q = lo_w + 4 * (int32_t)i + k;
                if (__funnelshift_r(a, b, 8 * k) == 0xBB6F72F8u && q >= 0 && (uint32_t)q + 4 <= n && es_off + (uint32_t)q >= 4)
                    sync_note(es_off + (uint32_t)q - 4, cnt_raw, slots, nslots);
            }
        }
    }
    const uint32_t tail0 = head + words * 4;
    if (tail0 + lane < n) dst[tail0 + lane] = ld_u8(src + tail0 + lane);
    if (edge_ok && edge_v == 0xBB6F72F8u && es_off + edge_q >= 4) sync_note(es_off + edge_q - 4, cnt_raw, slots, nslots);
}

// Zero bytes behind the stream (bit readers may run ahead), and the packet table's capacity:
// with more packets than rows the table is incomplete; the decode then runs on no packets at all
// and the host comes back with a larger table.
__global__ void k_es_tail(uint8_t *__restrict__ es, DecCounts *__restrict__ cnt, uint32_t rows,
                          TrackDev *__restrict__ tracks, uint32_t n_tracks, const __grid_constant__ TrackArgs ta)
{
    // (the track table's rows from the kernel arguments: this block is in the chain anyway)
    for (uint32_t i = threadIdx.x; i < n_tracks; i += blockDim.x) {
        TrackDev T;
        memset(&T, 0, sizeof T);
        T.first_sector = ta.t[i][0]; T.last_sector = ta.t[i][1]; T.pts_length = ta.t[i][2]; T.cont = ta.t[i][3];
        tracks[i] = T;
    }
    const uint64_t end = cnt->es_total;
    for (uint32_t i = threadIdx.x; i < DVDA_ES_PAD / 16; i += blockDim.x)
        reinterpret_cast<uint4 *>(es + ((end + 15) & ~15ull))[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x < 16 && end + threadIdx.x < ((end + 15) & ~15ull)) es[end + threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x == 0 && cnt->np > rows) {
        cnt->overflow |= CAP_ROWS;
        cnt->need_rows = cnt->np;
        cnt->np = 0;
        cnt->es_total = 0;
    }
}

// ---------------------------------------------------------------------- PCM

// Sample order inside a chunk (two frames), as group lists: see
// build_pcm_tables() in engine.cu.  tab[i] = destination byte (little-endian
// sample bytes, sample-major) of chunk byte i — the same permutation the
// reference applies (src/pcm.c:103-166), rebuilt from the layout rule.
__constant__ uint8_t c_pcm_perm[2][6][36];

// One warp per PCM packet (many packets in flight per SM: the work of a packet is a
// chain of small table loads followed by 2 KiB of payload).  The payload is staged
// in shared memory with 16-byte loads; then every lane assembles whole samples (the
// two or three bytes of a sample are found through the inverse of the chunk
// permutation) and consecutive lanes store consecutive samples.
#define PCM_WARPS 8
__global__ void __launch_bounds__(PCM_WARPS * 32)
k_pcm_unpack(const uint8_t *__restrict__ sectors, PacketTable pt, const DecCounts *__restrict__ cnt,
             const uint32_t *__restrict__ status,
             const uint64_t *__restrict__ pk_pf, const TrackDev *__restrict__ tracks,
             const uint32_t *__restrict__ trk_pk_lo, uint32_t n_tracks,
             int32_t *__restrict__ pcm)
{
    __shared__ uint4 stage_all[PCM_WARPS][DVDA_SECTOR / 16 + 2];
    __shared__ uint8_t inv_all[PCM_WARPS][40];
    const uint32_t wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t i = blockIdx.x * PCM_WARPS + wib;
    // (queued before the host has seen the batch's status: nothing is written into an output buffer
    // that turned out too small)
    if (!cnt->any_pcm || (*status & STATUS_PCM_SMALL) || i >= cnt->np || pt.codec[i] != CODEC_PCM) return;
    // owning track: the last one whose first packet is <= i
    const uint32_t t = upper_bound_dev(trk_pk_lo, n_tracks, i);
    if (t == 0) return;
    const TrackDev &T = tracks[t - 1];
    if (T.status != 0 || T.codec != 0 || i < T.pk_lo || i >= T.pcm_pk_end) return;
    uint4 *stage = stage_all[wib];
    uint8_t *inv = inv_all[wib];
    const uint32_t ch = T.channels, bytes = T.bits >> 3, chunk = T.pcm_chunk;
    const uint32_t nchunks = pt.pcm_frames[i] >> 1;
    const uint8_t *src = sectors + (uint64_t)pt.sector[i] * DVDA_SECTOR + pt.off[i] + pt.pad2[i];
    int32_t *dst = pcm + T.out_base + (pk_pf[i] - T.pcm_frame0) * ch;
    const uint8_t *perm = c_pcm_perm[bytes == 3][ch - 1];
    for (uint32_t b = lane; b < chunk; b += 32) inv[perm[b]] = (uint8_t)b;
    // the 16-byte pieces covering the payload (sectors are 16-byte aligned, a packet stays inside its sector)
    const uint32_t used = nchunks * chunk;
    const uint32_t mis = (uint32_t)((uintptr_t)src & 15);
    const uint4 *base = reinterpret_cast<const uint4 *>(src - mis);
    for (uint32_t k = lane; k * 16 < mis + used; k += 32) stage[k] = __ldg(base + k);
    __syncwarp();
    const uint8_t *sb = reinterpret_cast<const uint8_t *>(stage) + mis;
    const uint32_t per_chunk = 2 * ch, nsamples = nchunks * per_chunk;
    // lane -> (chunk, sample in chunk), advanced by 32 samples per step without dividing
    uint32_t k = lane / per_chunk, w = lane - k * per_chunk;
    const uint32_t dk = 32 / per_chunk, dw = 32 - dk * per_chunk;
    for (uint32_t smp = lane; smp < nsamples; smp += 32) {
        const uint8_t *c = sb + k * chunk;
        int32_t v;
        if (bytes == 2) v = (int16_t)(c[inv[2 * w]] | (c[inv[2 * w + 1]] << 8));
        else {
            v = c[inv[3 * w]] | (c[inv[3 * w + 1]] << 8) | (c[inv[3 * w + 2]] << 16);
            v = (v << 8) >> 8;
        }
        dst[smp] = v;
        k += dk; w += dw;
        if (w >= per_chunk) { w -= per_chunk; k++; }
    }
}

int upload_pcm_tables(const uint8_t *tables)
{
    CUDA_TRY(cudaMemcpyToSymbol(c_pcm_perm, tables, 2 * 6 * 36));
    return 0;
}

// ------------------------------------------------------------ host launchers

int launch_sector_count(const uint8_t *sectors, uint32_t n_sectors, uint32_t *sec_cnt, uint32_t *sec_bad, cudaStream_t s)
{
    LAUNCH(k_sector_count, div_up_u32(n_sectors, 128), 128, 0, s, sectors, n_sectors, sec_cnt, sec_bad);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
int launch_packet_fill(const uint8_t *sectors, uint32_t n_sectors, const uint32_t *sec_base, PacketTable pt, uint32_t rows,
                       uint32_t *nonmlp, uint32_t *pcm_stop, cudaStream_t s)
{
    LAUNCH(k_packet_fill, div_up_u32(n_sectors, 128), 128, 0, s, sectors, n_sectors, sec_base, pt, rows);
    if (rows) LAUNCH(k_packet_flags, div_up_u32(rows, 256), 256, 0, s, pt, rows, sec_base + n_sectors, nonmlp, pcm_stop);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
// any_mlp false: the previous decode of this shape met no MLP packet — the gather is left out (an
// MLP track then raises CAP_SHAPE in the track set-up and the decode is repeated with it).
// targs: the track descriptors, written into `tracks` on the way (nullptr: the caller has done that).
int launch_es_gather(const uint8_t *sectors, PacketTable pt, uint32_t rows, const uint64_t *pk_es, uint8_t *es,
                     DecCounts *cnt, uint32_t *cnt_raw, uint16_t *slots, uint32_t nslots, bool any_mlp,
                     TrackDev *tracks, uint32_t n_tracks, const TrackArgs *targs, cudaStream_t s)
{
    if (rows && any_mlp) LAUNCH(k_es_gather, div_up_u32((uint64_t)rows * 32, 256), 256, 0, s, sectors, pt, rows, cnt, pk_es, es, cnt_raw, slots, nslots < 2 ? nslots : 2u);
    static const TrackArgs none = {};
    LAUNCH(k_es_tail, 1, 256, 0, s, es, cnt, rows, tracks, targs ? n_tracks : 0u, targs ? *targs : none);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
int launch_pcm_unpack(const uint8_t *sectors, PacketTable pt, uint32_t rows, const DecCounts *cnt, const uint32_t *status,
                      const uint64_t *pk_pf, const TrackDev *tracks, const uint32_t *trk_pk_lo, uint32_t n_tracks, int32_t *pcm, cudaStream_t s)
{
    if (!rows) return 0;
    LAUNCH(k_pcm_unpack, div_up_u32(rows, PCM_WARPS), PCM_WARPS * 32, 0, s, sectors, pt, cnt, status, pk_pf, tracks, trk_pk_lo, n_tracks, pcm);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
