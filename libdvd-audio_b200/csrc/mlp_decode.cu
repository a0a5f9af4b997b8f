// mlp_decode.cu — MLP access-unit decode kernels.
//
// Replaces (reference tree, src/mlp.c unless noted):
//   :670-712, :1360-1399  read_substream + checkdata_callback   -> k_checkdata
//   :407-612              decode_mlp_frame (sync, directory)    -> au_layout()
//   :714-807              decode_substream / decode_block       -> decode_segment()
//   :809-854              decode_restart_header                 -> restart_header()
//   :856-1120             decoding / matrix / FIR / IIR params  -> decoding_params()
//   :1122-1241            decode_residual_data + src/bitstream.c:1806-1833
//                         br_read_huffman_code                  -> decode_block() entropy loop
//   :1243-1306            filter_channel                        -> decode_block() filter step
//   :1308-1358            rematrix_channels                     -> k_rematrix
//   :514-538, :584-608    output shift, RIFF WAVE order, append -> k_rematrix
//   src/dvd-audio.c:781-792 dvda_read interleave                 -> k_rematrix (writes interleaved)
//
// Parallel layout.  MLP state is re-initialised at a restart header, so a track
// is cut into restart-delimited segments (mlp_index.cu).  32 consecutive
// segments form a group; one warp decodes one (group, substream): lane l parses
// and filters segment l sequentially, all lanes in lock step, and writes its
// samples to the group's tile, laid out [frame][channel][lane] so that the 32
// lanes of a warp store 128 contiguous bytes.  The rematrix kernel transposes
// 32x32 (frame x lane) patches through shared memory, applies the per-AU
// matrices, LSB bypass and output shift, and writes interleaved frames.
#include "common.cuh"
#include "kernels.cuh"

// ------------------------------------------------------------- check data

__constant__ uint8_t c_crc8[256];   // CRC-8 poly 0x63, built by the engine

// which bytes of an access unit belong to which substream
struct AuLayout {
    uint32_t total;        // AU bytes incl. the 4-byte header
    uint32_t data0;        // offset (from AU start) of the substream data area
    uint32_t end[2];       // cumulative substream ends, relative to data0
    uint32_t chk0;         // substream 0's checkdata bit (governs both, mlp.c:543-545)
    bool has_sync, params_differ, ok;
};

__device__ AuLayout au_layout(const uint8_t *es, uint64_t pos, const TrackDev &T)
{
    AuLayout L;
    L.ok = false; L.has_sync = false; L.params_differ = false;
    L.total = (((ld_u8(es + pos) & 15u) << 8) | ld_u8(es + pos + 1)) * 2;
    uint32_t q = 4;
    // major sync: needs all 28 bytes inside the AU (mlp.c:614-654)
    if (L.total >= 32 && ld_be32(es + pos + 4) == 0xF8726FBBu) {
        const uint32_t ns = ld_u8(es + pos + 20) >> 4;
        if (ns == 1 || ns == 2) {
            L.has_sync = true;
            const uint32_t b8 = ld_u8(es + pos + 8), b9 = ld_u8(es + pos + 9), asg = ld_u8(es + pos + 11) & 31;
            L.params_differ = (b8 >> 4) != T.g0_bps || (b8 & 15) != T.g1_bps || (b9 >> 4) != T.g0_rate ||
                              (b9 & 15) != T.g1_rate || asg != T.assignment;
            q = 32;
        }
    }
    L.end[0] = L.end[1] = 0; L.chk0 = 0;
    for (uint32_t k = 0; k < T.nss; k++) {
        if (q + 2 > L.total) return L;
        const uint32_t b0 = ld_u8(es + pos + q);
        L.end[k] = (((b0 & 15u) << 8) | ld_u8(es + pos + q + 1)) * 2;
        if (k == 0) L.chk0 = (b0 >> 5) & 1;
        q += 2 + ((b0 >> 7) ? 2 : 0);
    }
    L.data0 = q;
    if (q > L.total) return L;
    for (uint32_t k = 0; k < T.nss; k++) {
        const uint32_t start = k ? L.end[0] : 0;
        if (L.end[k] < start || q + L.end[k] > L.total) return L;
        if (L.chk0 && L.end[k] - start < 2) return L;
    }
    L.ok = true;
    return L;
}

// One thread per access unit: parity and CRC-8 of each substream.
// au_err: 0 ok, 1 = drop silently (stream parameters changed, mlp.c:452-455),
// else ERR_* bits.
__global__ void k_checkdata(MlpTables m, const uint32_t *__restrict__ seg_au_base)
{
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= m.nau) return;
    const uint32_t si = upper_bound_dev(seg_au_base, m.nseg, a) - 1;
    const TrackDev &T = m.tracks[m.segs[si].track];
    const uint64_t pos = m.au_pos[a];
    const AuLayout L = au_layout(m.es, pos, T);
    uint32_t err = 0;
    if (L.has_sync && L.params_differ) err = 1;
    else if (!L.ok) err = ERR_SYNTAX;
    else if (L.chk0) {
        for (uint32_t k = 0; k < T.nss && !err; k++) {
            const uint32_t start = k ? L.end[0] : 0;
            const uint8_t *p = m.es + pos + L.data0 + start;
            const uint32_t n = L.end[k] - start - 2;
            uint32_t parity = 0, crc = 0x3C, fin = 0;
            for (uint32_t i = 0; i < n; i++) {
                const uint32_t b = ld_u8(p + i);
                parity ^= b;
                fin = crc ^ b;
                crc = c_crc8[fin];
            }
            if (((ld_u8(p + n) ^ parity) & 0xFF) != 0xA9) err = ERR_PARITY;
            else if (ld_u8(p + n + 1) != fin) err = ERR_CRC;
        }
    }
    m.au_err[a] = (uint8_t)err;
}

int upload_crc_table(const uint8_t *t)
{
    CUDA_TRY(cudaMemcpyToSymbol(c_crc8, t, 256));
    return 0;
}

int launch_checkdata(MlpTables m, const uint32_t *seg_au_base, cudaStream_t s)
{
    if (!m.nau) return 0;
    LAUNCH(k_checkdata, div_up_u32(m.nau, 128), 128, 0, s, m, seg_au_base);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------- bit reader

// MSB-first reader over the elementary stream (reference src/bitstream.c:1077-1111).
// `pos` counts bits from word `wbase`, so 32 bits suffice inside one access unit.
struct BitReader {
    const uint32_t *words;
    uint64_t wbase;
    uint32_t pos;
    uint32_t cw;           // word offset of the cached pair, 0xFFFFFFFE = none
    uint32_t hi, lo;
};

__device__ __forceinline__ void br_open(BitReader &b, const uint8_t *es, uint64_t byte_pos)
{
    b.words = reinterpret_cast<const uint32_t *>(es);
    b.wbase = byte_pos >> 2;
    b.pos = (uint32_t)(byte_pos & 3) * 8;
    b.cw = 0xFFFFFFFEu;                        // neither w nor w - 1 for any real word offset
}

// next n bits (1..32) without consuming them
__device__ __forceinline__ uint32_t br_peek(BitReader &b, uint32_t n)
{
    const uint32_t w = b.pos >> 5;
    if (w != b.cw) {
        b.hi = (w == b.cw + 1) ? b.lo : ld_be32_aligned(b.words + b.wbase + w);
        b.lo = ld_be32_aligned(b.words + b.wbase + w + 1);
        b.cw = w;
    }
    return __funnelshift_l(b.lo, b.hi, b.pos & 31) >> (32 - n);
}
__device__ __forceinline__ uint32_t br_get(BitReader &b, uint32_t n)
{
    if (!n) return 0;
    const uint32_t v = br_peek(b, n);
    b.pos += n;
    return v;
}
// two's complement, n in 1..32 (src/bitstream.c:1198-1206)
__device__ __forceinline__ int32_t br_get_s(BitReader &b, uint32_t n)
{
    const uint32_t v = br_get(b, n);
    return (int32_t)(v << (32 - n)) >> (32 - n);
}

// ------------------------------------------------------------ decoder state

struct ChanState {
    int32_t fir_c[8], iir_c[8];
    int32_t fst[8], ist[8];          // circular histories
    uint8_t fir_order, iir_order, fir_shift, iir_shift;
    uint8_t fhead, ihead;            // next write slot
    uint8_t flen, ilen;              // valid entries (<= 8)
    int32_t huff_offset;
    uint8_t codebook, huff_lsbs;
};

struct SubState {
    uint8_t min_ch, max_ch, mmc, noise_shift;
    uint32_t seed;
    uint8_t flags;                   // bit k = presence flag k
    uint8_t have_header, matrix_len, dirty;
    uint16_t block_size;
    int16_t coeff[DVDA_MAX_MAT][DVDA_MAX_CH];
    uint8_t mat_out[DVDA_MAX_MAT], mat_bypass[DVDA_MAX_MAT];
    uint8_t out_shift[DVDA_MAX_CH], q[DVDA_MAX_CH];
    ChanState ch[DVDA_MAX_CH];
};

// Huffman LUT: 9 peeked bits -> value | length << 8; 0xFFFF = invalid code.
// Built from the prefix structure of the three codebooks
// (src/mlp_codebook{1,2,3}.json): 0^z 1 -> 8 - z; 1 + literal -> 7 + literal;
// 01 0^k 1 -> hi_base + k.
__device__ uint16_t huff_entry(uint32_t cb, uint32_t v9)
{
    const uint32_t lit = 3 - cb;                       // literal bits behind a leading 1
    if (v9 & 0x100) return (uint16_t)((7 + ((v9 >> (8 - lit)) & ((1u << lit) - 1))) | ((1 + lit) << 8));
    if (v9 & 0x080) {
        const uint32_t rest = v9 & 0x7F;
        if (!rest) return 0xFFFF;
        const uint32_t k = __clz(rest) - 25;
        return (uint16_t)((7 + (1u << lit) + k) | ((3 + k) << 8));
    }
    if (!v9) return 0xFFFF;
    const uint32_t z = __clz(v9) - 23;
    return (uint16_t)((8 - z) | ((z + 1) << 8));
}

struct DecodeJob {
    uint32_t seg;          // global segment index
    uint32_t k;            // substream
    uint32_t lane;         // column in the group's tile
    bool exact_history;    // FIR history at segment start is the true one
};

// ---- parameter parsing (cold path) ------------------------------------------

__device__ bool restart_header(BitReader &b, SubState &s)
{
    const uint32_t sync = br_get(b, 13), noise_type = br_get(b, 1);
    b.pos += 16;
    s.min_ch = br_get(b, 4); s.max_ch = br_get(b, 4); s.mmc = br_get(b, 4);
    s.noise_shift = br_get(b, 4);
    s.seed = br_get(b, 23);
    b.pos += 19 + 1 + 8 + 16;
    if (sync != 0x18F5 || noise_type != 0) return false;
    if (s.max_ch < s.min_ch || s.mmc < s.max_ch || s.mmc >= DVDA_MAX_CH) return false;
    for (uint32_t c = 0; c <= s.mmc; c++) if (br_get(b, 6) > s.mmc) return false;
    b.pos += 8;
    s.have_header = 1;
    s.dirty = 1;
    return true;
}

__device__ bool filter_params(BitReader &b, ChanState &C, bool iir)
{
    const uint32_t order = br_get(b, 4);
    if (order > 8) return false;
    if (!order) {
        if (iir) { C.iir_order = 0; C.iir_shift = 0; C.ilen = 0; C.ihead = 0; }
        else { C.fir_order = 0; C.fir_shift = 0; }
        return true;
    }
    const uint32_t shift = br_get(b, 4), bits = br_get(b, 5);
    if (bits < 1 || bits > 16) return false;
    const uint32_t cshift = br_get(b, 3);
    if (bits + cshift > 16) return false;
    int32_t *coef = iir ? C.iir_c : C.fir_c;
    for (uint32_t i = 0; i < order; i++) coef[i] = (int32_t)((uint32_t)br_get_s(b, bits) << cshift);
    if (!iir) {
        C.fir_order = order; C.fir_shift = shift;
        if (br_get(b, 1)) return false;
    } else {
        C.iir_order = order; C.iir_shift = shift;
        C.ilen = 0; C.ihead = 0;
        if (br_get(b, 1)) {
            const uint32_t sbits = br_get(b, 4), sshift = br_get(b, 4);
            if (!sbits) return false;                           // reference underflows (G2)
            // first value sent pairs with coeff[0] = most recent (mlp.c:1103-1107):
            // store so that reading backwards from ihead yields sent[0], sent[1], ...
            for (uint32_t i = 0; i < order; i++)
                C.ist[(order - 1 - i) & 7] = (int32_t)((uint32_t)br_get_s(b, sbits) << sshift);
            C.ilen = order; C.ihead = order & 7;
        }
    }
    return true;
}

__device__ bool decoding_params(BitReader &b, SubState &s, bool restart)
{
    if (restart) {
        if (br_get(b, 1)) { uint32_t f = 0; for (int k = 0; k < 8; k++) f |= br_get(b, 1) << k; s.flags = f; }
        else s.flags = 0xFF;
    } else if ((s.flags & 1) && br_get(b, 1)) {
        uint32_t f = 0; for (int k = 0; k < 8; k++) f |= br_get(b, 1) << k; s.flags = f;
    }
    if ((s.flags & 0x80) && br_get(b, 1)) {
        s.block_size = br_get(b, 9);
        if (s.block_size < 8) return false;
    } else if (restart) s.block_size = 8;

    if ((s.flags & 0x40) && br_get(b, 1)) {
        s.dirty = 1;
        s.matrix_len = br_get(b, 4);
        if (s.matrix_len > DVDA_MAX_MAT || s.mmc + 3 > DVDA_MAX_CH) return false;
        for (uint32_t m = 0; m < s.matrix_len; m++) {
            if ((s.mat_out[m] = br_get(b, 4)) > s.mmc) return false;
            const uint32_t frac = br_get(b, 4);
            if (frac > 14) return false;
            s.mat_bypass[m] = br_get(b, 1);
            for (uint32_t c = 0; c < DVDA_MAX_CH; c++) s.coeff[m][c] = 0;
            for (uint32_t c = 0; c < (uint32_t)s.mmc + 3; c++)
                if (br_get(b, 1)) s.coeff[m][c] = (int16_t)((uint32_t)br_get_s(b, frac + 2) << (14 - frac));
        }
    } else if (restart) { s.matrix_len = 0; s.dirty = 1; }

    if ((s.flags & 0x20) && br_get(b, 1)) {
        s.dirty = 1;
        for (uint32_t c = 0; c <= s.mmc; c++) s.out_shift[c] = (uint8_t)(br_get_s(b, 4) & 31);
    } else if (restart) { for (int c = 0; c < DVDA_MAX_CH; c++) s.out_shift[c] = 0; s.dirty = 1; }

    if ((s.flags & 0x10) && br_get(b, 1)) {
        s.dirty = 1;
        for (uint32_t c = 0; c <= s.max_ch; c++) s.q[c] = br_get(b, 4);
    } else if (restart) { for (int c = 0; c < DVDA_MAX_CH; c++) s.q[c] = 0; s.dirty = 1; }

    for (uint32_t c = s.min_ch; c <= s.max_ch; c++) {
        ChanState &C = s.ch[c];
        if (br_get(b, 1)) {
            if ((s.flags & 0x08) && br_get(b, 1)) { if (!filter_params(b, C, false)) return false; }
            else if (restart) { C.fir_order = 0; C.fir_shift = 0; }
            if ((s.flags & 0x04) && br_get(b, 1)) { if (!filter_params(b, C, true)) return false; }
            else if (restart) { C.iir_order = 0; C.iir_shift = 0; C.ilen = 0; C.ihead = 0; }
            if ((s.flags & 0x02) && br_get(b, 1)) C.huff_offset = br_get_s(b, 15);
            else if (restart) C.huff_offset = 0;
            C.codebook = br_get(b, 2);
            C.huff_lsbs = br_get(b, 5);
            if (C.huff_lsbs > 24) return false;
        } else if (restart) {
            C.fir_order = 0; C.fir_shift = 0;
            C.iir_order = 0; C.iir_shift = 0; C.ilen = 0; C.ihead = 0;
            C.huff_offset = 0; C.codebook = 0; C.huff_lsbs = 24;
        }
    }
    return true;
}

// ---- one block: entropy decode + prediction filters (hot path) --------------

// returns frames decoded, 0 on a syntax error
__device__ uint32_t decode_block(const MlpTables &m, const GroupDev &G, const DecodeJob &job,
                                 SubState &s, BitReader &b, uint32_t end_bits, uint32_t frame0,
                                 uint32_t nch, bool governing, const uint16_t (*lut)[512], uint32_t &flags)
{
    if (br_get(b, 1)) {
        const bool restart = br_get(b, 1);
        if (restart && !restart_header(b, s)) return 0;
        if (!s.have_header) return 0;
        if (!decoding_params(b, s, restart)) return 0;
    }
    if (!s.have_header || b.pos > end_bits) return 0;

    const uint32_t n = s.block_size;
    int32_t sho[DVDA_MAX_CH];
    uint32_t lsb_bits[DVDA_MAX_CH], shift[DVDA_MAX_CH];
    for (uint32_t c = s.min_ch; c <= s.max_ch; c++) {
        const ChanState &C = s.ch[c];
        if (C.huff_lsbs < s.q[c]) return 0;
        const uint32_t nb = C.huff_lsbs - s.q[c];
        lsb_bits[c] = nb;
        if (C.codebook) {
            const int ss = (int)nb + 2 - (int)C.codebook;
            sho[c] = C.huff_offset - 7 * (1 << nb) - (ss >= 0 ? (1 << ss) : 0);
        } else {
            sho[c] = C.huff_offset - (nb >= 1 ? (1 << (nb - 1)) : 0);
        }
        if (C.fir_order + C.iir_order > 8) return 0;
        if (C.fir_shift > 0 && C.iir_shift > 0 && C.fir_shift != C.iir_shift) return 0;
        shift[c] = (C.fir_shift > 0 && C.iir_shift > 0) ? C.fir_shift : C.fir_order > 0 ? C.fir_shift : C.iir_shift;
        if (C.iir_order > C.ilen) return 0;                     // reference reads out of bounds (G2)
        if (C.fir_order > C.flen) {
            if (job.exact_history) return 0;                    // reference reads out of bounds (G1)
            flags |= SEG_NEEDS_CARRY;                           // history lives in the previous segment
        }
    }

    int32_t *tile = m.tiles + G.tile_off + job.lane;
    uint8_t *byp = m.bypass + G.byp_off + job.lane;
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t f = frame0 + i;
        const bool room = f < G.cap;
        if (!room) flags |= SEG_OVERFLOW;
        uint32_t bmask = 0;
        for (uint32_t k = 0; k < s.matrix_len; k++)
            if (s.mat_bypass[k]) bmask |= br_get(b, 1) << k;
        if (governing && room) byp[(uint64_t)f * DVDA_LANES] = (uint8_t)bmask;
        for (uint32_t c = s.min_ch; c <= s.max_ch; c++) {
            ChanState &C = s.ch[c];
            int32_t msb = 0;
            if (C.codebook) {
                const uint32_t e = lut[C.codebook - 1][br_peek(b, 9)];
                if (e == 0xFFFF) return 0;
                msb = e & 0xFF;
                b.pos += e >> 8;
            }
            const int32_t lsb = (int32_t)br_get(b, lsb_bits[c]);
            const int32_t res = (int32_t)((uint32_t)((msb << lsb_bits[c]) + lsb + sho[c]) << s.q[c]);
            // prediction: FIR over previous outputs, IIR over previous residual-ish state
            long long sum = 0;
            for (uint32_t j = 0; j < C.fir_order; j++) sum += (long long)C.fir_c[j] * C.fst[(C.fhead - 1 - j) & 7];
            for (uint32_t j = 0; j < C.iir_order; j++) sum += (long long)C.iir_c[j] * C.ist[(C.ihead - 1 - j) & 7];
            const int32_t ssum = (int32_t)(sum >> shift[c]);
            int32_t v = (int32_t)((uint32_t)ssum + (uint32_t)res);
            const uint32_t q = s.q[c];
            v = (v >> q) << q;
            C.fst[C.fhead] = v; C.fhead = (C.fhead + 1) & 7; if (C.flen < 8) C.flen++;
            C.ist[C.ihead] = (int32_t)((uint32_t)v - (uint32_t)ssum); C.ihead = (C.ihead + 1) & 7; if (C.ilen < 8) C.ilen++;
            if (room) tile[((uint64_t)f * nch + c) * DVDA_LANES] = v;
        }
        if (b.pos > end_bits) return 0;
    }
    return n;
}

__device__ __forceinline__ uint32_t noise_step(uint32_t seed)
{
    const uint32_t sh = (seed >> 7) & 0xFFFF;
    return (seed << 16) ^ sh ^ (sh << 5);
}

// Decodes substream job.k of segment job.seg, access unit by access unit.
__device__ void decode_segment(const MlpTables &m, const DecodeJob &job, const uint16_t (*lut)[512],
                               const int32_t *init_hist)
{
    SegDev &S = m.segs[job.seg];
    const TrackDev &T = m.tracks[S.track];
    const GroupDev &G = m.groups[T.grp_base + (job.seg - T.seg_base) / DVDA_LANES];
    const bool governing = job.k + 1 == T.nss;
    const uint32_t nch = T.channels;

    SubState s;
    memset(&s, 0, sizeof s);
    s.flags = 0xFF;
    if (init_hist) {
        for (int c = 0; c < DVDA_MAX_CH; c++) {
            for (int j = 0; j < 8; j++) s.ch[c].fst[j] = init_hist[c * 8 + j];
            s.ch[c].flen = 8; s.ch[c].fhead = 0;
        }
    }

    uint32_t frames = 0, flags = 0, err = 0, stop_au = 0xFFFFFFFFu, pset = 0xFFFFFFFFu;
    for (uint32_t a = 0; a < S.n_au; a++) {
        const uint32_t A = S.au_base + a;
        const uint64_t pos = m.au_pos[A];
        const uint32_t e = m.au_err[A];
        const AuLayout L = au_layout(m.es, pos, T);
        if (pos + L.total > T.es_cut) { stop_au = a; break; }     // end of track, not an error
        if (e == 1) {                                             // dropped access unit
            if (governing) { AuDev R = {frames, 0, s.seed, pset}; m.au[A] = R; }
            m.au_frames_ss[job.k * m.nau + A] = 0;
            continue;
        }
        if (e) { err |= e; stop_au = a; break; }
        const uint32_t start = job.k ? L.end[0] : 0;
        const uint32_t len = L.end[job.k] - start - (L.chk0 ? 2 : 0);
        BitReader b;
        br_open(b, m.es, pos + L.data0 + start);
        const uint32_t end_bits = b.pos + len * 8;
        const uint32_t au_frame0 = frames;
        bool bad = false;
        for (;;) {
            const uint32_t n = decode_block(m, G, job, s, b, end_bits, frames, nch, governing, lut, flags);
            if (!n) { bad = true; break; }
            frames += n;
            const uint32_t last = br_get(b, 1);
            if (b.pos > end_bits) { bad = true; break; }
            if (last) break;
        }
        if (bad) { err |= ERR_SYNTAX; stop_au = a; frames = au_frame0; break; }
        const uint32_t nf = frames - au_frame0;
        // a restart header inside the AU reloads the noise seed before the AU is
        // rematrixed (mlp.c:828, 504-512): take it after the blocks
        const uint32_t seed0 = s.seed;
        m.au_frames_ss[job.k * m.nau + A] = nf;
        if (governing) {
            // parameters in force after the AU's last block govern the whole AU (mlp.c:504-525)
            if (s.dirty || pset == 0xFFFFFFFFu) {
                ParamSet P;
                memset(&P, 0, sizeof P);
                P.matrix_len = s.matrix_len; P.mmc = s.mmc; P.noise_shift = s.noise_shift;
                uint32_t uses = 0;
                for (uint32_t k = 0; k < s.matrix_len; k++) {
                    P.out_ch[k] = s.mat_out[k];
                    for (int c = 0; c < DVDA_MAX_CH; c++) P.coeff[k][c] = s.coeff[k][c];
                    uses |= (s.coeff[k][s.mmc + 1] != 0) | (s.coeff[k][s.mmc + 2] != 0);
                }
                P.uses_noise = uses;
                for (int c = 0; c < DVDA_MAX_CH; c++) { P.q[c] = s.q[c]; P.out_shift[c] = s.out_shift[c]; }
                m.psets[A] = P;
                pset = A;
                s.dirty = 0;
            }
            AuDev R = {au_frame0, nf, seed0, pset};
            m.au[A] = R;
            uint32_t seed = s.seed;
            for (uint32_t i = 0; i < nf; i++) seed = noise_step(seed);
            s.seed = seed;
        }
    }

    // last 8 outputs per channel, oldest first is not needed: store most-recent-last order
    int32_t *tail = m.fir_tail + ((uint64_t)job.k * m.nseg + job.seg) * (DVDA_MAX_CH * 8);
    for (uint32_t c = s.min_ch; c <= s.max_ch && s.have_header; c++) {
        const ChanState &C = s.ch[c];
        // slot j of the stored tail = what a fresh circular buffer with fhead = 0 expects
        for (int j = 0; j < 8; j++) tail[c * 8 + j] = C.fst[(C.fhead + j) & 7];
    }
    m.ss_flags[job.k * m.nseg + job.seg] = flags;
    if (job.k == 0) S.frames = frames;
    if (err) atomicOr(&S.err, err);
    if (stop_au != 0xFFFFFFFFu) atomicMin(&S.err_au, stop_au);
}

#define DEC_WARPS 4

// one warp per (group, substream); lane = segment of the group
__global__ void __launch_bounds__(DEC_WARPS * 32) k_mlp_decode(MlpTables m)
{
    __shared__ uint16_t lut[3][512];
    for (uint32_t i = threadIdx.x; i < 3 * 512; i += blockDim.x) lut[i >> 9][i & 511] = huff_entry((i >> 9) + 1, i & 511);
    __syncthreads();
    const uint32_t warp = blockIdx.x * DEC_WARPS + (threadIdx.x >> 5);
    const uint32_t lane = threadIdx.x & 31;
    // work items: (group, substream) pairs, substream-major inside a group
    const uint32_t g = warp >> 1, k = warp & 1;
    if (g >= m.ngroups) return;
    const GroupDev &G = m.groups[g];
    const TrackDev &T = m.tracks[G.track];
    if (k >= T.nss || lane >= G.nseg) return;
    DecodeJob job;
    job.seg = G.seg0 + lane;
    job.k = k;
    job.lane = lane;
    job.exact_history = (job.seg == T.seg_base);      // a track starts with empty histories
    decode_segment(m, job, lut, nullptr);
}

int launch_mlp_decode(MlpTables m, cudaStream_t s)
{
    if (!m.ngroups) return 0;
    LAUNCH(k_mlp_decode, div_up_u32((uint64_t)m.ngroups * 2, DEC_WARPS), DEC_WARPS * 32, 0, s, m);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// Segments whose first filtered block needs FIR history from the previous
// segment (the reference never clears it, mlp.c:948-952) are decoded again, in
// order, run by run: one thread per (run head, substream).
__global__ void k_carry_fix(MlpTables m)
{
    __shared__ uint16_t lut[3][512];
    for (uint32_t i = threadIdx.x; i < 3 * 512; i += blockDim.x) lut[i >> 9][i & 511] = huff_entry((i >> 9) + 1, i & 511);
    __syncthreads();
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t seg = idx >> 1, k = idx & 1;
    if (seg >= m.nseg) return;
    const TrackDev &T = m.tracks[m.segs[seg].track];
    if (k >= T.nss) return;
    const uint32_t *fl = m.ss_flags_prev + (uint64_t)k * m.nseg;   // as they were before any fix-up
    if (!(fl[seg] & SEG_NEEDS_CARRY)) return;
    // run head: predecessor (same track) is not waiting for a carry itself
    if (seg > T.seg_base && (fl[seg - 1] & SEG_NEEDS_CARRY)) return;
    const uint32_t track_end = T.seg_base + T.nseg;
    for (uint32_t s = seg; s < track_end && (fl[s] & SEG_NEEDS_CARRY); s++) {
        DecodeJob job;
        job.seg = s; job.k = k; job.lane = (s - T.seg_base) % DVDA_LANES; job.exact_history = true;
        const int32_t *prev = m.fir_tail + ((uint64_t)k * m.nseg + (s - 1)) * (DVDA_MAX_CH * 8);
        // parsing does not depend on filter history, so the error bookkeeping of the
        // first pass (merged with atomics) is reproduced exactly
        decode_segment(m, job, lut, s > T.seg_base ? prev : nullptr);
    }
}

int launch_carry_fix(MlpTables m, cudaStream_t s)
{
    if (!m.nseg) return 0;
    LAUNCH(k_carry_fix, div_up_u32((uint64_t)m.nseg * 2, 64), 64, 0, s, m);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// ----------------------------------------------------------- bookkeeping

// one thread per segment: merge the substreams' verdicts, count frames
__global__ void k_seg_finalize(MlpTables m, uint32_t *__restrict__ seg_frames, uint32_t *__restrict__ status)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.nseg) return;
    SegDev &S = m.segs[i];
    TrackDev &T = m.tracks[S.track];
    // flags of this decode attempt; S.flags keeps SEG_OVERFLOW from an earlier one so
    // that a second attempt sizes the tile from the counted frames
    const uint32_t now = m.ss_flags[i] | (T.nss == 2 ? m.ss_flags[m.nseg + i] : 0);
    uint32_t flags = S.flags | now;
    uint32_t err = S.err, stop = S.err_au;
    uint32_t frames = 0;
    const uint32_t lim = min(stop, S.n_au);
    for (uint32_t a = 0; a < lim; a++) {
        const uint32_t A = S.au_base + a;
        const uint32_t nf = m.au_frames_ss[A];
        if (T.nss == 2 && m.au_frames_ss[m.nau + A] != nf) { err |= ERR_SYNTAX; stop = a; break; }
        frames += nf;
    }
    if ((flags & SEG_IRREGULAR) && stop == 0xFFFFFFFFu) { err |= ERR_SYNTAX; stop = S.n_au; }
    S.flags = flags & ~SEG_NEEDS_CARRY;
    if (now & SEG_OVERFLOW) atomicOr(status, SEG_OVERFLOW);
    S.err = err;
    S.err_au = stop;
    S.frames = frames;
    seg_frames[i] = frames;
    if (stop != 0xFFFFFFFFu) {
        atomicMin(&T.err_seg, i - T.seg_base);
        if (err) atomicOr((unsigned int *)&T.error_flags, err);
    }
}

// one thread per segment: position in the track's output; thread of the first
// segment also totals the track
__global__ void k_track_finalize(MlpTables m, const uint64_t *__restrict__ scan)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.nseg) return;
    SegDev &S = m.segs[i];
    TrackDev &T = m.tracks[S.track];
    const uint32_t local = i - T.seg_base;
    S.frame0 = scan[i] - scan[T.seg_base];
    if (local > T.err_seg) S.frames = 0;                  // behind the point where the track ended
    if (local == 0) {
        const uint32_t last = min(T.err_seg, T.nseg - 1);
        T.frames = scan[T.seg_base + last + 1] - scan[T.seg_base];
    }
}

int launch_seg_finalize(MlpTables m, uint32_t *seg_frames, uint32_t *status, cudaStream_t s)
{
    if (!m.nseg) return 0;
    LAUNCH(k_seg_finalize, div_up_u32(m.nseg, 128), 128, 0, s, m, seg_frames, status);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
int launch_track_finalize(MlpTables m, const uint64_t *seg_frame_scan, cudaStream_t s)
{
    if (!m.nseg) return 0;
    LAUNCH(k_track_finalize, div_up_u32(m.nseg, 128), 128, 0, s, m, seg_frame_scan);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------- rematrix

// RIFF WAVE slot of MLP channel c (table at mlp.c:416-438): identity except for
// assignments 0x12-0x14
__device__ __forceinline__ uint32_t wave_slot(uint32_t assignment, uint32_t c)
{
    if (assignment == 0x12 || assignment == 0x13) return (0x24310u >> (4 * c)) & 15;      // 0,1,3,4,2
    if (assignment == 0x14) return (0x325410u >> (4 * c)) & 15;                          // 0,1,4,5,2,3
    return c;
}

#define RM_THREADS 256

// One block per (group, 32-frame chunk).  Loads the [32 frames][nch][32 lanes]
// patch of the tile coalesced, then every warp takes segments (lanes of the
// patch) and its 32 threads take the 32 frames: noise, matrices, bypass, shift,
// channel order, interleaved store.
__global__ void __launch_bounds__(RM_THREADS) k_rematrix(MlpTables m, const uint64_t *__restrict__ grp_chunk_base)
{
    extern __shared__ int32_t sm[];                      // [nch][32][33] samples, then [32][33] bypass bytes as ints
    const uint64_t chunk = blockIdx.x;
    const uint32_t g = upper_bound_dev(grp_chunk_base, m.ngroups, chunk) - 1;
    const GroupDev &G = m.groups[g];
    const TrackDev &T = m.tracks[G.track];
    const uint32_t nch = T.channels;
    const uint32_t f0 = (uint32_t)(chunk - grp_chunk_base[g]) * 32;
    const uint32_t nf = min(32u, G.cap - f0);
    int32_t *bsm = sm + nch * 32 * 33;

    // coalesced load: consecutive threads read consecutive lanes
    const int32_t *src = m.tiles + G.tile_off + (uint64_t)f0 * nch * DVDA_LANES;
    for (uint32_t i = threadIdx.x; i < nf * nch * 32; i += RM_THREADS) {
        const uint32_t l = i & 31, c = (i >> 5) % nch, f = (i >> 5) / nch;
        sm[(c * 32 + f) * 33 + l] = src[i];
    }
    const uint8_t *bsrc = m.bypass + G.byp_off + (uint64_t)f0 * DVDA_LANES;
    for (uint32_t i = threadIdx.x; i < nf * 32; i += RM_THREADS) bsm[(i >> 5) * 33 + (i & 31)] = bsrc[i];
    __syncthreads();

    const uint32_t f = threadIdx.x & 31;                 // frame inside the chunk
    for (uint32_t l = threadIdx.x >> 5; l < G.nseg; l += RM_THREADS / 32) {
        const SegDev &S = m.segs[G.seg0 + l];
        const uint32_t F = f0 + f;                       // frame inside the segment
        if (F >= S.frames) continue;
        // access unit holding frame F: last one with frame0 <= F among the decoded ones
        uint32_t lo = 0, hi = min(S.n_au, S.err_au);
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (m.au[S.au_base + mid].frame0 <= F) lo = mid; else hi = mid;
        }
        // dropped AUs have no frames and share frame0 with their successor: take the last match
        const AuDev au = m.au[S.au_base + lo];
        const ParamSet &P = m.psets[au.pset];
        int32_t v[DVDA_MAX_CH];
#pragma unroll
        for (uint32_t c = 0; c < DVDA_MAX_CH; c++) v[c] = (c < nch) ? sm[(c * 32 + f) * 33 + l] : 0;
        const uint32_t ml = P.matrix_len;
        if (ml) {
            int32_t n0 = 0, n1 = 0;
            if (P.uses_noise) {
                uint32_t seed = au.seed;
                for (uint32_t i = au.frame0; i < F; i++) seed = noise_step(seed);
                const uint32_t sh = (seed >> 7) & 0xFFFF;
                n0 = (int32_t)((uint32_t)(int32_t)(int8_t)(seed >> 15) << P.noise_shift);
                n1 = (int32_t)((uint32_t)(int32_t)(int8_t)sh << P.noise_shift);
            }
            const uint32_t bm = (uint32_t)bsm[f * 33 + l];
            for (uint32_t k = 0; k < ml; k++) {
                long long sum = 0;
#pragma unroll
                for (uint32_t c = 0; c < DVDA_MAX_CH; c++)
                    if (c <= P.mmc) sum += (long long)v[c] * P.coeff[k][c];
                sum += (long long)n0 * P.coeff[k][P.mmc + 1];
                sum += (long long)n1 * P.coeff[k][P.mmc + 2];
                const uint32_t oc = P.out_ch[k], q = P.q[oc];
                const int32_t r = (((int32_t)(sum >> 14)) >> q << q) + (int32_t)((bm >> k) & 1);
#pragma unroll
                for (uint32_t c = 0; c < DVDA_MAX_CH; c++) if (c == oc) v[c] = r;
            }
        }
        int32_t *dst = m.pcm + T.out_base + (S.frame0 + F) * nch;
#pragma unroll
        for (uint32_t c = 0; c < DVDA_MAX_CH; c++) {
            if (c < nch) {
                int32_t x = v[c];
                if (c <= P.mmc) x = (int32_t)((uint32_t)x << P.out_shift[c]);
                dst[wave_slot(T.assignment, c)] = x;
            }
        }
    }
}

int launch_rematrix(MlpTables m, uint64_t total_chunks, const uint64_t *grp_chunk_base, cudaStream_t s)
{
    if (!total_chunks) return 0;
    const size_t smem = (size_t)(DVDA_MAX_CH + 1) * 32 * 33 * sizeof(int32_t);
    static bool attr_set = false;
    if (!attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(k_rematrix, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    LAUNCH(k_rematrix, (uint32_t)total_chunks, RM_THREADS, smem, s, m, grp_chunk_base);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
