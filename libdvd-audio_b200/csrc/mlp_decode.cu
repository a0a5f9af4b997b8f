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
#include "mlp_common.cuh"
#include "../../include/dvdagpu.h"


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

// BYTES: byte source, b(i) = byte i of the access unit
template <typename BYTES>
__device__ __forceinline__ AuLayout au_layout_from(const BYTES &b, const TrackDev &T)
{
    AuLayout L;
    L.ok = false; L.has_sync = false; L.params_differ = false;
    L.total = (((b(0) & 15u) << 8) | b(1)) * 2;
    uint32_t q = 4;
    // major sync: needs all 28 bytes inside the AU (mlp.c:614-654)
    if (L.total >= 32 && ((b(4) << 24) | (b(5) << 16) | (b(6) << 8) | b(7)) == 0xF8726FBBu) {
        const uint32_t ns = b(20) >> 4;
        if (ns == 1 || ns == 2) {
            L.has_sync = true;
            const uint32_t b8 = b(8), b9 = b(9), asg = b(11) & 31;
            L.params_differ = (b8 >> 4) != T.g0_bps || (b8 & 15) != T.g1_bps || (b9 >> 4) != T.g0_rate ||
                              (b9 & 15) != T.g1_rate || asg != T.assignment;
            q = 32;
        }
    }
    L.end[0] = L.end[1] = 0; L.chk0 = 0;
    for (uint32_t k = 0; k < T.nss; k++) {
        if (q + 2 > L.total) return L;
        const uint32_t b0 = b(q);
        L.end[k] = (((b0 & 15u) << 8) | b(q + 1)) * 2;
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

__device__ AuLayout au_layout(const uint8_t *es, uint64_t pos, const TrackDev &T)
{
    const uint8_t *p = es + pos;
    return au_layout_from([p](uint32_t i) { return ld_u8(p + i); }, T);
}

// One thread per access unit: parity and CRC-8 of each substream
// (mlp.c:670-712, 1360-1399: parity over all bytes but the last two, CRC-8
// (poly 0x63, start 0x3C) over the same bytes where the check byte is compared
// with the value *in front of* the last table step).  au_err: 0 ok, 1 = drop
// silently (stream parameters changed, mlp.c:452-455), else ERR_* bits.
//
// The access units of a warp's 32 lanes follow each other in the elementary stream, so the
// warp first copies their common byte range into its own window of shared memory (16-byte
// asynchronous copies, coalesced: every byte of the stream crosses L1 once and DRAM traffic is
// the stream itself), then every lane walks its own access unit there.  Lanes whose access
// unit does not fit the window behind the first pending one wait for the next round.  A lane
// works on four runs of 16-byte pieces at once (four independent table chains, four bytes per
// step each) and joins them by carrying each run's state over the length of what follows it.
// Tables: [0..3] "byte followed by k zero bytes", [4..11] "state carried over 16 << k zero bytes".
__device__ __align__(16) uint8_t g_chk_tab[12][256];
#define CHK_WARPS 4
#define CHK_THREADS (CHK_WARPS * 32)
#define CHK_WINDOW 16384          // bytes per warp; an access unit has at most 8190
#define CHK_TAB_BYTES (12 * 256)
#define CHK_SMEM_BYTES (CHK_TAB_BYTES + CHK_WARPS * CHK_WINDOW)
__device__ __forceinline__ void chk_cp16(uint32_t smem, const void *g)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem), "l"(g) : "memory");
}
__global__ void __launch_bounds__(CHK_THREADS) k_checkdata(MlpTables m, const uint32_t *__restrict__ seg_au_base)
{
    extern __shared__ __align__(16) uint8_t chk_sm[];
    uint8_t (*T)[256] = reinterpret_cast<uint8_t (*)[256]>(chk_sm);
    uint8_t (*ADV)[256] = T + 4;
    for (uint32_t i = threadIdx.x; i < CHK_TAB_BYTES / 16; i += CHK_THREADS)
        reinterpret_cast<uint4 *>(chk_sm)[i] = reinterpret_cast<const uint4 *>(&g_chk_tab[0][0])[i];
    __syncthreads();
#define CHK_STEP(c_, wv)                                                                                   \
    {                                                                                                      \
        const uint32_t w_ = (wv);                                                                          \
        pw ^= w_;                                                                                          \
        c_ = T[3][(c_ ^ w_) & 0xFF] ^ T[2][(w_ >> 8) & 0xFF] ^ T[1][(w_ >> 16) & 0xFF] ^ T[0][w_ >> 24];   \
    }
#define CHK_WORD(wv) CHK_STEP(crc, wv)
    // the state after `blocks` * 16 more zero bytes
    auto carry = [&](uint32_t c, uint32_t blocks) {
        for (uint32_t k = 0; blocks; k++, blocks >>= 1)
            if (blocks & 1) c = ADV[k][c];
        return c;
    };
    const uint32_t lane = threadIdx.x & 31;
    uint8_t *const win = chk_sm + CHK_TAB_BYTES + (threadIdx.x >> 5) * CHK_WINDOW;
    const uint32_t win_s = (uint32_t)__cvta_generic_to_shared(win);
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    const bool have = a < m.cnt->nau;
    uint64_t pos = 0;
    uint32_t total = 0;
    if (have) {
        pos = m.au_pos[a];
        total = max(2u, (((ld_u8(m.es + pos) & 15u) << 8) | ld_u8(m.es + pos + 1)) * 2);
    }
    uint32_t pending = __ballot_sync(0xFFFFFFFFu, have);
    while (pending) {
        const int first = __ffs(pending) - 1;
        const uint64_t base = __shfl_sync(0xFFFFFFFFu, pos, first) & ~15ull;
        const bool fits = ((pending >> lane) & 1) && pos >= base && pos + total <= base + CHK_WINDOW;
        const uint32_t now = __ballot_sync(0xFFFFFFFFu, fits);        // the first pending lane is always in
        const uint32_t span = __reduce_max_sync(0xFFFFFFFFu, fits ? (uint32_t)(pos + total - base) : 0u);
        const uint32_t n16 = (span + 15) >> 4;
        const uint8_t *src = m.es + base;                             // (the pad behind the stream covers the round-up)
        for (uint32_t i = lane; i < n16; i += 32) chk_cp16(win_s + i * 16, src + (uint64_t)i * 16);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        if (fits) {
            const TrackDev &Tr = m.tracks[m.segs[m.au_seg[a]].track];
            const uint8_t *au = win + (uint32_t)(pos - base);
            const AuLayout L = au_layout_from([au](uint32_t i) { return (uint32_t)au[i]; }, Tr);
            uint32_t err = 0;
            if (L.has_sync && L.params_differ) err = 1;
            else if (!L.ok) err = ERR_SYNTAX;
            else if (L.chk0) {
                for (uint32_t k = 0; k < Tr.nss && !err; k++) {
                    const uint32_t start = k ? L.end[0] : 0;
                    const uint8_t *p = au + L.data0 + start;          // same 16-byte phase as in the stream
                    const uint32_t n = L.end[k] - start - 2;          // bytes covered
                    uint32_t parity = 0, crc = 0x3C, fin = 0;
                    if (n) {
                        // all bytes but the last advance the CRC; the last one only forms `fin`
                        const uint32_t body = n - 1;
                        uint32_t i = 0;
                        const uint32_t head = min(body, (uint32_t)((16 - ((uintptr_t)p & 15)) & 15));
                        for (; i < head; i++) { const uint32_t b = p[i]; parity ^= b; crc = T[0][crc ^ b]; }
                        uint32_t pw = 0;
                        const uint32_t q = (body - i) >> 6;
                        if (q) {
                            uint32_t c1 = 0, c2 = 0, c3 = 0;
                            const uint4 *r0 = reinterpret_cast<const uint4 *>(p + i);
                            const uint4 *r1 = r0 + q, *r2 = r1 + q, *r3 = r2 + q;
                            for (uint32_t j = 0; j < q; j++) {
                                const uint4 v0 = r0[j], v1 = r1[j], v2 = r2[j], v3 = r3[j];
                                CHK_STEP(crc, v0.x) CHK_STEP(c1, v1.x) CHK_STEP(c2, v2.x) CHK_STEP(c3, v3.x)
                                CHK_STEP(crc, v0.y) CHK_STEP(c1, v1.y) CHK_STEP(c2, v2.y) CHK_STEP(c3, v3.y)
                                CHK_STEP(crc, v0.z) CHK_STEP(c1, v1.z) CHK_STEP(c2, v2.z) CHK_STEP(c3, v3.z)
                                CHK_STEP(crc, v0.w) CHK_STEP(c1, v1.w) CHK_STEP(c2, v2.w) CHK_STEP(c3, v3.w)
                            }
                            crc = carry(crc, q) ^ c1;
                            crc = carry(crc, q) ^ c2;
                            crc = carry(crc, q) ^ c3;
                            i += q * 64;
                        }
                        for (; i + 16 <= body; i += 16) {
                            const uint4 v0 = *reinterpret_cast<const uint4 *>(p + i);
                            CHK_WORD(v0.x) CHK_WORD(v0.y) CHK_WORD(v0.z) CHK_WORD(v0.w)
                        }
                        parity ^= (pw ^ (pw >> 8) ^ (pw >> 16) ^ (pw >> 24)) & 0xFF;
                        for (; i < body; i++) { const uint32_t b = p[i]; parity ^= b; crc = T[0][crc ^ b]; }
                        const uint32_t last = p[body];
                        parity ^= last;
                        fin = crc ^ last;
                    }
                    if (((p[n] ^ parity) & 0xFF) != 0xA9) err = ERR_PARITY;
                    else if (p[n + 1] != fin) err = ERR_CRC;
                }
            }
            m.au_err[a] = (uint8_t)err;
        }
        pending &= ~now;
        __syncwarp();                                                 // the window is overwritten by the next round
    }
#undef CHK_WORD
#undef CHK_STEP
}

// The same check with every lane reading its access unit straight from the elementary stream,
// 16 bytes at a time: for large access units (few of them fit a window, so the windowed kernel
// above runs many rounds with most lanes idle) this is the faster one.
#define CHKD_THREADS 128
__global__ void __launch_bounds__(CHKD_THREADS) k_checkdata_direct(MlpTables m, const uint32_t *__restrict__ seg_au_base)
{
    __shared__ uint8_t T[4][256];
    for (uint32_t i = threadIdx.x; i < 256; i += CHKD_THREADS) {
        const uint8_t t0 = c_crc8[i], t1 = c_crc8[t0], t2 = c_crc8[t1];
        T[0][i] = t0; T[1][i] = t1; T[2][i] = t2; T[3][i] = c_crc8[t2];
    }
    __syncthreads();
#define CHK_WORD(wv)                                                                                       \
    {                                                                                                      \
        const uint32_t w_ = (wv);                                                                          \
        pw ^= w_;                                                                                          \
        crc = T[3][(crc ^ w_) & 0xFF] ^ T[2][(w_ >> 8) & 0xFF] ^ T[1][(w_ >> 16) & 0xFF] ^ T[0][w_ >> 24]; \
    }
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= m.cnt->nau) return;
    const uint32_t si = upper_bound_dev(seg_au_base, m.cnt->nseg, a) - 1;
    const TrackDev &Tr = m.tracks[m.segs[si].track];
    const uint64_t pos = m.au_pos[a];
    const AuLayout L = au_layout(m.es, pos, Tr);
    uint32_t err = 0;
    if (L.has_sync && L.params_differ) err = 1;
    else if (!L.ok) err = ERR_SYNTAX;
    else if (L.chk0) {
        for (uint32_t k = 0; k < Tr.nss && !err; k++) {
            const uint32_t start = k ? L.end[0] : 0;
            const uint8_t *p = m.es + pos + L.data0 + start;
            const uint32_t n = L.end[k] - start - 2;          // bytes covered
            uint32_t parity = 0, crc = 0x3C, fin = 0;
            if (n) {
                // all bytes but the last advance the CRC; the last one only forms `fin`
                const uint32_t body = n - 1;
                uint32_t i = 0;
                const uint32_t head = min(body, (uint32_t)((16 - ((uintptr_t)p & 15)) & 15));
                for (; i < head; i++) { const uint32_t b = ld_u8(p + i); parity ^= b; crc = T[0][crc ^ b]; }
                uint32_t pw = 0;
                for (; i + 32 <= body; i += 32) {
                    const uint4 v0 = __ldg(reinterpret_cast<const uint4 *>(p + i));
                    const uint4 v1 = __ldg(reinterpret_cast<const uint4 *>(p + i + 16));
                    CHK_WORD(v0.x) CHK_WORD(v0.y) CHK_WORD(v0.z) CHK_WORD(v0.w)
                    CHK_WORD(v1.x) CHK_WORD(v1.y) CHK_WORD(v1.z) CHK_WORD(v1.w)
                }
                for (; i + 16 <= body; i += 16) {
                    const uint4 v0 = __ldg(reinterpret_cast<const uint4 *>(p + i));
                    CHK_WORD(v0.x) CHK_WORD(v0.y) CHK_WORD(v0.z) CHK_WORD(v0.w)
                }
                parity ^= (pw ^ (pw >> 8) ^ (pw >> 16) ^ (pw >> 24)) & 0xFF;
                for (; i < body; i++) { const uint32_t b = ld_u8(p + i); parity ^= b; crc = T[0][crc ^ b]; }
                const uint32_t last = ld_u8(p + body);
                parity ^= last;
                fin = crc ^ last;
            }
            if (((ld_u8(p + n) ^ parity) & 0xFF) != 0xA9) err = ERR_PARITY;
            else if (ld_u8(p + n + 1) != fin) err = ERR_CRC;
        }
    }
    m.au_err[a] = (uint8_t)err;
#undef CHK_WORD
}

__global__ void k_huff_lut_build();       // further down, next to the table it fills

int upload_crc_table(const uint8_t *t)
{
    CUDA_TRY(cudaMemcpyToSymbol(c_crc8, t, 256));
    k_huff_lut_build<<<1, 256>>>();
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaDeviceSynchronize());
    static uint8_t tab[12][256];
    for (int i = 0; i < 256; i++) {
        tab[0][i] = t[i];
        for (int k = 1; k < 4; k++) tab[k][i] = t[tab[k - 1][i]];
        uint8_t c = (uint8_t)i;
        for (int z = 0; z < 16; z++) c = t[c];               // a zero byte: state -> t[state]
        tab[4][i] = c;
    }
    for (int k = 5; k < 12; k++)
        for (int i = 0; i < 256; i++) tab[k][i] = tab[k - 1][tab[k - 1][i]];
    CUDA_TRY(cudaMemcpyToSymbol(g_chk_tab, tab, sizeof tab));
    return 0;
}

// access units of up to this many bytes on average go through the shared-memory windows
#define CHK_WINDOWED_MAX_AU 512
// windowed: access units of up to CHK_WINDOWED_MAX_AU bytes on average (the host decides from
// what it knows of the stream: both kernels are right for any input)
int launch_checkdata(MlpTables m, const uint32_t *seg_au_base, bool windowed, cudaStream_t s)
{
    if (!m.cap_au) return 0;
    if (!windowed) {
        LAUNCH(k_checkdata_direct, div_up_u32(m.cap_au, CHKD_THREADS), CHKD_THREADS, 0, s, m, seg_au_base);
        CUDA_TRY(cudaGetLastError());
        return 0;
    }
    static PerDeviceOnce attr_once;
    if (attr_once.run([&]() -> int { CUDA_TRY(cudaFuncSetAttribute(k_checkdata, cudaFuncAttributeMaxDynamicSharedMemorySize, CHK_SMEM_BYTES)); return 0; })) return -1;
    LAUNCH(k_checkdata, div_up_u32(m.cap_au, CHK_THREADS), CHK_THREADS, CHK_SMEM_BYTES, s, m, seg_au_base);
    CUDA_TRY(cudaGetLastError());
    return 0;
}


// ---- reader for the header-only passes.  The 128 bytes behind a seat are fetched with
// eight independent 16-byte loads into the thread's column of a shared-memory window
// (one memory latency for a whole parameter block instead of one per 32-byte sector);
// words beyond the window come straight from global memory.
#define GRD_WIN_WORDS 32
#define GRD_THREADS 128
struct GRd {
    const uint8_t *es;
    uint32_t *col;              // this thread's column of the window: word i at col[i * GRD_THREADS]
    uint32_t win_w0;            // absolute word index of window word 0
    uint64_t win;
    int32_t avail;
    uint32_t next_w, base_w;
};
// fetch the window: the 16-byte aligned 128 bytes from byte_pos on
__device__ __forceinline__ void grd_stage(GRd &r, const uint8_t *es, uint32_t *col, uint64_t byte_pos)
{
    r.es = es; r.col = col;
    const uint64_t base = byte_pos & ~15ull;
    r.win_w0 = (uint32_t)(base >> 2);
    const uint4 *src = reinterpret_cast<const uint4 *>(es + base);
    uint4 v[GRD_WIN_WORDS / 4];
#pragma unroll
    for (int i = 0; i < GRD_WIN_WORDS / 4; i++) v[i] = __ldg(src + i);
#pragma unroll
    for (int i = 0; i < GRD_WIN_WORDS / 4; i++) {
        col[(4 * i + 0) * GRD_THREADS] = v[i].x; col[(4 * i + 1) * GRD_THREADS] = v[i].y;
        col[(4 * i + 2) * GRD_THREADS] = v[i].z; col[(4 * i + 3) * GRD_THREADS] = v[i].w;
    }
}
// word w (absolute index), as stored (little endian)
__device__ __forceinline__ uint32_t grd_word(const GRd &r, uint32_t w)
{
    const uint32_t i = w - r.win_w0;
    return i < GRD_WIN_WORDS ? r.col[i * GRD_THREADS] : __ldg(reinterpret_cast<const uint32_t *>(r.es) + w);
}
__device__ __forceinline__ uint32_t grd_byte(const GRd &r, uint64_t byte_pos)
{
    return (grd_word(r, (uint32_t)(byte_pos >> 2)) >> (8 * (byte_pos & 3))) & 0xFF;
}
__device__ __forceinline__ void grd_seat(GRd &r, uint64_t byte_pos)
{
    r.next_w = r.base_w = (uint32_t)(byte_pos >> 2); r.win = 0; r.avail = 0;
}
// (out of line, arguments by value: the header parsers pull at some sixty places, and the size
// of their code is what limits them)
__device__ __noinline__ uint32_t grd_fetch_be(const uint32_t *col, uint32_t win_w0, const uint8_t *es, uint32_t w)
{
    const uint32_t i = w - win_w0;
    const uint32_t v = i < GRD_WIN_WORDS ? col[i * GRD_THREADS] : __ldg(reinterpret_cast<const uint32_t *>(es) + w);
    return __byte_perm(v, 0, 0x0123);
}
__device__ __forceinline__ void rd_pull(GRd &r)
{
    const uint32_t word = grd_fetch_be(r.col, r.win_w0, r.es, r.next_w);
    r.win |= (uint64_t)word << (32 - r.avail);
    r.avail += 32;
    r.next_w++;
}


// ------------------------------------------------------------ decoder state

struct ChanState {
    int32_t fir_c[8], iir_c[8];
    int32_t fst[8], ist[8];          // circular histories (generic path; IIR state as transmitted)
    uint8_t fir_order, iir_order, fir_shift, iir_shift;
    uint8_t fhead, ihead;            // next write slot
    uint8_t flen, ilen;              // valid entries (<= 8)
    uint8_t ist_new;                 // IIR history was replaced by a parameter block
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

// Huffman LUT [codebook 0..3][9 peeked bits] -> value | length << 8; 0xFFFF = invalid code.
// Built from the prefix structure of the three codebooks
// (src/mlp_codebook{1,2,3}.json): 0^z 1 -> 8 - z; 1 + literal -> 7 + literal;
// 01 0^k 1 -> hi_base + k.
__device__ uint16_t huff_entry(uint32_t cb, uint32_t v9)
{
    if (cb == 0) return 0;                               // codebook 0: no code, MSB = 0
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

// The same table for every block: built once per device (upload_crc_table), copied into shared
// memory by the kernels with 16-byte loads.
__device__ __align__(16) uint16_t g_huff_lut[4][512];
__global__ void k_huff_lut_build()
{
    for (uint32_t i = threadIdx.x; i < 4 * 512; i += blockDim.x) g_huff_lut[i >> 9][i & 511] = huff_entry(i >> 9, i & 511);
}
__device__ __forceinline__ void huff_lut_to_shared(uint16_t (*lut)[512])
{
    const uint4 *src = reinterpret_cast<const uint4 *>(&g_huff_lut[0][0]);
    uint4 *dst = reinterpret_cast<uint4 *>(&lut[0][0]);
    for (uint32_t i = threadIdx.x; i < 4 * 512 * 2 / 16; i += blockDim.x) dst[i] = __ldg(src + i);
}

struct DecodeJob {
    uint32_t seg;          // global segment index
    uint32_t k;            // substream
    uint32_t lane;         // column in the group's tile
    bool exact_history;    // FIR history at segment start is the true one
};

// ---- parameter parsing (cold path) ------------------------------------------

template <typename RD>
__device__ bool restart_header(RD &b, SubState &s)
{
    const uint32_t sync = rd_get(b, 13), noise_type = rd_get(b, 1);
    rd_skip(b, 16);
    s.min_ch = rd_get(b, 4); s.max_ch = rd_get(b, 4); s.mmc = rd_get(b, 4);
    s.noise_shift = rd_get(b, 4);
    s.seed = rd_get(b, 23);
    rd_skip(b, 19 + 1 + 8 + 16);
    if (sync != 0x18F5 || noise_type != 0) return false;
    if (s.max_ch < s.min_ch || s.mmc < s.max_ch || s.mmc >= DVDA_MAX_CH) return false;
    for (uint32_t c = 0; c <= s.mmc; c++) if (rd_get(b, 6) > s.mmc) return false;
    rd_skip(b, 8);
    s.have_header = 1;
    s.dirty = 1;
    return true;
}

template <typename RD>
__device__ bool filter_params(RD &b, ChanState &C, bool iir)
{
    const uint32_t order = rd_get(b, 4);
    if (order > 8) return false;
    if (!order) {
        if (iir) { C.iir_order = 0; C.iir_shift = 0; C.ilen = 0; C.ihead = 0; C.ist_new = 1; }
        else { C.fir_order = 0; C.fir_shift = 0; }
        return true;
    }
    const uint32_t shift = rd_get(b, 4), bits = rd_get(b, 5);
    if (bits < 1 || bits > 16) return false;
    const uint32_t cshift = rd_get(b, 3);
    if (bits + cshift > 16) return false;
    int32_t *coef = iir ? C.iir_c : C.fir_c;
    for (uint32_t i = 0; i < order; i++) coef[i] = (int32_t)((uint32_t)rd_get_s(b, bits) << cshift);
    if (!iir) {
        C.fir_order = order; C.fir_shift = shift;
        if (rd_get(b, 1)) return false;
    } else {
        C.iir_order = order; C.iir_shift = shift;
        C.ilen = 0; C.ihead = 0; C.ist_new = 1;
        if (rd_get(b, 1)) {
            const uint32_t sbits = rd_get(b, 4), sshift = rd_get(b, 4);
            if (!sbits) return false;                           // reference underflows (G2)
            // first value sent pairs with coeff[0] = most recent (mlp.c:1103-1107):
            // store so that reading backwards from ihead yields sent[0], sent[1], ...
            for (uint32_t i = 0; i < order; i++)
                C.ist[(order - 1 - i) & 7] = (int32_t)((uint32_t)rd_get_s(b, sbits) << sshift);
            C.ilen = order; C.ihead = order & 7;
        }
    }
    return true;
}

template <typename RD>
__device__ bool decoding_params(RD &b, SubState &s, bool restart)
{
    if (restart) {
        if (rd_get(b, 1)) { uint32_t f = 0; for (int k = 0; k < 8; k++) f |= rd_get(b, 1) << k; s.flags = f; }
        else s.flags = 0xFF;
    } else if ((s.flags & 1) && rd_get(b, 1)) {
        uint32_t f = 0; for (int k = 0; k < 8; k++) f |= rd_get(b, 1) << k; s.flags = f;
    }
    if ((s.flags & 0x80) && rd_get(b, 1)) {
        s.block_size = rd_get(b, 9);
        if (s.block_size < 8) return false;
    } else if (restart) s.block_size = 8;

    if ((s.flags & 0x40) && rd_get(b, 1)) {
        s.dirty = 1;
        s.matrix_len = rd_get(b, 4);
        if (s.matrix_len > DVDA_MAX_MAT || s.mmc + 3 > DVDA_MAX_CH) return false;
        for (uint32_t m = 0; m < s.matrix_len; m++) {
            if ((s.mat_out[m] = rd_get(b, 4)) > s.mmc) return false;
            const uint32_t frac = rd_get(b, 4);
            if (frac > 14) return false;
            s.mat_bypass[m] = rd_get(b, 1);
            for (uint32_t c = 0; c < DVDA_MAX_CH; c++) s.coeff[m][c] = 0;
            for (uint32_t c = 0; c < (uint32_t)s.mmc + 3; c++)
                if (rd_get(b, 1)) s.coeff[m][c] = (int16_t)((uint32_t)rd_get_s(b, frac + 2) << (14 - frac));
        }
    } else if (restart) { s.matrix_len = 0; s.dirty = 1; }

    if ((s.flags & 0x20) && rd_get(b, 1)) {
        s.dirty = 1;
        for (uint32_t c = 0; c <= s.mmc; c++) s.out_shift[c] = (uint8_t)(rd_get_s(b, 4) & 31);
    } else if (restart) { for (int c = 0; c < DVDA_MAX_CH; c++) s.out_shift[c] = 0; s.dirty = 1; }

    if ((s.flags & 0x10) && rd_get(b, 1)) {
        s.dirty = 1;
        for (uint32_t c = 0; c <= s.max_ch; c++) s.q[c] = rd_get(b, 4);
    } else if (restart) { for (int c = 0; c < DVDA_MAX_CH; c++) s.q[c] = 0; s.dirty = 1; }

    for (uint32_t c = s.min_ch; c <= s.max_ch; c++) {
        ChanState &C = s.ch[c];
        if (rd_get(b, 1)) {
            if ((s.flags & 0x08) && rd_get(b, 1)) { if (!filter_params(b, C, false)) return false; }
            else if (restart) { C.fir_order = 0; C.fir_shift = 0; }
            if ((s.flags & 0x04) && rd_get(b, 1)) { if (!filter_params(b, C, true)) return false; }
            else if (restart) { C.iir_order = 0; C.iir_shift = 0; C.ilen = 0; C.ihead = 0; C.ist_new = 1; }
            if ((s.flags & 0x02) && rd_get(b, 1)) C.huff_offset = rd_get_s(b, 15);
            else if (restart) C.huff_offset = 0;
            C.codebook = rd_get(b, 2);
            C.huff_lsbs = rd_get(b, 5);
            if (C.huff_lsbs > 24) return false;
        } else if (restart) {
            C.fir_order = 0; C.fir_shift = 0;
            C.iir_order = 0; C.iir_shift = 0; C.ilen = 0; C.ihead = 0; C.ist_new = 1;
            C.huff_offset = 0; C.codebook = 0; C.huff_lsbs = 24;
        }
    }
    return true;
}

// block header: optional restart header + decoding parameters (mlp.c:749-771)
template <typename RD>
__device__ __forceinline__ bool block_header(RD &b, SubState &s, bool &changed)
{
    changed = false;
    if (rd_get(b, 1)) {
        const bool restart = rd_get(b, 1);
        if (restart && !restart_header(b, s)) return false;
        if (!s.have_header) return false;
        if (!decoding_params(b, s, restart)) return false;
        changed = true;
    }
    return s.have_header;
}

// per-channel constants of one block (mlp.c:1151-1176, 1260-1270); false = syntax error
__device__ __forceinline__ bool channel_setup(const SubState &s, const ChanState &C, uint32_t q, bool exact_history,
                                              uint32_t &flags, uint32_t &lsb_bits, int32_t &sho, uint32_t &shift)
{
    if (C.huff_lsbs < q) return false;
    const uint32_t nb = C.huff_lsbs - q;
    lsb_bits = nb;
    if (C.codebook) {
        const int ss = (int)nb + 2 - (int)C.codebook;
        sho = C.huff_offset - 7 * (1 << nb) - (ss >= 0 ? (1 << ss) : 0);
    } else {
        sho = C.huff_offset - (nb >= 1 ? (1 << (nb - 1)) : 0);
    }
    if (C.fir_order + C.iir_order > 8) return false;
    if (C.fir_shift > 0 && C.iir_shift > 0 && C.fir_shift != C.iir_shift) return false;
    shift = (C.fir_shift > 0 && C.iir_shift > 0) ? C.fir_shift : C.fir_order > 0 ? C.fir_shift : C.iir_shift;
    if (C.iir_order > C.ilen) return false;                     // reference reads out of bounds (G2)
    if (C.fir_order > C.flen) {
        if (exact_history) return false;                        // reference reads out of bounds (G1)
        flags |= SEG_NEEDS_CARRY;                               // history lives in the previous segment
    }
    (void)s;
    return true;
}


// ---- one block, generic: any channel count, histories in local memory --------

// returns frames decoded, 0 on a syntax error
__device__ uint32_t decode_block_generic(const MlpTables &m, const GroupDev &G, const DecodeJob &job,
                                         SubState &s, Rd &b, uint32_t end_bits, uint32_t frame0,
                                         uint32_t nch, bool governing, const uint16_t (*lut)[512], uint32_t &flags)
{
    bool changed;
    if (!block_header(b, s, changed)) return 0;
    if (rd_pos(b) > end_bits) return 0;

    const uint32_t n = s.block_size;
    int32_t sho[DVDA_MAX_CH];
    uint32_t lsb_bits[DVDA_MAX_CH], shift[DVDA_MAX_CH];
    for (uint32_t c = s.min_ch; c <= s.max_ch; c++) {
        if (!channel_setup(s, s.ch[c], s.q[c], job.exact_history, flags, lsb_bits[c], sho[c], shift[c])) return 0;
        s.ch[c].ist_new = 0;
    }
    uint32_t want = 0;
    for (uint32_t k = 0; k < s.matrix_len; k++) want |= (uint32_t)(s.mat_bypass[k] != 0) << k;

    int32_t *tile = m.tiles + G.tile_off + job.lane;
    uint8_t *byp = m.bypass + G.byp_off + job.lane;
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t f = frame0 + i;
        const bool room = f < G.cap;
        if (!room) flags |= SEG_OVERFLOW;
        const uint32_t bmask = bypass_bits(b, want);
        if (governing && room) byp[(uint64_t)f * DVDA_LANES] = (uint8_t)bmask;
        for (uint32_t c = s.min_ch; c <= s.max_ch; c++) {
            ChanState &C = s.ch[c];
            int32_t msb = 0;
            if (C.codebook) {
                const uint32_t e = lut[C.codebook][rd_peek(b, 9)];
                if (e == 0xFFFF) return 0;
                msb = e & 0xFF;
                rd_drop(b, e >> 8);
            }
            const int32_t lsb = (int32_t)rd_get(b, lsb_bits[c]);
            const int32_t res = (int32_t)((uint32_t)((msb << lsb_bits[c]) + lsb + sho[c]) << s.q[c]);
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
        if (rd_pos(b) > end_bits) return 0;
    }
    return n;
}

// ---- one block, fast: NCH channels, filter state in registers ----------------
//
// Histories are kept as 8 registers per channel and filter.  The frame loop is
// unrolled by 8 so that "age a at frame j" is register (a - j) mod 8 with both
// indices known at compile time: no register moves, no local memory.  Taps
// beyond the transmitted orders carry zero coefficients, so every lane runs the
// same 16 multiply-adds whatever its filter orders are (no divergence).


template <int NCH>
struct Hot {
    int32_t fh[NCH][8], ih[NCH][8];      // histories, [0] = most recent at chunk boundaries
    int32_t cf[NCH][8], ci[NCH][8];      // coefficients, zero beyond the order
};

template <int NCH>
__device__ __forceinline__ void hot_load_params(Hot<NCH> &H, SubState &s)
{
#pragma unroll
    for (int cc = 0; cc < NCH; cc++) {
        ChanState &C = s.ch[s.min_ch + cc];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            H.cf[cc][j] = j < C.fir_order ? C.fir_c[j] : 0;
            H.ci[cc][j] = j < C.iir_order ? C.iir_c[j] : 0;
        }
        if (C.ist_new) {
            // IIR history replaced by the transmitted state (or emptied)
#pragma unroll
            for (int a = 0; a < 8; a++) H.ih[cc][a] = a < C.ilen ? C.ist[(C.ihead - 1 - a) & 7] : 0;
            C.ist_new = 0;
        }
    }
}

template <int NCH>
__device__ __forceinline__ uint32_t decode_block_fast(const MlpTables &m, const GroupDev &G, const DecodeJob &job,
                                      SubState &s, Hot<NCH> &H, Rd &b, uint32_t end_bits, uint32_t frame0,
                                      uint32_t nch, bool governing, const uint16_t (*lut)[512], uint32_t &flags)
{
    bool changed;
    if (!block_header(b, s, changed)) return 0;
    if (rd_pos(b) > end_bits) return 0;
    if ((uint32_t)(s.max_ch - s.min_ch + 1) != NCH) return 0xFFFFFFFFu;   // not the expected channel split: generic path
    if (changed) hot_load_params(H, s);

    const uint32_t n = s.block_size;
    int32_t sho[NCH];
    uint32_t lsb_bits[NCH], shift[NCH], cb[NCH], q[NCH];
#pragma unroll
    for (int cc = 0; cc < NCH; cc++) {
        ChanState &C = s.ch[s.min_ch + cc];
        q[cc] = s.q[s.min_ch + cc];
        cb[cc] = C.codebook;
        if (!channel_setup(s, C, q[cc], job.exact_history, flags, lsb_bits[cc], sho[cc], shift[cc])) return 0;
        // histories are full after any block (blocks hold >= 8 frames)
        C.flen = 8; C.ilen = 8;
    }
    uint32_t want = 0;
    for (uint32_t k = 0; k < s.matrix_len; k++) want |= (uint32_t)(s.mat_bypass[k] != 0) << k;

    int32_t *tile = m.tiles + G.tile_off + job.lane + ((uint64_t)frame0 * nch + s.min_ch) * DVDA_LANES;
    uint8_t *byp = m.bypass + G.byp_off + job.lane + (uint64_t)frame0 * DVDA_LANES;
    const uint32_t tile_step = nch * DVDA_LANES;
    const uint32_t cap = G.cap;
    uint32_t f = frame0, over = 0;

    // one frame with the histories rotated by J (compile time); no data-dependent
    // branch: invalid codes are collected in `bad` and looked at once per 8 frames
    uint32_t bad = 0;
#define DVDA_FRAME(J)                                                                              \
    {                                                                                              \
        const bool room = f < cap;                                                                 \
        over |= room ? 0u : SEG_OVERFLOW;                                                          \
        if (want) {                                                                                \
            const uint32_t bmask = bypass_bits(b, want);                                           \
            rd_hot_begin(b);                      /* the checked reader moved next_w */             \
            if (governing && room) *byp = (uint8_t)bmask;                                          \
        } else if (governing && room) *byp = 0;                                                    \
        _Pragma("unroll") for (int cc = 0; cc < NCH; cc++) {                                       \
            rd_top_up(b);                                                                          \
            const uint32_t e = lut[cb[cc]][(uint32_t)(b.win >> 55)];                               \
            bad |= e;                                                                              \
            const uint32_t hl = (e >> 8) & 15;                                                     \
            const int32_t msb = e & 0xFF;                                                          \
            b.win <<= hl;                                                                          \
            const int32_t lsb = (int32_t)(uint32_t)((b.win >> 1) >> (63 - lsb_bits[cc]));          \
            b.win <<= lsb_bits[cc];                                                                \
            b.avail -= hl + lsb_bits[cc];                                                          \
            const int32_t res = (int32_t)((uint32_t)((msb << lsb_bits[cc]) + lsb + sho[cc]) << q[cc]); \
            long long s0 = 0, s1 = 0;                                                              \
            /* oldest taps first: only the last multiply-add waits for the previous sample */      \
            _Pragma("unroll") for (int a = 7; a >= 0; a--) {                                       \
                s0 += (long long)H.cf[cc][a] * H.fh[cc][(a - (J)) & 7];                            \
                s1 += (long long)H.ci[cc][a] * H.ih[cc][(a - (J)) & 7];                            \
            }                                                                                      \
            const long long sum = s0 + s1;                                                         \
            const int32_t ssum = (int32_t)(sum >> shift[cc]);                                      \
            int32_t v = (int32_t)((uint32_t)ssum + (uint32_t)res);                                 \
            v = (v >> q[cc]) << q[cc];                                                             \
            H.fh[cc][(7 - (J)) & 7] = v;                                                           \
            H.ih[cc][(7 - (J)) & 7] = (int32_t)((uint32_t)v - (uint32_t)ssum);                     \
            if (room) tile[cc * DVDA_LANES] = v;                                                   \
        }                                                                                          \
        f++; tile += tile_step; byp += DVDA_LANES;                                                 \
    }

    // words 8 frames can consume at most (33 bits per sample, 6 bypass bits per frame), plus
    // the word fetched ahead
    const uint32_t need8 = (4 * (NCH * 33 + 6) + 31) / 32 + 2;      // per four frames (the ring is small)
    uint32_t i = 0;
#if DVDA_UNROLL == 8
    for (; i + 8 <= n; i += 8) {
        rd_prefetch(b, need8);
        rd_hot_begin(b);
        DVDA_FRAME(0) DVDA_FRAME(1) DVDA_FRAME(2) DVDA_FRAME(3)
        rd_prefetch(b, need8);
        rd_hot_begin(b);
        DVDA_FRAME(4) DVDA_FRAME(5) DVDA_FRAME(6) DVDA_FRAME(7)
        if ((bad & 0x8000) || rd_pos(b) > end_bits) return 0;
    }
#else
    // half the code: four frames with static rotation, then swap the register halves
    for (; i + 4 <= n; i += 4) {
        rd_prefetch(b, need8);
        rd_hot_begin(b);
        DVDA_FRAME(0) DVDA_FRAME(1) DVDA_FRAME(2) DVDA_FRAME(3)
#pragma unroll
        for (int cc = 0; cc < NCH; cc++) {
#pragma unroll
            for (int a = 0; a < 4; a++) {
                const int32_t tf = H.fh[cc][a], ti = H.ih[cc][a];
                H.fh[cc][a] = H.fh[cc][a + 4]; H.ih[cc][a] = H.ih[cc][a + 4];
                H.fh[cc][a + 4] = tf; H.ih[cc][a + 4] = ti;
            }
        }
        if ((bad & 0x8000) || rd_pos(b) > end_bits) return 0;
    }
#endif
    // leftover frames (block size not a multiple of 8): rotate the registers for real
    for (; i < n; i++) {
        rd_prefetch(b, need8);
        rd_hot_begin(b);
        DVDA_FRAME(0)
#pragma unroll
        for (int cc = 0; cc < NCH; cc++) {
            const int32_t nf = H.fh[cc][7], ni = H.ih[cc][7];
#pragma unroll
            for (int a = 7; a > 0; a--) { H.fh[cc][a] = H.fh[cc][a - 1]; H.ih[cc][a] = H.ih[cc][a - 1]; }
            H.fh[cc][0] = nf; H.ih[cc][0] = ni;
        }
        if ((bad & 0x8000) || rd_pos(b) > end_bits) return 0;
    }
#undef DVDA_FRAME
    flags |= over;
    return n;
}


// access-unit header through the reader: "4p 12u 16p", optional major sync,
// substream directory (mlp.c:392-394, 614-668).  Leaves the reader anywhere.
__device__ AuLayout au_layout_rd(Rd &b, uint64_t pos, const TrackDev &T)
{
    AuLayout L;
    L.ok = false; L.has_sync = false; L.params_differ = false;
    rd_seat(b, pos);
    rd_skip(b, (uint32_t)(pos & 3) * 8);
    L.total = ((rd_get(b, 32) >> 16) & 0xFFF) * 2;
    uint32_t q = 4;
    if (L.total >= 32 && rd_peek(b, 32) == 0xF8726FBBu) {
        rd_drop(b, 32);
        const uint32_t w8 = rd_get(b, 32);                 // bytes 8..11: formats, 11 skipped bits, assignment
        rd_skip(b, 64);                                    // bytes 12..19
        const uint32_t ns = rd_get(b, 4);                  // byte 20, high nibble
        if (ns == 1 || ns == 2) {
            L.has_sync = true;
            L.params_differ = (w8 >> 28) != T.g0_bps || ((w8 >> 24) & 15) != T.g1_bps || ((w8 >> 20) & 15) != T.g0_rate ||
                              ((w8 >> 16) & 15) != T.g1_rate || (w8 & 31) != T.assignment;
            rd_skip(b, 92);
            q = 32;
        } else {
            // not a major sync after all: the bytes are the directory (mlp.c:641-643)
            rd_seat(b, pos);
            rd_skip(b, (uint32_t)(pos & 3) * 8 + 32);
        }
    }
    L.end[0] = L.end[1] = 0; L.chk0 = 0;
    for (uint32_t k = 0; k < T.nss; k++) {
        if (q + 2 > L.total) return L;
        const uint32_t d = rd_get(b, 16);
        L.end[k] = (d & 0xFFF) * 2;
        if (k == 0) L.chk0 = (d >> 13) & 1;
        q += 2;
        if (d >> 15) { rd_skip(b, 16); q += 2; }
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

// Decodes substream job.k of segment job.seg, access unit by access unit.
// NCH > 0: fast path for exactly NCH channels in the substream; NCH = 0: generic.
template <int NCH>
__device__ __forceinline__ void decode_segment(const MlpTables &m, const DecodeJob &job, const uint16_t (*lut)[512],
                               uint32_t ring, const int32_t *init_hist)
{
    SegDev &S = m.segs[job.seg];
    const TrackDev &T = m.tracks[S.track];
    const GroupDev &G = m.groups[T.grp_base + (job.seg - T.seg_base) / DVDA_LANES];
    const bool governing = job.k + 1 == T.nss;
    const uint32_t nch = T.channels;

    SubState s;
    memset(&s, 0, sizeof s);
    s.flags = 0xFF;
    Hot<(NCH > 0 ? NCH : 1)> H;
    if (NCH > 0) {
#pragma unroll
        for (int cc = 0; cc < (NCH > 0 ? NCH : 1); cc++)
#pragma unroll
            for (int a = 0; a < 8; a++) { H.fh[cc][a] = 0; H.ih[cc][a] = 0; H.cf[cc][a] = 0; H.ci[cc][a] = 0; }
    } else if (init_hist) {
        for (int c = 0; c < DVDA_MAX_CH; c++) {
            for (int j = 0; j < 8; j++) s.ch[c].fst[j] = init_hist[c * 8 + j];
            s.ch[c].flen = 8; s.ch[c].fhead = 0;
        }
    }
    Rd b;
    rd_init(b, m.es, ring);

    uint32_t frames = 0, flags = 0, err = 0, stop_au = 0xFFFFFFFFu, pset = 0xFFFFFFFFu;
    bool abandon = false;          // fast path only: hand the segment to the generic fix-up pass
    uint64_t pos = S.n_au ? m.au_pos[S.au_base] : 0;
    for (uint32_t a = 0; a < S.n_au; a++) {
        const uint32_t A = S.au_base + a;
        const uint32_t e = m.au_err[A];
        const AuLayout L = au_layout_rd(b, pos, T);
        const uint64_t au_pos = pos;
        pos += L.total;                                           // the chain k_au_chase walked
        if (au_pos + L.total > T.es_cut) { stop_au = a; break; }  // end of track, not an error
        if (e == 1) {                                             // dropped access unit
            if (governing) { AuDev R = {frames, 0, s.seed, pset}; m.au[A] = R; }
            m.au_frames_ss[job.k * m.cap_au + A] = 0;
            continue;
        }
        if (e) { err |= e; stop_au = a; break; }
        const uint32_t start = job.k ? L.end[0] : 0;
        const uint32_t len = L.end[job.k] - start - (L.chk0 ? 2 : 0);
        const uint64_t data = au_pos + L.data0 + start;
        rd_seat(b, data);
        rd_skip(b, (uint32_t)(data & 3) * 8);
        const uint32_t end_bits = (uint32_t)(data & 3) * 8 + len * 8;
        const uint32_t au_frame0 = frames;
        bool bad = false;
        for (;;) {
            uint32_t n;
            if (NCH > 0) n = decode_block_fast<(NCH > 0 ? NCH : 1)>(m, G, job, s, H, b, end_bits, frames, nch, governing, lut, flags);
            else n = decode_block_generic(m, G, job, s, b, end_bits, frames, nch, governing, lut, flags);
            if (n == 0xFFFFFFFFu) { abandon = true; break; }
            if (!n) { bad = true; break; }
            frames += n;
            const uint32_t last = rd_get(b, 1);
            if (rd_pos(b) > end_bits) { bad = true; break; }
            if (last) break;
        }
        if (abandon) break;
        if (bad) { err |= ERR_SYNTAX; stop_au = a; frames = au_frame0; break; }
        const uint32_t nf = frames - au_frame0;
        // a restart header inside the AU reloads the noise seed before the AU is
        // rematrixed (mlp.c:828, 504-512): take it after the blocks
        const uint32_t seed0 = s.seed;
        m.au_frames_ss[job.k * m.cap_au + A] = nf;
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
                uint32_t shifts = 0;
                for (int c = 0; c < DVDA_MAX_CH; c++) { P.q[c] = s.q[c]; P.out_shift[c] = s.out_shift[c]; shifts |= s.out_shift[c]; }
                m.psets[A] = P;
                // top bit: nothing to do for these frames but to copy them
                pset = A | ((s.matrix_len == 0 && shifts == 0) ? 0x80000000u : 0u);
                s.dirty = 0;
            }
            AuDev R = {au_frame0, nf, seed0, pset};
            m.au[A] = R;
            uint32_t seed = s.seed;
            for (uint32_t i = 0; i < nf; i++) seed = noise_step(seed);
            s.seed = seed;
        }
    }
    cp_wait<0>();
    if (abandon) {
        m.ss_flags[job.k * m.cap_seg + job.seg] = SEG_NEEDS_CARRY;
        return;
    }

    // last 8 outputs per channel: slot 7 = most recent (what a circular buffer with fhead = 0 expects)
    int32_t *tail = m.fir_tail + ((uint64_t)job.k * m.cap_seg + job.seg) * (DVDA_MAX_CH * 8);
    if (s.have_header) {
        if (NCH > 0) {
#pragma unroll
            for (int cc = 0; cc < (NCH > 0 ? NCH : 1); cc++)
#pragma unroll
                for (int j = 0; j < 8; j++) tail[(s.min_ch + cc) * 8 + j] = H.fh[cc][7 - j];
        } else {
            for (uint32_t c = s.min_ch; c <= s.max_ch; c++) {
                const ChanState &C = s.ch[c];
                for (int j = 0; j < 8; j++) tail[c * 8 + j] = C.fst[(C.fhead + j) & 7];
            }
        }
    }
    else if (init_hist) {
        // nothing of this segment was decoded (its access units dropped): the history it was given goes on as it came
        for (int i = 0; i < DVDA_MAX_CH * 8; i++) tail[i] = init_hist[i];
    }
    m.ss_flags[job.k * m.cap_seg + job.seg] = flags;
    if (job.k == 0) S.frames = frames;
    if (err) atomicOr(&S.err, err);
    if (stop_au != 0xFFFFFFFFu) atomicMin(&S.err_au, stop_au);
}

// =============================================================================
// Fast path: three passes with access-unit parallelism.
//
// Decoding a segment in one thread (decode_segment above) is bound by the
// latency of one lane's dependent chain: the time does not shrink with the
// amount of work.  What actually chains across access units is small:
//   * decoding parameters inherit        -> pass A walks the *headers* of a
//     segment (first block of each AU; the AU positions are known from the
//     length chain, no residual has to be decoded to find them) and writes a
//     snapshot of what each AU needs;
//   * the entropy decode of an AU needs only that snapshot
//                                        -> pass B: one lane per (segment, AU),
//     16x more parallelism, residuals go to the tile;
//   * the prediction filters chain over frames but not over channels
//                                        -> pass C: one lane per (segment,
//     channel) runs the recurrence over the residuals in the tile, in place.
// The passes assume the common shape: at most 4 channels per substream,
// parameters only in the first block of an AU, AUs of the nominal length, no
// damage, no FIR history needed from the previous segment.  Anything else sets
// SEG_FALLBACK and the segment is decoded again by decode_segment (and
// k_carry_fix), which is the complete implementation.
// =============================================================================


// ---- pass A: headers ------------------------------------------------------------
//
// Three small kernels instead of one walk per segment:
//   A0  k_mlp_segctx    lane = (segment, substream): the restart header and the
//                       parameters of the segment's first access unit give the
//                       context every later parameter block is parsed in (channel
//                       range, matrix channel count, presence flags, noise seed);
//   A1  k_mlp_au_parse  lane = (segment, substream, access unit > 0): parses the
//                       AU's parameter block *as a delta* (what was transmitted,
//                       nothing resolved) and notes where the residuals begin;
//   A2  k_mlp_resolve   lane = (segment, substream): walks the heads of the deltas
//                       in order — no bit stream access — and writes what pass B
//                       needs per AU, the rematrix parameter sets and the noise seeds.
// The first access unit's "delta" states everything (defaults included), so the
// consumers treat all access units alike.  The filter passes read coefficients and
// histories straight from the deltas.  A parameter block that changes the presence
// flags, a restart header in the middle of a segment and everything malformed give
// the segment to the complete decoder.


// seat a reader on substream k of access unit A; false = not for the fast path
__device__ __forceinline__ bool au_seat(const MlpTables &m, const TrackDev &T, uint32_t A, uint32_t k, GRd &b, uint32_t *col,
                                        uint32_t &end_bits, uint64_t &origin)
{
    const uint64_t au_pos = m.au_pos[A];
    // the header and a typical parameter block in one go
    grd_stage(b, m.es, col, au_pos);
    const AuLayout L = au_layout_from([&b, au_pos](uint32_t i) { return grd_byte(b, au_pos + i); }, T);
    // damage and the end-of-track rules are the complete decoder's business (check data runs
    // beside the passes; k_flag_fallbacks hands the segments it objects to over afterwards)
    if (!L.ok || au_pos + L.total > T.es_cut) return false;
    const uint32_t start = k ? L.end[0] : 0;
    const uint32_t len = L.end[k] - start - (L.chk0 ? 2 : 0);
    const uint64_t data = au_pos + L.data0 + start;
    if (k) grd_stage(b, m.es, col, data);                 // the second substream sits further back
    grd_seat(b, data);
    rd_skip(b, (uint32_t)(data & 3) * 8);
    end_bits = (uint32_t)(data & 3) * 8 + len * 8;
    origin = (data & ~3ull) * 8;
    return true;
}

// One channel's FIR or IIR block of a delta (mlp.c:1029-1120).  The values are collected
// in registers (loops over all eight slots with compile-time indices): the record is
// written once, with 16-byte stores, by the caller.
template <typename RD>
__device__ __forceinline__ bool delta_filter(RD &b, bool iir, uint32_t &order_out, uint32_t &shift_out,
                                             int32_t (&coef)[8], int32_t (&ist)[8], uint32_t &present)
{
    const uint32_t order = rd_get(b, 4);
    if (order > 8) return false;
    uint32_t shift = 0;
    if (order) {
        shift = rd_get(b, 4);
        const uint32_t bits = rd_get(b, 5);
        if (bits < 1 || bits > 16) return false;
        const uint32_t cshift = rd_get(b, 3);
        if (bits + cshift > 16) return false;
#pragma unroll
        for (int i = 0; i < 8; i++) if ((uint32_t)i < order) coef[i] = (int32_t)((uint32_t)rd_get_s(b, bits) << cshift);
        if (rd_get(b, 1)) {
            if (!iir) return false;
            const uint32_t sbits = rd_get(b, 4), sshift = rd_get(b, 4);
            if (!sbits) return false;                           // reference underflows (G2)
#pragma unroll
            for (int i = 0; i < 8; i++) if ((uint32_t)i < order) ist[i] = (int32_t)((uint32_t)rd_get_s(b, sbits) << sshift);
            present |= CD_IIR_STATE;
        }
    }
    order_out = order; shift_out = shift;
    present |= iir ? CD_IIR : CD_FIR;
    return true;
}

__device__ __forceinline__ uint32_t pack16(int32_t lo, int32_t hi) { return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16); }

// decoding parameters of one block as a delta (mlp.c:856-993).  RESTART: the block
// follows a restart header, where everything not transmitted falls back to its
// default — the delta then states every field.
template <typename RD>
__device__ bool parse_delta(RD &b, const SegCtx &cx, AuDelta &D, ChanCoef *CF, MatCoef &MC, const bool RESTART)
{
    uint32_t present = 0, block_size = 8, matrix_len = 0;
    uint32_t mo[2] = {0, 0}, mb[2] = {0, 0};                    // mat_out / mat_bypass bytes 0-3, 4-5
    uint32_t os[2] = {0, 0}, qq[2] = {0, 0};                    // out_shift / q bytes 0-3, 4-7
    if (RESTART) {
        if (rd_get(b, 1)) rd_skip(b, 8);                        // the presence flags themselves: in cx.flags already
        present = AD_BLOCK | AD_MATRIX | AD_SHIFT | AD_Q;
    } else if ((cx.flags & 1) && rd_get(b, 1)) return false;    // new presence flags: complete decoder
    if ((cx.flags & 0x80) && rd_get(b, 1)) {
        block_size = rd_get(b, 9);
        if (block_size < 8) return false;
        present |= AD_BLOCK;
    }
    if ((cx.flags & 0x40) && rd_get(b, 1)) {
        matrix_len = rd_get(b, 4);
        if (matrix_len > DVDA_MAX_MAT || cx.mmc + 3 > DVDA_MAX_CH) return false;
        present |= AD_MATRIX;
        // (rolled loops: matrices are rare next to filter parameters, and the kernel's code
        // size matters — eight warps per scheduler run through different parts of it)
#pragma unroll 1
        for (uint32_t k = 0; k < matrix_len; k++) {
            const uint32_t out = rd_get(b, 4), frac = rd_get(b, 4);
            if (out > cx.mmc || frac > 14) return false;
            mo[k >> 2] |= out << (8 * (k & 3));
            mb[k >> 2] |= rd_get(b, 1) << (8 * (k & 3));
#pragma unroll 1
            for (uint32_t c = 0; c < DVDA_MAX_CH; c++) {
                int16_t v = 0;
                if (c < (uint32_t)cx.mmc + 3 && rd_get(b, 1)) v = (int16_t)((uint32_t)rd_get_s(b, frac + 2) << (14 - frac));
                MC.c[k][c] = v;
            }
        }
    }
    if ((cx.flags & 0x20) && rd_get(b, 1)) {
        present |= AD_SHIFT;
#pragma unroll
        for (int c = 0; c < DVDA_MAX_CH; c++) if ((uint32_t)c <= cx.mmc) os[c >> 2] |= ((uint32_t)rd_get_s(b, 4) & 31u) << (8 * (c & 3));
    }
    if ((cx.flags & 0x10) && rd_get(b, 1)) {
        present |= AD_Q;
#pragma unroll
        for (int c = 0; c < DVDA_MAX_CH; c++) if ((uint32_t)c <= cx.max_ch) qq[c >> 2] |= rd_get(b, 4) << (8 * (c & 3));
    }
    // head bytes 0..31: block_size, present, matrix_len | mat_out[6] mat_bypass[6] | out_shift[8] | q[8]
    static_assert(offsetof(AuDelta, mat_out) == 4 && offsetof(AuDelta, mat_bypass) == 10 && offsetof(AuDelta, out_shift) == 16 &&
                  offsetof(AuDelta, q) == 24 && offsetof(AuDelta, ch) == 32 && sizeof(ChanHead) == 12, "AuDelta head layout");
    uint4 *head = reinterpret_cast<uint4 *>(&D);
    head[0] = make_uint4(block_size | present << 16 | matrix_len << 24, mo[0],
                         (mo[1] & 0xFFFFu) | (mb[0] << 16), (mb[0] >> 16) | (mb[1] << 16));
    head[1] = make_uint4(os[0], os[1], qq[0], qq[1]);
    uint32_t *chw = reinterpret_cast<uint32_t *>(&D.ch[0]);
#pragma unroll 1
    for (uint32_t c = cx.min_ch; c <= cx.max_ch; c++) {
        const uint32_t cc = c - cx.min_ch;
        // defaults after a restart (mlp.c:904-989; the FIR history alone survives it)
        uint32_t p = RESTART ? (CD_PRESENT | CD_FIR | CD_IIR | CD_OFFSET) : 0u;
        uint32_t fo = 0, fsh = 0, io = 0, ish = 0, cb = 0, lsbs = 24;
        int32_t off = 0;
        int32_t fc[8], ic[8], st[8];
#pragma unroll
        for (int i = 0; i < 8; i++) { fc[i] = 0; ic[i] = 0; st[i] = 0; }
        if (rd_get(b, 1)) {
            p |= CD_PRESENT;
            if ((cx.flags & 0x08) && rd_get(b, 1) && !delta_filter(b, false, fo, fsh, fc, st, p)) return false;
            if ((cx.flags & 0x04) && rd_get(b, 1) && !delta_filter(b, true, io, ish, ic, st, p)) return false;
            if ((cx.flags & 0x02) && rd_get(b, 1)) { off = rd_get_s(b, 15); p |= CD_OFFSET; }
            cb = rd_get(b, 2);
            lsbs = rd_get(b, 5);
            if (lsbs > 24) return false;
        }
        chw[cc * 3 + 0] = (uint32_t)off;
        chw[cc * 3 + 1] = fo | fsh << 8 | io << 16 | ish << 24;
        chw[cc * 3 + 2] = cb | lsbs << 8 | p << 16;
        if (p & (CD_FIR | CD_IIR)) {
            uint4 *k = reinterpret_cast<uint4 *>(&CF[cc]);
            k[0] = make_uint4((uint32_t)st[0], (uint32_t)st[1], (uint32_t)st[2], (uint32_t)st[3]);
            k[1] = make_uint4((uint32_t)st[4], (uint32_t)st[5], (uint32_t)st[6], (uint32_t)st[7]);
            k[2] = make_uint4(pack16(fc[0], fc[1]), pack16(fc[2], fc[3]), pack16(fc[4], fc[5]), pack16(fc[6], fc[7]));
            k[3] = make_uint4(pack16(ic[0], ic[1]), pack16(ic[2], ic[3]), pack16(ic[4], ic[5]), pack16(ic[6], ic[7]));
        }
    }
    return true;
}

// restart header (mlp.c:809-854) into the segment context; the reader is left behind it
template <typename RD>
__device__ bool restart_ctx(RD &b, SegCtx &cx)
{
    const uint32_t sync = rd_get(b, 13), noise_type = rd_get(b, 1);
    rd_skip(b, 16);
    cx.min_ch = (uint8_t)rd_get(b, 4); cx.max_ch = (uint8_t)rd_get(b, 4); cx.mmc = (uint8_t)rd_get(b, 4);
    cx.noise_shift = (uint8_t)rd_get(b, 4);
    cx.seed = rd_get(b, 23);
    rd_skip(b, 19 + 1 + 8 + 16);
    if (sync != 0x18F5 || noise_type != 0) return false;
    if (cx.max_ch < cx.min_ch || cx.mmc < cx.max_ch || cx.mmc >= DVDA_MAX_CH) return false;
    for (uint32_t c = 0; c <= cx.mmc; c++) if (rd_get(b, 6) > cx.mmc) return false;
    rd_skip(b, 8);
    return true;
}

// A0: context of a segment's substream (no parameters yet: those are pass A1's)
__device__ __forceinline__ void segctx_segment(const MlpTables &m, const DecodeJob &job, uint32_t *col)
{
    const SegDev &S = m.segs[job.seg];
    const TrackDev &T = m.tracks[S.track];
    SegCtx cx;
    memset(&cx, 0, sizeof cx);
    if (S.n_au) {
        GRd b;
        uint32_t end_bits;
        uint64_t origin;
        if (au_seat(m, T, S.au_base, job.k, b, col, end_bits, origin) && rd_get(b, 1) && rd_get(b, 1) && restart_ctx(b, cx) &&
            cx.max_ch - cx.min_ch < 4) {
            // presence flags in force for the segment (all set unless the block says otherwise)
            uint32_t f = 0xFF;
            if (rd_get(b, 1)) { f = 0; for (int k = 0; k < 8; k++) f |= rd_get(b, 1) << k; }
            cx.flags = (uint8_t)f;
            cx.ok = rd_pos(b) <= end_bits;
        }
    }
    m.seg_ctx[job.k * m.cap_seg + job.seg] = cx;
}

// A1: parameter block of one access unit as a delta
__device__ __forceinline__ void parse_au(const MlpTables &m, const DecodeJob &job, uint32_t a, uint32_t *col)
{
    const SegDev &S = m.segs[job.seg];
    const TrackDev &T = m.tracks[S.track];
    const uint32_t A = S.au_base + a;
    AuSnap &sn = m.au_snap[(uint64_t)job.k * m.cap_au + A];
    const SegCtx cx = m.seg_ctx[job.k * m.cap_seg + job.seg];
    uint32_t state = 0;                                   // 0: not for the fast path, 1: no parameters, 2: delta written
    GRd b;
    uint32_t end_bits;
    uint64_t origin;
    if (cx.ok && au_seat(m, T, A, job.k, b, col, end_bits, origin)) {
        AuDelta &D = m.au_delta[(uint64_t)job.k * m.cap_au + A];
        if (a == 0) {
            // "parameters present", "restart header", the header itself (checked by pass A0)
            rd_skip(b, 2 + 113 + 6 * (cx.mmc + 1u) + 8);
            state = 2;
        } else {
            state = 1;
            // a restart header here would start a new run of parameters: complete decoder
            if (rd_get(b, 1)) state = rd_get(b, 1) ? 0 : 2;
        }
        // one call site: the parser is the bulk of this kernel's code
        if (state == 2 && !parse_delta(b, cx, D, m.au_cf + ((uint64_t)job.k * m.cap_au + A) * 4, m.au_mcoef[(uint64_t)job.k * m.cap_au + A], a == 0)) state = 0;
        if (rd_pos(b) > end_bits) state = 0;
        sn.bit0 = origin + rd_pos(b);
        sn.bit_end = origin + end_bits;
    }
    sn.valid = (uint8_t)state;
}

// A2: the segment's parameter chain, resolved per access unit
template <int NCH>
__device__ __forceinline__ void resolve_segment(const MlpTables &m, const DecodeJob &job)
{
    SegDev &S = m.segs[job.seg];
    const TrackDev &T = m.tracks[S.track];
    const bool governing = job.k + 1 == T.nss;
    const uint32_t nominal = T.au_nominal;
    const SegCtx cx = m.seg_ctx[job.k * m.cap_seg + job.seg];
    AuSnap *snaps = m.au_snap + (uint64_t)job.k * m.cap_au;
    const AuDelta *deltas = m.au_delta + (uint64_t)job.k * m.cap_au;
    uint8_t *fchg = m.au_fchg + (uint64_t)job.k * m.cap_au;

    // running state: what the entropy decoder needs, per channel of the substream
    uint32_t block_size = 8, want = 0, q8 = 0;            // q8: quant_step_size of channels 0..7, a nibble each
    uint64_t shift8 = 0;                                  // output shifts, a byte each
    uint32_t mat_src = 0;                                 // access unit whose delta holds the matrices in force
    uint32_t matrix_len = 0;
    int32_t offset[NCH];
    uint32_t cb[NCH], lsbs[NCH], fo[NCH], io[NCH], fs[NCH], is[NCH];
#pragma unroll
    for (int cc = 0; cc < NCH; cc++) { offset[cc] = 0; cb[cc] = 0; lsbs[cc] = 24; fo[cc] = io[cc] = fs[cc] = is[cc] = 0; }
    uint32_t seed = cx.seed, frames = 0, flags = 0, pset = 0xFFFFFFFFu;
    // (the channel layout the later passes count on: one substream from channel 0 on, or the stereo
    // pair in substream 0 and the rest, from channel 2 on, in substream 1)
    bool fallback = !cx.ok || (uint32_t)(cx.max_ch - cx.min_ch + 1) != NCH || cx.min_ch != (job.k ? 2u : 0u);

    bool first_cleared = false;
    uint32_t seed_at = 0;                                 // frame the seed belongs to (advanced only when some matrix uses noise)
    bool uses_noise = false;
    // The state byte and the head of the delta (five 16-byte loads) of the next access unit are
    // requested before the current one is worked on: one memory latency per step is hidden.
    struct Head { uint4 v[5]; };
    Head nextv;
    uint32_t next_state = 0;
    // (and, first of all, every access unit's record of the segment is asked into L2: the walk below
    // is a chain of dependent steps, each of which would otherwise wait for DRAM — 36 of the 40 cycles
    // between two instructions of this kernel went there)
    if (!fallback) {
        // (both tables are dense runs of memory for a segment: line by line)
        const uint8_t *p0 = reinterpret_cast<const uint8_t *>(deltas + S.au_base), *p1 = reinterpret_cast<const uint8_t *>(deltas + S.au_base + S.n_au);
        for (const uint8_t *p = p0; p < p1; p += 128) prefetch_l2(p);
        if (p1 > p0) prefetch_l2(p1 - 1);
        p0 = reinterpret_cast<const uint8_t *>(snaps + S.au_base); p1 = reinterpret_cast<const uint8_t *>(snaps + S.au_base + S.n_au);
        for (const uint8_t *p = p0; p < p1; p += 128) prefetch_l2(p);
        if (p1 > p0) prefetch_l2(p1 - 1);
    }
    if (S.n_au && !fallback) {
        next_state = snaps[S.au_base].valid;
        const uint4 *src = reinterpret_cast<const uint4 *>(deltas + S.au_base);
#pragma unroll
        for (int i = 0; i < 5; i++) nextv.v[i] = src[i];
    }
    for (uint32_t a = 0; a < S.n_au && !fallback; a++) {
        const uint32_t A = S.au_base + a;
        const uint32_t state = next_state;
        Head u = nextv;
        if (a + 1 < S.n_au) {
            next_state = snaps[A + 1].valid;
            const uint4 *src = reinterpret_cast<const uint4 *>(deltas + A + 1);
#pragma unroll
            for (int i = 0; i < 5; i++) nextv.v[i] = src[i];
        }
        if (!state) { fallback = true; break; }
        bool dirty = false;
        uint32_t chg = 0;                                 // channels whose filter set-up changes with this AU
        if (state == 2) {
            // words of the head by compile-time index (they stay in registers)
            uint32_t hw[20];
#pragma unroll
            for (int i = 0; i < 5; i++) { hw[4 * i] = u.v[i].x; hw[4 * i + 1] = u.v[i].y; hw[4 * i + 2] = u.v[i].z; hw[4 * i + 3] = u.v[i].w; }
#define HEAD_U8(off) ((hw[(off) >> 2] >> (8 * ((off) & 3))) & 0xFFu)
            const uint32_t present = HEAD_U8(offsetof(AuDelta, present));
            if (present & AD_BLOCK) block_size = hw[0] & 0xFFFFu;
            if (present & AD_MATRIX) {
                const uint32_t len_was = matrix_len;
                mat_src = A;
                matrix_len = HEAD_U8(offsetof(AuDelta, matrix_len));
                if (len_was | matrix_len) dirty = true;       // "no matrices" re-stated changes nothing
                want = 0;
#pragma unroll
                for (int k = 0; k < DVDA_MAX_MAT; k++)
                    want |= ((uint32_t)k < matrix_len && HEAD_U8(offsetof(AuDelta, mat_bypass) + k)) ? 1u << k : 0u;
            }
            const uint64_t shift8_was = shift8;
            const uint32_t q8_was = q8;
            if (present & AD_SHIFT) {
#pragma unroll
                for (int c = 0; c < DVDA_MAX_CH; c++)
                    if ((uint32_t)c <= cx.mmc) shift8 = (shift8 & ~(0xFFull << (8 * c))) | ((uint64_t)HEAD_U8(offsetof(AuDelta, out_shift) + c) << (8 * c));
            }
            if (present & AD_Q) {
#pragma unroll
                for (int c = 0; c < DVDA_MAX_CH; c++)
                    if ((uint32_t)c <= cx.max_ch) q8 = (q8 & ~(15u << (4 * c))) | (HEAD_U8(offsetof(AuDelta, q) + c) << (4 * c));
            }
            // re-stated values that did not change need no new parameter set
            if (shift8 != shift8_was || q8 != q8_was) dirty = true;
            if (q8 != q8_was) chg = (1u << NCH) - 1;
#pragma unroll
            for (int cc = 0; cc < NCH; cc++) {
                const int o = (int)offsetof(AuDelta, ch) + cc * (int)sizeof(ChanHead);
                const uint32_t p = HEAD_U8(o + offsetof(ChanHead, present));
                if (!(p & CD_PRESENT)) continue;
                if (p & CD_FIR) { fo[cc] = HEAD_U8(o + offsetof(ChanHead, fir_order)); fs[cc] = HEAD_U8(o + offsetof(ChanHead, fir_shift)); chg |= 1u << cc; }
                if (p & CD_IIR) {
                    io[cc] = HEAD_U8(o + offsetof(ChanHead, iir_order)); is[cc] = HEAD_U8(o + offsetof(ChanHead, iir_shift)); chg |= 1u << cc;
                    // the history is replaced by what was sent: too short = reference reads out of bounds (G2)
                    if (io[cc] && !(p & CD_IIR_STATE)) fallback = true;
                }
                if (p & CD_OFFSET) offset[cc] = (int32_t)hw[(o + offsetof(ChanHead, huff_offset)) >> 2];
                cb[cc] = HEAD_U8(o + offsetof(ChanHead, codebook)); lsbs[cc] = HEAD_U8(o + offsetof(ChanHead, huff_lsbs));
            }
#undef HEAD_U8
        }
        // (blocks of a multiple of eight frames: the entropy loops of the fast path step by eight)
        if (block_size > nominal || nominal % block_size || (block_size & 7)) { fallback = true; break; }

        // per-channel constants of the AU's blocks (mlp.c:1151-1176, 1260-1270), packed as the
        // five 64-bit words behind the positions in the snapshot: block_size, want, valid, min_ch,
        // nch, has_matrix, - | four times {sho, cb, lsb_bits, q, shift}
        uint64_t ow[5];
        ow[0] = (uint64_t)block_size | (uint64_t)want << 16 | 1ull << 24 | (uint64_t)cx.min_ch << 32 | (uint64_t)NCH << 40 |
                (uint64_t)(matrix_len != 0) << 48;
        ow[1] = ow[2] = ow[3] = ow[4] = 0;
#pragma unroll
        for (int cc = 0; cc < NCH; cc++) {
            const uint32_t q = (q8 >> (4 * (cx.min_ch + cc))) & 15;
            if (lsbs[cc] < q) { fallback = true; break; }
            const uint32_t nb = lsbs[cc] - q;
            int32_t sho;
            if (cb[cc]) {
                const int ss = (int)nb + 2 - (int)cb[cc];
                sho = offset[cc] - 7 * (1 << nb) - (ss >= 0 ? (1 << ss) : 0);
            } else {
                sho = offset[cc] - (nb >= 1 ? (1 << (nb - 1)) : 0);
            }
            if (fo[cc] + io[cc] > 8) { fallback = true; break; }
            if (fs[cc] > 0 && is[cc] > 0 && fs[cc] != is[cc]) { fallback = true; break; }
            if (a == 0 && fo[cc] > 0) {
                // FIR history of the previous segment is needed (the reference never clears it)
                if (!job.exact_history) flags |= SEG_WANTS_PREV;
                fallback = true; break;
            }
            const uint32_t shift = (fs[cc] > 0 && is[cc] > 0) ? fs[cc] : fo[cc] > 0 ? fs[cc] : is[cc];
            ow[1 + cc] = (uint64_t)(uint32_t)sho | (uint64_t)(cb[cc] | nb << 8 | q << 16 | shift << 24) << 32;
        }
        if (fallback) break;
        first_cleared = true;                             // (the first block starts no FIR from history this segment does not have)
        // bytes 16..55 of the snapshot (the positions in front are pass A1's)
        static_assert(offsetof(AuSnap, block_size) == 16 && offsetof(AuSnap, want) == 18 && offsetof(AuSnap, valid) == 19 &&
                      offsetof(AuSnap, min_ch) == 20 && offsetof(AuSnap, nch) == 21 && offsetof(AuSnap, has_matrix) == 22 &&
                      offsetof(AuSnap, ch) == 24 && sizeof(ChanSnap) == 8 && sizeof(AuSnap) == 56, "AuSnap layout");
        {
            uint64_t *dst = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(snaps + A) + 16);
#pragma unroll
            for (int i = 0; i < 5; i++) dst[i] = ow[i];
        }
        fchg[A] = (uint8_t)chg;

        const uint32_t au_frame0 = frames;
        frames += nominal;
        m.au_frames_ss[job.k * m.cap_au + A] = nominal;
        if (governing) {
            if (dirty || pset == 0xFFFFFFFFu) {
                ParamSet P;
                memset(&P, 0, sizeof P);
                P.matrix_len = (uint8_t)matrix_len; P.mmc = cx.mmc; P.noise_shift = cx.noise_shift;
                const AuDelta &M = m.au_delta[(uint64_t)job.k * m.cap_au + mat_src];
                const MatCoef &MC = m.au_mcoef[(uint64_t)job.k * m.cap_au + mat_src];
                uint32_t uses = 0;
                for (uint32_t k = 0; k < matrix_len; k++) {
                    P.out_ch[k] = M.mat_out[k];
                    for (int c = 0; c < DVDA_MAX_CH; c++) P.coeff[k][c] = MC.c[k][c];
                    uses |= (MC.c[k][cx.mmc + 1] != 0) | (MC.c[k][cx.mmc + 2] != 0);
                }
                P.uses_noise = uses;
                uses_noise = uses != 0;
                for (int c = 0; c < DVDA_MAX_CH; c++) { P.q[c] = (q8 >> (4 * c)) & 15; P.out_shift[c] = (uint8_t)(shift8 >> (8 * c)); }
                m.psets[A] = P;
                // top bit: nothing to do for these frames but to copy them
                pset = A | ((matrix_len == 0 && shift8 == 0) ? 0x80000000u : 0u);
            }
            // the generator steps once per frame; nobody looks at the seed of an AU whose matrices ignore the noise
            if (uses_noise) { seed = noise_advance(seed, au_frame0 - seed_at); seed_at = au_frame0; }
            AuDev R = {au_frame0, nominal, seed, pset};
            m.au[A] = R;
        }
    }
    // A segment handed to the complete decoder before its first block was looked at may turn out to
    // need the previous segment's FIR history there (the reference never clears it): the predecessor
    // then has to come from the complete decoder too, which stores its tail in time
    // (k_flag_fallbacks).  Only a first block seen to start without FIR taps rules that out.
    if (fallback && !first_cleared && !job.exact_history) flags |= SEG_WANTS_PREV;
    if (fallback) flags |= SEG_FALLBACK;
    m.ss_flags[job.k * m.cap_seg + job.seg] = flags;
    if (job.k == 0) S.frames = frames;
}


// ---- pass B: entropy decode of one access unit -------------------------------------
//
// The bit window is kept as two 32-bit halves (hi holds the next bits, MSB first):
// peeking is a shift of hi, consuming n <= 32 bits a funnel shift, topping up two
// funnel shifts — about half the instructions of 64-bit shifts.  All frames of the
// access unit fit the tile by construction (the group's capacity is its longest
// segment, access units have the nominal length here).
template <int NCH>
__device__ __forceinline__ void entropy_au(const MlpTables &m, const DecodeJob &job, uint32_t a,
                                           const uint16_t (*lut)[512], uint32_t ring)
{
    const SegDev &S = m.segs[job.seg];
    const TrackDev &T = m.tracks[S.track];
    const GroupDev &G = m.groups[T.grp_base + (job.seg - T.seg_base) / DVDA_LANES];
    const bool governing = job.k + 1 == T.nss;
    const uint32_t nch = T.channels, nominal = T.au_nominal;
    const uint32_t A = S.au_base + a;
    // pass A gave up on the segment: its later snapshots were never written
    if (m.ss_flags[job.k * m.cap_seg + job.seg] & SEG_FALLBACK) return;
    const AuSnap sn = m.au_snap[(uint64_t)job.k * m.cap_au + A];
    if (!sn.valid) return;
    const uint32_t frame0 = a * nominal;
    if (frame0 + nominal > G.cap) { atomicOr(&m.ss_flags[job.k * m.cap_seg + job.seg], SEG_OVERFLOW | SEG_FALLBACK); return; }

    Rd b;
    rd_init(b, m.es, ring);
    rd_seat(b, (sn.bit0 >> 5) << 2);
    rd_issue_ahead(b);
    rd_skip(b, (uint32_t)(sn.bit0 & 31));
    const uint32_t end_bits = (uint32_t)(sn.bit_end - ((sn.bit0 >> 5) << 5));

    int32_t sho[NCH];
    uint32_t lsb_bits[NCH], q[NCH];
    const uint16_t *cbt[NCH];
#pragma unroll
    for (int cc = 0; cc < NCH; cc++) { sho[cc] = sn.ch[cc].sho; lsb_bits[cc] = sn.ch[cc].lsb_bits; cbt[cc] = lut[sn.ch[cc].cb]; q[cc] = sn.ch[cc].q; }
    const uint32_t want = sn.want, bs = sn.block_size;
    // the bypass bits are looked at wherever the governing substream has matrices at all
    const bool store_byp = governing && sn.has_matrix;
    int32_t *tile = m.tiles + G.tile_off + job.lane + ((uint64_t)frame0 * nch + sn.min_ch) * DVDA_LANES;
    uint8_t *byp = m.bypass + G.byp_off + job.lane + (uint64_t)frame0 * DVDA_LANES;
    const uint32_t tile_step = nch * DVDA_LANES;
    constexpr uint32_t need8 = (8 * (NCH * 33 + 6) + 31) / 32 + 2;     // words eight frames can consume, plus the one fetched ahead
    constexpr uint32_t need1 = ((NCH * 33 + 6) + 31) / 32 + 2;
    uint32_t done = 0, bad = 0;
    uint32_t hi, lo;

#define DVDA_WIN_LOAD() { hi = (uint32_t)(b.win >> 32); lo = (uint32_t)b.win; }
#define DVDA_WIN_STORE() { b.win = ((uint64_t)hi << 32) | lo; }
    // frame FI of the current run (FI: compile-time offset into the tile)
#define DVDA_EFRAME(FI)                                                                            \
    {                                                                                              \
        if (want) {                                                                                \
            DVDA_WIN_STORE()                                                                       \
            const uint32_t bmask = bypass_bits(b, want);                                           \
            rd_hot_begin(b);                                                                       \
            DVDA_WIN_LOAD()                                                                        \
            if (store_byp) byp[(FI) * DVDA_LANES] = (uint8_t)bmask;                                \
        } else if (store_byp) byp[(FI) * DVDA_LANES] = 0;                                          \
        _Pragma("unroll") for (int cc = 0; cc < NCH; cc++) {                                       \
            if (b.avail <= 32) {                                                                   \
                hi |= __funnelshift_rc(b.ahead, 0u, (uint32_t)b.avail);                            \
                lo = __funnelshift_rc(0u, b.ahead, (uint32_t)b.avail);                             \
                b.avail += 32;                                                                     \
                b.next_w++;                                                                        \
            }                                                                                      \
            b.ahead = rd_ring_word(b, b.next_w);                                                   \
            const uint32_t e = cbt[cc][hi >> 23];                                                  \
            bad |= e;                                                                              \
            const uint32_t hl = (e >> 8) & 15;                                                     \
            const int32_t msb = e & 0xFF;                                                          \
            hi = __funnelshift_l(lo, hi, hl); lo <<= hl;                                           \
            const int32_t lsb = (int32_t)((hi >> 1) >> (31 - lsb_bits[cc]));                       \
            hi = __funnelshift_l(lo, hi, lsb_bits[cc]); lo <<= lsb_bits[cc];                       \
            b.avail -= hl + lsb_bits[cc];                                                          \
            tile[(FI) * tile_step + cc * DVDA_LANES] =                                             \
                (int32_t)((uint32_t)((msb << lsb_bits[cc]) + lsb + sho[cc]) << q[cc]);             \
        }                                                                                          \
    }

    bool ok = true;
    while (ok) {
        // one block of bs frames (bs divides the nominal AU length, bs >= 8)
        uint32_t i = 0;
        for (; i + 8 <= bs; i += 8) {
            rd_prefetch(b, need8);
            rd_hot_begin(b);
            DVDA_WIN_LOAD()
            DVDA_EFRAME(0) DVDA_EFRAME(1) DVDA_EFRAME(2) DVDA_EFRAME(3)
            DVDA_EFRAME(4) DVDA_EFRAME(5) DVDA_EFRAME(6) DVDA_EFRAME(7)
            DVDA_WIN_STORE()
            tile += 8 * tile_step; byp += 8 * DVDA_LANES;
        }
        for (; i < bs; i++) {
            rd_prefetch(b, need1);
            rd_hot_begin(b);
            DVDA_WIN_LOAD()
            DVDA_EFRAME(0)
            DVDA_WIN_STORE()
            tile += tile_step; byp += DVDA_LANES;
        }
        done += bs;
        if ((bad & 0x8000) || rd_pos(b) > end_bits) { ok = false; break; }
        const uint32_t last = rd_get(b, 1);
        if (rd_pos(b) > end_bits) { ok = false; break; }
        if (last) break;
        // another block: it must not bring parameters of its own, and must fit the AU
        if (done + bs > nominal || rd_get(b, 1)) { ok = false; break; }
    }
#undef DVDA_EFRAME
#undef DVDA_WIN_LOAD
#undef DVDA_WIN_STORE
    cp_wait<0>();
    if (!ok || done != nominal) atomicOr(&m.ss_flags[job.k * m.cap_seg + job.seg], SEG_FALLBACK);
}

// ---- pass C: prediction filters + rematrix + interleaved output, one or two substreams ------
//
// One lane per (segment, channel).  A segment of a track with n0 channels in substream 0 and n1
// in substream 1 (n1 = 0: one substream) takes n0 + n1 neighbouring lanes, a warp takes
// 32 / (n0 + n1) segments of a group, and all lanes step through the frames of their segments
// together (access units have the nominal length here).  Residuals come from the tile; the
// next 8 frames are loaded while the current 8 are filtered (the tile carries 16 frames of
// slack, no bounds tests).  Per access unit the warp picks the smallest compiled tap count that
// covers the filter orders of all its lanes (0, 4 or 8 taps each for FIR and IIR).  Filtered
// samples are parked in a shared-memory patch [segment][32 frames][channels], in output channel
// order.  When a patch is full the warp turns round: one lane per FRAME of a row applies what
// the access unit's parameters ask for beyond a copy — noise, the matrices in order, bypass
// bits, output shift (mlp.c:504-538, 1308-1358) — in place, each frame worked once with all its
// channels in registers; channels of both substreams meet there and substream 1's matrices
// govern all of them (mlp.c:575-595).  Then the patch leaves row by row as bulk copies
// shared -> global.  Runs after the frame counts are final (it writes straight into the PCM
// buffer).
#define OUT_WARPS 4
#ifndef OUT_MIN_BLOCKS
#define OUT_MIN_BLOCKS 5
#endif
#define OUT_PF 32                                       // frames per patch
#define OUT_ROW_PAD 4                                    // per row: a note for the matrix step on each run of 8 frames
#define OUT_PATCH_WORDS (OUT_PF * 32 + OUT_ROW_PAD * 32) // most a patch needs: spw rows of OUT_PF * lanes-per-segment + OUT_ROW_PAD words
#define OUT_WARP_WORDS (2 * OUT_PATCH_WORDS + 8 * 16)   // two patches, used in turn (bulk stores in flight) + per segment {output base lo, hi, frames, aligned, parameter sets of two units in turn}
#define NOTE_UNIT (1u << 23)                            // note of a run of 8 frames: noise seed (23 bits) | which of the two units | nothing to do
#define NOTE_SKIP (1u << 24)
#define OUT_SMEM_BYTES (OUT_WARPS * OUT_WARP_WORDS * 4)
#define OUT_MAX_LPS 8

// The matrix step of a patch of the output pass: lane = frame of a row.  The row's pad words say
// which parameters govern each run of 8 frames and where the noise generator stood at its start.

// Six-channel tracks (five rows of 196 words in a patch of 1152): the parameter sets of the two
// access units a patch can touch are kept in the patches' unused tails, brought there by the
// segment's lanes at the top of each unit — the frames of the matrix step then wait for shared
// memory, not for a chain of dependent global loads.
#define PSET_SM_OFF6 (5 * (OUT_PF * 6 + OUT_ROW_PAD))     // first free word of a patch
#define PSET_SM_WORDS 32                                  // a ParamSet
static_assert(PSET_SM_OFF6 + 5 * PSET_SM_WORDS <= OUT_PATCH_WORDS && (PSET_SM_OFF6 % 4) == 0, "room for five parameter sets behind the rows");
template <int NL>
__device__ __forceinline__ void matrix_rows(const ParamSet *__restrict__ psets, int32_t *pb, const int32_t *patch, const uint32_t *meta, const uint8_t *__restrict__ byp_rows,
                                         uint32_t row_words, uint32_t lps, uint32_t spw, uint32_t slotmap, uint32_t f0, uint32_t fend)
{
    constexpr bool SM = NL == 6;                          // parameter sets in shared memory
    auto ld8 = [](const void *p) { return SM ? *reinterpret_cast<const uint2 *>(p) : __ldg(reinterpret_cast<const uint2 *>(p)); };
    auto ld16 = [](const void *p) { return SM ? *reinterpret_cast<const uint4 *>(p) : __ldg(reinterpret_cast<const uint4 *>(p)); };
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t fr = f0 + lane;
    for (uint32_t row = 0; row < spw; row++) {
        int32_t *rw = pb + row * row_words;
        const uint32_t note = (uint32_t)rw[OUT_PF * lps + (lane >> 3)];
        if (fr >= fend || (note & NOTE_SKIP)) continue;
        const uint32_t unit = (note / NOTE_UNIT) & 1;
        const uint32_t ps = meta[row * 8 + 4 + unit];
        if (ps == 0xFFFFFFFFu) continue;
        const ParamSet *Q = SM ? reinterpret_cast<const ParamSet *>(patch + unit * OUT_PATCH_WORDS + PSET_SM_OFF6 + row * PSET_SM_WORDS) : &psets[ps];
        int32_t *px = rw + lane * lps;
        int32_t v[NL];
#pragma unroll
        for (int c = 0; c < NL; c++) v[c] = (uint32_t)c < lps ? px[(slotmap >> (4 * c)) & 15] : 0;
        const uint2 hd = ld8(Q->out_ch);                                           // out_ch[6], matrix_len, mmc
        const uint64_t ocs = (uint64_t)hd.y << 32 | hd.x;
        const uint32_t ml = (hd.y >> 16) & 0xFF, mmc = hd.y >> 24;
        if (ml) {
            uint32_t sd = note & (NOTE_UNIT - 1);
            for (uint32_t i = 0; i < (lane & 7); i++) sd = noise_step(sd);
            const uint32_t nsh = Q->noise_shift;
            const int32_t z0 = (int32_t)((uint32_t)(int32_t)(int8_t)(sd >> 15) << nsh);
            const int32_t z1 = (int32_t)((uint32_t)(int32_t)(int8_t)(sd >> 7) << nsh);
            const uint32_t bm = byp_rows[(uint64_t)fr * DVDA_LANES + row];
            const uint2 qw = ld8(Q->q);
            const uint64_t qs = (uint64_t)qw.y << 32 | qw.x;
            for (uint32_t mk = 0; mk < ml; mk++) {
                const uint4 cw = ld16(Q->coeff[mk]);
                const uint32_t w[4] = {cw.x, cw.y, cw.z, cw.w};
                long long sum = (long long)z0 * Q->coeff[mk][mmc + 1] + (long long)z1 * Q->coeff[mk][mmc + 2];
#pragma unroll
                for (int c = 0; c < NL; c++) {
                    const int32_t co = (int32_t)(int16_t)(w[c >> 1] >> (16 * (c & 1)));
                    if ((uint32_t)c <= mmc) sum += (long long)v[c] * co;
                }
                const uint32_t oc = (uint32_t)(ocs >> (8 * mk)) & 0xFF, qq = (uint32_t)(qs >> (8 * oc)) & 0xFF;
                const int32_t rr = (((int32_t)(sum >> 14)) >> qq << qq) + (int32_t)((bm >> mk) & 1);
#pragma unroll
                for (int c = 0; c < NL; c++) if ((uint32_t)c == oc) v[c] = rr;
            }
        }
        const uint2 ow = ld8(Q->out_shift);
        const uint64_t os = (uint64_t)ow.y << 32 | ow.x;
#pragma unroll
        for (int c = 0; c < NL; c++)
            if ((uint32_t)c < lps)
                px[(slotmap >> (4 * c)) & 15] = (uint32_t)c <= mmc ? (int32_t)((uint32_t)v[c] << ((uint32_t)(os >> (8 * c)) & 0xFF)) : v[c];
    }
}

// LPS: lanes per segment = channels of the track, fixed at compile time for the common layouts
// (2: stereo, 6: stereo pair + four more), 0: taken from the work list
template <int LPS>
__device__ __forceinline__ void filter_out_warp(const MlpTables &m, const OutWork &W, uint32_t warp, int32_t *out_sm)
{
    constexpr int NL = LPS ? LPS : OUT_MAX_LPS;             // channels that can meet in the matrix step
    const uint32_t wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const TrackDev &T = m.tracks[W.track];
    const uint32_t n0 = W.n0, lps = LPS ? (uint32_t)LPS : W.n0 + W.n1;            // lanes per segment = channels of the track
    const uint32_t spw = LPS ? 32 / lps : out_segs_per_warp(lps);   // segments per warp
    const uint32_t sub_n = (32 + spw - 1) / spw;           // warps per group
    const uint32_t rel = warp - W.warp0;
    const GroupDev &G = m.groups[T.grp_base + rel / sub_n];
    const uint32_t sub = rel % sub_n;
    const uint32_t sl = lane / lps, j = lane - sl * lps;   // segment in the warp, channel of the track
    const uint32_t k = j >= n0 ? 1u : 0u;                  // substream
    const uint32_t cc = k ? j - n0 : j;                    // channel inside the substream
    const uint32_t sg = sub * spw + sl;                    // segment inside the group = column of the tile
    const bool have = sl < spw && sg < G.nseg;
    const uint32_t seg = G.seg0 + (have ? sg : 0);
    const SegDev &S = m.segs[seg];
    uint32_t seg_flags = m.ss_flags_fast[seg];
    if (W.n1) seg_flags |= m.ss_flags_fast[m.cap_seg + seg];
    const bool mine = have && !(seg_flags & SEG_FALLBACK) && S.frames > 0;
    const uint32_t my_frames = mine ? S.frames : 0;
    const uint32_t max_frames = __reduce_max_sync(0xFFFFFFFFu, my_frames);
    if (!max_frames) return;
    const uint32_t nominal = T.au_nominal;

    const uint32_t row_words = OUT_PF * lps + OUT_ROW_PAD; // a segment's row in a patch (rows stay 16-byte aligned, banks spread)
    int32_t *patch = out_sm + (size_t)wib * OUT_WARP_WORDS;
    uint32_t *meta = reinterpret_cast<uint32_t *>(patch + 2 * OUT_PATCH_WORDS);
    if (j == 0 && sl < spw) {
        const uint64_t base = (mine ? S.frame0 : 0) * lps;
        meta[sl * 8 + 0] = (uint32_t)base; meta[sl * 8 + 1] = (uint32_t)(base >> 32); meta[sl * 8 + 2] = my_frames;
        meta[sl * 8 + 3] = ((T.out_base + base) & 3) == 0;        // the rows of this segment start on 16-byte boundaries
    }
    __syncwarp();

    const AuDelta *deltas = m.au_delta + (uint64_t)k * m.cap_au;
    const ChanCoef *coefs = m.au_cf + (uint64_t)k * m.cap_au * 4 + cc;      // + 4 * access unit
    const uint8_t *fchg = m.au_fchg + (uint64_t)k * m.cap_au;
    FiltSetup F = {0, 0, 0, 0, 0};
    int32_t fh[8], ih[8], cf[8], ci[8];
#pragma unroll
    for (int t = 0; t < 8; t++) { fh[t] = 0; ih[t] = 0; cf[t] = 0; ci[t] = 0; }
    uint32_t shift = 0, qmask = 0xFFFFFFFFu;
    const uint32_t tile_step = lps * DVDA_LANES;
    const int32_t *tp = m.tiles + G.tile_off + (have ? sg : 0) + (uint64_t)j * DVDA_LANES;
    int32_t *const pcm_row = m.pcm + T.out_base;
    const uint32_t out_slot = wave_slot(T.assignment, j);
    int32_t *const park = patch + (sl < spw ? sl : 0) * row_words + out_slot;
    uint32_t slotmap = 0;                                        // output slot of channel c, 4 bits each
#pragma unroll
    for (int c = 0; c < NL; c++) slotmap |= __shfl_sync(0xFFFFFFFFu, out_slot, c) << (4 * c);
    const uint8_t *const byp_rows = m.bypass + G.byp_off + sub * spw;   // + frame * DVDA_LANES + row
    bool patch_matrix = false;                                   // some access unit of the patch being filled wants the matrix step

    uint32_t seed = 0, pset = 0xFFFFFFFFu, f = 0, a = 0, cls = 0;
    bool trivial = true;
    int32_t nx[8];
#pragma unroll
    for (int t = 0; t < 8; t++) nx[t] = tp[t * tile_step];
    tp += 8 * tile_step;

    // The patch of 32 frames leaves row by row (a row = 32 frames of a segment, contiguous in the
    // output).  Whole, 16-byte aligned rows go out as bulk copies shared -> global, one instruction
    // per row issued by the row's lane; the copy engine reads the patch while the warp fills the
    // other one.  Ragged ends take plain stores.
    auto flush = [&](uint32_t f0, uint32_t fend) {
        int32_t *pb = patch + ((f0 / OUT_PF) & 1) * OUT_PATCH_WORDS;
        if (patch_matrix) {
            __syncwarp();
            matrix_rows<NL>(m.psets, pb, patch, meta, byp_rows, row_words, lps, spw, slotmap, f0, fend);
            patch_matrix = false;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the parked samples, for the async proxy
        __syncwarp();
        uint32_t slow = 0;
        if (lane < spw) {
            const uint4 mt = *reinterpret_cast<const uint4 *>(meta + lane * 8);     // base lo, hi, frames, aligned
            if (f0 < mt.z) {
                if (f0 + OUT_PF <= mt.z && mt.w) {
                    int32_t *dst = pcm_row + (((uint64_t)mt.y << 32 | mt.x) + (uint64_t)f0 * lps);
                    const uint32_t src = (uint32_t)__cvta_generic_to_shared(pb + lane * row_words);
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                 :: "l"(dst), "r"(src), "r"(OUT_PF * 4 * lps) : "memory");
                } else slow = 1;
            }
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        uint32_t rows = __ballot_sync(0xFFFFFFFFu, slow);
        while (rows) {
            const uint32_t row = __ffs(rows) - 1;
            rows &= rows - 1;
            const uint4 mt = *reinterpret_cast<const uint4 *>(meta + row * 8);
            const uint32_t n = min((uint32_t)OUT_PF, mt.z - f0) * lps;
            int32_t *dst = pcm_row + (((uint64_t)mt.y << 32 | mt.x) + (uint64_t)f0 * lps);
            for (uint32_t i = lane; i < n; i += 32) dst[i] = pb[row * row_words + i];
        }
        // the patch written one flush ago has been read by now: it is the one filled next
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
    };

    // the head of the access unit that comes next is always in registers already: what a lane
    // waits for at the top of an access unit is one round of four independent 16-byte loads
    // (prefetched into L1), not a chain of dependent ones
    const uint32_t au_base = mine ? S.au_base : 0;       // (a register: S lives in global memory)
    auto load_head = [&](uint32_t A) {
        DeltaHead H;
        const AuDelta &D = deltas[A];
        H.fchg = fchg[A];
        H.w0 = *reinterpret_cast<const uint32_t *>(&D);                  // block_size, present, matrix_len
        H.qv = D.q[j];
        const uint32_t *hw = reinterpret_cast<const uint32_t *>(&D.ch[cc]);
        H.h1 = hw[1]; H.h2 = hw[2];
        const uint2 sp = *reinterpret_cast<const uint2 *>(&m.au[A].seed);
        H.seed = sp.x; H.pset = sp.y;
        return H;
    };
    DeltaHead H = load_head(au_base);
    while (f < max_frames) {
        // ---- next access unit: this channel's filter parameters, the frame's rematrix parameters
        const bool au_act = f < my_frames;
        if (au_act) {
            const uint32_t A = au_base + a;
            const uint32_t An = min(A + 1, m.cap_au);        // (the tables have one spare entry)
            // (a prefetch brings a whole 128-byte line: the coefficients of a pair of channels — one lane asks)
            if (!(cc & 1)) prefetch_l1(&coefs[(uint64_t)An * 4]);
            if ((H.fchg >> cc) & 1) {
                filt_take_head(H, coefs[(uint64_t)A * 4], F, cf, ci, ih);
                shift = filt_shift(F); qmask = 0xFFFFFFFFu << F.q;
                cls = F.fo | F.io << 4;
            }
            seed = H.seed;
            if (H.pset != pset) {
                pset = H.pset;
                trivial = (pset & 0x80000000u) != 0;          // (a permuted channel order is dealt with when parking)
            }
            H = load_head(An);
        } else {
            cls = 0;
#pragma unroll
            for (int t = 0; t < 8; t++) { cf[t] = 0; ci[t] = 0; }
        }
        if (j == 0 && sl < spw) meta[sl * 8 + 4 + (a & 1)] = au_act && !trivial ? pset & 0x7FFFFFFFu : 0xFFFFFFFFu;
        if (LPS == 6 && au_act && !trivial && sl < spw) {
            // the unit's parameter set, eight pieces of 16 bytes, into the tail of the patch with the unit's parity
            const uint4 *src = reinterpret_cast<const uint4 *>(&m.psets[pset & 0x7FFFFFFFu]);
            uint4 *dst = reinterpret_cast<uint4 *>(patch + (a & 1) * OUT_PATCH_WORDS + PSET_SM_OFF6 + sl * PSET_SM_WORDS);
            dst[j] = __ldg(src + j);
            if (j < 2) dst[6 + j] = __ldg(src + 6 + j);
        }
        a++;
        const uint32_t nf = __reduce_max_sync(0xFFFFFFFFu, ((cls & 15) + 3) >> 2);
        const uint32_t ni = __reduce_max_sync(0xFFFFFFFFu, ((cls >> 4) + 3) >> 2);
        const uint32_t code = nf * 3 + ni;
        const bool any_matrix = __any_sync(0xFFFFFFFFu, au_act && !trivial);
        if (any_matrix && !patch_matrix) {
            // first unit of this patch to want the matrix step: the runs of 8 frames in front of it do not
            patch_matrix = true;
            if (j == 0 && sl < spw)
                for (uint32_t b = 0; b < ((f & (OUT_PF - 1)) >> 3); b++)
                    patch[((f / OUT_PF) & 1) * OUT_PATCH_WORDS + sl * row_words + OUT_PF * lps + b] = (int32_t)NOTE_SKIP;
        }

        for (uint32_t i = 0; i < nominal; i += 8) {
            int32_t r[8];
#pragma unroll
            for (int t = 0; t < 8; t++) r[t] = nx[t];
#pragma unroll
            for (int t = 0; t < 8; t++) nx[t] = tp[t * tile_step];
            tp += 8 * tile_step;
            switch (code) {
            case 0: filt8<0, 0>(cf, ci, fh, ih, r, shift, qmask); break;
            case 1: filt8<0, 4>(cf, ci, fh, ih, r, shift, qmask); break;
            case 2: filt8<0, 8>(cf, ci, fh, ih, r, shift, qmask); break;
            case 3: filt8<4, 0>(cf, ci, fh, ih, r, shift, qmask); break;
            case 4: filt8<4, 4>(cf, ci, fh, ih, r, shift, qmask); break;
            case 5: filt8<4, 8>(cf, ci, fh, ih, r, shift, qmask); break;
            case 6: filt8<8, 0>(cf, ci, fh, ih, r, shift, qmask); break;
            case 7: filt8<8, 4>(cf, ci, fh, ih, r, shift, qmask); break;
            default: filt8<8, 8>(cf, ci, fh, ih, r, shift, qmask); break;
            }
            // what the matrix step of this patch needs to know about these 8 frames
            if (patch_matrix && j == 0 && sl < spw)
                patch[((f / OUT_PF) & 1) * OUT_PATCH_WORDS + sl * row_words + OUT_PF * lps + ((f & (OUT_PF - 1)) >> 3)] =
                    (int32_t)((seed & (NOTE_UNIT - 1)) | (a & 1 ? 0u : NOTE_UNIT));        // (a counts the unit already)
            if (any_matrix) {
#pragma unroll
                for (int t = 0; t < 8; t++) seed = noise_step(seed);
            }
            if (au_act) {
                int32_t *pk = park + ((f / OUT_PF) & 1) * OUT_PATCH_WORDS + (f & (OUT_PF - 1)) * lps;
#pragma unroll
                for (int t = 0; t < 8; t++) pk[t * lps] = r[t];
            }
            f += 8;
            if ((f & (OUT_PF - 1)) == 0) { flush(f - OUT_PF, f); patch_matrix = any_matrix; }   // (the unit may go on into the next patch)
        }
    }
    if (f & (OUT_PF - 1)) flush(f & ~(uint32_t)(OUT_PF - 1), f);
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");    // shared memory is given back at exit
    // FIR tail for a following segment that needs it
    if (mine) {
        int32_t *tail = m.fir_tail + ((uint64_t)k * m.cap_seg + seg) * (DVDA_MAX_CH * 8);
#pragma unroll
        for (int t = 0; t < 8; t++) tail[j * 8 + t] = fh[7 - t];
    }
}

__global__ void __launch_bounds__(OUT_WARPS * 32, OUT_MIN_BLOCKS) k_mlp_filter_out(MlpTables m, const OutWork *__restrict__ work)
{
    extern __shared__ int32_t out_sm[];
    const uint32_t warp = blockIdx.x * OUT_WARPS + (threadIdx.x >> 5);
    if (warp >= m.cnt->nout_warps) return;
    // queued before the host has seen the batch's status: nothing to do if the batch is decoded
    // once more (tile overflow) or the output buffer sized in advance turned out too small
    if (*m.status & (SEG_OVERFLOW | STATUS_PCM_SMALL)) return;
    // ---- which track
    uint32_t lo = 0, hi = m.cnt->nout_work;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (work[mid].warp0 <= warp) lo = mid; else hi = mid;
    }
    const OutWork W = work[lo];
    switch (W.n0 + W.n1) {
    case 2: filter_out_warp<2>(m, W, warp, out_sm); break;
    case 6: filter_out_warp<6>(m, W, warp, out_sm); break;
    default: filter_out_warp<0>(m, W, warp, out_sm); break;
    }
}

// cap_warps: warps the grid covers (the work list and its length are on the device)
int launch_mlp_filter_out(MlpTables m, const OutWork *work, uint32_t cap_warps, cudaStream_t s)
{
    if (!cap_warps) return 0;
    static PerDeviceOnce attr_once;
    if (attr_once.run([&]() -> int { CUDA_TRY(cudaFuncSetAttribute(k_mlp_filter_out, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OUT_SMEM_BYTES)); return 0; })) return -1;
    LAUNCH(k_mlp_filter_out, div_up_u32(cap_warps, OUT_WARPS), OUT_WARPS * 32, OUT_SMEM_BYTES, s, m, work);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

#define DEC_WARPS 4
#ifndef DEC_MIN_BLOCKS
#define DEC_MIN_BLOCKS 3
#endif
#define DEC_SMEM_BYTES (DEC_WARPS * RING_SLOTS * DVDA_LANES * 16 + 4 * 512 * 2)

// One warp per (group, substream) — a row of the work list covers the groups of one track's
// substream; lane = segment of the group.  The kernels below look their warp up in the list (its
// length and the number of warps are on the device: the grids cover what the tables have room for)
// and dispatch on the substream's channel count: the per-channel-count code is compiled once per
// count (NCH = 0: generic, more than 4 channels), so that the common stereo case does not run with
// the unrolling of the 4-channel one.
__device__ __forceinline__ bool pair_job(const MlpTables &m, const DecWork *work, uint32_t warp, uint32_t lane,
                                         DecodeJob &job, uint32_t &nch)
{
    if (warp >= m.cnt->npairs) return false;
    uint32_t lo = 0, hi = m.cnt->nwork;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (work[mid].warp0 <= warp) lo = mid; else hi = mid;
    }
    const DecWork W = work[lo];
    const TrackDev &T = m.tracks[W.track];
    const GroupDev &G = m.groups[T.grp_base + (warp - W.warp0)];
    nch = W.nch;
    if (lane >= G.nseg) return false;
    job.seg = G.seg0 + lane;
    job.k = W.k;
    job.lane = lane;
    // a track starts with empty histories; a continued part does not
    job.exact_history = (job.seg == T.seg_base) && !(T.cont & TRACK_CONT_PREV);
    return true;
}

// the complete decoder: everything (without the fast path), or what the fast path gave up on
__global__ void __launch_bounds__(DEC_WARPS * 32, DEC_MIN_BLOCKS) k_mlp_decode(MlpTables m, const DecWork *__restrict__ work)
{
    extern __shared__ uint4 dyn_smem[];
    if (m.fast && !*m.any_fallback) return;              // the fast path kept every segment
    uint4 (*ring)[RING_SLOTS][DVDA_LANES] = reinterpret_cast<uint4 (*)[RING_SLOTS][DVDA_LANES]>(dyn_smem);
    uint16_t (*lut)[512] = reinterpret_cast<uint16_t (*)[512]>(dyn_smem + DEC_WARPS * RING_SLOTS * DVDA_LANES);
    huff_lut_to_shared(lut);
    __syncthreads();
    const uint32_t wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t warp = blockIdx.x * DEC_WARPS + wib;
    DecodeJob job;
    uint32_t nch;
    if (!pair_job(m, work, warp, lane, job, nch)) return;
    if (m.fast) {
        // after the fast path: only what it gave up on (any substream of the segment)
        const TrackDev &T = m.tracks[m.segs[job.seg].track];
        const uint32_t f = m.ss_flags_fast[job.seg] | (T.nss == 2 ? m.ss_flags_fast[m.cap_seg + job.seg] : 0);
        if (!(f & SEG_FALLBACK)) return;
    }
    const uint32_t rs = (uint32_t)__cvta_generic_to_shared(&ring[wib][0][lane]);
    switch (nch) {
    case 1: decode_segment<1>(m, job, lut, rs, nullptr); break;
    case 2: decode_segment<2>(m, job, lut, rs, nullptr); break;
    case 3: decode_segment<3>(m, job, lut, rs, nullptr); break;
    case 4: decode_segment<4>(m, job, lut, rs, nullptr); break;
    default: decode_segment<0>(m, job, lut, rs, nullptr); break;
    }
}

// ---- fast path kernels ------------------------------------------------------------

// pass A0: one warp per (group, substream), lane = segment
__global__ void __launch_bounds__(GRD_THREADS) k_mlp_segctx(MlpTables m, const DecWork *__restrict__ work)
{
    __shared__ uint32_t window[GRD_WIN_WORDS * GRD_THREADS];
    const uint32_t lane = threadIdx.x & 31, warp = blockIdx.x * 4 + (threadIdx.x >> 5);
    DecodeJob job;
    uint32_t nch;
    if (!pair_job(m, work, warp, lane, job, nch) || nch < 1 || nch > 4) return;
    segctx_segment(m, job, window + threadIdx.x);
}
// pass A2
__global__ void __launch_bounds__(128) k_mlp_resolve(MlpTables m, const DecWork *__restrict__ work)
{
    const uint32_t lane = threadIdx.x & 31, warp = blockIdx.x * 4 + (threadIdx.x >> 5);
    DecodeJob job;
    uint32_t nch;
    if (!pair_job(m, work, warp, lane, job, nch)) return;
    switch (nch) {
    case 1: resolve_segment<1>(m, job); break;
    case 2: resolve_segment<2>(m, job); break;
    case 3: resolve_segment<3>(m, job); break;
    case 4: resolve_segment<4>(m, job); break;
    default: break;                                     // more than four channels: the complete decoder's
    }
}
// pass A1: one thread per (access unit, substream), consecutive threads = consecutive access units
__global__ void __launch_bounds__(GRD_THREADS) k_mlp_au_parse(MlpTables m)
{
    __shared__ uint32_t window[GRD_WIN_WORDS * GRD_THREADS];
    const uint32_t A = blockIdx.x * GRD_THREADS + threadIdx.x, k = blockIdx.y;
    if (A >= m.cnt->nau) return;
    DecodeJob job;
    job.seg = m.au_seg[A];
    job.k = k;
    job.lane = 0;
    job.exact_history = false;
    const SegDev &S = m.segs[job.seg];
    if (k >= m.tracks[S.track].nss) return;
    parse_au(m, job, A - S.au_base, window + threadIdx.x);
}

// pass B: one warp per (group, substream, access unit index), lane = segment
__global__ void __launch_bounds__(DEC_WARPS * 32) k_mlp_entropy(MlpTables m, const DecWork *__restrict__ work)
{
    extern __shared__ uint4 dyn_smem[];
    const uint32_t wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t warp0 = blockIdx.x * DEC_WARPS;
    const uint32_t a = blockIdx.y;
    if (warp0 >= m.cnt->npairs || a >= m.cnt->max_au) return;            // (before the table is copied: most blocks of an oversized grid)
    uint4 (*ring)[RING_SLOTS][DVDA_LANES] = reinterpret_cast<uint4 (*)[RING_SLOTS][DVDA_LANES]>(dyn_smem);
    uint16_t (*lut)[512] = reinterpret_cast<uint16_t (*)[512]>(dyn_smem + DEC_WARPS * RING_SLOTS * DVDA_LANES);
    huff_lut_to_shared(lut);
    __syncthreads();
    DecodeJob job;
    uint32_t nch;
    if (!pair_job(m, work, warp0 + wib, lane, job, nch)) return;
    if (a >= m.segs[job.seg].n_au) return;
    const uint32_t rs = (uint32_t)__cvta_generic_to_shared(&ring[wib][0][lane]);
    switch (nch) {
    case 1: entropy_au<1>(m, job, a, lut, rs); break;
    case 2: entropy_au<2>(m, job, a, lut, rs); break;
    case 3: entropy_au<3>(m, job, a, lut, rs); break;
    case 4: entropy_au<4>(m, job, a, lut, rs); break;
    default: break;
    }
}

// A segment that needs its predecessor's FIR history is decoded by the complete decoder
// (+ k_carry_fix), which reads the predecessor's stored tail: have the complete decoder
// produce that one too (the output pass would store it too late).
// (the two sets only OR the one bit into other segments' words: their order does not matter)
__global__ void k_flag_fallbacks(MlpTables m)
{
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    {
        const uint32_t seg = idx >> 1, k = idx & 1;
        if (seg < m.cnt->nseg) {
            const uint32_t fl = m.ss_flags[k * m.cap_seg + seg];
            const TrackDev &T = m.tracks[m.segs[seg].track];
            // (does the complete decoder have anything to do?  It and the carry fix leave at once if not.)
            if (k < T.nss && (fl & (SEG_FALLBACK | SEG_WANTS_PREV))) *m.any_fallback = 1;
            if ((fl & SEG_WANTS_PREV) && seg > T.seg_base) {
                // ... and where the predecessor may deliver next to nothing (no access unit, or its first one
                // dropped or damaged: the check data has been through), the history reaches further back
                // (k_carry_fix takes such segments along): up to the first one that stands on its own
                uint32_t p = seg - 1;
                atomicOr(&m.ss_flags[k * m.cap_seg + p], SEG_FALLBACK);
                while (p > T.seg_base && (m.segs[p].n_au == 0 || m.au_err[m.segs[p].au_base] != 0)) {
                    p--;
                    atomicOr(&m.ss_flags[k * m.cap_seg + p], SEG_FALLBACK);
                }
            }
        }
    }
    const uint32_t A = idx;
    if (A < m.cnt->nau && m.au_err[A]) {
        const uint32_t seg = m.au_seg[A];
        atomicOr(&m.ss_flags[seg], SEG_FALLBACK);
        atomicOr(&m.ss_flags[m.cap_seg + seg], SEG_FALLBACK);
        *m.any_fallback = 1;
    }
}

size_t au_snap_bytes() { return sizeof(AuSnap); }
size_t seg_ctx_bytes() { return sizeof(SegCtx); }
size_t au_delta_bytes() { return sizeof(AuDelta) + 4 * sizeof(ChanCoef) + sizeof(MatCoef); }
void au_delta_split(void *base, size_t entries, MlpTables &m)
{
    uint8_t *p = static_cast<uint8_t *>(base);
    m.au_delta = reinterpret_cast<AuDelta *>(p);
    // (a pair of channels' coefficients = one 128-byte line)
    uint8_t *cf = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(p + entries * sizeof(AuDelta)) + 127) & ~(uintptr_t)127);
    m.au_cf = reinterpret_cast<ChanCoef *>(cf);
    m.au_mcoef = reinterpret_cast<MatCoef *>(cf + entries * 4 * sizeof(ChanCoef));
}

const uint16_t *huff_lut_device()
{
    void *p = nullptr;
    if (cudaGetSymbolAddress(&p, g_huff_lut) != cudaSuccess) return nullptr;
    return reinterpret_cast<const uint16_t *>(p);
}

// The fast path up to the tiles: header passes A0 .. A2, entropy pass B, then the flags that hand
// segments to the complete decoder.  cap_pairs: (group, substream) pairs the grids cover;
// lim_max_au: access units per segment the entropy grid covers; lim_nss: substreams the parse grid covers.
int launch_mlp_fast(MlpTables m, const DecWork *work, uint32_t cap_pairs, uint32_t lim_max_au, uint32_t lim_nss,
                    cudaEvent_t (*kev)[2], bool *kev_used, cudaEvent_t checked, cudaStream_t s)
{
    static PerDeviceOnce attr_once;
    if (attr_once.run([&]() -> int { CUDA_TRY(cudaFuncSetAttribute(k_mlp_entropy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DEC_SMEM_BYTES)); return 0; })) return -1;
    static const int slot[4] = {DVDAGPU_K_MLP_SEGCTX, DVDAGPU_K_MLP_AU_PARSE, DVDAGPU_K_MLP_RESOLVE, DVDAGPU_K_MLP_ENTROPY};
    const uint32_t small = div_up_u32(cap_pairs, 4);
    for (int pass = 0; pass < 4 && cap_pairs; pass++) {
        CUDA_TRY(record_timing(kev[slot[pass]][0], s));
        if (pass == 0) LAUNCH(k_mlp_segctx, small, 128, 0, s, m, work);
        else if (pass == 1) LAUNCH(k_mlp_au_parse, dim3(div_up_u32(m.cap_au, GRD_THREADS), lim_nss), GRD_THREADS, 0, s, m);
        else if (pass == 2) LAUNCH(k_mlp_resolve, small, 128, 0, s, m, work);
        else LAUNCH(k_mlp_entropy, dim3(div_up_u32(cap_pairs, DEC_WARPS), lim_max_au ? lim_max_au : 1), DEC_WARPS * 32, DEC_SMEM_BYTES, s, m, work);
        CUDA_TRY(record_timing(kev[slot[pass]][1], s));
        kev_used[slot[pass]] = true;
    }
    // check data ran beside all this: segments with a damaged or dropped access unit (parity, CRC,
    // changed stream parameters) go to the complete decoder, which knows where such a track ends —
    // as do the predecessors of segments that want their FIR history (one launch for both)
    CUDA_TRY(cudaStreamWaitEvent(s, checked, 0));
    if (m.cap_seg || m.cap_au) LAUNCH(k_flag_fallbacks, div_up_u32(max((uint64_t)m.cap_seg * 2, (uint64_t)m.cap_au), 256), 256, 0, s, m);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int launch_mlp_decode(MlpTables m, const DecWork *work, uint32_t cap_pairs, cudaStream_t s)
{
    if (!cap_pairs) return 0;
    static PerDeviceOnce attr_once;
    if (attr_once.run([&]() -> int { CUDA_TRY(cudaFuncSetAttribute(k_mlp_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DEC_SMEM_BYTES)); return 0; })) return -1;
    LAUNCH(k_mlp_decode, div_up_u32(cap_pairs, DEC_WARPS), DEC_WARPS * 32, DEC_SMEM_BYTES, s, m, work);
    CUDA_TRY(cudaGetLastError());
    return 0;
}


// Segments whose first filtered block needs FIR history from the previous
// segment (the reference never clears it, mlp.c:948-952), and segments the fast
// path could not take (unexpected channel split), are decoded again with the
// generic routine, in order, run by run: one thread per (run head, substream).
#define FIX_THREADS 32
__global__ void __launch_bounds__(FIX_THREADS) k_carry_fix(MlpTables m)
{
    __shared__ __align__(16) uint16_t lut[4][512];
    __shared__ uint4 ring[RING_SLOTS][DVDA_LANES];
    if (m.fast && !*m.any_fallback) return;              // nothing was handed to the complete decoder
    huff_lut_to_shared(lut);
    __syncthreads();
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t seg = idx >> 1, k = idx & 1;
    if (seg >= m.cnt->nseg) return;
    const TrackDev &T = m.tracks[m.segs[seg].track];
    if (k >= T.nss) return;
    const uint32_t *fl = m.ss_flags_prev + (uint64_t)k * m.cap_seg;   // as they were before any fix-up
    // A segment that delivered fewer than 8 frames (a dropped access unit, a track's stub) does not
    // own the FIR history behind it: what its successor needs reaches back into the segment before.
    // Such segments travel with the run — decoded again from their predecessor's tail they hand on
    // the right one — so a run is a stretch of segments that need a carry or are that short.
    auto in_run = [&](uint32_t s) { return (fl[s] & SEG_NEEDS_CARRY) != 0 || m.segs[s].frames < 8; };
    if (!in_run(seg)) return;
    // run head: predecessor (same track) is not part of a run itself
    if (seg > T.seg_base && in_run(seg - 1)) return;
    const uint32_t track_end = T.seg_base + T.nseg;
    {
        // (a run in which nobody needs the history is left alone)
        bool any = false;
        for (uint32_t s = seg; s < track_end && in_run(s) && !any; s++) any = (fl[s] & SEG_NEEDS_CARRY) != 0;
        if (!any) return;
    }
    if (seg == T.seg_base && (T.cont & TRACK_CONT_PREV)) {
        // the history lives in a part decoded elsewhere: the caller has to decode the parts together
        m.tracks[m.segs[seg].track].stopped = 2;
        return;
    }
    const uint32_t rs = (uint32_t)__cvta_generic_to_shared(&ring[0][threadIdx.x & 31]);
    for (uint32_t s = seg; s < track_end && in_run(s); s++) {
        DecodeJob job;
        job.seg = s; job.k = k; job.lane = (s - T.seg_base) % DVDA_LANES; job.exact_history = true;
        const int32_t *prev = m.fir_tail + ((uint64_t)k * m.cap_seg + (s - 1)) * (DVDA_MAX_CH * 8);
        // parsing does not depend on filter history, so the error bookkeeping of the
        // first pass (merged with atomics) is reproduced exactly
        decode_segment<0>(m, job, lut, rs, s > T.seg_base ? prev : nullptr);
    }
}

int launch_carry_fix(MlpTables m, cudaStream_t s)
{
    if (!m.cap_seg) return 0;
    LAUNCH(k_carry_fix, div_up_u32((uint64_t)m.cap_seg * 2, FIX_THREADS), FIX_THREADS, 0, s, m);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// ----------------------------------------------------------- bookkeeping

// one thread per segment: merge the substreams' verdicts, count frames
__global__ void k_seg_finalize(MlpTables m, uint32_t *__restrict__ seg_frames, uint32_t *__restrict__ status)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.cap_seg) return;
    if (i >= m.cnt->nseg) { seg_frames[i] = 0; return; }  // rows behind the last segment count nothing
    SegDev &S = m.segs[i];
    TrackDev &T = m.tracks[S.track];
    const uint32_t now = m.ss_flags[i] | (T.nss == 2 ? m.ss_flags[m.cap_seg + i] : 0);
    uint32_t flags = S.flags | now;
    uint32_t err = S.err, stop = S.err_au;
    uint32_t frames = 0;
    const uint32_t lim = min(stop, S.n_au);
    for (uint32_t a = 0; a < lim; a++) {
        const uint32_t A = S.au_base + a;
        const uint32_t nf = m.au_frames_ss[A];
        if (T.nss == 2 && m.au_frames_ss[m.cap_au + A] != nf) { err |= ERR_SYNTAX; stop = a; break; }
        frames += nf;
    }
    if ((flags & SEG_IRREGULAR) && stop == 0xFFFFFFFFu) { err |= ERR_SYNTAX; stop = S.n_au; }
    S.flags = flags & ~SEG_NEEDS_CARRY;
    // a segment longer than its tile: the decode is repeated with the frame count found here
    // (frames that do not count — behind an error — may have run over as well: no reason to come back)
    if ((now & SEG_OVERFLOW) && frames > m.groups[T.grp_base + (i - T.seg_base) / DVDA_LANES].cap) {
        atomicOr(status, SEG_OVERFLOW);
        m.seg_need[i] = max(m.seg_need[i], frames);
    }
    // anything left for k_rematrix once the fused filter + output pass has run?
    if (m.fast && frames && !seg_output_done(m, T, i)) atomicOr(status, STATUS_WANTS_REMATRIX);
    S.err = err;
    S.err_au = stop;
    S.frames = frames;
    seg_frames[i] = frames;
    if (stop != 0xFFFFFFFFu) {
        atomicMax(&T.stopped, 1u);
        atomicMin(&T.err_seg, i - T.seg_base);
        if (err) atomicOr((unsigned int *)&T.error_flags, err);
    }
}

// one thread per segment: position in the track's output; thread of the first
// segment also totals the track
__global__ void k_track_finalize(MlpTables m, const uint64_t *__restrict__ scan, const uint32_t *__restrict__ status)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.cnt->nseg) return;
    if (*status & SEG_OVERFLOW) return;                   // the batch is decoded once more: leave the counts alone
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
    if (!m.cap_seg) return 0;
    LAUNCH(k_seg_finalize, div_up_u32(m.cap_seg, 128), 128, 0, s, m, seg_frames, status);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
int launch_track_finalize(MlpTables m, const uint64_t *seg_frame_scan, const uint32_t *status, cudaStream_t s)
{
    if (!m.cap_seg) return 0;
    LAUNCH(k_track_finalize, div_up_u32(m.cap_seg, 128), 128, 0, s, m, seg_frame_scan, status);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------- rematrix

#define RM_THREADS 256

// A block takes (group, 32-frame chunk) items in turn (the grid is fixed: how many groups and
// chunks there are is known on the device only).  It loads the
// [32 frames][nch][32 lanes] patch of the tile coalesced, then every warp takes
// segments (lanes of the patch) and its 32 threads take the 32 frames: noise,
// matrices, bypass, shift, channel order, interleaved store.
//
// Per frame the access unit is found by division (AUs normally have the nominal
// length; a binary search covers the rest), and an AU whose parameters are
// trivial — no matrix, no output shift, identity channel order — skips the
// parameter set altogether: the kernel is then a pure transpose at HBM speed.
// access unit holding frame F of a segment
__device__ __forceinline__ AuDev au_of_frame(const MlpTables &m, const SegDev &S, uint32_t F, uint32_t nominal)
{
    const uint32_t n_ok = min(S.n_au, S.err_au);
    const uint32_t ai = min(F / nominal, n_ok - 1);
    AuDev au = m.au[S.au_base + ai];
    if (F < au.frame0 || F >= au.frame0 + au.nframes) {
        // irregular lengths or dropped AUs: last one with frame0 <= F (dropped AUs have no
        // frames and share frame0 with their successor)
        uint32_t lo = 0, hi = n_ok;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (m.au[S.au_base + mid].frame0 <= F) lo = mid; else hi = mid;
        }
        au = m.au[S.au_base + lo];
    }
    return au;
}

// one (group, chunk) item; every return is taken by the whole block
template <int NCH>
__device__ __forceinline__ void rematrix_chunk(const MlpTables &m, const GroupDev &G, const TrackDev &T, uint32_t f0, int32_t *sm)
{
    // sm: [nch][32][33] samples, then [32][33] bypass bytes as ints
    const uint32_t nch = NCH ? NCH : T.channels;
    const uint32_t nf = min(32u, G.cap - f0);
    int32_t *bsm = sm + nch * 32 * 33;
    if (m.fast) {
        // segments the fused output pass has written need nothing here
        const int left = (threadIdx.x < G.nseg) && !seg_output_done(m, T, G.seg0 + threadIdx.x);
        if (!__syncthreads_or(left)) return;
    }

    // coalesced load: consecutive threads read consecutive lanes
    const int32_t *src = m.tiles + G.tile_off + (uint64_t)f0 * nch * DVDA_LANES;
    for (uint32_t i = threadIdx.x; i < nf * nch * 32; i += RM_THREADS) {
        const uint32_t l = i & 31, c = (i >> 5) % nch, f = (i >> 5) / nch;
        sm[(c * 32 + f) * 33 + l] = src[i];
    }
    const uint8_t *bsrc = m.bypass + G.byp_off + (uint64_t)f0 * DVDA_LANES;
    for (uint32_t i = threadIdx.x; i < nf * 32; i += RM_THREADS) bsm[(i >> 5) * 33 + (i & 31)] = bsrc[i];
    const uint32_t nominal = T.au_nominal;
    // The noise generator steps once per frame from the seed at the start of the access unit.
    // Its state at the first frame of the chunk is worked out once per segment here (up to an
    // access unit of steps); a frame then needs at most 31 more.
    __shared__ uint32_t chunk_seed[32];
    if (threadIdx.x < G.nseg) {
        const SegDev &S = m.segs[G.seg0 + threadIdx.x];
        uint32_t seed = 0;
        if (f0 < S.frames && min(S.n_au, S.err_au) > 0) {
            const AuDev au = au_of_frame(m, S, f0, nominal);
            if (m.psets[au.pset & 0x7FFFFFFFu].uses_noise) {
                seed = au.seed;
                for (uint32_t i = au.frame0; i < f0; i++) seed = noise_step(seed);
            }
        }
        chunk_seed[threadIdx.x] = seed;
    }
    __syncthreads();

    const uint32_t f = threadIdx.x & 31;                 // frame inside the chunk
    const bool plain_order = !(T.assignment >= 0x12 && T.assignment <= 0x14);
    for (uint32_t l = threadIdx.x >> 5; l < G.nseg; l += RM_THREADS / 32) {
        const SegDev &S = m.segs[G.seg0 + l];
        const uint32_t F = f0 + f;                       // frame inside the segment
        if (F >= S.frames) continue;
        // already written by the fused filter + output pass of the fast path?
        if (seg_output_done(m, T, G.seg0 + l)) continue;
        const AuDev au = au_of_frame(m, S, F, nominal);
        int32_t v[DVDA_MAX_CH];
#pragma unroll
        for (uint32_t c = 0; c < DVDA_MAX_CH; c++) v[c] = (c < nch) ? sm[(c * 32 + f) * 33 + l] : 0;
        int32_t *dst = m.pcm + T.out_base + (S.frame0 + F) * nch;
        if ((au.pset & 0x80000000u) && plain_order) {
            // trivial parameters: straight copy
            if (NCH == 2) {
                *reinterpret_cast<int2 *>(dst) = make_int2(v[0], v[1]);     // out_base and nch are even
            } else {
#pragma unroll
                for (uint32_t c = 0; c < DVDA_MAX_CH; c++) if (c < nch) dst[c] = v[c];
            }
            continue;
        }
        const ParamSet &P = m.psets[au.pset & 0x7FFFFFFFu];
        const uint32_t ml = P.matrix_len;
        if (ml) {
            int32_t n0 = 0, n1 = 0;
            if (P.uses_noise) {
                // from the chunk's first frame if it lies in the same access unit, else from the unit's start
                const bool same = au.frame0 <= f0;
                uint32_t seed = same ? chunk_seed[l] : au.seed;
                for (uint32_t i = same ? f0 : au.frame0; i < F; i++) seed = noise_step(seed);
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

#define RM_SMEM_BYTES ((DVDA_MAX_CH + 1) * 32 * 33 * 4)
__global__ void __launch_bounds__(RM_THREADS) k_rematrix(MlpTables m)
{
    extern __shared__ int32_t rm_sm[];
    // with the fast path: only if some segment was left for this pass (k_seg_finalize)
    if (m.fast && !(*m.status & STATUS_WANTS_REMATRIX)) return;
    if (*m.status & (SEG_OVERFLOW | STATUS_PCM_SMALL)) return;
    const uint32_t chunks = m.cnt->max_chunks;
    const uint64_t items = (uint64_t)m.cnt->ngroups * chunks;
    for (uint64_t it = blockIdx.x; it < items; it += gridDim.x) {
        const GroupDev &G = m.groups[it / chunks];
        const uint32_t f0 = (uint32_t)(it % chunks) * 32;
        if (f0 < G.cap) {
            const TrackDev &T = m.tracks[G.track];
            if (T.channels == 1) rematrix_chunk<1>(m, G, T, f0, rm_sm);
            else if (T.channels == 2) rematrix_chunk<2>(m, G, T, f0, rm_sm);
            else rematrix_chunk<0>(m, G, T, f0, rm_sm);
        }
        __syncthreads();                                  // the patch is overwritten by the next item
    }
}

int launch_rematrix(MlpTables m, cudaStream_t s)
{
    if (!m.cap_grp) return 0;
    static PerDeviceOnce attr_once;
    if (attr_once.run([&]() -> int { CUDA_TRY(cudaFuncSetAttribute(k_rematrix, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RM_SMEM_BYTES)); return 0; })) return -1;
    LAUNCH(k_rematrix, 148 * 4, RM_THREADS, RM_SMEM_BYTES, s, m);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
