// mlp_fused.cu — the fused MLP pass: residual entropy decode + prediction filters + rematrix +
// output shift + interleaved PCM, in one kernel, with no intermediate in HBM.
//
// Replaces (reference tree, src/mlp.c unless noted), for everything the header passes of
// mlp_decode.cu (A0 .. A2) have resolved:
//   :1122-1241  decode_residual_data + src/bitstream.c:1806-1833 br_read_huffman_code
//   :1243-1306  filter_channel
//   :1308-1358  rematrix_channels, :504-538 / :575-608 output shift, RIFF WAVE order
//   src/dvd-audio.c:781-792  dvda_read interleave
//
// Why one pass.  The three-pass path decodes the residuals of every access unit into a tile in
// HBM (one lane per access unit), then runs the filter recurrences over the tile (one lane per
// channel): every sample crosses HBM twice more than it has to, and the filter pass sits waiting
// for its loads.  What really chains is only this: the bit position inside an access unit, and a
// channel's filter history across the access units of a segment.  So a lane here is one
// (segment, channel): it walks its segment access unit by access unit, frame by frame, and per
// frame steps over the codes of the *other* channels of its substream (a table look-up for the
// length, nothing else), decodes its own residual, and feeds it straight into its filter.  The
// look-ups of the other channels are redundant work — a few instructions per code — and buy a
// kernel with no tile traffic at all: the elementary stream is read once (through per-lane rings
// in shared memory, filled by cp.async), interleaved PCM is written once (bulk copies shared ->
// global).
//
// Lanes.  A segment of a track with n0 channels in substream 0 and n1 in substream 1 takes
// n0 + n1 neighbouring lanes (n1 = 0: one substream); a warp takes 32 / (n0 + n1) segments of a
// group.  Lanes of substream 0 read substream 0's bits, lanes of substream 1 theirs; both run
// through the frames of an access unit together, so the channels of a frame meet by shuffle
// where matrices ask for them (substream 1's matrices govern all channels, mlp.c:575-595).
//
// What this pass cannot take it notices while decoding (a block that brings parameters in the
// middle of an access unit, an invalid code, bits running past the substream): it flags the
// segment in ss_sticky and raises STATUS_REDO; the host repeats the decode stage with the
// segment handed to the complete decoder.
#include "mlp_common.cuh"
#include "../../include/dvdagpu.h"

#define FUSE_WARPS 4
#ifndef FUSE_MIN_BLOCKS
#define FUSE_MIN_BLOCKS 3
#endif
#define FUSE_PF 16                                   // frames per output patch (a row leaves as one bulk copy)
#define FUSE_PATCH_WORDS (FUSE_PF * 32 + 4 * 32)     // most a patch needs: spw rows of PF * lps + 4 words
#define FUSE_WARP_WORDS (2 * FUSE_PATCH_WORDS + 4 * 32)
#define FUSE_RING_BYTES (FUSE_WARPS * RING_SLOTS * DVDA_LANES * 16)
#define FUSE_LUT_BYTES (4 * 512 * 2)
#define FUSE_SMEM_BYTES (FUSE_RING_BYTES + FUSE_LUT_BYTES + FUSE_WARPS * FUSE_WARP_WORDS * 4)

uint32_t fused_warps_per_group(uint32_t n0, uint32_t n1)
{
    const uint32_t lps = n0 + n1, spw = 32 / lps;
    return (32 + spw - 1) / spw;
}

__device__ __forceinline__ uint32_t lds_u16(uint32_t addr)
{
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}

// M: most channels of one substream this instantiation steps over per frame (2 or 4)
template <int M>
__global__ void __launch_bounds__(FUSE_WARPS * 32, FUSE_MIN_BLOCKS)
k_mlp_fused(MlpTables m, const FusedWork *__restrict__ work, uint32_t n_work, uint32_t n_warps)
{
    extern __shared__ uint4 fuse_sm[];
    uint4 (*ring)[RING_SLOTS][DVDA_LANES] = reinterpret_cast<uint4 (*)[RING_SLOTS][DVDA_LANES]>(fuse_sm);
    uint16_t *lut = reinterpret_cast<uint16_t *>(fuse_sm + FUSE_RING_BYTES / 16);
    int32_t *patches = reinterpret_cast<int32_t *>(fuse_sm + (FUSE_RING_BYTES + FUSE_LUT_BYTES) / 16);
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(m.huff_lut);
        uint4 *dst = reinterpret_cast<uint4 *>(lut);
        for (uint32_t i = threadIdx.x; i < FUSE_LUT_BYTES / 16; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const uint32_t wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t warp = blockIdx.x * FUSE_WARPS + wib;
    if (warp >= n_warps) return;
    // queued before the host has seen the batch's status: nothing to do if the batch is decoded
    // once more (tile overflow) or the output buffer sized in advance turned out too small
    if (*m.status & (SEG_OVERFLOW | STATUS_PCM_SMALL)) return;

    // ---- which track, group, segment, substream, channel
    uint32_t lo = 0, hi_ = n_work;
    while (hi_ - lo > 1) {
        const uint32_t mid = (lo + hi_) >> 1;
        if (work[mid].warp0 <= warp) lo = mid; else hi_ = mid;
    }
    const FusedWork W = work[lo];
    const TrackDev &T = m.tracks[W.track];
    const uint32_t n0 = W.n0, lps = W.n0 + W.n1;            // lanes per segment
    const uint32_t spw = 32 / lps;                         // segments per warp
    const uint32_t sub_n = (32 + spw - 1) / spw;           // warps per group
    const uint32_t rel = warp - W.warp0;
    const GroupDev &G = m.groups[T.grp_base + rel / sub_n];
    const uint32_t sub = rel % sub_n;
    const uint32_t sl = lane / lps, j = lane - sl * lps;   // segment in the warp, channel of the track
    const uint32_t k = j >= n0 ? 1u : 0u;                  // substream
    const uint32_t cc = k ? j - n0 : j;                    // channel inside the substream
    const uint32_t nck = k ? W.n1 : n0;                    // channels of the substream
    const uint32_t sg = sub * spw + sl;                    // segment inside the group
    const bool have = sl < spw && sg < G.nseg;
    const uint32_t seg = G.seg0 + (have ? sg : 0);
    const SegDev &S = m.segs[seg];
    uint32_t seg_flags = m.ss_flags_fast[seg];
    if (W.n1) seg_flags |= m.ss_flags_fast[m.nseg + seg];
    const bool mine = have && !(seg_flags & SEG_FALLBACK) && S.frames > 0;
    const uint32_t my_frames = mine ? S.frames : 0;
    const uint32_t max_frames = __reduce_max_sync(0xFFFFFFFFu, my_frames);
    if (!max_frames) return;
    const uint32_t nominal = T.au_nominal;

    const uint32_t row_words = FUSE_PF * lps + 4;          // a segment's row in a patch (+4: 16-byte aligned, banks spread)
    const uint32_t patch_words = spw * row_words;
    int32_t *patch = patches + (size_t)wib * FUSE_WARP_WORDS;
    uint32_t *meta = reinterpret_cast<uint32_t *>(patch + 2 * FUSE_PATCH_WORDS);
    if (j == 0 && sl < spw) {
        const uint64_t base = (mine ? S.frame0 : 0) * lps;
        meta[sl * 4 + 0] = (uint32_t)base; meta[sl * 4 + 1] = (uint32_t)(base >> 32); meta[sl * 4 + 2] = my_frames;
        meta[sl * 4 + 3] = ((T.out_base + base) & 3) == 0;        // the rows of this segment start on 16-byte boundaries
    }
    __syncwarp();

    const AuSnap *snaps = m.au_snap + (uint64_t)k * m.nau;
    const AuDelta *deltas = m.au_delta + (uint64_t)k * m.nau;
    const uint8_t *fchg = m.au_fchg + (uint64_t)k * m.nau;
    const uint32_t au_base = mine ? S.au_base : 0;

    // ---- per-lane state
    Rd b;
    rd_init(b, m.es, (uint32_t)__cvta_generic_to_shared(&ring[wib][0][lane]));
    const uint32_t lut_s = (uint32_t)__cvta_generic_to_shared(lut);
    int32_t fh[8], ih[8], cf[8], ci[8];
#pragma unroll
    for (int t = 0; t < 8; t++) { fh[t] = 0; ih[t] = 0; cf[t] = 0; ci[t] = 0; }
    FiltSetup F = {0, 0, 0, 0, 0};
    uint32_t shift = 0, qmask = 0xFFFFFFFFu, q = 0, cls = 0;
    int32_t sho = 0;
    uint32_t lsbs[M], lutb[M];                             // of the substream's channels: LSB bits, table of the codebook
#pragma unroll
    for (int c = 0; c < M; c++) { lsbs[c] = 0; lutb[c] = lut_s; }
    uint32_t want = 0, blk = 8, blk_left = 0, end_bits = 0, bad = 0, au_done = 1, failed = 0;
    uint32_t seed = 0, pset = 0xFFFFFFFFu, f = 0, a = 0;
    const ParamSet *P = nullptr;
    bool trivial = true;
    int32_t *const pcm_row = m.pcm + T.out_base;
    const bool plain_order = !(T.assignment >= 0x12 && T.assignment <= 0x14);
    const uint32_t out_slot = wave_slot(T.assignment, j);
    const uint32_t group_lane0 = sl * lps;                 // first lane of this segment
    const uint32_t gov_lane = group_lane0 + (W.n1 ? n0 : 0);   // a lane of the governing substream
    int32_t *const park = patch + (sl < spw ? sl : 0) * row_words + out_slot;

    // The patch of FUSE_PF frames leaves row by row (a row = the frames of one segment, contiguous
    // in the output).  Whole, 16-byte aligned rows go out as bulk copies shared -> global issued by
    // the row's lane; the copy engine reads the patch while the warp fills the other one.
    auto flush = [&](uint32_t f0) {
        const int32_t *pb = patch + ((f0 / FUSE_PF) & 1) * FUSE_PATCH_WORDS;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the parked samples, for the async proxy
        __syncwarp();
        uint32_t slow = 0;
        if (lane < spw) {
            const uint4 mt = *reinterpret_cast<const uint4 *>(meta + lane * 4);     // base lo, hi, frames, aligned
            if (f0 < mt.z) {
                if (f0 + FUSE_PF <= mt.z && mt.w) {
                    int32_t *dst = pcm_row + (((uint64_t)mt.y << 32 | mt.x) + (uint64_t)f0 * lps);
                    const uint32_t src = (uint32_t)__cvta_generic_to_shared(pb + lane * row_words);
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                 :: "l"(dst), "r"(src), "r"(FUSE_PF * 4 * lps) : "memory");
                } else slow = 1;
            }
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        uint32_t rows = __ballot_sync(0xFFFFFFFFu, slow);
        while (rows) {
            const uint32_t row = __ffs(rows) - 1;
            rows &= rows - 1;
            const uint4 mt = *reinterpret_cast<const uint4 *>(meta + row * 4);
            const uint32_t n = min((uint32_t)FUSE_PF, mt.z - f0) * lps;
            int32_t *dst = pcm_row + (((uint64_t)mt.y << 32 | mt.x) + (uint64_t)f0 * lps);
            for (uint32_t i = lane; i < n; i += 32) dst[i] = pb[row * row_words + i];
        }
        // the patch written one flush ago has been read by now: it is the one filled next
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
    };

    constexpr uint32_t need8 = (8 * (M * 33 + 6) + 31) / 32 + 2;   // words eight frames can consume, plus the one fetched ahead

    while (f < max_frames) {
        // ---- next access unit: where its residuals are, this lane's entropy and filter parameters
        const bool au_act = f < my_frames;
        if (au_act) {
            const uint32_t A = au_base + a;
            const uint32_t An = min(A + 1, m.nau);          // (the tables have one spare entry)
            prefetch_l1(&snaps[An]);
            prefetch_l1(&deltas[An].cf[cc]);
            prefetch_l1(reinterpret_cast<const uint8_t *>(&deltas[An].cf[cc]) + 32);
            // the snapshot: positions, block size, bypass mask | per channel {sho, cb, lsb_bits, q, shift}
            const uint64_t *sw = reinterpret_cast<const uint64_t *>(&snaps[A]);
            const uint64_t bit0 = sw[0], bit_end = sw[1], w2 = sw[2];
            // the previous access unit must have ended properly (checked here, one access unit late,
            // and after the last one: see below)
            if (!au_done || (bad & 0x8000) || rd_pos(b) > end_bits) failed = 1;
            blk = (uint32_t)w2 & 0xFFFFu;
            want = (uint32_t)(w2 >> 16) & 0xFFu;
            if ((uint32_t)(w2 >> 32 & 0xFF) + cc != j || (uint32_t)(w2 >> 40 & 0xFF) != nck) failed = 1;   // channel layout as expected?
#pragma unroll
            for (int c = 0; c < M; c++) {
                if ((uint32_t)c < nck) {
                    const uint64_t cw = sw[3 + c];
                    const uint32_t hi32 = (uint32_t)(cw >> 32);             // cb, lsb_bits, q, shift
                    lsbs[c] = (hi32 >> 8) & 0xFF;
                    lutb[c] = lut_s + (hi32 & 0xFF) * 1024;
                    if ((uint32_t)c == cc) { sho = (int32_t)(uint32_t)cw; q = (hi32 >> 16) & 0xFF; shift = hi32 >> 24; }
                }
            }
            qmask = 0xFFFFFFFFu << q;
            if ((fchg[A] >> cc) & 1) {
                const AuDelta &D = deltas[A];
                const uint32_t *hw = reinterpret_cast<const uint32_t *>(&D.ch[cc]);
                DeltaHead H;
                H.fchg = 0; H.w0 = 0; H.qv = 0; H.seed = 0; H.pset = 0;
                H.h1 = hw[1]; H.h2 = hw[2];
                filt_take_head(H, D, cc, F, cf, ci, ih);
                cls = F.fo | F.io << 4;
            }
            {
                const uint2 sp = *reinterpret_cast<const uint2 *>(&m.au[A].seed);
                seed = sp.x;
                if (sp.y != pset) {
                    pset = sp.y;
                    P = &m.psets[pset & 0x7FFFFFFFu];
                    trivial = (pset & 0x80000000u) && plain_order;
                }
            }
            // seat the reader on the first residual bit
            rd_seat(b, (bit0 >> 5) << 2);
            rd_issue_ahead(b);
            rd_skip(b, (uint32_t)(bit0 & 31));
            end_bits = (uint32_t)(bit_end - ((bit0 >> 5) << 5));
            blk_left = blk;
            bad = 0;
            au_done = 0;
        } else {
            cls = 0;
#pragma unroll
            for (int t = 0; t < 8; t++) { cf[t] = 0; ci[t] = 0; }
        }
        a++;
        const uint32_t nf = __reduce_max_sync(0xFFFFFFFFu, ((cls & 15) + 3) >> 2);
        const uint32_t ni = __reduce_max_sync(0xFFFFFFFFu, ((cls >> 4) + 3) >> 2);
        const uint32_t code = nf * 3 + ni;
        const bool any_matrix = __any_sync(0xFFFFFFFFu, au_act && !trivial);
        const bool any_want = __any_sync(0xFFFFFFFFu, au_act && want);

        for (uint32_t i = 0; i < nominal; i += 8) {
            int32_t r[8];
            uint32_t bm[8];                                 // bypass bits of the eight frames (read by every lane of the substream)
#pragma unroll
            for (int t = 0; t < 8; t++) { r[t] = 0; bm[t] = 0; }
            if (au_act) {
                // ---- entropy: eight frames of this lane's substream; the lane keeps its own channel's residual
                rd_prefetch(b, need8);
                rd_hot_begin(b);
                uint32_t hi = (uint32_t)(b.win >> 32), lw = (uint32_t)b.win;
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    if (any_want && want) {
                        b.win = ((uint64_t)hi << 32) | lw;
                        bm[t] = bypass_bits(b, want);
                        rd_hot_begin(b);
                        hi = (uint32_t)(b.win >> 32); lw = (uint32_t)b.win;
                    }
#pragma unroll
                    for (int c = 0; c < M; c++) {
                        if ((uint32_t)c < nck) {
                            if (b.avail <= 32) {
                                hi |= __funnelshift_rc(b.ahead, 0u, (uint32_t)b.avail);
                                lw = __funnelshift_rc(0u, b.ahead, (uint32_t)b.avail);
                                b.avail += 32;
                                b.next_w++;
                            }
                            b.ahead = rd_ring_word(b, b.next_w);
                            const uint32_t e = lds_u16(lutb[c] + ((hi >> 23) << 1));
                            bad |= e;
                            const uint32_t hl = (e >> 8) & 15;
                            hi = __funnelshift_l(lw, hi, hl); lw <<= hl;
                            const uint32_t nb = lsbs[c];
                            const int32_t lsb = (int32_t)((hi >> 1) >> (31 - nb));
                            hi = __funnelshift_l(lw, hi, nb); lw <<= nb;
                            b.avail -= hl + nb;
                            if ((uint32_t)c == cc) r[t] = (int32_t)((uint32_t)((int32_t)((e & 0xFF) << nb) + lsb + sho) << q);
                        }
                    }
                }
                b.win = ((uint64_t)hi << 32) | lw;
                // ---- end of a block?  (blocks are multiples of eight frames here)
                blk_left -= 8;
                if (blk_left == 0) {
                    const uint32_t last = rd_get(b, 1);
                    if (last) au_done = (i + 8 == nominal);            // the access unit must end with its last frame
                    else if (i + 8 >= nominal || rd_get(b, 1)) failed = 1;   // more frames than nominal, or a block with parameters
                    blk_left = blk;
                    if (last && i + 8 != nominal) failed = 1;
                }
            }
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
            if (any_matrix) {
                // the shuffles need the whole warp: lanes without work just run along
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    int32_t v[2 * M];
#pragma unroll
                    for (int c = 0; c < 2 * M; c++) v[c] = __shfl_sync(0xFFFFFFFFu, r[t], group_lane0 + c);
                    const uint32_t bmask = __shfl_sync(0xFFFFFFFFu, bm[t], gov_lane);
                    if (au_act && !trivial) {
                        // noise, matrices in order, bypass bit, output shift (mlp.c:504-538, 1308-1358)
                        const uint32_t sh = (seed >> 7) & 0xFFFF;
                        const int32_t z0 = (int32_t)((uint32_t)(int32_t)(int8_t)(seed >> 15) << P->noise_shift);
                        const int32_t z1 = (int32_t)((uint32_t)(int32_t)(int8_t)sh << P->noise_shift);
                        const uint32_t ml = P->matrix_len, mmc = P->mmc;
                        for (uint32_t mk = 0; mk < ml; mk++) {
                            long long sum = 0;
#pragma unroll
                            for (int c = 0; c < 2 * M; c++) if ((uint32_t)c <= mmc && (uint32_t)c < lps) sum += (long long)v[c] * P->coeff[mk][c];
                            sum += (long long)z0 * P->coeff[mk][mmc + 1];
                            sum += (long long)z1 * P->coeff[mk][mmc + 2];
                            const uint32_t oc = P->out_ch[mk], qq = P->q[oc];
                            const int32_t rr = (((int32_t)(sum >> 14)) >> qq << qq) + (int32_t)((bmask >> mk) & 1);
#pragma unroll
                            for (int c = 0; c < 2 * M; c++) if ((uint32_t)c == oc) v[c] = rr;
                        }
                        int32_t mineval = 0;
#pragma unroll
                        for (int c = 0; c < 2 * M; c++) if ((uint32_t)c == j) mineval = v[c];
                        r[t] = j <= mmc ? (int32_t)((uint32_t)mineval << P->out_shift[j]) : mineval;
                    }
                    seed = noise_step(seed);
                }
            }
            if (au_act) {
                int32_t *pk = park + ((f / FUSE_PF) & 1) * FUSE_PATCH_WORDS + (f & (FUSE_PF - 1)) * lps;
#pragma unroll
                for (int t = 0; t < 8; t++) pk[t * lps] = r[t];
            }
            f += 8;
            if ((f & (FUSE_PF - 1)) == 0) flush(f - FUSE_PF);
        }
    }
    if (f & (FUSE_PF - 1)) flush(f & ~(uint32_t)(FUSE_PF - 1));
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");    // shared memory is given back at exit
    cp_wait<0>();
    if (mine) {
        // the last access unit's end
        if (!au_done || (bad & 0x8000) || rd_pos(b) > end_bits) failed = 1;
        if (failed) {
            // not for this pass after all: the decode stage runs again with the segment flagged
            atomicOr(&m.ss_sticky[seg], SEG_FALLBACK);
            atomicOr(&m.ss_sticky[m.nseg + seg], SEG_FALLBACK);
            atomicOr(m.status_rw, STATUS_REDO);
        }
        // FIR tail for a following segment that needs it
        int32_t *tail = m.fir_tail + ((uint64_t)k * m.nseg + seg) * (DVDA_MAX_CH * 8);
#pragma unroll
        for (int t = 0; t < 8; t++) tail[j * 8 + t] = fh[7 - t];
    }
}

template <int M>
static int launch_one_fused(MlpTables m, const FusedWork *work, uint32_t n_work, uint32_t n_warps, cudaStream_t s)
{
    if (!n_warps) return 0;
    static PerDeviceOnce attr_once;
    if (attr_once.run([&]() -> int { CUDA_TRY(cudaFuncSetAttribute(k_mlp_fused<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FUSE_SMEM_BYTES)); return 0; })) return -1;
    LAUNCH(k_mlp_fused<M>, div_up_u32(n_warps, FUSE_WARPS), FUSE_WARPS * 32, FUSE_SMEM_BYTES, s, m, work, n_work, n_warps);
    return 0;
}

int launch_mlp_fused(MlpTables m, const FusedWork *const work[2], const uint32_t n_work[2], const uint32_t n_warps[2], cudaStream_t s)
{
    if (launch_one_fused<2>(m, work[0], n_work[0], n_warps[0], s)) return -1;
    if (launch_one_fused<4>(m, work[1], n_work[1], n_warps[1], s)) return -1;
    CUDA_TRY(cudaGetLastError());
    return 0;
}
