// engine.cu — host side of the CUDA engine: device buffers, the kernel sequence
// of one decode, and the C ABI declared in include/dvdagpu.h.
//
// One decode = one pass of the whole hot path over a batch of tracks that share
// a sector buffer:
//
//   demux   k_sector_count -> scan -> k_packet_fill -> scans -> k_es_gather
//   index   k_sync_scan x2 -> k_track_setup -> k_segment_fill -> k_au_chase x2
//           -> k_yield_* -> k_group_setup
//   decode  k_checkdata -> k_mlp_decode (-> k_carry_fix) -> k_seg_finalize
//   output  k_rematrix (MLP), k_pcm_unpack (PCM)
//
// The host only sizes buffers between stages (a handful of 4- or 8-byte
// read-backs); no sample or stream byte is touched by the CPU.
#include "common.cuh"
#include "kernels.cuh"
#include "../../include/dvdagpu.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define MAP_BYTES (1u << 20)      // mapped pinned staging area for small transfers

thread_local uint32_t g_launch_count = 0;
thread_local bool g_capturing = false;
thread_local bool g_profiling = false;

// ---- DVDAGPU_TRACE=1: time line of one decode (debug aid, not on by default) ---------------
thread_local bool g_trace_on = false;
#define TRACE_MAX 512
struct TraceLog { int n; cudaEvent_t ev[TRACE_MAX]; const char *what[TRACE_MAX]; double host_us[TRACE_MAX]; bool gpu[TRACE_MAX]; int made; };
static thread_local TraceLog g_trace = {0, {}, {}, {}, {}, 0};
static double host_now_us()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}
void trace_mark(const char *what, cudaStream_t s)
{
    TraceLog &t = g_trace;
    if (t.n >= TRACE_MAX) return;
    if (t.n >= t.made) { cudaEventCreate(&t.ev[t.n]); t.made = t.n + 1; }
    t.what[t.n] = what; t.host_us[t.n] = host_now_us(); t.gpu[t.n] = s != (cudaStream_t)-1;
    if (t.gpu[t.n]) cudaEventRecord(t.ev[t.n], s);
    t.n++;
}
static void trace_host(const char *what) { if (g_trace_on) trace_mark(what, (cudaStream_t)-1); }
static void trace_dump()
{
    TraceLog &t = g_trace;
    if (t.n < 2) return;
    cudaDeviceSynchronize();
    int first_gpu = -1;
    float prev = 0;
    for (int i = 0; i < t.n; i++) {
        if (!t.gpu[i]) { fprintf(stderr, "[trace] %-28s host %9.1f us\n", t.what[i], t.host_us[i] - t.host_us[0]); continue; }
        if (first_gpu < 0) first_gpu = i;
        float ms = 0;
        cudaEventElapsedTime(&ms, t.ev[first_gpu], t.ev[i]);
        fprintf(stderr, "[trace] %-28s host %9.1f us   done on device %9.1f us  (+%.1f)\n", t.what[i],
                t.host_us[i] - t.host_us[0], ms * 1e3, (ms - prev) * 1e3);
        prev = ms;
    }
}
static thread_local char g_error[512] = "";

void dvdagpu_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof g_error, fmt, ap);
    va_end(ap);
}

extern "C" const char *dvdagpu_last_error(void) { return g_error; }

int upload_crc_table(const uint8_t *t);

// a device buffer that only ever grows
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    uint32_t gen = 0;                         // counts allocations
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return 0;
        gen++;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            dvdagpu_set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
            return -1;
        }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T *as() { return reinterpret_cast<T *>(p); }
};

enum {
    B_SECTORS, B_SECTORS2, B_PCM2, B_SEC_CNT, B_SEC_BAD, B_SEC_BASE, B_BAD_PREFIX,
    B_PK_SECTOR, B_PK_OFF, B_PK_LEN, B_PK_CODEC, B_PK_PAD2, B_PK_PARAMS, B_PK_MLPLEN, B_PK_PCMF,
    B_PK_ES, B_PK_PF, B_PK_NONMLP, B_PK_NM_PREFIX, B_PK_STOP, B_PK_STOP_PREFIX, B_PK_YIELD,
    B_ES, B_SYNC_SLOTS, B_SYNC_CNT_RAW, B_SYNC_BASE_RAW, B_SYNC_BASE_VALID, B_RAW, B_VALID,
    B_TRACKS, B_TRK_PK_LO, B_TRK_SEG_BASE, B_TRK_GRP_BASE,
    B_SEGS, B_SEG_NAU, B_SEG_AU_BASE, B_AU_POS, B_AU_ERR, B_AU, B_PSETS, B_AU_FRAMES,
    B_SS_FLAGS, B_SS_FLAGS_PREV, B_SS_FLAGS_FAST, B_FIR_TAIL,
    B_GROUPS, B_GRP_CELLS, B_CELL_BASE,
    B_DEC_WORK, B_OUT_WORK, B_COUNTS, B_SEG_NEED, B_AU_SEG, B_AU_SNAP, B_FILT_SNAP, B_AU_FCHG, B_SEG_CTX, B_AU_DELTA, B_TILES, B_BYPASS, B_SEG_FRAMES, B_SEG_FRAME_SCAN, B_SCAN_TMP, B_STATUS, B_AU_NOTED, B_PCM,
    B_COUNT
};

// capacities and launch limits of one decode (see shape_from_input / shape_from_last)
struct DecShape {
    uint64_t rows = 0, sync = 0, seg = 0, grp = 0, au = 0, cells = 0, pcm_fixed = 0;
    uint32_t max_au = 0, nss = 0, out_warps = 0;
    bool any_pcm = true, any_mlp = true, windowed = true;
};

// one decode in flight on a context
struct DecJob {
    const uint8_t *d_sectors = nullptr;
    uint32_t n_sectors = 0, n_tracks = 0, nslots = 2;
    std::vector<dvdagpu_track_desc> descs;
    std::vector<uint32_t> order;
    DecShape sh;
    int attempt = 0;
    bool use_fast = true, small_tables = false, pcm_small_seen = false, active = false;
    uint32_t launches = 0;                    // kernels launched for this decode (all attempts)
};

struct dvdagpu_ctx {
    int device;
    cudaStream_t own_stream;
    cudaStream_t stream;
    cudaStream_t h2d_stream, d2h_stream;      // copy engines of the pipelined path
    cudaStream_t aux_stream;                  // check data runs beside the header passes
    cudaEvent_t aux_ev[2];
    cudaStream_t aux_stream_hi = nullptr;     // side stream at the chain's own priority
    // what the previous decode of this context found and was sized for: the next one is sized from it
    DecCounts last;
    DecShape last_shape;
    uint32_t last_sectors = 0, last_tracks = 0;
    bool have_last = false;
    uint32_t seg_need_cap = 0;                // rows of the seg_need table that has been cleared
    bool profiling = false;                   // per-stage / per-kernel CUDA events (dvdagpu_set_profiling)
    DecJob job;                               // the decode in flight (decode_begin .. decode_end)
    cudaEvent_t job_done = nullptr;
    uint8_t *hback = nullptr, *dback = nullptr;   // mapped pinned memory the read-back of a decode lands in
    size_t back_bytes = 0;
    dvdagpu_ctx *peer = nullptr;              // a second context on the same device (pipelined path: two parts decode at once)
    cudaGraphExec_t graph_exec = nullptr;     // the decode sequence as an executable graph (updated in place from decode to decode)
    bool graph_ok = true, warmed = false, use_graph = false;
    uint32_t scan_tmp_gen = 0;                // allocation of the scan buffer that has been cleared
    cudaEvent_t pev[3][2];                    // [upload, decode, download][slot]
    int pcm_slot;                             // which PCM buffer the next decode writes
    uint8_t *hmap, *dmap;                     // mapped pinned staging area (host / device alias)
    size_t map_used;
    DevBuf buf[B_COUNT];
    cudaEvent_t ev[6];
    cudaEvent_t kev[16][2];
    bool kev_used[16];
    dvdagpu_stats stats;
    uint64_t pcm_samples;
    std::vector<TrackDev> h_tracks;
    const uint16_t *huff_lut = nullptr;       // device address of the Huffman table (per device)
};

// ---- constant tables, derived (not copied) --------------------------------

// PCM chunk permutation from the layout rule: a chunk is one or
// two groups of samples; 16-bit groups hold big-endian samples, 24-bit groups all
// (high, middle) byte pairs then all low bytes.  tab[i] = little-endian byte slot
// of stream byte i.
static void build_pcm_tables(uint8_t tab[2][6][36])
{
    memset(tab, 0, 2 * 6 * 36);
    for (int b24 = 0; b24 < 2; b24++) {
        for (int ch = 1; ch <= 6; ch++) {
            const int n = 2 * ch;
            int order[12], cnt = 0, first;
            const bool two = (ch == 6) || (b24 && ch >= 3);
            if (!two) { for (int i = 0; i < n; i++) order[cnt++] = i; first = n; }
            else {
                const int hi = ch == 6 ? 4 : ch;
                for (int f = 0; f < 2; f++) for (int c = 2; c < hi; c++) order[cnt++] = f * ch + c;
                first = cnt;
                for (int f = 0; f < 2; f++) for (int c = 0; c < ch; c++) if (c < 2 || c >= hi) order[cnt++] = f * ch + c;
            }
            uint8_t *t = tab[b24][ch - 1];
            int i = 0;
            for (int g = 0; g < 2; g++) {
                const int a = g ? first : 0, b = g ? n : first;
                if (!b24) for (int k = a; k < b; k++) { t[i++] = (uint8_t)(order[k] * 2 + 1); t[i++] = (uint8_t)(order[k] * 2); }
                else {
                    for (int k = a; k < b; k++) { t[i++] = (uint8_t)(order[k] * 3 + 2); t[i++] = (uint8_t)(order[k] * 3 + 1); }
                    for (int k = a; k < b; k++) t[i++] = (uint8_t)(order[k] * 3);
                }
            }
        }
    }
}

static void build_crc8(uint8_t t[256])
{
    for (unsigned i = 0; i < 256; i++) {
        unsigned c = i;
        for (int k = 0; k < 8; k++) c = (c & 0x80) ? ((c << 1) ^ 0x63) & 0xFF : (c << 1) & 0xFF;
        t[i] = (uint8_t)c;
    }
}

// ---- context -----------------------------------------------------------------

extern "C" int dvdagpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" dvdagpu_ctx *dvdagpu_create(int device)
{
    g_error[0] = 0;
    int n = dvdagpu_device_count();
    if (n <= 0) { dvdagpu_set_error("no CUDA device: this engine has no CPU path"); return nullptr; }
    if (device < 0 || device >= n) { dvdagpu_set_error("device %d out of range (%d present)", device, n); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { dvdagpu_set_error("cudaSetDevice(%d) failed", device); return nullptr; }
    dvdagpu_ctx *c = new dvdagpu_ctx();
    c->device = device;
    c->pcm_samples = 0;
    memset(&c->stats, 0, sizeof c->stats);
    // The decode chain runs at the highest stream priority, the side work (check data) at the
    // lowest: blocks of the chain are placed first whenever an SM has room, so a large side
    // kernel launched earlier fills the gaps instead of standing in front of the chain
    // (measured on one box, bench step: 1.88 ms without priorities, 1.84 ms with).
    int prio_least = 0, prio_greatest = 0;
    cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
    if (cudaStreamCreateWithPriority(&c->own_stream, cudaStreamNonBlocking, prio_greatest) != cudaSuccess) {
        dvdagpu_set_error("cudaStreamCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete c;
        return nullptr;
    }
    c->stream = c->own_stream;
    c->pcm_slot = 0;
    c->hmap = c->dmap = nullptr; c->map_used = 0;
    if (cudaHostAlloc((void **)&c->hmap, MAP_BYTES, cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void **)&c->dmap, c->hmap, 0) != cudaSuccess) {
        dvdagpu_set_error("cannot allocate mapped staging memory: %s", cudaGetErrorString(cudaGetLastError()));
        delete c;
        return nullptr;
    }
    cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking);
    cudaStreamCreateWithPriority(&c->aux_stream, cudaStreamNonBlocking, prio_least);
    cudaStreamCreateWithPriority(&c->aux_stream_hi, cudaStreamNonBlocking, prio_greatest);
    for (auto &e : c->aux_ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->job_done, cudaEventDisableTiming);
    for (auto &e : c->pev) { cudaEventCreateWithFlags(&e[0], cudaEventDisableTiming); cudaEventCreateWithFlags(&e[1], cudaEventDisableTiming); }
    for (auto &e : c->ev) cudaEventCreate(&e);
    for (auto &k : c->kev) { cudaEventCreate(&k[0]); cudaEventCreate(&k[1]); }
    uint8_t pcm_tab[2][6][36], crc[256];
    build_pcm_tables(pcm_tab);
    build_crc8(crc);
    if (upload_pcm_tables(&pcm_tab[0][0][0]) || upload_crc_table(crc)) { dvdagpu_destroy(c); return nullptr; }
    c->huff_lut = huff_lut_device();
    if (!c->huff_lut) { dvdagpu_set_error("no Huffman table on the device"); dvdagpu_destroy(c); return nullptr; }
    return c;
}

extern "C" void dvdagpu_destroy(dvdagpu_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->peer) { dvdagpu_destroy(c->peer); c->peer = nullptr; }
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    if (c->job_done) cudaEventDestroy(c->job_done);
    if (c->hback) cudaFreeHost(c->hback);
    for (auto &b : c->buf) b.release();
    for (auto &e : c->ev) cudaEventDestroy(e);
    for (auto &k : c->kev) { cudaEventDestroy(k[0]); cudaEventDestroy(k[1]); }
    for (auto &e : c->pev) { cudaEventDestroy(e[0]); cudaEventDestroy(e[1]); }
    if (c->hmap) cudaFreeHost(c->hmap);
    cudaStreamDestroy(c->h2d_stream);
    cudaStreamDestroy(c->d2h_stream);
    cudaStreamDestroy(c->aux_stream);
    if (c->aux_stream_hi) cudaStreamDestroy(c->aux_stream_hi);
    for (auto &e : c->aux_ev) cudaEventDestroy(e);
    cudaStreamDestroy(c->own_stream);
    delete c;
}

extern "C" int dvdagpu_set_stream(dvdagpu_ctx *c, void *cuda_stream)
{
    if (!c) return -1;
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return 0;
}

extern "C" int dvdagpu_set_profiling(dvdagpu_ctx *c, int on)
{
    if (!c) return -1;
    c->profiling = on != 0;
    return 0;
}

// pinned host memory handed out through the ABI: how much is live, and the most that ever was
// (the host library's readers promise a bound on it; the tests hold them to it)
#include <atomic>
#include <map>
#include <mutex>
static std::mutex g_host_lock;
static std::map<void *, size_t> g_host_sizes;
static uint64_t g_host_live = 0, g_host_peak = 0;

extern "C" void *dvdagpu_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        dvdagpu_set_error("cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    std::lock_guard<std::mutex> hold(g_host_lock);
    g_host_sizes[p] = bytes;
    g_host_live += bytes;
    if (g_host_live > g_host_peak) g_host_peak = g_host_live;
    return p;
}
extern "C" void dvdagpu_host_free(void *p)
{
    if (!p) return;
    {
        std::lock_guard<std::mutex> hold(g_host_lock);
        auto it = g_host_sizes.find(p);
        if (it != g_host_sizes.end()) { g_host_live -= it->second; g_host_sizes.erase(it); }
    }
    cudaFreeHost(p);
}
extern "C" void dvdagpu_host_usage(uint64_t *live_bytes, uint64_t *peak_bytes, int reset_peak)
{
    std::lock_guard<std::mutex> hold(g_host_lock);
    if (live_bytes) *live_bytes = g_host_live;
    if (peak_bytes) *peak_bytes = g_host_peak;
    if (reset_peak) g_host_peak = g_host_live;
}

extern "C" int dvdagpu_get_stats(dvdagpu_ctx *c, dvdagpu_stats *out)
{
    if (!c || !out) return -1;
    *out = c->stats;
    return 0;
}

extern "C" const void *dvdagpu_pcm_device(dvdagpu_ctx *c, uint64_t *n_samples)
{
    if (!c) return nullptr;
    if (n_samples) *n_samples = c->pcm_samples;
    return c->buf[c->pcm_slot ? B_PCM2 : B_PCM].p;
}

extern "C" int dvdagpu_fetch(dvdagpu_ctx *c, uint64_t offset, uint64_t count, int32_t *dst)
{
    if (!c) return -1;
    if (offset + count > c->pcm_samples) { dvdagpu_set_error("fetch beyond the decoded samples"); return -1; }
    if (!count) return 0;
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(dst, c->buf[c->pcm_slot ? B_PCM2 : B_PCM].as<int32_t>() + offset, count * sizeof(int32_t),
                             cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---- one decode ----------------------------------------------------------------

#define ENSURE(id, bytes) do { if (c->buf[id].ensure(bytes)) { dvdagpu_set_error("%s (buffer %s, %zu bytes)", g_error, #id, (size_t)(bytes)); return -1; } } while (0)
#define TRY(expr) do { if ((expr) != 0) return -1; } while (0)
// device time of one kernel (or a short run of kernels) into stats.kernel_ms[id]
#define TIMED(id, expr)                                                        \
    do {                                                                       \
        CUDA_TRY(record_timing(c->kev[id][0], s));                             \
        TRY(expr);                                                             \
        CUDA_TRY(record_timing(c->kev[id][1], s));                             \
        c->kev_used[id] = true;                                                \
    } while (0)

// Small transfers do not go through the copy engines: a bulk upload or download
// of another part may be queued there for milliseconds (the pipelined path), and
// a 4-byte read-back would wait behind it.  A one-block kernel moves the words
// to / from a mapped pinned staging area instead (plain loads/stores over PCIe).
__global__ void k_copy_words(uint32_t *__restrict__ dst, const uint32_t *__restrict__ src, uint32_t nwords)
{
    for (uint32_t i = threadIdx.x; i < nwords; i += blockDim.x) dst[i] = src[i];
}

// host -> device, asynchronous (the staging bytes stay untouched until the next decode).
// Uploads can be queued and sent by one launch (flush_h2d).
#define H2D_BATCH 10
struct CopyBatch { uint32_t *dst[H2D_BATCH]; const uint32_t *src[H2D_BATCH]; uint32_t nwords[H2D_BATCH]; uint32_t n; };
__global__ void k_copy_multi(CopyBatch b)
{
    for (uint32_t j = 0; j < b.n; j++)
        for (uint32_t i = threadIdx.x; i < b.nwords[j]; i += blockDim.x) b.dst[j][i] = b.src[j][i];
}
static thread_local CopyBatch g_h2d_batch;

static int flush_h2d(dvdagpu_ctx *c)
{
    CopyBatch &b = g_h2d_batch;
    if (!b.n) return 0;
    if (b.n == 1) LAUNCH(k_copy_words, 1, 256, 0, c->stream, b.dst[0], b.src[0], b.nwords[0]);
    else LAUNCH(k_copy_multi, 1, 256, 0, c->stream, b);
    b.n = 0;
    CUDA_TRY(cudaGetLastError());
    return 0;
}
static int queue_h2d(dvdagpu_ctx *c, void *dev, const void *host, size_t bytes)
{
    if (!bytes) return 0;
    if ((bytes & 3) || c->map_used + bytes > MAP_BYTES / 2) {
        TRY(flush_h2d(c));
        CUDA_TRY(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        return 0;
    }
    CopyBatch &b = g_h2d_batch;
    if (b.n == H2D_BATCH) TRY(flush_h2d(c));
    memcpy(c->hmap + c->map_used, host, bytes);
    b.dst[b.n] = (uint32_t *)dev; b.src[b.n] = (const uint32_t *)(c->dmap + c->map_used); b.nwords[b.n] = (uint32_t)(bytes / 4);
    b.n++;
    c->map_used += (bytes + 15) & ~(size_t)15;
    return 0;
}
static int small_h2d(dvdagpu_ctx *c, void *dev, const void *host, size_t bytes)
{
    TRY(queue_h2d(c, dev, host, bytes));
    return flush_h2d(c);
}

// Where every track's samples start in the output buffer (16-byte aligned tracks: vector and
// bulk stores), computed on the device so that the output pass can be queued without a round
// trip through the host.  `capacity`: samples the buffer was sized for in advance.
#define OB_THREADS 256
__global__ void __launch_bounds__(OB_THREADS) k_track_out_base(TrackDev *tracks, uint32_t n_tracks, uint32_t *status, uint64_t capacity)
{
    if (*status & SEG_OVERFLOW) return;
    uint64_t carry = 0;
    for (uint32_t base = 0; base < n_tracks; base += OB_THREADS) {
        const uint32_t i = base + threadIdx.x;
        uint64_t v = 0;
        if (i < n_tracks && tracks[i].status == 0) v = tracks[i].frames * tracks[i].channels;
        uint64_t total;
        const uint64_t ex = block_excl_scan<OB_THREADS>((v + 3) & ~3ull, &total);
        if (i < n_tracks) tracks[i].out_base = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0 && carry > capacity) atomicOr(status, STATUS_PCM_SMALL);
}

// ---- what the launches of a decode are sized for ------------------------------------------
//
// Nothing a decode finds out about its input goes back to the host before it is over, so every
// table and grid is sized beforehand: from bounds the input's size gives (first decode of a
// context, or an input unlike the previous one), or from what the previous decode needed, scaled
// to the new input's size plus a margin (parts of one track, tracks of one title set, the same
// input again: the common cases).  The kernels take the real counts from device memory; a table
// that turns out too small stops the stages behind it, raises its CAP_* bit, and the decode is
// repeated with that table at the size now known.
static void shape_from_input(DecShape &sh, uint32_t n_sectors, uint32_t n_tracks)
{
    const uint64_t es_cap = (uint64_t)n_sectors * DVDA_SECTOR;
    sh.rows = (uint64_t)n_sectors + n_sectors / 4 + 64;         // one audio packet per sector is the rule
    sh.sync = es_cap / 2048 + 1024;                             // a sync every 2 KiB of stream
    sh.seg = sh.sync + n_tracks;
    sh.grp = sh.seg / DVDA_LANES + n_tracks + 1;
    sh.au = es_cap / 256 + 1024;
    sh.cells = es_cap / DVDA_LANES + 4096;                      // one sample per stream byte
    sh.max_au = 64;
    sh.nss = 2;
    sh.out_warps = 8 * sh.grp;
    sh.any_pcm = true; sh.any_mlp = true; sh.windowed = true;
    sh.pcm_fixed = (uint64_t)n_sectors * (DVDA_SECTOR / 2) + 4ull * n_tracks + 64;   // 16-bit samples fill every sector
}
static uint64_t with_margin(uint64_t v, double scale) { return (uint64_t)((double)v * scale * 1.0625) + 64; }
static void shape_from_last(DecShape &sh, const DecCounts &k, const DecShape &used, uint32_t last_sectors, uint32_t n_sectors, uint32_t n_tracks)
{
    const double scale = last_sectors ? (double)n_sectors / (double)last_sectors : 1.0;
    sh.rows = with_margin(k.np, scale);
    sh.sync = with_margin(k.n_raw > k.n_valid ? k.n_raw : k.n_valid, scale);
    sh.seg = with_margin(k.nseg, scale);
    sh.grp = with_margin(k.ngroups, scale) + n_tracks;
    sh.au = with_margin(k.nau, scale);
    sh.cells = with_margin(k.cells, scale);
    sh.max_au = k.max_au ? k.max_au : 1;
    sh.nss = k.nss_max ? k.nss_max : 1;
    sh.out_warps = (uint32_t)with_margin(k.nout_warps, scale);
    sh.any_pcm = k.any_pcm != 0; sh.any_mlp = k.any_mlp != 0;
    sh.windowed = !k.nau || k.es_total / k.nau <= CHK_WINDOWED_MAX_AU_BYTES;
    sh.pcm_fixed = k.any_pcm ? with_margin(k.pcm_fixed, scale) + 4ull * n_tracks : 4ull * n_tracks + 64;
    (void)used;
}

static int decode_enqueue(dvdagpu_ctx *c);

// One decode = decode_begin (sizes the tables, enqueues everything) + decode_end (waits for it,
// reads counts, status and the track table, repeats the decode if a table was too small, fills the
// results).  Between the two the host is free: the pipelined path starts the next part's decode on
// the context's peer while this one runs.
static int decode_begin(dvdagpu_ctx *c, const uint8_t *d_sectors, uint64_t n_sectors64,
                        uint32_t n_tracks, const dvdagpu_track_desc *descs)
{
    g_error[0] = 0;
    g_trace_on = getenv("DVDAGPU_TRACE") != nullptr;
    g_trace.n = 0;
    if (g_trace_on) trace_mark("decode begins", c->stream);
    if (n_sectors64 == 0 || n_sectors64 > 0x7FFFFFFFull) { dvdagpu_set_error("bad sector count"); return -1; }
    DecJob &J = c->job;
    J.n_tracks = n_tracks; J.active = false;
    if (!n_tracks) { c->pcm_samples = 0; return 0; }
    const uint32_t n_sectors = (uint32_t)n_sectors64;
    memset(&c->stats, 0, sizeof c->stats);
    memset(c->kev_used, 0, sizeof c->kev_used);
    g_profiling = c->profiling || getenv("DVDAGPU_PROFILE") != nullptr;

    // tracks in sector order (the kernels binary-search them); results go back in caller order
    std::vector<uint32_t> &order = J.order;
    order.resize(n_tracks);
    for (uint32_t i = 0; i < n_tracks; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return descs[a].first_sector < descs[b].first_sector; });
    std::vector<TrackDev> &ht = c->h_tracks;
    ht.assign(n_tracks, TrackDev());

    // decode mode: the three-pass fast path with the complete decoder as its fall-back (default),
    // or the complete single-pass decoder alone (DVDAGPU_SINGLE_PASS=1); the GPU tests run both
    const bool use_fast = getenv("DVDAGPU_SINGLE_PASS") == nullptr;
    // (test hook: DVDAGPU_SMALL_TABLES=1 starts every table sized in advance with room for one entry,
    // so that each decode goes through the grow-and-repeat paths)
    const bool small_tables = getenv("DVDAGPU_SMALL_TABLES") != nullptr;
    // (test hook: DVDAGPU_SYNC_SLOTS=0 sends every chunk with a match through the re-search path)
    const uint32_t nslots = getenv("DVDAGPU_SYNC_SLOTS") ? (uint32_t)atoi(getenv("DVDAGPU_SYNC_SLOTS")) : 2u;

    DecShape &sh = J.sh;
    sh = DecShape();
    const bool similar = c->have_last && !small_tables && c->last_tracks == n_tracks &&
                         n_sectors >= c->last_sectors / 2 && n_sectors / 2 <= c->last_sectors;
    if (similar) shape_from_last(sh, c->last, c->last_shape, c->last_sectors, n_sectors, n_tracks);
    else shape_from_input(sh, n_sectors, n_tracks);
    if (small_tables) { sh.rows = sh.sync = sh.seg = sh.grp = sh.au = sh.cells = 1; sh.max_au = 1; sh.out_warps = 8; sh.pcm_fixed = 1; }
    J.d_sectors = d_sectors; J.n_sectors = n_sectors;
    J.descs.assign(descs, descs + n_tracks);
    J.use_fast = use_fast; J.small_tables = small_tables; J.nslots = nslots;
    J.attempt = 0; J.pcm_small_seen = false; J.active = true; J.launches = 0;
    return decode_enqueue(c);
}

// everything of one attempt, asynchronous on the context's stream(s), up to the read-back
static int decode_enqueue(dvdagpu_ctx *c)
{
    DecJob &J = c->job;
    cudaStream_t s = c->stream;
    const uint8_t *d_sectors = J.d_sectors;
    const uint32_t n_sectors = J.n_sectors, n_tracks = J.n_tracks, nslots = J.nslots;
    const dvdagpu_track_desc *descs = J.descs.data();
    const std::vector<uint32_t> &order = J.order;
    std::vector<TrackDev> &ht = c->h_tracks;
    const bool use_fast = J.use_fast, small_tables = J.small_tables, pcm_small_seen = J.pcm_small_seen;
    const int attempt = J.attempt;
    DecShape &sh = J.sh;
    const uint64_t es_cap = (uint64_t)n_sectors * DVDA_SECTOR;
    const uint32_t chunks_cap = div_up_u32(es_cap, SYNC_CHUNK);
    MlpTables m;
    g_profiling = c->profiling || getenv("DVDAGPU_PROFILE") != nullptr;
    const uint32_t launches_before = g_launch_count;
    struct CountLaunches { DecJob &J; uint32_t before; ~CountLaunches() { J.launches += g_launch_count - before; } } count_launches = {J, launches_before};
    {
        if (attempt == 16) { dvdagpu_set_error("internal: the tables keep overflowing"); return -1; }
        c->map_used = 0;
        g_h2d_batch.n = 0;
        for (uint32_t i = 0; i < n_tracks; i++) {
            memset(&ht[i], 0, sizeof(TrackDev));
            ht[i].first_sector = descs[order[i]].first_sector;
            ht[i].last_sector = descs[order[i]].last_sector;
            ht[i].pts_length = descs[order[i]].pts_length;
            ht[i].cont = descs[order[i]].flags & 7u;
        }
        const uint32_t rows = (uint32_t)sh.rows, cap_sync = (uint32_t)sh.sync, cap_seg = (uint32_t)sh.seg, cap_grp = (uint32_t)sh.grp;
        const uint32_t cap_au = (uint32_t)sh.au, cap_work = 2 * n_tracks, cap_pairs = 2 * cap_grp;
        const uint64_t cells = sh.cells;
        const uint32_t lim_nss = sh.nss;

        // ---------------- buffers (they only ever grow)
        ENSURE(B_COUNTS, sizeof(DecCounts) + 64);
        ENSURE(B_STATUS, 64);
        ENSURE(B_SEC_CNT, (size_t)n_sectors * 4); ENSURE(B_SEC_BAD, (size_t)n_sectors * 4);
        ENSURE(B_SEC_BASE, (size_t)(n_sectors + 1) * 4); ENSURE(B_BAD_PREFIX, (size_t)(n_sectors + 1) * 4);
        ENSURE(B_SCAN_TMP, scan_tmp_bytes(std::max<uint64_t>((uint64_t)n_sectors * 2048 / 32 + 4096, std::max<uint64_t>(rows, chunks_cap)) + 4096));
        const size_t npa = (size_t)rows + 1;
        ENSURE(B_PK_SECTOR, npa * 4); ENSURE(B_PK_OFF, npa * 2); ENSURE(B_PK_LEN, npa * 2);
        ENSURE(B_PK_CODEC, npa); ENSURE(B_PK_PAD2, npa); ENSURE(B_PK_PARAMS, npa * 4);
        ENSURE(B_PK_MLPLEN, npa * 4); ENSURE(B_PK_PCMF, npa * 4);
        ENSURE(B_PK_ES, npa * 8); ENSURE(B_PK_PF, npa * 8);
        ENSURE(B_PK_NONMLP, npa * 4); ENSURE(B_PK_NM_PREFIX, npa * 4);
        ENSURE(B_PK_STOP, npa * 4); ENSURE(B_PK_STOP_PREFIX, npa * 4);
        ENSURE(B_PK_YIELD, npa);
        ENSURE(B_ES, es_cap + 16 + DVDA_ES_PAD);
        ENSURE(B_SYNC_CNT_RAW, (size_t)chunks_cap * 8);           // raw counts, then valid counts
        ENSURE(B_SYNC_BASE_RAW, (size_t)(chunks_cap + 1) * 4); ENSURE(B_SYNC_BASE_VALID, (size_t)(chunks_cap + 1) * 4);
        ENSURE(B_SYNC_SLOTS, (size_t)chunks_cap * SYNC_SLOT_BYTES);
        ENSURE(B_RAW, ((size_t)cap_sync + 1) * 8); ENSURE(B_VALID, ((size_t)cap_sync + 1) * 8);
        ENSURE(B_TRACKS, (size_t)n_tracks * sizeof(TrackDev));
        ENSURE(B_TRK_PK_LO, (size_t)n_tracks * 4); ENSURE(B_TRK_SEG_BASE, (size_t)(n_tracks + 1) * 4);
        ENSURE(B_TRK_GRP_BASE, (size_t)(n_tracks + 1) * 4);
        ENSURE(B_DEC_WORK, ((size_t)cap_work + 1) * sizeof(DecWork)); ENSURE(B_OUT_WORK, ((size_t)n_tracks + 1) * sizeof(OutWork));
        ENSURE(B_SEGS, ((size_t)cap_seg + 1) * sizeof(SegDev));
        ENSURE(B_SEG_NAU, ((size_t)cap_seg + 1) * 4); ENSURE(B_SEG_AU_BASE, ((size_t)cap_seg + 1) * 4);
        ENSURE(B_AU_NOTED, au_noted_bytes(cap_seg));
        ENSURE(B_SEG_NEED, ((size_t)cap_seg + 1) * 4);
        ENSURE(B_GROUPS, ((size_t)cap_grp + 1) * sizeof(GroupDev));
        ENSURE(B_GRP_CELLS, ((size_t)cap_grp + 1) * 4); ENSURE(B_CELL_BASE, ((size_t)cap_grp + 1) * 8);
        const size_t naua = (size_t)cap_au + 1;
        ENSURE(B_AU_POS, naua * 8); ENSURE(B_AU_ERR, naua); ENSURE(B_AU, naua * sizeof(AuDev)); ENSURE(B_AU_SEG, naua * 4);
        ENSURE(B_PSETS, naua * sizeof(ParamSet)); ENSURE(B_AU_FRAMES, naua * 2 * 4);
        ENSURE(B_SS_FLAGS, ((size_t)cap_seg + 1) * 2 * 4); ENSURE(B_SS_FLAGS_PREV, ((size_t)cap_seg + 1) * 2 * 4);
        ENSURE(B_SS_FLAGS_FAST, ((size_t)cap_seg + 1) * 2 * 4);
        ENSURE(B_FIR_TAIL, ((size_t)cap_seg + 1) * 2 * DVDA_MAX_CH * 8 * 4);
        ENSURE(B_SEG_FRAMES, ((size_t)cap_seg + 1) * 4); ENSURE(B_SEG_FRAME_SCAN, ((size_t)cap_seg + 1) * 8);
        if (use_fast) {
            // per-substream tables: [2][cap_au] (the kernels index them as k * cap_au + A)
            ENSURE(B_AU_SNAP, naua * lim_nss * au_snap_bytes());
            ENSURE(B_AU_FCHG, naua * lim_nss); ENSURE(B_SEG_CTX, ((size_t)cap_seg + 1) * 2 * seg_ctx_bytes());
            ENSURE(B_AU_DELTA, naua * lim_nss * au_delta_bytes() + 256);      // (+ alignment of the second table)
        }
        ENSURE(B_TILES, (cells * DVDA_LANES + 64 + 16 * DVDA_MAX_CH * DVDA_LANES) * sizeof(int32_t));   // 16 frames of slack: the filter passes read ahead
        ENSURE(B_BYPASS, cells * DVDA_LANES + 64);
        const int pcm_buf = c->pcm_slot ? B_PCM2 : B_PCM;
        // (test hook: a capacity of one sample sends the decode through the "too small" path, once)
        const uint64_t pcm_capacity = small_tables && !pcm_small_seen ? 1 : cells * DVDA_LANES + sh.pcm_fixed;
        ENSURE(pcm_buf, (pcm_capacity + 64) * sizeof(int32_t));

        // ---------------- from here to the read-back everything is asynchronous on the stream(s): the
        // sequence can be captured into a CUDA graph and launched as one (the first decode of a context
        // runs it directly: kernel attributes are set on first use).  The pipelined path does so: a
        // captured decode costs the GPU's front end one submission instead of forty, and while bulk
        // copies saturate the link every separate launch waits its turn on it (16 parts end to end:
        // 14.1 ms launched one by one, 10.5 ms as graphs).  With the input resident the capture and
        // the update of the executable graph cost the host more than they save (1.71 against 1.66 ms).
        const bool graph = c->use_graph && c->graph_ok && c->warmed && !g_trace_on && getenv("DVDAGPU_NO_GRAPH") == nullptr;
        if (graph) {
            if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); c->graph_ok = false; }
            else g_capturing = true;
        }
        struct CaptureGuard {                                // (an error return inside the capture must end it)
            cudaStream_t s;
            ~CaptureGuard() { if (g_capturing) { cudaGraph_t g = nullptr; cudaStreamEndCapture(s, &g); if (g) cudaGraphDestroy(g); g_capturing = false; cudaGetLastError(); } }
        } capture_guard = {s};
        CUDA_TRY(record_timing(c->ev[0], s));
        void *tmp = c->buf[B_SCAN_TMP].p;
        const size_t tmp_bytes = c->buf[B_SCAN_TMP].cap;
        if (c->scan_tmp_gen != c->buf[B_SCAN_TMP].gen) {
            // the scans find their control words zero and leave them zero (scan.cu)
            CUDA_TRY(cudaMemsetAsync(tmp, 0, tmp_bytes, s));
            c->scan_tmp_gen = c->buf[B_SCAN_TMP].gen;
        }
        DecCounts *cnt = c->buf[B_COUNTS].as<DecCounts>();
        uint32_t *d_status = c->buf[B_STATUS].as<uint32_t>();
        CUDA_TRY(cudaMemsetAsync(cnt, 0, sizeof(DecCounts), s));
        CUDA_TRY(cudaMemsetAsync(d_status, 0, 64, s));
        // frame counts a tile overflow of an earlier attempt found: kept across the attempts of one decode
        uint32_t *seg_need = c->buf[B_SEG_NEED].as<uint32_t>();
        if (attempt == 0 || c->seg_need_cap != cap_seg) {
            CUDA_TRY(cudaMemsetAsync(seg_need, 0, ((size_t)cap_seg + 1) * 4, s));
            c->seg_need_cap = cap_seg;
        }

        // ---------------- demux
        uint32_t *sec_cnt = c->buf[B_SEC_CNT].as<uint32_t>(), *sec_bad = c->buf[B_SEC_BAD].as<uint32_t>();
        uint32_t *sec_base = c->buf[B_SEC_BASE].as<uint32_t>(), *bad_prefix = c->buf[B_BAD_PREFIX].as<uint32_t>();
        TRY(launch_sector_count(d_sectors, n_sectors, sec_cnt, sec_bad, s));
        {
            const uint32_t *in[2] = {sec_cnt, sec_bad}; void *out[2] = {sec_base, bad_prefix}; const bool wide[2] = {false, false};
            uint64_t *const copy[2] = {&cnt->np, nullptr};
            TRY(scan_batch(in, out, wide, 2, n_sectors, tmp, tmp_bytes, s, copy));
        }
        PacketTable pt;
        pt.sector = c->buf[B_PK_SECTOR].as<uint32_t>(); pt.off = c->buf[B_PK_OFF].as<uint16_t>();
        pt.len = c->buf[B_PK_LEN].as<uint16_t>(); pt.codec = c->buf[B_PK_CODEC].as<uint8_t>();
        pt.pad2 = c->buf[B_PK_PAD2].as<uint8_t>(); pt.params = c->buf[B_PK_PARAMS].as<uint32_t>();
        pt.mlp_len = c->buf[B_PK_MLPLEN].as<uint32_t>(); pt.pcm_frames = c->buf[B_PK_PCMF].as<uint32_t>();
        uint64_t *pk_es = c->buf[B_PK_ES].as<uint64_t>(), *pk_pf = c->buf[B_PK_PF].as<uint64_t>();
        uint32_t *nonmlp = c->buf[B_PK_NONMLP].as<uint32_t>(), *nm_prefix = c->buf[B_PK_NM_PREFIX].as<uint32_t>();
        uint32_t *pstop = c->buf[B_PK_STOP].as<uint32_t>(), *stop_prefix = c->buf[B_PK_STOP_PREFIX].as<uint32_t>();
        TRY(launch_packet_fill(d_sectors, n_sectors, sec_base, pt, rows, nonmlp, pstop, s));
        {
            const uint32_t *in[4] = {pt.mlp_len, pt.pcm_frames, nonmlp, pstop};
            void *out[4] = {pk_es, pk_pf, nm_prefix, stop_prefix};
            const bool wide[4] = {true, true, false, false};
            uint64_t *const copy[4] = {&cnt->es_total, nullptr, nullptr, nullptr};
            TRY(scan_batch(in, out, wide, 4, rows, tmp, tmp_bytes, s, copy));
        }
        uint8_t *es = c->buf[B_ES].as<uint8_t>();
        uint16_t *sync_slots = c->buf[B_SYNC_SLOTS].as<uint16_t>();
        // (the two count tables are one buffer: cleared together; the gather counts the sync patterns it meets)
        uint32_t *cnt_raw = c->buf[B_SYNC_CNT_RAW].as<uint32_t>(), *cnt_valid = cnt_raw + chunks_cap;
        if (sh.any_mlp) CUDA_TRY(cudaMemsetAsync(cnt_raw, 0, (size_t)chunks_cap * 8, s));
        // The track table's inputs.  Up to TRACKS_BY_ARG tracks travel as kernel arguments: no load
        // from host memory stands in the decode chain (while bulk copies run on the link such a load
        // waits behind them for a long time).
        TrackDev *d_tracks = c->buf[B_TRACKS].as<TrackDev>();
        TrackArgs ta0;
        const bool by_arg = n_tracks <= TRACKS_BY_ARG;
        if (by_arg)
            for (uint32_t i = 0; i < n_tracks; i++) { ta0.t[i][0] = ht[i].first_sector; ta0.t[i][1] = ht[i].last_sector; ta0.t[i][2] = ht[i].pts_length; ta0.t[i][3] = ht[i].cont; }
        TIMED(DVDAGPU_K_ES_GATHER, launch_es_gather(d_sectors, pt, rows, pk_es, es, cnt, cnt_raw, sync_slots, nslots, sh.any_mlp,
                                                    d_tracks, n_tracks, by_arg ? &ta0 : nullptr, s));
        CUDA_TRY(record_timing(c->ev[1], s));

        // ---------------- index
        uint32_t *base_raw = c->buf[B_SYNC_BASE_RAW].as<uint32_t>(), *base_valid = c->buf[B_SYNC_BASE_VALID].as<uint32_t>();
        uint64_t *raw = c->buf[B_RAW].as<uint64_t>(), *valid = c->buf[B_VALID].as<uint64_t>();
        if (sh.any_mlp) {
            TIMED(DVDAGPU_K_SYNC_SCAN, launch_sync_validate(es, cnt, chunks_cap, cnt_raw, cnt_valid, sync_slots, nslots, s));
            const uint32_t *in[2] = {cnt_raw, cnt_valid}; void *out[2] = {base_raw, base_valid}; const bool wide[2] = {false, false};
            uint64_t *const copy[2] = {&cnt->n_raw, &cnt->n_valid};
            TRY(scan_batch(in, out, wide, 2, chunks_cap, tmp, tmp_bytes, s, copy));
            TRY(launch_sync_fill(es, cnt, chunks_cap, cnt_raw, sync_slots, nslots, base_raw, base_valid, raw, cap_sync, valid, cap_sync, s));
        }
        if (!by_arg) TRY(small_h2d(c, d_tracks, ht.data(), n_tracks * sizeof(TrackDev)));
        TrackSetupArgs ta;
        ta.es = es; ta.cnt = cnt; ta.n_sectors = n_sectors; ta.sec_base = sec_base; ta.bad_prefix = bad_prefix;
        ta.pt = pt; ta.pk_es = pk_es; ta.pk_pf = pk_pf; ta.pk_nonmlp = nm_prefix; ta.pk_pcm_stop = stop_prefix;
        ta.raw = raw; ta.cap_raw = cap_sync; ta.valid = valid; ta.cap_valid = cap_sync;
        ta.mlp_searched = sh.any_mlp ? 1u : 0u;
        TRY(launch_track_setup(ta, d_tracks, n_tracks, s));
        uint32_t *trk_pk_lo = c->buf[B_TRK_PK_LO].as<uint32_t>(), *trk_seg_base = c->buf[B_TRK_SEG_BASE].as<uint32_t>();
        uint32_t *trk_grp_base = c->buf[B_TRK_GRP_BASE].as<uint32_t>();
        DecWork *d_work = c->buf[B_DEC_WORK].as<DecWork>();
        OutWork *d_out_work = c->buf[B_OUT_WORK].as<OutWork>();
        const PlanLimits lim = {cap_au, cells, sh.max_au, lim_nss, sh.out_warps, sh.any_pcm ? 1u : 0u, sh.any_mlp ? 1u : 0u};
        TRY(launch_track_plan(d_tracks, n_tracks, trk_pk_lo, trk_seg_base, trk_grp_base, d_work, d_out_work, cap_work, cap_seg, cap_grp,
                              cap_sync, cnt, sh.any_mlp ? nullptr : &lim, s));

        memset(&m, 0, sizeof m);
        m.es = es; m.pk_es = pk_es; m.cnt = cnt; m.cap_seg = cap_seg; m.cap_au = cap_au; m.cap_grp = cap_grp;
        m.tracks = d_tracks; m.n_tracks = n_tracks;
        m.status = d_status; m.status_rw = d_status; m.any_fallback = d_status + 8;
        m.huff_lut = c->huff_lut;
        m.seg_need = seg_need;
        m.segs = c->buf[B_SEGS].as<SegDev>();
        m.groups = c->buf[B_GROUPS].as<GroupDev>();
        m.au_pos = c->buf[B_AU_POS].as<uint64_t>(); m.au_err = c->buf[B_AU_ERR].as<uint8_t>(); m.au_seg = c->buf[B_AU_SEG].as<uint32_t>();
        m.au = c->buf[B_AU].as<AuDev>(); m.psets = c->buf[B_PSETS].as<ParamSet>();
        m.au_frames_ss = c->buf[B_AU_FRAMES].as<uint32_t>();
        m.ss_flags = c->buf[B_SS_FLAGS].as<uint32_t>(); m.ss_flags_prev = c->buf[B_SS_FLAGS_PREV].as<uint32_t>(); m.ss_flags_fast = c->buf[B_SS_FLAGS_FAST].as<uint32_t>();
        m.fir_tail = c->buf[B_FIR_TAIL].as<int32_t>();
        m.tiles = c->buf[B_TILES].as<int32_t>(); m.bypass = c->buf[B_BYPASS].as<uint8_t>();
        m.pcm = c->buf[pcm_buf].as<int32_t>();
        if (use_fast) {
            m.au_snap = reinterpret_cast<AuSnap *>(c->buf[B_AU_SNAP].p);
            m.au_fchg = c->buf[B_AU_FCHG].as<uint8_t>();
            m.seg_ctx = reinterpret_cast<SegCtx *>(c->buf[B_SEG_CTX].p);
            au_delta_split(c->buf[B_AU_DELTA].p, naua * lim_nss, m);
        }
        if (sh.any_mlp) {
            uint32_t *seg_nau = c->buf[B_SEG_NAU].as<uint32_t>(), *seg_au_base = c->buf[B_SEG_AU_BASE].as<uint32_t>();
            uint32_t *au_noted = c->buf[B_AU_NOTED].as<uint32_t>();
            uint32_t *grp_cells = c->buf[B_GRP_CELLS].as<uint32_t>();
            uint64_t *cell_base = c->buf[B_CELL_BASE].as<uint64_t>();
            TRY(launch_segment_fill(d_tracks, n_tracks, trk_seg_base, valid, m.segs, cap_seg, cnt, s));
            TRY(launch_au_chase(es, m.segs, cap_seg, cnt, d_tracks, seg_nau, nullptr, nullptr, seg_au_base, au_noted, 0, s));
            {
                const uint32_t *in[1] = {seg_nau}; void *out[1] = {seg_au_base}; const bool wide[1] = {false};
                uint64_t *const copy[1] = {&cnt->nau};
                TRY(scan_batch(in, out, wide, 1, cap_seg, tmp, tmp_bytes, s, copy));
            }
            // the groups (tile sizes) follow from the access-unit counts alone
            TRY(launch_group_setup(d_tracks, n_tracks, trk_grp_base, m.segs, m.groups, cap_grp, cnt, grp_cells, seg_need, s));
            {
                const uint32_t *in[1] = {grp_cells}; void *out[1] = {cell_base}; const bool wide[1] = {true};
                uint64_t *const copy[1] = {&cnt->cells};
                TRY(scan_batch(in, out, wide, 1, cap_grp, tmp, tmp_bytes, s, copy));
            }
            TRY(launch_plan_check(cnt, lim, s));
            TRY(launch_au_chase(es, m.segs, cap_seg, cnt, d_tracks, seg_nau, m.au_pos, m.au_seg, seg_au_base, au_noted, 1, s));
            // Parity / CRC-8 on a second stream, beside the rest of the index chain and the header passes
            // (it needs the access units' positions, nothing else).  Small
            // access units: the windowed kernel on the low-priority stream (it fills what the chain
            // leaves free).  Large ones: the direct kernel at the chain's own priority, which then runs
            // first and lets the header passes follow.  Both pairings, and starting the check beside the
            // entropy pass or on its own in between, were measured: DESIGN.md section 6.
            // (a caller's stream has the default = lowest priority, like aux_stream)
            cudaStream_t chk_stream = (sh.windowed || s != c->own_stream) ? c->aux_stream : c->aux_stream_hi;
            CUDA_TRY(cudaEventRecord(c->aux_ev[0], s));
            CUDA_TRY(cudaStreamWaitEvent(chk_stream, c->aux_ev[0], 0));
            CUDA_TRY(record_timing(c->kev[DVDAGPU_K_CHECKDATA][0], chk_stream));
            TRY(launch_checkdata(m, seg_au_base, sh.windowed, chk_stream));
            CUDA_TRY(record_timing(c->kev[DVDAGPU_K_CHECKDATA][1], chk_stream));
            c->kev_used[DVDAGPU_K_CHECKDATA] = true;
            CUDA_TRY(cudaEventRecord(c->aux_ev[1], chk_stream));
            bool any_parts = false;
            for (uint32_t i = 0; i < n_tracks; i++) any_parts |= (ht[i].cont & TRACK_CONT_NEXT) != 0;
            TRY(launch_yield(m, rows, seg_au_base, pt, trk_pk_lo, c->buf[B_PK_YIELD].as<uint8_t>(), any_parts, s));
            CUDA_TRY(record_timing(c->ev[2], s));

            // ---------------- decode
            uint32_t *seg_frames = c->buf[B_SEG_FRAMES].as<uint32_t>();
            uint64_t *seg_frame_scan = c->buf[B_SEG_FRAME_SCAN].as<uint64_t>();
            // (the group offsets' kernel also gives the fast path's segment tables their start values)
            TRY(launch_group_offsets(m.groups, cap_grp, cnt, cell_base,
                                     use_fast ? (void *)m.seg_ctx : nullptr, use_fast ? ((size_t)cap_seg + 1) * 2 * seg_ctx_bytes() : 0,
                                     use_fast ? m.ss_flags : nullptr, use_fast ? (cap_seg + 1) * 2 : 0u, s));
            m.fast = 0;
            if (use_fast) {
                // The three-pass path (access-unit parallel) decodes what has the common shape; what it
                // gives up on is flagged ...
                TRY(launch_mlp_fast(m, d_work, cap_pairs, sh.max_au, lim_nss, c->kev, c->kev_used, c->aux_ev[1], s));
                // (the flags as the fast path left them: what the complete decoder takes, what the output pass leaves alone)
                CUDA_TRY(cudaMemcpyAsync(m.ss_flags_fast, m.ss_flags, ((size_t)cap_seg + 1) * 2 * 4, cudaMemcpyDeviceToDevice, s));
                m.fast = 1;
            }
            CUDA_TRY(cudaStreamWaitEvent(s, c->aux_ev[1], 0));
            // ... and decoded by the complete single-pass decoder (everything, without the fast path)
            TIMED(DVDAGPU_K_MLP_DECODE, launch_mlp_decode(m, d_work, cap_pairs, s));
            CUDA_TRY(cudaMemcpyAsync(m.ss_flags_prev, m.ss_flags, ((size_t)cap_seg + 1) * 2 * 4, cudaMemcpyDeviceToDevice, s));
            TIMED(DVDAGPU_K_CARRY_FIX, launch_carry_fix(m, s));
            TRY(launch_seg_finalize(m, seg_frames, d_status, s));
            TRY(scan_u32_to_u64(seg_frames, seg_frame_scan, cap_seg, tmp, tmp_bytes, s));
            TRY(launch_track_finalize(m, seg_frame_scan, d_status, s));
        } else {
            CUDA_TRY(record_timing(c->ev[2], s));          // (the plan's block has checked the limits itself)
        }
        // The output buffer was sized before the frame counts are known (every MLP sample has a place
        // in the tiles, so the tiles' size bounds them); the tracks' places in it are computed on the device.
        LAUNCH(k_track_out_base, 1, OB_THREADS, 0, s, d_tracks, n_tracks, d_status, pcm_capacity);
        CUDA_TRY(record_timing(c->ev[3], s));

        // ---------------- output
        if (sh.any_mlp) {
            if (m.fast) TIMED(DVDAGPU_K_MLP_FILTER_OUT, launch_mlp_filter_out(m, d_out_work, sh.out_warps, s));
            TIMED(DVDAGPU_K_REMATRIX, launch_rematrix(m, s));
        }
        if (sh.any_pcm) TIMED(DVDAGPU_K_PCM_UNPACK, launch_pcm_unpack(d_sectors, pt, rows, cnt, d_status, pk_pf, d_tracks, trk_pk_lo, n_tracks, m.pcm, s));
        CUDA_TRY(record_timing(c->ev[4], s));

        if (g_capturing) {
            cudaGraph_t captured = nullptr;
            g_capturing = false;
            CUDA_TRY(cudaStreamEndCapture(s, &captured));
            bool ready = false;
            if (c->graph_exec) {
                cudaGraphExecUpdateResultInfo info;
                if (cudaGraphExecUpdate(c->graph_exec, captured, &info) == cudaSuccess) ready = true;
                else { cudaGetLastError(); cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
            }
            if (!ready) {
                const cudaError_t e = cudaGraphInstantiate(&c->graph_exec, captured, 0);
                if (e != cudaSuccess) { cudaGraphDestroy(captured); dvdagpu_set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(e)); return -1; }
            }
            cudaGraphDestroy(captured);
            CUDA_TRY(cudaGraphLaunch(c->graph_exec, s));
        }
        c->warmed = true;
        // ---------------- the read-back of the one round trip: counts, status, track table go to
        // mapped host memory behind everything else; decode_end() waits for the event
        {
            const size_t need = sizeof(DecCounts) + 64 + (size_t)n_tracks * sizeof(TrackDev);
            if (need > c->back_bytes) {
                if (c->hback) cudaFreeHost(c->hback);
                c->hback = c->dback = nullptr; c->back_bytes = 0;
                CUDA_TRY(cudaHostAlloc((void **)&c->hback, need + need / 2, cudaHostAllocMapped));
                CUDA_TRY(cudaHostGetDevicePointer((void **)&c->dback, c->hback, 0));
                c->back_bytes = need + need / 2;
            }
            CopyBatch bb;
            bb.n = 3;
            bb.dst[0] = (uint32_t *)c->dback; bb.src[0] = (const uint32_t *)cnt; bb.nwords[0] = sizeof(DecCounts) / 4;
            bb.dst[1] = (uint32_t *)(c->dback + sizeof(DecCounts)); bb.src[1] = d_status; bb.nwords[1] = 1;
            bb.dst[2] = (uint32_t *)(c->dback + sizeof(DecCounts) + 64); bb.src[2] = (const uint32_t *)d_tracks; bb.nwords[2] = (uint32_t)(n_tracks * sizeof(TrackDev) / 4);
            LAUNCH(k_copy_multi, 1, 256, 0, s, bb);
            CUDA_TRY(cudaGetLastError());
        }
        CUDA_TRY(cudaEventRecord(c->job_done, s));
    }
    return 0;
}

// waits for the decode begun by decode_begin, repeats it if a table was too small, fills the results
static int decode_end(dvdagpu_ctx *c, dvdagpu_track_result *results)
{
    DecJob &J = c->job;
    const uint32_t n_tracks = J.n_tracks;
    if (!n_tracks) return 0;
    if (!J.active) { dvdagpu_set_error("internal: no decode in flight"); return -1; }
    const uint32_t n_sectors = J.n_sectors;
    const std::vector<uint32_t> &order = J.order;
    std::vector<TrackDev> &ht = c->h_tracks;
    DecShape &sh = J.sh;
    DecCounts k;
    uint32_t status = 0;
    uint64_t total_samples = 0;
    for (;;) {
        const int attempt = J.attempt;
        CUDA_TRY(cudaEventSynchronize(c->job_done));
        memcpy(&k, c->hback, sizeof(DecCounts));
        memcpy(&status, c->hback + sizeof(DecCounts), 4);
        memcpy(ht.data(), c->hback + sizeof(DecCounts) + 64, (size_t)n_tracks * sizeof(TrackDev));
        if (g_trace_on) { trace_host("decode done"); trace_dump(); g_trace_on = false; }
        if (getenv("DVDAGPU_DEBUG")) {
            fprintf(stderr, "[dvdagpu] attempt %d: sectors=%u packets=%llu es=%llu raw=%llu valid=%llu segs=%u groups=%u aus=%llu cells=%llu max_au=%u overflow=%x status=%x\n",
                    attempt, n_sectors, (unsigned long long)k.np, (unsigned long long)k.es_total, (unsigned long long)k.n_raw,
                    (unsigned long long)k.n_valid, k.nseg, k.ngroups, (unsigned long long)k.nau, (unsigned long long)k.cells, k.max_au, k.overflow, status);
            for (uint32_t i = 0; i < n_tracks; i++) {
                const TrackDev &T = ht[i];
                fprintf(stderr, "[dvdagpu] track %u: status=%d codec=%d err=%x ch=%u pk=[%u,%u) pk_x=%u pk_open=%u es=[%llu,%llu) cut=%llu nss=%u nseg=%u err_seg=%u frames=%llu trunc=%u\n",
                        i, T.status, T.codec, T.error_flags, T.channels, T.pk_lo, T.pk_hi, T.pk_x, T.pk_open,
                        (unsigned long long)T.es_start, (unsigned long long)T.es_end, (unsigned long long)T.es_cut,
                        T.nss, T.nseg, T.err_seg, (unsigned long long)T.frames, T.truncated);
                fprintf(stderr, "[dvdagpu]          cont=%u check=[%u,%u) stopped=%u\n", T.cont, T.pk_check, T.pk_check_end, T.stopped);
            }
            if (k.nseg && k.nseg <= 4096 && c->buf[B_SEGS].p && c->buf[B_SS_FLAGS].p && c->buf[B_SS_FLAGS_FAST].p) {
                // the segment table with each substream's flags: after the fast path / at the end
                const uint32_t cap_seg = (uint32_t)sh.seg;
                std::vector<SegDev> hs(k.nseg);
                std::vector<uint32_t> f_fast(2 * ((size_t)cap_seg + 1)), f_end(2 * ((size_t)cap_seg + 1));
                cudaMemcpy(hs.data(), c->buf[B_SEGS].p, k.nseg * sizeof(SegDev), cudaMemcpyDeviceToHost);
                cudaMemcpy(f_fast.data(), c->buf[B_SS_FLAGS_FAST].p, f_fast.size() * 4, cudaMemcpyDeviceToHost);
                cudaMemcpy(f_end.data(), c->buf[B_SS_FLAGS].p, f_end.size() * 4, cudaMemcpyDeviceToHost);
                std::vector<AuDev> ha(k.nau <= 8192 ? k.nau : 0);
                std::vector<ParamSet> hp(ha.size());
                if (!ha.empty()) {
                    cudaMemcpy(ha.data(), c->buf[B_AU].p, ha.size() * sizeof(AuDev), cudaMemcpyDeviceToHost);
                    cudaMemcpy(hp.data(), c->buf[B_PSETS].p, hp.size() * sizeof(ParamSet), cudaMemcpyDeviceToHost);
                }
                for (uint32_t i = 0; i < k.nseg; i++) {
                    fprintf(stderr, "[dvdagpu] seg %u: track %u frame0=%llu frames=%u n_au=%u au_base=%u flags=%x err=%x err_au=%u | fast ss0=%x ss1=%x | end ss0=%x ss1=%x\n",
                            i, hs[i].track, (unsigned long long)hs[i].frame0, hs[i].frames, hs[i].n_au, hs[i].au_base, hs[i].flags, hs[i].err, hs[i].err_au,
                            f_fast[i], f_fast[cap_seg + i], f_end[i], f_end[cap_seg + i]);
                    for (uint32_t a = 0; a < hs[i].n_au && hs[i].au_base + a < ha.size(); a++) {
                        const AuDev &u = ha[hs[i].au_base + a];
                        const uint32_t pi = u.pset & 0x7FFFFFFFu;
                        fprintf(stderr, "[dvdagpu]     au %u: frame0=%u n=%u seed=%06x pset=%u%s", hs[i].au_base + a, u.frame0, u.nframes, u.seed & 0x7FFFFF, pi, (u.pset >> 31) ? " (trivial)" : "");
                        if (pi < hp.size()) {
                            const ParamSet &P = hp[pi];
                            fprintf(stderr, " | ml=%u mmc=%u nsh=%u noise=%u", P.matrix_len, P.mmc, P.noise_shift, P.uses_noise);
                            for (uint32_t mk = 0; mk < P.matrix_len && mk < DVDA_MAX_MAT; mk++) {
                                fprintf(stderr, " m%u->%u:", mk, P.out_ch[mk]);
                                for (int cc = 0; cc < DVDA_MAX_CH; cc++) fprintf(stderr, "%d,", P.coeff[mk][cc]);
                            }
                            fprintf(stderr, " q=%u,%u sh=%u,%u", P.q[0], P.q[1], P.out_shift[0], P.out_shift[1]);
                        }
                        fprintf(stderr, "\n");
                    }
                }
            }
        }
        // ---------------- a table too small, a kernel left out that had work, a tile or the output buffer too small?
        if (!k.overflow && !(status & (SEG_OVERFLOW | STATUS_PCM_SMALL))) break;
        DecShape in;
        shape_from_input(in, n_sectors, n_tracks);
        auto grow = [](uint64_t &cap, uint64_t need, uint64_t bound) { cap = std::max<uint64_t>(std::max<uint64_t>(need + need / 16 + 64, std::min<uint64_t>(cap * 4, bound)), cap); };
        if (k.overflow & CAP_ROWS) grow(sh.rows, k.need_rows, in.rows);
        if (k.overflow & CAP_SYNC) grow(sh.sync, k.need_sync, in.sync);
        if (k.overflow & CAP_SEG) grow(sh.seg, k.need_seg, in.seg);
        if (k.overflow & CAP_GRP) grow(sh.grp, k.need_grp, in.grp);
        if (k.overflow & CAP_AU) grow(sh.au, k.need_au, in.au);
        if (k.overflow & CAP_CELLS) grow(sh.cells, k.need_cells, in.cells);
        if (k.overflow & CAP_MAX_AU) sh.max_au = std::max(k.max_au, sh.max_au * 2);
        if (k.overflow & CAP_SHAPE) {
            // an input unlike the previous one: everything at least as the input's size bounds it
            sh.rows = std::max(sh.rows, in.rows); sh.sync = std::max(sh.sync, in.sync); sh.seg = std::max(sh.seg, in.seg);
            sh.grp = std::max(sh.grp, in.grp); sh.au = std::max(sh.au, in.au); sh.cells = std::max(sh.cells, in.cells);
            sh.max_au = std::max(sh.max_au, in.max_au); sh.out_warps = std::max(sh.out_warps, (uint32_t)(8 * sh.grp));
            sh.pcm_fixed = std::max(sh.pcm_fixed, in.pcm_fixed);
            sh.any_pcm = sh.any_mlp = true; sh.nss = 2;
        }
        if (status & STATUS_PCM_SMALL) {
            J.pcm_small_seen = true;
            uint64_t want = 64;
            for (uint32_t i = 0; i < n_tracks; i++) if (ht[i].status == 0) want += ht[i].frames * ht[i].channels + 4;
            sh.pcm_fixed = std::max<uint64_t>(sh.pcm_fixed, want);   // (on top of what the tiles bound)
        }
        J.attempt++;
        TRY(decode_enqueue(c));
    }
    // ---------------- results
    total_samples = 0;
    for (uint32_t i = 0; i < n_tracks; i++) {
        total_samples = (total_samples + 3) & ~3ull;          // 16-byte aligned tracks (vector stores)
        if (ht[i].out_base != total_samples) { dvdagpu_set_error("internal: output layout"); return -1; }   // k_track_out_base
        if (ht[i].status == 0) total_samples += ht[i].frames * ht[i].channels;
    }
    c->pcm_samples = total_samples;
    J.active = false;
    c->last = k; c->last_shape = sh; c->last_sectors = n_sectors; c->last_tracks = n_tracks; c->have_last = true;
    uint64_t es_used = 0;
    for (uint32_t i = 0; i < n_tracks; i++) {
        const TrackDev &T = ht[i];
        dvdagpu_track_result &R = results[order[i]];
        memset(&R, 0, sizeof R);
        R.status = T.status;
        if (T.status != 0) continue;
        R.error_flags = T.error_flags; R.codec = T.codec;
        R.group_0_bps = T.g0_bps; R.group_1_bps = T.g1_bps; R.group_0_rate = T.g0_rate; R.group_1_rate = T.g1_rate;
        R.channel_assignment = T.assignment; R.channels = T.channels; R.bits_per_sample = T.bits; R.sample_rate = T.rate;
        R.frames = T.frames; R.pcm_offset = T.out_base; R.truncated = T.truncated; R.stopped = T.stopped;
        if (T.codec == 1) es_used += T.es_end - T.es_start;
    }
    g_profiling = c->profiling || getenv("DVDAGPU_PROFILE") != nullptr;
    if (g_profiling) {
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]); c->stats.demux_ms = ms;
        cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]); c->stats.index_ms = ms;
        cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]); c->stats.decode_ms = ms;
        cudaEventElapsedTime(&ms, c->ev[3], c->ev[4]); c->stats.output_ms = ms;
        cudaEventElapsedTime(&ms, c->ev[0], c->ev[4]); c->stats.total_ms = ms;
        for (int kk = 0; kk < 16; kk++) {
            if (c->kev_used[kk] && cudaEventElapsedTime(&ms, c->kev[kk][0], c->kev[kk][1]) == cudaSuccess) c->stats.kernel_ms[kk] = ms;
        }
        cudaGetLastError();
    }
    c->stats.launches = J.launches;
    c->stats.segments = k.nseg;
    c->stats.access_units = k.nau;
    c->stats.es_bytes = es_used;
    c->stats.samples = total_samples;
    return 0;
}


static int decode_on_device(dvdagpu_ctx *c, const uint8_t *d_sectors, uint64_t n_sectors64,
                            uint32_t n_tracks, const dvdagpu_track_desc *descs, dvdagpu_track_result *results)
{
    TRY(decode_begin(c, d_sectors, n_sectors64, n_tracks, descs));
    return decode_end(c, results);
}

extern "C" int dvdagpu_decode_device(dvdagpu_ctx *c, const void *device_sectors, uint64_t n_sectors,
                                     uint32_t n_tracks, const dvdagpu_track_desc *tracks, dvdagpu_track_result *results)
{
    if (!c || !device_sectors || !tracks || !results) { dvdagpu_set_error("null argument"); return -1; }
    CUDA_TRY(cudaSetDevice(c->device));
    return decode_on_device(c, (const uint8_t *)device_sectors, n_sectors, n_tracks, tracks, results);
}

extern "C" int dvdagpu_decode_host(dvdagpu_ctx *c, const uint8_t *sectors, uint64_t n_sectors,
                                   uint32_t n_tracks, const dvdagpu_track_desc *tracks, dvdagpu_track_result *results)
{
    if (!c || !sectors || !tracks || !results) { dvdagpu_set_error("null argument"); return -1; }
    CUDA_TRY(cudaSetDevice(c->device));
    ENSURE(B_SECTORS, n_sectors * DVDA_SECTOR + 256);
    CUDA_TRY(cudaMemcpyAsync(c->buf[B_SECTORS].p, sectors, n_sectors * DVDA_SECTOR, cudaMemcpyHostToDevice, c->stream));
    return decode_on_device(c, c->buf[B_SECTORS].as<uint8_t>(), n_sectors, n_tracks, tracks, results);
}


// ---- one long track, pipelined ----------------------------------------------------
//
// The track is decoded in parts of `part_sectors` sectors.  Each part is a
// "track" of its own (DVDAGPU_PART_* flags): the cut lands on the first major
// sync behind the part's last sector, exactly like the disc's own track
// boundaries, so the parts' outputs concatenate to the whole track.  Upload of
// part i+1 (copy engine, h2d_stream) and download of part i-1 (d2h_stream) run
// while part i is being decoded; sector and PCM buffers are double-buffered, all
// other buffers are reused because decodes are serial.

static void add_stats(dvdagpu_stats &a, const dvdagpu_stats &b)
{
    a.demux_ms += b.demux_ms; a.index_ms += b.index_ms; a.decode_ms += b.decode_ms; a.output_ms += b.output_ms;
    a.total_ms += b.total_ms; a.launches += b.launches; a.segments += b.segments; a.access_units += b.access_units;
    a.es_bytes += b.es_bytes; a.samples += b.samples;
    for (int k = 0; k < 16; k++) a.kernel_ms[k] += b.kernel_ms[k];
}

extern "C" int dvdagpu_decode_track_pipelined(dvdagpu_ctx *c, const uint8_t *sectors, uint64_t n_sectors,
                                              const dvdagpu_track_desc *track, uint32_t part_sectors,
                                              int32_t *pcm_host, uint64_t pcm_capacity, dvdagpu_track_result *result)
{
    if (!c || !sectors || !track || !result || !pcm_host) { dvdagpu_set_error("null argument"); return -1; }
    CUDA_TRY(cudaSetDevice(c->device));
    const uint64_t first = track->first_sector;
    const uint64_t last = track->last_sector < n_sectors ? track->last_sector : n_sectors - 1;
    if (!part_sectors && getenv("DVDAGPU_PART_SECTORS")) part_sectors = (uint32_t)atoi(getenv("DVDAGPU_PART_SECTORS"));   // (tuning hook)
    // Where the parts begin.  The download of the samples is the longest leg, so the parts are small
    // enough for it to start early (a sixteenth of the track, at most 32 MB of AOB: the device
    // buffers stay bounded however long the track is) and large enough for the latency floor of a
    // decode — a chain of small dependent launches, about half a millisecond — not to outlast the
    // download of the part before (two parts decode at once, see below).
    std::vector<uint64_t> starts;
    {
        const uint64_t n = last >= first && first < n_sectors ? last - first + 1 : 0;
        if (!part_sectors) part_sectors = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(n / 16, 8192), 16384);
        if (n >= 2ull * part_sectors) for (uint64_t s0 = first; s0 <= last; s0 += part_sectors) starts.push_back(s0);
        // (a last part of a few sectors would be all latency: the part before takes it)
        if (starts.size() > 2 && last + 1 - starts.back() < part_sectors / 4) starts.pop_back();
    }
    const uint64_t margin = 64;
    dvdagpu_stats total_stats;
    memset(&total_stats, 0, sizeof total_stats);

    bool fallback = starts.size() < 2;
    uint64_t total_samples = 0, total_frames = 0;
    dvdagpu_track_result merged;
    memset(&merged, 0, sizeof merged);
    if (!fallback) {
        // Two contexts on the device take the parts in turn: a decode is a chain of dependent
        // launches, most of them small, so two chains side by side cost little more than one, and
        // while the host waits for part i - 1 part i is already running.
        dvdagpu_ctx *X[2] = {c, c};
        if (!getenv("DVDAGPU_PIPE_ONE_CONTEXT")) {
            if (!c->peer) c->peer = dvdagpu_create(c->device);
            if (c->peer) { X[1] = c->peer; c->peer->profiling = c->profiling; }
        }
        const bool two = X[1] != c;
        struct GraphMode {                                   // decodes of the parts are launched as graphs
            dvdagpu_ctx *a, *b;
            GraphMode(dvdagpu_ctx *a_, dvdagpu_ctx *b_) : a(a_), b(b_) { a->use_graph = b->use_graph = true; }
            ~GraphMode() { a->use_graph = b->use_graph = false; }
        } graph_mode(X[0], X[1]);
        uint32_t parts = (uint32_t)starts.size();
        // A PCM track (known once part 0 is decoded) goes on in windows of whole sectors: every packet
        // stands alone, what carries over is the frame budget (dvd-audio.c:1016-1082) — so the parts
        // follow one another (each is told the frames still to come), run on behind the track's last
        // sector while the budget lasts, and only uploads and downloads overlap with the decodes.
        bool pcm = false;
        uint64_t pcm_budget = 0;
        uint32_t restart_at = 0;                             // parts begun before the codec was known are decoded again
        auto window = [&](uint32_t i, uint64_t &s0, uint64_t &len, uint64_t &e_rel) {
            s0 = starts[i];
            if (pcm && i) {
                len = std::min<uint64_t>(part_sectors, n_sectors - s0);
                e_rel = len - 1;
                return;
            }
            const uint64_t e = i + 1 == parts ? last : starts[i + 1] - 1;
            uint64_t stop = (i + 1 == parts) ? n_sectors : e + 1 + margin;
            if (stop > n_sectors) stop = n_sectors;
            len = stop - s0;
            e_rel = e - s0;
        };
        // part i: context i & 1 (with two contexts), sector / PCM slot of that context in turn
        auto ctx_of = [&](uint32_t i) { return two ? X[i & 1] : c; };
        auto slot_of = [&](uint32_t i) { return two ? (int)((i >> 1) & 1) : (int)(i & 1); };
        const uint32_t reuse = two ? 4 : 2;                 // part i reuses the buffers of part i - reuse
        // DVDAGPU_PIPE_TRACE=1: when every leg of every part began and ended (debug aid)
        const bool ptrace = getenv("DVDAGPU_PIPE_TRACE") != nullptr;
        std::vector<cudaEvent_t> tev;
        std::vector<char> tkind;
        auto tmark = [&](cudaStream_t st, char kind) { if (ptrace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); tev.push_back(e); tkind.push_back(kind); } };
        auto upload = [&](uint32_t i) -> int {
            uint64_t s0, len, e_rel;
            window(i, s0, len, e_rel);
            dvdagpu_ctx *x = ctx_of(i);
            const int slot = slot_of(i), buf = slot ? B_SECTORS2 : B_SECTORS;
            if (x->buf[buf].ensure(len * DVDA_SECTOR + 256)) return -1;
            if (i >= reuse) CUDA_TRY(cudaStreamWaitEvent(c->h2d_stream, x->pev[1][slot], 0));   // part i - reuse decoded
            tmark(c->h2d_stream, 'u');
            CUDA_TRY(cudaMemcpyAsync(x->buf[buf].p, sectors + s0 * DVDA_SECTOR, len * DVDA_SECTOR,
                                     cudaMemcpyHostToDevice, c->h2d_stream));
            CUDA_TRY(cudaEventRecord(x->pev[0][slot], c->h2d_stream));
            tmark(c->h2d_stream, 'U');
            return 0;
        };
        auto begin = [&](uint32_t i) -> int {
            uint64_t s0, len, e_rel;
            window(i, s0, len, e_rel);
            dvdagpu_ctx *x = ctx_of(i);
            const int slot = slot_of(i);
            CUDA_TRY(cudaStreamWaitEvent(x->stream, x->pev[0][slot], 0));
            if (i >= reuse) CUDA_TRY(cudaStreamWaitEvent(x->stream, x->pev[2][slot], 0));       // PCM slot downloaded
            x->pcm_slot = slot;
            dvdagpu_track_desc d = {0, (uint32_t)e_rel, track->pts_length,
                                    (i ? (uint32_t)DVDAGPU_PART_CONTINUES_PREVIOUS : (track->flags & 1u)) |
                                    (i + 1 < parts ? (uint32_t)DVDAGPU_PART_CONTINUED_BY_NEXT : (track->flags & 2u))};
            if (pcm && i) { d.pts_length = (uint32_t)std::min<uint64_t>(pcm_budget, 0xFFFFFFFFull); d.flags = DVDAGPU_PCM_BUDGET_IN_FRAMES; }
            tmark(x->stream, 'd');
            return decode_begin(x, x->buf[slot ? B_SECTORS2 : B_SECTORS].as<uint8_t>(), len, 1, &d);
        };
        uint32_t begun = 0, ended = 0;
        bool stop = false;                                   // the track ended inside a part: the parts behind are dropped
        // waits for part i, queues its download; sets fallback / stop
        auto end = [&](uint32_t i) -> int {
            dvdagpu_ctx *x = ctx_of(i);
            const int slot = slot_of(i);
            dvdagpu_track_result r;
            TRY(decode_end(x, &r));
            CUDA_TRY(cudaEventRecord(x->pev[1][slot], x->stream));
            tmark(x->stream, 'D');
            add_stats(total_stats, x->stats);
            if (stop || fallback) return 0;                  // (drained only)
            const bool same_params = !i || (r.channels == merged.channels && r.sample_rate == merged.sample_rate &&
                                            r.bits_per_sample == merged.bits_per_sample && r.channel_assignment == merged.channel_assignment);
            if (pcm) {
                // a window that does not open with a PCM packet of the same parameters: the track ended in front of it
                if (r.status != 0 || r.codec != 0 || !same_params) { stop = true; merged.truncated = 0; return 0; }
            } else if (i == 0 && r.status == 0 && r.codec == 0 && !(track->flags & 7u)) {
                pcm = true;
                const uint64_t total = (uint64_t)llround((double)track->pts_length * (double)r.sample_rate / 90000.0);
                pcm_budget = total;
                // the windows behind part 0 (all of its window was delivered, margin included)
                uint64_t s0, len, e_rel;
                window(0, s0, len, e_rel);
                starts.resize(1);
                for (uint64_t s1 = s0 + len; s1 < n_sectors; s1 += part_sectors) starts.push_back(s1);
                parts = (uint32_t)starts.size();
                restart_at = 1;
            } else if (r.status != 0 || r.codec != 1 || r.stopped == 2 || (r.truncated && i + 1 < parts) || !same_params) {
                // anything the parts cannot express: decode in one piece instead
                fallback = true;
                return 0;
            }
            if (i == 0) merged = r;
            if (pcm) {
                pcm_budget = r.frames >= pcm_budget ? 0 : pcm_budget - r.frames;
                if (!pcm_budget || !r.truncated) stop = true;      // used up, or the window ended early
            }
            const uint64_t n = r.frames * r.channels;
            if (total_samples + n > pcm_capacity) return 3;
            tmark(c->d2h_stream, 'o');
            if (n) CUDA_TRY(cudaMemcpyAsync(pcm_host + total_samples, x->buf[slot ? B_PCM2 : B_PCM].as<int32_t>() + r.pcm_offset,
                                            n * sizeof(int32_t), cudaMemcpyDeviceToHost, c->d2h_stream));
            CUDA_TRY(cudaEventRecord(x->pev[2][slot], c->d2h_stream));
            tmark(c->d2h_stream, 'O');
            total_samples += n;
            total_frames += r.frames;
            merged.error_flags |= r.error_flags;
            merged.truncated = r.truncated;
            if (r.stopped == 1) { merged.stopped = 1; stop = true; }
            return 0;
        };
        // whatever happens, nothing may stay in flight behind this call (the buffers are reused)
        auto drain = [&]() {
            while (ended < begun) { dvdagpu_track_result r; decode_end(ctx_of(ended), &r); ended++; }
            cudaStreamSynchronize(c->d2h_stream);
            cudaStreamSynchronize(c->h2d_stream);
        };
        int rc = upload(0);
        for (uint32_t i = 0; i < parts && !rc && !fallback && !stop; i++) {
            if (i + 1 < parts) rc = upload(i + 1);
            if (!rc) { rc = begin(i); if (!rc) begun++; }
            // with two contexts the host now waits for the part before this one, with one (and for PCM parts,
            // which need their predecessor's frame count) for this one
            while (!rc && ended < begun && (begun - ended > (two && !pcm ? 1u : 0u) || i + 1 == parts) && !fallback && !restart_at) { rc = end(ended); ended++; }
            if (restart_at && !rc) {
                // the track is PCM: whatever was begun behind part 0 is waited for and decoded again as PCM windows
                while (ended < begun) { dvdagpu_track_result r; decode_end(ctx_of(ended), &r); add_stats(total_stats, ctx_of(ended)->stats); ended++; }
                begun = ended = restart_at;
                i = restart_at - 1;
                restart_at = 0;
                if (!stop && i + 1 < parts) rc = upload(i + 1);      // (its window changed; the next round begins it)
            }
        }
        while (!rc && ended < begun) { rc = end(ended); ended++; }
        drain();
        if (ptrace && !fallback && !rc) {
            // u/U upload begins / ends, d/D decode, o/O download, in the order they were queued
            fprintf(stderr, "[pipe] %u parts on %d context(s); milliseconds from the first upload's begin:", parts, two ? 2 : 1);
            for (size_t e = 0; e < tev.size(); e++) {
                float ms = 0;
                cudaEventElapsedTime(&ms, tev[0], tev[e]);
                fprintf(stderr, " %c%.2f", tkind[e], ms);
            }
            fprintf(stderr, "\n");
        }
        for (auto e : tev) cudaEventDestroy(e);
        if (rc == 3) { dvdagpu_set_error("PCM buffer too small"); memset(result, 0, sizeof *result); return 3; }
        if (rc) { memset(result, 0, sizeof *result); return -1; }
    }
    if (fallback) {
        c->pcm_slot = 0;
        dvdagpu_track_result r;
        TRY(dvdagpu_decode_host(c, sectors, n_sectors, 1, track, &r));
        add_stats(total_stats, c->stats);
        *result = r;
        if (r.status == 0) {
            const uint64_t n = r.frames * r.channels;
            if (n > pcm_capacity) { dvdagpu_set_error("PCM buffer too small"); return 3; }
            TRY(dvdagpu_fetch(c, r.pcm_offset, n, pcm_host));
        }
        c->stats = total_stats;
        return 0;
    }
    merged.frames = total_frames;
    merged.pcm_offset = 0;
    *result = merged;
    c->stats = total_stats;
    c->stats.samples = total_samples;
    return 0;
}
