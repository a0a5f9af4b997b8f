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
    B_ES, B_SYNC_SLOTS, B_SYNC_CNT_RAW, B_SYNC_CNT_VALID, B_SYNC_BASE_RAW, B_SYNC_BASE_VALID, B_RAW, B_VALID,
    B_TRACKS, B_TRK_PK_LO, B_TRK_SEG_BASE, B_TRK_GRP_BASE,
    B_SEGS, B_SEG_NAU, B_SEG_AU_BASE, B_AU_POS, B_AU_ERR, B_AU, B_PSETS, B_AU_FRAMES,
    B_SS_FLAGS, B_SS_FLAGS_PREV, B_SS_FLAGS_FAST, B_FIR_TAIL,
    B_GROUPS, B_GRP_CELLS, B_CELL_BASE, B_GRP_CHUNKS, B_GRP_CHUNK_BASE,
    B_DEC_WORK, B_FUSED_WORK, B_SS_STICKY, B_AU_SEG, B_AU_SNAP, B_FILT_SNAP, B_AU_FCHG, B_SEG_CTX, B_AU_DELTA, B_TILES, B_BYPASS, B_SEG_FRAMES, B_SEG_FRAME_SCAN, B_SCAN_TMP, B_STATUS, B_AU_NOTED, B_PCM,
    B_COUNT
};

struct dvdagpu_ctx {
    int device;
    cudaStream_t own_stream;
    cudaStream_t stream;
    cudaStream_t h2d_stream, d2h_stream;      // copy engines of the pipelined path
    cudaStream_t aux_stream;                  // check data runs beside the header passes
    cudaEvent_t aux_ev[2];
    cudaStream_t aux_stream_hi = nullptr;     // side stream at the chain's own priority
    uint32_t pk_last_sectors = 0, pk_last_np = 0;   // sector and packet count of the previous decode
    uint64_t sync_last_es = ~0ull;                  // stream size and sync counts of the previous decode
    uint32_t sync_last_raw = 0, sync_last_valid = 0;
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

extern "C" void *dvdagpu_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        dvdagpu_set_error("cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return p;
}
extern "C" void dvdagpu_host_free(void *p) { if (p) cudaFreeHost(p); }

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

#define ENSURE(id, bytes) do { if (c->buf[id].ensure(bytes)) return -1; } while (0)
#define TRY(expr) do { if ((expr) != 0) return -1; } while (0)
// device time of one kernel (or a short run of kernels) into stats.kernel_ms[id]
#define TIMED(id, expr)                                                        \
    do {                                                                       \
        CUDA_TRY(cudaEventRecord(c->kev[id][0], s));                           \
        TRY(expr);                                                             \
        CUDA_TRY(cudaEventRecord(c->kev[id][1], s));                           \
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

// device -> host, synchronous
static int small_d2h(dvdagpu_ctx *c, void *host, const void *dev, size_t bytes)
{
    if (bytes > MAP_BYTES / 2 || (bytes & 3)) {
        CUDA_TRY(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        return 0;
    }
    // the upper half of the staging area is for read-backs
    LAUNCH(k_copy_words, 1, 256, 0, c->stream, (uint32_t *)(c->dmap + MAP_BYTES / 2), (const uint32_t *)dev, (uint32_t)(bytes / 4));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    trace_host("  (host has the read-back)");
    memcpy(host, c->hmap + MAP_BYTES / 2, bytes);
    return 0;
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

// Totals the scans leave in the mapped staging area themselves (scan_batch's total_copy): slot j
// as the device sees it, and the host's read once the stream has drained.
static uint64_t *mapped_total_slot(dvdagpu_ctx *c, int j) { return reinterpret_cast<uint64_t *>(c->dmap + MAP_BYTES / 2) + j; }
static int mapped_totals(dvdagpu_ctx *c, int n, uint64_t *host)
{
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    trace_host("  (host has the totals)");
    for (int j = 0; j < n; j++) host[j] = reinterpret_cast<volatile uint64_t *>(c->hmap + MAP_BYTES / 2)[j];
    return 0;
}

template <typename T>
static int read_back(dvdagpu_ctx *c, const T *dev, T *host)
{
    return small_d2h(c, host, dev, sizeof(T));
}

// several read-backs, one round trip (each one costs tens of microseconds, more while bulk copies run)
static int small_d2h_multi(dvdagpu_ctx *c, int n, void *const host[], const void *const dev[], const size_t bytes[])
{
    size_t off[H2D_BATCH], total = 0;
    bool ok = n <= H2D_BATCH;
    for (int i = 0; i < n && ok; i++) {
        off[i] = total;
        total += (bytes[i] + 15) & ~(size_t)15;
        if (bytes[i] & 3) ok = false;
    }
    if (!ok || total > MAP_BYTES / 2) {
        for (int i = 0; i < n; i++) TRY(small_d2h(c, host[i], dev[i], bytes[i]));
        return 0;
    }
    CopyBatch b;
    b.n = (uint32_t)n;
    for (int i = 0; i < n; i++) {
        b.dst[i] = (uint32_t *)(c->dmap + MAP_BYTES / 2 + off[i]); b.src[i] = (const uint32_t *)dev[i]; b.nwords[i] = (uint32_t)(bytes[i] / 4);
    }
    LAUNCH(k_copy_multi, 1, 256, 0, c->stream, b);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    trace_host("  (host has the read-backs)");
    for (int i = 0; i < n; i++) memcpy(host[i], c->hmap + MAP_BYTES / 2 + off[i], bytes[i]);
    return 0;
}
static int small_d2h_pair(dvdagpu_ctx *c, void *host_a, const void *dev_a, size_t bytes_a,
                          void *host_b, const void *dev_b, size_t bytes_b)
{
    void *const host[2] = {host_a, host_b};
    const void *const dev[2] = {dev_a, dev_b};
    const size_t bytes[2] = {bytes_a, bytes_b};
    return small_d2h_multi(c, 2, host, dev, bytes);
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

static int decode_on_device(dvdagpu_ctx *c, const uint8_t *d_sectors, uint64_t n_sectors64,
                            uint32_t n_tracks, const dvdagpu_track_desc *descs, dvdagpu_track_result *results)
{
    g_error[0] = 0;
    g_launch_count = 0;
    g_trace_on = getenv("DVDAGPU_TRACE") != nullptr;
    g_trace.n = 0;
    if (g_trace_on) trace_mark("decode begins", c->stream);
    c->map_used = 0;
    g_h2d_batch.n = 0;
    cudaStream_t s = c->stream;
    if (n_sectors64 == 0 || n_sectors64 > 0x7FFFFFFFull) { dvdagpu_set_error("bad sector count"); return -1; }
    if (!n_tracks) { c->pcm_samples = 0; return 0; }
    const uint32_t n_sectors = (uint32_t)n_sectors64;
    memset(&c->stats, 0, sizeof c->stats);
    memset(c->kev_used, 0, sizeof c->kev_used);
    CUDA_TRY(cudaEventRecord(c->ev[0], s));

    // tracks in sector order (the kernels binary-search them); results go back in caller order
    std::vector<uint32_t> order(n_tracks);
    for (uint32_t i = 0; i < n_tracks; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return descs[a].first_sector < descs[b].first_sector; });
    std::vector<TrackDev> &ht = c->h_tracks;
    ht.assign(n_tracks, TrackDev());
    for (uint32_t i = 0; i < n_tracks; i++) {
        memset(&ht[i], 0, sizeof(TrackDev));
        ht[i].first_sector = descs[order[i]].first_sector;
        ht[i].last_sector = descs[order[i]].last_sector;
        ht[i].pts_length = descs[order[i]].pts_length;
        ht[i].cont = descs[order[i]].flags & 3u;
    }

    // ---------------- demux
    ENSURE(B_SEC_CNT, (size_t)n_sectors * 4);
    ENSURE(B_SEC_BAD, (size_t)n_sectors * 4);
    ENSURE(B_SEC_BASE, (size_t)(n_sectors + 1) * 4);
    ENSURE(B_BAD_PREFIX, (size_t)(n_sectors + 1) * 4);
    ENSURE(B_SCAN_TMP, scan_tmp_bytes((uint64_t)n_sectors * 2048 / 32 + 4096));
    void *tmp = c->buf[B_SCAN_TMP].p;
    const size_t tmp_bytes = c->buf[B_SCAN_TMP].cap;
    if (c->scan_tmp_gen != c->buf[B_SCAN_TMP].gen) {
        // the scans find their control words zero and leave them zero (scan.cu)
        CUDA_TRY(cudaMemsetAsync(tmp, 0, tmp_bytes, s));
        c->scan_tmp_gen = c->buf[B_SCAN_TMP].gen;
    }
    uint32_t *sec_cnt = c->buf[B_SEC_CNT].as<uint32_t>(), *sec_bad = c->buf[B_SEC_BAD].as<uint32_t>();
    uint32_t *sec_base = c->buf[B_SEC_BASE].as<uint32_t>(), *bad_prefix = c->buf[B_BAD_PREFIX].as<uint32_t>();
    TRY(launch_sector_count(d_sectors, n_sectors, sec_cnt, sec_bad, s));
    {
        const uint32_t *in[2] = {sec_cnt, sec_bad}; void *out[2] = {sec_base, bad_prefix}; const bool wide[2] = {false, false};
        uint64_t *const copy[2] = {mapped_total_slot(c, 0), nullptr};
        TRY(scan_batch(in, out, wide, 2, n_sectors, tmp, tmp_bytes, s, copy));
    }
    // The packet table is sized before the host knows the packet count (one audio packet per
    // sector is the rule; room for a quarter more, and whatever earlier decodes needed): the
    // table is filled, its prefix sums taken over all its rows, and only then does the host
    // fetch the packet count and the size of the elementary stream, in one round trip.  If the
    // table was too small it grows and the step is repeated.
    uint32_t np = 0;
    uint64_t es_total = 0;
    PacketTable pt;
    uint64_t *pk_es = nullptr, *pk_pf = nullptr;
    uint32_t *nonmlp = nullptr, *nm_prefix = nullptr, *pstop = nullptr, *stop_prefix = nullptr;
    size_t rows = (size_t)n_sectors + n_sectors / 4 + 64;
    if (c->pk_last_sectors == n_sectors) rows = std::max<size_t>(rows, (size_t)c->pk_last_np + 64);   // (the same input again)
    // (test hook: DVDAGPU_SMALL_TABLES=1 starts every table sized in advance with room for one entry,
    // so that each decode goes through the grow-and-repeat paths)
    const bool small_tables = getenv("DVDAGPU_SMALL_TABLES") != nullptr;
    if (small_tables) rows = 1;
    for (int attempt = 0;; attempt++) {
        const size_t npa = rows + 1;
        ENSURE(B_PK_SECTOR, npa * 4); ENSURE(B_PK_OFF, npa * 2); ENSURE(B_PK_LEN, npa * 2);
        ENSURE(B_PK_CODEC, npa); ENSURE(B_PK_PAD2, npa); ENSURE(B_PK_PARAMS, npa * 4);
        ENSURE(B_PK_MLPLEN, npa * 4); ENSURE(B_PK_PCMF, npa * 4);
        ENSURE(B_PK_ES, npa * 8); ENSURE(B_PK_PF, npa * 8);
        ENSURE(B_PK_NONMLP, npa * 4); ENSURE(B_PK_NM_PREFIX, npa * 4);
        ENSURE(B_PK_STOP, npa * 4); ENSURE(B_PK_STOP_PREFIX, npa * 4);
        ENSURE(B_PK_YIELD, npa);
        pt.sector = c->buf[B_PK_SECTOR].as<uint32_t>(); pt.off = c->buf[B_PK_OFF].as<uint16_t>();
        pt.len = c->buf[B_PK_LEN].as<uint16_t>(); pt.codec = c->buf[B_PK_CODEC].as<uint8_t>();
        pt.pad2 = c->buf[B_PK_PAD2].as<uint8_t>(); pt.params = c->buf[B_PK_PARAMS].as<uint32_t>();
        pt.mlp_len = c->buf[B_PK_MLPLEN].as<uint32_t>(); pt.pcm_frames = c->buf[B_PK_PCMF].as<uint32_t>();
        pk_es = c->buf[B_PK_ES].as<uint64_t>(); pk_pf = c->buf[B_PK_PF].as<uint64_t>();
        nonmlp = c->buf[B_PK_NONMLP].as<uint32_t>(); nm_prefix = c->buf[B_PK_NM_PREFIX].as<uint32_t>();
        pstop = c->buf[B_PK_STOP].as<uint32_t>(); stop_prefix = c->buf[B_PK_STOP_PREFIX].as<uint32_t>();
        TRY(launch_packet_fill(d_sectors, n_sectors, sec_base, pt, (uint32_t)rows, nonmlp, pstop, s));
        {
            const uint32_t *in[4] = {pt.mlp_len, pt.pcm_frames, nonmlp, pstop};
            void *out[4] = {pk_es, pk_pf, nm_prefix, stop_prefix};
            const bool wide[4] = {true, true, false, false};
            uint64_t *const copy[4] = {mapped_total_slot(c, 1), nullptr, nullptr, nullptr};
            TRY(scan_batch(in, out, wide, 4, rows, tmp, tmp_bytes, s, copy));
        }
        uint64_t totals[2] = {0, 0};
        TRY(mapped_totals(c, 2, totals));
        np = (uint32_t)totals[0];
        es_total = totals[1];
        c->pk_last_sectors = n_sectors; c->pk_last_np = np;
        if (np <= rows) break;
        if (attempt) { dvdagpu_set_error("internal: packet table"); return -1; }
        rows = (size_t)np + np / 8;
    }

    ENSURE(B_ES, es_total + DVDA_ES_PAD);
    uint8_t *es = c->buf[B_ES].as<uint8_t>();
    CUDA_TRY(cudaMemsetAsync(es + es_total, 0, DVDA_ES_PAD, s));
    TIMED(DVDAGPU_K_ES_GATHER, launch_es_gather(d_sectors, pt, np, pk_es, es, s));
    CUDA_TRY(cudaEventRecord(c->ev[1], s));

    // ---------------- index
    const uint32_t chunks = div_up_u32(es_total ? es_total : 1, SYNC_CHUNK);
    ENSURE(B_SYNC_CNT_RAW, (size_t)chunks * 4); ENSURE(B_SYNC_CNT_VALID, (size_t)chunks * 4);
    ENSURE(B_SYNC_BASE_RAW, (size_t)(chunks + 1) * 4); ENSURE(B_SYNC_BASE_VALID, (size_t)(chunks + 1) * 4);
    ENSURE(B_SYNC_SLOTS, (size_t)chunks * SYNC_SLOT_BYTES);
    uint16_t *sync_slots = c->buf[B_SYNC_SLOTS].as<uint16_t>();
    // (test hook: DVDAGPU_SYNC_SLOTS=0 sends every chunk with a match through the re-search path)
    const uint32_t nslots = getenv("DVDAGPU_SYNC_SLOTS") ? (uint32_t)atoi(getenv("DVDAGPU_SYNC_SLOTS")) : 2u;
    uint32_t *cnt_raw = c->buf[B_SYNC_CNT_RAW].as<uint32_t>(), *cnt_valid = c->buf[B_SYNC_CNT_VALID].as<uint32_t>();
    uint32_t *base_raw = c->buf[B_SYNC_BASE_RAW].as<uint32_t>(), *base_valid = c->buf[B_SYNC_BASE_VALID].as<uint32_t>();
    uint32_t n_raw = 0, n_valid = 0;
    if (es_total) {
        TIMED(DVDAGPU_K_SYNC_SCAN, launch_sync_count(es, es_total, cnt_raw, cnt_valid, sync_slots, nslots, s));
        const uint32_t *in[2] = {cnt_raw, cnt_valid}; void *out[2] = {base_raw, base_valid}; const bool wide[2] = {false, false};
        TRY(scan_batch(in, out, wide, 2, chunks, tmp, tmp_bytes, s));
    } else {
        CUDA_TRY(cudaMemsetAsync(base_raw + chunks, 0, 4, s));
        CUDA_TRY(cudaMemsetAsync(base_valid + chunks, 0, 4, s));
    }
    ENSURE(B_TRACKS, (size_t)n_tracks * sizeof(TrackDev));
    ENSURE(B_TRK_PK_LO, (size_t)n_tracks * 4); ENSURE(B_TRK_SEG_BASE, (size_t)(n_tracks + 1) * 4);
    ENSURE(B_TRK_GRP_BASE, (size_t)(n_tracks + 1) * 4);
    TrackDev *d_tracks = c->buf[B_TRACKS].as<TrackDev>();
    TRY(small_h2d(c, d_tracks, ht.data(), n_tracks * sizeof(TrackDev)));
    // The sync lists are sized before their lengths are known to the host (a sync every 2 KiB of
    // stream, or what the same input needed last time): they are filled, the tracks set up from
    // them, and the lengths come back together with the track table.  Lists that were too small
    // grow and the step is repeated.
    uint64_t *raw = nullptr, *valid = nullptr;
    size_t cap_raw = (size_t)(es_total / 2048) + 1024, cap_valid = cap_raw;
    if (c->sync_last_es == es_total) { cap_raw = std::max<size_t>(cap_raw, c->sync_last_raw); cap_valid = std::max<size_t>(cap_valid, c->sync_last_valid); }
    if (small_tables) cap_raw = cap_valid = 1;
    for (int attempt = 0;; attempt++) {
        ENSURE(B_RAW, (cap_raw + 1) * 8); ENSURE(B_VALID, (cap_valid + 1) * 8);
        raw = c->buf[B_RAW].as<uint64_t>(); valid = c->buf[B_VALID].as<uint64_t>();
        if (es_total) TRY(launch_sync_fill(es, es_total, cnt_raw, sync_slots, nslots, base_raw, base_valid,
                                           raw, (uint32_t)cap_raw, valid, (uint32_t)cap_valid, s));
        TrackSetupArgs ta;
        ta.es = es; ta.es_total = es_total; ta.n_sectors = n_sectors; ta.sec_base = sec_base; ta.bad_prefix = bad_prefix;
        ta.pt = pt; ta.np = np; ta.pk_es = pk_es; ta.pk_pf = pk_pf; ta.pk_nonmlp = nm_prefix; ta.pk_pcm_stop = stop_prefix;
        ta.raw = raw; ta.n_raw = base_raw + chunks; ta.cap_raw = (uint32_t)cap_raw;
        ta.valid = valid; ta.n_valid = base_valid + chunks; ta.cap_valid = (uint32_t)cap_valid;
        TRY(launch_track_setup(ta, d_tracks, n_tracks, s));
        {
            void *const host[3] = {&n_raw, &n_valid, ht.data()};
            const void *const dev[3] = {base_raw + chunks, base_valid + chunks, d_tracks};
            const size_t bytes[3] = {4, 4, n_tracks * sizeof(TrackDev)};
            TRY(small_d2h_multi(c, 3, host, dev, bytes));
        }
        c->sync_last_es = es_total; c->sync_last_raw = n_raw; c->sync_last_valid = n_valid;
        if (n_raw <= cap_raw && n_valid <= cap_valid) break;
        if (attempt) { dvdagpu_set_error("internal: sync lists"); return -1; }
        cap_raw = n_raw; cap_valid = n_valid;
    }

    std::vector<uint32_t> h_pk_lo(n_tracks), h_seg_base(n_tracks + 1), h_grp_base(n_tracks + 1);
    uint32_t nseg = 0, ngroups = 0;
    for (uint32_t i = 0; i < n_tracks; i++) {
        h_pk_lo[i] = ht[i].pk_lo;
        if (ht[i].status != 0 || ht[i].codec != 1) { ht[i].nseg = 0; ht[i].ngrp = 0; }
        ht[i].seg_base = nseg; ht[i].grp_base = ngroups;
        h_seg_base[i] = nseg; h_grp_base[i] = ngroups;
        nseg += ht[i].nseg; ngroups += ht[i].ngrp;
    }
    h_seg_base[n_tracks] = nseg; h_grp_base[n_tracks] = ngroups;
    // decode work lists, one per channel-count class (0 = generic, more than 4 channels):
    // substream 0 of a two-substream stream carries the stereo pair (DVD-Audio layout)
    std::vector<DecWork> h_work[5];
    uint32_t n_warps[5] = {0, 0, 0, 0, 0};
    for (uint32_t i = 0; i < n_tracks; i++) {
        if (!ht[i].nseg) continue;
        for (uint32_t k = 0; k < ht[i].nss; k++) {
            const uint32_t nch = ht[i].nss == 1 ? ht[i].channels : (k == 0 ? 2 : ht[i].channels - 2);
            const uint32_t cls = (nch >= 1 && nch <= 4) ? nch : 0;
            DecWork w = {n_warps[cls], i, k, 0};
            h_work[cls].push_back(w);
            n_warps[cls] += ht[i].ngrp;
        }
    }
    uint32_t *trk_pk_lo = c->buf[B_TRK_PK_LO].as<uint32_t>(), *trk_seg_base = c->buf[B_TRK_SEG_BASE].as<uint32_t>();
    uint32_t *trk_grp_base = c->buf[B_TRK_GRP_BASE].as<uint32_t>();
    TRY(queue_h2d(c, d_tracks, ht.data(), n_tracks * sizeof(TrackDev)));
    TRY(queue_h2d(c, trk_pk_lo, h_pk_lo.data(), n_tracks * 4));
    TRY(queue_h2d(c, trk_seg_base, h_seg_base.data(), (n_tracks + 1) * 4));
    TRY(queue_h2d(c, trk_grp_base, h_grp_base.data(), (n_tracks + 1) * 4));

    // decode mode: the three-pass path (default), the header passes + the fused entropy / filter /
    // output pass (DVDAGPU_FUSED=1: no tiles in HBM, but measured slower — DESIGN.md section 6), or the
    // complete single-pass decoder alone (DVDAGPU_SINGLE_PASS=1); the GPU tests run all three
    const int mode = getenv("DVDAGPU_SINGLE_PASS") ? 0 : getenv("DVDAGPU_FUSED") ? 2 : 1;
    // work lists of the fused pass, by class (0: at most two channels per substream, 1: up to four):
    // a run of warps per track
    std::vector<FusedWork> h_fused[2];
    uint32_t n_fwarps[2] = {0, 0};
    if (mode == 2) {
        for (uint32_t i = 0; i < n_tracks; i++) {
            if (!ht[i].nseg) continue;
            uint32_t n0 = ht[i].channels, n1 = 0;
            if (ht[i].nss == 2) { n0 = 2; n1 = ht[i].channels > 2 ? ht[i].channels - 2 : 0; if (!n1) continue; }
            else if (ht[i].nss != 1) continue;
            if (n0 < 1 || n0 > 4 || n1 > 4) continue;
            const int cl = (n0 > 2 || n1 > 2) ? 1 : 0;
            FusedWork w = {n_fwarps[cl], i, n0, n1};
            h_fused[cl].push_back(w);
            n_fwarps[cl] += ht[i].ngrp * fused_warps_per_group(n0, n1);
        }
    }
    const FusedWork *d_fused[2] = {nullptr, nullptr};
    uint32_t n_fused[2] = {0, 0};
    {
        ENSURE(B_FUSED_WORK, (h_fused[0].size() + h_fused[1].size() + 1) * sizeof(FusedWork));
        FusedWork *base = c->buf[B_FUSED_WORK].as<FusedWork>();
        size_t off = 0;
        for (int cl = 0; cl < 2; cl++) {
            n_fused[cl] = (uint32_t)h_fused[cl].size();
            d_fused[cl] = base + off;
            if (n_fused[cl]) TRY(queue_h2d(c, base + off, h_fused[cl].data(), n_fused[cl] * sizeof(FusedWork)));
            off += n_fused[cl];
        }
    }
    const DecWork *d_work[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    uint32_t n_work[5] = {0, 0, 0, 0, 0};
    {
        size_t total = 0;
        for (int cl = 0; cl < 5; cl++) total += h_work[cl].size();
        ENSURE(B_DEC_WORK, (total + 1) * sizeof(DecWork));
        DecWork *base = c->buf[B_DEC_WORK].as<DecWork>();
        size_t off = 0;
        for (int cl = 0; cl < 5; cl++) {
            n_work[cl] = (uint32_t)h_work[cl].size();
            d_work[cl] = base + off;
            if (n_work[cl])
                TRY(queue_h2d(c, base + off, h_work[cl].data(), n_work[cl] * sizeof(DecWork)));
            off += n_work[cl];
        }
        TRY(flush_h2d(c));
    }
    MlpTables m;
    memset(&m, 0, sizeof m);
    m.es = es; m.es_total = es_total; m.pk_es = pk_es; m.np = np;
    m.tracks = d_tracks; m.n_tracks = n_tracks; m.nseg = nseg; m.ngroups = ngroups;
    uint32_t nau = 0;
    uint32_t max_chunks = 0, status = 0;
    ENSURE(B_STATUS, 64);
    uint32_t *d_status = c->buf[B_STATUS].as<uint32_t>();
    m.status = d_status;
    m.status_rw = d_status;
    m.any_fallback = d_status + 8;
    m.huff_lut = c->huff_lut;
    // samples of the PCM tracks (known since the track set-up) and the alignment gaps
    uint64_t pcm_fixed = 4ull * n_tracks + 64;
    for (uint32_t i = 0; i < n_tracks; i++) if (ht[i].status == 0 && ht[i].codec == 0) pcm_fixed += ht[i].frames * ht[i].channels;
    const int pcm_buf = c->pcm_slot ? B_PCM2 : B_PCM;
    bool out_queued = false;
    if (nseg) {
        ENSURE(B_SEGS, (size_t)nseg * sizeof(SegDev));
        ENSURE(B_SEG_NAU, (size_t)nseg * 4); ENSURE(B_SEG_AU_BASE, (size_t)(nseg + 1) * 4);
        m.segs = c->buf[B_SEGS].as<SegDev>();
        uint32_t *seg_nau = c->buf[B_SEG_NAU].as<uint32_t>(), *seg_au_base = c->buf[B_SEG_AU_BASE].as<uint32_t>();
        TRY(launch_segment_fill(d_tracks, n_tracks, trk_seg_base, valid, m.segs, nseg, s));
        ENSURE(B_AU_NOTED, au_noted_bytes(nseg));
        uint32_t *au_noted = c->buf[B_AU_NOTED].as<uint32_t>();
        TRY(launch_au_chase(es, m.segs, nseg, d_tracks, seg_nau, nullptr, nullptr, seg_au_base, au_noted, 0, s));
        TRY(scan_u32_to_u32(seg_nau, seg_au_base, nseg, tmp, tmp_bytes, s));
        // the groups (tile sizes) follow from the access-unit counts alone: set them up now and
        // fetch all the sizes the next allocations need in one round trip
        ENSURE(B_GROUPS, (size_t)ngroups * sizeof(GroupDev));
        ENSURE(B_GRP_CELLS, (size_t)ngroups * 4); ENSURE(B_CELL_BASE, (size_t)(ngroups + 1) * 8);
        ENSURE(B_GRP_CHUNKS, (size_t)ngroups * 4);
        m.groups = c->buf[B_GROUPS].as<GroupDev>();
        uint32_t *grp_cells = c->buf[B_GRP_CELLS].as<uint32_t>(), *grp_chunks = c->buf[B_GRP_CHUNKS].as<uint32_t>();
        uint64_t *cell_base = c->buf[B_CELL_BASE].as<uint64_t>();
        uint64_t cells = 0;
        struct { uint32_t au, chunks; } most = {0, 0};
        CUDA_TRY(cudaMemsetAsync(d_status, 0, 64, s));
        TRY(launch_group_setup(d_tracks, n_tracks, trk_grp_base, m.segs, m.groups, ngroups, grp_cells, grp_chunks, d_status + 1, s));
        TRY(scan_u32_to_u64(grp_cells, cell_base, ngroups, tmp, tmp_bytes, s));
        {
            void *const host[3] = {&nau, &cells, &most};
            const void *const dev[3] = {seg_au_base + nseg, cell_base + ngroups, d_status + 1};
            const size_t bytes[3] = {4, 8, sizeof most};
            TRY(small_d2h_multi(c, 3, host, dev, bytes));
        }
        m.nau = nau;
        const size_t naua = (size_t)nau + 1;
        ENSURE(B_AU_POS, naua * 8); ENSURE(B_AU_ERR, naua); ENSURE(B_AU, naua * sizeof(AuDev)); ENSURE(B_AU_SEG, naua * 4);
        ENSURE(B_PSETS, naua * sizeof(ParamSet)); ENSURE(B_AU_FRAMES, naua * 2 * 4);
        ENSURE(B_SS_FLAGS, (size_t)nseg * 2 * 4); ENSURE(B_SS_FLAGS_PREV, (size_t)nseg * 2 * 4); ENSURE(B_SS_FLAGS_FAST, (size_t)nseg * 2 * 4);
        ENSURE(B_FIR_TAIL, (size_t)nseg * 2 * DVDA_MAX_CH * 8 * 4);
        m.au_pos = c->buf[B_AU_POS].as<uint64_t>(); m.au_err = c->buf[B_AU_ERR].as<uint8_t>(); m.au_seg = c->buf[B_AU_SEG].as<uint32_t>();
        m.au = c->buf[B_AU].as<AuDev>(); m.psets = c->buf[B_PSETS].as<ParamSet>();
        m.au_frames_ss = c->buf[B_AU_FRAMES].as<uint32_t>();
        m.ss_flags = c->buf[B_SS_FLAGS].as<uint32_t>(); m.ss_flags_prev = c->buf[B_SS_FLAGS_PREV].as<uint32_t>(); m.ss_flags_fast = c->buf[B_SS_FLAGS_FAST].as<uint32_t>();
        m.fir_tail = c->buf[B_FIR_TAIL].as<int32_t>();
        TRY(launch_au_chase(es, m.segs, nseg, d_tracks, seg_nau, m.au_pos, m.au_seg, seg_au_base, au_noted, 1, s));
        TRY(launch_yield(m, seg_au_base, pt, trk_pk_lo, c->buf[B_PK_YIELD].as<uint8_t>(), s));
        CUDA_TRY(cudaEventRecord(c->ev[2], s));

        // ---------------- decode
        // Parity / CRC-8 on a second stream, beside the group set-up and the header passes.  Small
        // access units: the windowed kernel on the low-priority stream (it fills what the chain
        // leaves free).  Large ones: the direct kernel at the chain's own priority, which then runs
        // first and lets the header passes follow.  Both pairings, and starting the check beside the
        // entropy pass or on its own in between, were measured: DESIGN.md section 6.
        // (a caller's stream has the default = lowest priority, like aux_stream)
        cudaStream_t chk_stream = (checkdata_windowed(m) || s != c->own_stream) ? c->aux_stream : c->aux_stream_hi;
        CUDA_TRY(cudaEventRecord(c->aux_ev[0], s));
        CUDA_TRY(cudaStreamWaitEvent(chk_stream, c->aux_ev[0], 0));
        CUDA_TRY(cudaEventRecord(c->kev[DVDAGPU_K_CHECKDATA][0], chk_stream));
        TRY(launch_checkdata(m, seg_au_base, chk_stream));
        CUDA_TRY(cudaEventRecord(c->kev[DVDAGPU_K_CHECKDATA][1], chk_stream));
        c->kev_used[DVDAGPU_K_CHECKDATA] = true;
        CUDA_TRY(cudaEventRecord(c->aux_ev[1], chk_stream));
        ENSURE(B_SEG_FRAMES, (size_t)nseg * 4); ENSURE(B_SEG_FRAME_SCAN, (size_t)(nseg + 1) * 8);
        uint32_t *seg_frames = c->buf[B_SEG_FRAMES].as<uint32_t>();
        uint64_t *seg_frame_scan = c->buf[B_SEG_FRAME_SCAN].as<uint64_t>();

        // The three-pass path (access-unit parallel) decodes what has the common shape; the complete
        // single-pass decoder takes the rest.  DVDAGPU_SINGLE_PASS=1 gives everything to the latter
        // (the GPU tests run both ways).
        const bool use_fast = mode != 0;
        if (use_fast) {
            ENSURE(B_SS_STICKY, (size_t)nseg * 2 * 4);
            m.ss_sticky = c->buf[B_SS_STICKY].as<uint32_t>();
            CUDA_TRY(cudaMemsetAsync(m.ss_sticky, 0, (size_t)nseg * 2 * 4, s));
            m.nss_max = 1;
            for (uint32_t i = 0; i < n_tracks; i++) if (ht[i].nseg && ht[i].nss > m.nss_max) m.nss_max = ht[i].nss;
            // per-substream tables: [nss_max][nau] (the kernels index them as k * nau + A)
            ENSURE(B_AU_SNAP, naua * m.nss_max * au_snap_bytes());
            m.au_snap = reinterpret_cast<AuSnap *>(c->buf[B_AU_SNAP].p);
            ENSURE(B_AU_FCHG, naua * m.nss_max); ENSURE(B_SEG_CTX, (size_t)nseg * 2 * seg_ctx_bytes());
            ENSURE(B_AU_DELTA, naua * m.nss_max * au_delta_bytes());
            m.au_fchg = c->buf[B_AU_FCHG].as<uint8_t>();
            m.seg_ctx = reinterpret_cast<SegCtx *>(c->buf[B_SEG_CTX].p);
            m.au_delta = reinterpret_cast<AuDelta *>(c->buf[B_AU_DELTA].p);
            // contexts exist only where pass A0 goes (substreams of up to four channels)
            CUDA_TRY(cudaMemsetAsync(m.seg_ctx, 0, (size_t)nseg * 2 * seg_ctx_bytes(), s));
        }
        for (int attempt = 0;; attempt++) {
            if (attempt) {
                // after a tile overflow: the groups again, from the frame counts now known
                // (after a STATUS_REDO of the fused pass the same steps are simply repeated)
                CUDA_TRY(cudaMemsetAsync(d_status, 0, 64, s));
                TRY(launch_group_setup(d_tracks, n_tracks, trk_grp_base, m.segs, m.groups, ngroups, grp_cells, grp_chunks, d_status + 1, s));
                TRY(scan_u32_to_u64(grp_cells, cell_base, ngroups, tmp, tmp_bytes, s));
                TRY(small_d2h_pair(c, &cells, cell_base + ngroups, 8, &most, d_status + 1, sizeof most));
            }
            CUDA_TRY(cudaMemsetAsync(d_status, 0, 4, s));         // (the maxima behind it stay)
            m.max_au = most.au; max_chunks = most.chunks;
            ENSURE(B_TILES, (cells * DVDA_LANES + 64 + 16 * DVDA_MAX_CH * DVDA_LANES) * sizeof(int32_t));   // 16 frames of slack: the filter passes read ahead
            ENSURE(B_BYPASS, cells * DVDA_LANES + 64);
            m.tiles = c->buf[B_TILES].as<int32_t>(); m.bypass = c->buf[B_BYPASS].as<uint8_t>();
            TRY(launch_group_offsets(m.groups, ngroups, cell_base, s));
            m.fast = 0;
            if (use_fast) {
                // three passes with access-unit parallelism; what they give up on is flagged ...
                // (start with every segment flagged: substreams with more than 4 channels
                // are not visited by the fast path at all)
                CUDA_TRY(cudaMemsetAsync(m.ss_flags, SEG_FALLBACK, (size_t)nseg * 2 * 4, s));
                TRY(launch_mlp_fast(m, d_work, n_work, n_warps, c->kev, c->kev_used, c->aux_ev[1], mode == 2, s));
                CUDA_TRY(cudaMemcpyAsync(m.ss_flags_prev, m.ss_flags, (size_t)nseg * 2 * 4, cudaMemcpyDeviceToDevice, s));
                CUDA_TRY(cudaMemcpyAsync(m.ss_flags_fast, m.ss_flags, (size_t)nseg * 2 * 4, cudaMemcpyDeviceToDevice, s));
                m.fast = (uint32_t)mode;
            }
            CUDA_TRY(cudaStreamWaitEvent(s, c->aux_ev[1], 0));
            // ... and decoded by the complete single-pass decoder (everything, without the fast path)
            TIMED(DVDAGPU_K_MLP_DECODE, launch_mlp_decode(m, d_work, n_work, n_warps, s));
            CUDA_TRY(cudaMemcpyAsync(m.ss_flags_prev, m.ss_flags, (size_t)nseg * 2 * 4, cudaMemcpyDeviceToDevice, s));
            TIMED(DVDAGPU_K_CARRY_FIX, launch_carry_fix(m, s));
            TRY(launch_seg_finalize(m, seg_frames, d_status, s));
            // the track totals are computed before the status is known (one round trip for both);
            // after an overflow they are simply computed again
            TRY(scan_u32_to_u64(seg_frames, seg_frame_scan, nseg, tmp, tmp_bytes, s));
            TRY(launch_track_finalize(m, seg_frame_scan, d_status, s));
            // The output buffer is sized before the frame counts are known (every MLP sample has a
            // place in the tiles, so the tiles' size bounds them), the tracks' places in it are
            // computed on the device, and the fused output pass of the fast path is queued right
            // here: the round trip below then costs the device nothing.
            // (test hook: a capacity of one sample sends every decode through the "too small" path)
            const uint64_t pcm_capacity = small_tables ? 1 : cells * DVDA_LANES + pcm_fixed;
            ENSURE(pcm_buf, (pcm_capacity + 64) * sizeof(int32_t));
            m.pcm = c->buf[pcm_buf].as<int32_t>();
            LAUNCH(k_track_out_base, 1, OB_THREADS, 0, s, d_tracks, n_tracks, d_status, pcm_capacity);
            CUDA_TRY(cudaEventRecord(c->ev[3], s));
            out_queued = m.fast != 0;
            if (m.fast == 1) TIMED(DVDAGPU_K_MLP_FILTER_OUT, launch_mlp_filter_out(m, d_work, n_work, n_warps, s));
            if (m.fast == 2) TIMED(DVDAGPU_K_MLP_FUSED, launch_mlp_fused(m, d_fused, n_fused, n_fwarps, s));
            TRY(small_d2h_pair(c, &status, d_status, 4, ht.data(), d_tracks, n_tracks * sizeof(TrackDev)));
            if (!(status & (SEG_OVERFLOW | STATUS_REDO))) break;
            // (a redo flags at least one more segment each time; an overflow is settled by the second attempt)
            if (attempt == 8) { dvdagpu_set_error((status & SEG_OVERFLOW) ? "tile overflow persists" : "the fused pass keeps asking for another attempt"); return -1; }
        }
    } else {
        CUDA_TRY(cudaEventRecord(c->ev[2], s));
        CUDA_TRY(cudaMemsetAsync(d_status, 0, 4, s));
        LAUNCH(k_track_out_base, 1, OB_THREADS, 0, s, d_tracks, n_tracks, d_status, pcm_fixed);
        TRY(small_d2h(c, ht.data(), d_tracks, n_tracks * sizeof(TrackDev)));
        CUDA_TRY(cudaEventRecord(c->ev[3], s));
    }
    if (getenv("DVDAGPU_DEBUG")) {
        fprintf(stderr, "[dvdagpu] sectors=%u packets=%u es=%llu raw=%u valid=%u segs=%u groups=%u aus=%u\n",
                n_sectors, np, (unsigned long long)es_total, n_raw, n_valid, nseg, ngroups, nau);
        for (uint32_t i = 0; i < n_tracks; i++) {
            const TrackDev &T = ht[i];
            fprintf(stderr, "[dvdagpu] track %u: status=%d codec=%d err=%x ch=%u pk=[%u,%u) pk_x=%u pk_open=%u es=[%llu,%llu) cut=%llu nss=%u nseg=%u err_seg=%u frames=%llu trunc=%u\n",
                    i, T.status, T.codec, T.error_flags, T.channels, T.pk_lo, T.pk_hi, T.pk_x, T.pk_open,
                    (unsigned long long)T.es_start, (unsigned long long)T.es_end, (unsigned long long)T.es_cut,
                    T.nss, T.nseg, T.err_seg, (unsigned long long)T.frames, T.truncated);
            fprintf(stderr, "[dvdagpu]          cont=%u check=[%u,%u) stopped=%u\n", T.cont, T.pk_check, T.pk_check_end, T.stopped);
        }
        if (nseg) {
            std::vector<SegDev> hs(std::min<uint32_t>(nseg, 8));
            cudaMemcpy(hs.data(), m.segs, hs.size() * sizeof(SegDev), cudaMemcpyDeviceToHost);
            for (size_t i = 0; i < hs.size(); i++)
                fprintf(stderr, "[dvdagpu] seg %zu: es=[%llu,%llu) n_au=%u au_base=%u flags=%x frames=%u err=%x err_au=%u frame0=%llu\n",
                        i, (unsigned long long)hs[i].es_pos, (unsigned long long)hs[i].es_limit, hs[i].n_au, hs[i].au_base,
                        hs[i].flags, hs[i].frames, hs[i].err, hs[i].err_au, (unsigned long long)hs[i].frame0);
        }
    }

    // ---------------- output
    uint64_t total_samples = 0;
    bool any_pcm = false;
    uint32_t mlp_channel_mask = 0;
    for (uint32_t i = 0; i < n_tracks; i++) if (ht[i].status == 0 && ht[i].codec == 1) mlp_channel_mask |= 1u << ht[i].channels;
    for (uint32_t i = 0; i < n_tracks; i++) {
        total_samples = (total_samples + 3) & ~3ull;          // 16-byte aligned tracks (vector stores)
        if (ht[i].out_base != total_samples) { dvdagpu_set_error("internal: output layout"); return -1; }   // k_track_out_base
        if (ht[i].status == 0) total_samples += ht[i].frames * ht[i].channels;
        any_pcm |= ht[i].status == 0 && ht[i].codec == 0;
    }
    ENSURE(pcm_buf, (total_samples + 64) * sizeof(int32_t));
    m.pcm = c->buf[pcm_buf].as<int32_t>();
    c->pcm_samples = total_samples;
    if (status & STATUS_PCM_SMALL) {
        // (does not happen as long as the tiles bound the samples; the buffer has its real size now)
        CUDA_TRY(cudaMemsetAsync(d_status, 0, 4, s));
        out_queued = false;
    }
    if (nseg && m.fast && !out_queued) {
        // fast path: filters (+ entropy decode) + rematrix + interleaved output in one pass
        if (m.fast == 1) TIMED(DVDAGPU_K_MLP_FILTER_OUT, launch_mlp_filter_out(m, d_work, n_work, n_warps, s));
        else TIMED(DVDAGPU_K_MLP_FUSED, launch_mlp_fused(m, d_fused, n_fused, n_fwarps, s));
    }
    if (nseg && max_chunks && (!m.fast || (status & STATUS_WANTS_REMATRIX))) TIMED(DVDAGPU_K_REMATRIX, launch_rematrix(m, max_chunks, mlp_channel_mask, s));
    if (any_pcm) TIMED(DVDAGPU_K_PCM_UNPACK, launch_pcm_unpack(d_sectors, pt, np, pk_pf, d_tracks, trk_pk_lo, n_tracks, m.pcm, s));
    CUDA_TRY(cudaEventRecord(c->ev[4], s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (g_trace_on) { trace_host("decode done"); trace_dump(); g_trace_on = false; }

    // ---------------- results
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
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]); c->stats.demux_ms = ms;
    cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]); c->stats.index_ms = ms;
    cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]); c->stats.decode_ms = ms;
    cudaEventElapsedTime(&ms, c->ev[3], c->ev[4]); c->stats.output_ms = ms;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[4]); c->stats.total_ms = ms;
    for (int k = 0; k < 16; k++) {
        if (c->kev_used[k] && cudaEventElapsedTime(&ms, c->kev[k][0], c->kev[k][1]) == cudaSuccess) c->stats.kernel_ms[k] = ms;
    }
    c->stats.launches = g_launch_count;
    c->stats.segments = nseg;
    c->stats.access_units = nau;
    c->stats.es_bytes = es_used;
    c->stats.samples = total_samples;
    return 0;
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
    if (!part_sectors) {
        // a decode has a latency floor of a few milliseconds whatever its size, so few, large
        // parts: about 75 MB of AOB each, between 2 and 8 of them
        const uint64_t n = last >= first ? last - first + 1 : 0;
        uint64_t parts = (n + 19000) / 38000;
        parts = parts < 2 ? 2 : parts > 8 ? 8 : parts;
        part_sectors = (uint32_t)((n + parts - 1) / parts);
        if (part_sectors < 8192) part_sectors = 8192;
    }
    const uint64_t margin = 64;
    dvdagpu_stats total_stats;
    memset(&total_stats, 0, sizeof total_stats);

    bool fallback = first >= n_sectors || last < first || (last - first + 1) < 2ull * part_sectors;
    uint64_t total_samples = 0, total_frames = 0;
    dvdagpu_track_result merged;
    memset(&merged, 0, sizeof merged);
    if (!fallback) {
        const uint32_t parts = (uint32_t)((last - first + 1 + part_sectors - 1) / part_sectors);
        auto window = [&](uint32_t i, uint64_t &s0, uint64_t &len, uint64_t &e_rel) {
            s0 = first + (uint64_t)i * part_sectors;
            uint64_t e = s0 + part_sectors - 1;
            if (e > last || i + 1 == parts) e = last;
            uint64_t stop = (i + 1 == parts) ? n_sectors : e + 1 + margin;
            if (stop > n_sectors) stop = n_sectors;
            len = stop - s0;
            e_rel = e - s0;
        };
        auto upload = [&](uint32_t i) -> int {
            uint64_t s0, len, e_rel;
            window(i, s0, len, e_rel);
            const int slot = i & 1, buf = slot ? B_SECTORS2 : B_SECTORS;
            ENSURE(buf, len * DVDA_SECTOR + 256);
            if (i >= 2) CUDA_TRY(cudaStreamWaitEvent(c->h2d_stream, c->pev[1][slot], 0));   // part i-2 decoded
            CUDA_TRY(cudaMemcpyAsync(c->buf[buf].p, sectors + s0 * DVDA_SECTOR, len * DVDA_SECTOR,
                                     cudaMemcpyHostToDevice, c->h2d_stream));
            CUDA_TRY(cudaEventRecord(c->pev[0][slot], c->h2d_stream));
            return 0;
        };
        TRY(upload(0));
        for (uint32_t i = 0; i < parts && !fallback; i++) {
            if (i + 1 < parts) TRY(upload(i + 1));
            uint64_t s0, len, e_rel;
            window(i, s0, len, e_rel);
            const int slot = i & 1;
            CUDA_TRY(cudaStreamWaitEvent(c->stream, c->pev[0][slot], 0));
            if (i >= 2) CUDA_TRY(cudaStreamWaitEvent(c->stream, c->pev[2][slot], 0));       // PCM slot downloaded
            c->pcm_slot = slot;
            dvdagpu_track_desc d = {0, (uint32_t)e_rel, track->pts_length,
                                    (i ? (uint32_t)DVDAGPU_PART_CONTINUES_PREVIOUS : (track->flags & 1u)) |
                                    (i + 1 < parts ? (uint32_t)DVDAGPU_PART_CONTINUED_BY_NEXT : (track->flags & 2u))};
            dvdagpu_track_result r;
            TRY(decode_on_device(c, c->buf[slot ? B_SECTORS2 : B_SECTORS].as<uint8_t>(), len, 1, &d, &r));
            CUDA_TRY(cudaEventRecord(c->pev[1][slot], c->stream));
            add_stats(total_stats, c->stats);
            // anything the parts cannot express: decode in one piece instead
            if (r.status != 0 || r.codec != 1 || r.stopped == 2 || (r.truncated && i + 1 < parts) ||
                (i && (r.channels != merged.channels || r.sample_rate != merged.sample_rate))) { fallback = true; break; }
            if (i == 0) merged = r;
            const uint64_t n = r.frames * r.channels;
            if (total_samples + n > pcm_capacity) {
                cudaStreamSynchronize(c->d2h_stream);
                dvdagpu_set_error("PCM buffer too small");
                result->frames = 0;
                return 3;
            }
            CUDA_TRY(cudaStreamWaitEvent(c->d2h_stream, c->pev[1][slot], 0));
            if (n) CUDA_TRY(cudaMemcpyAsync(pcm_host + total_samples, c->buf[slot ? B_PCM2 : B_PCM].as<int32_t>() + r.pcm_offset,
                                            n * sizeof(int32_t), cudaMemcpyDeviceToHost, c->d2h_stream));
            CUDA_TRY(cudaEventRecord(c->pev[2][slot], c->d2h_stream));
            total_samples += n;
            total_frames += r.frames;
            merged.error_flags |= r.error_flags;
            merged.truncated = r.truncated;
            if (r.stopped == 1) { merged.stopped = 1; break; }      // the track ended inside this part
        }
        CUDA_TRY(cudaStreamSynchronize(c->d2h_stream));
        CUDA_TRY(cudaStreamSynchronize(c->h2d_stream));
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
