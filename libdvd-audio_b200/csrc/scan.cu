// scan.cu — exclusive prefix sums over the engine's index tables.
//
// The tables these run on (packets per sector, payload bytes per packet, syncs
// per chunk, access units / frames per segment) are a few MB at most.  A scan
// is one launch: every block sums its tile, publishes the sum, adds up the
// sums its predecessors have published (one predecessor per thread, spinning
// until it is there) and writes its tile's prefix sums.  Blocks number
// themselves with a ticket as they start, so a block only ever waits for
// blocks that are already running.  The last block to finish clears the
// tickets and flags for the next scan; the control words must be zero before
// the first one (the engine clears the buffer when it allocates it).
#include "common.cuh"

#define SCAN_THREADS 512
#define SCAN_ITEMS 4
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)
#define SCAN_MAX_JOBS 4

// up to four tables of the same length are scanned by one set of launches (blockIdx.y = table)
struct ScanBatch {
    const uint32_t *in[SCAN_MAX_JOBS];
    void *out[SCAN_MAX_JOBS];
    uint32_t out64;                 // bit j: table j's prefix sums are 64-bit
    uint64_t *total_copy[SCAN_MAX_JOBS];   // where else to leave table j's total (mapped host memory), or null
};

// the table fits one tile (SINGLE): no block sums needed
template <bool SINGLE>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(ScanBatch b, uint64_t n, const uint64_t *__restrict__ all_sums, uint32_t nblocks)
{
    const uint32_t *__restrict__ in = b.in[blockIdx.y];
    const uint64_t *sums = all_sums + (uint64_t)blockIdx.y * (nblocks + 2);
    // items of one thread are contiguous so the per-thread prefix is a running sum
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t x[SCAN_ITEMS];
    uint64_t v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        x[k] = (base + k < n) ? in[base + k] : 0;
        v += x[k];
    }
    uint64_t total;
    uint64_t ex = block_excl_scan<SCAN_THREADS>(v, &total) + (SINGLE ? 0ull : sums[blockIdx.x]);
    const uint64_t grand = SINGLE ? total : sums[nblocks];
    if ((b.out64 >> blockIdx.y) & 1) {
        uint64_t *out = static_cast<uint64_t *>(b.out[blockIdx.y]);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            if (base + k < n) out[base + k] = ex;
            ex += x[k];
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = grand;
    } else {
        uint32_t *out = static_cast<uint32_t *>(b.out[blockIdx.y]);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            if (base + k < n) out[base + k] = (uint32_t)ex;
            ex += x[k];
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = (uint32_t)grand;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && b.total_copy[blockIdx.y]) *b.total_copy[blockIdx.y] = grand;
}

// control words of one table: [0] next ticket, [1] blocks done, then per block {sum is there, sum}
__device__ __forceinline__ uint64_t *scan_ctl(uint64_t *all, uint32_t table, uint32_t nblocks)
{
    return all + (uint64_t)table * (2 + 2 * (uint64_t)nblocks);
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_chained(ScanBatch b, uint64_t n, uint64_t *__restrict__ ctl_all, uint32_t nblocks)
{
    __shared__ uint32_t s_ticket, s_last;
    uint64_t *ctl = scan_ctl(ctl_all, blockIdx.y, nblocks);
    volatile uint64_t *slots = ctl + 2;
    if (threadIdx.x == 0) s_ticket = atomicAdd(reinterpret_cast<unsigned int *>(ctl), 1u);
    __syncthreads();
    const uint32_t bid = s_ticket;
    const uint32_t *__restrict__ in = b.in[blockIdx.y];
    const uint64_t base = (uint64_t)bid * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t x[SCAN_ITEMS];
    uint64_t v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        x[k] = (base + k < n) ? in[base + k] : 0;
        v += x[k];
    }
    uint64_t total;
    uint64_t ex = block_excl_scan<SCAN_THREADS>(v, &total);
    if (threadIdx.x == 0) {
        slots[2 * bid + 1] = total;
        __threadfence();
        slots[2 * bid] = 1;
    }
    // the sums of all earlier tiles, one per thread (and per round of SCAN_THREADS tiles)
    uint64_t before = 0;
    for (int64_t j = (int64_t)bid - 1 - (int64_t)threadIdx.x; j >= 0; j -= SCAN_THREADS) {
        while (slots[2 * j] == 0) __nanosleep(32);
        __threadfence();
        before += slots[2 * j + 1];
    }
    uint64_t before_all;
    block_excl_scan<SCAN_THREADS>(before, &before_all);
    ex += before_all;
    if ((b.out64 >> blockIdx.y) & 1) {
        uint64_t *out = static_cast<uint64_t *>(b.out[blockIdx.y]);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            if (base + k < n) out[base + k] = ex;
            ex += x[k];
        }
        if (bid == nblocks - 1 && threadIdx.x == 0) out[n] = before_all + total;
    } else {
        uint32_t *out = static_cast<uint32_t *>(b.out[blockIdx.y]);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            if (base + k < n) out[base + k] = (uint32_t)ex;
            ex += x[k];
        }
        if (bid == nblocks - 1 && threadIdx.x == 0) out[n] = (uint32_t)(before_all + total);
    }
    if (bid == nblocks - 1 && threadIdx.x == 0 && b.total_copy[blockIdx.y]) *b.total_copy[blockIdx.y] = before_all + total;
    // whoever finishes last leaves the control words as they were found
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(reinterpret_cast<unsigned int *>(ctl + 1), 1u) == nblocks - 1;
    }
    __syncthreads();
    if (s_last) {
        // (the sums too: the next scan may have another number of tiles and lay its words out differently)
        for (uint32_t i = threadIdx.x; i < 2 * nblocks; i += SCAN_THREADS) slots[i] = 0;
        if (threadIdx.x == 0) { ctl[0] = 0; ctl[1] = 0; }
    }
}

size_t scan_tmp_bytes(uint64_t n)
{
    return (size_t)SCAN_MAX_JOBS * (2 * (size_t)div_up_u32(n ? n : 1, SCAN_TILE) + 2) * sizeof(uint64_t);
}

// exclusive prefix sums of njobs (<= 4) tables of n entries each; out[j] has n + 1 entries
int scan_batch(const uint32_t *const in[], void *const out[], const bool out64[], int njobs, uint64_t n,
               void *tmp, size_t tmp_bytes, cudaStream_t s, uint64_t *const total_copy[])
{
    if (njobs < 1 || njobs > SCAN_MAX_JOBS || tmp_bytes < scan_tmp_bytes(n)) {
        dvdagpu_set_error("scan: bad batch or temporary buffer too small");
        return -1;
    }
    ScanBatch b;
    b.out64 = 0;
    for (int j = 0; j < SCAN_MAX_JOBS; j++) {
        b.in[j] = in[j < njobs ? j : 0]; b.out[j] = out[j < njobs ? j : 0];
        if (j < njobs && out64[j]) b.out64 |= 1u << j;
        b.total_copy[j] = (j < njobs && total_copy) ? total_copy[j] : nullptr;
    }
    uint64_t *sums = (uint64_t *)tmp;
    const uint32_t nblocks = div_up_u32(n ? n : 1, SCAN_TILE);
    if (nblocks == 1) {
        LAUNCH(k_scan_apply<true>, dim3(1, njobs), SCAN_THREADS, 0, s, b, n, sums, nblocks);
    } else {
        LAUNCH(k_scan_chained, dim3(nblocks, njobs), SCAN_THREADS, 0, s, b, n, sums, nblocks);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int scan_u32_to_u64(const uint32_t *in, uint64_t *out, uint64_t n, void *tmp, size_t tmp_bytes, cudaStream_t s)
{
    const uint32_t *i[1] = {in}; void *o[1] = {out}; const bool w[1] = {true};
    return scan_batch(i, o, w, 1, n, tmp, tmp_bytes, s, nullptr);
}
int scan_u32_to_u32(const uint32_t *in, uint32_t *out, uint64_t n, void *tmp, size_t tmp_bytes, cudaStream_t s)
{
    const uint32_t *i[1] = {in}; void *o[1] = {out}; const bool w[1] = {false};
    return scan_batch(i, o, w, 1, n, tmp, tmp_bytes, s, nullptr);
}
