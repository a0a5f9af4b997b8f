// scan.cu — exclusive prefix sums over the engine's index tables.
//
// The tables these run on (packets per sector, payload bytes per packet, syncs
// per chunk, access units / frames per segment) are a few MB at most, so a plain
// three-pass reduce / scan-of-sums / rescan is enough: every pass is coalesced
// and the middle pass is one block.
#include "common.cuh"

#define SCAN_THREADS 512
#define SCAN_ITEMS 4
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const uint32_t *__restrict__ in, uint64_t n, uint64_t *__restrict__ sums)
{
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
    uint64_t v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        const uint64_t i = base + (uint64_t)k * SCAN_THREADS + threadIdx.x;
        if (i < n) v += in[i];
    }
    uint64_t total;
    block_excl_scan<SCAN_THREADS>(v, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// one block: exclusive scan of the block sums in place, total behind them
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_sums(uint64_t *sums, uint32_t nblocks)
{
    uint64_t carry = 0;
    for (uint32_t base = 0; base < nblocks; base += SCAN_THREADS) {
        const uint32_t i = base + threadIdx.x;
        const uint64_t v = i < nblocks ? sums[i] : 0;
        uint64_t total;
        const uint64_t ex = block_excl_scan<SCAN_THREADS>(v, &total);
        if (i < nblocks) sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) sums[nblocks] = carry;
}

template <typename OutT>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const uint32_t *__restrict__ in, uint64_t n, const uint64_t *__restrict__ sums, uint32_t nblocks, OutT *__restrict__ out)
{
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
    uint64_t ex = block_excl_scan<SCAN_THREADS>(v, &total) + sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < n) out[base + k] = (OutT)ex;
        ex += x[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = (OutT)sums[nblocks];
}

size_t scan_tmp_bytes(uint64_t n)
{
    return (size_t)(div_up_u32(n ? n : 1, SCAN_TILE) + 2) * sizeof(uint64_t);
}

template <typename OutT>
static int scan_impl(const uint32_t *in, OutT *out, uint64_t n, void *tmp, size_t tmp_bytes, cudaStream_t s)
{
    if (tmp_bytes < scan_tmp_bytes(n)) {
        dvdagpu_set_error("scan: temporary buffer too small");
        return -1;
    }
    uint64_t *sums = (uint64_t *)tmp;
    const uint32_t nblocks = div_up_u32(n ? n : 1, SCAN_TILE);
    LAUNCH(k_scan_reduce, nblocks, SCAN_THREADS, 0, s, in, n, sums);
    LAUNCH(k_scan_sums, 1, SCAN_THREADS, 0, s, sums, nblocks);
    LAUNCH(k_scan_apply<OutT>, nblocks, SCAN_THREADS, 0, s, in, n, sums, nblocks, out);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int scan_u32_to_u64(const uint32_t *in, uint64_t *out, uint64_t n, void *tmp, size_t tmp_bytes, cudaStream_t s)
{
    return scan_impl<uint64_t>(in, out, n, tmp, tmp_bytes, s);
}
int scan_u32_to_u32(const uint32_t *in, uint32_t *out, uint64_t n, void *tmp, size_t tmp_bytes, cudaStream_t s)
{
    return scan_impl<uint32_t>(in, out, n, tmp, tmp_bytes, s);
}
