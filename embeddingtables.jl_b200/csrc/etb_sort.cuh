// etb_sort.cuh -- hand-written stable LSD radix sort of (key, int32 value) pairs for index! (K4).
//
// Why not a library sort: the keys of index! are (table slot << row_bits | row), typically 20-27 bits.
// With digits of up to 9 bits, chosen per call, C2's 25-bit keys take 3 passes (9+8+8) where an 8-bit
// library sort takes 4 (or needs the ensemble split into groups).  Stability is what carries the
// reference's occurrence order into the buckets (src/utils.jl:481-511), so every step keeps it.
//
// One pass = three kernels over tiles of 4096 consecutive positions:
//   rs_hist_kernel     per-tile digit counts                          -> tile_hist[digit][tile]
//   rs_scan_kernel     per digit: exclusive scan over tiles (in place) + digit totals
//   rs_scatter_kernel  per tile: stable rank of every element among equal digits (warp-ordered,
//                      ballot-built peer masks + per-warp counters), reorder the tile by digit in shared
//                      memory, then write each digit run to its global position -- consecutive
//                      threads write consecutive addresses inside a run, so stores stay coalesced.
// No inter-CTA waiting anywhere (no decoupled look-back): kernels are ordinary grids.
#pragma once
#include "etb_common.cuh"

namespace etb {

constexpr int kRsThreads = 256, kRsWarps = 8, kRsItems = 16;
constexpr int kRsTile = kRsThreads * kRsItems;  // 4096 positions per CTA
constexpr int kRsMaxBits = 9, kRsMaxBins = 1 << kRsMaxBits;

// digit widths of the passes for `bits` key bits: ceil(bits / 9) passes, widths as even as possible
inline int rs_plan(int bits, int (&width)[8]) {
    const int passes = std::max(1, (bits + kRsMaxBits - 1) / kRsMaxBits);
    for (int p = 0; p < passes; ++p) width[p] = bits / passes + (p < bits % passes ? 1 : 0);
    return passes;
}

inline size_t rs_scratch_bytes(int64_t n) {
    const int64_t ntiles = (n + kRsTile - 1) / kRsTile;
    return (size_t)(ntiles * kRsMaxBins + kRsMaxBins) * sizeof(uint32_t);
}

template <typename KeyT>
__global__ void __launch_bounds__(kRsThreads) rs_hist_kernel(const KeyT* __restrict__ keys, int64_t n, int shift, int nb,
                                                             uint32_t* __restrict__ tile_hist, int ntiles) {
    __shared__ uint32_t cnt[kRsMaxBins];
    for (int d = threadIdx.x; d < nb; d += kRsThreads) cnt[d] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kRsTile;
    KeyT k[kRsItems];
#pragma unroll
    for (int i = 0; i < kRsItems; ++i) {  // all loads first
        const int64_t p = base + i * kRsThreads + threadIdx.x;
        k[i] = p < n ? __ldg(keys + p) : (KeyT)0;
    }
#pragma unroll
    for (int i = 0; i < kRsItems; ++i) {  // plain shared-memory atomics (measured: __match_any_sync aggregation is 3x slower)
        const int64_t p = base + i * kRsThreads + threadIdx.x;
        if (p < n) atomicAdd(&cnt[(uint32_t)(k[i] >> shift) & (uint32_t)(nb - 1)], 1u);
    }
    __syncthreads();
    for (int d = threadIdx.x; d < nb; d += kRsThreads) tile_hist[(size_t)d * ntiles + blockIdx.x] = cnt[d];
}

template <typename KeyT>
struct alignas(sizeof(KeyT) == 4 ? 8 : 16) RsPair {
    KeyT k;
    int32_t v;
};

// exclusive scan of `v` over the block (256 threads); returns the exclusive prefix, total in *total
__device__ __forceinline__ uint32_t rs_block_exscan(uint32_t v, uint32_t* warp_tot /* [8] smem */, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kRsWarps; ++w) {
        const uint32_t t = warp_tot[w];
        if (w < warp) wbase += t;
        tot += t;
    }
    __syncthreads();
    *total = tot;
    return wbase + incl - v;
}

// one block per digit: exclusive scan of its tile counts in place, digit total out
__global__ void __launch_bounds__(kRsThreads) rs_scan_kernel(uint32_t* __restrict__ tile_hist, int ntiles,
                                                             uint32_t* __restrict__ digit_total) {
    __shared__ uint32_t warp_tot[kRsWarps];
    uint32_t* h = tile_hist + (size_t)blockIdx.x * ntiles;
    uint32_t running = 0;
    for (int t0 = 0; t0 < ntiles; t0 += kRsThreads) {
        const int t = t0 + threadIdx.x;
        const uint32_t v = t < ntiles ? h[t] : 0u;
        uint32_t tot;
        const uint32_t ex = rs_block_exscan(v, warp_tot, &tot);
        if (t < ntiles) h[t] = running + ex;
        running += tot;
    }
    if (threadIdx.x == 0) digit_total[blockIdx.x] = running;
}

#ifndef ETB_RS_MIN_BLOCKS
#define ETB_RS_MIN_BLOCKS 4
#endif
template <typename KeyT>
__global__ void __launch_bounds__(kRsThreads, ETB_RS_MIN_BLOCKS) rs_scatter_kernel(const KeyT* __restrict__ kin, const int32_t* __restrict__ vin,
                                                                KeyT* __restrict__ kout, int32_t* __restrict__ vout, int64_t n,
                                                                int shift, int nb, const uint32_t* __restrict__ tile_hist,
                                                                const uint32_t* __restrict__ digit_total, int ntiles) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    RsPair<KeyT>* spair = (RsPair<KeyT>*)rs_smem;                           // [kRsTile] (key, value) staged together:
                                                                           // one shared-memory store / load per element
    uint16_t* wcount = (uint16_t*)(spair + kRsTile);                       // [kRsWarps][kRsMaxBins] counts, then warp prefixes
    uint32_t* tstart = (uint32_t*)(wcount + kRsWarps * kRsMaxBins);        // [kRsMaxBins] tile-local start of each digit
    uint32_t* gbase = tstart + kRsMaxBins;                                 // [kRsMaxBins] global position of that start
    uint32_t* warp_tot = gbase + kRsMaxBins;                               // [kRsWarps]

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t dmask = (uint32_t)(nb - 1);
    const int nbits = 31 - __clz(nb);
    for (int i = threadIdx.x; i < kRsWarps * kRsMaxBins; i += kRsThreads) wcount[i] = 0;
    __syncthreads();

    const int64_t tile_base = (int64_t)blockIdx.x * kRsTile;
    const int tile_n = (int)min((int64_t)kRsTile, n - tile_base);
    // Position order inside the tile: warp w owns positions [w*512, (w+1)*512), 32 consecutive ones per step.
    KeyT k[kRsItems];
    uint32_t dr[kRsItems];  // digit << 16 | rank of the element among equal digits of its warp
    uint16_t* mycount = wcount + warp * kRsMaxBins;
#pragma unroll
    for (int i = 0; i < kRsItems; ++i) {  // all key loads first: the ranking loop's __syncwarp() would pin them in place
        const int q = warp * (32 * kRsItems) + i * 32 + lane;
        k[i] = q < tile_n ? __ldg(kin + tile_base + q) : (KeyT)0;
    }
#pragma unroll
    for (int i = 0; i < kRsItems; ++i) {
        const int q = warp * (32 * kRsItems) + i * 32 + lane;
        const bool valid = q < tile_n;
        const uint32_t d = valid ? ((uint32_t)(k[i] >> shift) & dmask) : 0xffffu;  // invalid lanes form their own group
        // lanes with my digit: one ballot per digit bit (__match_any_sync is several times slower on sm_100)
        unsigned peers = __ballot_sync(0xffffffffu, valid);
        if (!valid) peers = ~peers;
        for (int b = 0; b < nbits; ++b) {
            const unsigned bal = __ballot_sync(0xffffffffu, (d >> b) & 1u);
            peers &= ((d >> b) & 1u) ? bal : ~bal;
        }
        const int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (valid && lane == leader) {
            base = mycount[d];
            mycount[d] = (uint16_t)(base + __popc(peers));
        }
        base = __shfl_sync(0xffffffffu, base, leader);
        dr[i] = (d << 16) | (base + __popc(peers & ((1u << lane) - 1u)));
        __syncwarp();
    }
    __syncthreads();
    // per digit: exclusive prefix over the warps (in place) and the tile's count
    uint32_t my_cnt[kRsMaxBins / kRsThreads];
#pragma unroll
    for (int j = 0; j < kRsMaxBins / kRsThreads; ++j) {
        const int d = threadIdx.x + j * kRsThreads;
        uint32_t run = 0;
        if (d < nb) {
#pragma unroll
            for (int w = 0; w < kRsWarps; ++w) {
                const uint32_t c = wcount[w * kRsMaxBins + d];
                wcount[w * kRsMaxBins + d] = (uint16_t)run;
                run += c;
            }
        }
        my_cnt[j] = run;
    }
    // tile-local digit starts and global digit bases: two exclusive scans over the digits (digit d = tid + j*256,
    // scanned j-major so that the order is d = 0 .. nb-1)
    uint32_t carry_t = 0, carry_g = 0;
#pragma unroll
    for (int j = 0; j < kRsMaxBins / kRsThreads; ++j) {
        const int d = threadIdx.x + j * kRsThreads;
        uint32_t tot;
        const uint32_t ex_t = rs_block_exscan(my_cnt[j], warp_tot, &tot);
        if (d < nb) tstart[d] = carry_t + ex_t;
        carry_t += tot;
        const uint32_t dt = d < nb ? __ldg(digit_total + d) : 0u;
        const uint32_t ex_g = rs_block_exscan(dt, warp_tot, &tot);
        // gbase[d] = (global position of the digit's run for this tile) - (its tile-local start): the
        // write-out then needs one lookup per element, position = gbase[d] + q
        if (d < nb) gbase[d] = carry_g + ex_g + __ldg(tile_hist + (size_t)d * ntiles + blockIdx.x) - tstart[d];
        carry_g += tot;
    }
    __syncthreads();
    // reorder the tile by digit in shared memory (stable).  The values are only needed now and the keys are read
    // again (from L2: this CTA loaded them a moment ago) instead of being held in registers across the ranking and
    // the scans -- 64 registers per thread then hold everything without spilling.  Two halves of 8 loads each.
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        KeyT k2[kRsItems / 2];
        int32_t v[kRsItems / 2];
#pragma unroll
        for (int i = 0; i < kRsItems / 2; ++i) {
            const int q = warp * (32 * kRsItems) + (h * (kRsItems / 2) + i) * 32 + lane;
            k2[i] = q < tile_n ? __ldg(kin + tile_base + q) : (KeyT)0;
            v[i] = q < tile_n ? __ldg(vin + tile_base + q) : 0;
        }
#pragma unroll
        for (int i = 0; i < kRsItems / 2; ++i) {
            const uint32_t dri = dr[h * (kRsItems / 2) + i];
            const uint32_t d = dri >> 16;
            if (d != 0xffffu) {
                const uint32_t q2 = tstart[d] + wcount[warp * kRsMaxBins + d] + (dri & 0xffffu);
                spair[q2] = RsPair<KeyT>{k2[i], v[i]};
            }
        }
    }
    __syncthreads();
    // write the digit runs out: consecutive threads, consecutive addresses inside a run
    for (int q = threadIdx.x; q < tile_n; q += kRsThreads) {
        const RsPair<KeyT> pr = spair[q];
        const uint32_t d = (uint32_t)(pr.k >> shift) & dmask;
        const uint32_t g = gbase[d] + (uint32_t)q;  // modular uint32 arithmetic: gbase may have wrapped below zero
        kout[g] = pr.k;
        vout[g] = pr.v;
    }
}

template <typename KeyT>
constexpr size_t rs_scatter_smem() {
    return (size_t)kRsTile * sizeof(RsPair<KeyT>) + (size_t)kRsWarps * kRsMaxBins * sizeof(uint16_t) +
           2 * kRsMaxBins * sizeof(uint32_t) + kRsWarps * sizeof(uint32_t);
}

// Sorts n pairs by the low `bits` bits of the key, stable.  keys/vals are ping-pong buffers; on return
// *result is the index (0/1) of the buffer holding the sorted pairs.  `scratch` has rs_scratch_bytes(n).
template <typename KeyT>
int32_t radix_sort_pairs(KeyT* keys[2], int32_t* vals[2], int64_t n, int bits, uint32_t* scratch, cudaStream_t stream,
                         int* result) {
    ETB_CUDA(cudaFuncSetAttribute(rs_scatter_kernel<KeyT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)rs_scatter_smem<KeyT>()));  // > 48 KB of dynamic shared memory
    ETB_CUDA(cudaFuncSetAttribute(rs_scatter_kernel<KeyT>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int width[8];
    const int passes = rs_plan(bits, width);
    const int ntiles = (int)((n + kRsTile - 1) / kRsTile);
    uint32_t* tile_hist = scratch;
    uint32_t* digit_total = scratch + (size_t)ntiles * kRsMaxBins;
    int cur = 0, shift = 0;
    for (int p = 0; p < passes; ++p) {
        const int nb = 1 << width[p];
        rs_hist_kernel<KeyT><<<ntiles, kRsThreads, 0, stream>>>(keys[cur], n, shift, nb, tile_hist, ntiles);
        ETB_LAUNCHED();
        rs_scan_kernel<<<nb, kRsThreads, 0, stream>>>(tile_hist, ntiles, digit_total);
        ETB_LAUNCHED();
        rs_scatter_kernel<KeyT><<<ntiles, kRsThreads, rs_scatter_smem<KeyT>(), stream>>>(
            keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], n, shift, nb, tile_hist, digit_total, ntiles);
        ETB_LAUNCHED();
        cur ^= 1;
        shift += width[p];
    }
    *result = cur;
    return ETB_OK;
}

}  // namespace etb
