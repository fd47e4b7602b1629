// etb_index.cuh -- index! (K4) as a per-table segmented, stable LSD radix sort.
//
// What index! needs (reference src/utils.jl:131-314): for every table of the ensemble, the occurrences of each
// distinct row grouped together, members in occurrence order.  Every occurrence p of table t is the pair
// (row-1, delta column of p); the pairs of ONE table are sorted by row, stably, inside that table's segment of
// the concatenated arrays -- tables are already partitioned, so the table slot is not part of the sort key
// (round 1 sorted slot << row_bits | row: 25 key bits = 3 passes on C2 where 20 bits = 2 passes of 10 suffice).
//
// One pass = three ordinary kernels over tiles of 4096 consecutive positions of one table (no inter-CTA waiting;
// two designs with waiting -- per-tile decoupled look-back, and chunks of tiles walked by one CTA -- were built
// and measured slower on B200, see profiles/README.md):
//   ix_hist_kernel     per-tile digit counts                                   -> tile_hist[digit][tile]
//   ix_scan_kernel     per digit: exclusive scan over the tiles of each table (in place) + the table's digit totals
//   ix_scatter_kernel  per tile: stable rank of every element among equal digits (warp-ordered), the tile
//                      reordered by digit in shared memory, every digit run written with consecutive threads on
//                      consecutive addresses.  The first pass reads the caller's index arrays themselves (there is
//                      no key/value formation pass).
// Bucket heads -> records: ix_count_heads_kernel (per tile), ix_scan_counts_kernel (one block), ix_write_records_kernel:
// one etb_bucket_record per bucket, numbered over all tables in (table, row) order.
#pragma once
#include <type_traits>

#include "etb_layout.cuh"

namespace etb {

constexpr int kIxMaxItems = 96;            // tables per launch (descriptors travel in kernel parameters)
constexpr int kIxMaxBits = 11;             // widest digit
constexpr int kIxMaxBins = 1 << kIxMaxBits;
constexpr int kIxMaxPasses = 6;            // 64-bit rows
constexpr int kIxItems = 16;               // positions per thread and tile
constexpr int kIxThreads = 256, kIxWarps = kIxThreads / 32;  // the record kernels; the sort kernels take THREADS
constexpr int kIxTile = kIxThreads * kIxItems;               // 4096 positions per CTA of the record kernels

struct IxItem {  // 48 bytes
    const void* idx;
    uint64_t ld_idx;         // elements between index columns (1 for a vector of indices)
    uint32_t n;              // occurrences (< 2^30)
    uint32_t bag;            // divisor of the flat position (1 for a vector of indices)
    uint32_t magic, mshift;  // delta column of flat position p: (p * magic) >> (31 + mshift)
    uint32_t seg_start;      // first position of this table's segment in the concatenated arrays
    uint32_t tile_start;     // first sort tile of this table in the launch's numbering (tile = 16 * THREADS positions)
    uint32_t rec_tile_start; // first record tile (4096 positions) of this table in the launch's numbering
    uint32_t pad;
};

struct IxParams {
    IxItem item[kIxMaxItems];
    const void* kin;        // keys in  (passes after the first; the record kernels: the sorted keys)
    const int32_t* vin;
    void* kout;
    int32_t* vout;
    uint32_t* tile_hist;    // [ntiles][nb] digit counts of this pass, then exclusive prefixes within each table
    uint32_t* digit_total;  // [n_items][nb]
    BucketRec* recs;
    int64_t* nnz;
    uint32_t* rec_counts;   // [tiles of the call + 1] bucket heads per tile, then exclusive offsets (rec_inline: counts)
    int32_t rec_inline;     // few tiles: ix_write_records_kernel adds up the counts before its tile itself (no scan kernel)
    int32_t rec_total;      // record tiles of the whole call
    int32_t n_items, ntiles, rec_tiles, pass, slot0, row_bits;
    uint32_t tile0;         // record tiles of the earlier launches of this call
    uint8_t width[8], shift[8];
};

// digit widths: ceil(bits / 11) passes, as even as possible (20 bits = 10 + 10, 24 = 8 + 8 + 8, 17 = 9 + 8)
inline int ix_plan(int bits, uint8_t (&width)[8], uint8_t (&shift)[8]) {
    const int passes = std::max(1, (bits + kIxMaxBits - 1) / kIxMaxBits);
    int sh = 0;
    for (int p = 0; p < passes; ++p) {
        width[p] = (uint8_t)(bits / passes + (p < bits % passes ? 1 : 0));
        if (width[p] == 0) width[p] = 1;
        shift[p] = (uint8_t)sh;
        sh += width[p];
    }
    return passes;
}

// exact p / d for p < 2^31: (p * magic) >> (31 + mshift)
inline void ix_magic(uint32_t d, uint32_t* magic, uint32_t* mshift) {
    uint32_t L = 0;
    while ((1ull << L) < d) ++L;
    *mshift = L;
    *magic = (uint32_t)(((1ull << (31 + L)) + d - 1) / d);
}

#ifdef __CUDACC__
// exclusive scans of `a` and `b` over the block with one pair of barriers; totals in *ta, *tb
template <int THREADS>
__device__ __forceinline__ void ix_block_exscan2(uint32_t& a, uint32_t& b, uint32_t* warp_tot /* [2][THREADS/32] smem */,
                                                 uint32_t* ta, uint32_t* tb) {
    constexpr int kWarps = THREADS / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t ia = a, ib = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t x = __shfl_up_sync(0xffffffffu, ia, o), y = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= o) { ia += x; ib += y; }
    }
    if (lane == 31) { warp_tot[warp] = ia; warp_tot[kWarps + warp] = ib; }
    __syncthreads();
    uint32_t wa = 0, wb = 0, sa = 0, sb = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
        const uint32_t x = warp_tot[w], y = warp_tot[kWarps + w];
        if (w < warp) { wa += x; wb += y; }
        sa += x; sb += y;
    }
    __syncthreads();
    *ta = sa; *tb = sb;
    a = wa + ia - a;
    b = wb + ib - b;
}

// table of launch-local tile `tile`: the last item whose tile_start <= tile (items without tiles share their
// tile_start with the next item and are skipped that way)
template <bool REC>
__device__ __forceinline__ int ix_find_item(const IxParams& P, uint32_t tile) {
    int lo = 0, hi = P.n_items - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if ((REC ? P.item[mid].rec_tile_start : P.item[mid].tile_start) <= tile) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

// Flat position p of a table's occurrences -> element of the caller's index array and delta column: column-major
// traversal (reference `columns`, src/utils.jl:73-81); the delta column of position p is p / bag.  Branch-free,
// so that the loads of a tile are all issued before the first one is used.
__device__ __forceinline__ size_t ix_src_pos(const IxItem& it, uint32_t p, uint32_t* col_out) {
    const uint32_t col = (uint32_t)(((uint64_t)p * it.magic) >> (31 + it.mshift));
    *col_out = col;
    return (size_t)col * it.ld_idx + (p - col * it.bag);
}

// key (row - 1) of position p of the table's segment: from the index array in the first pass (SrcT = its element
// type), else (SrcT = void) from the previous pass's output.  p must be a valid position (callers clamp).
template <typename KeyT, typename SrcT>
__device__ __forceinline__ KeyT ix_load_key(const IxItem& it, const KeyT* kin, uint32_t p) {
    if constexpr (std::is_void<SrcT>::value) {
        return __ldg(kin + p);
    } else {
        uint32_t col;
        return (KeyT)((int64_t)__ldg((const SrcT*)it.idx + ix_src_pos(it, p, &col)) - 1);
    }
}

// ------------------------------------------------------------------------------------ per-tile digit counts
template <typename KeyT, typename SrcT, int THREADS>
__global__ void __launch_bounds__(THREADS) ix_hist_kernel(const __grid_constant__ IxParams P) {
    pdl_begin();
    __shared__ uint32_t cnt[kIxMaxBins];
    constexpr uint32_t kTile = THREADS * kIxItems;
    const int nb = 1 << P.width[P.pass], shift = P.shift[P.pass];
    const uint32_t dmask = (uint32_t)nb - 1u;
    const int item = ix_find_item<false>(P, blockIdx.x);
    const IxItem& it = P.item[item];
    const uint32_t base = (blockIdx.x - it.tile_start) * kTile;
    const KeyT* kin = (const KeyT*)P.kin + it.seg_start;
    for (int j = threadIdx.x; j < nb; j += THREADS) cnt[j] = 0;
    __syncthreads();
    KeyT k[kIxItems];
#pragma unroll
    for (int i = 0; i < kIxItems; ++i)  // all loads first (positions past the end repeat the last one)
        k[i] = ix_load_key<KeyT, SrcT>(it, kin, min(base + i * THREADS + threadIdx.x, it.n - 1));
#pragma unroll
    for (int i = 0; i < kIxItems; ++i)
        if (base + i * THREADS + threadIdx.x < it.n) atomicAdd(&cnt[(uint32_t)(k[i] >> shift) & dmask], 1u);
    __syncthreads();
    uint32_t* g = P.tile_hist + (size_t)blockIdx.x * nb;  // one contiguous row per tile
    for (int d = threadIdx.x; d < nb; d += THREADS) g[d] = cnt[d];
}

// one thread per (table, digit): exclusive scan of the digit's counts over the table's tiles (in place; consecutive
// threads = consecutive digits, so every access is coalesced) and the table's total
__global__ void __launch_bounds__(256) ix_scan_kernel(const __grid_constant__ IxParams P, int tile_size) {
    pdl_begin();
    const int nb = 1 << P.width[P.pass];
    const int d = blockIdx.x * 256 + threadIdx.x;
    if (d >= nb) return;
    const IxItem& it = P.item[blockIdx.y];
    const int nt = (int)((it.n + tile_size - 1) / tile_size);
    uint32_t* h = P.tile_hist + (size_t)it.tile_start * nb + d;
    uint32_t running = 0;
    int t = 0;
    for (; t + 8 <= nt; t += 8) {  // 8 loads in flight
        uint32_t v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = h[(size_t)(t + u) * nb];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            h[(size_t)(t + u) * nb] = running;
            running += v[u];
        }
    }
    for (; t < nt; ++t) {
        const uint32_t v = h[(size_t)t * nb];
        h[(size_t)t * nb] = running;
        running += v;
    }
    P.digit_total[(size_t)blockIdx.y * nb + d] = running;
}

// The same with one WARP per (table, digit), for calls with few (table, digit) pairs (one table of C3: 256 threads
// walking 128 tiles in a chain of dependent round trips, 13 us of index!'s 80): lane l takes tiles l, l + 32, ...; 32
// tiles per shuffle scan.  The accesses are not coalesced (the counts come from L2, written by ix_hist_kernel just
// before), so the thread-per-digit kernel stays for ensembles (C2: 0.350 vs 0.368 ms).
__global__ void __launch_bounds__(256) ix_scan_warp_kernel(const __grid_constant__ IxParams P, int tile_size) {
    pdl_begin();
    const int nb = 1 << P.width[P.pass];
    const int lane = threadIdx.x & 31;
    const int d = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (d >= nb) return;
    const IxItem& it = P.item[blockIdx.y];
    const int nt = (int)((it.n + tile_size - 1) / tile_size);
    uint32_t* h = P.tile_hist + (size_t)it.tile_start * nb + d;
    uint32_t running = 0;
    for (int t0 = 0; t0 < nt; t0 += 64) {  // two rounds of 32 tiles in flight
        const int ta = t0 + lane, tb = t0 + 32 + lane;
        const uint32_t va = ta < nt ? h[(size_t)ta * nb] : 0u;
        const uint32_t vb = tb < nt ? h[(size_t)tb * nb] : 0u;
        uint32_t ia = va, ib = vb;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t na = __shfl_up_sync(0xffffffffu, ia, off), nb2 = __shfl_up_sync(0xffffffffu, ib, off);
            if (lane >= off) {
                ia += na;
                ib += nb2;
            }
        }
        const uint32_t ta_total = __shfl_sync(0xffffffffu, ia, 31), tb_total = __shfl_sync(0xffffffffu, ib, 31);
        if (ta < nt) h[(size_t)ta * nb] = running + ia - va;
        if (tb < nt) h[(size_t)tb * nb] = running + ta_total + ib - vb;
        running += ta_total + tb_total;
    }
    if (lane == 0) P.digit_total[(size_t)blockIdx.y * nb + d] = running;
}

// ------------------------------------------------------------------------------------ scatter
template <typename KeyT>
struct alignas(sizeof(KeyT) == 4 ? 8 : 16) IxPair {
    KeyT k;
    int32_t v;
};

template <typename KeyT, int THREADS>
constexpr size_t ix_scatter_smem(int nb) {
    return (size_t)THREADS * kIxItems * sizeof(IxPair<KeyT>)  // the tile, reordered by digit
           + (size_t)(THREADS / 32) * nb * sizeof(uint16_t)   // per-warp digit counters, then warp prefixes
           + (size_t)nb * sizeof(uint32_t)                    // tile-local digit starts, then write-out bases
           + 2 * (THREADS / 32) * sizeof(uint32_t);           // scan scratch
}

// Two or more lanes of a 32-key row share a digit: their shared-memory atomics were served in an unspecified order.
// Peer masks from one ballot per digit bit (bits above the digit's width are zero in every lane and change nothing)
// give each lane its place in lane order: rank = (count after the row) - (peers) + (peers in lower lanes).
// Out of line: 16 rows share one copy of the unrolled ballots.
static __device__ __noinline__ uint32_t ix_rank_shared_digit(uint32_t d, bool valid, uint32_t now) {
    const int lane = threadIdx.x & 31;
    unsigned peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int b = 0; b < kIxMaxBits; ++b) {
        const bool bit = (d >> b) & 1u;
        const unsigned bal = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? bal : ~bal;
    }
    return now - __popc(peers) + __popc(peers & ((1u << lane) - 1u));
}

// Stable rank of every element of a tile among the equal digits of its warp.  Warp w owns positions
// [w * 32 * kIxItems, ...), 32 consecutive ones per step (element i of lane l = position w*512 + i*32 + l), so ranks
// follow position order.  wcount: [warps][nb] 16-bit counters, zero on entry; on exit the per-warp digit counts.
// RANK = 0: peer masks from one ballot per digit bit for every row, counters updated by the leader of each peer
// group.  RANK = 1: shared-memory atomics on the counters; ballots only for rows in which two lanes share a digit.
// (Both are stable; which one is faster is a measurement, ETB_IX_RANK selects.)  rk: two 16-bit ranks per register.
template <typename KeyT, int RANK>
__device__ __forceinline__ void ix_rank_rows(const KeyT (&k)[kIxItems], uint32_t (&rk)[kIxItems / 2], uint16_t* wcount, int nb,
                                             int nbits, int shift, int tile_n) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t dmask = (uint32_t)nb - 1u;
    if constexpr (RANK == 1) {
        uint32_t* cw = (uint32_t*)(wcount + (size_t)warp * nb);  // two 16-bit counters per word
#pragma unroll
        for (int i = 0; i < kIxItems; ++i) {
            const int q = warp * (32 * kIxItems) + i * 32 + lane;
            const bool valid = q < tile_n;
            const uint32_t d = (uint32_t)(k[i] >> shift) & dmask;
            const int half = (int)(d & 1u) * 16;
            uint32_t old = 0, now = 0;
            if (valid) old = (atomicAdd(&cw[d >> 1], 1u << half) >> half) & 0xffffu;
            __syncwarp();
            if (valid) now = (cw[d >> 1] >> half) & 0xffffu;
            // A lane whose digit is unique in this row got the exact count of earlier equal digits.  Lanes that
            // share a digit were served in an unspecified order; at least one of them sees now - old != 1.
            uint32_t rank = old;
            if (__any_sync(0xffffffffu, valid && now - old != 1u)) rank = ix_rank_shared_digit(d, valid, now);
            if (i & 1) rk[i >> 1] |= rank << 16;
            else rk[i >> 1] = rank;
            __syncwarp();
        }
    } else {
        uint16_t* mycount = wcount + (size_t)warp * nb;
#pragma unroll
        for (int i = 0; i < kIxItems; ++i) {
            const int q = warp * (32 * kIxItems) + i * 32 + lane;
            const bool valid = q < tile_n;
            const uint32_t d = (uint32_t)(k[i] >> shift) & dmask;
            unsigned peers = __ballot_sync(0xffffffffu, valid);
            if (!valid) peers = ~peers;  // invalid lanes form their own group
#pragma unroll 1
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
            const uint32_t rank = base + __popc(peers & ((1u << lane) - 1u));
            if (i & 1) rk[i >> 1] |= rank << 16;
            else rk[i >> 1] = rank;
            __syncwarp();
        }
    }
}

template <typename KeyT, typename SrcT, int RANK, int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 256 ? 4 : 2) ix_scatter_kernel(const __grid_constant__ IxParams P) {
    pdl_begin();
    constexpr bool kFirst = !std::is_void<SrcT>::value;
    constexpr int kIxThreads = THREADS, kIxWarps = THREADS / 32, kIxTile = THREADS * kIxItems;  // shadow the record kernels' constants
    constexpr int kIxMaxBpt = kIxMaxBins / THREADS;  // digits per thread in the per-digit steps
    extern __shared__ __align__(16) unsigned char ix_smem[];
    const int nbits = P.width[P.pass], shift = P.shift[P.pass];
    const int nb = 1 << nbits;
    const uint32_t dmask = (uint32_t)nb - 1u;
    const int bpt = max(1, nb / kIxThreads);  // digits per thread: d = tid * bpt + j
    IxPair<KeyT>* spair = (IxPair<KeyT>*)ix_smem;
    uint16_t* wcount = (uint16_t*)(spair + kIxTile);                 // [kIxWarps][nb]
    uint32_t* tstart = (uint32_t*)(wcount + (size_t)kIxWarps * nb);  // [nb] tile-local start of each digit; after the
    uint32_t* gbase = tstart;                                        // reorder: output position of that start, minus it
    uint32_t* warp_tot = tstart + nb;                                // [2 * kIxWarps]

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kIxWarps * nb / 2; i += kIxThreads) ((uint32_t*)wcount)[i] = 0;
    const int item = ix_find_item<false>(P, blockIdx.x);
    const IxItem& it = P.item[item];
    const uint32_t tbase = (blockIdx.x - it.tile_start) * kIxTile;  // first position of the tile in the table's segment
    const int tile_n = (int)min((uint32_t)kIxTile, it.n - tbase);
    const KeyT* kin = (const KeyT*)P.kin + it.seg_start;
    const int32_t* vin = P.vin + it.seg_start;
    // ---- keys of the tile, all loads first.  Warp w owns positions [w * 512, (w + 1) * 512), 32 consecutive ones per
    // step, so ranks follow position order.
    KeyT k[kIxItems];
#pragma unroll
    for (int i = 0; i < kIxItems; ++i) {
        const int q = warp * (32 * kIxItems) + i * 32 + lane;
        k[i] = ix_load_key<KeyT, SrcT>(it, kin, tbase + min(q, tile_n - 1));
    }
    __syncthreads();
    // ---- stable rank of every element among the equal digits of its warp
    uint32_t rk[kIxItems / 2];  // two 16-bit ranks per register
    ix_rank_rows<KeyT, RANK>(k, rk, wcount, nb, nbits, shift, tile_n);
    __syncthreads();
    // ---- per digit: exclusive prefix over the warps (in place), the tile's count; tile-local digit starts and the
    // table's digit bases (two block scans sharing their barriers).  The write-out bases stay in registers until the
    // reorder has read the tile-local starts, whose array they then take over.
    uint32_t gb[kIxMaxBpt];
    {
        uint32_t cnt[kIxMaxBpt], tsum = 0, gsum = 0;
        const uint32_t* dtot = P.digit_total + (size_t)item * nb;
        const uint32_t* th = P.tile_hist + (size_t)blockIdx.x * nb;
#pragma unroll
        for (int j = 0; j < kIxMaxBpt; ++j) {
            const int d = threadIdx.x * bpt + j;
            uint32_t run = 0;
            if (j < bpt && d < nb) {
#pragma unroll
                for (int w = 0; w < kIxWarps; ++w) {
                    const uint32_t c = wcount[w * nb + d];
                    wcount[w * nb + d] = (uint16_t)run;
                    run += c;
                }
                gsum += __ldg(dtot + d);
            }
            cnt[j] = run;
            tsum += run;
        }
        uint32_t ta, tb;
        ix_block_exscan2<THREADS>(tsum, gsum, warp_tot, &ta, &tb);
#pragma unroll
        for (int j = 0; j < kIxMaxBpt; ++j) {
            const int d = threadIdx.x * bpt + j;
            gb[j] = 0;
            if (j < bpt && d < nb) {
                tstart[d] = tsum;
                // write-out position of tile element q of digit d = gbase[d] + q
                gb[j] = it.seg_start + gsum + __ldg(th + d) - tsum;
                tsum += cnt[j];
                gsum += __ldg(dtot + d);
            }
        }
    }
    __syncthreads();
    // ---- reorder the tile by digit in shared memory (stable); values: the delta column of the position in the first
    // pass, the previous pass's output otherwise (two halves of 8 loads)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        int32_t v[kIxItems / 2];
#pragma unroll
        for (int i = 0; i < kIxItems / 2; ++i) {
            const int q = warp * (32 * kIxItems) + (hh * (kIxItems / 2) + i) * 32 + lane;
            if constexpr (kFirst) {
                uint32_t col;
                ix_src_pos(it, tbase + q, &col);
                v[i] = (int32_t)col;
            } else {
                v[i] = __ldg(vin + tbase + min(q, tile_n - 1));
            }
        }
#pragma unroll
        for (int i = 0; i < kIxItems / 2; ++i) {
            const int ii = hh * (kIxItems / 2) + i;
            const int q = warp * (32 * kIxItems) + ii * 32 + lane;
            if (q < tile_n) {
                const uint32_t d = (uint32_t)(k[ii] >> shift) & dmask;
                const uint32_t rank = (rk[ii >> 1] >> ((ii & 1) * 16)) & 0xffffu;
                spair[tstart[d] + wcount[warp * nb + d] + rank] = IxPair<KeyT>{k[ii], v[i]};
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kIxMaxBpt; ++j) {
        const int d = threadIdx.x * bpt + j;
        if (j < bpt && d < nb) gbase[d] = gb[j];
    }
    __syncthreads();
    // ---- write the digit runs out: consecutive threads, consecutive addresses inside a run
    KeyT* kout = (KeyT*)P.kout;
    for (int q = threadIdx.x; q < tile_n; q += kIxThreads) {
        const IxPair<KeyT> pr = spair[q];
        const uint32_t d = (uint32_t)(pr.k >> shift) & dmask;
        const uint32_t g = gbase[d] + (uint32_t)q;  // modular uint32 arithmetic: gbase may have wrapped below zero
        kout[g] = pr.k;
        P.vout[g] = pr.v;
    }
}

// ------------------------------------------------------------------------------------ small tables
// Every table of the launch has at most one tile (4096 occurrences): one CTA sorts a whole table in shared memory --
// all passes back to back, no histogram / scan kernels, nothing but the sorted pairs and the head count goes to
// global memory.  C1 (26 tables x 2048 indices) is 2 launches this way instead of 9.
template <typename KeyT>
constexpr size_t ix_small_smem(int nb) {
    return (size_t)kIxTile * sizeof(IxPair<KeyT>) + (size_t)kIxWarps * nb * sizeof(uint16_t) + (size_t)nb * sizeof(uint32_t) +
           2 * kIxWarps * sizeof(uint32_t);
}

template <typename KeyT, typename SrcT, int RANK>
__global__ void __launch_bounds__(kIxThreads) ix_small_kernel(const __grid_constant__ IxParams P, int npasses) {
    pdl_begin();
    constexpr int kMaxBpt = kIxMaxBins / kIxThreads;
    extern __shared__ __align__(16) unsigned char ix_smem[];
    const int nb_max = 1 << P.width[0];
    IxPair<KeyT>* spair = (IxPair<KeyT>*)ix_smem;
    uint16_t* wcount = (uint16_t*)(spair + kIxTile);                     // [kIxWarps][nb]
    uint32_t* tstart = (uint32_t*)(wcount + (size_t)kIxWarps * nb_max);  // [nb]
    uint32_t* warp_tot = tstart + nb_max;                                // [2 * kIxWarps]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const IxItem& it = P.item[blockIdx.x];
    const int tile_n = (int)it.n;
    if (tile_n == 0) return;
    KeyT k[kIxItems];
    int32_t v[kIxItems];
#pragma unroll
    for (int i = 0; i < kIxItems; ++i) {  // element i of lane l = position w*512 + i*32 + l
        const int q = min(warp * (32 * kIxItems) + i * 32 + lane, tile_n - 1);
        uint32_t col;
        k[i] = (KeyT)((int64_t)__ldg((const SrcT*)it.idx + ix_src_pos(it, (uint32_t)q, &col)) - 1);
        v[i] = (int32_t)col;
    }
    for (int pass = 0; pass < npasses; ++pass) {
        const int nbits = P.width[pass], shift = P.shift[pass];
        const int nb = 1 << nbits;
        const uint32_t dmask = (uint32_t)nb - 1u;
        const int bpt = max(1, nb / kIxThreads);
        for (int i = threadIdx.x; i < kIxWarps * nb / 2; i += kIxThreads) ((uint32_t*)wcount)[i] = 0;
        __syncthreads();
        uint32_t rk[kIxItems / 2];
        ix_rank_rows<KeyT, RANK>(k, rk, wcount, nb, nbits, shift, tile_n);
        __syncthreads();
        {   // per digit: exclusive prefix over the warps (in place) and the digit's start
            uint32_t cnt[kMaxBpt], tsum = 0, dummy = 0;
#pragma unroll
            for (int j = 0; j < kMaxBpt; ++j) {
                const int d = threadIdx.x * bpt + j;
                uint32_t run = 0;
                if (j < bpt && d < nb) {
#pragma unroll
                    for (int w = 0; w < kIxWarps; ++w) {
                        const uint32_t c = wcount[w * nb + d];
                        wcount[w * nb + d] = (uint16_t)run;
                        run += c;
                    }
                }
                cnt[j] = run;
                tsum += run;
            }
            uint32_t ta, tb;
            ix_block_exscan2<kIxThreads>(tsum, dummy, warp_tot, &ta, &tb);
#pragma unroll
            for (int j = 0; j < kMaxBpt; ++j) {
                const int d = threadIdx.x * bpt + j;
                if (j < bpt && d < nb) {
                    tstart[d] = tsum;
                    tsum += cnt[j];
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kIxItems; ++i) {
            const int q = warp * (32 * kIxItems) + i * 32 + lane;
            if (q < tile_n) {
                const uint32_t d = (uint32_t)(k[i] >> shift) & dmask;
                const uint32_t rank = (rk[i >> 1] >> ((i & 1) * 16)) & 0xffffu;
                spair[tstart[d] + wcount[warp * nb + d] + rank] = IxPair<KeyT>{k[i], v[i]};
            }
        }
        __syncthreads();
        if (pass + 1 < npasses) {  // the next pass ranks the elements in their new order
#pragma unroll
            for (int i = 0; i < kIxItems; ++i) {
                const IxPair<KeyT> pr = spair[min(warp * (32 * kIxItems) + i * 32 + lane, tile_n - 1)];
                k[i] = pr.k;
                v[i] = pr.v;
            }
        }
    }
    // sorted pairs out (coalesced) and the number of bucket heads of this table = its record tile
    KeyT* kout = (KeyT*)P.kout + it.seg_start;
    int32_t* vout = P.vout + it.seg_start;
    uint32_t heads = 0;
    for (int q = threadIdx.x; q < tile_n; q += kIxThreads) {
        const IxPair<KeyT> pr = spair[q];
        kout[q] = pr.k;
        vout[q] = pr.v;
        heads += (q == 0 || spair[q - 1].k != pr.k) ? 1u : 0u;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) heads += __shfl_xor_sync(0xffffffffu, heads, o);
    __syncthreads();
    if (lane == 0) warp_tot[warp] = heads;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int w = 0; w < kIxWarps; ++w) t += warp_tot[w];
        P.rec_counts[P.tile0 + it.rec_tile_start] = t;
    }
}

// ------------------------------------------------------------------------------------ bucket records
// Position p of a table's sorted segment is a bucket head when it is the segment's first position or its key
// differs from the previous one.  Positions are striped over the block (p = base + i*256 + tid) so every load is
// coalesced; ranks follow position order (i-major, then warp, then lane) via warp ballots.
template <typename KeyT>
__device__ __forceinline__ uint32_t ix_head_flags(const KeyT* __restrict__ keys, uint32_t base, uint32_t n, KeyT (&k)[kIxItems]) {
    KeyT prev[kIxItems];
#pragma unroll
    for (int i = 0; i < kIxItems; ++i) {  // all loads first, unconditionally (positions past the end repeat the last one)
        const uint32_t p = min(base + i * kIxThreads + threadIdx.x, n - 1);
        k[i] = __ldg(keys + p);
        prev[i] = __ldg(keys + (p ? p - 1 : 0));
    }
    uint32_t flags = 0;
#pragma unroll
    for (int i = 0; i < kIxItems; ++i) {
        const uint32_t p = base + i * kIxThreads + threadIdx.x;
        if (p < n && (p == 0 || prev[i] != k[i])) flags |= 1u << i;
    }
    return flags;
}

template <typename KeyT>
__global__ void __launch_bounds__(kIxThreads) ix_count_heads_kernel(const __grid_constant__ IxParams P) {
    pdl_begin();
    __shared__ uint32_t warp_sums[kIxWarps];
    const int item = ix_find_item<true>(P, blockIdx.x);
    const IxItem& it = P.item[item];
    KeyT k[kIxItems];
    uint32_t c = __popc(ix_head_flags((const KeyT*)P.kin + it.seg_start, (blockIdx.x - it.rec_tile_start) * kIxTile, it.n, k));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int w = 0; w < kIxWarps; ++w) t += warp_sums[w];
        P.rec_counts[P.tile0 + blockIdx.x] = t;
    }
}

// exclusive scan of the tile counts of the whole call in place (one block of 1024 threads), total -> nnz
__global__ void __launch_bounds__(1024) ix_scan_counts_kernel(uint32_t* __restrict__ counts, int ntiles, int64_t* __restrict__ nnz) {
    pdl_begin();
    __shared__ uint32_t warp_tot[32];
    const int per = (ntiles + 1023) / 1024;
    const int lo = min((int)threadIdx.x * per, ntiles), hi = min(lo + per, ntiles);
    uint32_t sum = 0;
    for (int i = lo; i < hi; ++i) sum += counts[i];
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += v;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t w = warp_tot[threadIdx.x], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, wi, o);
            if (threadIdx.x >= o) wi += v;
        }
        warp_tot[threadIdx.x] = wi - w;  // exclusive warp offsets
        if (threadIdx.x == 31) *nnz = (int64_t)wi;
    }
    __syncthreads();
    uint32_t run = warp_tot[threadIdx.x >> 5] + incl - sum;
    for (int i = lo; i < hi; ++i) {
        const uint32_t c = counts[i];
        counts[i] = run;
        run += c;
    }
}

template <typename KeyT>
__global__ void __launch_bounds__(kIxThreads) ix_write_records_kernel(const __grid_constant__ IxParams P) {
    pdl_begin();
    __shared__ uint32_t cnt[kIxItems][kIxWarps];  // heads per (item row, warp), then exclusive offsets
    __shared__ uint32_t warp_sums[kIxWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int item = ix_find_item<true>(P, blockIdx.x);
    const IxItem& it = P.item[item];
    const uint32_t tbase = (blockIdx.x - it.rec_tile_start) * kIxTile;
    const uint32_t gt = P.tile0 + blockIdx.x;  // tile number within the whole call
    uint32_t off = 0;
    if (P.rec_inline) {  // buckets before this tile
        for (uint32_t j = threadIdx.x; j < gt; j += kIxThreads) off += P.rec_counts[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) off += __shfl_xor_sync(0xffffffffu, off, o);
        if (lane == 0) warp_sums[warp] = off;
    }
    KeyT k[kIxItems];
    const uint32_t flags = ix_head_flags((const KeyT*)P.kin + it.seg_start, tbase, it.n, k);
    uint32_t before[kIxItems];  // heads of lower lanes in my warp, per item row
#pragma unroll
    for (int i = 0; i < kIxItems; ++i) {
        const uint32_t ballot = __ballot_sync(0xffffffffu, (flags >> i) & 1u);
        before[i] = __popc(ballot & ((1u << lane) - 1u));
        if (lane == 0) cnt[i][warp] = __popc(ballot);
    }
    __syncthreads();
    if (threadIdx.x < 32) {  // exclusive scan of the 16 x 8 counts in position order (4 per lane)
        constexpr int kCells = kIxItems * kIxWarps, kPer = kCells / 32;
        uint32_t* flat = &cnt[0][0];
        uint32_t v[kPer], sum = 0;
#pragma unroll
        for (int j = 0; j < kPer; ++j) { v[j] = flat[threadIdx.x * kPer + j]; sum += v[j]; }
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        uint32_t run = incl - sum;
#pragma unroll
        for (int j = 0; j < kPer; ++j) { flat[threadIdx.x * kPer + j] = run; run += v[j]; }
        if (P.rec_inline && lane == 31 && gt == (uint32_t)P.rec_total - 1) {  // the last tile of the call: nnz
            uint32_t before_me = 0;
#pragma unroll
            for (int w = 0; w < kIxWarps; ++w) before_me += warp_sums[w];
            *P.nnz = (int64_t)(before_me + incl);
        }
    }
    __syncthreads();
    uint32_t tile_off = 0;
    if (P.rec_inline) {
#pragma unroll
        for (int w = 0; w < kIxWarps; ++w) tile_off += warp_sums[w];
    } else {
        tile_off = P.rec_counts[gt];
    }
    const uint64_t slot = (uint64_t)(P.slot0 + item) << P.row_bits;
    int32_t m0[kIxItems];
#pragma unroll
    for (int i = 0; i < kIxItems; ++i)  // the first member of every position, heads or not: 16 coalesced loads in flight
        m0[i] = __ldg(P.vin + it.seg_start + min(tbase + i * kIxThreads + threadIdx.x, it.n - 1));
#pragma unroll
    for (int i = 0; i < kIxItems; ++i) {
        if ((flags >> i) & 1u) {
            BucketRec r;
            r.start = it.seg_start + tbase + i * kIxThreads + threadIdx.x;
            r.m0 = m0[i];
            r.key = slot | (uint64_t)k[i];
            P.recs[tile_off + cnt[i][warp] + before[i]] = r;
        }
    }
}
#endif  // __CUDACC__

}  // namespace etb
