// etb_index.cu -- host side of index! (K4): workspace layout and the launches of etb_index.cuh.
//
// Replaces the reference's Indexer (histogram!/prefixsum!/remap!, src/utils.jl:131-314 -- a stable counting sort
// of occurrence -> delta column by table row), for every table of an ensemble in one call
// (src/sparseupdate.jl:211-213).
#include <algorithm>
#include <stdlib.h>
#include <vector>

#include "etb_index.cuh"

namespace etb {

static int bits_for(uint64_t count) {  // bits needed to represent 0 .. count-1
    int b = 0;
    while (b < 63 && (1ull << b) < count) ++b;
    return b;
}

static size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// ranking of the scatter kernel (etb_index.cuh): ETB_IX_RANK=0 ballots for every row, 1 atomics + ballots on demand
static int ix_rank_mode() {
    static const int mode = [] {
        const char* e = getenv("ETB_IX_RANK");
        return e ? atoi(e) : 1;
    }();
    return mode;
}

// tile of the sort kernels: 16 positions per thread, ETB_IX_THREADS=256|512 threads per CTA (default: 512 from
// 1 M occurrences up -- per-tile work is amortised over twice the elements -- and 256 below, for more CTAs)
static int ix_threads_for(int64_t n_total) {
    static const int forced = [] {
        const char* e = getenv("ETB_IX_THREADS");
        return e ? atoi(e) : 0;
    }();
    if (forced == 256 || forced == 512) return forced;
    return n_total >= (1 << 20) ? 512 : 256;
}

int32_t make_layout(const etb_update_item* items, int32_t n_items, IndexLayout& L) {
    ETB_REQUIRE(n_items >= 0, "etb_index: negative item count");
    ETB_REQUIRE(n_items == 0 || items, "etb_index: null items");
    int64_t n_total = 0, max_rows = 1;
    size_t max_row_bytes = 16;
    for (int i = 0; i < n_items; ++i) {
        const etb_update_item& it = items[i];
        if (int32_t st = validate_table(it.table, "etb_index")) return st;
        ETB_REQUIRE(idx_elt_valid(it.idx_elt), "etb_index: item %d: index type must be ETB_I32/ETB_I64", i);
        ETB_REQUIRE(it.batch >= 0 && it.batch < 0x7fffffffll, "etb_index: item %d: bad batch %lld", i, (long long)it.batch);
        ETB_REQUIRE(it.bag >= 0 && it.bag <= 0x7fffffffll, "etb_index: item %d: bad bag %lld", i, (long long)it.bag);
        ETB_REQUIRE(it.bag == 0 || it.ld_idx >= it.bag, "etb_index: item %d: ld_idx < bag", i);
        const int64_t n_i = it.batch * (it.bag ? it.bag : 1);
        ETB_REQUIRE(n_i < (1ll << 30), "etb_index: item %d: %lld occurrences exceed the 2^30 limit per table and call", i, (long long)n_i);
        n_total += n_i;
        max_rows = std::max(max_rows, it.table.nrows);
        // partial rows of long buckets are kept in the arithmetic type (Float32 for the half types)
        max_row_bytes = std::max(max_row_bytes, (size_t)it.table.dim * std::max<size_t>(4, elt_bytes(it.table.elt)));
    }
    ETB_REQUIRE(n_total < 0x7fffffffll, "etb_index: %lld occurrences exceed the 2^31 limit of one call", (long long)n_total);
    L.n_total = n_total;
    L.row_bits = std::max(1, bits_for((uint64_t)max_rows));
    L.slot_bits = bits_for((uint64_t)std::max(1, n_items));
    ETB_REQUIRE(L.row_bits + L.slot_bits <= 64, "etb_index: (table, row) does not fit 64 bits");
    L.key_bytes = L.row_bits <= 32 ? 4 : 8;  // the sort key is the row alone: tables are sorted segment by segment
    L.npasses = ix_plan(L.row_bits, L.width, L.shift);
    L.nb_max = 1 << L.width[0];  // the widest digit comes first
    L.threads = L.key_bytes == 8 ? 256 : ix_threads_for(n_total);
    L.total_tiles = L.rec_tiles = 0;
    for (int i = 0; i < n_items; ++i) {
        const int64_t n_i = items[i].batch * (items[i].bag ? items[i].bag : 1);
        L.total_tiles += (n_i + L.threads * kIxItems - 1) / (L.threads * kIxItems);
        L.rec_tiles += (n_i + kIxTile - 1) / kIxTile;
    }
    const size_t n = (size_t)std::max<int64_t>(n_total, 1);
    L.max_long = n / kLongThreshold + 1;
    L.max_medium = n / (kShortMax + 1) + 1;
    L.max_chunks = n / kLongChunk + L.max_long;  // partial rows
    L.max_tasks = L.max_chunks + L.max_medium;    // task list = long chunks + medium buckets
    L.partial_pitch = align_up(max_row_bytes, 16);
    size_t off = 0;
    for (int b = 0; b < 2; ++b) { L.off_keys[b] = off; off = align_up(off + n * L.key_bytes); }
    for (int b = 0; b < 2; ++b) { L.off_vals[b] = off; off = align_up(off + n * sizeof(int32_t)); }
    L.off_recs = off; off = align_up(off + (n + 1) * sizeof(BucketRec));
    L.off_nnz = off; off = align_up(off + sizeof(int64_t));
    L.off_tile_hist = off; off = align_up(off + (size_t)(L.total_tiles + 1) * L.nb_max * sizeof(uint32_t));
    L.off_digit_total = off; off = align_up(off + (size_t)std::max(1, std::min(n_items, kIxMaxItems)) * L.nb_max * sizeof(uint32_t));
    L.off_rec_counts = off; off = align_up(off + (size_t)(L.rec_tiles + 1) * sizeof(uint32_t));
    L.off_counters = off; off = align_up(off + sizeof(LongCounters));
    L.off_long = off; off = align_up(off + L.max_long * sizeof(LongRec));
    L.off_chunks = off; off = align_up(off + L.max_tasks * sizeof(ChunkRec));
    L.off_partials = off; off = align_up(off + L.max_chunks * L.partial_pitch);
    L.total = align_up(off);
    return ETB_OK;
}

// launches of one pass: one instantiation per (key type, index type, ranking, tile size)
template <typename KeyT, typename SrcT, int RANK, int THREADS>
static cudaError_t ix_launch_pass(const IxParams& P, cudaStream_t s) {
    const int nb = 1 << P.width[P.pass];
    launch_k(ix_hist_kernel<KeyT, SrcT, THREADS>, P.ntiles, THREADS, 0, s, P);
    ++launch_counter();
    if ((int64_t)nb * P.n_items < 8192)  // few (table, digit) pairs: a warp each
        launch_k(ix_scan_warp_kernel, dim3((nb + 7) / 8, P.n_items), 256, 0, s, P, THREADS * kIxItems);
    else
        launch_k(ix_scan_kernel, dim3((nb + 255) / 256, P.n_items), 256, 0, s, P, THREADS * kIxItems);
    ++launch_counter();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const size_t smem = ix_scatter_smem<KeyT, THREADS>(nb);
    static thread_local size_t configured = 0;  // opt in to > 48 KB of dynamic shared memory once per size
    if (smem > configured) {
        e = cudaFuncSetAttribute(ix_scatter_kernel<KeyT, SrcT, RANK, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(ix_scatter_kernel<KeyT, SrcT, RANK, THREADS>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    launch_k(ix_scatter_kernel<KeyT, SrcT, RANK, THREADS>, P.ntiles, THREADS, smem, s, P);
    ++launch_counter();
    return cudaGetLastError();
}

template <typename KeyT, typename SrcT>
static cudaError_t ix_launch_pass_cfg(const IxParams& P, int threads, cudaStream_t s) {
    if constexpr (sizeof(KeyT) == 8) {
        return ix_launch_pass<KeyT, SrcT, 1, 256>(P, s);
    } else {
        const int rank = ix_rank_mode();
        if (threads == 512) return rank ? ix_launch_pass<KeyT, SrcT, 1, 512>(P, s) : ix_launch_pass<KeyT, SrcT, 0, 512>(P, s);
        return rank ? ix_launch_pass<KeyT, SrcT, 1, 256>(P, s) : ix_launch_pass<KeyT, SrcT, 0, 256>(P, s);
    }
}

template <typename KeyT>
static cudaError_t ix_launch_pass_any(const IxParams& P, int idx_elt, int threads, cudaStream_t s) {
    if (P.pass > 0) return ix_launch_pass_cfg<KeyT, void>(P, threads, s);
    return idx_elt == ETB_I64 ? ix_launch_pass_cfg<KeyT, long long>(P, threads, s) : ix_launch_pass_cfg<KeyT, int>(P, threads, s);
}

int32_t index_impl(void* ws, size_t ws_bytes, const etb_update_item* items, int32_t n_items,
                   etb_index_view* view, cudaStream_t stream) {
    IndexLayout L;
    if (int32_t st = make_layout(items, n_items, L)) return st;
    ETB_REQUIRE(ws != nullptr, "etb_index: null workspace");
    ETB_REQUIRE(((uintptr_t)ws % 256) == 0, "etb_index: workspace must be 256-byte aligned");
    if (ws_bytes < L.total)
        return fail(ETB_ERR_WORKSPACE, "etb_index: workspace has %zu bytes, needs %zu", ws_bytes, L.total);
    char* base = (char*)ws;
    BucketRec* recs = (BucketRec*)(base + L.off_recs);
    int64_t* nnz = (int64_t*)(base + L.off_nnz);
    int32_t* vals[2] = {(int32_t*)(base + L.off_vals[0]), (int32_t*)(base + L.off_vals[1])};
    void* keys[2] = {base + L.off_keys[0], base + L.off_keys[1]};
    const int fin = (L.npasses - 1) & 1;  // pass q writes buffer q & 1

    if (L.n_total == 0) {
        ETB_CUDA(cudaMemsetAsync(nnz, 0, sizeof(int64_t), stream));
    } else {
        static thread_local std::vector<IxParams> groups;  // one launch group per kIxMaxItems tables
        groups.clear();
        int64_t seg = 0, tile0 = 0;
        const int idx_elt = items[0].idx_elt;
        for (int i0 = 0; i0 < n_items; i0 += kIxMaxItems) {
            const int n = std::min(kIxMaxItems, n_items - i0);
            groups.emplace_back();
            IxParams& P = groups.back();
            uint32_t nt = 0, rt = 0;
            const uint32_t tile = (uint32_t)L.threads * kIxItems;
            for (int j = 0; j < n; ++j) {
                const etb_update_item& it = items[i0 + j];
                IxItem& d = P.item[j];
                d.idx = it.idx;
                d.n = (uint32_t)(it.batch * (it.bag ? it.bag : 1));
                d.bag = it.bag ? (uint32_t)it.bag : 1u;  // a vector of indices is a bag-1 matrix with unit stride
                d.ld_idx = it.bag ? (uint64_t)it.ld_idx : 1ull;
                ix_magic(d.bag, &d.magic, &d.mshift);
                d.seg_start = (uint32_t)seg;
                d.tile_start = nt;
                d.rec_tile_start = rt;
                d.pad = 0;
                seg += d.n;
                nt += (d.n + tile - 1) / tile;
                rt += (d.n + kIxTile - 1) / kIxTile;
                ETB_REQUIRE(it.idx_elt == idx_elt, "etb_index: all items of one call must share the index element type");
                ETB_REQUIRE(d.n == 0 || it.idx, "etb_index: item %d: null indices", i0 + j);
            }
            P.n_items = n;
            P.ntiles = (int32_t)nt;
            P.rec_tiles = (int32_t)rt;
            P.slot0 = i0;
            P.row_bits = L.row_bits;
            memcpy(P.width, L.width, sizeof(P.width));
            memcpy(P.shift, L.shift, sizeof(P.shift));
            P.tile_hist = (uint32_t*)(base + L.off_tile_hist);
            P.digit_total = (uint32_t*)(base + L.off_digit_total);
            P.recs = recs;
            P.nnz = nnz;
            P.rec_counts = (uint32_t*)(base + L.off_rec_counts);
            P.tile0 = (uint32_t)tile0;
            tile0 += rt;
        }
        const bool rec_inline = L.rec_tiles <= 2048;  // few record tiles: no scan kernel
        // small tables (every table at most one tile): one CTA sorts a table in shared memory, all passes at once
        int64_t max_n = 0;
        for (int i = 0; i < n_items; ++i) max_n = std::max<int64_t>(max_n, items[i].batch * (items[i].bag ? items[i].bag : 1));
        const bool small = max_n <= kIxTile && L.key_bytes == 4;
        for (IxParams& P : groups) {
            P.rec_inline = rec_inline;
            P.rec_total = (int32_t)L.rec_tiles;
            if (P.ntiles == 0) continue;
            if (small) {
                P.pass = 0;
                P.kin = nullptr;
                P.vin = nullptr;
                P.kout = keys[fin];
                P.vout = vals[fin];
                const size_t smem = ix_small_smem<uint32_t>(L.nb_max);
                const int rank = ix_rank_mode();
#define ETB_SMALL(SRC, RK)                                                                                                 \
    do {                                                                                                                   \
        static thread_local size_t configured = 0;                                                                         \
        if (smem > configured) {                                                                                           \
            ETB_CUDA(cudaFuncSetAttribute(ix_small_kernel<uint32_t, SRC, RK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            configured = smem;                                                                                             \
        }                                                                                                                  \
        launch_k(ix_small_kernel<uint32_t, SRC, RK>, P.n_items, kIxThreads, smem, stream, P, L.npasses);                         \
    } while (0)
                if (idx_elt == ETB_I64) { if (rank) ETB_SMALL(long long, 1); else ETB_SMALL(long long, 0); }
                else { if (rank) ETB_SMALL(int, 1); else ETB_SMALL(int, 0); }
#undef ETB_SMALL
                ETB_LAUNCHED();
                P.kin = keys[fin];
                P.vin = vals[fin];
                continue;
            }
            // K4a: the partition passes, least significant digit first; pass q reads buffer (q - 1) & 1 (the first one
            // reads the index arrays) and writes buffer q & 1
            for (int q = 0; q < L.npasses; ++q) {
                P.pass = q;
                P.kin = q ? keys[(q - 1) & 1] : nullptr;
                P.vin = q ? vals[(q - 1) & 1] : nullptr;
                P.kout = keys[q & 1];
                P.vout = vals[q & 1];
                const cudaError_t e = L.key_bytes == 4 ? ix_launch_pass_any<uint32_t>(P, idx_elt, L.threads, stream)
                                                       : ix_launch_pass_any<uint64_t>(P, idx_elt, L.threads, stream);
                if (e != cudaSuccess) return fail(ETB_ERR_CUDA, "etb_index: %s", cudaGetErrorString(e));
            }
            // K4b: bucket heads per tile
            P.kin = keys[fin];
            P.vin = vals[fin];
            if (L.key_bytes == 4) launch_k(ix_count_heads_kernel<uint32_t>, P.rec_tiles, kIxThreads, 0, stream, P);
            else launch_k(ix_count_heads_kernel<uint64_t>, P.rec_tiles, kIxThreads, 0, stream, P);
            ETB_LAUNCHED();
        }
        if (!rec_inline) {
            launch_k(ix_scan_counts_kernel, 1, 1024, 0, stream, (uint32_t*)(base + L.off_rec_counts), (int)L.rec_tiles, nnz);
            ETB_LAUNCHED();
        }
        for (IxParams& P : groups) {  // one record per bucket head, numbered over the whole call
            if (P.ntiles == 0) continue;
            if (L.key_bytes == 4) launch_k(ix_write_records_kernel<uint32_t>, P.rec_tiles, kIxThreads, 0, stream, P);
            else launch_k(ix_write_records_kernel<uint64_t>, P.rec_tiles, kIxThreads, 0, stream, P);
            ETB_LAUNCHED();
        }
    }
    if (view) {
        view->keys = keys[fin];
        view->map = vals[fin];
        view->records = recs;
        view->nnz = nnz;
        view->scratch = base + L.off_counters;
        view->n_total = L.n_total;
        view->key_bytes = L.key_bytes;
        view->row_bits = L.row_bits;
        view->num_splits = 0;
        view->this_split = 0;
    }
    return ETB_OK;
}

}  // namespace etb

using namespace etb;

extern "C" {

int32_t etb_index_workspace_bytes(const etb_update_item* items_host, int32_t n_items, size_t* bytes_host) {
    ETB_REQUIRE(bytes_host, "etb_index_workspace_bytes: null output");
    IndexLayout L;
    if (int32_t st = make_layout(items_host, n_items, L)) return st;
    *bytes_host = L.total;
    return ETB_OK;
}

int32_t etb_index(void* workspace, size_t workspace_bytes, const etb_update_item* items_host, int32_t n_items,
                  etb_index_view* view_host, void* stream) {
    ETB_API_RANGE();
    launch_counter() = 0;
    return index_impl(workspace, workspace_bytes, items_host, n_items, view_host, (cudaStream_t)stream);
}

}  // extern "C"
