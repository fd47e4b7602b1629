// etb_layout.cuh -- workspace layout of index! / update! shared by etb_index.cu (K4) and etb_update.cu (K5).
#pragma once
#include "etb_common.cuh"

namespace etb {

// One record per bucket, written by K4 and read by K5 with ONE coalesced load per 32 buckets
// (the reference's `cumulative` entry (col, offset), src/utils.jl:101-106, plus the first member).
struct alignas(16) BucketRec {
    uint32_t start;  // first sorted position of the bucket
    int32_t m0;      // delta column of its first member
    uint64_t key;    // slot << row_bits | (row - 1)
};

// Bucket classes of the update: SHORT (<= kShortMax members) finish inside the main kernel; MEDIUM
// (kShortMax < members <= kLongThreshold, and every larger bucket in strict mode) become tasks of
// bucket_tasks_kernel and are reduced strictly in order with 8 rows in flight; LONG (> kLongThreshold,
// ETB_UPDATE_SPLIT_LONG only) are cut into kLongChunk-member chunk tasks whose partial rows
// long_combine_kernel adds in a fixed order.
constexpr int kShortMax = 4;
constexpr int kLongThreshold = 128;  // buckets with more members than this are "long"
constexpr int kLongChunk = 128;      // members per partial sum
struct LongCounters { uint32_t n_long, n_chunks /* task cursor */, n_partials, pad; };
struct LongRec { uint32_t bucket, chunk_base, nchunks, pad; };
struct ChunkRec { uint32_t long_id, chunk; };  // long_id == kMediumTask: a MEDIUM bucket, chunk = its bucket index
constexpr uint32_t kMediumTask = 0xffffffffu;

struct IndexLayout {
    int64_t n_total;
    int32_t row_bits, slot_bits, key_bytes;
    // the sort (etb_index.cuh)
    int32_t npasses, nb_max, threads;
    uint8_t width[8], shift[8];
    int64_t total_tiles, rec_tiles;  // sort tiles (16 * threads positions) and record tiles (4096) over all items
    size_t off_tile_hist, off_digit_total, off_rec_counts;
    size_t max_long, max_chunks, max_medium, max_tasks, partial_pitch;
    size_t off_keys[2], off_vals[2], off_recs, off_nnz, off_counters, off_long, off_chunks, off_partials, total;
};


// the layout is a pure function of the items (and of ETB_IX_THREADS): K5 recomputes it to find K4's scratch regions
int32_t make_layout(const etb_update_item* items, int32_t n_items, IndexLayout& L);
int32_t index_impl(void* ws, size_t ws_bytes, const etb_update_item* items, int32_t n_items, etb_index_view* view,
                   cudaStream_t stream);

}  // namespace etb
