/*
 * embtab_b200.h -- C ABI of libembtab_b200.so
 *
 * The B200 (sm_100a) replacement for the data-parallel hot path of
 * darchr/EmbeddingTables.jl:
 *
 *     lookup / lookup!                       (reference src/lookup.jl:35-182)
 *     maplookup / maplookup!  (3 strategies) (reference src/lookup.jl:220-371)
 *     SparseEmbeddingUpdate + index!         (reference src/sparseupdate.jl:6-32, src/utils.jl:131-314)
 *     update!(::Descent, table, grad)        (reference src/sparseupdate.jl:46-238)
 *
 * A Julia host binds these with `ccall((:etb_xxx, libembtab_b200), Cint, (...), ...)`;
 * the Python host mirror (embeddingtables.jl_b200/embtab) binds them with ctypes.
 * See INTEGRATION.md for the reference-side stubs.
 *
 * Conventions
 *   - every function returns an int32 status: 0 = ok, nonzero = etb_status;
 *     etb_last_error() returns a thread-local message for the last failure.
 *   - every pointer is a DEVICE pointer unless its name ends in `_host` or the
 *     comment says "host array" (descriptor arrays are host arrays: the library
 *     copies them into kernel parameters, so a call never allocates, never
 *     synchronises and can be captured into a CUDA graph).
 *   - matrices are column-major with an explicit leading dimension in ELEMENTS,
 *     exactly Julia's layout: one embedding row == one Julia column == `dim`
 *     contiguous elements (reference README.md:304-307).
 *   - indices are 1-BASED (Julia), int64 by default (ETB_I64) or int32 (ETB_I32).
 *     As in the reference there is NO bounds check on the hot path
 *     (`@inbounds`, reference src/lookup.jl:57-84): an out-of-range index is UB.
 *   - every compute call is asynchronous on `stream` (a cudaStream_t passed as void*).
 *   - the library never frees or retains caller memory; workspaces are
 *     caller-owned after a *_workspace_bytes query (the GPU analogue of the
 *     reference's caller-owned `Indexer`, src/utils.jl:288-304).
 */
#ifndef EMBTAB_B200_H
#define EMBTAB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define ETB_VERSION 100 /* 0.1.0 */

typedef enum etb_status {
    ETB_OK = 0,
    ETB_ERR_INVALID = 1,   /* bad argument (null pointer, negative size, unsupported dtype) */
    ETB_ERR_CUDA = 2,      /* a CUDA runtime call failed; see etb_last_error() */
    ETB_ERR_WORKSPACE = 3, /* caller workspace too small */
    ETB_ERR_UNSUPPORTED = 4
} etb_status;

/* element types of tables / outputs / deltas, and of index arrays (I32 / I64 only) */
typedef enum etb_dtype {
    ETB_F32 = 0,
    ETB_F64 = 1,
    ETB_I32 = 2,
    ETB_I64 = 3,
    /* Extension (no reference counterpart -- the reference's tables are Float32/Float64/integer):
     * half-precision storage with Float32 arithmetic.  Pooled sums and SGD updates convert each element to
     * Float32, accumulate in the reference's order, and round to nearest-even ONCE when the result is stored
     * (output / cotangent have the table's element type).  The update epilogue is `row - eta*acc` in Float32
     * (ETB_UPDATE_FMA is ignored).  dim and every leading dimension must be even (4-byte aligned rows). */
    ETB_F16 = 4,
    ETB_BF16 = 5
} etb_dtype;

/* flags of etb_sgd_update* */
enum {
    /* epilogue `row = fma(-eta, acc, row)` -- what the reference's specialised kernel
     * computes (`muladd`, src/sparseupdate.jl:123-127, src/simd.jl:54-59).  Without the
     * flag the epilogue is `row - eta*acc` with separate roundings (the generic kernel,
     * src/sparseupdate.jl:88). */
    ETB_UPDATE_FMA = 1,
    /* allow long duplicate runs (hot Zipf rows) to be reduced as fixed-size chunks that
     * are combined in a fixed order: deterministic, atomics-free, but not bit-identical
     * to the strictly sequential order of the reference.  Off = strictly sequential. */
    ETB_UPDATE_SPLIT_LONG = 2
};

/*
 * Device table descriptor: replaces `columnpointer(table, i)` for
 *   SimpleEmbedding (reference src/simple.jl:52-55): base + (i-1)*ld*sizeof(T)
 *   SplitEmbedding  (reference src/split.jl:59-65,81-86):
 *       chunks[(i-1) / shard_rows] + ((i-1) % shard_rows)*ld*sizeof(T)
 */
typedef struct etb_table {
    void* base;          /* Simple: address of embedding row 1.  Split: NULL.  Cached: row 1 of the HOST-resident
                          * table (page-locked, device-addressable)                      */
    void* const* chunks; /* Split: DEVICE array of chunk base pointers.  Simple: NULL.  Cached: HOST pointer to
                          * an etb_cache_desc                                            */
    int64_t nrows;       /* embedding rows (Julia size(table, 2))                       */
    int64_t shard_rows;  /* Split: rows per chunk (cols_per_shard).  Simple: 0.  Cached: ETB_TABLE_CACHED */
    int32_t dim;         /* featuresize (Julia size(table, 1)), elements                */
    int32_t ld;          /* elements between consecutive embedding rows (>= dim)        */
    int32_t elt;         /* etb_dtype                                                   */
    int32_t reserved;
} etb_table;

/*
 * Host-tier table with an HBM row cache (no reference counterpart -- the hook the reference leaves for
 * it is the IndexingContext argument of columnpointer, src/EmbeddingTables.jl:74-77, 87-93, which lets a table
 * resolve a row differently in the Forward and Update phases).  The whole table lives in page-locked host memory
 * that the GPU can address (`base`); `rows` caches up to `capacity` of its rows in HBM and `slot_of_row` says which.
 * Every kernel resolves a row as  slot_of_row[i-1] >= 0 ? rows + slot*ld : base + (i-1)*ld,  so results are bit for
 * bit those of an all-HBM table.  A cached row is authoritative in HBM (update! writes it there) until
 * etb_cache_flush copies it back.  etb_cache_admit, run after update! (the Update phase), admits the rows the batch
 * touched most often: it histograms the occurrence counts of the rows that are still on the host and admits those at
 * or above the count at which they fit into the free slots (never below `min_count`).
 */
#define ETB_TABLE_CACHED (-1)
typedef struct etb_cache_desc { /* a HOST struct of DEVICE pointers */
    void* rows;            /* capacity x ld elements                                       */
    int32_t* slot_of_row;  /* nrows entries: cache slot of each row, -1 = lives on the host */
    int32_t* row_of_slot;  /* capacity entries: 0-based row held by each slot in use        */
    int32_t* cursor;       /* one int32: slots in use                                       */
    int64_t capacity;
    int32_t* hist;         /* ETB_CACHE_HIST_BINS int32 of scratch for etb_cache_admit       */
} etb_cache_desc;
#define ETB_CACHE_HIST_BINS 64

/*
 * One table's share of an ensemble lookup: replaces one `lookup!(out[i], x[i], I[i])`
 * of maplookup! (reference src/lookup.jl:233-241, 263-276) or one
 * `lookup!(view(dst, rows_i, :), x[i], I[i])` of the PreallocationStrategy
 * (reference src/lookup.jl:334-367) -- there `dst` points at row
 * `prependrows + sum(featuresize(x[1:i-1]))` of the concatenated matrix and
 * `ld_dst` is that matrix's row count.
 */
typedef struct etb_lookup_item {
    etb_table table;
    const void* idx; /* bag == 0: `batch` indices.  bag >= 1: bag x batch, column-major */
    void* dst;       /* dim x batch, column-major, leading dimension ld_dst             */
    int64_t ld_dst;
    int64_t batch;   /* output columns                                                  */
    int64_t bag;     /* 0 = non-reducing gather; >= 1 = pooled sum over `bag` rows      */
    int64_t ld_idx;  /* elements between index columns (>= bag); ignored when bag == 0  */
    int32_t idx_elt; /* ETB_I64 or ETB_I32                                              */
    int32_t reserved;
} etb_lookup_item;

/*
 * One table's share of an index/update: the `(delta, indices)` pair of a
 * SparseEmbeddingUpdate (reference src/sparseupdate.jl:6-13) plus its table.
 * `delta` is dim x batch with leading dimension ld_delta (a row-slice view of the
 * concatenated cotangent in the Preallocation pullback, reference src/lookup.jl:383-386).
 * Meaning (reference src/sparseupdate.jl:16-32): G[:, c] += delta[:, j] for every
 * c in indices[:, j] (matrix) or c = indices[j] (vector).
 */
typedef struct etb_update_item {
    etb_table table;
    const void* delta;
    int64_t ld_delta;
    const void* idx; /* same layout as etb_lookup_item.idx */
    int64_t batch;
    int64_t bag;     /* 0 = vector of indices (one per delta column) */
    int64_t ld_idx;
    int32_t idx_elt;
    int32_t flags;   /* per-table ETB_UPDATE_FMA (OR-ed with the call's flags): the reference picks the
                      * epilogue per table type (src/sparseupdate.jl:131-154), so an ensemble may mix them */
} etb_update_item;

/* ---------------------------------------------------------------- runtime ------------- */
/* Replaces Julia `Array` storage management for HBM-resident tables. */
int32_t etb_version(void);
const char* etb_last_error(void);
int32_t etb_device_count(int32_t* count_host);
int32_t etb_init(int32_t device); /* cudaSetDevice + context warm-up */
int32_t etb_malloc(void** ptr_host, size_t bytes);
int32_t etb_free(void* ptr);
int32_t etb_malloc_host(void** ptr_host, size_t bytes); /* pinned host memory */
int32_t etb_free_host(void* ptr_host);
int32_t etb_memcpy_h2d(void* dst, const void* src_host, size_t bytes, void* stream);
int32_t etb_memcpy_d2h(void* dst_host, const void* src, size_t bytes, void* stream);
int32_t etb_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream);
/* Strided (2-D) copies: `height` runs of `width_bytes` contiguous bytes, run k at base + k * pitch.  This is the
 * row-slice view of a column-major matrix -- e.g. one table's rows of the concatenated cotangent
 * (reference src/utils.jl:50-63 Slicer, src/lookup.jl:374-389) -- moved without staging. */
int32_t etb_memcpy2d_h2d(void* dst, size_t dst_pitch, const void* src_host, size_t src_pitch, size_t width_bytes,
                         size_t height, void* stream);
int32_t etb_memcpy2d_d2h(void* dst_host, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes,
                         size_t height, void* stream);
/* device-to-device, also between a local buffer and a peer buffer mapped with etb_ipc_import: the copy engines move
 * row blocks over NVLink while the SMs keep computing (the multi-GPU exchange of embtab/dist.py, exchange="copy") */
int32_t etb_memcpy2d_d2d(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes,
                         size_t height, void* stream);
int32_t etb_memset(void* dst, int32_t byte, size_t bytes, void* stream);
int32_t etb_stream_create(void** stream_host);
int32_t etb_stream_sync(void* stream);
int32_t etb_stream_destroy(void* stream);

/* ---------------------------------------------------------------- lookup -------------- */
/* K1. O[:, j] = A[:, I[j]], bit copy for any element type.
 * Replaces lookup!(dst, src, indices::AbstractVector), reference src/lookup.jl:51-102. */
int32_t etb_gather(void* dst, int64_t ld_dst, const etb_table* table_host, const void* idx,
                   int32_t idx_elt, int64_t n, void* stream);

/* K2. O[:, j] = ((A[:, I[1,j]] + A[:, I[2,j]]) + ...) + A[:, I[bag,j]], strictly in bag order,
 * accumulator seeded with the first row.  Replaces lookup!(O, A, I::AbstractMatrix),
 * reference src/lookup.jl:108-182 (lookup_generic! / lookup_static_inner / lookup_static!). */
int32_t etb_pooled_sum(void* dst, int64_t ld_dst, const etb_table* table_host, const void* idx,
                       int32_t idx_elt, int64_t bag, int64_t batch, int64_t ld_idx, void* stream);

/* K3. The whole ensemble in one launch per kernel class (one launch when all tables share
 * dim/dtype, the DLRM case).  Replaces maplookup! for DefaultStrategy, SimpleParallelStrategy
 * and PreallocationStrategy, reference src/lookup.jl:233-241, 263-276, 316-371.
 * `items_host` is a host array. */
int32_t etb_maplookup(const etb_lookup_item* items_host, int32_t n_items, void* stream);

/* Number of kernel launches the last etb_maplookup / etb_sgd_update / etb_index call made on
 * this thread (bench.py's gpu_launches claim is counted, not guessed). */
int32_t etb_last_launch_count(void);

/* ---------------------------------------------------------------- index! -------------- */
/* K4. Group the occurrences of each distinct (table, row) of an ensemble, stable in
 * occurrence order (column-major traversal of the index matrix; the delta column of flat
 * position p is p / bag).  Replaces index!(indexer, indices, nrows) = histogram! +
 * prefixsum! + remap!, reference src/utils.jl:131-314, for all tables of an ensemble at once
 * (reference src/sparseupdate.jl:211-213).  Buckets come out in ascending (table, row) order
 * instead of first-seen order; bucket members keep the reference's order.
 *
 * The result is described by an etb_index_view (a host POD of device pointers into the
 * caller's workspace -- the GPU analogue of the Indexer's `cumulative` and `map` fields,
 * reference src/utils.jl:288-293):
 *   keys[n_total]  sorted composite keys (table_slot << row_bits | row-1), uint32 or uint64
 *   map[n_total]   delta column (0-based) of each sorted position        (reference `map`)
 *   records[nnz]   one etb_bucket_record per bucket                      (reference `cumulative`)
 *   nnz            number of buckets (device int64; bucket s ends at records[s+1].start, or at
 *                  n_total for the last one)
 */
typedef struct etb_bucket_record {
    uint32_t start; /* first sorted position of the bucket               */
    int32_t m0;     /* delta column (0-based) of the bucket's first member */
    uint64_t key;   /* table_slot << row_bits | row-1                      */
} etb_bucket_record;

typedef struct etb_index_view {
    const void* keys;       /* device */
    const int32_t* map;     /* device */
    const void* records;    /* device: etb_bucket_record[nnz] */
    const int64_t* nnz;     /* device */
    void* scratch;          /* device: internal scratch of ETB_UPDATE_SPLIT_LONG */
    int64_t n_total;        /* sum of occurrences over items */
    int32_t key_bytes;      /* 4 or 8 */
    int32_t row_bits;       /* key = slot << row_bits | (row-1) */
    /* IndexerView(indexer, num_splits, this_split), reference src/utils.jl:320-338: restrict an
     * update to buckets (this_split-1)*s .. min(this_split*s, nnz)-1 with s = cdiv(nnz+1,
     * num_splits).  num_splits == 0 (what etb_index writes) = all buckets. */
    int32_t num_splits;
    int32_t this_split;     /* 1-based */
} etb_index_view;

int32_t etb_index_workspace_bytes(const etb_update_item* items_host, int32_t n_items,
                                  size_t* bytes_host);
int32_t etb_index(void* workspace, size_t workspace_bytes, const etb_update_item* items_host,
                  int32_t n_items, etb_index_view* view_host, void* stream);

/* ---------------------------------------------------------------- update! ------------- */
/* K5. For every bucket (distinct row k of table t): acc = 0; acc += delta_t[:, col] for the
 * bucket's members in order; A_t[:, k] = fma(-eta, acc, A_t[:, k]) (ETB_UPDATE_FMA, in the call's
 * flags for every table or in etb_update_item.flags per table) or A_t[:, k] - eta*acc.
 * No atomics on table data: one bucket = one table row = one writer.
 * Replaces update!(table, update, indexer, alpha) reference src/sparseupdate.jl:57-154 and the
 * ensemble form :199-238.  `view_host` must be the result of etb_index on the SAME items.
 * eta is converted to the table's element type (reference src/sparseupdate.jl:173). */
int32_t etb_sgd_update(const etb_index_view* view_host, const etb_update_item* items_host,
                       int32_t n_items, double eta, int32_t flags, void* stream);

/* Optimiser extension (the reference's update! exists for Flux.Descent only,
 * src/sparseupdate.jl:160-189): row-wise Adagrad on the same buckets.  `states_host` is a host array of n_items
 * device pointers; states_host[t] holds one element per row of table t, in the table's arithmetic type
 * (Float32 for Float32 / Float16 / BFloat16 tables, Float64 for Float64 tables), zero before the first step.
 * For every bucket, with g = its summed cotangent (accumulated exactly as etb_sgd_update does):
 *     h = state[k] + (sum_d g_d^2) / dim;   state[k] = h;   A[:, k] -= (eta / (sqrt(h) + eps)) * g
 * every operation rounded separately, the sum of squares in a fixed order (per lane, then an XOR butterfly over
 * the lanes of the row -- oracle/oracle.py adagrad_update restates it), so results are deterministic.
 * Rows must fit one pass of the kernel (up to 128 vectors: 2 KB with 16-byte alignment), else ETB_ERR_UNSUPPORTED.
 * ETB_UPDATE_FMA is ignored; ETB_UPDATE_SPLIT_LONG applies to the summation of g as usual. */
int32_t etb_adagrad_update(const etb_index_view* view_host, const etb_update_item* items_host,
                           void* const* states_host, int32_t n_items, double eta, double eps, int32_t flags,
                           void* stream);

/* update!(opt, table(s), grad(s)): etb_index followed by etb_sgd_update
 * (reference src/sparseupdate.jl:160-178). */
int32_t etb_index_and_update(void* workspace, size_t workspace_bytes,
                             const etb_update_item* items_host, int32_t n_items, double eta,
                             int32_t flags, void* stream);

/* Host-tier tables (etb_cache_desc above).  etb_cache_admit: for every bucket of `view` (the result of etb_index on
 * these items) whose table is ETB_TABLE_CACHED, whose row is not cached yet and which has >= min_count members, take
 * a free slot, copy the row from host memory into it and publish it in slot_of_row.  Call it when no kernel that
 * uses the tables is in flight on another stream (it is stream-ordered on `stream`).  etb_cache_flush: copy every
 * cached row back to the host table (the cache stays valid). */
int32_t etb_cache_admit(const etb_index_view* view_host, const etb_update_item* items_host, int32_t n_items,
                        int32_t min_count, void* stream);
int32_t etb_cache_flush(const etb_table* table_host, void* stream);

/* Debug/test helper: dense gradient of a SparseEmbeddingUpdate.
 * dst (dim x ncols, ld_dst) must be zeroed by the caller; accumulates in occurrence order.
 * Replaces uncompress(), reference src/sparseupdate.jl:16-32.  f32 / f64 only. */
int32_t etb_uncompress(void* dst, int64_t ld_dst, int32_t dim, int32_t elt, const void* delta,
                       int64_t ld_delta, const void* idx, int32_t idx_elt, int64_t bag,
                       int64_t batch, int64_t ld_idx, void* stream);

/* ---------------------------------------------------------------- multi-GPU ----------- */
/* Table-wise sharded ensembles exchange pooled outputs / cotangents between ranks.  The
 * transport (NCCL all-to-all) is driven by the host layer through torch.distributed; these
 * two kernels are the device-side pack/unpack at either end of it.  No reference counterpart
 * (the reference is single-process).
 *
 * etb_a2a_unpack: recv buffer holds `nranks` blocks; block r is (rows_r x batch_local)
 * column-major dense; it lands at rows [row_off_r, row_off_r + rows_r) of dst (ld_dst).
 * etb_a2a_pack is the inverse (gathers row-blocks of src into a dense send buffer). */
int32_t etb_a2a_unpack(void* dst, int64_t ld_dst, const void* recv, const int64_t* rows_host,
                       const int64_t* row_off_host, int32_t nranks, int64_t batch_local,
                       int32_t elt, void* stream);
int32_t etb_a2a_pack(void* send, const void* src, int64_t ld_src, const int64_t* rows_host,
                     const int64_t* row_off_host, int32_t nranks, int64_t batch_local, int32_t elt,
                     void* stream);

/* Fused exchange over NVLink peer memory (one process per GPU, buffers shared with CUDA IPC).
 * etb_ipc_export/import/close wrap cudaIpcGetMemHandle / cudaIpcOpenMemHandle / cudaIpcCloseMemHandle
 * for buffers allocated with etb_malloc; `handle_host` is ETB_IPC_HANDLE_BYTES bytes of host memory
 * that the host layer ships to the peers (torch.distributed all_gather).
 *
 * With peer-mapped destinations the forward needs no collective at all: etb_maplookup's `dst`
 * pointers are simply addresses inside the PEERS' feature matrices.  etb_a2a_scatter is the backward
 * counterpart: row block r of `src` (rows_host[r] x batch_local at row row_off_host[r]) is stored as a
 * dense rows_host[r] x batch_local matrix at dst_ptrs_host[r] -- a peer address inside owner r's
 * (rows_r x B_global) cotangent buffer.  The caller orders the stores against the consumers with a
 * stream-ordered barrier (kernel completion makes peer stores visible system-wide). */
#define ETB_IPC_HANDLE_BYTES 64
int32_t etb_ipc_export(void* ptr, void* handle_host);
int32_t etb_ipc_import(const void* handle_host, void** ptr_host);
int32_t etb_ipc_close(void* ptr);
int32_t etb_a2a_scatter(void* const* dst_ptrs_host, const void* src, int64_t ld_src,
                        const int64_t* rows_host, const int64_t* row_off_host, int32_t nranks,
                        int64_t batch_local, int32_t elt, void* stream);

/* etb_a2a_scatter with a leading dimension per destination: block r lands as a rows_host[r] x batch_local matrix
 * whose columns are dst_ld_host[r] elements apart -- a row block INSIDE the owner's cotangent buffer, so the backward
 * exchange can run table group by table group while the owner already updates the groups that have arrived. */
int32_t etb_a2a_scatter_ld(void* const* dst_ptrs_host, const int64_t* dst_ld_host, const void* src, int64_t ld_src,
                           const int64_t* rows_host, const int64_t* row_off_host, int32_t nranks,
                           int64_t batch_local, int32_t elt, void* stream);

/* Stream-ordered barrier over peer memory (no collective library call): flag_ptrs_host[r] is rank r's flag array
 * (nranks zero-initialised uint32, allocated with etb_malloc by r and mapped here with etb_ipc_import; my own entry
 * is my local array).  One tiny kernel stores `epoch` into slot `rank` of every rank's array with release
 * semantics at system scope and spins until all slots of my own array have reached `epoch`.  Kernels enqueued before
 * it on `stream` are complete when it signals, so their peer stores are visible to every rank that passes the
 * barrier.  `epoch` must grow by one per barrier (wrap-around is handled).  Every rank must call it. */
int32_t etb_peer_barrier(void* const* flag_ptrs_host, int32_t rank, int32_t nranks, uint32_t epoch, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* EMBTAB_B200_H */
