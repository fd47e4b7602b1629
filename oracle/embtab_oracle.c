/*
 * embtab_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C restatement of the algorithm of darchr/EmbeddingTables.jl's hot path, used
 *   (1) by tests/ as the checker the CUDA path is compared against,
 *   (2) by __graft_entry__.smoke() as the checker, and
 *   (3) by bench.py's `cpu_baseline` / `--impl reference` legs as the timed CPU arm
 *       ("kind": "port" -- the reference itself is Julia and cannot run in this image).
 * Nothing under embeddingtables.jl_b200/ may import, link or call this file.
 *
 * Parity pinning: the reference holds golden vectors only for histogram!/index!
 * (test/misc.jl:33-110) and worked examples in README.md:32-73,113-160,190-232; all of them
 * are replayed against this file by tests/test_oracle_golden.py.  The floating-point
 * behaviour below follows the reference source line by line (order of additions, where the
 * accumulator starts, where a fused multiply-add is used); the one thing the reference's
 * own tests do not pin is FMA-vs-separate rounding in the update epilogue (they assert
 * isapprox with rtol 3.45e-4, test/update.jl:45,61,82) -- see DESIGN.md "Oracle".
 *
 * Every function cites the reference file:line it restates (paths relative to the
 * reference checkout).  Layout convention = Julia's: column-major, one embedding row is one
 * Julia column of `dim` contiguous elements; indices are 1-based int64.
 */
#define _GNU_SOURCE
#include <immintrin.h>
#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { ETBO_F32 = 0, ETBO_F64 = 1, ETBO_I32 = 2, ETBO_I64 = 3 };

typedef struct etbo_table {
    void* base;         /* SimpleEmbedding: parent matrix (src/simple.jl:50-51)            */
    void** chunks;      /* SplitEmbedding: vector of chunk matrices (src/split.jl:4)       */
    int64_t nrows;      /* size(table, 2)                                                   */
    int64_t shard_rows; /* SplitEmbedding matrixsize[2] (src/split.jl:9,52)                */
    int32_t dim;        /* featuresize                                                      */
    int32_t ld;         /* strides(A)[2] (src/EmbeddingTables.jl:83-85)                     */
    int32_t elt;
    int32_t is_static;  /* Static{N} vs Dynamic tag (src/EmbeddingTables.jl:60-63)          */
} etbo_table;

static inline size_t elt_size(int elt) { return (elt == ETBO_F32 || elt == ETBO_I32) ? 4 : 8; }

/* columnpointer(table, i): src/simple.jl:52-55 (Simple), src/split.jl:59-65,81-86 (Split:
 * `_divrem_index` on the 1-based index, then the inner matrix's columnpointer). */
static inline char* columnpointer(const etbo_table* t, int64_t i) {
    size_t stride = (size_t)t->ld * elt_size(t->elt);
    if (t->chunks) {
        int64_t z = i - 1;
        int64_t chunk = z / t->shard_rows; /* sdiv_int */
        int64_t col = z % t->shard_rows;   /* srem_int */
        return (char*)t->chunks[chunk] + (size_t)col * stride;
    }
    return (char*)t->base + (size_t)(i - 1) * stride;
}

static int g_has_avx512 = -1;
static int has_avx512(void) {
    if (g_has_avx512 < 0) g_has_avx512 = __builtin_cpu_supports("avx512f") ? 1 : 0;
    return g_has_avx512;
}
/* tests force the portable loops so both code paths are compared with each other */
void etbo_force_portable(int on) { g_has_avx512 = on ? 0 : (__builtin_cpu_supports("avx512f") ? 1 : 0); }
int etbo_uses_avx512(void) { return has_avx512(); }

/* ---------------------------------------------------------------------------------------
 * Benchmark support (bench.py's CPU arm): one worker thread per core, pinned -- the state
 * Polyester's `@batch per=thread` loops run in with JULIA_NUM_THREADS = cores and pinned
 * threads -- and a threaded fill for the 13 GB of synthetic tables.
 * ------------------------------------------------------------------------------------- */
static int g_pin_threads = 0;
void etbo_set_pinning(int on) { g_pin_threads = on; }

/* number of CPUs this process may run on (the honest `cores` of the baseline) */
int etbo_allowed_cpus(void) {
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) != 0) return 1;
    int n = CPU_COUNT(&set);
    return n > 0 ? n : 1;
}

static void pin_worker(int tid) {
    if (!g_pin_threads) return;
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) != 0) return;
    int n = CPU_COUNT(&set);
    if (n <= 0) return;
    int want = tid % n, seen = 0;
    for (int c = 0; c < CPU_SETSIZE; ++c) {
        if (!CPU_ISSET(c, &set)) continue;
        if (seen++ == want) {
            cpu_set_t one;
            CPU_ZERO(&one);
            CPU_SET(c, &one);
            pthread_setaffinity_np(pthread_self(), sizeof(one), &one);
            return;
        }
    }
}

typedef struct { float* dst; size_t n; uint64_t seed; int tid; } fill_job;
static void* fill_worker(void* p) {
    fill_job* j = (fill_job*)p;
    pin_worker(j->tid);
    uint64_t x = j->seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull * (uint64_t)(j->tid + 1);
    for (size_t i = 0; i < j->n; ++i) { /* xorshift64*: uniform [0, 1) with 24 random bits, like rand(Float32) */
        x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
        j->dst[i] = (float)((x * 0x2545F4914F6CDD1Dull) >> 40) * (1.0f / 16777216.0f);
    }
    return NULL;
}
void etbo_fill_uniform(float* dst, size_t n, uint64_t seed, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
    fill_job* jobs = (fill_job*)malloc(sizeof(fill_job) * nthreads);
    size_t per = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; ++t) {
        size_t lo = (size_t)t * per, hi = lo + per < n ? lo + per : n;
        jobs[t] = (fill_job){dst + (lo < n ? lo : n), lo < n ? hi - lo : 0, seed, t};
        pthread_create(&th[t], NULL, fill_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
}

/* ---------------------------------------------------------------------------------------
 * Non-reducing lookup.  lookup_generic!(dst, src, indices::AbstractVector) src/lookup.jl:51-67
 * and lookup_static! :70-87 -- both are an element-wise / whole-vector copy of column
 * indices[j] into destination column j.
 * ------------------------------------------------------------------------------------- */
void etbo_gather(void* dst, int64_t ld_dst, const etbo_table* t, const int64_t* idx, int64_t n) {
    size_t es = elt_size(t->elt), rowbytes = (size_t)t->dim * es;
    for (int64_t j = 0; j < n; ++j)
        memcpy((char*)dst + (size_t)j * (size_t)ld_dst * es, columnpointer(t, idx[j]), rowbytes);
}

/* ---------------------------------------------------------------------------------------
 * Reducing lookup.  lookup_generic!(O, A, I::AbstractMatrix) src/lookup.jl:108-132:
 * copy the first looked-up row into the output column, then `vO[k] += vA[k]` for bag
 * entries 2..bag, in order.  The static path (lookup_static_inner :134-147 +
 * lookup_static! :149-165) seeds a register tile with the first row and adds the others in
 * the same order, then does one (non-temporal) store: identical arithmetic per element.
 * ------------------------------------------------------------------------------------- */
#define DEF_POOLED(NAME, T)                                                                      \
    static void NAME(T* dst, int64_t ld_dst, const etbo_table* t, const int64_t* idx,            \
                     int64_t bag, int64_t c0, int64_t c1, int64_t ld_idx) {                     \
        const int dim = t->dim;                                                                  \
        for (int64_t j = c0; j < c1; ++j) {                                                      \
            T* o = dst + (size_t)j * (size_t)ld_dst;                                             \
            const int64_t* col = idx + (size_t)j * (size_t)ld_idx;                               \
            const T* a = (const T*)columnpointer(t, col[0]);                                     \
            for (int k = 0; k < dim; ++k) o[k] = a[k];                                           \
            for (int64_t i = 1; i < bag; ++i) {                                                  \
                a = (const T*)columnpointer(t, col[i]);                                          \
                for (int k = 0; k < dim; ++k) o[k] += a[k];                                      \
            }                                                                                    \
        }                                                                                        \
    }
DEF_POOLED(pooled_f32, float)
DEF_POOLED(pooled_f64, double)
/* Julia integer `+` wraps; unsigned arithmetic gives the same bits without C UB */
DEF_POOLED(pooled_i32, uint32_t)
DEF_POOLED(pooled_i64, uint64_t)

/* The reference's register-tile kernel as generated for an AVX-512 host: K = dim/16 tiles of
 * Vec{16,Float32} (src/simd.jl:1-29, src/lookup.jl:176-178), sequential accumulate
 * (src/lookup.jl:139-146), aligned non-temporal store (src/simd.jl:31-45, Val(true) at
 * src/lookup.jl:161).  Same numbers as pooled_f32; it exists so the CPU baseline is timed
 * with the reference's instruction mix, not a scalar loop. */
__attribute__((target("avx512f"))) static void pooled_f32_avx512(float* dst, int64_t ld_dst,
                                                                 const etbo_table* t,
                                                                 const int64_t* idx, int64_t bag,
                                                                 int64_t c0, int64_t c1,
                                                                 int64_t ld_idx) {
    const int K = t->dim / 16;
    __m512 acc[16]; /* dim*4 <= MAX_ACCUMULATOR_SIZE = 1024 B  ->  K <= 16 (src/lookup.jl:30-32) */
    for (int64_t j = c0; j < c1; ++j) {
        float* o = dst + (size_t)j * (size_t)ld_dst;
        const int64_t* col = idx + (size_t)j * (size_t)ld_idx;
        const float* a = (const float*)columnpointer(t, col[0]);
        for (int k = 0; k < K; ++k) acc[k] = _mm512_loadu_ps(a + 16 * k);
        for (int64_t i = 1; i < bag; ++i) {
            a = (const float*)columnpointer(t, col[i]);
            for (int k = 0; k < K; ++k) acc[k] = _mm512_add_ps(acc[k], _mm512_loadu_ps(a + 16 * k));
        }
        if (((uintptr_t)o & 63) == 0)
            for (int k = 0; k < K; ++k) _mm512_stream_ps(o + 16 * k, acc[k]);
        else
            for (int k = 0; k < K; ++k) _mm512_storeu_ps(o + 16 * k, acc[k]);
    }
    _mm_sfence(); /* sfence(), src/utils.jl:16-22, called at src/lookup.jl:163 */
}

/* lookup!(dst, src, I::AbstractMatrix) dispatch, src/lookup.jl:167-182: the static tile kernel
 * when the table is Static{N} and N*sizeof(T) <= 1024, else the generic loop.  Deviation kept
 * on purpose: the reference's tile kernel silently drops the last N % 16
 * features; here a Static table whose N is not a multiple of 16 uses the generic loop, i.e. the
 * full-N sum README.md:22-25 defines. */
static void pooled_range(void* dst, int64_t ld_dst, const etbo_table* t, const int64_t* idx,
                         int64_t bag, int64_t c0, int64_t c1, int64_t ld_idx) {
    switch (t->elt) {
        case ETBO_F32:
            if (has_avx512() && t->is_static && t->dim % 16 == 0 && t->dim * 4 <= 1024)
                pooled_f32_avx512((float*)dst, ld_dst, t, idx, bag, c0, c1, ld_idx);
            else
                pooled_f32((float*)dst, ld_dst, t, idx, bag, c0, c1, ld_idx);
            break;
        case ETBO_F64: pooled_f64((double*)dst, ld_dst, t, idx, bag, c0, c1, ld_idx); break;
        case ETBO_I32: pooled_i32((uint32_t*)dst, ld_dst, t, idx, bag, c0, c1, ld_idx); break;
        default: pooled_i64((uint64_t*)dst, ld_dst, t, idx, bag, c0, c1, ld_idx); break;
    }
}

void etbo_pooled_sum(void* dst, int64_t ld_dst, const etbo_table* t, const int64_t* idx,
                     int64_t bag, int64_t batch, int64_t ld_idx) {
    pooled_range(dst, ld_dst, t, idx, bag, 0, batch, ld_idx);
}

/* ---------------------------------------------------------------------------------------
 * maplookup! -- src/lookup.jl:233-241 (DefaultStrategy: serial map), :263-276
 * (SimpleParallelStrategy: Polyester @batch per=thread over tables = static contiguous split),
 * :316-371 (PreallocationStrategy: worksize_div batch chunks x tables behind an atomic
 * counter; k -> (j, i) = _divrem_index(k, ntables); chunk j covers columns
 * (j-1)*worksize+1 : min(j*worksize, batch) of table i).
 * ------------------------------------------------------------------------------------- */
typedef struct etbo_lookup_item {
    etbo_table table;
    const int64_t* idx;
    void* dst;
    int64_t ld_dst;
    int64_t batch;
    int64_t bag; /* 0 = non-reducing */
    int64_t ld_idx;
} etbo_lookup_item;

static void lookup_cols(const etbo_lookup_item* it, int64_t c0, int64_t c1) {
    size_t es = elt_size(it->table.elt);
    if (c1 <= c0) return;
    if (it->bag == 0) {
        etbo_gather((char*)it->dst + (size_t)c0 * (size_t)it->ld_dst * es, it->ld_dst, &it->table,
                    it->idx + c0, c1 - c0);
    } else {
        pooled_range(it->dst, it->ld_dst, &it->table, it->idx, it->bag, c0, c1, it->ld_idx);
    }
}

typedef struct {
    const etbo_lookup_item* items;
    int n_items, tid, nthreads, worksize_div, strategy;
    atomic_long* counter;
} map_job;

static void* map_worker(void* p) {
    map_job* job = (map_job*)p;
    if (job->nthreads > 1) pin_worker(job->tid);
    if (job->strategy == 1) { /* static split of tables across threads */
        int per = (job->n_items + job->nthreads - 1) / job->nthreads;
        int lo = job->tid * per, hi = lo + per < job->n_items ? lo + per : job->n_items;
        for (int i = lo; i < hi; ++i) lookup_cols(&job->items[i], 0, job->items[i].batch);
    } else { /* dynamic queue */
        long len = (long)job->worksize_div * job->n_items;
        for (;;) {
            long k = atomic_fetch_add(job->counter, 1); /* 1-based like Threads.atomic_add! */
            if (k > len) break;
            long j = (k - 1) / job->n_items + 1, i = (k - 1) % job->n_items; /* _divrem_index */
            const etbo_lookup_item* it = &job->items[i];
            int64_t worksize = 1 + (it->batch - 1) / job->worksize_div; /* cdiv, :301-302 */
            int64_t start = (j - 1) * worksize, stop = j * worksize < it->batch ? j * worksize : it->batch;
            lookup_cols(it, start, stop);
        }
    }
    return NULL;
}

/* strategy: 0 Default, 1 SimpleParallel, 2 Preallocation */
void etbo_maplookup(const etbo_lookup_item* items, int n_items, int strategy, int nthreads,
                    int worksize_div) {
    if (strategy == 0 || nthreads <= 1) {
        if (strategy == 2) { /* same work-item order as the queue, on one thread */
            atomic_long c = 1;
            map_job job = {items, n_items, 0, 1, worksize_div, 2, &c};
            map_worker(&job);
        } else {
            for (int i = 0; i < n_items; ++i) lookup_cols(&items[i], 0, items[i].batch);
        }
        return;
    }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
    map_job* jobs = (map_job*)malloc(sizeof(map_job) * nthreads);
    atomic_long counter = 1;
    for (int t = 0; t < nthreads; ++t) {
        jobs[t] = (map_job){items, n_items, t, nthreads, worksize_div, strategy, &counter};
        pthread_create(&th[t], NULL, map_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
}

/* ---------------------------------------------------------------------------------------
 * Indexer.  index!(I, A, maxindex) src/utils.jl:306-314 = histogram! + prefixsum! + remap!.
 *
 * Traversal order of A is `columns(A)` (src/utils.jl:73-81): a vector yields (i, A[i]); a
 * matrix is walked column-major and yields (column, A[row, column]) -- i.e. flat position p
 * (0-based) belongs to delta column p / bag (+1).
 *
 * Dense variant (DenseIndexer, src/utils.jl:154-167, 190-239, 259-272): histogram is an array
 * of (order, count) over 1..maxindex.  Sparse variant (SparseIndexer: a Dictionaries.jl
 * insertion-ordered hash, src/utils.jl:136-152, 170-188, 242-257): restated with an
 * open-addressing table that records first-seen order, which is all the algorithm uses of it.
 *
 * Outputs (1-based, like the Julia structs): cum_col[nnz+1], cum_off[nnz+1] =
 * Vector{ColOffset} incl. the (0, n+1) terminator; map[n]; hist_order/hist_count (optional,
 * dense only, for the known-answer test of histogram!, test/misc.jl:33-72).
 * Returns nnz.
 * ------------------------------------------------------------------------------------- */
typedef struct { int64_t order, count; } order_count;

int64_t etbo_histogram_dense(const int64_t* A, int64_t n, int64_t maxindex, int64_t* order_out,
                             int64_t* count_out) {
    /* shallow_empty! (:364-368) then unsafe_histogram! (:393-406) */
    for (int64_t i = 0; i < maxindex; ++i) order_out[i] = count_out[i] = 0;
    int64_t order = 0;
    for (int64_t p = 0; p < n; ++p) {
        int64_t a = A[p] - 1;
        if (order_out[a] == 0) order_out[a] = ++order;
        count_out[a] += 1;
    }
    return order;
}

int64_t etbo_index_dense(const int64_t* A, int64_t n, int64_t bag, int64_t maxindex,
                         int64_t* cum_col, int64_t* cum_off, int64_t* map) {
    order_count* h = (order_count*)calloc((size_t)maxindex, sizeof(order_count));
    int64_t nnz = 0;
    for (int64_t p = 0; p < n; ++p) { /* histogram!, :393-406 */
        order_count* e = &h[A[p] - 1];
        if (e->order == 0) e->order = ++nnz;
        e->count += 1;
    }
    /* prefixsum! dense, :429-478: loop 1 places (key, count) at its first-seen slot ... */
    for (int64_t k = 0; k < maxindex; ++k)
        if (h[k].order) { cum_col[h[k].order - 1] = k + 1; cum_off[h[k].order - 1] = h[k].count; }
    /* ... loop 2 turns counts into 1-based offsets, then the terminator */
    int64_t next = 1;
    for (int64_t i = 0; i < nnz; ++i) { int64_t c = cum_off[i]; cum_off[i] = next; next += c; }
    cum_col[nnz] = 0; cum_off[nnz] = next;
    /* remap!, :498-511: map[next_bucket_start - remaining] = delta column; remaining-- */
    if (bag < 1) bag = 1;
    for (int64_t p = 0; p < n; ++p) {
        order_count* e = &h[A[p] - 1];
        map[cum_off[e->order] - e->count - 1] = p / bag + 1;
        e->count -= 1;
    }
    free(h);
    return nnz;
}

static inline uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

int64_t etbo_index_sparse(const int64_t* A, int64_t n, int64_t bag, int64_t maxindex,
                          int64_t* cum_col, int64_t* cum_off, int64_t* map) {
    (void)maxindex;
    size_t cap = 16;
    while (cap < (size_t)n * 2 + 2) cap <<= 1;
    int64_t* keys = (int64_t*)calloc(cap, sizeof(int64_t)); /* 0 = empty (indices are >= 1) */
    order_count* vals = (order_count*)calloc(cap, sizeof(order_count));
    int64_t nnz = 0;
#define SLOT_OF(key, s)                                                    \
    do {                                                                   \
        s = mix64((uint64_t)(key)) & (cap - 1);                            \
        while (keys[s] != 0 && keys[s] != (key)) s = (s + 1) & (cap - 1);  \
    } while (0)
    for (int64_t p = 0; p < n; ++p) { /* unsafe_histogram! on a Dictionary, :375-391 */
        size_t s; SLOT_OF(A[p], s);
        if (keys[s] == 0) { keys[s] = A[p]; vals[s].order = ++nnz; vals[s].count = 1; }
        else vals[s].count += 1;
    }
    /* prefixsum! on a Dictionary, :409-427: pairs() iterates in insertion (= first-seen) order */
    for (size_t s = 0; s < cap; ++s)
        if (keys[s]) { cum_col[vals[s].order - 1] = keys[s]; cum_off[vals[s].order - 1] = vals[s].count; }
    int64_t next = 1;
    for (int64_t i = 0; i < nnz; ++i) { int64_t c = cum_off[i]; cum_off[i] = next; next += c; }
    cum_col[nnz] = 0; cum_off[nnz] = next;
    if (bag < 1) bag = 1;
    for (int64_t p = 0; p < n; ++p) { /* remap!, :481-496 */
        size_t s; SLOT_OF(A[p], s);
        map[cum_off[vals[s].order] - vals[s].count - 1] = p / bag + 1;
        vals[s].count -= 1;
    }
#undef SLOT_OF
    free(keys);
    free(vals);
    return nnz;
}

/* IndexerView(I, num_splits, this_split), src/utils.jl:325-333: range over `cumulative`
 * (length nnz+1 incl. terminator).  Returns 1-based [start, stop]; entries processed by
 * update! are start .. stop-1 (the loop is over length(cumulative)-1, sparseupdate.jl:69,110). */
void etbo_indexer_view(int64_t cum_len, int64_t num_splits, int64_t this_split, int64_t* start,
                       int64_t* stop) {
    int64_t split = 1 + (cum_len - 1) / num_splits; /* cdiv */
    *start = (this_split - 1) * split + 1;
    int64_t s = this_split * split + 1;
    *stop = s < cum_len ? s : cum_len;
}

/* ---------------------------------------------------------------------------------------
 * update!(table, update, indexer, alpha).
 *
 * Specialised kernel, _update_specialized_impl! src/sparseupdate.jl:97-129 -- chosen by the
 * @generated dispatch :131-154 when the table is Static{N} and N*sizeof(T) <= 512; simdtype
 * (src/simd.jl:5-12) restricts it to Float32 with N % 16 == 0:
 *     accum = zero; accum += delta[:, map[i]] for i = start..stop (in order);
 *     row   = muladd(-alpha, accum, row)          -> one fused multiply-add per element
 * Generic kernel, _update_generic_impl! :57-95:
 *     scratch = 0; scratch[k] += delta[k, map[i]];  row = row - alpha*scratch
 * (LoopVectorization.vmap(nt)!(f, ...) with f(x,y) = x - alpha*y, :88-93.  LoopVectorization
 * 0.12.118 is not vendored; whether it contracts x - alpha*y into an FMA is not pinned by any
 * reference test.  This oracle evaluates it with two roundings, the literal reading.)
 *
 * entries [e0, e1) are 0-based positions in cum_col/cum_off (an IndexerView range).
 * ------------------------------------------------------------------------------------- */
static int update_is_specialized(const etbo_table* t) {
    return t->is_static && t->elt == ETBO_F32 && t->dim % 16 == 0 && t->dim * 4 <= 512;
}

__attribute__((target("avx512f,fma"))) static void update_f32_avx512(
    const etbo_table* t, const float* delta, int64_t ld_delta, const int64_t* cum_col,
    const int64_t* cum_off, const int64_t* map, int64_t e0, int64_t e1, float alpha0) {
    const int K = t->dim / 16;
    const __m512 alpha = _mm512_set1_ps(-alpha0); /* alpha = -convert(T, alpha0), :108 */
    __m512 acc[8];
    for (int64_t e = e0; e < e1; ++e) {
        int64_t start = cum_off[e], stop = cum_off[e + 1] - 1;
        for (int k = 0; k < K; ++k) acc[k] = _mm512_setzero_ps();
        for (int64_t i = start; i <= stop; ++i) {
            const float* g = delta + (size_t)(map[i - 1] - 1) * (size_t)ld_delta;
            for (int k = 0; k < K; ++k) acc[k] = _mm512_add_ps(acc[k], _mm512_loadu_ps(g + 16 * k));
        }
        float* row = (float*)columnpointer(t, cum_col[e]);
        if (((uintptr_t)row & 63) == 0)
            for (int k = 0; k < K; ++k)
                _mm512_stream_ps(row + 16 * k, _mm512_fmadd_ps(alpha, acc[k], _mm512_load_ps(row + 16 * k)));
        else
            for (int k = 0; k < K; ++k)
                _mm512_storeu_ps(row + 16 * k, _mm512_fmadd_ps(alpha, acc[k], _mm512_loadu_ps(row + 16 * k)));
    }
}

void etbo_update(const etbo_table* t, const void* delta, int64_t ld_delta, const int64_t* cum_col,
                 const int64_t* cum_off, const int64_t* map, int64_t e0, int64_t e1, double eta) {
    const int dim = t->dim;
    if (t->elt == ETBO_F32) {
        const float alpha0 = (float)eta; /* convert(eltype(table), opt.eta), :173 */
        const float* d = (const float*)delta;
        if (update_is_specialized(t)) {
            if (has_avx512()) {
                update_f32_avx512(t, d, ld_delta, cum_col, cum_off, map, e0, e1, alpha0);
                return;
            }
            float* acc = (float*)malloc(sizeof(float) * dim);
            const float alpha = -alpha0;
            for (int64_t e = e0; e < e1; ++e) {
                int64_t start = cum_off[e], stop = cum_off[e + 1] - 1;
                for (int k = 0; k < dim; ++k) acc[k] = 0.0f;
                for (int64_t i = start; i <= stop; ++i) {
                    const float* g = d + (size_t)(map[i - 1] - 1) * (size_t)ld_delta;
                    for (int k = 0; k < dim; ++k) acc[k] += g[k];
                }
                float* row = (float*)columnpointer(t, cum_col[e]);
                for (int k = 0; k < dim; ++k) row[k] = fmaf(alpha, acc[k], row[k]);
            }
            free(acc);
            return;
        }
        float* scratch = (float*)malloc(sizeof(float) * dim);
        for (int64_t e = e0; e < e1; ++e) {
            int64_t start = cum_off[e], stop = cum_off[e + 1] - 1;
            for (int k = 0; k < dim; ++k) scratch[k] = 0.0f; /* zero!(scratchspace), :72 */
            for (int64_t i = start; i <= stop; ++i) {
                const float* g = d + (size_t)(map[i - 1] - 1) * (size_t)ld_delta;
                for (int k = 0; k < dim; ++k) scratch[k] += g[k];
            }
            float* row = (float*)columnpointer(t, cum_col[e]);
            for (int k = 0; k < dim; ++k) {
                volatile float prod = alpha0 * scratch[k]; /* two roundings, no contraction */
                row[k] = row[k] - prod;
            }
        }
        free(scratch);
        return;
    }
    /* Float64 tables always take the generic kernel (simdtype is Float32-only) */
    const double alpha0 = eta;
    const double* d = (const double*)delta;
    double* scratch = (double*)malloc(sizeof(double) * dim);
    for (int64_t e = e0; e < e1; ++e) {
        int64_t start = cum_off[e], stop = cum_off[e + 1] - 1;
        for (int k = 0; k < dim; ++k) scratch[k] = 0.0;
        for (int64_t i = start; i <= stop; ++i) {
            const double* g = d + (size_t)(map[i - 1] - 1) * (size_t)ld_delta;
            for (int k = 0; k < dim; ++k) scratch[k] += g[k];
        }
        double* row = (double*)columnpointer(t, cum_col[e]);
        for (int k = 0; k < dim; ++k) {
            volatile double prod = alpha0 * scratch[k];
            row[k] = row[k] - prod;
        }
    }
    free(scratch);
}

/* ---------------------------------------------------------------------------------------
 * Ensemble update!, src/sparseupdate.jl:199-238: phase 1 index every table (@batch over
 * tables, :211-213); phase 2 a dynamic queue over num_splits x ntables work items, k ->
 * (i, j) = _divrem_index(k, num_splits): table i, IndexerView split j (:217-237).
 * Scratch for the Indexer outputs is allocated here (the reference reuses caller-owned
 * Indexers; allocation is outside what the reference times either, see bench.py).
 * ------------------------------------------------------------------------------------- */
typedef struct etbo_update_item {
    etbo_table table;
    const void* delta;
    int64_t ld_delta;
    const int64_t* idx;
    int64_t batch;
    int64_t bag; /* 0 = vector */
    int64_t ld_idx; /* must equal bag (indices are traversed flat, utils.jl:76-81) */
    /* caller-owned Indexer storage: cum_col/cum_off have n+1 slots, map has n */
    int64_t* cum_col;
    int64_t* cum_off;
    int64_t* map;
    int64_t nnz; /* out */
} etbo_update_item;

typedef struct {
    etbo_update_item* items;
    int n_items, tid, nthreads, num_splits, phase, dense;
    double eta;
    atomic_long* counter;
} upd_job;

static void* upd_worker(void* p) {
    upd_job* job = (upd_job*)p;
    if (job->nthreads > 1) pin_worker(job->tid);
    if (job->phase == 1) {
        int per = (job->n_items + job->nthreads - 1) / job->nthreads;
        int lo = job->tid * per, hi = lo + per < job->n_items ? lo + per : job->n_items;
        for (int i = lo; i < hi; ++i) {
            etbo_update_item* it = &job->items[i];
            int64_t n = it->batch * (it->bag ? it->bag : 1);
            it->nnz = job->dense ? etbo_index_dense(it->idx, n, it->bag, it->table.nrows, it->cum_col,
                                                    it->cum_off, it->map)
                                 : etbo_index_sparse(it->idx, n, it->bag, it->table.nrows, it->cum_col,
                                                     it->cum_off, it->map);
        }
    } else {
        long len = (long)job->num_splits * job->n_items;
        for (;;) {
            long k = atomic_fetch_add(job->counter, 1);
            if (k > len) break;
            long i = (k - 1) / job->num_splits, j = (k - 1) % job->num_splits + 1;
            etbo_update_item* it = &job->items[i];
            int64_t start, stop;
            etbo_indexer_view(it->nnz + 1, job->num_splits, j, &start, &stop);
            if (stop > start)
                etbo_update(&it->table, it->delta, it->ld_delta, it->cum_col, it->cum_off, it->map,
                            start - 1, stop - 1, job->eta);
        }
    }
    return NULL;
}

void etbo_update_ensemble(etbo_update_item* items, int n_items, double eta, int num_splits,
                          int nthreads, int dense_indexer) {
    if (nthreads < 1) nthreads = 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
    upd_job* jobs = (upd_job*)malloc(sizeof(upd_job) * nthreads);
    atomic_long counter = 1;
    for (int phase = 1; phase <= 2; ++phase) {
        for (int t = 0; t < nthreads; ++t) {
            jobs[t] = (upd_job){items, n_items, t, nthreads, num_splits, phase, dense_indexer, eta, &counter};
            if (nthreads == 1) upd_worker(&jobs[t]);
            else pthread_create(&th[t], NULL, upd_worker, &jobs[t]);
        }
        if (nthreads > 1)
            for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    }
    free(th);
    free(jobs);
}

/* ---------------------------------------------------------------------------------------
 * uncompress(x, dstcols), src/sparseupdate.jl:16-32: dense gradient.  For each delta column
 * (in order) and each index of that column (in order): dst[:, c] .+= delta[:, column].
 * dst must be zeroed by the caller (the reference allocates zeros, :22).
 * ------------------------------------------------------------------------------------- */
void etbo_uncompress(void* dst, int64_t ld_dst, int32_t dim, int32_t elt, const void* delta,
                     int64_t ld_delta, const int64_t* idx, int64_t bag, int64_t batch, int64_t ld_idx) {
    int64_t per = bag ? bag : 1;
    for (int64_t j = 0; j < batch; ++j)
        for (int64_t i = 0; i < per; ++i) {
            int64_t c = bag ? idx[(size_t)j * (size_t)ld_idx + i] : idx[j];
            if (elt == ETBO_F32) {
                float* o = (float*)dst + (size_t)(c - 1) * (size_t)ld_dst;
                const float* g = (const float*)delta + (size_t)j * (size_t)ld_delta;
                for (int k = 0; k < dim; ++k) o[k] += g[k];
            } else {
                double* o = (double*)dst + (size_t)(c - 1) * (size_t)ld_dst;
                const double* g = (const double*)delta + (size_t)j * (size_t)ld_delta;
                for (int k = 0; k < dim; ++k) o[k] += g[k];
            }
        }
}
