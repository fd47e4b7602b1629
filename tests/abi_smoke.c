/*
 * abi_smoke.c -- the C ABI of libembtab_b200.so driven from plain C: no Python, no torch, nothing but
 * include/embtab_b200.h.  This is what a `ccall` from the reference's host language amounts to.
 *
 *   gcc -std=c99 -Iinclude tests/abi_smoke.c -Lembeddingtables.jl_b200/lib -lembtab_b200 \
 *       -Wl,-rpath,$PWD/embeddingtables.jl_b200/lib -lm -o abi_smoke && ./abi_smoke
 *
 * Small SimpleEmbedding / SplitEmbedding tables, 1-based int64 indices:
 *   etb_gather          O[:, j] = A[:, I[j]]                            (reference src/lookup.jl:51-102)
 *   etb_pooled_sum      O[:, j] = sum_i A[:, I[i, j]], in bag order     (reference src/lookup.jl:108-182)
 *   etb_maplookup       PreallocationStrategy form, prepend rows kept   (reference src/lookup.jl:316-371)
 *   etb_index + etb_sgd_update / etb_index_and_update                   (reference src/sparseupdate.jl:57-238)
 * Expected values are computed right here with the reference's order of operations (sequential sums,
 * accumulator from zero, row - eta * acc with separate roundings for these Dynamic tables) and compared bit for bit.
 * Exit code 0 = all equal.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "embtab_b200.h"

#define CHECK(call)                                                                          \
    do {                                                                                     \
        int32_t st_ = (call);                                                                \
        if (st_ != ETB_OK) {                                                                 \
            fprintf(stderr, "%s:%d: %s -> %d: %s\n", __FILE__, __LINE__, #call, st_, etb_last_error()); \
            exit(2);                                                                         \
        }                                                                                    \
    } while (0)

enum { DIM = 24, NROWS = 301, BAG = 5, BATCH = 64, SHARD = 64, PREPEND = 3, NT = 2 };

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd(void) {
    rng_state ^= rng_state >> 12; rng_state ^= rng_state << 25; rng_state ^= rng_state >> 27;
    return (uint32_t)((rng_state * 0x2545F4914F6CDD1Dull) >> 32);
}
static float rndf(void) { return (float)(rnd() >> 8) * (1.0f / 16777216.0f) - 0.5f; }

static void* to_device(const void* host, size_t bytes, void* stream) {
    void* d = NULL;
    CHECK(etb_malloc(&d, bytes));
    CHECK(etb_memcpy_h2d(d, host, bytes, stream));
    return d;
}

static int same(const char* what, const float* got, const float* want, size_t n) {
    if (memcmp(got, want, n * sizeof(float)) == 0) {
        printf("ok   %s\n", what);
        return 0;
    }
    for (size_t i = 0; i < n; ++i)
        if (memcmp(&got[i], &want[i], sizeof(float)) != 0) {
            printf("FAIL %s: element %zu is %.9g, expected %.9g\n", what, i, got[i], want[i]);
            break;
        }
    return 1;
}

int main(void) {
    int32_t ndev = 0;
    CHECK(etb_device_count(&ndev));
    if (ndev < 1) {
        fprintf(stderr, "abi_smoke: no CUDA device\n");
        return 3;
    }
    CHECK(etb_init(0));
    if (etb_version() != ETB_VERSION) {
        fprintf(stderr, "abi_smoke: header version %d, library version %d\n", ETB_VERSION, etb_version());
        return 2;
    }
    void* stream = NULL;
    CHECK(etb_stream_create(&stream));
    int bad = 0;

    /* ---- host data: two tables (column-major DIM x NROWS, one embedding row = DIM contiguous floats) */
    static float A[NT][NROWS * DIM];
    static int64_t I[NT][BAG * BATCH]; /* BAG x BATCH, column-major, 1-based */
    static float delta[(PREPEND + NT * DIM) * BATCH];
    for (int t = 0; t < NT; ++t) {
        for (int i = 0; i < NROWS * DIM; ++i) A[t][i] = rndf();
        for (int i = 0; i < BAG * BATCH; ++i) /* table 1: heavy duplicates, row 5 has > 128 occurrences (a "long" bucket) */
            I[t][i] = t == 0 ? 1 + (int64_t)(rnd() % NROWS) : (rnd() % 10 < 7 ? 5 : 1 + (int64_t)(rnd() % 17));
    }
    for (size_t i = 0; i < sizeof(delta) / sizeof(float); ++i) delta[i] = rndf();

    /* ---- device tables: table 0 = SimpleEmbedding, table 1 = SplitEmbedding in chunks of SHARD rows (last ragged) */
    etb_table tab[NT];
    memset(tab, 0, sizeof(tab));
    void* d_A0 = to_device(A[0], sizeof(A[0]), stream);
    tab[0].base = d_A0; tab[0].nrows = NROWS; tab[0].dim = DIM; tab[0].ld = DIM; tab[0].elt = ETB_F32;
    enum { NCHUNK = (NROWS + SHARD - 1) / SHARD };
    void* chunk_ptrs[NCHUNK];
    for (int c = 0; c < NCHUNK; ++c) {
        int rows = (c + 1) * SHARD <= NROWS ? SHARD : NROWS - c * SHARD;
        chunk_ptrs[c] = to_device(&A[1][(size_t)c * SHARD * DIM], (size_t)rows * DIM * sizeof(float), stream);
    }
    void* d_chunks = to_device(chunk_ptrs, sizeof(chunk_ptrs), stream);
    tab[1].chunks = (void* const*)d_chunks; tab[1].nrows = NROWS; tab[1].shard_rows = SHARD;
    tab[1].dim = DIM; tab[1].ld = DIM; tab[1].elt = ETB_F32;
    void* d_I[NT];
    for (int t = 0; t < NT; ++t) d_I[t] = to_device(I[t], sizeof(I[t]), stream);

    /* ---- K1 gather: the first BATCH indices of table 0 as a vector */
    {
        static float got[DIM * BATCH], want[DIM * BATCH];
        void* d_out = NULL;
        CHECK(etb_malloc(&d_out, sizeof(got)));
        CHECK(etb_gather(d_out, DIM, &tab[0], d_I[0], ETB_I64, BATCH, stream));
        CHECK(etb_memcpy_d2h(got, d_out, sizeof(got), stream));
        CHECK(etb_stream_sync(stream));
        for (int j = 0; j < BATCH; ++j) memcpy(&want[j * DIM], &A[0][(I[0][j] - 1) * DIM], DIM * sizeof(float));
        bad += same("etb_gather (SimpleEmbedding)", got, want, DIM * BATCH);
        CHECK(etb_free(d_out));
    }
    /* ---- K2 pooled sum on the split table: accumulator = first row, then + in bag order */
    static float pooled[NT][DIM * BATCH];
    for (int t = 0; t < NT; ++t)
        for (int j = 0; j < BATCH; ++j)
            for (int k = 0; k < DIM; ++k) {
                float acc = A[t][(I[t][j * BAG] - 1) * DIM + k];
                for (int i = 1; i < BAG; ++i) acc = acc + A[t][(I[t][j * BAG + i] - 1) * DIM + k];
                pooled[t][j * DIM + k] = acc;
            }
    {
        static float got[DIM * BATCH];
        void* d_out = NULL;
        CHECK(etb_malloc(&d_out, sizeof(got)));
        CHECK(etb_pooled_sum(d_out, DIM, &tab[1], d_I[1], ETB_I64, BAG, BATCH, BAG, stream));
        CHECK(etb_memcpy_d2h(got, d_out, sizeof(got), stream));
        CHECK(etb_stream_sync(stream));
        bad += same("etb_pooled_sum (SplitEmbedding)", got, pooled[1], DIM * BATCH);
        CHECK(etb_free(d_out));
    }
    /* ---- K3 fused ensemble lookup, PreallocationStrategy(PREPEND) layout; prepend rows must stay untouched */
    {
        enum { ROWS = PREPEND + NT * DIM };
        static float got[ROWS * BATCH], want[ROWS * BATCH];
        for (int i = 0; i < ROWS * BATCH; ++i) want[i] = got[i] = -7.0f;
        void* d_out = to_device(got, sizeof(got), stream);
        etb_lookup_item items[NT];
        memset(items, 0, sizeof(items));
        for (int t = 0; t < NT; ++t) {
            items[t].table = tab[t];
            items[t].idx = d_I[t];
            items[t].dst = (char*)d_out + (size_t)(PREPEND + t * DIM) * sizeof(float);
            items[t].ld_dst = ROWS; items[t].batch = BATCH; items[t].bag = BAG; items[t].ld_idx = BAG;
            items[t].idx_elt = ETB_I64;
            for (int j = 0; j < BATCH; ++j) memcpy(&want[j * ROWS + PREPEND + t * DIM], &pooled[t][j * DIM], DIM * sizeof(float));
        }
        CHECK(etb_maplookup(items, NT, stream));
        CHECK(etb_memcpy_d2h(got, d_out, sizeof(got), stream));
        CHECK(etb_stream_sync(stream));
        bad += same("etb_maplookup (PreallocationStrategy, prepend rows untouched)", got, want, ROWS * BATCH);
        CHECK(etb_free(d_out));
    }
    /* ---- K4 + K5: ensemble update!(Descent(eta)) with per-table row slices of one cotangent matrix */
    {
        enum { ROWS = PREPEND + NT * DIM };
        const double eta = 0.37;
        const float eta_f = (float)eta;
        void* d_delta = to_device(delta, sizeof(delta), stream);
        etb_update_item items[NT];
        memset(items, 0, sizeof(items));
        for (int t = 0; t < NT; ++t) {
            items[t].table = tab[t];
            items[t].delta = (char*)d_delta + (size_t)(PREPEND + t * DIM) * sizeof(float);
            items[t].ld_delta = ROWS; items[t].idx = d_I[t]; items[t].batch = BATCH; items[t].bag = BAG;
            items[t].ld_idx = BAG; items[t].idx_elt = ETB_I64; items[t].flags = 0; /* Dynamic tables: row - eta*acc */
        }
        size_t ws_bytes = 0;
        CHECK(etb_index_workspace_bytes(items, NT, &ws_bytes));
        void* ws = NULL;
        CHECK(etb_malloc(&ws, ws_bytes));
        etb_index_view view;
        CHECK(etb_index(ws, ws_bytes, items, NT, &view, stream));
        CHECK(etb_sgd_update(&view, items, NT, eta, 0, stream)); /* strict order: bit-identical to the reference */
        /* expected: per row, acc = 0 + members in occurrence order (column-major traversal), then row - eta*acc */
        static float want[NT][NROWS * DIM];
        for (int t = 0; t < NT; ++t) {
            memcpy(want[t], A[t], sizeof(A[t]));
            for (int row = 1; row <= NROWS; ++row)
                for (int k = 0; k < DIM; ++k) {
                    float acc = 0.0f;
                    int hit = 0;
                    for (int p = 0; p < BAG * BATCH; ++p)
                        if (I[t][p] == row) {
                            acc = acc + delta[(size_t)(p / BAG) * ROWS + PREPEND + t * DIM + k];
                            hit = 1;
                        }
                    if (hit) {
                        volatile float prod = eta_f * acc; /* two roundings, like the reference's generic kernel */
                        want[t][(row - 1) * DIM + k] = A[t][(row - 1) * DIM + k] - prod;
                    }
                }
        }
        static float got[NROWS * DIM];
        CHECK(etb_memcpy_d2h(got, d_A0, sizeof(got), stream));
        CHECK(etb_stream_sync(stream));
        bad += same("etb_index + etb_sgd_update (SimpleEmbedding)", got, want[0], NROWS * DIM);
        for (int c = 0; c < NCHUNK; ++c) {
            int rows = (c + 1) * SHARD <= NROWS ? SHARD : NROWS - c * SHARD;
            CHECK(etb_memcpy_d2h(&got[(size_t)c * SHARD * DIM], chunk_ptrs[c], (size_t)rows * DIM * sizeof(float), stream));
        }
        CHECK(etb_stream_sync(stream));
        bad += same("etb_index + etb_sgd_update (SplitEmbedding, heavy duplicates)", got, want[1], NROWS * DIM);
        /* a second step through the one-call form, on table 0 only */
        CHECK(etb_index_and_update(ws, ws_bytes, items, 1, eta, 0, stream));
        for (int row = 1; row <= NROWS; ++row)
            for (int k = 0; k < DIM; ++k) {
                float acc = 0.0f;
                int hit = 0;
                for (int p = 0; p < BAG * BATCH; ++p)
                    if (I[0][p] == row) { acc = acc + delta[(size_t)(p / BAG) * ROWS + PREPEND + k]; hit = 1; }
                if (hit) {
                    volatile float prod = eta_f * acc;
                    want[0][(row - 1) * DIM + k] = want[0][(row - 1) * DIM + k] - prod;
                }
            }
        CHECK(etb_memcpy_d2h(got, d_A0, sizeof(got), stream));
        CHECK(etb_stream_sync(stream));
        bad += same("etb_index_and_update (second step)", got, want[0], NROWS * DIM);
        CHECK(etb_free(ws));
        CHECK(etb_free(d_delta));
    }
    /* ---- argument validation comes back as a status, never a crash */
    if (etb_gather(NULL, DIM, NULL, d_I[0], ETB_I64, BATCH, stream) == ETB_OK) {
        printf("FAIL etb_gather accepted a null table\n");
        ++bad;
    } else {
        printf("ok   null table rejected: %s\n", etb_last_error());
    }
    for (int t = 0; t < NT; ++t) CHECK(etb_free(d_I[t]));
    for (int c = 0; c < NCHUNK; ++c) CHECK(etb_free(chunk_ptrs[c]));
    CHECK(etb_free(d_chunks));
    CHECK(etb_free(d_A0));
    CHECK(etb_stream_destroy(stream));
    printf(bad ? "abi_smoke: %d check(s) FAILED\n" : "abi_smoke: all checks passed\n", bad);
    return bad ? 1 : 0;
}
