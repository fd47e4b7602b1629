// Ceiling microbenchmark: what does B200 HBM deliver for the update kernel's access pattern?
//   rmw      : random 512-byte rows: read, modify, write back (u rows of a 13.3 GB array)
//   rmw+read : the same plus one more random 512-byte read per row (the delta row, from a 226 MB array)
//   read     : random 512-byte row reads only (the pooled kernel's pattern)
// Each warp handles rows in batches of U with all loads in flight before the first store.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>
#include <random>

template <int U, int MODE>
__global__ void __launch_bounds__(256) k(float4* __restrict__ table, const float4* __restrict__ delta,
                                         const uint32_t* __restrict__ rows, const uint32_t* __restrict__ cols, int64_t n,
                                         float* sink) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t i0 = warp * U;
    if (i0 >= n) return;
    float4 v[U], d[U];
    uint32_t r[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int64_t i = min(i0 + u, n - 1);
        r[u] = rows[i];
        v[u] = table[(int64_t)r[u] * 32 + lane];
        if (MODE == 1) d[u] = __ldg(&delta[(int64_t)cols[i] * 32 + lane]);
    }
    float acc = 0.f;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (MODE == 2) { acc += v[u].x + v[u].w; continue; }
        if (MODE == 1) { v[u].x = fmaf(-0.01f, d[u].x, v[u].x); v[u].y = fmaf(-0.01f, d[u].y, v[u].y); v[u].z = fmaf(-0.01f, d[u].z, v[u].z); v[u].w = fmaf(-0.01f, d[u].w, v[u].w); }
        else { v[u].x += 1.f; v[u].y += 1.f; v[u].z += 1.f; v[u].w += 1.f; }
        if (i0 + u < n) table[(int64_t)r[u] * 32 + lane] = v[u];
    }
    if (MODE == 2 && acc == 123.456f) *sink = acc;
}

template <int U, int MODE>
float run(float4* table, float4* delta, uint32_t* rows, uint32_t* cols, int64_t n, float* sink) {
    const int64_t warps = (n + U - 1) / U;
    const int grid = (int)((warps * 32 + 255) / 256);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9f;
    for (int it = 0; it < 6; ++it) {
        cudaEventRecord(a);
        k<U, MODE><<<grid, 256>>>(table, delta, rows, cols, n, sink);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (it > 0) best = std::min(best, ms);
    }
    return best;
}

int main() {
    const int64_t nrows = 26ll * 1000000, ncols = 26ll * 16384;   // 13.3 GB table, 218 MB delta
    const int64_t n = 26ll * 407964;                               // distinct rows updated per C2 step
    float4 *table, *delta; uint32_t *rows, *cols; float* sink;
    cudaMalloc(&table, nrows * 512); cudaMalloc(&delta, ncols * 512);
    cudaMalloc(&rows, n * 4); cudaMalloc(&cols, n * 4); cudaMalloc(&sink, 4);
    cudaMemset(table, 0, nrows * 512); cudaMemset(delta, 0, ncols * 512);
    // sorted distinct rows (the update visits rows in ascending order), random delta columns within a table
    std::mt19937_64 g(1);
    std::vector<uint32_t> hr(n), hc(n);
    for (int t = 0; t < 26; ++t) {
        std::vector<uint32_t> p(1000000);
        for (uint32_t i = 0; i < 1000000; ++i) p[i] = i;
        std::shuffle(p.begin(), p.end(), g);
        std::sort(p.begin(), p.begin() + 407964);
        for (int i = 0; i < 407964; ++i) { hr[t * 407964ll + i] = t * 1000000u + p[i]; hc[t * 407964ll + i] = t * 16384u + (uint32_t)(g() % 16384); }
    }
    cudaMemcpy(rows, hr.data(), n * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(cols, hc.data(), n * 4, cudaMemcpyHostToDevice);
    printf("{\"rows\": %lld", (long long)n);
#define R(U, MODE, NAME, BYTES) { float ms = run<U, MODE>(table, delta, rows, cols, n, sink); \
        printf(", \"%s_U%d\": {\"ms\": %.4f, \"gbs\": %.1f}", NAME, U, ms, (double)(BYTES) / ms / 1e6); }
    R(4, 0, "rmw", n * 1024.0) R(8, 0, "rmw", n * 1024.0) R(16, 0, "rmw", n * 1024.0)
    R(4, 1, "rmw_plus_delta", n * 1024.0 + ncols * 512.0) R(8, 1, "rmw_plus_delta", n * 1024.0 + ncols * 512.0)
    R(8, 2, "read", n * 512.0) R(16, 2, "read", n * 512.0)
    printf("}\n");
    return 0;
}
