// ubench_rows.cu -- the HBM ceiling of the update's access pattern as a function of the ROW LENGTH: random (sorted,
// distinct) rows of RB bytes read, modified and written back, plus one random RB-byte delta row read per row.
// A group of G = RB / 16 lanes owns a row (G = 32 with 20 active lanes for 320 bytes); every warp keeps U row batches
// (32 / G rows each) in flight before the first store.  Answers: is 0.4 - 0.6 of peak at dim 16 / 32 / 80 the kernel
// or the memory system?      nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_rows ubench_rows.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <algorithm>
#include <random>
#include <vector>

template <int G, int ACTIVE, int U>
__global__ void __launch_bounds__(256) k(float4* __restrict__ table, const float4* __restrict__ delta,
                                         const uint32_t* __restrict__ rows, const uint32_t* __restrict__ cols, int64_t n) {
    constexpr int RPW = 32 / G;  // rows per warp and batch
    const int lane = threadIdx.x & 31, gl = lane % G, grp = lane / G;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t i0 = warp * (U * RPW);
    if (i0 >= n) return;
    float4 v[U], d[U];
    int64_t a[U];
    bool on[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + u * RPW + grp;
        on[u] = i < n && gl < ACTIVE;
        const int64_t ii = min(i, n - 1);
        a[u] = (int64_t)rows[ii] * ACTIVE + gl;
        if (on[u]) {
            v[u] = table[a[u]];
            d[u] = __ldg(&delta[(int64_t)cols[ii] * ACTIVE + gl]);
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (on[u]) {
            v[u].x = fmaf(-0.01f, d[u].x, v[u].x); v[u].y = fmaf(-0.01f, d[u].y, v[u].y);
            v[u].z = fmaf(-0.01f, d[u].z, v[u].z); v[u].w = fmaf(-0.01f, d[u].w, v[u].w);
            table[a[u]] = v[u];
        }
    }
}

template <int G, int ACTIVE, int U>
void run(float4* table, float4* delta, uint32_t* rows, uint32_t* cols, int64_t n, int64_t ncols, double peak) {
    constexpr int RPW = 32 / G;
    const int64_t warps = (n + U * RPW - 1) / (U * RPW);
    const int grid = (int)((warps * 32 + 255) / 256);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9f;
    for (int it = 0; it < 5; ++it) {
        cudaEventRecord(a);
        k<G, ACTIVE, U><<<grid, 256>>>(table, delta, rows, cols, n);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (it > 0) best = std::min(best, ms);
    }
    const double rb = ACTIVE * 16.0, bytes = n * 2 * rb + ncols * rb + n * 8.0;
    printf("{\"row_bytes\": %d, \"dim_f32\": %d, \"U\": %d, \"rows_in_flight_per_warp\": %d, \"ms\": %.4f, \"gbs\": %.1f, \"frac_of_measured_peak\": %.3f, \"err\": \"%s\"}\n",
           (int)rb, (int)rb / 4, U, U * RPW, best, bytes / best / 1e6, bytes / best / 1e6 / peak, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv) {
    const double peak = argc > 1 ? atof(argv[1]) : 6546.9;
    const int T = 26;
    const int64_t nrows_t = 1000000, batch = 16384, u_t = 407964;  // C2's shape with the row length varied
    const int64_t n = T * u_t, ncols = T * batch;
    float4 *table, *delta; uint32_t *rows, *cols;
    cudaMalloc(&table, T * nrows_t * 512); cudaMalloc(&delta, ncols * 512);
    cudaMalloc(&rows, n * 4); cudaMalloc(&cols, n * 4);
    cudaMemset(table, 0, T * nrows_t * 512); cudaMemset(delta, 0, ncols * 512);
    std::mt19937_64 g(1);
    std::vector<uint32_t> hr(n), hc(n);
    for (int t = 0; t < T; ++t) {
        std::vector<uint32_t> p(nrows_t);
        for (uint32_t i = 0; i < nrows_t; ++i) p[i] = i;
        std::shuffle(p.begin(), p.end(), g);
        std::sort(p.begin(), p.begin() + u_t);
        for (int i = 0; i < u_t; ++i) { hr[t * u_t + i] = t * (uint32_t)nrows_t + p[i]; hc[t * u_t + i] = t * (uint32_t)batch + (uint32_t)(g() % batch); }
    }
    cudaMemcpy(rows, hr.data(), n * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(cols, hc.data(), n * 4, cudaMemcpyHostToDevice);
    run<4, 4, 4>(table, delta, rows, cols, n, ncols, peak);   run<4, 4, 8>(table, delta, rows, cols, n, ncols, peak);
    run<8, 8, 4>(table, delta, rows, cols, n, ncols, peak);   run<8, 8, 8>(table, delta, rows, cols, n, ncols, peak);
    run<16, 16, 4>(table, delta, rows, cols, n, ncols, peak); run<16, 16, 8>(table, delta, rows, cols, n, ncols, peak);
    run<32, 20, 4>(table, delta, rows, cols, n, ncols, peak); run<32, 20, 8>(table, delta, rows, cols, n, ncols, peak);
    run<32, 32, 4>(table, delta, rows, cols, n, ncols, peak); run<32, 32, 8>(table, delta, rows, cols, n, ncols, peak);
    return 0;
}
