// ubench_chain.cu -- what does ONE warp pay per member for a strictly ordered sum fed from shared memory?
// (the adding warp of long_strict_sliced_kernel).  nvcc -arch=sm_100a -O3 -o ubench_chain ubench_chain.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ROWS = 128, ITER = 400;
__device__ __forceinline__ float lds32(unsigned a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ float4 lds128(unsigned a) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v; }

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// SPIN = 1: the other warps wait on an mbarrier (try_wait loop) until the adding warp is done, like the producers of the
// kernel that wait for a stage to be handed back; SPIN = 2: they wait at a named hardware barrier instead
template <int MODE, int U, int SPIN = 0>
__global__ void k(float* out, long long* cyc) {
    __shared__ __align__(16) float s[ROWS * 32];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    for (int i = threadIdx.x; i < ROWS * 32; i += blockDim.x) s[i] = 1.0f + i * 1e-7f;
    __syncthreads();
    if (threadIdx.x >= 32) {
        if (SPIN == 1) {
            unsigned ok;
            do {
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
                             : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
            } while (!ok);
        } else if (SPIN == 2) {
            asm volatile("bar.sync 1, %0;" ::"r"((int)blockDim.x) : "memory");
        }
        return;
    }
    const int lane = threadIdx.x;
    float acc = 0.f;
    const unsigned base = (unsigned)__cvta_generic_to_shared(s) + lane * (MODE == 2 ? 16 : 4);
    const long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
        if constexpr (MODE == 0) {  // chain only
#pragma unroll
            for (int r = 0; r < ROWS; ++r) acc = acc + 1.5f * (it == -1 ? 0.f : 1.f) + 0.f * r;
        } else if constexpr (MODE == 1) {  // 4-byte loads, groups of U, double buffered (the kernel's loop)
            float va[U], vb[U];
#pragma unroll
            for (int u = 0; u < U; ++u) va[u] = lds32(base + u * 128);
            int r = 0;
            for (; r + 3 * U <= ROWS; r += 2 * U) {
#pragma unroll
                for (int u = 0; u < U; ++u) vb[u] = lds32(base + (r + U + u) * 128);
#pragma unroll
                for (int u = 0; u < U; ++u) acc = acc + va[u];
#pragma unroll
                for (int u = 0; u < U; ++u) va[u] = lds32(base + (r + 2 * U + u) * 128);
#pragma unroll
                for (int u = 0; u < U; ++u) acc = acc + vb[u];
            }
#pragma unroll
            for (int u = 0; u < U; ++u) acc = acc + va[u];
            r += U;
            for (; r < ROWS; ++r) acc = acc + lds32(base + r * 128);
        } else if constexpr (MODE == 2) {  // 16-byte loads of 4 members (transposed layout)
            float4 a0 = lds128(base), a1 = lds128(base + 512), b0, b1;
            int q = 0;
            for (; q + 6 <= ROWS / 4; q += 4) {
                b0 = lds128(base + (q + 2) * 512); b1 = lds128(base + (q + 3) * 512);
                acc += a0.x; acc += a0.y; acc += a0.z; acc += a0.w; acc += a1.x; acc += a1.y; acc += a1.z; acc += a1.w;
                a0 = lds128(base + (q + 4) * 512); a1 = lds128(base + (q + 5) * 512);
                acc += b0.x; acc += b0.y; acc += b0.z; acc += b0.w; acc += b1.x; acc += b1.y; acc += b1.z; acc += b1.w;
            }
            acc += a0.x; acc += a0.y; acc += a0.z; acc += a0.w; acc += a1.x; acc += a1.y; acc += a1.z; acc += a1.w;
        } else if constexpr (MODE == 3) {  // all loads of the batch first, then the chain (no overlap)
            float v[32];
            for (int r = 0; r < ROWS; r += 32) {
#pragma unroll
                for (int u = 0; u < 32; ++u) v[u] = lds32(base + (r + u) * 128);
#pragma unroll
                for (int u = 0; u < 32; ++u) acc = acc + v[u];
            }
        }
    }
    const long long t1 = clock64();
    if (lane == 0) *cyc = t1 - t0;
    out[lane] = acc;
    if (SPIN == 1 && lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
    if (SPIN == 2) asm volatile("bar.sync 1, %0;" ::"r"((int)blockDim.x) : "memory");
}

template <int MODE, int U, int SPIN = 0>
void run(const char* name, int nthreads, int members_per_iter, int nblocks = 1) {
    float* out; long long* cyc;
    cudaMalloc(&out, 128); cudaMalloc(&cyc, 8);
    k<MODE, U, SPIN><<<nblocks, nthreads>>>(out, cyc);
    k<MODE, U, SPIN><<<nblocks, nthreads>>>(out, cyc);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-46s %6.2f cycles/member (%s)\n", name, (double)h / ITER / members_per_iter, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    run<1, 16>("lds32 + fadd, U=16, 1 warp in the CTA", 32, ROWS);
    run<1, 16>("lds32 + fadd, U=16, 8 warps launched (7 exit)", 256, ROWS);
    run<1, 4>("lds32 + fadd, U=4", 32, ROWS);
    run<1, 8>("lds32 + fadd, U=8", 32, ROWS);
    run<1, 32>("lds32 + fadd, U=32", 32, ROWS);
    run<2, 1>("lds128 (4 members) + 4 fadd", 32, (ROWS / 4 - 2) / 4 * 4 * 4 + 8);
    run<3, 1>("32 x lds32, then 32 x fadd", 32, ROWS);
    run<1, 16, 1>("lds32 + fadd, 7 warps in mbarrier.try_wait", 256, ROWS);
    run<1, 16, 1>("... 3 such CTAs per SM (444 CTAs)", 256, ROWS, 444);
    run<1, 16, 2>("lds32 + fadd, 7 warps at bar.sync", 256, ROWS);
    run<1, 16, 2>("... 3 such CTAs per SM (444 CTAs)", 256, ROWS, 444);
    run<1, 16, 0>("lds32 + fadd, 444 CTAs of one warp", 32, ROWS, 444);
    return 0;
}
