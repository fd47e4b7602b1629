// etb_common.cuh -- shared host/device helpers of libembtab_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/embtab_b200.h"

namespace etb {

// ---------------------------------------------------------------- errors / bookkeeping
char* error_buffer();  // thread-local, 512 bytes
int32_t fail(int32_t status, const char* fmt, ...);
int32_t& launch_counter();  // thread-local: kernels launched by the current API call

#define ETB_CUDA(expr)                                                                     \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess)                                                             \
            return ::etb::fail(ETB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,               \
                               cudaGetErrorString(_e), __FILE__, __LINE__);                \
    } while (0)

#define ETB_REQUIRE(cond, ...)                                                             \
    do {                                                                                   \
        if (!(cond)) return ::etb::fail(ETB_ERR_INVALID, __VA_ARGS__);                     \
    } while (0)

// after a kernel launch: count it and surface launch-configuration errors
#define ETB_LAUNCHED()                                                                     \
    do {                                                                                   \
        ++::etb::launch_counter();                                                         \
        ETB_CUDA(cudaGetLastError());                                                      \
    } while (0)

// One NVTX range per API call (domain "embtab"): `nsys profile --trace=cuda,nvtx` shows every etb_* call with the
// kernels it enqueued underneath.  Header-only NVTX v3: without a profiler attached a range is one untaken branch.
struct ApiRange {
    explicit ApiRange(const char* name) { nvtxRangePushA(name); }
    ~ApiRange() { nvtxRangePop(); }
};
#define ETB_API_RANGE() ::etb::ApiRange etb_api_range_(__func__)

inline size_t elt_bytes(int32_t elt) {
    return (elt == ETB_F16 || elt == ETB_BF16) ? 2 : ((elt == ETB_F32 || elt == ETB_I32) ? 4 : 8);
}
inline bool elt_valid(int32_t elt) { return elt >= ETB_F32 && elt <= ETB_BF16; }
inline bool elt_is_float(int32_t elt) { return elt == ETB_F32 || elt == ETB_F64 || elt == ETB_F16 || elt == ETB_BF16; }
inline bool idx_elt_valid(int32_t e) { return e == ETB_I32 || e == ETB_I64; }

int num_sms();  // multiprocessors of the current device (148 on B200), queried once per thread and device

// ---------------------------------------------------------------- programmatic dependent launch
// Every kernel of the library starts with pdl_begin(): it lets the NEXT kernel of the stream be scheduled as soon as all
// CTAs of this one are resident (its CTAs fill the SMs that this kernel's last wave leaves) and then waits until the
// PREVIOUS kernel has completed and its writes are visible -- before the first global access.  Chains of small kernels
// (index! is 9 to 11 launches) then pay the launch latency once instead of per kernel.  Launches go through launch_k(),
// which sets the programmatic-stream-serialization attribute when ETB_PDL=1.  MEASURED AND LEFT OFF BY DEFAULT: eager
// launches gain (C3 update! 144 -> 122 us at n = 65536), but the numbers that count are CUDA-graph replays, and there
// the programmatic edges lose (C3 index! 67 -> 106 us, C1 49 -> 50 us, C2 unchanged).  Without the attribute the two
// instructions of pdl_begin() do nothing.
bool pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_begin() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);  // errors surface in cudaGetLastError()
}
#endif

inline int pow2ceil(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

// ---------------------------------------------------------------- device table addressing
// Compact device-side form of etb_table: the GPU `columnpointer`
// (reference src/simple.jl:52-55, src/split.jl:59-65,81-86).
struct DevTable {
    const char* base;           // Simple; Cached: the host-resident table
    const char* const* chunks;  // Split (device array)
    int64_t row_stride;         // bytes between embedding rows (ld * sizeof(T))
    uint32_t shard_rows;        // 0 = Simple
    uint32_t pad;
    const char* cache_rows;     // Cached: the HBM row cache ...
    const int32_t* slot_of_row; // ... and the slot of every row (-1 = on the host); null for Simple / Split
};

inline DevTable make_dev_table(const etb_table& t) {
    DevTable d;
    d.base = (const char*)t.base;
    d.chunks = (const char* const*)t.chunks;
    d.row_stride = (int64_t)t.ld * (int64_t)elt_bytes(t.elt);
    d.shard_rows = t.chunks ? (uint32_t)t.shard_rows : 0u;
    d.pad = 0;
    d.cache_rows = nullptr;
    d.slot_of_row = nullptr;
    if (t.chunks && t.shard_rows == ETB_TABLE_CACHED) {  // `chunks` is a host pointer to the cache descriptor
        const etb_cache_desc* c = (const etb_cache_desc*)t.chunks;
        d.chunks = nullptr;
        d.shard_rows = 0xffffffffu;
        d.cache_rows = (const char*)c->rows;
        d.slot_of_row = c->slot_of_row;
    }
    return d;
}

int32_t validate_table(const etb_table& t, const char* who);

#ifdef __CUDACC__
// address of embedding row `i1` (1-based, as the host passes it)
__device__ __forceinline__ const char* row_ptr(const DevTable& t, int64_t i1) {
    uint64_t z = (uint64_t)(i1 - 1);
    if (t.shard_rows == 0) return t.base + z * (uint64_t)t.row_stride;
    if (t.slot_of_row) {  // host-tier table: the row's HBM copy when it has one
        const int32_t s = __ldg(t.slot_of_row + z);
        return s >= 0 ? t.cache_rows + (uint64_t)s * (uint64_t)t.row_stride : t.base + z * (uint64_t)t.row_stride;
    }
    uint64_t chunk, within;
    if (z <= 0xffffffffull) {  // 32-bit divide is ~4x cheaper than the 64-bit one
        uint32_t z32 = (uint32_t)z;
        chunk = z32 / t.shard_rows;
        within = z32 - (uint32_t)chunk * t.shard_rows;
    } else {
        chunk = z / t.shard_rows;
        within = z - chunk * t.shard_rows;
    }
    return (const char*)__ldg((const unsigned long long*)t.chunks + chunk) + within * (uint64_t)t.row_stride;
}

__device__ __forceinline__ const char* shfl_ptr(const char* p, int src, int width) {
    unsigned long long v = (unsigned long long)p;
    unsigned lo = __shfl_sync(0xffffffffu, (unsigned)v, src, width);
    unsigned hi = __shfl_sync(0xffffffffu, (unsigned)(v >> 32), src, width);
    return (const char*)(((unsigned long long)hi << 32) | lo);
}

// ---------------------------------------------------------------- vector access
// A VB-byte register vector of NE = VB/sizeof(T) elements of T.
template <typename T, int VB>
struct alignas(VB) Vec {
    static constexpr int NE = VB / (int)sizeof(T);
    T e[NE];
};

// Storage type -> arithmetic type.  Half-precision tables (ETB_F16 / ETB_BF16, an extension: the reference has
// Float32/Float64/integer tables only) are summed and updated in Float32 and rounded to nearest-even once, when
// the result is stored; every other type computes in itself.
template <typename T>
struct AccOf {
    using type = T;
};
template <>
struct AccOf<__half> {
    using type = float;
};
template <>
struct AccOf<__nv_bfloat16> {
    using type = float;
};
template <typename T>
using acc_t = typename AccOf<T>::type;

template <typename T>
__device__ __forceinline__ acc_t<T> to_acc(T x) {
    return x;
}
template <>
__device__ __forceinline__ float to_acc<__half>(__half x) {
    return __half2float(x);
}
template <>
__device__ __forceinline__ float to_acc<__nv_bfloat16>(__nv_bfloat16 x) {
    return __bfloat162float(x);
}
template <typename T>
__device__ __forceinline__ T from_acc(acc_t<T> x) {
    return x;
}
template <>
__device__ __forceinline__ __half from_acc<__half>(float x) {
    return __float2half_rn(x);
}
template <>
__device__ __forceinline__ __nv_bfloat16 from_acc<__nv_bfloat16>(float x) {
    return __float2bfloat16_rn(x);
}

// accumulator twin of Vec<T, VB>: the same NE elements, in the arithmetic type (identical to Vec for
// Float32/Float64/integers, twice the bytes for the half types)
template <typename T, int VB>
struct alignas(VB) AccVec {
    static constexpr int NE = VB / (int)sizeof(T);
    acc_t<T> e[NE];
};

template <typename T, int VB>
__device__ __forceinline__ void acc_add(AccVec<T, VB>& acc, const Vec<T, VB>& v) {
#pragma unroll
    for (int k = 0; k < Vec<T, VB>::NE; ++k) acc.e[k] = acc.e[k] + to_acc<T>(v.e[k]);
}
template <typename T, int VB>
__device__ __forceinline__ void acc_add(AccVec<T, VB>& acc, const AccVec<T, VB>& v) {
#pragma unroll
    for (int k = 0; k < Vec<T, VB>::NE; ++k) acc.e[k] = acc.e[k] + v.e[k];
}
template <typename T, int VB>
__device__ __forceinline__ void acc_fill(AccVec<T, VB>& acc, acc_t<T> x) {
#pragma unroll
    for (int k = 0; k < Vec<T, VB>::NE; ++k) acc.e[k] = x;
}
template <typename T, int VB>
__device__ __forceinline__ Vec<T, VB> acc_round(const AccVec<T, VB>& acc) {
    Vec<T, VB> out;
#pragma unroll
    for (int k = 0; k < Vec<T, VB>::NE; ++k) out.e[k] = from_acc<T>(acc.e[k]);
    return out;
}

// Read-only row loads through the non-coherent path with the DEFAULT L2 policy.  Measured on
// B200 (profiles/r1a): adding `.L1::no_allocate` makes the sectors evict_first in L2, and rows
// that are re-read (the cotangent slice in update!, hot Zipf rows in lookups) then miss L2 --
// +5.6 GB of DRAM reads on C2's update.
template <int VB>
__device__ __forceinline__ void ld_row(void* out, const char* p);
template <>
__device__ __forceinline__ void ld_row<16>(void* out, const char* p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    *(uint4*)out = v;
}
template <>
__device__ __forceinline__ void ld_row<8>(void* out, const char* p) {
    uint2 v;
    asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    *(uint2*)out = v;
}
template <>
__device__ __forceinline__ void ld_row<4>(void* out, const char* p) {
    uint32_t v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    *(uint32_t*)out = v;
}

// streaming store (written once, not re-read by this kernel): the GPU analogue of the
// reference's non-temporal stores (src/simd.jl:31-45)
template <int VB>
__device__ __forceinline__ void st_stream(char* p, const void* in);
template <>
__device__ __forceinline__ void st_stream<16>(char* p, const void* in) {
    uint4 v = *(const uint4*)in;
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
template <>
__device__ __forceinline__ void st_stream<8>(char* p, const void* in) {
    uint2 v = *(const uint2*)in;
    asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
template <>
__device__ __forceinline__ void st_stream<4>(char* p, const void* in) {
    uint32_t v = *(const uint32_t*)in;
    asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// coherent global vector load/store for the read-modify-write of table rows (explicit
// ld.global/st.global: a pointer fetched from a descriptor would otherwise compile to a generic LD/ST)
template <int VB>
__device__ __forceinline__ void ld_plain(void* out, const char* p) {
    if constexpr (VB == 16) {
        uint4 v;
        asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
        *(uint4*)out = v;
    } else if constexpr (VB == 8) {
        uint2 v;
        asm volatile("ld.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
        *(uint2*)out = v;
    } else {
        uint32_t v;
        asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
        *(uint32_t*)out = v;
    }
}
template <int VB>
__device__ __forceinline__ void st_plain(char* p, const void* in) {
    if constexpr (VB == 16) {
        const uint4 v = *(const uint4*)in;
        asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    } else if constexpr (VB == 8) {
        const uint2 v = *(const uint2*)in;
        asm volatile("st.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
    } else {
        const uint32_t v = *(const uint32_t*)in;
        asm volatile("st.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    }
}

template <typename IdxT>
__device__ __forceinline__ int64_t ld_index(const void* idx, int64_t pos) {
    return (int64_t)__ldg((const IdxT*)idx + pos);
}
#endif  // __CUDACC__

}  // namespace etb
