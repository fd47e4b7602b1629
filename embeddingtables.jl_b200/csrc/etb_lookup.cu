// etb_lookup.cu -- K1 gather, K2 pooled sum, K3 fused multi-table lookup (sm_100a).
//
// Replaces the reference's lookup!/maplookup! (src/lookup.jl:42-182, 233-241, 263-276,
// 316-371) and its SIMD register-tile primitives (src/simd.jl:1-52).
//
// Shape of the kernels.  An embedding row is `dim` contiguous elements = nvec vectors of
// VB bytes (VB = 16 when everything is 16-byte aligned, else 8 or 4).  A GROUP of G lanes
// (G = power of two <= 32, G*VPL >= nvec when the row fits one pass) owns one output
// column; lane gl of the group holds vectors gl, gl+G, ... of the accumulator in registers
// -- the GPU form of the reference's TiledSIMD{K,16,Float32} tile.  For dim 128 f32 a group
// is exactly one warp and every row read is one coalesced 512-byte LDG.128; for dim 64 two
// columns share a warp, for dim 16 eight do.
//
// Bit-exactness.  The pooled sum of one feature element is computed by ONE lane as
// ((a1 + a2) + a3) + ... in bag order, accumulator seeded with the first row -- the same
// association as lookup_static_inner (src/lookup.jl:139-146) and lookup_generic!
// (:112-128), so f32/f64 results equal the reference's bit for bit; there is no multiply, so
// nothing can be contracted into an FMA.  Latency is hidden by issuing U independent row
// loads before the U dependent adds (addresses depend only on the indices, which the group
// fetched with one coalesced load and passes around with shuffles).
//
// All `items` of one launch share (element type, VB, VPL, G, nvec, index type); an ensemble
// whose tables share dim and dtype -- the DLRM case -- is ONE launch, whatever the strategy:
// the three reference strategies differ only in where `dst` points.
#include <algorithm>
#include <vector>

#include "etb_common.cuh"

namespace etb {

constexpr int kThreads = 256;
constexpr int kMaxItems = 256;   // descriptors per launch (kernel-parameter space: 256*88 B = 22 KB of 32 KB)
constexpr int kGatherCols = 4;   // columns per group in the gather kernel (loads in flight)

struct LookupDesc {  // 88 bytes
    DevTable table;
    const void* idx;
    char* dst;
    int64_t ld_dst_bytes;
    uint32_t batch;
    uint32_t ld_idx;
    uint32_t bag;
    uint32_t pad;
};

struct LookupParams {
    LookupDesc item[kMaxItems];
    int32_t G;     // lanes per group
    int32_t nvec;  // VB-byte vectors per embedding row
};

template <typename T>
__device__ __forceinline__ T additive_identity() {
    return T(0);
}
template <>
__device__ __forceinline__ float additive_identity<float>() {
    return -0.0f;
}
template <>
__device__ __forceinline__ double additive_identity<double>() {
    return -0.0;
}

// ------------------------------------------------------------------------------------ K2/K3
// Occupancy over per-warp depth (measured, profiles/README.md): with 4 rows in flight per lane the VPL = 1
// kernel needs 32 registers -> 64 warps/SM; uniform C2 is at the DRAM limit either way (0.98 ms), the
// L2-bound Zipf case gains 17 % (0.48 -> 0.40 ms) over 8 rows in flight at 54 registers.  Rows narrower
// than 512 bytes keep 8 in flight (dim 16/32 lose 10-20 % with 4).
// DEEP = false (rows of >= 512 bytes, G = 32): 4 rows in flight per lane, 32 registers, 64 warps/SM.
// DEEP = true (narrower rows, several columns per warp): 8 rows in flight -- small rows need more of them.
template <typename T, int VB, int VPL, typename IdxT, bool DEEP>
__global__ void __launch_bounds__(kThreads, (VPL == 1 && !DEEP) ? (sizeof(T) == 4 ? 8 : 5) : (VPL <= 2 ? 4 : 3))
pooled_kernel(const __grid_constant__ LookupParams P) {
    pdl_begin();
    constexpr int U = (VPL == 1 && !DEEP) ? 4 : ((8 / VPL) > 1 ? (8 / VPL) : 1);  // row loads in flight per lane batch
    using V = Vec<T, VB>;
    using A = AccVec<T, VB>;  // == V except for half-precision tables (Float32 accumulation)
    const LookupDesc& d = P.item[blockIdx.y];
    const int G = P.G;
    const int nvec = P.nvec;
    const int gl = threadIdx.x & (G - 1);
    const uint32_t col_raw = blockIdx.x * (kThreads / G) + threadIdx.x / G;
    if (blockIdx.x * (kThreads / G) >= d.batch) return;  // whole block past this item's batch
    const bool active = col_raw < d.batch;
    // inactive groups recompute the last column (never stored): every load below stays in
    // bounds and every shuffle stays warp-uniform without a single branch in the hot loop
    const uint32_t col = active ? col_raw : d.batch - 1;
    const IdxT* ip = (const IdxT*)d.idx + (size_t)col * d.ld_idx;
    const uint32_t bag = d.bag;
    char* out = d.dst + (size_t)col * d.ld_dst_bytes;

    for (int pass0 = 0; pass0 < nvec; pass0 += G * VPL) {
        int vi[VPL];
#pragma unroll
        for (int p = 0; p < VPL; ++p) vi[p] = min(pass0 + gl + p * G, nvec - 1) * VB;
        // Seed with the additive identity that leaves the first row's bits untouched:
        // (-0.0) + x == x for every x (incl. x = +-0), so this equals "accumulator = first
        // row" of the reference (src/lookup.jl:139-140) without a special first iteration.
        A acc[VPL];
#pragma unroll
        for (int p = 0; p < VPL; ++p) acc_fill(acc[p], additive_identity<acc_t<T>>());
        for (uint32_t i0 = 0; i0 < bag; i0 += G) {
            const int m = (int)min((uint32_t)G, bag - i0);
            // one coalesced index load per group; each lane resolves one row address
            const char* myrow = row_ptr(d.table, (int64_t)__ldg(ip + i0 + min(gl, m - 1)));
            for (int j0 = 0; j0 < m; j0 += U) {
                const char* r[U];
#pragma unroll
                for (int u = 0; u < U; ++u) r[u] = shfl_ptr(myrow, min(j0 + u, m - 1), G);
                V v[U][VPL];
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int p = 0; p < VPL; ++p) ld_row<VB>(&v[u][p], r[u] + vi[p]);
                if (j0 + U <= m) {  // full batch: U ordered adds
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int p = 0; p < VPL; ++p) acc_add(acc[p], v[u][p]);
                } else {  // ragged tail: the clamped duplicates are loaded but not added
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (j0 + u < m)
#pragma unroll
                            for (int p = 0; p < VPL; ++p) acc_add(acc[p], v[u][p]);
                }
            }
        }
#pragma unroll
        for (int p = 0; p < VPL; ++p) {
            const int v = pass0 + gl + p * G;
            if (active && v < nvec) {
                const V res = acc_round(acc[p]);
                st_stream<VB>(out + (size_t)v * VB, &res);
            }
        }
    }
}

// ------------------------------------------------------------------------------------ K2 small bags
// With a bag of 1..4 rows one column gives a group only 1..4 loads to keep in flight and the kernel
// becomes latency-bound (measured: 19 % of peak at bag 1, dim 128).  This variant lets a group own C
// columns at once (C * BAG = 8 row loads in flight per lane at VPL = 1): one coalesced load fetches
// the C*BAG indices, all rows are requested, then each column is summed in bag order from the same
// identity seed -- the arithmetic per column is exactly pooled_kernel's.
template <typename T, int VPL, typename IdxT, int BAG, int C>
__global__ void __launch_bounds__(kThreads)
pooled_smallbag_kernel(const __grid_constant__ LookupParams P) {
    pdl_begin();
    constexpr int VB = 16;
    using V = Vec<T, VB>;
    const LookupDesc& d = P.item[blockIdx.y];
    const int G = P.G;            // C * BAG <= G is guaranteed by the host
    const int nvec = P.nvec;      // == G * VPL (exact fit), single pass
    const int gl = threadIdx.x & (G - 1);
    const uint32_t col0 = (blockIdx.x * (kThreads / G) + threadIdx.x / G) * C;
    if (blockIdx.x * (kThreads / G) * C >= d.batch) return;
    // lane gl < C*BAG resolves the row of (column col0 + gl / BAG, bag entry gl % BAG); columns past
    // the batch are clamped to the last one (computed, never stored)
    const int e = min(gl, C * BAG - 1);
    const uint32_t mycol = min(col0 + e / BAG, d.batch - 1);
    const char* myrow = row_ptr(d.table, (int64_t)__ldg((const IdxT*)d.idx + (size_t)mycol * d.ld_idx + e % BAG));
    V v[C * BAG][VPL];
#pragma unroll
    for (int j = 0; j < C * BAG; ++j) {
        const char* r = shfl_ptr(myrow, j, G);
#pragma unroll
        for (int p = 0; p < VPL; ++p) ld_row<VB>(&v[j][p], r + (gl + p * G) * VB);
    }
    (void)nvec;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        AccVec<T, VB> acc[VPL];
#pragma unroll
        for (int p = 0; p < VPL; ++p) acc_fill(acc[p], additive_identity<acc_t<T>>());
#pragma unroll
        for (int i = 0; i < BAG; ++i)
#pragma unroll
            for (int p = 0; p < VPL; ++p) acc_add(acc[p], v[c * BAG + i][p]);
        if (col0 + c < d.batch) {
            char* out = d.dst + (size_t)(col0 + c) * d.ld_dst_bytes;
#pragma unroll
            for (int p = 0; p < VPL; ++p) {
                const V res = acc_round(acc[p]);
                st_stream<VB>(out + (size_t)(gl + p * G) * VB, &res);
            }
        }
    }
}

// ------------------------------------------------------------------------------------ K1
// Non-reducing gather: a bit copy, element type irrelevant.  Each group copies kGatherCols
// columns with all their row loads in flight before the first store.
template <int VB, int VPL, typename IdxT>
__global__ void __launch_bounds__(kThreads)
gather_kernel(const __grid_constant__ LookupParams P) {
    pdl_begin();
    using V = Vec<uint32_t, VB>;
    const LookupDesc& d = P.item[blockIdx.y];
    const int G = P.G;
    const int nvec = P.nvec;
    const int gl = threadIdx.x & (G - 1);
    const int groups = kThreads / G;
    const uint32_t col_base = blockIdx.x * (groups * kGatherCols) + threadIdx.x / G;
    if (col_base >= d.batch) return;

    for (int pass0 = 0; pass0 < nvec; pass0 += G * VPL) {
        const char* r[kGatherCols];
#pragma unroll
        for (int c = 0; c < kGatherCols; ++c) {
            const uint32_t col = min(col_base + c * groups, d.batch - 1);
            r[c] = row_ptr(d.table, ld_index<IdxT>(d.idx, col));
        }
        V v[kGatherCols][VPL];
#pragma unroll
        for (int c = 0; c < kGatherCols; ++c)
#pragma unroll
            for (int p = 0; p < VPL; ++p)
                ld_row<VB>(&v[c][p], r[c] + (size_t)min(pass0 + gl + p * G, nvec - 1) * VB);
#pragma unroll
        for (int c = 0; c < kGatherCols; ++c) {
            const uint32_t col = col_base + c * groups;
#pragma unroll
            for (int p = 0; p < VPL; ++p) {
                const int vv = pass0 + gl + p * G;
                if (col < d.batch && vv < nvec)
                    st_stream<VB>(d.dst + (size_t)col * d.ld_dst_bytes + (size_t)vv * VB, &v[c][p]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------ host
struct LookupClass {
    int32_t elt, vb, vpl, G, nvec, idx_elt;
    bool pooled;
    bool operator==(const LookupClass& o) const {
        return elt == o.elt && vb == o.vb && vpl == o.vpl && G == o.G && nvec == o.nvec &&
               idx_elt == o.idx_elt && pooled == o.pooled;
    }
};

static bool aligned_to(const void* p, size_t a) { return ((uintptr_t)p % a) == 0; }

// widest vector every address of this item is aligned to
static int pick_vb(const etb_lookup_item& it) {
    const size_t es = elt_bytes(it.table.elt);
    const size_t rowbytes = (size_t)it.table.dim * es;
    const size_t stride = (size_t)it.table.ld * es;
    const size_t ldd = (size_t)it.ld_dst * es;
    for (size_t vb : {(size_t)16, (size_t)8}) {
        if (vb < es) continue;
        // Split tables: chunk bases live in device memory; the ABI requires them to be
        // 16-byte aligned (every CUDA allocation is 256-byte aligned).
        if (rowbytes % vb == 0 && stride % vb == 0 && ldd % vb == 0 && aligned_to(it.table.base, vb) &&
            aligned_to(it.dst, vb))
            return (int)vb;
    }
    if (es == 2)  // half-precision rows must at least be 4-byte aligned (even dim and leading dimensions)
        return (rowbytes % 4 == 0 && stride % 4 == 0 && ldd % 4 == 0 && aligned_to(it.table.base, 4) && aligned_to(it.dst, 4)) ? 4 : 0;
    return (int)es == 8 ? 8 : 4;
}

static LookupClass classify(const etb_lookup_item& it) {
    LookupClass c;
    c.elt = it.table.elt;
    c.pooled = it.bag > 0;
    c.idx_elt = it.idx_elt;
    c.vb = pick_vb(it);
    if (c.vb == 0) {  // misaligned half-precision rows: rejected by the caller
        c.nvec = c.G = c.vpl = 0;
        return c;
    }
    c.nvec = (int)((size_t)it.table.dim * elt_bytes(it.table.elt) / c.vb);
    c.G = std::min(32, pow2ceil(c.nvec));
    const int per_lane = (c.nvec + c.G - 1) / c.G;
    c.vpl = per_lane >= 4 ? 4 : (per_lane >= 2 ? 2 : 1);
    return c;
}

template <typename T, int VB, typename IdxT>
static cudaError_t launch_pooled_vpl(int vpl, dim3 grid, cudaStream_t s, const LookupParams& P) {
    switch (vpl) {
        case 1:
            if (P.G == 32) launch_k(pooled_kernel<T, VB, 1, IdxT, false>, grid, kThreads, 0, s, P);
            else launch_k(pooled_kernel<T, VB, 1, IdxT, true>, grid, kThreads, 0, s, P);
            break;
        case 2: launch_k(pooled_kernel<T, VB, 2, IdxT, true>, grid, kThreads, 0, s, P); break;
        default: launch_k(pooled_kernel<T, VB, 4, IdxT, true>, grid, kThreads, 0, s, P); break;
    }
    return cudaGetLastError();
}

template <typename T, typename IdxT>
static cudaError_t launch_pooled_vb(const LookupClass& c, dim3 grid, cudaStream_t s, const LookupParams& P) {
    if constexpr (sizeof(T) <= 4) {
        if (c.vb == 4) return launch_pooled_vpl<T, 4, IdxT>(c.vpl, grid, s, P);
    }
    if (c.vb == 8) return launch_pooled_vpl<T, 8, IdxT>(c.vpl, grid, s, P);
    return launch_pooled_vpl<T, 16, IdxT>(c.vpl, grid, s, P);
}

// small-bag variant: VB = 16, exact fit (nvec == G * VPL), every item of the launch has the same bag
// in {1, 2, 4} and C * BAG <= G.  Returns the columns-per-group it used (0 = not applicable).
template <typename T, typename IdxT>
static int launch_smallbag_t(const LookupClass& c, int bag, uint32_t max_batch, int n, cudaStream_t s,
                             const LookupParams& P, cudaError_t* err) {
    if (c.vb != 16 || c.nvec != c.G * c.vpl) return 0;
    const int groups = kThreads / c.G;
#define ETB_SB(VPLV, BAGV, CV)                                                                          \
    if (c.vpl == VPLV && bag == BAGV && CV * BAGV <= c.G) {                                             \
        dim3 grid((max_batch + groups * CV - 1) / (groups * CV), (unsigned)n, 1);                       \
        launch_k(pooled_smallbag_kernel<T, VPLV, IdxT, BAGV, CV>, grid, kThreads, 0, s, P);                   \
        *err = cudaGetLastError();                                                                      \
        return CV;                                                                                      \
    }
    ETB_SB(1, 1, 8) ETB_SB(1, 2, 4) ETB_SB(1, 4, 2) ETB_SB(1, 1, 4) ETB_SB(1, 2, 2) ETB_SB(1, 1, 2)
    ETB_SB(2, 1, 4) ETB_SB(2, 2, 2) ETB_SB(4, 1, 2)
#undef ETB_SB
    return 0;
}

template <typename IdxT>
static int launch_smallbag(const LookupClass& c, int bag, uint32_t max_batch, int n, cudaStream_t s,
                           const LookupParams& P, cudaError_t* err) {
    switch (c.elt) {
        case ETB_F32: return launch_smallbag_t<float, IdxT>(c, bag, max_batch, n, s, P, err);
        case ETB_F64: return launch_smallbag_t<double, IdxT>(c, bag, max_batch, n, s, P, err);
        case ETB_I32: return launch_smallbag_t<uint32_t, IdxT>(c, bag, max_batch, n, s, P, err);
        case ETB_F16: return launch_smallbag_t<__half, IdxT>(c, bag, max_batch, n, s, P, err);
        case ETB_BF16: return launch_smallbag_t<__nv_bfloat16, IdxT>(c, bag, max_batch, n, s, P, err);
        default: return launch_smallbag_t<unsigned long long, IdxT>(c, bag, max_batch, n, s, P, err);
    }
}

template <typename IdxT>
static cudaError_t launch_pooled(const LookupClass& c, dim3 grid, cudaStream_t s, const LookupParams& P) {
    switch (c.elt) {
        case ETB_F32: return launch_pooled_vb<float, IdxT>(c, grid, s, P);
        case ETB_F64: return launch_pooled_vb<double, IdxT>(c, grid, s, P);
        case ETB_I32: return launch_pooled_vb<uint32_t, IdxT>(c, grid, s, P);  // Julia ints wrap
        case ETB_F16: return launch_pooled_vb<__half, IdxT>(c, grid, s, P);
        case ETB_BF16: return launch_pooled_vb<__nv_bfloat16, IdxT>(c, grid, s, P);
        default: return launch_pooled_vb<unsigned long long, IdxT>(c, grid, s, P);
    }
}

template <int VB, typename IdxT>
static cudaError_t launch_gather_vpl(int vpl, dim3 grid, cudaStream_t s, const LookupParams& P) {
    switch (vpl) {
        case 1: launch_k(gather_kernel<VB, 1, IdxT>, grid, kThreads, 0, s, P); break;
        case 2: launch_k(gather_kernel<VB, 2, IdxT>, grid, kThreads, 0, s, P); break;
        default: launch_k(gather_kernel<VB, 4, IdxT>, grid, kThreads, 0, s, P); break;
    }
    return cudaGetLastError();
}

template <typename IdxT>
static cudaError_t launch_gather(const LookupClass& c, dim3 grid, cudaStream_t s, const LookupParams& P) {
    switch (c.vb) {
        case 4: return launch_gather_vpl<4, IdxT>(c.vpl, grid, s, P);
        case 8: return launch_gather_vpl<8, IdxT>(c.vpl, grid, s, P);
        default: return launch_gather_vpl<16, IdxT>(c.vpl, grid, s, P);
    }
}

static int32_t maplookup_impl(const etb_lookup_item* items, int32_t n_items, cudaStream_t stream) {
    launch_counter() = 0;
    ETB_REQUIRE(n_items >= 0, "etb_maplookup: negative item count");
    if (n_items == 0) return ETB_OK;
    ETB_REQUIRE(items, "etb_maplookup: null items");
    std::vector<LookupClass> cls((size_t)n_items);
    std::vector<char> done((size_t)n_items, 0);
    for (int i = 0; i < n_items; ++i) {
        const etb_lookup_item& it = items[i];
        if (int32_t st = validate_table(it.table, "etb_maplookup")) return st;
        ETB_REQUIRE(idx_elt_valid(it.idx_elt), "etb_maplookup: item %d: index type must be ETB_I32/ETB_I64", i);
        ETB_REQUIRE(it.batch >= 0 && it.batch <= 0xffffffffll, "etb_maplookup: item %d: bad batch %lld", i, (long long)it.batch);
        ETB_REQUIRE(it.bag >= 0 && it.bag <= 0xffffffffll, "etb_maplookup: item %d: bad bag %lld", i, (long long)it.bag);
        if (it.batch == 0) { done[i] = 1; continue; }  // empty lookup: nothing to write
        ETB_REQUIRE(it.idx && it.dst, "etb_maplookup: item %d: null idx/dst", i);
        ETB_REQUIRE(it.ld_dst >= it.table.dim, "etb_maplookup: item %d: ld_dst (%lld) < dim (%d)", i, (long long)it.ld_dst, it.table.dim);
        ETB_REQUIRE(it.bag == 0 || (it.ld_idx >= it.bag && it.ld_idx <= 0xffffffffll), "etb_maplookup: item %d: ld_idx (%lld) < bag (%lld)", i, (long long)it.ld_idx, (long long)it.bag);
        cls[i] = classify(it);
        if (cls[i].vb == 0)
            return fail(ETB_ERR_UNSUPPORTED, "etb_maplookup: item %d: half-precision rows must be 4-byte aligned (even dim, ld, ld_dst)", i);
    }
    static thread_local LookupParams P;  // 7 KB: keep it off the stack
    for (int i = 0; i < n_items; ++i) {
        if (done[i]) continue;
        const LookupClass c = cls[i];
        int n = 0;
        uint32_t max_batch = 0;
        int64_t common_bag = items[i].bag;  // small-bag kernel needs one bag for the whole launch
        for (int j = i; j < n_items && n < kMaxItems; ++j) {
            if (done[j] || !(cls[j] == c)) continue;
            const etb_lookup_item& it = items[j];
            LookupDesc& d = P.item[n++];
            d.table = make_dev_table(it.table);
            d.idx = it.idx;
            d.dst = (char*)it.dst;
            d.ld_dst_bytes = it.ld_dst * (int64_t)elt_bytes(it.table.elt);
            d.batch = (uint32_t)it.batch;
            d.ld_idx = (uint32_t)it.ld_idx;
            d.bag = (uint32_t)it.bag;
            d.pad = 0;
            max_batch = std::max(max_batch, d.batch);
            if (it.bag != common_bag) common_bag = -1;
            done[j] = 1;
        }
        P.G = c.G;
        P.nvec = c.nvec;
        const uint32_t cols_per_block = (uint32_t)(kThreads / c.G) * (c.pooled ? 1u : (uint32_t)kGatherCols);
        dim3 grid((max_batch + cols_per_block - 1) / cols_per_block, (unsigned)n, 1);
        cudaError_t e = cudaSuccess;
        if (c.pooled && (common_bag == 1 || common_bag == 2 || common_bag == 4) &&
            (c.idx_elt == ETB_I64 ? launch_smallbag<long long>(c, (int)common_bag, max_batch, n, stream, P, &e)
                                  : launch_smallbag<int>(c, (int)common_bag, max_batch, n, stream, P, &e)) > 0) {
            // launched by the small-bag variant
        } else if (c.pooled)
            e = c.idx_elt == ETB_I64 ? launch_pooled<long long>(c, grid, stream, P) : launch_pooled<int>(c, grid, stream, P);
        else
            e = c.idx_elt == ETB_I64 ? launch_gather<long long>(c, grid, stream, P) : launch_gather<int>(c, grid, stream, P);
        ++launch_counter();
        if (e != cudaSuccess) return fail(ETB_ERR_CUDA, "etb_maplookup: kernel launch failed: %s", cudaGetErrorString(e));
    }
    return ETB_OK;
}

}  // namespace etb

using namespace etb;

extern "C" {

int32_t etb_maplookup(const etb_lookup_item* items_host, int32_t n_items, void* stream) {
    ETB_API_RANGE();
    return maplookup_impl(items_host, n_items, (cudaStream_t)stream);
}

int32_t etb_gather(void* dst, int64_t ld_dst, const etb_table* table_host, const void* idx,
                   int32_t idx_elt, int64_t n, void* stream) {
    ETB_API_RANGE();
    ETB_REQUIRE(table_host, "etb_gather: null table");
    etb_lookup_item it;
    memset(&it, 0, sizeof(it));
    it.table = *table_host;
    it.idx = idx;
    it.dst = dst;
    it.ld_dst = ld_dst;
    it.batch = n;
    it.bag = 0;
    it.ld_idx = 0;
    it.idx_elt = idx_elt;
    return maplookup_impl(&it, 1, (cudaStream_t)stream);
}

int32_t etb_pooled_sum(void* dst, int64_t ld_dst, const etb_table* table_host, const void* idx,
                       int32_t idx_elt, int64_t bag, int64_t batch, int64_t ld_idx, void* stream) {
    ETB_API_RANGE();
    ETB_REQUIRE(table_host, "etb_pooled_sum: null table");
    ETB_REQUIRE(bag >= 1, "etb_pooled_sum: bag must be >= 1 (got %lld)", (long long)bag);
    etb_lookup_item it;
    memset(&it, 0, sizeof(it));
    it.table = *table_host;
    it.idx = idx;
    it.dst = dst;
    it.ld_dst = ld_dst;
    it.batch = batch;
    it.bag = bag;
    it.ld_idx = ld_idx;
    it.idx_elt = idx_elt;
    return maplookup_impl(&it, 1, (cudaStream_t)stream);
}

}  // extern "C"
