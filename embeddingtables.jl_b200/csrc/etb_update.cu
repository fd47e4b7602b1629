// etb_update.cu -- K5 fused segment-reduce + SGD over the buckets of index!, the a2a helpers, the host-tier cache (sm_100a).
//
// Replaces the reference's update! kernels (src/sparseupdate.jl:57-154, ensemble form :199-238), which walk the
// buckets of its Indexer (histogram!/prefixsum!/remap!, src/utils.jl:131-314: a stable counting sort of
// occurrence -> delta column by table row).
//
// K4 (index!) lives in etb_index.cu / etb_index.cuh: a per-table segmented, stable LSD radix sort of
// (row, delta column) pairs and one 16-byte record per bucket (start, first delta column, slot << row_bits | row).
// Members of a bucket stay in occurrence order = the order remap! records (src/utils.jl:242-272); buckets come
// out in ascending (table, row) order instead of the reference's first-seen order -- buckets are disjoint table
// rows, so results do not depend on that order.
//
// K5.  A group of G lanes owns one bucket: acc = 0; acc += delta[:, map[i]] for the bucket's
// members in order (one lane per feature vector, like the lookup kernels, so the sum has the
// reference's association, src/sparseupdate.jl:114-120); then one read-modify-write of the
// table row with a fused multiply-add (muladd, src/sparseupdate.jl:123-127) or two roundings
// (:88) -- a per-table choice, like the reference's dispatch.  One bucket = one row = one writer:
// no atomics touch table data.  Kernels: sgd_update_exact_kernel / sgd_update_kernel (a warp per
// tile of 32 buckets; buckets of up to 4 members finish here), bucket_tasks_kernel (buckets of 5..128
// members and, in the opt-in split order, the 128-member chunks of longer ones), long_strict_sliced_kernel
// (strict order, the default: a bucket of more than 128 members is cut into slices of 16 feature elements,
// every slice streamed and added on its own SM; long_strict_kernel, one CTA per bucket, for Adagrad),
// long_combine_kernel (split order: chunk partials of a long bucket, added in a fixed order).  The only
// atomics are worklist cursors; nothing that reaches a result depends on their order.
#include <algorithm>
#include <type_traits>
#include <vector>

#include "etb_common.cuh"
#include "etb_layout.cuh"

namespace etb {

constexpr int kUThreads = 256;
// Tuning constants, all measured on C2 / B200 (profiles/README.md).  What matters for this access pattern is
// occupancy and the absence of register spills (a spill of freshly loaded rows serialises the loads);
// more bytes in flight per warp, bigger tiles and L2 bulk prefetch of the tile's rows did nothing.
#define ETB_UPDATE_UB 8               /* generic kernel: buckets in flight per group (124 registers, 2 CTAs/SM) */
#define ETB_UPDATE_U 1                /* generic kernel: extra member rows in flight in the short-duplicates loop */
#define ETB_UPDATE_RPL 1              /* generic kernel: bucket records per lane (tile = 32 * RPL buckets; 2 gains 1.5 %) */
#define ETB_UPDATE_MIN_BLOCKS 2
#define ETB_UPDATE_EXACT_UB 4         /* exact-fit kernel: buckets in flight per group (64 registers, 4 CTAs/SM) */
#define ETB_UPDATE_EXACT_MIN_BLOCKS 4
#define ETB_UPDATE_USE_EXACT 1
constexpr int kUMaxItems = 96;

// ------------------------------------------------------------------------------------ K5
struct UpdDesc {  // 64 bytes
    DevTable table;
    const char* delta;
    int64_t ld_delta_bytes;
};
struct UpdParams {
    UpdDesc item[kUMaxItems];
    const BucketRec* recs;
    const int32_t* map;
    const int64_t* nnz;
    LongCounters* counters;
    LongRec* longs;
    ChunkRec* chunks;
    char* partials;
    int64_t partial_pitch;
    int64_t n_total;
    double eta;
    int32_t row_bits;
    int32_t slot0, nslots;  // this launch handles slots [slot0, slot0 + nslots)
    int32_t G, nvec, vb;
    int32_t fma, split_long;
    int32_t strict_long;  // strict order and the rows fit one pass: buckets of > kLongThreshold members go to long_strict_kernel
    int32_t num_splits, this_split;  // IndexerView, reference src/utils.jl:325-333
    // row-wise Adagrad only (OPT == kOptAdagrad): per-item state vectors (one acc_t element per table row)
    double eps;
    char* state[kUMaxItems];
};

enum { kOptSgd = 0, kOptAdagrad = 1 };

template <typename T>
__device__ __forceinline__ T sgd_epilogue(T row, T acc, T eta, bool fma);
template <>
__device__ __forceinline__ float sgd_epilogue<float>(float row, float acc, float eta, bool fma) {
    // fma:   muladd(-eta, acc, row), one rounding (reference src/sparseupdate.jl:108,123-127)
    // else:  row - eta*acc, two roundings (reference :88); intrinsics forbid contraction
    return fma ? __fmaf_rn(-eta, acc, row) : __fsub_rn(row, __fmul_rn(eta, acc));
}
template <>
__device__ __forceinline__ double sgd_epilogue<double>(double row, double acc, double eta, bool fma) {
    return fma ? __fma_rn(-eta, acc, row) : __dsub_rn(row, __dmul_rn(eta, acc));
}

// table element in, table element out; arithmetic in acc_t<T> (Float32 for the half types, one rounding at the end)
template <typename T>
__device__ __forceinline__ T sgd_apply(T row, acc_t<T> acc, acc_t<T> eta, bool fma) {
    return from_acc<T>(sgd_epilogue<acc_t<T>>(to_acc<T>(row), acc, eta, fma));
}

// partial sums of long buckets are rows of AccVec (arithmetic type): AS = sizeof(AccVec) / VB pieces of VB bytes
template <typename T, int VB>
__device__ __forceinline__ void ld_acc(AccVec<T, VB>& a, const char* p) {
#pragma unroll
    for (int k = 0; k < (int)sizeof(AccVec<T, VB>) / VB; ++k) ld_plain<VB>((char*)&a + k * VB, p + k * VB);
}
template <typename T, int VB>
__device__ __forceinline__ void st_acc(char* p, const AccVec<T, VB>& a) {
#pragma unroll
    for (int k = 0; k < (int)sizeof(AccVec<T, VB>) / VB; ++k) st_plain<VB>(p + k * VB, (const char*)&a + k * VB);
}

// correctly rounded, never contracted (the parity tests repeat these operations one by one on the CPU)
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float sqrt_rn(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ double sqrt_rn(double a) { return __dsqrt_rn(a); }

// Row-wise Adagrad (extension; the reference has Descent only, src/sparseupdate.jl:160-189).
// g = the bucket's summed cotangent (same order as the SGD path).  One state element per table row:
//     h = state[row] + (sum_d g_d^2) / dim;  state[row] = h;  row -= (eta / (sqrt(h) + eps)) * g
// The sum of squares has a fixed order: every lane adds the squares of its own elements (vector p ascending,
// element ascending), then the G lanes of the group are combined by an XOR butterfly (offsets G/2 ... 1); a + b is
// commutative, so every lane ends with the same bits.  Lanes past the row (`on[p]` false) contribute nothing.
// Called by all lanes of the group together; the row must fit one pass (nvec <= G * VPL, checked by the host).
// The caller loads `state_old` together with the table row (one bucket = one writer, so nothing changes it in
// between) -- loading it here would put a dependent DRAM round trip at the end of every bucket.
template <typename T, int VB, int VPL>
__device__ __forceinline__ void adagrad_apply(Vec<T, VB> (&out)[VPL], const Vec<T, VB> (&old)[VPL],
                                              const AccVec<T, VB> (&g)[VPL], const bool (&on)[VPL], acc_t<T>* state,
                                              acc_t<T> state_old, acc_t<T> eta, acc_t<T> eps, int dim, int G, int gl,
                                              unsigned gmask) {
    using A = acc_t<T>;
    A s = A(0);
#pragma unroll
    for (int p = 0; p < VPL; ++p)
        if (on[p])
#pragma unroll
            for (int e = 0; e < Vec<T, VB>::NE; ++e) s = add_rn(s, mul_rn(g[p].e[e], g[p].e[e]));
    for (int off = G >> 1; off > 0; off >>= 1) s = add_rn(s, __shfl_xor_sync(gmask, s, off));
    const A h = add_rn(state_old, div_rn(s, (A)dim));  // state_old: loaded by the caller with the row, long before
    if (gl == 0) *state = h;
    const A scale = div_rn(eta, add_rn(sqrt_rn(h), eps));
#pragma unroll
    for (int p = 0; p < VPL; ++p)
#pragma unroll
        for (int e = 0; e < Vec<T, VB>::NE; ++e)
            out[p].e[e] = from_acc<T>(sub_rn(to_acc<T>(old[p].e[e]), mul_rn(scale, g[p].e[e])));
}

__device__ __forceinline__ unsigned group_mask(int G, int lane) {
    return (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
}

__device__ __forceinline__ const char* shfl_ptr_mask(unsigned mask, const char* p, int src_lane) {
    unsigned long long v = (unsigned long long)p;
    unsigned lo = __shfl_sync(mask, (unsigned)v, src_lane);
    unsigned hi = __shfl_sync(mask, (unsigned)(v >> 32), src_lane);
    return (const char*)(((unsigned long long)hi << 32) | lo);
}

// acc += delta[:, map[i]] for i in [i0, stop), strictly in order, U rows in flight.
// All lanes of the calling group are converged; shuffles stay inside the group.
template <typename T, int VB, int VPL, int U>
__device__ __forceinline__ void accumulate_members(AccVec<T, VB> (&acc)[VPL], const UpdDesc& d, const int32_t* map,
                                                   int64_t i0, int64_t stop, const int (&vi)[VPL], int G, int gl,
                                                   int lane, unsigned gmask) {
    using V = Vec<T, VB>;
    for (; i0 < stop; i0 += G) {
        const int m = (int)min((int64_t)G, stop - i0);
        const char* mine_d = d.delta + (int64_t)__ldg(map + i0 + min(gl, m - 1)) * d.ld_delta_bytes;
        for (int j0 = 0; j0 < m; j0 += U) {
            V v[U][VPL];
#pragma unroll
            for (int w = 0; w < U; ++w) {
                const char* r = shfl_ptr_mask(gmask, mine_d, (lane & ~(G - 1)) + min(j0 + w, m - 1));
                if (w == 0 || j0 + w < m)  // no duplicate traffic for the clamped tail
#pragma unroll
                    for (int p = 0; p < VPL; ++p) ld_row<VB>(&v[w][p], r + vi[p]);
            }
#pragma unroll
            for (int w = 0; w < U; ++w)
                if (j0 + w < m)
#pragma unroll
                    for (int p = 0; p < VPL; ++p) acc_add(acc[p], v[w][p]);
        }
    }
}

// One warp owns a TILE of 32 consecutive buckets.  Lane l fetches bucket l's record with one
// coalesced load (start, key, first member), resolves the row and first-delta addresses and parks
// them in shared memory -- one metadata round trip per 32 buckets, one LDS.128 per bucket later.
// Each group of G lanes then walks the G buckets of its own lanes:
// UB buckets at a time, old table row + first delta row of all UB in flight together; buckets with
// duplicates then add their remaining members strictly in order (those rows come from L2).
// Tiles are taken in bucket order = (table, row) order, one tile per warp (no grid-stride loop):
// the warps resident at any moment then work inside ONE table, so the delta rows they re-read
// (that table's slice of the cotangent) stay L2-resident.  nnz lives on the device: the grid is
// sized for the upper bound n_total and surplus warps exit.
// hand a bucket to the long path: fixed-size chunks, combined in chunk order (cold; kept out of line
// so that it costs the hot loop no registers)
// strict order: a long bucket is one job of long_strict_kernel (one CTA streams its member rows through shared memory)
__device__ __noinline__ void register_strict_long_bucket(LongCounters* counters, LongRec* longs, uint32_t bucket) {
    longs[atomicAdd(&counters->n_long, 1u)] = LongRec{bucket, 0u, 0u, 0u};
}

// ... and of long_strict_sliced_kernel: the record carries everything a job needs (start, members, key), so that a job
// starts after ONE dependent load instead of three
__device__ __noinline__ void register_sliced_long_bucket(LongCounters* counters, LongRec* longs, uint32_t start, int cnt,
                                                         uint64_t key) {
    longs[atomicAdd(&counters->n_long, 1u)] = LongRec{start, (uint32_t)cnt, (uint32_t)key, (uint32_t)(key >> 32)};
}

__device__ __noinline__ void register_long_bucket(LongCounters* counters, LongRec* longs, ChunkRec* chunks,
                                                  uint32_t bucket, int cnt) {
    const uint32_t nch = (uint32_t)((cnt + kLongChunk - 1) / kLongChunk);
    const uint32_t j = atomicAdd(&counters->n_long, 1u);
    const uint32_t pb = atomicAdd(&counters->n_partials, nch);  // rows of the partial-sum buffer
    const uint32_t tb = atomicAdd(&counters->n_chunks, nch);    // slots of the task list
    longs[j] = LongRec{bucket, pb, nch, 0u};
    for (uint32_t c = 0; c < nch; ++c) chunks[tb + c] = ChunkRec{j, c};
}

struct alignas(16) TileMeta {
    const char* row;  // table row of the bucket
    const char* d0;   // delta row of its first member
};
struct alignas(16) TileMeta2 {
    uint32_t start;
    int32_t cnt;   // members; 0 = nothing to do for this launch (invalid / other class / long)
    int32_t slot;
    int32_t pad;
};

template <typename T, int VB, int VPL, int OPT>
__global__ void __launch_bounds__(kUThreads, ETB_UPDATE_MIN_BLOCKS)
sgd_update_kernel(const __grid_constant__ UpdParams P) {
    pdl_begin();
    constexpr int UB = (ETB_UPDATE_UB / VPL) > 1 ? (ETB_UPDATE_UB / VPL) : 1;  // buckets in flight per group
    constexpr int U = (ETB_UPDATE_U / VPL) > 1 ? (ETB_UPDATE_U / VPL) : 1;  // extra member rows in flight
    constexpr int RPL = ETB_UPDATE_RPL;                                       // records (buckets) per lane
    constexpr int TILE = 32 * RPL;                                            // buckets per warp
    using V = Vec<T, VB>;
    __shared__ TileMeta s_meta[kUThreads * RPL];
    __shared__ TileMeta2 s_meta2[kUThreads * RPL];
    const int G = P.G, nvec = P.nvec;
    const int lane = threadIdx.x & 31;
    const int gl = lane & (G - 1);
    const int wbase = (threadIdx.x & ~31) * RPL;    // this warp's slice of the shared arrays
    const int goff = lane & ~(G - 1);               // this group's first lane
    const unsigned gmask = group_mask(G, lane);
    const int64_t nnz = *P.nnz;
    const uint64_t row_mask = (P.row_bits >= 64) ? ~0ull : ((1ull << P.row_bits) - 1ull);
    const acc_t<T> eta = (acc_t<T>)P.eta;  // convert(eltype(table), opt.eta), reference src/sparseupdate.jl:173

    int64_t s_begin = 0, s_end = nnz;
    if (P.num_splits > 0) {  // cumulative has nnz+1 entries; split_size = cdiv(nnz+1, num_splits)
        const int64_t split = nnz / P.num_splits + 1;
        s_begin = (int64_t)(P.this_split - 1) * split;
        s_end = min((int64_t)P.this_split * split, nnz);
    }
    const int64_t warp_id = (int64_t)blockIdx.x * (kUThreads / 32) + threadIdx.x / 32;
    const int64_t t0 = s_begin + warp_id * TILE;
    if (t0 >= s_end) return;

    // ---- tile metadata: lane l describes buckets t0 + l (+ 32, ...)
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const int64_t sb = t0 + r * 32 + lane;
        const bool valid = sb < s_end;
        const int64_t s = valid ? sb : s_end - 1;
        const uint4 raw = __ldg((const uint4*)(P.recs + s));
        const int64_t start = raw.x;
        const int64_t stop = (s + 1 < nnz) ? (int64_t)__ldg(&P.recs[s + 1].start) : P.n_total;
        const uint64_t key = ((uint64_t)raw.w << 32) | raw.z;
        const int slot = (int)(key >> P.row_bits) - P.slot0;
        const bool mine = valid && slot >= 0 && slot < P.nslots;  // else another launch's class
        const UpdDesc& md = P.item[mine ? slot : 0];
        int cnt = mine ? (int)(stop - start) : 0;
        if (cnt > kShortMax) {
            if (P.split_long && cnt > kLongThreshold) register_long_bucket(P.counters, P.longs, P.chunks, (uint32_t)s, cnt);
            else if (P.strict_long == 2 && cnt > kLongThreshold) register_sliced_long_bucket(P.counters, P.longs, (uint32_t)start, cnt, key);
            else if (P.strict_long && cnt > kLongThreshold) register_strict_long_bucket(P.counters, P.longs, (uint32_t)s);
            else P.chunks[atomicAdd(&P.counters->n_chunks, 1u)] = ChunkRec{kMediumTask, (uint32_t)s};
            cnt = 0;
        }
        // addresses only for buckets this launch finishes itself: a padding lane or a bucket of another kernel
        // class would pair item[0]'s table with a row of a different table (a Split table would then read its
        // chunk-pointer array out of bounds)
        TileMeta m{nullptr, nullptr};
        if (cnt > 0) {
            m.row = row_ptr(md.table, (int64_t)(key & row_mask) + 1);
            m.d0 = md.delta + (int64_t)(int32_t)raw.y * md.ld_delta_bytes;
        }
        s_meta[wbase + r * 32 + lane] = m;
        // pad: the SGD epilogue's FMA flag, or (Adagrad) the row number for the state vector
        s_meta2[wbase + r * 32 + lane] = TileMeta2{raw.x, cnt, mine ? slot : 0,
                                                   OPT == kOptSgd ? (int32_t)md.table.pad : (int32_t)(key & row_mask)};
    }
    __syncwarp();

    for (int pass0 = 0; pass0 < nvec; pass0 += G * VPL) {
        int vi[VPL];
#pragma unroll
        for (int p = 0; p < VPL; ++p) vi[p] = min(pass0 + gl + p * G, nvec - 1) * VB;
        for (int r = 0; r < RPL; ++r) {
            const int gbase = wbase + r * 32 + goff;  // my group's buckets of this 32-record slab
            for (int k0 = 0; k0 < G; k0 += UB) {
                // the old table row and the first delta row of UB buckets, all in flight together
                V old[UB][VPL], v0[UB][VPL];
                acc_t<T> st_old[OPT == kOptAdagrad ? UB : 1];
                (void)st_old;
#pragma unroll
                for (int u = 0; u < UB; ++u) {
                    if (k0 + u < G && s_meta2[gbase + k0 + u].cnt > 0) {
                        const TileMeta m = s_meta[gbase + k0 + u];
                        if constexpr (OPT == kOptAdagrad) {
                            const TileMeta2 q = s_meta2[gbase + k0 + u];
                            st_old[u] = *((const acc_t<T>*)P.state[q.slot] + q.pad);
                        }
#pragma unroll
                        for (int p = 0; p < VPL; ++p) {
                            ld_plain<VB>(&old[u][p], m.row + vi[p]);
                            ld_row<VB>(&v0[u][p], m.d0 + vi[p]);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < UB; ++u) {
                    if (k0 + u < G) {
                        const TileMeta2 m2 = s_meta2[gbase + k0 + u];
                        if (m2.cnt > 0) {
                            AccVec<T, VB> acc[VPL];  // accum = zero(Tiled), then += members in order (src/sparseupdate.jl:114-120)
#pragma unroll
                            for (int p = 0; p < VPL; ++p)
#pragma unroll
                                for (int e = 0; e < V::NE; ++e) acc[p].e[e] = acc_t<T>(0) + to_acc<T>(v0[u][p].e[e]);
                            if (m2.cnt > 1)  // a few duplicates (<= kShortMax members): the rest, strictly in order
                                accumulate_members<T, VB, VPL, U>(acc, P.item[m2.slot], P.map, (int64_t)m2.start + 1,
                                                                  (int64_t)m2.start + m2.cnt, vi, G, gl, lane, gmask);
                            char* row = const_cast<char*>(s_meta[gbase + k0 + u].row);
                            if constexpr (OPT == kOptSgd) {
#pragma unroll
                                for (int p = 0; p < VPL; ++p) {
                                    if (pass0 + gl + p * G < nvec) {
                                        V out;
#pragma unroll
                                        for (int e = 0; e < V::NE; ++e)
                                            out.e[e] = sgd_apply<T>(old[u][p].e[e], acc[p].e[e], eta, m2.pad != 0);
                                        st_plain<VB>(row + vi[p], &out);
                                    }
                                }
                            } else {
                                bool on[VPL];
#pragma unroll
                                for (int p = 0; p < VPL; ++p) on[p] = gl + p * G < nvec;
                                V out[VPL];
                                adagrad_apply<T, VB, VPL>(out, old[u], acc, on, (acc_t<T>*)P.state[m2.slot] + m2.pad,
                                                          st_old[OPT == kOptAdagrad ? u : 0], eta, (acc_t<T>)P.eps,
                                                          nvec * V::NE, G, gl, gmask);
#pragma unroll
                                for (int p = 0; p < VPL; ++p)
                                    if (on[p]) st_plain<VB>(row + vi[p], &out[p]);
                            }
                        }
                    }
                }
            }
        }
    }
}

// Single-pass specialisation of the main kernel: VB = 16 and the row fits G*VPL vectors (every dim that is
// a multiple of 4 floats up to 512, e.g. dim 128 f32 -> G = 32, VPL = 1; dim 80 -> 20 of 32 lanes active).  G, the vector offsets and the single pass
// are compile-time, the few-duplicates loop needs no shuffles (every lane reads the same map entry:
// one broadcast transaction), so the kernel fits in few registers and runs at 3-4x the occupancy of
// the generic one -- which is what the random 512-byte read-modify-write pattern wants: the HBM
// ceiling (6.5 TB/s, tools/ubench_rmw.cu) is reached with only 4 rows in flight per warp but needs
// ~all warp slots filled.
template <typename T, int VPL, int G, int UB, int OPT>
__global__ void __launch_bounds__(kUThreads, ETB_UPDATE_EXACT_MIN_BLOCKS)
sgd_update_exact_kernel(const __grid_constant__ UpdParams P) {
    pdl_begin();
    constexpr int VB = 16;
    using V = Vec<T, VB>;
    __shared__ TileMeta s_meta[kUThreads];
    __shared__ TileMeta2 s_meta2[kUThreads];
    const int lane = threadIdx.x & 31;
    const int voff = (lane & (G - 1)) * VB;                  // my first vector of a row
    // rows narrower than G*VPL vectors (e.g. dim 80 = 20 vectors on G = 32): the surplus lanes idle
    bool on[VPL];
#pragma unroll
    for (int p = 0; p < VPL; ++p) on[p] = (lane & (G - 1)) + p * G < P.nvec;
    const int gbase = threadIdx.x & ~(G - 1);                // my group's slice of the shared arrays
    const int64_t nnz = *P.nnz;
    int64_t s_begin = 0, s_end = nnz;
    if (P.num_splits > 0) {
        const int64_t split = nnz / P.num_splits + 1;
        s_begin = (int64_t)(P.this_split - 1) * split;
        s_end = min((int64_t)P.this_split * split, nnz);
    }
    const int64_t t0 = s_begin + (((int64_t)blockIdx.x * kUThreads + threadIdx.x) >> 5) * 32;
    if (t0 >= s_end) return;
    {   // ---- tile metadata: lane l describes bucket t0 + l
        const bool valid = t0 + lane < s_end;
        const int64_t s = valid ? t0 + lane : s_end - 1;
        const uint4 raw = __ldg((const uint4*)(P.recs + s));
        const int64_t stop = (s + 1 < nnz) ? (int64_t)__ldg(&P.recs[s + 1].start) : P.n_total;
        const uint64_t key = ((uint64_t)raw.w << 32) | raw.z;
        const int slot = (int)(key >> P.row_bits) - P.slot0;
        const bool mine = valid && slot >= 0 && slot < P.nslots;
        const UpdDesc& md = P.item[mine ? slot : 0];
        int cnt = mine ? (int)(stop - (int64_t)raw.x) : 0;
        if (cnt > kShortMax) {
            if (P.split_long && cnt > kLongThreshold) register_long_bucket(P.counters, P.longs, P.chunks, (uint32_t)s, cnt);
            else if (P.strict_long == 2 && cnt > kLongThreshold) register_sliced_long_bucket(P.counters, P.longs, raw.x, cnt, key);
            else if (P.strict_long && cnt > kLongThreshold) register_strict_long_bucket(P.counters, P.longs, (uint32_t)s);
            else P.chunks[atomicAdd(&P.counters->n_chunks, 1u)] = ChunkRec{kMediumTask, (uint32_t)s};
            cnt = 0;
        }
        const uint64_t row_mask = (P.row_bits >= 64) ? ~0ull : ((1ull << P.row_bits) - 1ull);
        TileMeta m{nullptr, nullptr};  // see sgd_update_kernel: no address arithmetic for buckets that are not mine
        if (cnt > 0) {
            m.row = row_ptr(md.table, (int64_t)(key & row_mask) + 1);
            m.d0 = md.delta + (int64_t)(int32_t)raw.y * md.ld_delta_bytes;
        }
        s_meta[threadIdx.x] = m;
        s_meta2[threadIdx.x] = TileMeta2{raw.x, cnt, mine ? slot : 0,
                                         OPT == kOptSgd ? (int32_t)md.table.pad : (int32_t)(key & row_mask)};
    }
    __syncwarp();
    const acc_t<T> eta = (acc_t<T>)P.eta;
#pragma unroll 1
    for (int k0 = 0; k0 < G; k0 += UB) {
        // the first member's delta row is loaded straight into the accumulator registers when the arithmetic type
        // is the storage type; the half types stage it and convert
        constexpr bool kSame = sizeof(AccVec<T, VB>) == sizeof(V);
        V old[UB][VPL];
        AccVec<T, VB> acc[UB][VPL];
        V first[kSame ? 1 : UB][kSame ? 1 : VPL];
        (void)first;
        // Adagrad: the state element comes in with the row (a load at the end of the bucket would be a dependent
        // DRAM round trip); the half types have no registers to spare for that (measured: a 16-byte spill costs more)
        constexpr bool kEarlyState = OPT == kOptAdagrad && sizeof(T) >= 4;
        acc_t<T> st_old[kEarlyState ? UB : 1];
        (void)st_old;
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            if (s_meta2[gbase + k0 + u].cnt > 0) {
                const TileMeta m = s_meta[gbase + k0 + u];
                if constexpr (kEarlyState) {
                    const TileMeta2 q = s_meta2[gbase + k0 + u];
                    st_old[u] = *((const acc_t<T>*)P.state[q.slot] + q.pad);
                }
#pragma unroll
                for (int p = 0; p < VPL; ++p) {
                    if (on[p]) {
                        ld_plain<VB>(&old[u][p], m.row + voff + p * G * VB);
                        if constexpr (kSame) ld_row<VB>(&acc[u][p], m.d0 + voff + p * G * VB);
                        else ld_row<VB>(&first[u][p], m.d0 + voff + p * G * VB);
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const TileMeta2 m2 = s_meta2[gbase + k0 + u];
            if (m2.cnt > 0) {
#pragma unroll
                for (int p = 0; p < VPL; ++p)
#pragma unroll
                    for (int e = 0; e < V::NE; ++e) {  // accum = zero + first
                        if constexpr (kSame) acc[u][p].e[e] = acc_t<T>(0) + acc[u][p].e[e];
                        else acc[u][p].e[e] = acc_t<T>(0) + to_acc<T>(first[u][p].e[e]);
                    }
                if (m2.cnt > 1) {  // up to kShortMax-1 more members, strictly in order
                    const UpdDesc& d = P.item[m2.slot];
                    for (int i = 1; i < m2.cnt; ++i) {
                        const char* r = d.delta + (int64_t)__ldg(P.map + m2.start + i) * d.ld_delta_bytes + voff;
#pragma unroll
                        for (int p = 0; p < VPL; ++p) {
                            if (on[p]) {
                                V v;
                                ld_row<VB>(&v, r + p * G * VB);
                                acc_add(acc[u][p], v);
                            }
                        }
                    }
                }
                char* row = const_cast<char*>(s_meta[gbase + k0 + u].row) + voff;
                if constexpr (OPT == kOptSgd) {
#pragma unroll
                    for (int p = 0; p < VPL; ++p) {
                        if (on[p]) {
                            V out;
#pragma unroll
                            for (int e = 0; e < V::NE; ++e) out.e[e] = sgd_apply<T>(old[u][p].e[e], acc[u][p].e[e], eta, m2.pad != 0);
                            st_plain<VB>(row + p * G * VB, &out);
                        }
                    }
                } else {
                    V out[VPL];
                    acc_t<T>* state = (acc_t<T>*)P.state[m2.slot] + m2.pad;
                    adagrad_apply<T, VB, VPL>(out, old[u], acc[u], on, state, kEarlyState ? st_old[kEarlyState ? u : 0] : *state,
                                              eta, (acc_t<T>)P.eps, P.nvec * V::NE, G, lane & (G - 1), group_mask(G, lane));
#pragma unroll
                    for (int p = 0; p < VPL; ++p)
                        if (on[p]) st_plain<VB>(row + p * G * VB, &out[p]);
                }
            }
        }
    }
}

// Task kernel: MEDIUM buckets and the chunks of LONG buckets are the same job -- sum a run of members
// strictly in occurrence order from zero, 8 rows in flight -- so they share one worklist and one
// launch (they overlap instead of running back to back).  A medium task finishes its table row; a
// chunk task writes its partial row.
template <typename T, int VB, int VPL, int OPT>
__global__ void __launch_bounds__(kUThreads, 2)  // up to 128 registers: a spill of loaded rows serialises the loads
bucket_tasks_kernel(const __grid_constant__ UpdParams P) {
    pdl_begin();
    constexpr int U = (8 / VPL) > 1 ? (8 / VPL) : 1;
    using V = Vec<T, VB>;
    const int G = P.G, nvec = P.nvec;
    const int lane = threadIdx.x & 31;
    const int gl = lane & (G - 1);
    const unsigned gmask = group_mask(G, lane);
    const uint32_t n_tasks = P.counters->n_chunks;
    const int64_t nnz = *P.nnz;
    const uint64_t row_mask = (P.row_bits >= 64) ? ~0ull : ((1ull << P.row_bits) - 1ull);
    const acc_t<T> eta = (acc_t<T>)P.eta;
    constexpr int AS = (int)sizeof(AccVec<T, VB>) / VB;  // a partial row holds AccVecs
    const uint32_t groups_total = gridDim.x * (kUThreads / G);
    for (uint32_t i = blockIdx.x * (kUThreads / G) + threadIdx.x / G; i < n_tasks; i += groups_total) {
        const ChunkRec cr = P.chunks[i];
        const bool medium = cr.long_id == kMediumTask;
        uint32_t bucket, pbase = 0;
        if (medium) {
            bucket = cr.chunk;
        } else {
            const LongRec lr = P.longs[cr.long_id];
            bucket = lr.bucket;
            pbase = lr.chunk_base;
        }
        const BucketRec rec = P.recs[bucket];
        const int64_t stop_all = ((int64_t)bucket + 1 < nnz) ? (int64_t)P.recs[bucket + 1].start : P.n_total;
        const int64_t a = medium ? (int64_t)rec.start : (int64_t)rec.start + (int64_t)cr.chunk * kLongChunk;
        const int64_t b = medium ? stop_all : min(a + kLongChunk, stop_all);
        const UpdDesc& d = P.item[(int)(rec.key >> P.row_bits) - P.slot0];
        char* row = const_cast<char*>(row_ptr(d.table, (int64_t)(rec.key & row_mask) + 1));
        char* part = P.partials + (int64_t)(pbase + cr.chunk) * P.partial_pitch;
        for (int pass0 = 0; pass0 < nvec; pass0 += G * VPL) {
            int vi[VPL];
#pragma unroll
            for (int p = 0; p < VPL; ++p) vi[p] = min(pass0 + gl + p * G, nvec - 1) * VB;
            V old[VPL];
            AccVec<T, VB> acc[VPL];
            acc_t<T>* state = nullptr;
            acc_t<T> st_old = acc_t<T>(0);
            if (OPT == kOptAdagrad && medium) {
                state = (acc_t<T>*)P.state[(int)(rec.key >> P.row_bits) - P.slot0] + (int64_t)(rec.key & row_mask);
                st_old = *state;
            }
#pragma unroll
            for (int p = 0; p < VPL; ++p) {
                if (medium) ld_plain<VB>(&old[p], row + vi[p]);
                acc_fill(acc[p], acc_t<T>(0));
            }
            accumulate_members<T, VB, VPL, U>(acc, d, P.map, a, b, vi, G, gl, lane, gmask);
            if (OPT == kOptAdagrad && medium) {  // single pass (host-checked); the whole group is here together
                bool on[VPL];
#pragma unroll
                for (int p = 0; p < VPL; ++p) on[p] = gl + p * G < nvec;
                V out[VPL];
                adagrad_apply<T, VB, VPL>(out, old, acc, on, state, st_old, eta, (acc_t<T>)P.eps, nvec * V::NE, G, gl, gmask);
#pragma unroll
                for (int p = 0; p < VPL; ++p)
                    if (on[p]) st_plain<VB>(row + vi[p], &out[p]);
                continue;
            }
#pragma unroll
            for (int p = 0; p < VPL; ++p) {
                if (pass0 + gl + p * G < nvec) {
                    if (medium) {
                        V out;
#pragma unroll
                        for (int k = 0; k < V::NE; ++k) out.e[k] = sgd_apply<T>(old[p].e[k], acc[p].e[k], eta, d.table.pad != 0);
                        st_plain<VB>(row + vi[p], &out);
                    } else {
                        st_acc<T, VB>(part + (int64_t)vi[p] * AS, acc[p]);
                    }
                }
            }
        }
    }
}

// LONG buckets, phase B: one CTA per long bucket.  Its kUThreads/G groups each add a contiguous range of
// the bucket's partial rows in chunk order (8 in flight); the group sums are then added in group order
// through shared memory and the epilogue is applied.  The split depends only on the chunk count:
// deterministic.
template <typename T, int VB, int VPL, int OPT>
__global__ void __launch_bounds__(kUThreads)
long_combine_kernel(const __grid_constant__ UpdParams P) {
    pdl_begin();
    constexpr int U = (8 / VPL) > 1 ? (8 / VPL) : 1;
    using V = Vec<T, VB>;
    using A = AccVec<T, VB>;
    constexpr int AB = (int)sizeof(A);  // bytes of one accumulator vector (= VB except for the half types)
    __shared__ __align__(16) char s_sum[kUThreads * VPL * AB];  // [group][G * VPL vectors]
    const int G = P.G, nvec = P.nvec;
    const int gl = threadIdx.x & (G - 1);
    const int grp = threadIdx.x / G, ngroups = kUThreads / G;
    const uint32_t n_long = P.counters->n_long;
    const uint64_t row_mask = (P.row_bits >= 64) ? ~0ull : ((1ull << P.row_bits) - 1ull);
    const acc_t<T> eta = (acc_t<T>)P.eta;
    for (uint32_t j = blockIdx.x; j < n_long; j += gridDim.x) {
        const LongRec lr = P.longs[j];
        const BucketRec rec = P.recs[lr.bucket];
        const UpdDesc& d = P.item[(int)(rec.key >> P.row_bits) - P.slot0];
        char* row = const_cast<char*>(row_ptr(d.table, (int64_t)(rec.key & row_mask) + 1));
        const char* part = P.partials + (int64_t)lr.chunk_base * P.partial_pitch;
        const uint32_t per = (lr.nchunks + ngroups - 1) / ngroups;
        const uint32_t c_lo = min((uint32_t)grp * per, lr.nchunks), c_hi = min(c_lo + per, lr.nchunks);
        const int used_groups = (int)((lr.nchunks + per - 1) / per);
        for (int pass0 = 0; pass0 < nvec; pass0 += G * VPL) {
            int vi[VPL];
#pragma unroll
            for (int p = 0; p < VPL; ++p) vi[p] = min(pass0 + gl + p * G, nvec - 1) * VB;
            A acc[VPL];
#pragma unroll
            for (int p = 0; p < VPL; ++p) acc_fill(acc[p], acc_t<T>(0));
            for (uint32_t c0 = c_lo; c0 < c_hi; c0 += U) {
                A v[U][VPL];
#pragma unroll
                for (int w = 0; w < U; ++w)
                    if (c0 + w < c_hi)
#pragma unroll
                        for (int p = 0; p < VPL; ++p)
                            ld_acc<T, VB>(v[w][p], part + (int64_t)(c0 + w) * P.partial_pitch + (int64_t)vi[p] * (AB / VB));
#pragma unroll
                for (int w = 0; w < U; ++w)
                    if (c0 + w < c_hi)
#pragma unroll
                        for (int p = 0; p < VPL; ++p) acc_add(acc[p], v[w][p]);
            }
#pragma unroll
            for (int p = 0; p < VPL; ++p) *(A*)(s_sum + ((size_t)(grp * VPL + p) * G + gl) * AB) = acc[p];
            __syncthreads();
            if (grp == 0) {
                V old[VPL];
                A tot[VPL];
                bool on[VPL];
#pragma unroll
                for (int p = 0; p < VPL; ++p) {
                    tot[p] = *(const A*)(s_sum + ((size_t)p * G + gl) * AB);
                    on[p] = pass0 + gl + p * G < nvec;
                    if (on[p]) ld_plain<VB>(&old[p], row + vi[p]);
                    for (int g2 = 1; g2 < used_groups; ++g2) {  // groups without chunks are skipped (no 0 + -0)
                        const A v = *(const A*)(s_sum + ((size_t)(g2 * VPL + p) * G + gl) * AB);
                        acc_add(tot[p], v);
                    }
                }
                if constexpr (OPT == kOptSgd) {
#pragma unroll
                    for (int p = 0; p < VPL; ++p) {
                        if (on[p]) {
                            V out;
#pragma unroll
                            for (int k = 0; k < V::NE; ++k) out.e[k] = sgd_apply<T>(old[p].e[k], tot[p].e[k], eta, d.table.pad != 0);
                            st_plain<VB>(row + vi[p], &out);
                        }
                    }
                } else {  // single pass (host-checked); group 0 = the first G lanes of warp 0
                    V out[VPL];
                    acc_t<T>* state = (acc_t<T>*)P.state[(int)(rec.key >> P.row_bits) - P.slot0] + (int64_t)(rec.key & row_mask);
                    adagrad_apply<T, VB, VPL>(out, old, tot, on, state, *state, eta, (acc_t<T>)P.eps, nvec * V::NE, G, gl,
                                              group_mask(G, threadIdx.x & 31));
#pragma unroll
                    for (int p = 0; p < VPL; ++p)
                        if (on[p]) st_plain<VB>(row + vi[p], &out[p]);
                }
            }
            __syncthreads();
        }
    }
}

// LONG buckets in STRICT order (the reference's: every member added one after the other, src/sparseupdate.jl:72-84,
// 114-120).  The sum of one feature element is a serial chain of additions, so a hot row's time is members x (one
// add); what a GPU can do is make sure the chain never waits for memory: one CTA per long bucket, all 8 warps stream
// the member rows into shared memory with cp.async (4 stages of up to 64 rows), and the first G lanes add them from
// there in order (C3's hottest row has 45 k members: measurements in profiles/README.md).
constexpr int kStrictStageBytes = 24 * 1024, kStrictStages = 4, kStrictMaxRows = 48;  // 96 KB: two CTAs per SM
constexpr int kStrictPieces = 7;  // VB-byte pieces per producer thread and batch (a batch has at most 7 * 224 of them)

template <int VB>
__device__ __forceinline__ void cp_async(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    if constexpr (VB == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
    else if constexpr (VB == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}

// shared-memory vector load the compiler may not sink towards its use: the consumer wants U of them in flight
template <int VB>
__device__ __forceinline__ void lds_vec(void* out, const char* p) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    if constexpr (VB == 16) {
        uint4 v;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
        *(uint4*)out = v;
    } else if constexpr (VB == 8) {
        uint2 v;
        asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
        *(uint2*)out = v;
    } else {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
        *(uint32_t*)out = v;
    }
}

template <typename T, int VB, int VPL, int OPT>
__global__ void __launch_bounds__(kUThreads)
long_strict_kernel(const __grid_constant__ UpdParams P) {
    pdl_begin();
    using V = Vec<T, VB>;
    using A = AccVec<T, VB>;
    extern __shared__ __align__(16) char s_rows[];  // [kStrictStages][rows_per_stage][nvec * VB]
    const int G = P.G, nvec = P.nvec;
    const int row_bytes = nvec * VB;
    const int rows_per_stage = min(min(kStrictMaxRows, kStrictStageBytes / row_bytes), kStrictPieces * (kUThreads - 32) / nvec);
    const int gl = threadIdx.x & (G - 1);
    const bool consumer = threadIdx.x < G;  // group 0 = the first G lanes of warp 0 add; warp 0 never loads,
    const bool producer = threadIdx.x >= 32;  // warps 1..7 never add
    constexpr int kProducers = kUThreads - 32;
    // a producer's pieces of a batch: piece pc = (tid - 32) + k * 224 is vector pv[k] of the batch's row pr[k] -- the
    // same for every batch, so the divisions happen once per kernel
    int pr[kStrictPieces], pvb[kStrictPieces], poff[kStrictPieces];  // row in the batch, byte in the row, byte in the stage
#pragma unroll
    for (int k = 0; k < kStrictPieces; ++k) {
        const int pc = (int)threadIdx.x - 32 + k * kProducers;
        pr[k] = producer ? pc / nvec : rows_per_stage;  // warp 0: no pieces
        pvb[k] = (pc - pr[k] * nvec) * VB;
        poff[k] = pr[k] * row_bytes + pvb[k];
    }
    const bool full = G * VPL == nvec;  // every lane of the group holds VPL vectors of the row
    const uint32_t n_long = P.counters->n_long;
    const int64_t nnz = *P.nnz;
    const uint64_t row_mask = (P.row_bits >= 64) ? ~0ull : ((1ull << P.row_bits) - 1ull);
    const acc_t<T> eta = (acc_t<T>)P.eta;
    for (uint32_t j = blockIdx.x; j < n_long; j += gridDim.x) {
        const LongRec lr = P.longs[j];
        const BucketRec rec = P.recs[lr.bucket];
        const int64_t start = rec.start;
        const int64_t stop = ((int64_t)lr.bucket + 1 < nnz) ? (int64_t)P.recs[lr.bucket + 1].start : P.n_total;
        const int slot = (int)(rec.key >> P.row_bits) - P.slot0;
        const UpdDesc& d = P.item[slot];
        char* row = const_cast<char*>(row_ptr(d.table, (int64_t)(rec.key & row_mask) + 1));
        const int nbatch = (int)((stop - start + rows_per_stage - 1) / rows_per_stage);
        // producers: delta columns of my pieces' members in batch b.  The loads are issued kMapAhead batches before
        // the batch is requested (a ring of register sets): one iteration is shorter than a load's round trip
        constexpr int kMapAhead = 4;
        int32_t col[kMapAhead][kStrictPieces];
        const int32_t* bmap = P.map + start;
        const int members = (int)(stop - start);
        auto load_map = [&](int b, int32_t (&c)[kStrictPieces]) {
            const int rows = min(rows_per_stage, members - b * rows_per_stage);  // <= 0 past the last batch
#pragma unroll
            for (int k = 0; k < kStrictPieces; ++k) c[k] = pr[k] < rows ? __ldg(bmap + b * rows_per_stage + pr[k]) : 0;
        };
        // producers: member rows of batch b -> stage b % kStrictStages, one VB-byte piece per thread and step
        const char* dbase = d.delta;
        const int64_t ldb = d.ld_delta_bytes;
        auto issue = [&](int b, const int32_t (&c)[kStrictPieces]) {
            const int rows = min(rows_per_stage, members - b * rows_per_stage);
            char* stage = s_rows + (b % kStrictStages) * (rows_per_stage * row_bytes);
#pragma unroll
            for (int k = 0; k < kStrictPieces; ++k)
                if (pr[k] < rows) cp_async<VB>(stage + poff[k], dbase + (int64_t)c[k] * ldb + pvb[k]);
            asm volatile("cp.async.commit_group;" ::: "memory");  // one group per batch, empty ones included
        };
        A acc[VPL];
        V old[VPL];
        acc_t<T> st_old = acc_t<T>(0);
        acc_t<T>* state = nullptr;
        bool on[VPL];
#pragma unroll
        for (int p = 0; p < VPL; ++p) {
            acc_fill(acc[p], acc_t<T>(0));  // accum = zero, then += members in order
            on[p] = consumer && gl + p * G < nvec;
            if (on[p]) ld_plain<VB>(&old[p], row + (size_t)(gl + p * G) * VB);
        }
        if (OPT == kOptAdagrad && consumer) {
            state = (acc_t<T>*)P.state[slot] + (int64_t)(rec.key & row_mask);
            st_old = *state;
        }
        for (int b = 0; b < kStrictStages - 1; ++b) {  // the first stages: request as soon as the columns are here
            load_map(b, col[0]);
            issue(b, col[0]);
        }
#pragma unroll
        for (int q = 0; q < kMapAhead; ++q) load_map(kStrictStages - 1 + q, col[q]);
        for (int b0 = 0; b0 < nbatch; b0 += kMapAhead) {
#pragma unroll
          for (int q = 0; q < kMapAhead; ++q) {  // batch b0 + q; every thread runs all kMapAhead steps (barriers)
            const int b = b0 + q;
            asm volatile("cp.async.wait_group %0;" ::"n"(kStrictStages - 2) : "memory");  // my pieces of batch b have landed
            __syncthreads();  // everybody's have; the stage of batch b - 1 is free again
            issue(b + kStrictStages - 1, col[q]);                     // columns loaded kMapAhead steps ago
            load_map(b + kStrictStages - 1 + kMapAhead, col[q]);      // in flight for the next kMapAhead steps
            if (consumer && b < nbatch) {
                const int rows = min(rows_per_stage, members - b * rows_per_stage);
                const char* stage = s_rows + (b % kStrictStages) * (rows_per_stage * row_bytes);
                // U rows from shared memory into registers, then their additions in order: the shared-memory latency is
                // paid once per U rows instead of once per row
                constexpr int U = (8 / VPL) > 1 ? (8 / VPL) : 1;
                const char* mine_s = stage + (size_t)gl * VB;
                int r = 0;
                if (full) {  // no lane predicates: U * VPL independent loads, then the ordered additions
                    for (; r + U <= rows; r += U) {
                        V v[U][VPL];
#pragma unroll
                        for (int u = 0; u < U; ++u)
#pragma unroll
                            for (int p = 0; p < VPL; ++p) lds_vec<VB>(&v[u][p], mine_s + (r + u) * row_bytes + p * G * VB);
#pragma unroll
                        for (int u = 0; u < U; ++u)
#pragma unroll
                            for (int p = 0; p < VPL; ++p) acc_add(acc[p], v[u][p]);
                    }
                }
                for (; r < rows; ++r) {
#pragma unroll
                    for (int p = 0; p < VPL; ++p) {
                        if (on[p]) {
                            const V v = *(const V*)(mine_s + (size_t)r * row_bytes + (size_t)p * G * VB);
                            acc_add(acc[p], v);
                        }
                    }
                }
            }
          }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (consumer) {
            if constexpr (OPT == kOptSgd) {
#pragma unroll
                for (int p = 0; p < VPL; ++p) {
                    if (on[p]) {
                        V out;
#pragma unroll
                        for (int k = 0; k < V::NE; ++k) out.e[k] = sgd_apply<T>(old[p].e[k], acc[p].e[k], eta, d.table.pad != 0);
                        st_plain<VB>(row + (size_t)(gl + p * G) * VB, &out);
                    }
                }
            } else {
                V out[VPL];
                adagrad_apply<T, VB, VPL>(out, old, acc, on, state, st_old, eta, (acc_t<T>)P.eps, nvec * V::NE, G, gl,
                                          group_mask(G, threadIdx.x & 31));
#pragma unroll
                for (int p = 0; p < VPL; ++p)
                    if (on[p]) st_plain<VB>(row + (size_t)(gl + p * G) * VB, &out[p]);
            }
        }
        __syncthreads();  // the stages are reused by the next bucket
    }
}

// The same job SLICED BY FEATURES (SGD; the default).  The strict order binds the members of ONE feature element into a
// chain; different elements are independent.  One CTA streaming whole 512-byte rows is bound by what a single SM pulls
// from L2 / HBM and by four adds per member in its one adding warp (measured: 15 ns per member), far from the chain's
// own cost (one dependent add per member: 4 cycles = 2 ns).  So a long bucket becomes ceil(dim / 16) jobs, one per
// slice of 16 consecutive elements; the slices of a bucket run on different SMs at the same time.  In a job
//   * warps 1..7 stream that slice of every member row into shared memory (cp.async, VB-byte pieces, 16 KB stages of
//     up to 256 members) and warp 0 adds: lane l owns element l of the slice -- one shared-memory load and ONE dependent
//     add per member, the loads of the next 16 members in flight beside the adds (tools/ubench_chain.cu: 4.8 cycles per
//     member for that loop alone; in the kernel 6.2 + 300 per batch = 7.4 cycles per member, clock64-instrumented with
//     -DETB_SLICE_PROFILE).
//   * The members' delta columns (the map) travel like the rows: batch b's copies also fetch the map of batch b + S into
//     a ring of 2 S slots, so it has landed when the stage is handed back for batch b + S.  (Prefetched into registers,
//     every batch waited for the NEWEST map load -- the loads share scoreboards -- i.e. one DRAM round trip per batch.)
//   * Stages are handed over through mbarriers: s_full[st] gets one arrival per producer thread, delivered by the copy
//     unit when that thread's pieces have landed (cp.async.mbarrier.arrive.noinc: nobody waits for a landing but the
//     adding warp); s_empty[st] one arrival from the adding warp.  The ring runs on across the CTA's jobs: the
//     producers are already loading the next job's first batches while warp 0 finishes the current one.
// Any row length (slices are independent), any alignment class, any floating-point element type.
// Measured (C3, hottest row 45 280 members): 0.80 ms -> 0.19 ms for the long buckets; a variant with the stage transposed
// in quads of members (one 16-byte shared-memory load per 4 members, 4.3 cycles per member in the microbenchmark) was
// built and was bound by its 4-byte cp.async producers instead (27 cycles per member): not kept.
#ifndef ETB_SLICE_STAGES
#define ETB_SLICE_STAGES 4
#endif
#ifndef ETB_SLICE_CTAS
#define ETB_SLICE_CTAS 3  /* CTAs per SM of the grid = what fits (73 KB of shared memory, 76 registers): one wave, every job starts at once; 4 measured the same */
#endif
#ifndef ETB_SLICE_U
#define ETB_SLICE_U 16
#endif
#ifndef ETB_SLICE_ELEMS
#define ETB_SLICE_ELEMS 16  /* 16 elements x 256 members per 16 KB stage: 7.4 cycles per member; 32 x 128: 8.6; 8 x 512: 7.2 */
#endif
constexpr int kSliceElems = ETB_SLICE_ELEMS, kSliceStages = ETB_SLICE_STAGES, kSliceStageBytes = 16 * 1024,
              kSliceMaxRows = kSliceStageBytes / (kSliceElems * 4), kSlicePieces = 5;
static_assert(kSliceElems == 8 || kSliceElems == 16 || kSliceElems == 32, "a slice is a fraction of a warp");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    } while (!ok);
}
// the mbarrier gets one arrival from this thread once all cp.async it has issued so far have landed (no waiting here)
__device__ __forceinline__ void cp_async_arrive(uint64_t* b) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}
template <typename T>
__device__ __forceinline__ void lds_elem(T* out, const char* p) {  // volatile: the adding warp wants them issued in a batch
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    if constexpr (sizeof(T) == 8) {
        unsigned long long v;
        asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a));
        *(unsigned long long*)out = v;
    } else if constexpr (sizeof(T) == 4) {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
        *(uint32_t*)out = v;
    } else {
        unsigned short v;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
        *(unsigned short*)out = v;
    }
}

template <typename T, int VB>
__global__ void __launch_bounds__(kUThreads)
long_strict_sliced_kernel(const __grid_constant__ UpdParams P) {
    pdl_begin();
    using A = acc_t<T>;
    extern __shared__ __align__(16) char s_rows[];  // [S][R][SB]
    constexpr int kProducers = kUThreads - 32;
    constexpr int SB = kSliceElems * (int)sizeof(T);  // bytes of a full slice (64 / 128 / 256): the row pitch of a stage
    constexpr int U = ETB_SLICE_U;
    constexpr int S = kSliceStages;
    __shared__ __align__(8) uint64_t s_full[S], s_empty[S];
    __shared__ int32_t s_map[2 * S][kSliceMaxRows];
    const int row_bytes = P.nvec * VB;
    const int dim = row_bytes / (int)sizeof(T);
    const int nsl = (dim + kSliceElems - 1) / kSliceElems;
    const int lane = threadIdx.x & 31;
    const bool consumer = threadIdx.x < 32;  // warp 0 adds, warps 1..7 load
    const uint64_t n_jobs = (uint64_t)P.counters->n_long * (uint64_t)nsl;
    const uint64_t row_mask = (P.row_bits >= 64) ? ~0ull : ((1ull << P.row_bits) - 1ull);
    const A eta = (A)P.eta;
    if (threadIdx.x < S) {
        mbar_init(&s_full[threadIdx.x], kProducers);
        mbar_init(&s_empty[threadIdx.x], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // gb counts this CTA's batches over all its jobs: batch gb lives in stage gb % S and is that stage's (gb / S)-th use
    // (= the barrier phase).  Both sides walk the same (job, batch) sequence, each at its own pace.
    uint32_t gb = 0;
    uint4 lr_next = make_uint4(0, 0, 0, 0);  // the next job's record is fetched while this job runs
    if (blockIdx.x < n_jobs) lr_next = __ldg((const uint4*)(P.longs + blockIdx.x / (uint32_t)nsl));
    for (uint64_t job = blockIdx.x; job < n_jobs; job += gridDim.x) {
        const uint64_t j = job / (uint64_t)nsl;
        const int sl = (int)(job - j * (uint64_t)nsl);
        const int sbytes = min(SB, row_bytes - sl * SB);  // this slice's bytes of a row (a multiple of VB)
        const int pps = sbytes / VB;                      // pieces per member row
        const int R = min(min(kSliceMaxRows, kSliceStageBytes / SB), kSlicePieces * kProducers / pps);
        const uint4 lr = lr_next;  // {start, members, key} (register_sliced_long_bucket)
        if (job + gridDim.x < n_jobs) lr_next = __ldg((const uint4*)(P.longs + (job + gridDim.x) / (uint64_t)nsl));
        const int64_t start = lr.x;
        const int members = (int)lr.y;
        const uint64_t key = ((uint64_t)lr.w << 32) | lr.z;
        const int slot = (int)(key >> P.row_bits) - P.slot0;
        const UpdDesc& d = P.item[slot];
        const int nbatch = (members + R - 1) / R;
        if (consumer) {
            const bool on = lane < kSliceElems && sl * kSliceElems + lane < dim;
            T* mine_row = nullptr;
            T old = T();
            if (on) {
                mine_row = (T*)const_cast<char*>(row_ptr(d.table, (int64_t)(key & row_mask) + 1)) + sl * kSliceElems + lane;
                old = *mine_row;
            }
            A acc = A(0);  // accum = zero, then += members in order
#ifdef ETB_SLICE_PROFILE
            long long prof_wait = 0, prof_add = 0;
#endif
            for (int b = 0; b < nbatch; ++b, ++gb) {
                const int st = (int)(gb % S);
#ifdef ETB_SLICE_PROFILE
                const long long tp0 = clock64();
#endif
                mbar_wait(&s_full[st], (gb / S) & 1u);  // batch b has landed
#ifdef ETB_SLICE_PROFILE
                const long long tp1 = clock64();
                prof_wait += tp1 - tp0;
#endif
                const int rows = min(R, members - b * R);
                const char* mine = s_rows + st * kSliceStageBytes + min(lane, kSliceElems - 1) * (int)sizeof(T);
                // Groups of U members, double buffered: the loads of one group are in flight beside the other group's
                // chain of adds.  Every load of the batch is issued at least one group ahead of its add -- also the last
                // full group and the partial one (a member-by-member tail pays the shared-memory latency per member:
                // measured 9.2 cycles per member instead of 4.9, tools/ubench_chain.cu).
                T va[U], vb2[U];
                int r = 0;
                const int left = rows % U;  // members of the partial group, loaded (predicated) beside the last full one
                auto load_group = [&](T (&v)[U], int r0) {
#pragma unroll
                    for (int u = 0; u < U; ++u) lds_elem<T>(&v[u], mine + (r0 + u) * SB);
                };
                auto load_partial = [&](T (&v)[U], int r0) {
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (u < left) lds_elem<T>(&v[u], mine + (r0 + u) * SB);
                };
                auto add_group = [&](const T (&v)[U]) {
#pragma unroll
                    for (int u = 0; u < U; ++u) acc = acc + to_acc<T>(v[u]);
                };
                auto add_partial = [&](const T (&v)[U]) {
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (u < left) acc = acc + to_acc<T>(v[u]);
                };
                if (rows < U) {
                    load_partial(va, 0);
                    add_partial(va);
                } else {
                    load_group(va, 0);
                    for (;;) {  // va holds the group at r
                        if (r + 2 * U > rows) {  // the last full group
                            if (left) load_partial(vb2, r + U);
                            add_group(va);
                            if (left) add_partial(vb2);
                            break;
                        }
                        load_group(vb2, r + U);
                        add_group(va);
                        r += U;  // vb2 holds the group at r
                        if (r + 2 * U > rows) {
                            if (left) load_partial(va, r + U);
                            add_group(vb2);
                            if (left) add_partial(va);
                            break;
                        }
                        load_group(va, r + U);
                        add_group(vb2);
                        r += U;
                    }
                }
                __syncwarp();  // the stage may be overwritten (by this CTA's batch gb + S, if there is one)
                if (lane == 0) mbar_arrive(&s_empty[st]);
#ifdef ETB_SLICE_PROFILE
                prof_add += clock64() - tp1;
#endif
            }
#ifdef ETB_SLICE_PROFILE
            if (lane == 0 && members > 20000)
                printf("job %llu slice %d members %d batches %d: wait %lld add %lld cycles\n", (unsigned long long)job, sl, members, nbatch, prof_wait, prof_add);
#endif
            if (on) *mine_row = sgd_apply<T>(old, acc, eta, d.table.pad != 0);
        } else {
            // a producer's pieces of a batch: piece pc = (tid - 32) + k * 224 is piece pc % pps of the batch's row pc / pps
            int pr[kSlicePieces], poff[kSlicePieces], pvb[kSlicePieces];
            const int t = (int)threadIdx.x - 32;
#pragma unroll
            for (int k = 0; k < kSlicePieces; ++k) {
                const int pc = t + k * kProducers;
                pr[k] = pc / pps;  // rows >= R never exist
                pvb[k] = (pc - pr[k] * pps) * VB;
                poff[k] = pr[k] * SB + pvb[k];
            }
            const int32_t* bmap = P.map + start;
            const char* dbase = d.delta + (size_t)sl * SB;
            const int64_t ldb = d.ld_delta_bytes;
            // the map of the job's first S batches: one round trip per job (their slots were last read 2 S batches ago)
            for (int i = t; i < min(members, S * R); i += kProducers) s_map[(gb + i / R) % (2 * S)][i % R] = __ldg(bmap + i);
            asm volatile("bar.sync 1, %0;" ::"n"(kProducers) : "memory");  // the producer warps only
            for (int b = 0; b < nbatch; ++b, ++gb) {
                const int st = (int)(gb % S);
                if (gb >= S) mbar_wait(&s_empty[st], (gb / S - 1) & 1u);  // the stage's previous batch has been added
                const int rows = min(R, members - b * R);
                const int32_t* m = s_map[gb % (2 * S)];  // filled above or with this CTA's batch gb - S
                int32_t c[kSlicePieces];
#pragma unroll
                for (int k = 0; k < kSlicePieces; ++k) c[k] = pr[k] < rows ? m[pr[k]] : 0;
                char* stage = s_rows + st * kSliceStageBytes;
#pragma unroll
                for (int k = 0; k < kSlicePieces; ++k)
                    if (pr[k] < rows) cp_async<VB>(stage + poff[k], dbase + (int64_t)c[k] * ldb + pvb[k]);
                for (int i = t; i < R && (b + S) * R + i < members; i += kProducers)  // the map of batch b + S
                    cp_async<4>(&s_map[(gb + S) % (2 * S)][i], bmap + (b + S) * R + i);
                cp_async_arrive(&s_full[st]);  // every producer thread, with or without pieces in this batch
            }
        }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
}

// The same job with Blackwell's bulk-copy engine (rows of 16-byte vectors): warp 1 asks the TMA unit for one member
// row per lane -- cp.async.bulk, 32 rows = one stage, completion counted in bytes on the stage's mbarrier -- and
// warp 0 adds the rows of a stage once its barrier has flipped, then hands the stage back through a second
// mbarrier.  No thread copies data or computes piece addresses; 6 stages of 32 rows (up to 80 KB) are in flight.
// MEASURED SLOWER than the cp.async kernel above (C3: 2.08 vs 1.01 ms): a bulk request per 512-byte row pays the
// copy engine's per-request cost 45 000 times.  Kept behind ETB_STRICT_BULK=1 as the evidence.
constexpr int kBulkStages = 6, kBulkRows = 32, kBulkThreads = 64;

__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <typename T, int VPL, int OPT>
__global__ void __launch_bounds__(kBulkThreads)
long_strict_bulk_kernel(const __grid_constant__ UpdParams P) {
    pdl_begin();
    constexpr int VB = 16;
    using V = Vec<T, VB>;
    using A = AccVec<T, VB>;
    extern __shared__ __align__(128) char s_rows[];  // [kBulkStages][kBulkRows][row_bytes]
    __shared__ __align__(8) uint64_t s_full[kBulkStages], s_empty[kBulkStages];
    const int G = P.G, nvec = P.nvec;
    const int row_bytes = nvec * VB;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane & (G - 1);
    const bool adder = warp == 0 && lane < G;  // group 0 = the first G lanes of warp 0
    const bool full = G * VPL == nvec;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < kBulkStages; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&s_empty[i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t n_long = P.counters->n_long;
    const int64_t nnz = *P.nnz;
    const uint64_t row_mask = (P.row_bits >= 64) ? ~0ull : ((1ull << P.row_bits) - 1ull);
    const acc_t<T> eta = (acc_t<T>)P.eta;
    uint32_t it = 0;  // batches handled so far by this CTA (both warps count alike): stage = it % S, use = it / S
    for (uint32_t j = blockIdx.x; j < n_long; j += gridDim.x) {
        const LongRec lr = P.longs[j];
        const BucketRec rec = P.recs[lr.bucket];
        const int64_t start = rec.start;
        const int64_t stop = ((int64_t)lr.bucket + 1 < nnz) ? (int64_t)P.recs[lr.bucket + 1].start : P.n_total;
        const int slot = (int)(rec.key >> P.row_bits) - P.slot0;
        const UpdDesc& d = P.item[slot];
        const int members = (int)(stop - start);
        const int nbatch = (members + kBulkRows - 1) / kBulkRows;
        if (warp == 1) {
            // ---- producer: one member row per lane and batch; the delta column of the next batch is fetched while
            // this one is issued
            const int32_t* bmap = P.map + start;
            constexpr int kAhead = 8;  // the delta column of a batch is fetched 8 batches before it is requested
            int32_t col[kAhead];
#pragma unroll
            for (int q = 0; q < kAhead; ++q) col[q] = q * kBulkRows + lane < members ? __ldg(bmap + q * kBulkRows + lane) : 0;
            for (int b0 = 0; b0 < nbatch; b0 += kAhead) {
#pragma unroll
                for (int q = 0; q < kAhead; ++q) {
                    const int b = b0 + q;
                    if (b < nbatch) {
                        const uint32_t stage = it % kBulkStages, use = it / kBulkStages;
                        const int rows = min(kBulkRows, members - b * kBulkRows);
                        if (use > 0) mbar_wait(&s_empty[stage], (use - 1) & 1u);  // the adders are done with this stage
                        if (lane == 0) mbar_expect_tx(&s_full[stage], (uint32_t)(rows * row_bytes));
                        __syncwarp();
                        if (lane < rows)
                            bulk_g2s(s_rows + ((size_t)stage * kBulkRows + lane) * row_bytes,
                                     d.delta + (int64_t)col[q] * d.ld_delta_bytes, (uint32_t)row_bytes, &s_full[stage]);
                        const int nxt = (b + kAhead) * kBulkRows + lane;
                        col[q] = nxt < members ? __ldg(bmap + nxt) : 0;
                        ++it;
                    }
                }
            }
        } else {
            // ---- consumer: the member rows of a stage, added in order
            char* row = const_cast<char*>(row_ptr(d.table, (int64_t)(rec.key & row_mask) + 1));
            A acc[VPL];
            V old[VPL];
            acc_t<T> st_old = acc_t<T>(0);
            acc_t<T>* state = nullptr;
            bool on[VPL];
#pragma unroll
            for (int p = 0; p < VPL; ++p) {
                acc_fill(acc[p], acc_t<T>(0));  // accum = zero, then += members in order
                on[p] = adder && gl + p * G < nvec;
                if (on[p]) ld_plain<VB>(&old[p], row + (size_t)(gl + p * G) * VB);
            }
            if (OPT == kOptAdagrad && adder) {
                state = (acc_t<T>*)P.state[slot] + (int64_t)(rec.key & row_mask);
                st_old = *state;
            }
            for (int b = 0; b < nbatch; ++b, ++it) {
                const uint32_t stage = it % kBulkStages, use = it / kBulkStages;
                const int rows = min(kBulkRows, members - b * kBulkRows);
                mbar_wait(&s_full[stage], use & 1u);  // the bytes of this stage have landed
                if (adder) {
                    constexpr int U = (8 / VPL) > 1 ? (8 / VPL) : 1;
                    const char* mine_s = s_rows + (size_t)stage * kBulkRows * row_bytes + (size_t)gl * VB;
                    int r = 0;
                    if (full) {
                        for (; r + U <= rows; r += U) {
                            V v[U][VPL];
#pragma unroll
                            for (int u = 0; u < U; ++u)
#pragma unroll
                                for (int p = 0; p < VPL; ++p) lds_vec<VB>(&v[u][p], mine_s + (r + u) * row_bytes + p * G * VB);
#pragma unroll
                            for (int u = 0; u < U; ++u)
#pragma unroll
                                for (int p = 0; p < VPL; ++p) acc_add(acc[p], v[u][p]);
                        }
                    }
                    for (; r < rows; ++r) {
#pragma unroll
                        for (int p = 0; p < VPL; ++p) {
                            if (on[p]) {
                                V v;
                                lds_vec<VB>(&v, mine_s + r * row_bytes + p * G * VB);
                                acc_add(acc[p], v);
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_empty[stage]);  // the stage may be refilled
            }
            if (adder) {
                if constexpr (OPT == kOptSgd) {
#pragma unroll
                    for (int p = 0; p < VPL; ++p) {
                        if (on[p]) {
                            V out;
#pragma unroll
                            for (int k = 0; k < V::NE; ++k) out.e[k] = sgd_apply<T>(old[p].e[k], acc[p].e[k], eta, d.table.pad != 0);
                            st_plain<VB>(row + (size_t)(gl + p * G) * VB, &out);
                        }
                    }
                } else {
                    V out[VPL];
                    adagrad_apply<T, VB, VPL>(out, old, acc, on, state, st_old, eta, (acc_t<T>)P.eps, nvec * V::NE, G, gl,
                                              group_mask(G, lane));
#pragma unroll
                    for (int p = 0; p < VPL; ++p)
                        if (on[p]) st_plain<VB>(row + (size_t)(gl + p * G) * VB, &out[p]);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------ host
static int pick_vb_update(const etb_update_item& it) {
    const size_t es = elt_bytes(it.table.elt);
    const size_t rowbytes = (size_t)it.table.dim * es;
    const size_t stride = (size_t)it.table.ld * es;
    const size_t ldd = (size_t)it.ld_delta * es;
    for (size_t vb : {(size_t)16, (size_t)8}) {
        if (vb < es) continue;
        if (rowbytes % vb == 0 && stride % vb == 0 && ldd % vb == 0 && ((uintptr_t)it.table.base % vb) == 0 &&
            ((uintptr_t)it.delta % vb) == 0)
            return (int)vb;
    }
    if (es == 2)  // half-precision rows must at least be 4-byte aligned
        return (rowbytes % 4 == 0 && stride % 4 == 0 && ldd % 4 == 0 && ((uintptr_t)it.table.base % 4) == 0 &&
                ((uintptr_t)it.delta % 4) == 0) ? 4 : 0;
    return es == 8 ? 8 : 4;
}

struct UpdClass {
    int32_t elt, vb, vpl, G, nvec;
    bool operator==(const UpdClass& o) const {
        return elt == o.elt && vb == o.vb && vpl == o.vpl && G == o.G && nvec == o.nvec;
    }
};

static UpdClass classify_update(const etb_update_item& it) {
    UpdClass c;
    c.elt = it.table.elt;
    c.vb = pick_vb_update(it);
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

enum { kKernelMain = 0, kKernelTasks = 1, kKernelCombine = 2, kKernelStrictLong = 3 };

// fork / join of the task kernel (update_impl): one side stream and two events per host thread and device
struct ForkJoin {
    cudaStream_t side = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
static ForkJoin* fork_join() {
    static thread_local ForkJoin fj[16];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    ForkJoin& f = fj[dev];
    if (!f.side) {
        if (cudaStreamCreateWithFlags(&f.side, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&f.fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&f.join, cudaEventDisableTiming) != cudaSuccess) {
            f.side = nullptr;
            return nullptr;
        }
    }
    return &f;
}
static bool update_fork_enabled() {
    static const bool on = [] { const char* e = getenv("ETB_UPDATE_FORK"); return !(e && e[0] == '0'); }();
    return on;
}

// ETB_STRICT_SLICED=0: round 2's one-CTA-per-bucket kernel for the long buckets of the strict order (kept for comparison)
static bool strict_sliced() {
    static const bool on = [] { const char* e = getenv("ETB_STRICT_SLICED"); return !(e && e[0] == '0'); }();
    return on;
}

template <typename T, int VB, int VPL, int OPT>
static void launch_update_one(int which, int grid, cudaStream_t s, const UpdParams& P) {
    if (which == kKernelMain) launch_k(sgd_update_kernel<T, VB, VPL, OPT>, grid, kUThreads, 0, s, P);
    else if (which == kKernelTasks) launch_k(bucket_tasks_kernel<T, VB, VPL, OPT>, grid, kUThreads, 0, s, P);
    else if (which == kKernelStrictLong && VB == 16 && P.nvec * VB * kBulkRows * kBulkStages <= 200 * 1024 &&
             getenv("ETB_STRICT_BULK")) {  // measured 2x slower than the cp.async variant for 512-byte rows: opt-in only
        if constexpr (VB == 16) {  // the bulk-copy (TMA) variant
            const int smem = P.nvec * VB * kBulkRows * kBulkStages;
            static thread_local int configured = 0;
            if (smem > configured) {
                cudaFuncSetAttribute(long_strict_bulk_kernel<T, VPL, OPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                configured = smem;
            }
            launch_k(long_strict_bulk_kernel<T, VPL, OPT>, grid, kBulkThreads, smem, s, P);
        }
    } else if (which == kKernelStrictLong && P.strict_long == 2) {  // the feature-sliced kernel (SGD)
        constexpr int kSmem = kSliceStages * kSliceStageBytes;  // 64 KB of dynamic shared memory: three CTAs per SM
        static thread_local bool configured = false;
        if (!configured) {
            cudaFuncSetAttribute(long_strict_sliced_kernel<T, VB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
            configured = true;
        }
        launch_k(long_strict_sliced_kernel<T, VB>, grid, kUThreads, kSmem, s, P);
    } else if (which == kKernelStrictLong) {
        constexpr int kSmem = kStrictStages * kStrictStageBytes;  // 96 KB of dynamic shared memory: opt in once
        static thread_local bool configured = false;
        if (!configured) {
            cudaFuncSetAttribute(long_strict_kernel<T, VB, VPL, OPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
            configured = true;
        }
        launch_k(long_strict_kernel<T, VB, VPL, OPT>, grid, kUThreads, kSmem, s, P);
    } else launch_k(long_combine_kernel<T, VB, VPL, OPT>, grid, kUThreads, 0, s, P);
}

template <typename T, int VB, int OPT>
static void launch_update_vpl(int which, int vpl, int grid, cudaStream_t s, const UpdParams& P) {
    switch (vpl) {
        case 1: launch_update_one<T, VB, 1, OPT>(which, grid, s, P); break;
        case 2: launch_update_one<T, VB, 2, OPT>(which, grid, s, P); break;
        default: launch_update_one<T, VB, 4, OPT>(which, grid, s, P); break;
    }
}

template <typename T, int OPT>
static void launch_update_vb(int which, const UpdClass& c, int grid, cudaStream_t s, const UpdParams& P) {
    if constexpr (sizeof(T) <= 4) {
        if (c.vb == 4) return launch_update_vpl<T, 4, OPT>(which, c.vpl, grid, s, P);
    }
    if (c.vb == 8) return launch_update_vpl<T, 8, OPT>(which, c.vpl, grid, s, P);
    return launch_update_vpl<T, 16, OPT>(which, c.vpl, grid, s, P);
}

template <typename T, int OPT>
static bool launch_update_exact(const UpdClass& c, int grid, cudaStream_t s, const UpdParams& P) {
    constexpr int UB = ETB_UPDATE_EXACT_UB;
    if (c.vb != 16 || c.nvec > c.G * c.vpl) return false;  // 16-byte vectors, row fits one pass
    if (sizeof(T) == 2 && c.vpl > 1) return false;           // half types: 8 accumulators per vector, VPL > 1 would spill
#define ETB_EXACT(VPLV, GV, UBV)                                                         \
    if (c.vpl == VPLV && c.G == GV) {                                                    \
        launch_k(sgd_update_exact_kernel<T, VPLV, GV, (UBV), OPT>, grid, kUThreads, 0, s, P);  \
        return true;                                                                     \
    }
    ETB_EXACT(1, 32, UB) ETB_EXACT(1, 16, UB) ETB_EXACT(1, 8, UB) ETB_EXACT(1, 4, UB < 4 ? UB : 4)
    ETB_EXACT(2, 32, UB > 2 ? UB / 2 : 1) ETB_EXACT(4, 32, 1)
#undef ETB_EXACT
    return false;
}

template <int OPT>
static void launch_update_opt(int which, const UpdClass& c, int grid, cudaStream_t s, const UpdParams& P) {
    if (which == kKernelMain && ETB_UPDATE_USE_EXACT) {
        bool done;
        switch (c.elt) {
#ifndef ETB_DEV_F32_ONLY
            case ETB_F16: done = launch_update_exact<__half, OPT>(c, grid, s, P); break;
            case ETB_BF16: done = launch_update_exact<__nv_bfloat16, OPT>(c, grid, s, P); break;
            case ETB_F64: done = launch_update_exact<double, OPT>(c, grid, s, P); break;
#endif
            default: done = launch_update_exact<float, OPT>(c, grid, s, P); break;
        }
        if (done) return;
    }
#ifdef ETB_DEV_F32_ONLY  // experiment builds (tools/): Float32 tables only, a quarter of the compile time
    return launch_update_vb<float, OPT>(which, c, grid, s, P);
#endif
    switch (c.elt) {
        case ETB_F32: launch_update_vb<float, OPT>(which, c, grid, s, P); break;
        case ETB_F16: launch_update_vb<__half, OPT>(which, c, grid, s, P); break;
        case ETB_BF16: launch_update_vb<__nv_bfloat16, OPT>(which, c, grid, s, P); break;
        default: launch_update_vb<double, OPT>(which, c, grid, s, P); break;
    }
}

static void launch_update(int opt, int which, const UpdClass& c, int grid, cudaStream_t s, const UpdParams& P) {
    if (opt == kOptAdagrad) launch_update_opt<kOptAdagrad>(which, c, grid, s, P);
    else launch_update_opt<kOptSgd>(which, c, grid, s, P);
}

static int32_t update_impl(const etb_index_view* view, const etb_update_item* items, int32_t n_items, double eta,
                           int32_t flags, cudaStream_t stream, int opt = kOptSgd, void* const* states = nullptr,
                           double eps = 0.0) {
    ETB_REQUIRE(view, "etb_sgd_update: null index view");
    ETB_REQUIRE(n_items >= 0 && (n_items == 0 || items), "etb_sgd_update: bad items");
    if (n_items == 0 || view->n_total == 0) return ETB_OK;
    IndexLayout L;  // the scratch region's layout is a pure function of the items
    if (int32_t st = make_layout(items, n_items, L)) return st;
    ETB_REQUIRE(L.n_total == view->n_total, "etb_sgd_update: items do not match the index view");
    std::vector<UpdClass> cls((size_t)n_items);
    for (int i = 0; i < n_items; ++i) {
        const etb_update_item& it = items[i];
        // update! on integer tables throws in the reference too (InexactError at
        // convert(eltype(table), opt.eta), src/sparseupdate.jl:173)
        ETB_REQUIRE(elt_is_float(it.table.elt), "etb_sgd_update: item %d: only floating-point tables can be updated", i);
        ETB_REQUIRE(it.batch == 0 || it.delta, "etb_sgd_update: item %d: null delta", i);
        ETB_REQUIRE(it.ld_delta >= it.table.dim, "etb_sgd_update: item %d: ld_delta < dim", i);
        cls[i] = classify_update(it);
        if (cls[i].vb == 0)
            return fail(ETB_ERR_UNSUPPORTED, "etb_sgd_update: item %d: half-precision rows must be 4-byte aligned (even dim, ld, ld_delta)", i);
        if (opt == kOptAdagrad) {
            ETB_REQUIRE(states && states[i], "etb_adagrad_update: item %d: null state vector", i);
            ETB_REQUIRE(it.table.nrows <= 0x7fffffffll, "etb_adagrad_update: item %d: more than 2^31 rows", i);
            if (cls[i].nvec > cls[i].G * cls[i].vpl)
                return fail(ETB_ERR_UNSUPPORTED, "etb_adagrad_update: item %d: rows of more than %d vectors of %d bytes are not supported",
                            i, cls[i].G * cls[i].vpl, cls[i].vb);
        }
    }
    ETB_REQUIRE(view->num_splits >= 0 && (view->num_splits == 0 || (view->this_split >= 1 && view->this_split <= view->num_splits)),
                "etb_sgd_update: bad IndexerView split %d of %d", view->this_split, view->num_splits);
    static thread_local UpdParams P;
    char* scratch = (char*)view->scratch;
    P.recs = (const BucketRec*)view->records;
    P.map = view->map;
    P.nnz = view->nnz;
    P.counters = (LongCounters*)scratch;
    P.longs = (LongRec*)(scratch + (L.off_long - L.off_counters));
    P.chunks = (ChunkRec*)(scratch + (L.off_chunks - L.off_counters));
    P.partials = scratch + (L.off_partials - L.off_counters);
    P.partial_pitch = (int64_t)L.partial_pitch;
    P.n_total = view->n_total;
    P.eta = eta;
    P.eps = eps;
    P.row_bits = view->row_bits;
    P.fma = (flags & ETB_UPDATE_FMA) ? 1 : 0;
    P.split_long = (flags & ETB_UPDATE_SPLIT_LONG) ? 1 : 0;
    P.num_splits = view->num_splits;
    P.this_split = view->this_split;
    // one launch per run of consecutive items sharing a kernel class (all tables of a DLRM
    // ensemble share it: one launch)
    for (int i0 = 0; i0 < n_items;) {
        const UpdClass c = cls[i0];
        int n = 0;
        while (i0 + n < n_items && n < kUMaxItems && cls[i0 + n] == c) {
            const etb_update_item& it = items[i0 + n];
            UpdDesc& d = P.item[n];
            d.table = make_dev_table(it.table);
            // per-table epilogue; the half types always take `row - eta*acc` in Float32 (see ETB_F16 in the header)
            d.table.pad = (((flags | it.flags) & ETB_UPDATE_FMA) && elt_bytes(it.table.elt) >= 4) ? 1u : 0u;
            d.delta = (const char*)it.delta;
            d.ld_delta_bytes = it.ld_delta * (int64_t)elt_bytes(it.table.elt);
            P.state[n] = opt == kOptAdagrad ? (char*)states[i0 + n] : nullptr;
            ++n;
        }
        P.slot0 = i0;
        P.nslots = n;
        P.G = c.G;
        P.nvec = c.nvec;
        P.vb = c.vb;
        // strict order: long buckets whose rows fit one pass (and one 32 KB stage) are streamed by long_strict_kernel
        // (the feature-sliced SGD kernel takes rows of any length)
        const bool sliced = opt == kOptSgd && strict_sliced();
        P.strict_long = P.split_long ? 0 : (sliced ? 2 : ((c.nvec <= c.G * c.vpl && c.nvec * c.vb <= kStrictStageBytes) ? 1 : 0));
        ETB_CUDA(cudaMemsetAsync(P.counters, 0, sizeof(LongCounters), stream));
        const int64_t buckets_per_block = (kUThreads / 32) * 32 * ETB_UPDATE_RPL;  // one tile per warp
        const int grid = (int)((view->n_total + buckets_per_block - 1) / buckets_per_block);
        launch_update(opt, kKernelMain, c, grid, stream, P);
        ETB_LAUNCHED();
        const int64_t per_block = kUThreads / c.G;
        // Strict order, SGD: the task kernel (buckets of 5..128 members) and the sliced long-bucket kernel work on
        // disjoint buckets and both depend only on the main kernel's worklists, so the task kernel runs on an internal
        // side stream beside the long buckets (a fork and a join of events: capturable, two parallel branches in a CUDA
        // graph).  C3: the 45 us of tasks disappear behind the hottest row's chain.  ETB_UPDATE_FORK=0: one after the other.
        ForkJoin* fj = (P.strict_long == 2 && view->n_total > kLongThreshold && update_fork_enabled()) ? fork_join() : nullptr;
        auto launch_tasks = [&](cudaStream_t ts) -> int32_t {  // medium buckets + long-bucket chunks: one task kernel
            const int64_t max_tasks = view->n_total / (kShortMax + 1) + 1;
            const int gridT = (int)std::min<int64_t>((max_tasks + per_block - 1) / per_block, (int64_t)num_sms() * 8);
            launch_update(opt, kKernelTasks, c, gridT, ts, P);
            ETB_LAUNCHED();
            return ETB_OK;
        };
        if (fj) ETB_CUDA(cudaEventRecord(fj->fork, stream));  // behind the main kernel
        if (!fj && view->n_total > kShortMax)
            if (int32_t st = launch_tasks(stream)) return st;
        if (P.strict_long && view->n_total > kLongThreshold) {
            const int64_t max_long = view->n_total / kLongThreshold + 1;
            const int64_t nsl = sliced ? ((int64_t)c.nvec * c.vb / (int64_t)elt_bytes(c.elt) + kSliceElems - 1) / kSliceElems : 1;
            const int gridS = (int)std::min<int64_t>(max_long * nsl, (int64_t)num_sms() * (sliced ? ETB_SLICE_CTAS : 2));  // one CTA per job
            launch_update(opt, kKernelStrictLong, c, gridS, stream, P);
            ETB_LAUNCHED();
        }
        if (fj) {  // the long buckets were launched first: their CTAs take the SMs, the tasks fill in as the short jobs end
            ETB_CUDA(cudaStreamWaitEvent(fj->side, fj->fork, 0));
            if (int32_t st = launch_tasks(fj->side)) return st;
            ETB_CUDA(cudaEventRecord(fj->join, fj->side));
        }
        if (P.split_long && view->n_total > kLongThreshold) {
            const int64_t max_long = view->n_total / kLongThreshold + 1;
            const int gridB = (int)std::min<int64_t>(max_long, (int64_t)num_sms() * 4);  // one CTA per long bucket
            launch_update(opt, kKernelCombine, c, gridB, stream, P);
            ETB_LAUNCHED();
        }
        if (fj) ETB_CUDA(cudaStreamWaitEvent(stream, fj->join, 0));  // the caller's stream continues behind both kernels
        i0 += n;
    }
    return ETB_OK;
}

// ------------------------------------------------------------------------------------ uncompress
template <typename T, typename IdxT>
__global__ void uncompress_kernel(T* dst, int64_t ld_dst, int dim, const T* delta, int64_t ld_delta, const IdxT* idx,
                                  int64_t bag, int64_t batch, int64_t ld_idx) {
    pdl_begin();
    // one thread per feature element walks every occurrence in order: deterministic, and the
    // same association as the reference's `columnview(dst, c) .+= update` loop
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= dim) return;
    const int64_t per = bag ? bag : 1;
    for (int64_t j = 0; j < batch; ++j) {
        const T g = delta[j * ld_delta + k];
        for (int64_t i = 0; i < per; ++i) {
            const int64_t c = (int64_t)(bag ? idx[j * ld_idx + i] : idx[j]);
            dst[(c - 1) * ld_dst + k] += g;
        }
    }
}

// ------------------------------------------------------------------------------------ a2a pack / unpack
struct A2ABlocks {
    int64_t rows[16];
    int64_t row_off[16];
    int64_t dense_ld_bytes[16];  // bytes between the columns of block r's destination (rows[r] * es when dense)
    char* dense[16];  // block r's (rows[r] x batch_local) matrix: local buffer or peer memory
};

// One CTA walks whole columns of block r (blockIdx.y): a column of the block is one contiguous run on both sides, so
// there is no division per vector, and a thread has up to four 16-byte loads in flight before its first store (the
// stores go over NVLink when the destination is peer memory).
template <int VB, bool UNPACK>
__global__ void __launch_bounds__(256)
a2a_copy_kernel(char* strided, int64_t ld_bytes, const __grid_constant__ A2ABlocks B, int64_t batch_local, int es) {
    pdl_begin();
    constexpr int U = 4;
    const int r = blockIdx.y;
    const int64_t row_bytes = B.rows[r] * es;
    const int vec_per_col = (int)(row_bytes / VB);
    char* dbase = B.dense[r];
    char* sbase = strided + B.row_off[r] * es;
    const int64_t dld = B.dense_ld_bytes[r];
    // L threads per column (a power of two, at most the CTA): narrow blocks put several columns side by side
    int L = 256;
    while (L > 1 && (L >> 1) >= vec_per_col) L >>= 1;
    const int cpi = 256 / L, cl = threadIdx.x / L, vl = threadIdx.x & (L - 1);
    for (int64_t col = (int64_t)blockIdx.x * cpi + cl; col < batch_local; col += (int64_t)gridDim.x * cpi) {
        char* s = sbase + col * ld_bytes;
        char* d = dbase + col * dld;
        char* from = UNPACK ? d : s;
        char* to = UNPACK ? s : d;
        for (int v0 = vl; v0 < vec_per_col; v0 += U * L) {
            Vec<uint32_t, VB> x[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (v0 + u * L < vec_per_col) ld_row<VB>(&x[u], from + (int64_t)(v0 + u * L) * VB);
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (v0 + u * L < vec_per_col) st_stream<VB>(to + (int64_t)(v0 + u * L) * VB, &x[u]);
        }
    }
}

// dense == nullptr: per-block destinations in dense_ptrs (peer memory); else blocks laid end to end in `dense`
static int32_t a2a_copy(bool unpack, void* strided, int64_t ld, void* dense, void* const* dense_ptrs, const int64_t* rows,
                        const int64_t* row_off, int32_t nranks, int64_t batch_local, int32_t elt, cudaStream_t stream,
                        const int64_t* dense_ld = nullptr) {
    launch_counter() = 0;
    ETB_REQUIRE(nranks >= 1 && nranks <= 16, "etb_a2a: nranks must be in 1..16");
    ETB_REQUIRE(elt_valid(elt), "etb_a2a: bad element type");
    ETB_REQUIRE(strided && (dense || dense_ptrs) && rows && row_off, "etb_a2a: null pointer");
    if (batch_local == 0) return ETB_OK;
    const int es = (int)elt_bytes(elt);
    A2ABlocks B;
    int64_t off = 0, max_rows = 0;
    int vb = 16;
    for (int r = 0; r < nranks; ++r) {
        B.rows[r] = rows[r];
        B.row_off[r] = row_off[r];
        B.dense[r] = dense ? (char*)dense + off * es : (char*)dense_ptrs[r];
        B.dense_ld_bytes[r] = (dense_ld ? dense_ld[r] : rows[r]) * es;
        ETB_REQUIRE(B.dense[r] || rows[r] == 0, "etb_a2a: null block pointer");
        ETB_REQUIRE(!dense_ld || dense_ld[r] >= rows[r], "etb_a2a: destination leading dimension smaller than the block");
        off += rows[r] * batch_local;
        max_rows = std::max(max_rows, rows[r]);
        while (vb > es && ((rows[r] * es) % vb || (row_off[r] * es) % vb || (uintptr_t)B.dense[r] % vb || B.dense_ld_bytes[r] % vb)) vb >>= 1;
    }
    while (vb > es && ((ld * es) % vb || (uintptr_t)strided % vb)) vb >>= 1;
    if (vb < 4) return fail(ETB_ERR_UNSUPPORTED, "etb_a2a: half-precision blocks must be 4-byte aligned (even row counts and offsets)");
    if (max_rows == 0) return ETB_OK;
    int lanes = 256;  // threads per column of the widest block (the kernel derives each block's own)
    while (lanes > 1 && (lanes >> 1) >= max_rows * es / vb) lanes >>= 1;
    const int64_t col_groups = (batch_local + 256 / lanes - 1) / (256 / lanes);
    dim3 grid((unsigned)std::min<int64_t>(col_groups, (int64_t)num_sms() * 8), (unsigned)nranks);
    const int64_t ldb = ld * es;
#define ETB_A2A(VBV)                                                                                          \
    if (unpack) launch_k(a2a_copy_kernel<VBV, true>, grid, 256, 0, stream, (char*)strided, ldb, B, batch_local, es); \
    else launch_k(a2a_copy_kernel<VBV, false>, grid, 256, 0, stream, (char*)strided, ldb, B, batch_local, es);
    if (vb == 16) { ETB_A2A(16) } else if (vb == 8) { ETB_A2A(8) } else { ETB_A2A(4) }
#undef ETB_A2A
    ETB_LAUNCHED();
    return ETB_OK;
}

// ------------------------------------------------------------------------------------ host-tier tables
// Admission (Update phase): one warp per bucket.  A bucket of a cached table whose row still lives on the host and
// that had >= min_count members takes the next free slot; the warp copies the row into it and only then publishes
// the slot (no other kernel touches the table meanwhile: stream order).
struct CacheItem {
    const char* host_rows;
    char* cache_rows;
    int32_t* slot_of_row;
    int32_t* row_of_slot;
    int32_t* cursor;
    int32_t* hist;
    int64_t row_stride, row_bytes;
    int32_t capacity, pad;
};
struct CacheParams {
    CacheItem item[kUMaxItems];
    const BucketRec* recs;
    const int64_t* nnz;
    int64_t n_total;
    int32_t row_bits, min_count, n_items;
};

// occurrence counts (clamped to the last bin) of the rows of host-tier tables that are not cached yet
__global__ void __launch_bounds__(256) cache_count_kernel(const __grid_constant__ CacheParams P) {
    pdl_begin();
    const int64_t nnz = *P.nnz;
    for (int64_t b = (int64_t)blockIdx.x * 256 + threadIdx.x; b < nnz; b += (int64_t)gridDim.x * 256) {
        const BucketRec rec = P.recs[b];
        const int slot_id = (int)(rec.key >> P.row_bits);
        if (slot_id >= P.n_items || !P.item[slot_id].slot_of_row) continue;
        const CacheItem& c = P.item[slot_id];
        const uint64_t row_mask = (P.row_bits >= 64) ? ~0ull : ((1ull << P.row_bits) - 1ull);
        if (c.slot_of_row[rec.key & row_mask] >= 0) continue;
        const int64_t stop = (b + 1 < nnz) ? (int64_t)P.recs[b + 1].start : P.n_total;
        atomicAdd(c.hist + (int)min((int64_t)ETB_CACHE_HIST_BINS - 1, stop - (int64_t)rec.start), 1);
    }
}

// the smallest count t >= min_count such that the uncached rows with >= t occurrences fit into the free slots
__device__ __forceinline__ int cache_threshold(const CacheItem& c, int min_count) {
    const int free_slots = c.capacity - min(*c.cursor, c.capacity);
    int t = ETB_CACHE_HIST_BINS - 1, fit = 0;
    for (int k = ETB_CACHE_HIST_BINS - 1; k >= min_count; --k) {
        fit += c.hist[k];
        if (fit > free_slots) break;
        t = k;
    }
    return t;
}

__global__ void __launch_bounds__(256) cache_admit_kernel(const __grid_constant__ CacheParams P) {
    pdl_begin();
    __shared__ int s_threshold[kUMaxItems];
    for (int i = threadIdx.x; i < P.n_items; i += 256)
        s_threshold[i] = P.item[i].slot_of_row ? cache_threshold(P.item[i], P.min_count) : 0x7fffffff;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t nnz = *P.nnz;
    const uint64_t row_mask = (P.row_bits >= 64) ? ~0ull : ((1ull << P.row_bits) - 1ull);
    for (int64_t b = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); b < nnz; b += (int64_t)gridDim.x * 8) {
        const BucketRec rec = P.recs[b];
        const int slot_id = (int)(rec.key >> P.row_bits);
        if (slot_id >= P.n_items) continue;
        const CacheItem& c = P.item[slot_id];
        if (!c.slot_of_row) continue;  // not a host-tier table
        const int64_t stop = (b + 1 < nnz) ? (int64_t)P.recs[b + 1].start : P.n_total;
        const int64_t row = (int64_t)(rec.key & row_mask);
        int s = -1;
        if (lane == 0 && stop - (int64_t)rec.start >= s_threshold[slot_id] && c.slot_of_row[row] < 0) {
            s = atomicAdd(c.cursor, 1);
            if (s >= c.capacity) {  // full: undo (the cursor never runs away)
                atomicSub(c.cursor, 1);
                s = -1;
            }
        }
        s = __shfl_sync(0xffffffffu, s, 0);
        if (s < 0) continue;
        const char* src = c.host_rows + row * c.row_stride;
        char* dst = c.cache_rows + (int64_t)s * c.row_stride;
        if (((c.row_bytes | c.row_stride | (int64_t)(uintptr_t)c.host_rows | (int64_t)(uintptr_t)c.cache_rows) & 15) == 0)
            for (int64_t o = lane * 16; o < c.row_bytes; o += 32 * 16) *(uint4*)(dst + o) = *(const uint4*)(src + o);
        else
            for (int64_t o = lane; o < c.row_bytes; o += 32) dst[o] = src[o];
        __syncwarp();
        if (lane == 0) {
            c.row_of_slot[s] = (int32_t)row;
            __threadfence();
            c.slot_of_row[row] = s;
        }
    }
}

// write every cached row back to the host table: one warp per slot
__global__ void __launch_bounds__(256) cache_flush_kernel(CacheItem c) {
    pdl_begin();
    const int lane = threadIdx.x & 31;
    const int used = min(*c.cursor, c.capacity);
    for (int s = blockIdx.x * 8 + (threadIdx.x >> 5); s < used; s += gridDim.x * 8) {
        const char* src = c.cache_rows + (int64_t)s * c.row_stride;
        char* dst = const_cast<char*>(c.host_rows) + (int64_t)c.row_of_slot[s] * c.row_stride;
        if (((c.row_bytes | c.row_stride | (int64_t)(uintptr_t)c.host_rows | (int64_t)(uintptr_t)c.cache_rows) & 15) == 0)
            for (int64_t o = lane * 16; o < c.row_bytes; o += 32 * 16) *(uint4*)(dst + o) = *(const uint4*)(src + o);
        else
            for (int64_t o = lane; o < c.row_bytes; o += 32) dst[o] = src[o];
    }
}

static bool make_cache_item(const etb_table& t, CacheItem& c) {
    memset(&c, 0, sizeof(c));
    if (!(t.chunks && t.shard_rows == ETB_TABLE_CACHED)) return false;
    const etb_cache_desc* d = (const etb_cache_desc*)t.chunks;
    c.host_rows = (const char*)t.base;
    c.cache_rows = (char*)d->rows;
    c.slot_of_row = d->slot_of_row;
    c.row_of_slot = d->row_of_slot;
    c.cursor = d->cursor;
    c.hist = d->hist;
    c.row_stride = (int64_t)t.ld * (int64_t)elt_bytes(t.elt);
    c.row_bytes = (int64_t)t.dim * (int64_t)elt_bytes(t.elt);
    c.capacity = (int32_t)d->capacity;
    return true;
}

// ------------------------------------------------------------------------------------ peer-memory barrier
// One warp: lane r tells rank r "rank `me` has reached `epoch`" with a release store into r's flag array (peer
// memory over NVLink), then waits until rank r has told me the same.  The kernels before this one on the stream
// have completed, so their peer stores are ordered before the flag; the kernels after it see every peer's data.
struct PeerFlags { uint32_t* flags[16]; };
__global__ void peer_barrier_kernel(const __grid_constant__ PeerFlags F, int me, int nranks, uint32_t epoch) {
    pdl_begin();
    const int r = threadIdx.x;
    if (r < nranks) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(F.flags[r] + me), "r"(epoch) : "memory");
        uint32_t v;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(F.flags[me] + r) : "memory");
        } while ((int32_t)(v - epoch) < 0);
    }
}

}  // namespace etb

using namespace etb;

extern "C" {

int32_t etb_sgd_update(const etb_index_view* view_host, const etb_update_item* items_host, int32_t n_items, double eta,
                       int32_t flags, void* stream) {
    ETB_API_RANGE();
    launch_counter() = 0;
    return update_impl(view_host, items_host, n_items, eta, flags, (cudaStream_t)stream);
}

int32_t etb_adagrad_update(const etb_index_view* view_host, const etb_update_item* items_host, void* const* states_host,
                           int32_t n_items, double eta, double eps, int32_t flags, void* stream) {
    ETB_API_RANGE();
    launch_counter() = 0;
    ETB_REQUIRE(eps >= 0.0, "etb_adagrad_update: negative eps");
    return update_impl(view_host, items_host, n_items, eta, flags & ~ETB_UPDATE_FMA, (cudaStream_t)stream, kOptAdagrad,
                       states_host, eps);
}

int32_t etb_index_and_update(void* workspace, size_t workspace_bytes, const etb_update_item* items_host,
                             int32_t n_items, double eta, int32_t flags, void* stream) {
    ETB_API_RANGE();
    launch_counter() = 0;
    etb_index_view view;
    if (int32_t st = index_impl(workspace, workspace_bytes, items_host, n_items, &view, (cudaStream_t)stream)) return st;
    return update_impl(&view, items_host, n_items, eta, flags, (cudaStream_t)stream);
}

int32_t etb_uncompress(void* dst, int64_t ld_dst, int32_t dim, int32_t elt, const void* delta, int64_t ld_delta,
                       const void* idx, int32_t idx_elt, int64_t bag, int64_t batch, int64_t ld_idx, void* stream) {
    ETB_API_RANGE();
    launch_counter() = 0;
    ETB_REQUIRE(elt == ETB_F32 || elt == ETB_F64, "etb_uncompress: only Float32/Float64");
    ETB_REQUIRE(idx_elt_valid(idx_elt), "etb_uncompress: bad index type");
    ETB_REQUIRE(dim > 0 && batch >= 0 && bag >= 0, "etb_uncompress: bad sizes");
    if (batch == 0) return ETB_OK;
    ETB_REQUIRE(dst && delta && idx, "etb_uncompress: null pointer");
    const int threads = 128, blocks = (dim + threads - 1) / threads;
    cudaStream_t s = (cudaStream_t)stream;
    if (elt == ETB_F32) {
        if (idx_elt == ETB_I64) launch_k(uncompress_kernel<float, long long>, blocks, threads, 0, s, (float*)dst, ld_dst, dim, (const float*)delta, ld_delta, (const long long*)idx, bag, batch, ld_idx);
        else launch_k(uncompress_kernel<float, int>, blocks, threads, 0, s, (float*)dst, ld_dst, dim, (const float*)delta, ld_delta, (const int*)idx, bag, batch, ld_idx);
    } else {
        if (idx_elt == ETB_I64) launch_k(uncompress_kernel<double, long long>, blocks, threads, 0, s, (double*)dst, ld_dst, dim, (const double*)delta, ld_delta, (const long long*)idx, bag, batch, ld_idx);
        else launch_k(uncompress_kernel<double, int>, blocks, threads, 0, s, (double*)dst, ld_dst, dim, (const double*)delta, ld_delta, (const int*)idx, bag, batch, ld_idx);
    }
    ETB_LAUNCHED();
    return ETB_OK;
}

int32_t etb_a2a_unpack(void* dst, int64_t ld_dst, const void* recv, const int64_t* rows_host, const int64_t* row_off_host,
                       int32_t nranks, int64_t batch_local, int32_t elt, void* stream) {
    ETB_API_RANGE();
    return a2a_copy(true, dst, ld_dst, const_cast<void*>(recv), nullptr, rows_host, row_off_host, nranks, batch_local,
                    elt, (cudaStream_t)stream);
}

int32_t etb_a2a_pack(void* send, const void* src, int64_t ld_src, const int64_t* rows_host, const int64_t* row_off_host,
                     int32_t nranks, int64_t batch_local, int32_t elt, void* stream) {
    ETB_API_RANGE();
    return a2a_copy(false, const_cast<void*>(src), ld_src, send, nullptr, rows_host, row_off_host, nranks, batch_local,
                    elt, (cudaStream_t)stream);
}

int32_t etb_a2a_scatter(void* const* dst_ptrs_host, const void* src, int64_t ld_src, const int64_t* rows_host,
                        const int64_t* row_off_host, int32_t nranks, int64_t batch_local, int32_t elt, void* stream) {
    ETB_API_RANGE();
    return a2a_copy(false, const_cast<void*>(src), ld_src, nullptr, dst_ptrs_host, rows_host, row_off_host, nranks,
                    batch_local, elt, (cudaStream_t)stream);
}

int32_t etb_cache_admit(const etb_index_view* view_host, const etb_update_item* items_host, int32_t n_items,
                        int32_t min_count, void* stream) {
    ETB_API_RANGE();
    launch_counter() = 0;
    ETB_REQUIRE(view_host && (n_items == 0 || items_host) && n_items >= 0, "etb_cache_admit: bad arguments");
    if (view_host->n_total == 0) return ETB_OK;
    static thread_local CacheParams P;
    for (int i0 = 0; i0 < n_items; i0 += kUMaxItems) {  // bucket keys carry the table slot: one launch per 96 tables
        const int n = std::min(kUMaxItems, n_items - i0);
        bool any = false;
        // the kernel decodes slot = key >> row_bits over the whole call; give it this group's tables at their slots
        for (int j = 0; j < n; ++j) {
            if (int32_t st = validate_table(items_host[i0 + j].table, "etb_cache_admit")) return st;
            any = make_cache_item(items_host[i0 + j].table, P.item[j]) || any;
        }
        if (!any) continue;
        ETB_REQUIRE(i0 == 0, "etb_cache_admit: host-tier tables must be among the first %d tables of an ensemble", kUMaxItems);
        P.recs = (const BucketRec*)view_host->records;
        P.nnz = view_host->nnz;
        P.n_total = view_host->n_total;
        P.row_bits = view_host->row_bits;
        P.min_count = std::max(1, min_count);
        P.n_items = n;
        for (int j = 0; j < n; ++j)
            if (P.item[j].hist) ETB_CUDA(cudaMemsetAsync(P.item[j].hist, 0, ETB_CACHE_HIST_BINS * sizeof(int32_t), (cudaStream_t)stream));
        const int grid = (int)std::min<int64_t>((view_host->n_total + 7) / 8, (int64_t)num_sms() * 8);
        launch_k(cache_count_kernel, grid, 256, 0, (cudaStream_t)stream, P);
        ETB_LAUNCHED();
        launch_k(cache_admit_kernel, grid, 256, 0, (cudaStream_t)stream, P);
        ETB_LAUNCHED();
    }
    return ETB_OK;
}

int32_t etb_cache_flush(const etb_table* table_host, void* stream) {
    ETB_API_RANGE();
    launch_counter() = 0;
    ETB_REQUIRE(table_host, "etb_cache_flush: null table");
    if (int32_t st = validate_table(*table_host, "etb_cache_flush")) return st;
    CacheItem c;
    ETB_REQUIRE(make_cache_item(*table_host, c), "etb_cache_flush: not an ETB_TABLE_CACHED table");
    if (c.capacity == 0) return ETB_OK;
    const int grid = std::min((c.capacity + 7) / 8, num_sms() * 8);
    launch_k(cache_flush_kernel, grid, 256, 0, (cudaStream_t)stream, c);
    ETB_LAUNCHED();
    return ETB_OK;
}

int32_t etb_a2a_scatter_ld(void* const* dst_ptrs_host, const int64_t* dst_ld_host, const void* src, int64_t ld_src,
                           const int64_t* rows_host, const int64_t* row_off_host, int32_t nranks, int64_t batch_local,
                           int32_t elt, void* stream) {
    ETB_API_RANGE();
    ETB_REQUIRE(dst_ld_host, "etb_a2a_scatter_ld: null leading dimensions");
    return a2a_copy(false, const_cast<void*>(src), ld_src, nullptr, dst_ptrs_host, rows_host, row_off_host, nranks,
                    batch_local, elt, (cudaStream_t)stream, dst_ld_host);
}

int32_t etb_peer_barrier(void* const* flag_ptrs_host, int32_t rank, int32_t nranks, uint32_t epoch, void* stream) {
    ETB_API_RANGE();
    launch_counter() = 0;
    ETB_REQUIRE(nranks >= 1 && nranks <= 16 && rank >= 0 && rank < nranks, "etb_peer_barrier: bad rank %d of %d", rank, nranks);
    ETB_REQUIRE(flag_ptrs_host, "etb_peer_barrier: null flag pointers");
    PeerFlags F;
    for (int r = 0; r < 16; ++r) F.flags[r] = r < nranks ? (uint32_t*)flag_ptrs_host[r] : nullptr;
    for (int r = 0; r < nranks; ++r) ETB_REQUIRE(F.flags[r], "etb_peer_barrier: null flag array of rank %d", r);
    launch_k(peer_barrier_kernel, 1, 32, 0, (cudaStream_t)stream, F, rank, nranks, epoch);
    ETB_LAUNCHED();
    return ETB_OK;
}

}  // extern "C"
