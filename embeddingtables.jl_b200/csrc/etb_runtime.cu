// etb_runtime.cu -- runtime group of the C ABI (include/embtab_b200.h): HBM storage management
// that replaces Julia `Array` storage for tables, plus error reporting.
#include <stdarg.h>

#include <stdlib.h>

#include "etb_common.cuh"

namespace etb {

char* error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}

int32_t fail(int32_t status, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(error_buffer(), 512, fmt, ap);
    va_end(ap);
    return status;
}

int32_t& launch_counter() {
    static thread_local int32_t n = 0;
    return n;
}

bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("ETB_PDL"); return e && e[0] == '1'; }();
    return on;
}

int num_sms() {
    static thread_local int dev_cached = -1, sms = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return sms;
    if (dev != dev_cached) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sms = v;
        dev_cached = dev;
    }
    return sms;
}

int32_t validate_table(const etb_table& t, const char* who) {
    ETB_REQUIRE(elt_valid(t.elt), "%s: unsupported table element type %d", who, t.elt);
    ETB_REQUIRE(t.dim > 0, "%s: table dim must be positive (got %d)", who, t.dim);
    ETB_REQUIRE(t.ld >= t.dim, "%s: table ld (%d) < dim (%d)", who, t.ld, t.dim);
    ETB_REQUIRE(t.nrows >= 0, "%s: negative nrows", who);
    if (t.chunks && t.shard_rows == ETB_TABLE_CACHED) {
        const etb_cache_desc* c = (const etb_cache_desc*)t.chunks;
        ETB_REQUIRE(t.base && c->rows && c->slot_of_row && c->row_of_slot && c->cursor && c->hist && c->capacity >= 0,
                    "%s: cached table needs base, rows, slot_of_row, row_of_slot, cursor and hist", who);
        ETB_REQUIRE(t.nrows <= 0x7fffffffll && c->capacity <= 0x7fffffffll, "%s: cached table: more than 2^31 rows / slots", who);
    } else if (t.chunks) {
        ETB_REQUIRE(t.shard_rows > 0 && t.shard_rows <= 0xffffffffll,
                    "%s: split table needs 0 < shard_rows < 2^32 (got %lld)", who, (long long)t.shard_rows);
    } else {
        ETB_REQUIRE(t.base != nullptr || t.nrows == 0, "%s: table has neither base nor chunks", who);
    }
    return ETB_OK;
}

}  // namespace etb

using namespace etb;

extern "C" {

int32_t etb_version(void) { return ETB_VERSION; }

const char* etb_last_error(void) { return error_buffer(); }

int32_t etb_last_launch_count(void) { return launch_counter(); }

int32_t etb_device_count(int32_t* count_host) {
    ETB_REQUIRE(count_host, "etb_device_count: null output");
    int n = 0;
    ETB_CUDA(cudaGetDeviceCount(&n));
    *count_host = n;
    return ETB_OK;
}

int32_t etb_init(int32_t device) {
    ETB_CUDA(cudaSetDevice(device));
    ETB_CUDA(cudaFree(0));
    cudaDeviceProp prop;
    ETB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(ETB_ERR_UNSUPPORTED, "etb_init: device %d is sm_%d%d; this library is built for sm_100a only",
                    device, prop.major, prop.minor);
    return ETB_OK;
}

int32_t etb_malloc(void** ptr_host, size_t bytes) {
    ETB_REQUIRE(ptr_host, "etb_malloc: null output");
    *ptr_host = nullptr;
    if (bytes == 0) return ETB_OK;
    ETB_CUDA(cudaMalloc(ptr_host, bytes));
    return ETB_OK;
}

int32_t etb_free(void* ptr) {
    if (ptr) ETB_CUDA(cudaFree(ptr));
    return ETB_OK;
}

int32_t etb_malloc_host(void** ptr_host, size_t bytes) {
    ETB_REQUIRE(ptr_host, "etb_malloc_host: null output");
    *ptr_host = nullptr;
    if (bytes == 0) return ETB_OK;
    ETB_CUDA(cudaMallocHost(ptr_host, bytes));
    return ETB_OK;
}

int32_t etb_free_host(void* ptr_host) {
    if (ptr_host) ETB_CUDA(cudaFreeHost(ptr_host));
    return ETB_OK;
}

int32_t etb_memcpy_h2d(void* dst, const void* src_host, size_t bytes, void* stream) {
    ETB_API_RANGE();
    if (bytes == 0) return ETB_OK;
    ETB_REQUIRE(dst && src_host, "etb_memcpy_h2d: null pointer");
    ETB_CUDA(cudaMemcpyAsync(dst, src_host, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return ETB_OK;
}

int32_t etb_memcpy_d2h(void* dst_host, const void* src, size_t bytes, void* stream) {
    ETB_API_RANGE();
    if (bytes == 0) return ETB_OK;
    ETB_REQUIRE(dst_host && src, "etb_memcpy_d2h: null pointer");
    ETB_CUDA(cudaMemcpyAsync(dst_host, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return ETB_OK;
}

int32_t etb_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream) {
    if (bytes == 0) return ETB_OK;
    ETB_REQUIRE(dst && src, "etb_memcpy_d2d: null pointer");
    ETB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return ETB_OK;
}

int32_t etb_memcpy2d_h2d(void* dst, size_t dst_pitch, const void* src_host, size_t src_pitch, size_t width_bytes,
                         size_t height, void* stream) {
    ETB_API_RANGE();
    if (width_bytes == 0 || height == 0) return ETB_OK;
    ETB_REQUIRE(dst && src_host, "etb_memcpy2d_h2d: null pointer");
    ETB_REQUIRE(dst_pitch >= width_bytes && src_pitch >= width_bytes, "etb_memcpy2d_h2d: pitch smaller than the run");
    ETB_CUDA(cudaMemcpy2DAsync(dst, dst_pitch, src_host, src_pitch, width_bytes, height, cudaMemcpyHostToDevice,
                               (cudaStream_t)stream));
    return ETB_OK;
}

int32_t etb_memcpy2d_d2h(void* dst_host, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes,
                         size_t height, void* stream) {
    ETB_API_RANGE();
    if (width_bytes == 0 || height == 0) return ETB_OK;
    ETB_REQUIRE(dst_host && src, "etb_memcpy2d_d2h: null pointer");
    ETB_REQUIRE(dst_pitch >= width_bytes && src_pitch >= width_bytes, "etb_memcpy2d_d2h: pitch smaller than the run");
    ETB_CUDA(cudaMemcpy2DAsync(dst_host, dst_pitch, src, src_pitch, width_bytes, height, cudaMemcpyDeviceToHost,
                               (cudaStream_t)stream));
    return ETB_OK;
}

int32_t etb_memcpy2d_d2d(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes,
                         size_t height, void* stream) {
    ETB_API_RANGE();
    if (width_bytes == 0 || height == 0) return ETB_OK;
    ETB_REQUIRE(dst && src, "etb_memcpy2d_d2d: null pointer");
    ETB_REQUIRE(dst_pitch >= width_bytes && src_pitch >= width_bytes, "etb_memcpy2d_d2d: pitch smaller than the run");
    ETB_CUDA(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width_bytes, height, cudaMemcpyDefault, (cudaStream_t)stream));
    return ETB_OK;
}

int32_t etb_memset(void* dst, int32_t byte, size_t bytes, void* stream) {
    if (bytes == 0) return ETB_OK;
    ETB_REQUIRE(dst, "etb_memset: null pointer");
    ETB_CUDA(cudaMemsetAsync(dst, byte, bytes, (cudaStream_t)stream));
    return ETB_OK;
}

int32_t etb_ipc_export(void* ptr, void* handle_host) {
    static_assert(sizeof(cudaIpcMemHandle_t) == ETB_IPC_HANDLE_BYTES, "IPC handle size");
    ETB_REQUIRE(ptr && handle_host, "etb_ipc_export: null pointer");
    cudaIpcMemHandle_t h;
    ETB_CUDA(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle_host, &h, sizeof(h));
    return ETB_OK;
}

int32_t etb_ipc_import(const void* handle_host, void** ptr_host) {
    ETB_REQUIRE(handle_host && ptr_host, "etb_ipc_import: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle_host, sizeof(h));
    ETB_CUDA(cudaIpcOpenMemHandle(ptr_host, h, cudaIpcMemLazyEnablePeerAccess));
    return ETB_OK;
}

int32_t etb_ipc_close(void* ptr) {
    if (ptr) ETB_CUDA(cudaIpcCloseMemHandle(ptr));
    return ETB_OK;
}

int32_t etb_stream_create(void** stream_host) {
    ETB_REQUIRE(stream_host, "etb_stream_create: null output");
    cudaStream_t s;
    ETB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream_host = (void*)s;
    return ETB_OK;
}

int32_t etb_stream_sync(void* stream) {
    ETB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return ETB_OK;
}

int32_t etb_stream_destroy(void* stream) {
    if (stream) ETB_CUDA(cudaStreamDestroy((cudaStream_t)stream));
    return ETB_OK;
}

}  // extern "C"
