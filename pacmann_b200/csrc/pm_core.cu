// Handle management, error reporting and device initialisation for libpacmann_cuda.so.
#include <cstring>
#include <new>

#include <algorithm>

#include "pm_common.cuh"

namespace pm {

int upload_tables();  // pm_pir.cu

static thread_local char t_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
    return code;
}
void count_launch(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

constexpr int MAX_DEV = 64;
static std::mutex g_dev_mu;
static bool g_dev_ready[MAX_DEV] = {};
static int g_sm_count[MAX_DEV] = {};
static void *g_zero[MAX_DEV] = {};

// device < 0: keep the calling thread's current device
int ensure_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return set_error(PM_ERR_CUDA, "no usable CUDA device (%s); libpacmann_cuda has no CPU fallback",
                         e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0) {
        PM_CUDA(cudaGetDevice(&device));
    } else {
        if (device >= n || device >= MAX_DEV) return set_error(PM_ERR_ARG, "device %d out of range (%d devices)", device, n);
        PM_CUDA(cudaSetDevice(device));
    }
    std::lock_guard<std::mutex> lock(g_dev_mu);
    if (!g_dev_ready[device]) {
        cudaDeviceProp prop;
        PM_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10)
            return set_error(PM_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                             prop.major, prop.minor);
        g_sm_count[device] = prop.multiProcessorCount;
        int rc = upload_tables();
        if (rc) return rc;
        PM_CUDA(cudaMalloc(&g_zero[device], 4096));
        PM_CUDA(cudaMemset(g_zero[device], 0, 4096));
        g_dev_ready[device] = true;
    }
    return PM_OK;
}
int sm_count(int device) { return g_sm_count[device]; }
unsigned int *sync_counter(pm_db *db) { return db->sync_pool + 32 * (db->sync_next.fetch_add(1) % 64); }

struct DevWork {
    std::mutex mu;
    cudaStream_t stream = nullptr;
    void *buf = nullptr;
    size_t bytes = 0;
};
static DevWork g_work[MAX_DEV];
int dev_work(int device, size_t bytes, void **ptr, cudaStream_t *stream, std::unique_lock<std::mutex> *lock) {
    if (device < 0) PM_CUDA(cudaGetDevice(&device));
    DevWork &w = g_work[device];
    *lock = std::unique_lock<std::mutex>(w.mu);
    if (!w.stream) PM_CUDA(cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking));
    if (w.bytes < bytes) {
        if (w.buf) {
            PM_CUDA(cudaStreamSynchronize(w.stream));
            PM_CUDA(cudaFree(w.buf));
            w.buf = nullptr;
            w.bytes = 0;
        }
        // grow-only, geometrically and from 8 MB: re-growing means cudaFree + cudaMalloc (milliseconds, device-wide
        // synchronisation), and callers such as the per-step distance calls of a search ask for slowly creeping sizes
        size_t want = std::max<size_t>(2 * bytes, (size_t)8 << 20);
        cudaError_t e = cudaMalloc(&w.buf, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return set_error(PM_ERR_NOMEM, "cudaMalloc(%zu) for the device workspace failed: %s", want, cudaGetErrorString(e));
        }
        w.bytes = want;
    }
    *ptr = w.buf;
    *stream = w.stream;
    return PM_OK;
}
const void *zero_page(int device) { return g_zero[device]; }

int scratch(pm_db *db, int slot, size_t bytes, void **out) {
    if (bytes == 0) bytes = 256;
    if (db->scratch_bytes[slot] < bytes) {
        if (db->scratch[slot]) {
            PM_CUDA(cudaStreamSynchronize(db->stream));
            PM_CUDA(cudaStreamSynchronize(db->copy_stream));
            PM_CUDA(cudaFree(db->scratch[slot]));
            db->scratch[slot] = nullptr;
            db->scratch_bytes[slot] = 0;
        }
        size_t want = bytes + bytes / 8;
        cudaError_t e = cudaMalloc(&db->scratch[slot], want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return set_error(PM_ERR_NOMEM, "cudaMalloc(%zu) for scratch failed: %s", want, cudaGetErrorString(e));
        }
        db->scratch_bytes[slot] = want;
    }
    *out = db->scratch[slot];
    return PM_OK;
}

}  // namespace pm

using namespace pm;

PM_EXPORT const char *pm_version(void) { return "pacmann-b200 0.1 (sm_100a)"; }
PM_EXPORT const char *pm_last_error(void) { return t_err; }
PM_EXPORT uint64_t pm_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

PM_EXPORT int pm_device_count(int *count) {
    if (!count) return set_error(PM_ERR_ARG, "pm_device_count: null pointer");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return set_error(PM_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *count = n;
    return PM_OK;
}

static int db_new(void *borrowed, uint64_t n_rows, uint64_t entry_u64, int device, pm_db **out);

PM_EXPORT int pm_db_create_empty(uint64_t n_rows, uint64_t entry_u64, int device, pm_db **out) {
    return db_new(nullptr, n_rows, entry_u64, device, out);
}
PM_EXPORT int pm_db_wrap(void *device_rows, uint64_t n_rows, uint64_t entry_u64, int device, pm_db **out) {
    if (!device_rows) return set_error(PM_ERR_ARG, "pm_db_wrap: null device pointer");
    if ((uintptr_t)device_rows % 16) return set_error(PM_ERR_ARG, "pm_db_wrap: device pointer must be 16-byte aligned");
    return db_new(device_rows, n_rows, entry_u64, device, out);
}

static int db_new(void *borrowed, uint64_t n_rows, uint64_t entry_u64, int device, pm_db **out) {
    if (!out) return set_error(PM_ERR_ARG, "pm_db_create: null out pointer");
    *out = nullptr;
    if (entry_u64 == 0) return set_error(PM_ERR_ARG, "pm_db_create: entry_u64 must be > 0");
    int rc = ensure_device(device);
    if (rc) return rc;
    if (device < 0) PM_CUDA(cudaGetDevice(&device));
    pm_db *db = new (std::nothrow) pm_db();
    if (!db) return set_error(PM_ERR_NOMEM, "out of host memory");
    db->device = device;
    db->n_rows = n_rows;
    db->entry_u64 = entry_u64;
    db->sm_count = sm_count(device);
    for (int i = 0; i < 4; i++) { db->scratch[i] = nullptr; db->scratch_bytes[i] = 0; }
    size_t bytes = n_rows * entry_u64 * 8;
    cudaError_t e = cudaSuccess;
    db->owns_rows = (borrowed == nullptr);
    if (borrowed) {
        db->d_rows = (uint64_t *)borrowed;
    } else {
        e = cudaMalloc(&db->d_rows, bytes ? bytes : 256);
        if (e != cudaSuccess) {
            cudaGetLastError();
            delete db;
            return set_error(PM_ERR_NOMEM, "cudaMalloc(%zu) for the table failed: %s", bytes, cudaGetErrorString(e));
        }
    }
    db->sync_next = 0;
    e = cudaMalloc(&db->sync_pool, 64 * 128);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&db->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&db->copy_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 4 && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&db->ev[i], cudaEventDisableTiming);
    if (e != cudaSuccess) {
        if (db->owns_rows) cudaFree(db->d_rows);
        delete db;
        return set_error(PM_ERR_CUDA, "stream/event creation failed: %s", cudaGetErrorString(e));
    }
    *out = db;
    return PM_OK;
}

PM_EXPORT int pm_db_upload(pm_db *db, uint64_t row0, uint64_t n_rows, const uint64_t *rows_host) {
    if (!db || (n_rows && !rows_host)) return set_error(PM_ERR_ARG, "pm_db_upload: null pointer");
    if (row0 + n_rows > db->n_rows) return set_error(PM_ERR_ARG, "pm_db_upload: rows [%llu,%llu) exceed the table",
                                                     (unsigned long long)row0, (unsigned long long)(row0 + n_rows));
    int rc = ensure_device(db->device);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(db->mu);
    PM_CUDA(cudaMemcpyAsync(db->d_rows + row0 * db->entry_u64, rows_host, n_rows * db->entry_u64 * 8, cudaMemcpyHostToDevice, db->stream));
    PM_CUDA(cudaStreamSynchronize(db->stream));
    return PM_OK;
}

PM_EXPORT int pm_db_create(const uint64_t *rows_host, uint64_t n_rows, uint64_t entry_u64, int device, pm_db **out) {
    if (n_rows && !rows_host) return set_error(PM_ERR_ARG, "pm_db_create: null rows");
    int rc = pm_db_create_empty(n_rows, entry_u64, device, out);
    if (rc) return rc;
    rc = pm_db_upload(*out, 0, n_rows, rows_host);
    if (rc) {
        pm_db_destroy(*out);
        *out = nullptr;
    }
    return rc;
}

PM_EXPORT int pm_db_info(const pm_db *db, uint64_t *n_rows, uint64_t *entry_u64, int *device, void **device_ptr) {
    if (!db) return set_error(PM_ERR_ARG, "pm_db_info: null handle");
    if (n_rows) *n_rows = db->n_rows;
    if (entry_u64) *entry_u64 = db->entry_u64;
    if (device) *device = db->device;
    if (device_ptr) *device_ptr = db->d_rows;
    return PM_OK;
}

PM_EXPORT int pm_db_sync(pm_db *db) {
    if (!db) return set_error(PM_ERR_ARG, "pm_db_sync: null handle");
    int rc = ensure_device(db->device);
    if (rc) return rc;
    PM_CUDA(cudaStreamSynchronize(db->stream));
    PM_CUDA(cudaStreamSynchronize(db->copy_stream));
    return PM_OK;
}

PM_EXPORT int pm_db_destroy(pm_db *db) {
    if (!db) return PM_OK;
    if (ensure_device(db->device) == PM_OK) {
        cudaStreamSynchronize(db->stream);
        cudaStreamSynchronize(db->copy_stream);
        for (int i = 0; i < 4; i++) {
            if (db->scratch[i]) cudaFree(db->scratch[i]);
            cudaEventDestroy(db->ev[i]);
        }
        if (db->owns_rows) cudaFree(db->d_rows);
        cudaFree(db->sync_pool);
        cudaStreamDestroy(db->stream);
        cudaStreamDestroy(db->copy_stream);
    }
    delete db;
    return PM_OK;
}

// ---- plain device buffers that can be shared between the per-GPU processes of one box (CUDA IPC) ----
// Used for the multi-GPU form of hint generation: rank 0 owns the full parity table, every other rank maps it and
// its hint kernel stores its shard straight into rank 0's HBM over NVLink -- no separate gather step.
PM_EXPORT int pm_buf_alloc(uint64_t bytes, int device, void **dev_ptr) {
    if (!dev_ptr) return set_error(PM_ERR_ARG, "pm_buf_alloc: null pointer");
    int rc = ensure_device(device);
    if (rc) return rc;
    cudaError_t e = cudaMalloc(dev_ptr, bytes ? bytes : 256);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_error(PM_ERR_NOMEM, "pm_buf_alloc: cudaMalloc(%llu) failed: %s", (unsigned long long)bytes, cudaGetErrorString(e));
    }
    return PM_OK;
}
PM_EXPORT int pm_buf_free(void *dev_ptr, int device) {
    if (!dev_ptr) return PM_OK;
    int rc = ensure_device(device);
    if (rc) return rc;
    PM_CUDA(cudaFree(dev_ptr));
    return PM_OK;
}
PM_EXPORT int pm_buf_ipc_export(void *dev_ptr, int device, uint8_t handle[64]) {
    if (!dev_ptr || !handle) return set_error(PM_ERR_ARG, "pm_buf_ipc_export: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
    int rc = ensure_device(device);
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    PM_CUDA(cudaIpcGetMemHandle(&h, dev_ptr));
    memcpy(handle, &h, 64);
    return PM_OK;
}
PM_EXPORT int pm_buf_ipc_open(const uint8_t handle[64], int device, void **dev_ptr) {
    if (!dev_ptr || !handle) return set_error(PM_ERR_ARG, "pm_buf_ipc_open: null pointer");
    int rc = ensure_device(device);
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    PM_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return PM_OK;
}
PM_EXPORT int pm_buf_ipc_close(void *dev_ptr, int device) {
    if (!dev_ptr) return PM_OK;
    int rc = ensure_device(device);
    if (rc) return rc;
    PM_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return PM_OK;
}
