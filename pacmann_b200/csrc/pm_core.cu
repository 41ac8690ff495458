// Handle management, error reporting and device initialisation for libpacmann_cuda.so.
#include <cstdlib>
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

// ---- launch-time tuning knobs ---------------------------------------------------------------------------------
// Every launch heuristic that the parity tests must be able to force (the benchmarked variants included) is a knob
// here: read at every launch, initialised once from the environment (PM_HG_SYNC, ...), changed with pm_tuning_set().
struct TuneKnob { const char *name, *env; int dflt; std::atomic<int> value; std::atomic<bool> init; };
static TuneKnob g_knobs[T_COUNT] = {
    {"hg_sync", "PM_HG_SYNC", -1, {0}, {false}},       // round barrier of the hint kernel: -1 auto, 0 off, 1 on
    {"hg_warps", "PM_HG_WARPS", 0, {0}, {false}},      // CTA width of the hint kernel in warps: 0 auto, 1..16
    {"hg_ntab", "PM_HG_NTAB", 0, {0}, {false}},        // AES T-tables in shared memory: 0 auto, 1, 4
    {"hg_tail_split", "PM_HG_TAIL_SPLIT", -1, {0}, {false}},   // last partial round shared by all CTAs (stream-K): -1 auto (on), 0 off, 1 on
    {"hg_serpentine", "PM_HG_SERPENTINE", 1, {0}, {false}},   // alternate the sweep direction between rounds: 0 off, 1 on
    {"hg_xbytes", "PM_HG_XBYTES", 0, {0}, {false}},    // varying chunk-id bytes assumed by the hoisted PRF rounds: 0 auto, 1, 2, 4
    {"ans_split", "PM_ANS_SPLIT", 0, {0}, {false}},    // CTAs per sub-query of the answer kernel: 0 auto, 1..8
    {"hg_d2h_groups", "PM_HG_D2H_GROUPS", 0, {0}, {false}},   // launch groups of the host-buffer pm_hintgen: 0 auto, 1..16
    {"search_ans_stream", "PM_SEARCH_ANS_STREAM", 1, {0}, {false}},   // device search: answer kernel on a low-priority stream of its own (0 = one stream)
};
int tune(Tune t) {
    TuneKnob &k = g_knobs[t];
    if (!k.init.load(std::memory_order_acquire)) {
        const char *v = getenv(k.env);
        k.value.store(v && *v ? atoi(v) : k.dflt, std::memory_order_relaxed);
        k.init.store(true, std::memory_order_release);
    }
    return k.value.load(std::memory_order_relaxed);
}

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

PM_EXPORT int pm_tuning_set(const char *name, int value) {
    if (!name) return set_error(PM_ERR_ARG, "pm_tuning_set: null name");
    for (int t = 0; t < T_COUNT; t++)
        if (!strcmp(name, g_knobs[t].name)) {
            g_knobs[t].value.store(value, std::memory_order_relaxed);
            g_knobs[t].init.store(true, std::memory_order_release);
            return PM_OK;
        }
    return set_error(PM_ERR_ARG, "pm_tuning_set: unknown knob '%s'", name);
}
PM_EXPORT int pm_tuning_get(const char *name, int *value) {
    if (!name || !value) return set_error(PM_ERR_ARG, "pm_tuning_get: null pointer");
    for (int t = 0; t < T_COUNT; t++)
        if (!strcmp(name, g_knobs[t].name)) {
            *value = tune((Tune)t);
            return PM_OK;
        }
    return set_error(PM_ERR_ARG, "pm_tuning_get: unknown knob '%s'", name);
}

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

PM_EXPORT int pm_buf_upload(void *dev_ptr, const void *host, uint64_t bytes, int device) {
    if (bytes && (!dev_ptr || !host)) return set_error(PM_ERR_ARG, "pm_buf_upload: null pointer");
    int rc = ensure_device(device);
    if (rc) return rc;
    if (bytes) PM_CUDA(cudaMemcpy(dev_ptr, host, bytes, cudaMemcpyHostToDevice));
    return PM_OK;
}
PM_EXPORT int pm_buf_download(void *host, const void *dev_ptr, uint64_t bytes, int device) {
    if (bytes && (!dev_ptr || !host)) return set_error(PM_ERR_ARG, "pm_buf_download: null pointer");
    int rc = ensure_device(device);
    if (rc) return rc;
    if (bytes) PM_CUDA(cudaMemcpy(host, dev_ptr, bytes, cudaMemcpyDeviceToHost));
    return PM_OK;
}
// asynchronous device-to-device copy (unified addressing: either side may be a peer GPU's IPC-mapped buffer), on `stream`
PM_EXPORT int pm_buf_copy_dev(void *dst, const void *src, uint64_t bytes, int device, void *stream) {
    if (bytes && (!dst || !src)) return set_error(PM_ERR_ARG, "pm_buf_copy_dev: null pointer");
    int rc = ensure_device(device);
    if (rc) return rc;
    if (bytes) PM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return PM_OK;
}
PM_EXPORT int pm_buf_zero(void *dev_ptr, uint64_t bytes, int device) {
    if (bytes && !dev_ptr) return set_error(PM_ERR_ARG, "pm_buf_zero: null pointer");
    int rc = ensure_device(device);
    if (rc) return rc;
    if (bytes) PM_CUDA(cudaMemset(dev_ptr, 0, bytes));
    return PM_OK;
}
PM_EXPORT int pm_host_register(void *host, uint64_t bytes) {
    if (!host || !bytes) return set_error(PM_ERR_ARG, "pm_host_register: null pointer");
    int rc = ensure_device(-1);
    if (rc) return rc;
    PM_CUDA(cudaHostRegister(host, bytes, cudaHostRegisterPortable));
    return PM_OK;
}
PM_EXPORT int pm_host_unregister(void *host) {
    if (!host) return PM_OK;
    PM_CUDA(cudaHostUnregister(host));
    return PM_OK;
}

// ---- completion flags in (peer-mapped) device memory ---------------------------------------------------------------
// The exchange step of multi-GPU hint generation: every rank's kernel stores its parities straight into the consumer's
// table over NVLink; what is left of a "gather" is telling the consumer that the stores have landed.  A rank signals by
// adding 1 to a counter in the consumer's memory (system-scope release, stream-ordered behind its hint kernel); the
// consumer's stream waits until the counter reaches the expected epoch value.  No NCCL call, no host round trip.
namespace pm {
__global__ void flag_signal_kernel(unsigned int *flag) {
    __threadfence_system();
    asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(flag) : "memory");
}
// thread i waits for flags[32*i] (one 128-byte line per signalling rank: a fast rank running ahead cannot stand in for a
// slow one); a timeout is reported in flags[32*n] and the stream goes on -- never hang the GPU
__global__ void flag_wait_kernel(unsigned int *flags, unsigned int target, unsigned long long timeout_ns) {
    unsigned int *flag = flags + 32 * threadIdx.x, *timed_out = flags + 32 * blockDim.x;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        unsigned int seen;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
        if ((int)(seen - target) >= 0) return;
        __nanosleep(200);
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > timeout_ns) {
            *timed_out = 1;
            return;
        }
    }
}
}  // namespace pm
PM_EXPORT int pm_flag_signal_dev(void *flag, int device, void *stream) {
    if (!flag) return set_error(PM_ERR_ARG, "pm_flag_signal_dev: null pointer");
    int rc = ensure_device(device);
    if (rc) return rc;
    flag_signal_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned int *)flag);
    PM_CHECK_LAUNCH();
    count_launch();
    return PM_OK;
}
PM_EXPORT int pm_flag_wait_dev(void *flags, uint32_t n_flags, uint32_t target, uint32_t timeout_ms, int device, void *stream) {
    if (!flags || n_flags == 0 || n_flags > 1024) return set_error(PM_ERR_ARG, "pm_flag_wait_dev: bad argument");
    int rc = ensure_device(device);
    if (rc) return rc;
    if (timeout_ms == 0) {
        // Unbounded form: a stream memory operation per counter (cuStreamWaitValue32, wrap-safe >=) instead of a polling
        // kernel.  It occupies no SM -- a resident polling CTA keeps a cooperative launch of one-CTA-per-SM kernels (the hint
        // kernel takes a whole register file) from starting until the wait is over, which serialised the consumer's next
        // preprocessing behind the previous gather.
        typedef int (*wait32_t)(cudaStream_t, unsigned long long, unsigned int, unsigned int);
        static wait32_t wait32 = nullptr;
        if (!wait32) {
            void *fn = nullptr;
            cudaDriverEntryPointQueryResult qres;
            PM_CUDA(cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &qres));
            if (!fn || qres != cudaDriverEntryPointSuccess) return set_error(PM_ERR_UNSUPPORTED, "pm_flag_wait_dev: cuStreamWaitValue32 is not available");
            wait32 = (wait32_t)fn;
        }
        for (uint32_t i = 0; i < n_flags; i++) {
            const int e = wait32((cudaStream_t)stream, (unsigned long long)(uintptr_t)((unsigned int *)flags + 32 * i), target, 0x0 /* CU_STREAM_WAIT_VALUE_GEQ */);
            if (e != 0) return set_error(PM_ERR_CUDA, "pm_flag_wait_dev: cuStreamWaitValue32 failed (%d)", e);
        }
        return PM_OK;
    }
    flag_wait_kernel<<<1, n_flags, 0, (cudaStream_t)stream>>>((unsigned int *)flags, target, (unsigned long long)timeout_ms * 1000000ull);
    PM_CHECK_LAUNCH();
    count_launch();
    return PM_OK;
}
