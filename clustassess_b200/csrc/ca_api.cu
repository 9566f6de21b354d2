// ca_api.cu -- C ABI entry points, per-device contexts, device-resident label sets (one GPU or sharded over several)
// and the batched pipeline.  See include/clustassess_b200.h for the contract and the reference lines every entry replaces.
#include <dlfcn.h>
#include <nccl.h>  // types and enum values only: the library is loaded with dlopen at the first multi-GPU ca_init
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <memory>
#include <cmath>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>

#include "ca_internal.h"

namespace ca {

// ---- errors ---------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- host worker pool -------------------------------------------------------------------------------------
// The item planner runs between two GPU stages of every batched call; creating its threads (and their thread-local
// scratch) per call cost as much as the planning.  A pool's threads live until ca_shutdown / process exit.
namespace {
class WorkerPool {
  public:
    // runs job() on `n` pool threads and returns immediately; wait() blocks until all of them have returned
    void start(int n, std::function<void()> job) {
        std::unique_lock<std::mutex> g(mu_);
        while ((int)threads_.size() < n) threads_.emplace_back([this, id = (int)threads_.size()] { loop(id); });
        job_ = std::move(job);
        want_ = n;
        running_ = n;
        generation_++;
        cv_.notify_all();
    }
    void wait() {
        std::unique_lock<std::mutex> g(mu_);
        done_cv_.wait(g, [this] { return running_ == 0; });
    }
    void shutdown() {
        {
            std::unique_lock<std::mutex> g(mu_);
            stop_ = true;
            cv_.notify_all();
        }
        for (auto& t : threads_) t.join();
        threads_.clear();
        stop_ = false;
    }
    ~WorkerPool() { shutdown(); }

  private:
    void loop(int id) {
        unsigned long seen = 0;
        for (;;) {
            std::function<void()> job;
            {
                std::unique_lock<std::mutex> g(mu_);
                cv_.wait(g, [&] { return stop_ || (generation_ != seen && id < want_); });
                if (stop_) return;
                seen = generation_;
                job = job_;
            }
            job();
            std::unique_lock<std::mutex> g(mu_);
            if (--running_ == 0) done_cv_.notify_all();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    std::vector<std::thread> threads_;
    std::function<void()> job_;
    unsigned long generation_ = 0;
    int want_ = 0, running_ = 0;
    bool stop_ = false;
};
}  // namespace

// grow-only pinned host buffer: cudaMemcpyAsync from pageable memory first synchronises with the stream's earlier work,
// which would make the host wait for the previous batch's sort before it can hand over the next batch's plan
struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return CA_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        const size_t want = bytes + bytes / 8 + 4096;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e != cudaSuccess) {
            p = nullptr;
            set_error("cudaHostAlloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
            return CA_E_NOMEM;
        }
        cap = want;
        return CA_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T* as() { return static_cast<T*>(p); }
};

struct Workspace {
    PinnedBuf h_cstart, h_crow, h_items, h_small;
    DevBuf perm, cstart, crow, items, parts, w, wu_prefix, qinfo;
    void release() {
        perm.release(); cstart.release(); crow.release(); items.release(); parts.release(); w.release(); wu_prefix.release(); qinfo.release();
        h_cstart.release(); h_crow.release(); h_items.release(); h_small.release();
    }
};

// ---- host -> device upload: pinned staging slots (see upload_segments) --------------------------------------------
namespace {
struct Stager {
    static constexpr int MAXT = 8, SLOTS = 2;
    static constexpr size_t CHUNK = 4u << 20;
    char* pinned = nullptr;
    cudaEvent_t ev[MAXT][SLOTS] = {};
    bool used[MAXT][SLOTS] = {};
    int ensure() {
        if (pinned) return CA_OK;
        cudaError_t e = cudaHostAlloc((void**)&pinned, (size_t)MAXT * SLOTS * CHUNK, cudaHostAllocDefault);
        if (e != cudaSuccess) {
            pinned = nullptr;
            (void)cudaGetLastError();
            return CA_E_NOMEM;  // the caller falls back to the direct copy
        }
        for (int t = 0; t < MAXT; t++)
            for (int k = 0; k < SLOTS; k++) {
                if (cudaEventCreateWithFlags(&ev[t][k], cudaEventDisableTiming) != cudaSuccess) {
                    (void)cudaGetLastError();
                    release();
                    return CA_E_CUDA;
                }
                used[t][k] = false;
            }
        return CA_OK;
    }
    void release() {
        for (int t = 0; t < MAXT; t++)
            for (int k = 0; k < SLOTS; k++) {
                if (ev[t][k]) cudaEventDestroy(ev[t][k]);
                ev[t][k] = nullptr;
                used[t][k] = false;
            }
        if (pinned) cudaFreeHost(pinned);
        pinned = nullptr;
    }
    char* slot(int t, int k) { return pinned + ((size_t)t * SLOTS + k) * CHUNK; }
};
struct CacheBlock {
    void* p;
    size_t bytes;
};
constexpr size_t CACHE_LIMIT = 96ull << 30;  // parked bytes beyond this go back to the driver

// ---- per-device state ------------------------------------------------------------------------------------------
// Everything that lives on (or is pinned for) one GPU: stream and events, the caching allocator's parked blocks, the batched
// pipeline's workspace, the upload stager, the scratch buffers and a planner pool.  The library keeps one Dev per device
// selected by ca_init; a host thread works on the device its thread-local pointer names (the primary one by default).
struct Dev {
    Ctx c;
    int index = 0;  // position in the selected device list
    std::mutex cache_mu;
    std::vector<CacheBlock> cache;
    Workspace W;
    Stager S;
    DevBuf scr[SCR_COUNT];
    WorkerPool pool;
    ncclComm_t comm = nullptr;
    DevBuf red_ecc, red_sim;  // multi-GPU jobs: this device's partial sums
    cudaEvent_t ev_lab = nullptr;  // multi-GPU jobs: this device's transposition (with its stores into the peers' memory) has finished
};
std::vector<std::unique_ptr<Dev>>& devs() {
    static std::vector<std::unique_ptr<Dev>> d;
    return d;
}
std::mutex g_init_mu;
std::atomic<int> g_live_handles{0};
thread_local Dev* tl_dev = nullptr;
Dev& cur() {
    if (tl_dev) return *tl_dev;
    auto& d = devs();
    if (d.empty()) d.emplace_back(new Dev());  // an uninitialised placeholder: ensure_init() fills it in
    return *d[0];
}
struct DevScope {  // work on device g for the lifetime of the object
    Dev* prev;
    explicit DevScope(Dev* d) : prev(tl_dev) {
        tl_dev = d;
        if (d && d->c.inited) cudaSetDevice(d->c.device);
    }
    ~DevScope() {
        tl_dev = prev;
        if (prev && prev->c.inited) cudaSetDevice(prev->c.device);
    }
};
}  // namespace

Ctx& ctx() { return cur().c; }
DevBuf& scratch(int which) { return cur().scr[which]; }
static Workspace& ws() { return cur().W; }

// ---- device memory: a small caching allocator (per device) ------------------------------------------------------
// The host-pointer entry points (the ones the R shim binds) create and drop a label set per call; cudaMalloc /
// cudaFree of multi-gigabyte buffers cost more than the kernels of a small job and cudaFree synchronises the
// device.  Freed blocks are parked here and handed out again (best fit, at most 2x oversized); ca_shutdown()
// returns everything to the driver.
cudaError_t dev_alloc(void** out, size_t bytes) {
    Dev& D = cur();
    {
        std::lock_guard<std::mutex> g(D.cache_mu);
        auto& c = D.cache;
        int best = -1;
        for (int i = 0; i < (int)c.size(); i++)
            if (c[i].bytes >= bytes && c[i].bytes <= 2 * bytes + (1u << 20) && (best < 0 || c[i].bytes < c[best].bytes)) best = i;
        if (best >= 0) {
            *out = c[best].p;
            c.erase(c.begin() + best);
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(out, bytes);
    if (e != cudaSuccess) {  // out of memory: drop the cache and retry once
        (void)cudaGetLastError();
        dev_cache_release();
        e = cudaMalloc(out, bytes);
    }
    return e;
}
void dev_free(void* p, size_t bytes) {
    if (!p) return;
    Dev& D = cur();
    std::lock_guard<std::mutex> g(D.cache_mu);
    size_t parked = bytes;
    for (auto& b : D.cache) parked += b.bytes;
    if (parked > CACHE_LIMIT) {
        cudaFree(p);
        return;
    }
    D.cache.push_back({p, bytes});
}
void dev_cache_release() {
    Dev& D = cur();
    std::lock_guard<std::mutex> g(D.cache_mu);
    for (auto& b : D.cache) cudaFree(b.p);
    D.cache.clear();
}

int DevBuf::ensure(size_t bytes) {
    if (bytes <= cap) return CA_OK;
    if (p) dev_free(p, cap);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = dev_alloc(&p, want);
    if (e != cudaSuccess) {
        p = nullptr;
        set_error("cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
        return CA_E_NOMEM;
    }
    cap = want;
    return CA_OK;
}
void DevBuf::release() {
    if (p) dev_free(p, cap);
    p = nullptr;
    cap = 0;
}

// ---- NCCL, loaded at run time --------------------------------------------------------------------------------
// The reduction of the per-device partial sums is ONE ncclAllReduce (64-bit integer sums, see ca_internal.h) over
// NVLink; the communicators are created by ncclCommInitAll in ca_init.  libnccl is opened with dlopen so that a
// single-GPU user (and a process that already carries its own copy, e.g. PyTorch's) never depends on it at load time.
namespace {
struct Nccl {
    void* handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load() {
        if (handle) return true;
        const char* names[] = {getenv("CA_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            if (!nm || !*nm) continue;
            handle = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
            if (handle) break;
        }
        if (!handle) return false;
        CommInitAll = (decltype(CommInitAll))dlsym(handle, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(handle, "ncclCommDestroy");
        AllReduce = (decltype(AllReduce))dlsym(handle, "ncclAllReduce");
        GroupStart = (decltype(GroupStart))dlsym(handle, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(handle, "ncclGroupEnd");
        GetErrorString = (decltype(GetErrorString))dlsym(handle, "ncclGetErrorString");
        if (!CommInitAll || !CommDestroy || !AllReduce || !GroupStart || !GroupEnd || !GetErrorString) {
            dlclose(handle);
            handle = nullptr;
            return false;
        }
        return true;
    }
};
Nccl& nccl() {
    static Nccl n;
    return n;
}
bool g_have_comms = false;
}  // namespace

// ---- initialisation -----------------------------------------------------------------------------------------------
static int init_device(Dev& D, int device) {
    Ctx& c = D.c;
    CA_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    CA_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return CA_E_NO_DEVICE;
    }
    c.device = device;
    c.sm_count = prop.multiProcessorCount;
    c.smem_optin = prop.sharedMemPerBlockOptin;
    if (!c.own_stream) CA_CUDA(cudaStreamCreateWithFlags(&c.own_stream, cudaStreamNonBlocking));
    c.stream = c.own_stream;
    for (int i = 0; i < 10; i++)
        if (!c.ev[i]) CA_CUDA(cudaEventCreate(&c.ev[i]));
    if (!D.ev_lab) CA_CUDA(cudaEventCreateWithFlags(&D.ev_lab, cudaEventDisableTiming));
    if (!c.up_stream) CA_CUDA(cudaStreamCreateWithFlags(&c.up_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 4; i++) {
        if (!c.peer_stream[i]) CA_CUDA(cudaStreamCreateWithFlags(&c.peer_stream[i], cudaStreamNonBlocking));
        if (!c.peer_ev[i]) CA_CUDA(cudaEventCreateWithFlags(&c.peer_ev[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < 8; i++)
        if (!c.up_ev[i]) CA_CUDA(cudaEventCreateWithFlags(&c.up_ev[i], cudaEventDisableTiming));
    c.inited = true;
    return CA_OK;
}

static void release_device(Dev& D) {
    if (!D.c.inited) return;
    DevScope scope(&D);
    cudaDeviceSynchronize();
    D.W.release();
    D.S.release();
    for (auto& b : D.scr) b.release();
    D.red_ecc.release();
    D.red_sim.release();
    dev_cache_release();
    D.pool.shutdown();
    if (D.comm && nccl().CommDestroy) nccl().CommDestroy(D.comm);
    D.comm = nullptr;
    for (int i = 0; i < 10; i++)
        if (D.c.ev[i]) { cudaEventDestroy(D.c.ev[i]); D.c.ev[i] = nullptr; }
    if (D.ev_lab) { cudaEventDestroy(D.ev_lab); D.ev_lab = nullptr; }
    for (int i = 0; i < 8; i++)
        if (D.c.up_ev[i]) { cudaEventDestroy(D.c.up_ev[i]); D.c.up_ev[i] = nullptr; }
    if (D.c.up_stream) { cudaStreamDestroy(D.c.up_stream); D.c.up_stream = nullptr; }
    for (int i = 0; i < 4; i++) {
        if (D.c.peer_ev[i]) { cudaEventDestroy(D.c.peer_ev[i]); D.c.peer_ev[i] = nullptr; }
        if (D.c.peer_stream[i]) { cudaStreamDestroy(D.c.peer_stream[i]); D.c.peer_stream[i] = nullptr; }
    }
    if (D.c.own_stream) { cudaStreamDestroy(D.c.own_stream); D.c.own_stream = nullptr; }
    D.c.stream = nullptr;
    D.c.inited = false;
}

static void shutdown_all() {
    for (auto& d : devs()) release_device(*d);
    devs().clear();
    g_have_comms = false;
    tl_dev = nullptr;
}

// Select `ndev` devices.  The same list again is a no-op; a different list tears the old state down first (refused while
// label sets are alive: their memory belongs to the old devices).
static int init_devices(const int* list, int ndev) {
    std::lock_guard<std::mutex> g(g_init_mu);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        (void)cudaGetLastError();
        set_error("no CUDA device available (%s); this library has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return CA_E_NO_DEVICE;
    }
    if (ndev > CA_MAX_DEVICES) {
        set_error("ca_init: %d devices requested, at most %d are supported", ndev, CA_MAX_DEVICES);
        return CA_E_ARG;
    }
    for (int i = 0; i < ndev; i++) {
        if (list[i] < 0 || list[i] >= count) {
            set_error("ca_init: device %d out of range (have %d)", list[i], count);
            return CA_E_ARG;
        }
        for (int j = 0; j < i; j++)
            if (list[j] == list[i]) {
                set_error("ca_init: device %d listed twice", list[i]);
                return CA_E_ARG;
            }
    }
    auto& D = devs();
    bool same = (int)D.size() == ndev;
    for (int i = 0; same && i < ndev; i++) same = D[i]->c.inited && D[i]->c.device == list[i];
    if (same) return CA_OK;
    bool any_inited = false;
    for (auto& d : D) any_inited |= d->c.inited;
    if (any_inited) {
        if (g_live_handles.load() > 0) {
            set_error("ca_init: cannot switch devices while %d label set(s) are alive; destroy them (or call ca_shutdown) first", g_live_handles.load());
            return CA_E_ARG;
        }
        shutdown_all();
    }
    D.clear();
    for (int i = 0; i < ndev; i++) {
        D.emplace_back(new Dev());
        D[i]->index = i;
        int rc = init_device(*D[i], list[i]);
        if (rc != CA_OK) {
            shutdown_all();
            return rc;
        }
    }
    if (ndev > 1) {  // every device stores into every other's label array and reads nothing remote afterwards: peer access both ways
        for (int i = 0; i < ndev; i++) {
            cudaSetDevice(list[i]);
            for (int j = 0; j < ndev; j++) {
                if (i == j) continue;
                int can = 0;
                cudaDeviceCanAccessPeer(&can, list[i], list[j]);
                if (!can) {
                    set_error("ca_init: device %d cannot access device %d's memory (no NVLink / PCIe peer path); multi-GPU jobs need it", list[i], list[j]);
                    shutdown_all();
                    return CA_E_UNSUPPORTED;
                }
                cudaError_t pe = cudaDeviceEnablePeerAccess(list[j], 0);
                if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) {
                    set_error("cudaDeviceEnablePeerAccess(%d -> %d) failed: %s", list[i], list[j], cudaGetErrorString(pe));
                    shutdown_all();
                    return CA_E_CUDA;
                }
                (void)cudaGetLastError();
            }
        }
        // NCCL communicators (CA_REDUCE=p2p skips NCCL: the partial sums are then added by a kernel reading peer memory)
        const char* red = getenv("CA_REDUCE");
        if (!(red && strcmp(red, "p2p") == 0) && nccl().load()) {
            std::vector<ncclComm_t> comms(ndev);
            ncclResult_t r = nccl().CommInitAll(comms.data(), ndev, list);
            if (r == ncclSuccess) {
                for (int i = 0; i < ndev; i++) D[i]->comm = comms[i];
                g_have_comms = true;
            } else {
                fprintf(stderr, "clustassess_b200: ncclCommInitAll failed (%s); falling back to the peer-memory reduction\n", nccl().GetErrorString(r));
            }
        }
        cudaSetDevice(list[0]);
    }
    tl_dev = nullptr;
    return CA_OK;
}

int ensure_init() {
    auto& D = devs();
    if (!D.empty() && D[0]->c.inited) {
        CA_CUDA(cudaSetDevice(cur().c.device));
        return CA_OK;
    }
    const int zero = 0;
    return init_devices(&zero, 1);
}
static int ndevices() { return (int)devs().size(); }

// ---- label preparation ------------------------------------------------------------------------------
static constexpr int64_t RANGE_LIMIT = 1ll << 24;        // per-partition label range (max-min+1)
static constexpr int64_t TOTAL_RANGE_LIMIT = 1ll << 28;  // summed over partitions
static constexpr int32_t K_LIMIT = 65535;                // compact labels are uint16

bool engine_supports(int32_t kmax) {
    if (kmax > K_LIMIT) return false;
    int32_t max_rows = 0;
    tile_smem_budget(kmax, &max_rows);
    return max_rows >= 1;
}

// stage A: min / max, label histogram, dense ranks, cluster counts of the partitions held here
static int ranks_stage_a(Labels& L) {
    Ctx& c = ctx();
    cudaStream_t s = c.stream;
    const int32_t m = L.m;
    const int64_t n = L.n;
    const bool dbg = getenv("CA_PIPE_DEBUG") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (dbg) fprintf(stderr, "    [stage A] %-18s +%.3f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    };
    L.h_min.assign(m, 0);
    L.h_max.assign(m, 0);
    L.h_k.assign(m, 0);
    L.h_roff.assign(m + 1, 0);
    L.kmax = 0;
    L.identity = true;
    if (m == 0) return CA_OK;
    CA_TRY(L.d_min.ensure((size_t)m * 4));
    CA_TRY(L.d_max.ensure((size_t)m * 4));
    lap("buffers");
    CA_TRY(launch_minmax(L.d_raw, n, m, L.d_min.as<int32_t>(), L.d_max.as<int32_t>(), s));
    lap("minmax launched");
    CA_CUDA(cudaMemcpyAsync(L.h_min.data(), L.d_min.p, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
    CA_CUDA(cudaMemcpyAsync(L.h_max.data(), L.d_max.p, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
    CA_CUDA(cudaStreamSynchronize(s));
    lap("minmax synced");
    int64_t max_range = 0;
    for (int32_t p = 0; p < m; p++) {
        if (L.h_min[p] == INT32_MIN) {
            set_error("partition %d contains NA_integer_ (INT_MIN); NA labels are not supported", L.gid_of(p));
            return CA_E_NA_LABEL;
        }
        int64_t range = (int64_t)L.h_max[p] - (int64_t)L.h_min[p] + 1;
        if (range > RANGE_LIMIT) {
            set_error("partition %d: label range %lld exceeds the supported %lld", L.gid_of(p), (long long)range, (long long)RANGE_LIMIT);
            return CA_E_UNSUPPORTED;
        }
        L.h_roff[p + 1] = L.h_roff[p] + range;
        max_range = std::max(max_range, range);
    }
    if (L.h_roff[m] > TOTAL_RANGE_LIMIT) {
        set_error("summed label range %lld exceeds the supported %lld", (long long)L.h_roff[m], (long long)TOTAL_RANGE_LIMIT);
        return CA_E_UNSUPPORTED;
    }
    const size_t tot = (size_t)L.h_roff[m];
    CA_TRY(L.d_roff.ensure((size_t)(m + 1) * 8));
    CA_TRY(L.d_cnt.ensure(tot * 4));
    CA_TRY(L.d_rank.ensure(tot * 4));
    CA_TRY(L.d_kcount.ensure((size_t)m * 4));
    CA_CUDA(cudaMemcpyAsync(L.d_roff.p, L.h_roff.data(), (size_t)(m + 1) * 8, cudaMemcpyHostToDevice, s));
    CA_CUDA(cudaMemsetAsync(L.d_cnt.p, 0, tot * 4, s));
    CA_TRY(launch_count(L.d_raw, n, m, L.d_min.as<int32_t>(), L.d_roff.as<int64_t>(), nullptr, (int32_t)max_range,
                        L.d_cnt.as<uint32_t>(), s));
    CA_TRY(launch_rank(L.d_cnt.as<uint32_t>(), L.d_roff.as<int64_t>(), nullptr, nullptr, m, L.d_rank.as<uint32_t>(),
                       L.d_kcount.as<int32_t>(), s));
    CA_CUDA(cudaMemcpyAsync(L.h_k.data(), L.d_kcount.p, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
    lap("count/rank queued");
    CA_CUDA(cudaStreamSynchronize(s));
    lap("count/rank synced");
    for (int32_t p = 0; p < m; p++) {
        L.kmax = std::max(L.kmax, L.h_k[p]);
        if ((int64_t)L.h_k[p] != L.h_roff[p + 1] - L.h_roff[p]) L.identity = false;
    }
    return CA_OK;
}

// stage B: compact cluster sizes of the partitions held here, at the row stride of the whole set (kmax over all shards)
static int ranks_stage_b(Labels& L, int32_t kmax_all, int32_t kp_fixed = 0) {
    cudaStream_t s = ctx().stream;
    const int32_t m = L.m;
    L.kmax = kmax_all;
    L.kp = kp_fixed ? kp_fixed : std::max(8, (L.kmax + 1 + 7) & ~7);  // + the pad column
    L.mpad = (L.mg() + 31) & ~31;
    if (m == 0) return CA_OK;
    int64_t max_range = 0;
    for (int32_t p = 0; p < m; p++) max_range = std::max(max_range, L.h_roff[p + 1] - L.h_roff[p]);
    CA_TRY(L.d_sizes.ensure((size_t)m * L.kp * 4));
    CA_CUDA(cudaMemsetAsync(L.d_sizes.p, 0, (size_t)m * L.kp * 4, s));
    CA_TRY(launch_sizes(L.d_cnt.as<uint32_t>(), L.d_rank.as<uint32_t>(), L.d_roff.as<int64_t>(), nullptr, nullptr, m,
                        (int32_t)max_range, L.kp, L.d_sizes.as<int32_t>(), s));
    L.h_sizes.resize((size_t)m * L.kp);
    CA_CUDA(cudaMemcpyAsync(L.h_sizes.data(), L.d_sizes.p, (size_t)m * L.kp * 4, cudaMemcpyDeviceToHost, s));
    CA_CUDA(cudaStreamSynchronize(s));
    return CA_OK;
}

// stage C: what the wave kernel reads per PARTNER partition -- initial table entries, K | flags, table templates -- from the
// cluster sizes and counts of the whole set
static int ranks_stage_c(Labels& L) {
    cudaStream_t s = ctx().stream;
    const int32_t mg = L.mg();
    CA_TRY(L.d_spack.ensure((size_t)mg * L.kp * 4));
    CA_TRY(L.d_kinfo.ensure((size_t)L.mpad * 4));
    CA_CUDA(cudaMemsetAsync(L.d_kinfo.p, 0, (size_t)L.mpad * 4, s));
    CA_TRY(launch_spack(L.sizes_all(), L.kcount_all(), mg, L.kp, L.d_spack.as<uint32_t>(), L.d_kinfo.as<int32_t>(), s));
    int32_t max_rows = 0, half = 0;
    tile_smem_budget(L.kmax, &max_rows, &half);
    L.tmpl_bytes = std::max(half, 16);  // one table buffer of the wave kernel per partition
    CA_TRY(L.d_tmpl.ensure((size_t)mg * (size_t)L.tmpl_bytes));
    CA_TRY(launch_table_template(L.d_spack.as<uint32_t>(), L.kcount_all(), mg, L.kp, L.tmpl_bytes, std::max(max_rows, 1),
                                 L.d_tmpl.as<uint32_t>(), s));
    return CA_OK;
}

int prepare_ranks(Labels& L) {
    if (L.ranked) return CA_OK;
    CA_TRY(ranks_stage_a(L));
    CA_TRY(ranks_stage_b(L, L.kmax));
    L.ranked = true;
    return CA_OK;
}

int prepare_labels(Labels& L) {
    if (L.prepared) return CA_OK;
    Ctx& c = ctx();
    cudaStream_t s = c.stream;
    const int32_t m = L.m;
    const int64_t n = L.n;
    CA_TRY(prepare_ranks(L));
    if (!engine_supports(L.kmax)) {
        set_error("a partition has %d clusters: beyond the batched engine's tables (uint16 labels, one table row per shared-memory buffer)", L.kmax);
        return CA_E_UNSUPPORTED;
    }
    CA_TRY(ranks_stage_c(L));
    CA_CUDA(cudaEventRecord(c.ev[1], s));
    CA_TRY(L.d_lab16.ensure((size_t)(n + 1) * L.mpad * 2));  // n cells + the pad cell
    LabDst dst;
    dst.n = 1;
    dst.p[0] = L.d_lab16.as<uint16_t>();
    CA_TRY(launch_relabel_transpose(L.d_raw, n, m, L.mpad / 32, nullptr, L.d_min.as<int32_t>(), L.identity ? nullptr : L.d_rank.as<uint32_t>(),
                                    L.d_roff.as<int64_t>(), L.d_kcount.as<int32_t>(), dst, s));
    CA_CUDA(cudaEventRecord(c.ev[2], s));  // no synchronisation: the host goes on to plan while the transposition runs
    L.prepared = true;
    return CA_OK;
}

// zigzag owner of partition i among `world` ranks: 0..w-1, w-1..0, ... (balances sum of M-1-i)
static inline int owner_of(int32_t i, int32_t world) {
    int32_t r = i % (2 * world);
    return r < world ? r : 2 * world - 1 - r;
}

// ---- host -> device upload ----------------------------------------------------------------------------
// The R shim hands over R's own vectors, i.e. PAGEABLE memory.  cudaMemcpyAsync from pageable memory is staged by
// the driver through one small pinned buffer on one thread and reaches a fraction of the PCIe rate; at config 5
// (2 GB of labels) that copy would cost more than the whole computation.  Large pageable uploads are therefore
// staged here: up to `max_threads` host threads copy 4 MB chunks into their own pair of pinned slots and enqueue the
// slot -> device DMA on the caller's stream, so the host memcpys of all threads overlap each other and the DMA.
// Pinned / registered sources (bench.py's e2e buffers, torch pinned tensors) and small uploads go straight to
// cudaMemcpyAsync.  CA_UPLOAD_DIRECT=1 forces the direct path (for A/B measurements).
struct UploadSeg {
    const void* src;
    void* dst;
    size_t bytes;
};
namespace {
bool is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}
}  // namespace

// paced = true (the pipelined path): at most PACE_BATCH chunks are handed to the copy engine before it is allowed to run empty,
// so that small copies other streams issue meanwhile are served between two batches (see upload_throttled).
static int upload_segments(const UploadSeg* segs, size_t nseg, cudaStream_t s, int max_threads = Stager::MAXT, bool paced = false) {
    size_t total = 0;
    for (size_t i = 0; i < nseg; i++) total += segs[i].bytes;
    if (total == 0) return CA_OK;
    Dev& D = cur();
    bool staged = total >= 2 * Stager::CHUNK && getenv("CA_UPLOAD_DIRECT") == nullptr;
    for (size_t i = 0; staged && i < nseg; i++)
        if (segs[i].bytes && !is_pageable(segs[i].src)) staged = false;
    if (staged && D.S.ensure() != CA_OK) staged = false;
    if (!staged) {
        for (size_t i = 0; i < nseg; i++)
            if (segs[i].bytes) CA_CUDA(cudaMemcpyAsync(segs[i].dst, segs[i].src, segs[i].bytes, cudaMemcpyHostToDevice, s));
        return CA_OK;
    }
    Stager& S = D.S;
    std::vector<UploadSeg> chunks;
    chunks.reserve(total / Stager::CHUNK + nseg);
    for (size_t i = 0; i < nseg; i++)
        for (size_t off = 0; off < segs[i].bytes; off += Stager::CHUNK)
            chunks.push_back({(const char*)segs[i].src + off, (char*)segs[i].dst + off, std::min(Stager::CHUNK, segs[i].bytes - off)});
    const int nthreads = (int)std::max<size_t>(1, std::min<size_t>({(size_t)Stager::MAXT, (size_t)std::max(1, max_threads),
                                                                      (size_t)std::thread::hardware_concurrency() / 2, chunks.size()}));
    const int device = D.c.device;
    std::atomic<size_t> next{0};
    std::atomic<int> failed{(int)cudaSuccess};
    static const int PACE_BATCH = getenv("CA_PIPE_BATCH") ? std::max(1, atoi(getenv("CA_PIPE_BATCH"))) : 4;
    std::mutex pace_mu;  // paced uploads: one thread at a time talks to the copy engine
    cudaEvent_t pace_last = nullptr;
    int pace_count = 0;
    auto work = [&](int t) {
        if (cudaError_t e0 = cudaSetDevice(device)) { failed = (int)e0; return; }
        int k = 0;
        for (size_t c; (c = next.fetch_add(1)) < chunks.size() && failed == (int)cudaSuccess; k ^= 1) {
            cudaError_t e = cudaSuccess;
            if (S.used[t][k]) {  // the slot's previous DMA has read it
                if (paced)
                    while ((e = cudaEventQuery(S.ev[t][k])) == cudaErrorNotReady) std::this_thread::sleep_for(std::chrono::microseconds(20));
                else
                    e = cudaEventSynchronize(S.ev[t][k]);
            }
            if (e == cudaSuccess) {
                memcpy(S.slot(t, k), chunks[c].src, chunks[c].bytes);
                std::unique_lock<std::mutex> lk(pace_mu, std::defer_lock);
                if (paced) {
                    lk.lock();
                    if (pace_count >= PACE_BATCH) {  // let the engine drain
                        while ((e = cudaEventQuery(pace_last)) == cudaErrorNotReady) std::this_thread::sleep_for(std::chrono::microseconds(20));
                        pace_count = 0;
                    }
                }
                if (e == cudaSuccess) e = cudaMemcpyAsync(chunks[c].dst, S.slot(t, k), chunks[c].bytes, cudaMemcpyHostToDevice, s);
                if (e == cudaSuccess) e = cudaEventRecord(S.ev[t][k], s);
                pace_last = S.ev[t][k];
                pace_count++;
            }
            if (e != cudaSuccess) { failed = (int)e; return; }
            S.used[t][k] = true;
        }
    };
    std::vector<std::thread> helpers;
    helpers.reserve(nthreads);
    try {
        for (int t = 1; t < nthreads; t++) helpers.emplace_back(work, t);
    } catch (...) {  // no more threads to be had: the ones that started (and this one) take all chunks
    }
    work(0);
    for (auto& th : helpers) th.join();
    if (failed != (int)cudaSuccess) {
        (void)cudaGetLastError();
        set_error("host-to-device copy failed: %s", cudaGetErrorString((cudaError_t)failed.load()));
        return CA_E_CUDA;
    }
    return CA_OK;
}

// ---- which way through the wave kernel (ca_tile.cu) a job takes: pure host logic, also exported for the CPU tests (ca_plan_route) ----
// rt       ratio tables (convert_tables): one division per table ENTRY and a gather that only adds.  Measured (tools/sim_scan_sweep.py,
//          profiles/r2_table_routes.txt): the extra pass and barrier per phase cost what the shorter gather saves -- except where
//          clusters are HUGE (64 x 1M cells with K = 8 .. 64: 14.1 -> 9.5 ms): there most partners have clusters beyond the 16-bit
//          size field and most owner clusters need 32-bit counters, and the plain gather looks every size up in global memory.
//          So: at most 96 clusters of more than 8192 cells on average.  Needs ecc (the matrix alone has no gather).  CA_RT=0/1 forces it.
// sim_scan the ECS matrix summed over the table entries (scan_sim) instead of per (cell, partner).  Reading the tables costs about
//          (ceil(K / 32) * 42 + 30) warp instructions per (partner, row), i.e. per n / K cells; the per-cell way about 8 per 32
//          (cell, partner) units on top of a gather that element_sim_matrix (no ecc) would not need at all (17 more).
//          CA_SIM_SCAN=0/1 forces it.  Ratio-table launches sum the matrix while they convert the tables.
struct Route {
    bool rt, sim_scan;
};
static Route route_for(int32_t kmax, int64_t n, bool want_ecc, bool want_sim, bool allow_rt = true) {
    Route r{false, false};
    r.rt = want_ecc && kmax >= 1 && kmax <= 96 && (int64_t)kmax * 8192 <= n;
    if (const char* e = getenv("CA_RT")) r.rt = want_ecc && atoi(e) != 0;
    if (!allow_rt) r.rt = false;  // (the tables of one partner do not fit a third of a buffer)
    if (want_sim && !r.rt) {
        if (const char* e = getenv("CA_SIM_SCAN")) {
            r.sim_scan = atoi(e) != 0;
        } else {
            const double est = (std::ceil(kmax / 32.0) * 42.0 + 30.0) * 32.0 * (double)std::max(kmax, 1) / (double)std::max<int64_t>(n, 1);
            r.sim_scan = est < (want_ecc ? 6.0 : 18.0);
        }
    }
    return r;
}
extern "C" int ca_plan_route(int32_t kmax, int64_t n, int want_ecc, int want_sim, int* ratio_tables_out, int* sim_from_tables_out) {
    if (kmax < 0 || n < 0 || !ratio_tables_out || !sim_from_tables_out) {
        set_error("ca_plan_route: bad argument");
        return CA_E_ARG;
    }
    const Route r = route_for(kmax, n, want_ecc != 0, want_sim != 0);
    *ratio_tables_out = r.rt;
    *sim_from_tables_out = r.sim_scan;
    return CA_OK;
}

// The all-pairs job of one device: every owned partition i against its partners j in (i, m) -- or, in agreement mode,
// partition `ref_only` against all j in [0, m_partners).  Owned = the partitions i with zigzag(i) == rank of a whole label
// set, or every partition a SHARD holds (one process, several GPUs).  Sums are added to d_ecc / d_sim as fixed-point
// integers (ca_internal.h).  finish_sync = false leaves the stream running (the caller enqueues the reduction behind it and
// reads the stage timings with read_profile() after its own synchronisation).
static int read_profile(Ctx& c);
static int run_pairs_fallback(Labels& L, const double* weights_host, int32_t rank, int32_t world, double* d_ecc, double* d_sim, int32_t ref_only,
                              int32_t m_partners);
static int run_allpairs(Labels& L, const double* weights_host, int32_t rank, int32_t world, double* d_ecc, double* d_sim,
                        int32_t ref_only, int32_t m_partners, int plan_threads = 0, bool finish_sync = true,
                        const std::function<int()>* before_wave = nullptr) {
    Ctx& c = ctx();
    cudaStream_t s = c.stream;
    const int32_t m = L.mg();  // partitions of the whole set: the partner axis
    const int64_t n = L.n;
    auto t_call0 = std::chrono::steady_clock::now();
    for (double& v : c.prof) v = 0.0;
    c.launches = 0;
    CA_CUDA(cudaEventRecord(c.ev[0], s));
    const bool was_prepared = L.prepared;
    if (!was_prepared) {
        CA_TRY(prepare_ranks(L));
        if (!engine_supports(L.kmax)) return run_pairs_fallback(L, weights_host, rank, world, d_ecc, d_sim, ref_only, m_partners);
    }
    CA_TRY(prepare_labels(L));
    if (was_prepared) {
        CA_CUDA(cudaEventRecord(c.ev[1], s));
        CA_CUDA(cudaEventRecord(c.ev[2], s));
    }
    if (n == 0) return CA_OK;

    // ---- host planning ----------------------------------------------------------------------------
    auto t_plan0 = std::chrono::steady_clock::now();
    const bool plan_dbg = getenv("CA_PLAN_DEBUG") != nullptr;
    auto lap = [&](const char* what) {
        if (plan_dbg) fprintf(stderr, "[plan dev %d] %-28s %.3f ms\n", c.device, what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_plan0).count());
    };
    const Route route = route_for(L.kmax, n, d_ecc != nullptr, d_sim != nullptr);
    bool use_rt = route.rt;
    int32_t max_rows = 0;
    int smem_bytes = tile_smem_budget(L.kmax, &max_rows, nullptr, use_rt);
    if (use_rt && max_rows < 1) {
        use_rt = false;
        smem_bytes = tile_smem_budget(L.kmax, &max_rows);
    }
    if (max_rows < 1) {
        set_error("%d clusters per partition need more shared memory per table row than one SM has", L.kmax);
        return CA_E_UNSUPPORTED;
    }
    std::vector<int32_t> parts;    // owned partitions, in slot order: index into this label set's arrays ...
    std::vector<int32_t> parts_g;  // ... and their ids in the whole set
    if (ref_only >= 0) {
        parts.push_back(ref_only);
    } else if (L.m_glob) {
        for (int32_t k = 0; k < L.m; k++)
            if (L.gid[k] < m - 1) parts.push_back(k);
    } else {
        for (int32_t i = 0; i < m - 1; i++)
            if (owner_of(i, world) == rank) parts.push_back(i);
    }
    const int32_t nparts = (int32_t)parts.size();
    parts_g.resize(nparts);
    for (int32_t t = 0; t < nparts; t++) parts_g[t] = L.gid_of(parts[t]);
    bool qinfo_done = false;          // the sort wrote the quad info too (chunk sort)
    std::vector<size_t> off_kind[2];  // items of each kind before slot sl
    Item* items_kind[2] = {nullptr, nullptr};  // [0] resident, [1] streaming: views into the pinned staging buffer, in slot order
    size_t n_kind[2] = {0, 0};
    // every cluster pads to a multiple of TL_CPT positions: an upper bound that does not wait for the planner
    const int64_t perm_stride = ((int64_t)n + (int64_t)(TL_CPT - 1) * L.kmax + TL_CPT + 3) & ~3LL;
    if (perm_stride > 0x7fffff00LL) {
        set_error("%lld padded positions exceed the int32 position range", (long long)perm_stride);
        return CA_E_UNSUPPORTED;
    }
    Workspace& W = ws();
    // the previous call's copies out of the pinned staging buffers have completed: every batched call ends with a
    // stream synchronisation.  Entries beyond a partition's K are never read, so the buffers are not cleared.
    const size_t ncl = (size_t)std::max(nparts, 1) * L.kp;
    CA_TRY(W.h_cstart.ensure(ncl * 4));
    CA_TRY(W.h_crow.ensure(ncl * 2));
    int32_t* const h_cstart = W.h_cstart.as<int32_t>();
    uint16_t* const h_crow = W.h_crow.as<uint16_t>();
    CA_TRY(W.qinfo.ensure((size_t)std::max(nparts, 1) * perm_stride * 2));
    CA_TRY(W.perm.ensure((size_t)std::max(nparts, 1) * perm_stride * 4));
    CA_TRY(W.cstart.ensure(ncl * 4));
    CA_TRY(W.crow.ensure(ncl * 2));
    CA_TRY(W.parts.ensure((size_t)std::max(nparts, 1) * 4));
    // small host arrays go through one pinned block: [parts | wu_prefix of both kinds | weights]
    const size_t small_bytes = (size_t)std::max(nparts, 1) * 4 + 2 * ((size_t)(m + TL_WU_SPAN) / TL_WU_SPAN + 2) * 4 + 64 + ((size_t)L.mpad + 8) * 8;
    CA_TRY(W.h_small.ensure(small_bytes));
    if (nparts) {
        memcpy(W.h_small.p, parts.data(), (size_t)nparts * 4);
        CA_CUDA(cudaMemcpyAsync(W.parts.p, W.h_small.p, (size_t)nparts * 4, cudaMemcpyHostToDevice, s));
    }
    CA_CUDA(cudaEventRecord(c.ev[3], s));
    lap("buffers ready");
    {   // Plan every owned partition; partitions are independent, so the host threads share them.  The slots are cut
        // into batches: as soon as a batch is planned its cluster starts go to the device and its counting sort is
        // launched, so the GPU sorts batch b while the host plans batch b + 1.
        int nthreads = std::max(1, std::min<int>({(int)std::thread::hardware_concurrency(), 32, (nparts + 7) / 8}));
        if (plan_threads > 0) nthreads = std::max(1, std::min(nthreads, plan_threads));
        const int nbatch = nthreads > 1 ? (nparts >= 256 ? 8 : nparts >= 64 ? 4 : nparts >= 24 ? 2 : 1) : 1;
        WorkerPool& pool = cur().pool;
        std::vector<std::vector<Item>> items_s(nparts);  // per slot, the long (streaming) items first
        std::vector<int32_t> nstream(std::max(nparts, 1), 0);  // per slot: how many of them are streaming items
        std::vector<int32_t> bnd(nbatch + 1);
        for (int b = 0; b <= nbatch; b++) bnd[b] = (int32_t)((int64_t)b * nparts / nbatch);
        std::vector<std::atomic<int>> left(nbatch);
        for (int b = 0; b < nbatch; b++) left[b].store(bnd[b + 1] - bnd[b]);
        std::atomic<int> next{0}, err{CA_OK};
        auto work = [&]() {
            PlanOut po;
            for (;;) {
                const int32_t sl = next.fetch_add(1);
                if (sl >= nparts) break;
                int b = 0;
                while (sl >= bnd[b + 1]) b++;
                if (err.load() == CA_OK) {
                    const int32_t p = parts[sl], pg = parts_g[sl];
                    int rc = plan_blocks(L.h_sizes.data() + (size_t)p * L.kp, L.h_k[p], TL_CPT, TL_CAP, max_rows, po);
                    if (rc == CA_OK && po.total_pos > perm_stride) rc = CA_E_UNSUPPORTED;
                    if (rc != CA_OK) {
                        err.store(rc);
                    } else {
                        for (int32_t k = 0; k < L.h_k[p]; k++) {
                            h_cstart[(size_t)sl * L.kp + k] = po.cstart[k];
                            h_crow[(size_t)sl * L.kp + k] = (uint16_t)po.crow[k];
                        }
                        std::vector<Item>& v = items_s[sl];
                        v.reserve(po.pos0.size());
                        for (size_t t = 0; t < po.pos0.size(); t++) {
                            Item it;
                            it.m = pg;
                            it.pos0 = po.pos0[t];
                            it.npos = po.npos[t];
                            it.nrows = po.nrows[t] | (po.wide[t] ? (1 << 30) : 0);  // bit 30: 32-bit counters
                            it.jbeg = ref_only >= 0 ? 0 : pg + 1;
                            it.jend = ref_only >= 0 ? m_partners : m;
                            it.sa = po.sa[t];
                            it.pad_ = sl;  // perm slot
                            v.push_back(it);
                        }
                        // (plan_blocks emits the streaming items first, longest first, then the resident ones; work units are handed
                        // out first come first served, so the order of the nearly equal resident items does not matter)
                        int32_t ns = 0;
                        for (const Item& it : v) ns += it.npos > TL_CAP ? 1 : 0;
                        nstream[sl] = ns;
                    }
                }
                left[b].fetch_sub(1, std::memory_order_release);
            }
        };
        if (nthreads > 1)
            pool.start(nthreads, work);
        else
            work();
        int launch_rc = CA_OK;
        for (int b = 0; b < nbatch; b++) {
            while (left[b].load(std::memory_order_acquire) > 0) std::this_thread::yield();
            const int32_t b0 = bnd[b], nb = bnd[b + 1] - bnd[b];
            if (err.load() != CA_OK || launch_rc != CA_OK || nb == 0) continue;
            cudaError_t e = cudaMemcpyAsync(W.cstart.as<int32_t>() + (size_t)b0 * L.kp, h_cstart + (size_t)b0 * L.kp, (size_t)nb * L.kp * 4,
                                            cudaMemcpyHostToDevice, s);
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(W.crow.as<uint16_t>() + (size_t)b0 * L.kp, h_crow + (size_t)b0 * L.kp, (size_t)nb * L.kp * 2,
                                    cudaMemcpyHostToDevice, s);
            if (e != cudaSuccess) {
                set_error("uploading the plan failed: %s", cudaGetErrorString(e));
                launch_rc = CA_E_CUDA;
                continue;
            }
            launch_rc = launch_sort(L.d_raw, n, L.d_min.as<int32_t>(), L.identity ? nullptr : L.d_rank.as<uint32_t>(), L.d_roff.as<int64_t>(),
                                    L.d_kcount.as<int32_t>(), W.cstart.as<int32_t>() + (size_t)b0 * L.kp, L.kp, W.parts.as<int32_t>() + b0, nb,
                                    W.perm.as<int32_t>() + (size_t)b0 * perm_stride, perm_stride, W.crow.as<uint16_t>() + (size_t)b0 * L.kp,
                                    L.d_sizes.as<int32_t>(), W.qinfo.as<uint32_t>() + (size_t)b0 * (perm_stride >> 2) * 2, &qinfo_done, s);
        }
        lap("last batch launched");
        if (nthreads > 1) pool.wait();
        lap("threads joined");
        if (err.load() != CA_OK) {
            set_error("item planning failed (code %d)", err.load());
            return err.load();
        }
        CA_TRY(launch_rc);
        CA_CUDA(cudaEventRecord(c.ev[4], s));
        // Work unit = (item, span of TL_WU_SPAN partners), numbered span-major.  With the items in the order
        // of their first partner (`parts` is ascending, so slot order is that order), the items that reach into
        // span s are a prefix of the list, and a work-unit number decodes with one prefix array.
        // Two kinds of items, two launches of the wave kernel (each kind gets its own block size and register allocation):
        // streaming items (one cluster with more than TL_CAP cells) first, resident items second.  Both lists keep the order of
        // the first partner, so "the items that reach into span s" is a prefix of each.  They are written straight
        // into the pinned staging buffer: [resident items][streaming items].
        // per-slot offsets of both kinds (the workers counted their slots' streaming items), then the pool copies
        off_kind[0].assign(nparts + 1, 0);
        off_kind[1].assign(nparts + 1, 0);
        for (int32_t sl = 0; sl < nparts; sl++) {
            off_kind[1][sl + 1] = off_kind[1][sl] + (size_t)nstream[sl];
            off_kind[0][sl + 1] = off_kind[0][sl] + (items_s[sl].size() - (size_t)nstream[sl]);
        }
        n_kind[0] = off_kind[0][nparts];
        n_kind[1] = off_kind[1][nparts];
        CA_TRY(W.h_items.ensure(std::max<size_t>(n_kind[0] + n_kind[1], 1) * sizeof(Item)));
        items_kind[0] = W.h_items.as<Item>();
        items_kind[1] = items_kind[0] + n_kind[0];
        std::atomic<int> next_slot{0};
        auto fill = [&]() {
            for (;;) {
                const int32_t s0 = next_slot.fetch_add(16);
                if (s0 >= nparts) break;
                for (int32_t sl = s0; sl < std::min(nparts, s0 + 16); sl++) {
                    size_t o[2] = {off_kind[0][sl], off_kind[1][sl]};
                    for (const Item& it : items_s[sl]) {
                        const int kind = it.npos > TL_CAP ? 1 : 0;
                        items_kind[kind][o[kind]++] = it;
                    }
                }
            }
        };
        if (nthreads > 1 && nparts >= 64) {
            pool.start(nthreads, fill);
            pool.wait();
        } else {
            fill();
        }
        lap("items merged");
    }
    // work units before span sp = items whose first partner lies before the end of span sp, summed over the spans so far.
    // All items of a slot share their partner range, so this walks the slots (off_kind = items before a slot), not the items.
    std::vector<int32_t> h_wu_prefix_kind[2];
    for (int kind = 0; kind < 2; kind++) {
        std::vector<int32_t>& pre = h_wu_prefix_kind[kind];
        pre.assign(1, 0);
        const int32_t jmax = n_kind[kind] ? (ref_only >= 0 ? m_partners : m) : 0;
        const int32_t nsuper = (jmax + TL_WU_SPAN - 1) / TL_WU_SPAN;
        int32_t t = 0;  // slots whose items have started
        int64_t tot = 0;
        for (int32_t sp = 0; sp < nsuper; sp++) {
            while (t < nparts && (ref_only >= 0 ? 0 : parts_g[t] + 1) < (sp + 1) * TL_WU_SPAN) t++;
            tot += (int64_t)off_kind[kind][t];
            if (tot > 0x7fffff00LL) {
                set_error("too many work units (%lld)", (long long)tot);
                return CA_E_UNSUPPORTED;
            }
            pre.push_back((int32_t)tot);
        }
    }
    std::vector<double> h_w(L.mpad + 8, 0.0);
    double wsum = 0.0;
    const int32_t mw = ref_only >= 0 ? m_partners : m;
    for (int32_t p = 0; p < mw; p++) {
        h_w[p] = weights_host ? weights_host[p] : 1.0;
        wsum += h_w[p];
    }
    double inv_norm;
    if (ref_only >= 0) {
        h_w[ref_only] = 1.0;           // agreement: (1/m) * sum_j ECS(ref, j)
        inv_norm = 1.0 / (double)m_partners;
    } else {
        inv_norm = 1.0 / (wsum * (wsum - 1.0) / 2.0);  // R/ECS.R:1471-1472
    }
    lap("kinds, prefixes, weights");
    c.prof[2] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_plan0).count();

    // ---- upload the item lists (cluster starts and the sort went out batch by batch above) -----------------------
    const size_t npre0 = h_wu_prefix_kind[0].size(), npre1 = h_wu_prefix_kind[1].size();
    CA_TRY(W.wu_prefix.ensure((npre0 + npre1) * 4));
    CA_TRY(W.items.ensure(std::max<size_t>(n_kind[0] + n_kind[1], 1) * sizeof(Item)));
    CA_TRY(W.w.ensure(h_w.size() * 8));
    {
        char* hs = static_cast<char*>(W.h_small.p) + (((size_t)std::max(nparts, 1) * 4 + 63) & ~(size_t)63);
        if ((size_t)(hs - static_cast<char*>(W.h_small.p)) + (npre0 + npre1) * 4 + 64 + h_w.size() * 8 > W.h_small.cap) {
            set_error("internal: pinned staging block too small");
            return CA_E_ARG;
        }
        memcpy(hs, h_wu_prefix_kind[0].data(), npre0 * 4);
        memcpy(hs + npre0 * 4, h_wu_prefix_kind[1].data(), npre1 * 4);
        CA_CUDA(cudaMemcpyAsync(W.wu_prefix.p, hs, (npre0 + npre1) * 4, cudaMemcpyHostToDevice, s));
        char* hw = hs + (((npre0 + npre1) * 4 + 63) & ~(size_t)63);
        memcpy(hw, h_w.data(), h_w.size() * 8);
        CA_CUDA(cudaMemcpyAsync(W.w.p, hw, h_w.size() * 8, cudaMemcpyHostToDevice, s));
    }
    if (n_kind[0] + n_kind[1])
        CA_CUDA(cudaMemcpyAsync(W.items.p, W.h_items.p, (n_kind[0] + n_kind[1]) * sizeof(Item), cudaMemcpyHostToDevice, s));

    // ---- the wave kernel ----------------------------------------------------------------------------
    // multi-GPU jobs: the stream waits here for the peers' stores (labels, cluster sizes) and builds the table templates; planning and
    // sort ran meanwhile
    if (before_wave) CA_TRY((*before_wave)());
    TileParams T;
    T.lab16 = L.lab16_all();
    T.n = n;
    T.perm = W.perm.as<int32_t>();
    T.sizes = L.sizes_all();
    T.tmpl = L.tmpl_all();
    T.tmpl_bytes = L.tmpl_bytes_all();
    T.kinfo = L.kinfo_all();
    {
        DevBuf& zeros = scratch(SCR_ZEROS);
        if (!zeros.p) {
            CA_TRY(zeros.ensure(256 * 1024));
            CA_CUDA(cudaMemsetAsync(zeros.p, 0, 256 * 1024, s));
        }
        T.zeros = zeros.p;
    }
    T.crow = W.crow.as<uint16_t>();
    T.w = W.w.as<double>();
    T.items = W.items.as<Item>();
    T.wu_prefix = W.wu_prefix.as<int32_t>();
    T.qinfo = W.qinfo.as<uint32_t>();
    T.nsuper = T.nwu = T.streaming = 0;
    T.ecc = reinterpret_cast<unsigned long long*>(d_ecc);  // fixed-point partial sums until launch_finish converts them
    T.sim = reinterpret_cast<unsigned long long*>(d_sim);
    T.sim_magic = sim_fix_magic(n);
    T.sim_scale = sim_fix_scale(n);
    T.rt = use_rt ? 1 : 0;
    T.sim_scan = route_for(L.kmax, n, d_ecc != nullptr, d_sim != nullptr, use_rt).sim_scan ? 1 : 0;
    T.perm_stride = perm_stride;
    T.m = m;
    T.mpad = L.mpad;
    T.kp = L.kp;
    T.smem_bytes = smem_bytes;
    T.inv_norm = inv_norm;
    T.unit_w = 1;
    T.dbg = 0;
    T.sel[0] = 0x4410u;
    T.sel[1] = 0x4432u;
#ifdef CA_ABLATE
    if (const char* e = getenv("CA_ABLATE")) T.dbg = atoi(e);
#endif
    for (int32_t p = 0; p < mw; p++)
        if (h_w[p] != 1.0) T.unit_w = 0;
    if (!qinfo_done && nparts > 0) {
        if (L.m_glob) {
            set_error("CA_SORT_LEGACY is not available for sharded label sets");
            return CA_E_UNSUPPORTED;
        }
        CA_TRY(launch_qinfo(T, W.parts.as<int32_t>(), nparts, s));
    }
    CA_CUDA(cudaEventRecord(c.ev[4], s));
    DevBuf& wu_ctr = scratch(SCR_WUCTR);  // one hand-out counter per launch
    CA_TRY(wu_ctr.ensure(8));
    CA_CUDA(cudaMemsetAsync(wu_ctr.p, 0, 8, s));
    const bool wu_static = getenv("CA_WU_STATIC") != nullptr;  // tuning: work units dealt round-robin as in round 1
    for (int kind = 1; kind >= 0; kind--) {  // the long streaming items first
        T.wu_counter = wu_static ? nullptr : wu_ctr.as<int32_t>() + kind;
        T.items = W.items.as<Item>() + (kind == 1 ? n_kind[0] : 0);
        T.wu_prefix = W.wu_prefix.as<int32_t>() + (kind == 1 ? npre0 : 0);
        T.nsuper = (int32_t)h_wu_prefix_kind[kind].size() - 1;
        T.nwu = h_wu_prefix_kind[kind].back();
        T.streaming = kind;
        CA_TRY(launch_tile(T, s));
        if (kind == 1) CA_CUDA(cudaEventRecord(c.ev[8], s));  // between the streaming and the resident launch
    }
    CA_CUDA(cudaEventRecord(c.ev[5], s));
    c.prof[6] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call0).count();
    if (!finish_sync) return CA_OK;
    CA_CUDA(cudaStreamSynchronize(s));
    CA_TRY(read_profile(c));
    c.prof[6] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call0).count();
    return CA_OK;
}

// stage timings of the last run_allpairs on this device (its stream must have been synchronised)
static int read_profile(Ctx& c) {
    float ms;
    CA_CUDA(cudaEventElapsedTime(&ms, c.ev[0], c.ev[1]));
    c.prof[0] = ms;
    CA_CUDA(cudaEventElapsedTime(&ms, c.ev[1], c.ev[2]));
    c.prof[1] = ms;
    CA_CUDA(cudaEventElapsedTime(&ms, c.ev[3], c.ev[4]));
    c.prof[3] = ms;
    CA_CUDA(cudaEventElapsedTime(&ms, c.ev[4], c.ev[5]));
    c.prof[4] = ms;
    if (cudaEventElapsedTime(&ms, c.ev[4], c.ev[8]) == cudaSuccess) c.prof[8] = ms;  // the streaming launch ...
    if (cudaEventElapsedTime(&ms, c.ev[8], c.ev[5]) == cudaSuccess) c.prof[9] = ms;  // ... and the resident one
    (void)cudaGetLastError();
    c.prof[7] = (double)c.launches;  // every kernel of this call (pack, transpose, sort, wave)
    return CA_OK;
}

static double ident_constant(const double* weights_host, int32_t m) {
    // sum_i w_i/norm*(w_i-1)/2  (R/ECS.R:1475,1491)
    double wsum = 0.0;
    for (int32_t p = 0; p < m; p++) wsum += weights_host ? weights_host[p] : 1.0;
    const double norm = wsum * (wsum - 1.0) / 2.0;
    double c0 = 0.0;
    for (int32_t p = 0; p < m; p++) {
        double w = weights_host ? weights_host[p] : 1.0;
        c0 += w / norm * (w - 1.0) / 2.0;
    }
    return c0;
}

// ---- cluster counts beyond the wave engine: a loop over the pairs ---------------------------------------------------
// The batched engine keeps one table row per shared-memory buffer and uint16 labels: K <= 24 799 (one row of 4 (K + 1) bytes in
// half of 195 KB).  The reference has no such limit (src/calculate_ecs.cpp:7-19 allocates whatever K1 x K2 needs), so beyond
// it every owned pair (i, j) is processed on its own on the device -- no re-upload, dense ranks as labels, a hashed contingency
// table in global memory (ca_pair.cu: the table of two partitions with 10^5 clusters each would be 40 GB dense, but has at
// most N non-empty cells).  O(pairs x N) global atomics: slow next to the engine, but any cluster count that fits memory works.
static int run_pairs_fallback(Labels& L, const double* weights_host, int32_t rank, int32_t world, double* d_ecc, double* d_sim, int32_t ref_only,
                              int32_t m_partners) {
    Ctx& c = ctx();
    cudaStream_t s = c.stream;
    if (L.m_glob) {
        set_error("internal: the pair-loop fallback runs on whole label sets only");
        return CA_E_UNSUPPORTED;
    }
    const int32_t m = L.m;
    const int64_t n = L.n;
    if (n == 0) return CA_OK;
    const int32_t mw = ref_only >= 0 ? m_partners : m;
    double wsum = 0.0;
    for (int32_t p = 0; p < mw; p++) wsum += weights_host ? weights_host[p] : 1.0;
    const double inv_norm = ref_only >= 0 ? 1.0 / (double)m_partners : 1.0 / (wsum * (wsum - 1.0) / 2.0);
    const double sim_magic = sim_fix_magic(n);
    const uint32_t* rk = L.identity ? nullptr : L.d_rank.as<uint32_t>();
    auto one = [&](int32_t i, int32_t j, double wij) -> int {
        return pair_hashed_accumulate(L.d_raw + (size_t)i * n, L.d_raw + (size_t)j * n, n, L.d_min.as<int32_t>() + i, L.d_min.as<int32_t>() + j,
                                      rk ? rk + L.h_roff[i] : nullptr, rk ? rk + L.h_roff[j] : nullptr, L.d_sizes.as<int32_t>() + (size_t)i * L.kp,
                                      L.d_sizes.as<int32_t>() + (size_t)j * L.kp, wij * inv_norm,
                                      reinterpret_cast<unsigned long long*>(d_ecc),
                                      d_sim ? reinterpret_cast<unsigned long long*>(d_sim) + (size_t)i * m + j : nullptr, sim_magic, s);
    };
    if (ref_only >= 0) {
        for (int32_t j = 0; j < m_partners; j++) CA_TRY(one(ref_only, j, 1.0));
    } else {
        for (int32_t i = 0; i < m - 1; i++) {
            if (owner_of(i, world) != rank) continue;
            for (int32_t j = i + 1; j < m; j++)
                CA_TRY(one(i, j, (weights_host ? weights_host[i] : 1.0) * (weights_host ? weights_host[j] : 1.0)));
        }
    }
    CA_CUDA(cudaEventRecord(c.ev[1], s));
    CA_CUDA(cudaEventRecord(c.ev[2], s));
    CA_CUDA(cudaEventRecord(c.ev[3], s));
    CA_CUDA(cudaEventRecord(c.ev[4], s));
    CA_CUDA(cudaEventRecord(c.ev[8], s));
    CA_CUDA(cudaEventRecord(c.ev[5], s));
    CA_CUDA(cudaStreamSynchronize(s));
    return read_profile(c);
}

}  // namespace ca

using namespace ca;

// A label set: on one device (L), or sharded over the selected devices by 32-partition spans (sh[g] = device g's share).
struct ca_labels {
    Labels L;
    bool sharded = false;
    int64_t n = 0;
    int32_t m = 0;
    std::vector<std::unique_ptr<Labels>> sh;
    std::vector<int32_t> owner_of_span;
};
namespace ca {
Labels& labels_of(ca_labels_t* h) { return h->L; }

namespace {
// reusable barrier for the device threads of one multi-GPU job
class Barrier {
  public:
    explicit Barrier(int n) : n_(n) {}
    void wait() {
        std::unique_lock<std::mutex> g(mu_);
        const unsigned long gen = gen_;
        if (++count_ == n_) {
            count_ = 0;
            gen_++;
            cv_.notify_all();
        } else {
            cv_.wait(g, [&] { return gen_ != gen; });
        }
    }

  private:
    std::mutex mu_;
    std::condition_variable cv_;
    int n_, count_ = 0;
    unsigned long gen_ = 0;
};

// spans (32 partitions) -> devices: longest-processing-time greedy on the pair counts sum_{i in span} (m - 1 - i).  The spans'
// weights decrease with the span number, so this deals them 0..G-1, G-1..0, ...: the zigzag of ca_shard_owner at span granularity.
std::vector<int32_t> assign_spans(int32_t m, int ndev) {
    const int32_t nspan = (m + 31) / 32;
    std::vector<int32_t> owner(nspan, 0);
    std::vector<int64_t> load(ndev, 0);
    for (int32_t sp = 0; sp < nspan; sp++) {
        int64_t w = 0;
        for (int32_t i = sp * 32; i < std::min(m, sp * 32 + 32); i++) w += m - 1 - i;
        int best = 0;
        for (int g = 1; g < ndev; g++)
            if (load[g] < load[best]) best = g;
        owner[sp] = best;
        load[best] += w + 1;  // + 1: spans without pairs still spread out
    }
    return owner;
}
}  // namespace
}  // namespace ca

// Runs fn(g) on one host thread per selected device (thread-local device selection set up), returns the first error.
static int for_each_device(const std::function<int(int)>& fn) {
    const int G = (int)devs().size();
    std::vector<int> rcs(G, CA_OK);
    std::vector<std::string> msgs(G);
    std::vector<std::thread> th;
    th.reserve(G);
    auto body = [&](int g) {
        DevScope scope(devs()[g].get());
        rcs[g] = fn(g);
        if (rcs[g] != CA_OK) msgs[g] = g_err;
    };
    for (int g = 1; g < G; g++) th.emplace_back(body, g);
    body(0);
    for (auto& t : th) t.join();
    for (int g = 0; g < G; g++)
        if (rcs[g] != CA_OK) {
            set_error("%s", msgs[g].c_str());
            return rcs[g];
        }
    return CA_OK;
}

// ---- the multi-GPU job: one process, one host thread per device ----------------------------------------------------------
// Device g holds the raw labels of the spans it owns.  Per call (when the set is not prepared yet):
//   A  min / max, histograms, ranks of its partitions                                   | barrier: global K_max known, every lab16 allocated
//   B  cluster sizes (to the host, and by peer stores into every device's whole-set arrays), relabel + transpose of its spans into
//      its own lab16 array, then DMA copies of the finished slabs into the lab16 array of EVERY peer  | barrier: copies queued everywhere
//   C  all partitions' sizes / counts to the device, table templates
// then items, sort and the wave kernel for its own partitions against all partners, ONE ncclAllReduce (64-bit integer sums, so the
// result does not depend on the reduction order) on the same stream, and on the primary device finish + the copy to the host.
static int multi_allpairs(ca_labels* H, const double* weights, double* ecc_out, double* sim_out) {
    const int G = (int)devs().size();
    const int64_t n = H->n;
    const int32_t m = H->m;
    Barrier bar(G);
    std::atomic<int> failed{CA_OK};
    std::vector<int32_t> kmax_loc(G, 0);
    const int plan_threads = std::max(1, (int)std::thread::hardware_concurrency() / G);
    bool need_prepare = false;
    for (auto& sh : H->sh) need_prepare |= !sh->prepared;
    const int32_t mpad = (m + 31) & ~31;
    int rc = for_each_device([&](int g) -> int {
        Dev& D = *devs()[g];
        Ctx& c = D.c;
        cudaStream_t s = c.stream;
        Labels& L = *H->sh[g];
        int my = CA_OK;
        auto fail = [&](int code) {
            if (code != CA_OK && my == CA_OK) {
                my = code;
                int exp = CA_OK;
                failed.compare_exchange_strong(exp, code);
            }
        };
        auto t0 = std::chrono::steady_clock::now();
        bool hook_called = false;
        const bool dbg = getenv("CA_PIPE_DEBUG") != nullptr && (g == 0 || g == G - 1);
        auto lap = [&](const char* what) {
            if (dbg) fprintf(stderr, "[multi dev %d] %-28s %.3f ms\n", g, what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
        };
        if (need_prepare) {
            // ---- A ----
            fail(ranks_stage_a(L));
            if (my == CA_OK) fail(L.d_lab16.ensure((size_t)(n + 1) * mpad * 2));
            kmax_loc[g] = L.kmax;
            lap("A done");
            bar.wait();
            lap("A barrier");
            int32_t kmax_all = 0;
            for (int q = 0; q < G; q++) kmax_all = std::max(kmax_all, kmax_loc[q]);
            if (failed.load() == CA_OK && !engine_supports(kmax_all)) {
                set_error("a partition has %d clusters: beyond the batched engine's tables; run this job on one device (ca_init with one device) for the pair-loop fallback", kmax_all);
                fail(CA_E_UNSUPPORTED);
            }
            // ---- B ----
            if (failed.load() == CA_OK) {
                fail(ranks_stage_b(L, kmax_all));
                if (my == CA_OK) fail(L.d_sizes_g.ensure((size_t)m * L.kp * 4));
                if (my == CA_OK) fail(L.d_kcount_g.ensure((size_t)m * 4));
            }
            lap("B sizes");
            bar.wait();  // every device's whole-set arrays exist
            if (failed.load() == CA_OK) {
                if (L.m > 0) {
                    // This device's rows of the size / count arrays go to EVERY device by peer stores (small).  Its slabs of the label
                    // array are written locally by the transposition kernel and then copied into the peers' arrays by the DMA engines
                    // (one 64 MB copy per span and peer, spread over four copy streams): SM-issued 16-byte stores into seven peers reached
                    // 290 GB/s of NVLink egress, which put 3 ms in front of every wave kernel at 8 GPUs; the copy engines move the same
                    // bytes at line rate while the SMs sort.  CA_PEER_STORES=1 restores the store path.
                    static const bool by_stores = getenv("CA_PEER_STORES") != nullptr;
                    SizeDst sd;
                    LabDst dst;
                    sd.n = dst.n = 0;
                    for (int q = 0; q < G; q++) {
                        const int qq = (g + q) % G;  // own memory first
                        sd.sizes[sd.n] = H->sh[qq]->d_sizes_g.as<int32_t>();
                        sd.kcount[sd.n++] = H->sh[qq]->d_kcount_g.as<int32_t>();
                        if (by_stores || q == 0) dst.p[dst.n++] = H->sh[qq]->d_lab16.as<uint16_t>();
                    }
                    fail(launch_scatter_sizes(L.d_sizes.as<int32_t>(), L.d_kcount.as<int32_t>(), L.d_span.as<int32_t>(), L.m, L.kp, sd, s));
                    cudaError_t e = cudaEventRecord(c.ev[1], s);
                    if (e == cudaSuccess && my == CA_OK)
                        fail(launch_relabel_transpose(L.d_raw, n, L.m, (int32_t)L.spans.size(), L.d_span.as<int32_t>(), L.d_min.as<int32_t>(),
                                                      L.identity ? nullptr : L.d_rank.as<uint32_t>(), L.d_roff.as<int64_t>(), L.d_kcount.as<int32_t>(), dst, s));
                    if (e == cudaSuccess) e = cudaEventRecord(c.ev[2], s);
                    if (e == cudaSuccess && !by_stores && G > 1) {
                        e = cudaEventRecord(c.peer_ev[0], s);  // the slabs are complete in this device's memory
                        for (int k = 0; k < 4 && e == cudaSuccess; k++) e = cudaStreamWaitEvent(c.peer_stream[k], c.peer_ev[0], 0);
                        const size_t slab = (size_t)2 * 32 * (size_t)(n + 1);  // bytes of one span: two super-tiles of 32 (n + 1) bytes
                        int k = 0;
                        for (int q = 1; q < G && e == cudaSuccess; q++) {
                            const int qq = (g + q) % G;
                            for (size_t t = 0; t < L.spans.size() && e == cudaSuccess; t++, k++) {
                                const size_t off = (size_t)L.spans[t] * slab;
                                e = cudaMemcpyPeerAsync(static_cast<char*>(H->sh[qq]->d_lab16.p) + off, devs()[qq]->c.device, static_cast<const char*>(L.d_lab16.p) + off,
                                                        c.device, slab, c.peer_stream[k & 3]);
                            }
                        }
                        for (int k2 = 1; k2 < 4 && e == cudaSuccess; k2++) {  // stream 0 collects the others: one event tells the peers everything has landed
                            e = cudaEventRecord(c.peer_ev[k2], c.peer_stream[k2]);
                            if (e == cudaSuccess) e = cudaStreamWaitEvent(c.peer_stream[0], c.peer_ev[k2], 0);
                        }
                        if (e == cudaSuccess) e = cudaStreamWaitEvent(c.peer_stream[0], c.ev[2], 0);  // ... and the size stores issued on the main stream
                        if (e == cudaSuccess) e = cudaEventRecord(D.ev_lab, c.peer_stream[0]);
                    } else if (e == cudaSuccess) {
                        e = cudaEventRecord(D.ev_lab, s);
                    }
                    if (e != cudaSuccess) {
                        set_error("relabel/transpose into peer memory failed: %s", cudaGetErrorString(e));
                        fail(CA_E_CUDA);
                    }
                } else if (cudaEventRecord(D.ev_lab, s) != cudaSuccess) {  // nothing to send: the event still has to exist for the peers' waits
                    set_error("cudaEventRecord failed");
                    fail(CA_E_CUDA);
                }
            }
            lap("B peer stores queued");
        }
        const double prep_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        lap("C done");
        // ---- this device's pairs ----
        if (failed.load() == CA_OK) {
            if (ecc_out) fail(D.red_ecc.ensure((size_t)std::max<int64_t>(n, 1) * 8));
            if (sim_out && my == CA_OK) fail(D.red_sim.ensure((size_t)m * m * 8));
            cudaError_t e = cudaSuccess;
            if (my == CA_OK && ecc_out) e = cudaMemsetAsync(D.red_ecc.p, 0, (size_t)n * 8, s);
            if (my == CA_OK && e == cudaSuccess && sim_out) e = cudaMemsetAsync(D.red_sim.p, 0, (size_t)m * m * 8, s);
            if (e != cudaSuccess) {
                set_error("cudaMemsetAsync failed: %s", cudaGetErrorString(e));
                fail(CA_E_CUDA);
            }
            // the wave kernel reads every device's labels: its stream waits for all the label events (recorded above, or already
            // complete when the set was prepared by an earlier call); items and sort do not, they run while the stores are in flight
            // Everything the wave kernel reads from the peers -- compact labels, cluster sizes and counts -- arrives by their stores;
            // its stream waits for all the peers' events (a host barrier first: an event must have been recorded before it can be waited
            // for), then builds the table templates from the complete size arrays (stage C).  Items and sort need none of it: they
            // were planned and queued while the stores were in flight.
            const std::function<int()> wait_peers = [&]() -> int {
                if (!need_prepare) return CA_OK;
                hook_called = true;
                bar.wait();
                if (failed.load() != CA_OK) return failed.load();
                for (int q = 0; q < G; q++)
                    if (q != g) CA_CUDA(cudaStreamWaitEvent(s, devs()[q]->ev_lab, 0));
                CA_TRY(ranks_stage_c(L));
                L.ranked = L.prepared = true;
                return CA_OK;
            };
            if (my == CA_OK && L.m > 0 && n > 0) {
                L.prepared = need_prepare ? true : L.prepared;  // run_allpairs must not prepare a share on its own; stage C follows in the hook
                fail(run_allpairs(L, weights, 0, 1, ecc_out ? D.red_ecc.as<double>() : nullptr, sim_out ? D.red_sim.as<double>() : nullptr, -1, 0,
                                  plan_threads, /*finish_sync=*/false, &wait_peers));
                if (my != CA_OK) L.prepared = L.ranked = false;
            }
        }
        if (need_prepare && !hook_called) {  // a device without pairs (or one that failed early) still meets the others at the barrier,
            bar.wait();                       // and still builds its templates: a later call on the prepared set may give it work
            if (failed.load() == CA_OK) {
                for (int q = 0; q < G; q++)
                    if (q != g && cudaStreamWaitEvent(s, devs()[q]->ev_lab, 0) != cudaSuccess) fail(CA_E_CUDA);
                if (my == CA_OK) fail(ranks_stage_c(L));
                if (my == CA_OK) L.ranked = L.prepared = true;
            }
        }
        lap("wave queued");
        bar.wait();  // nobody enters the collective unless everybody will
        lap("collective barrier");
        if (failed.load() != CA_OK) return my;
        // ---- reduction ----
        if (g_have_comms) {
            ncclResult_t r = ncclSuccess;
            if (ecc_out) r = nccl().AllReduce(D.red_ecc.p, D.red_ecc.p, (size_t)n, ncclInt64, ncclSum, D.comm, s);
            if (r == ncclSuccess && sim_out) r = nccl().AllReduce(D.red_sim.p, D.red_sim.p, (size_t)m * m, ncclInt64, ncclSum, D.comm, s);
            if (r != ncclSuccess) {
                set_error("ncclAllReduce failed: %s", nccl().GetErrorString(r));
                return CA_E_CUDA;
            }
        } else {  // no NCCL: the primary device adds the peers' buffers through peer memory
            cudaError_t e = cudaStreamSynchronize(s);
            bar.wait();
            if (g == 0 && e == cudaSuccess) {
                PeerSrc src;
                src.n = 0;
                for (int q = 1; q < G; q++) src.p[src.n++] = static_cast<const unsigned long long*>(devs()[q]->red_ecc.p);
                if (ecc_out) fail(launch_peer_sum(static_cast<unsigned long long*>(D.red_ecc.p), src, n, s));
                src.n = 0;
                for (int q = 1; q < G; q++) src.p[src.n++] = static_cast<const unsigned long long*>(devs()[q]->red_sim.p);
                if (sim_out) fail(launch_peer_sum(static_cast<unsigned long long*>(D.red_sim.p), src, (int64_t)m * m, s));
                e = cudaStreamSynchronize(s);
            }
            bar.wait();  // the peers keep their buffers until the primary has read them
            if (e != cudaSuccess) {
                set_error("peer-memory reduction failed: %s", cudaGetErrorString(e));
                return CA_E_CUDA;
            }
            if (my != CA_OK) return my;
        }
        if (g == 0) {
            CA_CUDA(cudaEventRecord(c.ev[6], s));
            CA_TRY(launch_finish(ecc_out ? D.red_ecc.as<double>() : nullptr, n, ident_constant(weights, m), sim_out ? D.red_sim.as<double>() : nullptr, m, s));
            CA_CUDA(cudaEventRecord(c.ev[7], s));
            if (ecc_out && n > 0) CA_CUDA(cudaMemcpyAsync(ecc_out, D.red_ecc.p, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
            if (sim_out) CA_CUDA(cudaMemcpyAsync(sim_out, D.red_sim.p, (size_t)m * m * 8, cudaMemcpyDeviceToHost, s));
        }
        lap("reduce + finish queued");
        CA_CUDA(cudaStreamSynchronize(s));
        lap("stream done");
        if (L.m > 0 && n > 0) CA_TRY(read_profile(c));
        if (need_prepare) c.prof[0] = prep_ms;  // stages A-C including their barriers (host clock)
        c.prof[6] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (g == 0) {
            float ms;
            CA_CUDA(cudaEventElapsedTime(&ms, c.ev[6], c.ev[7]));
            c.prof[5] = ms;
        }
        return CA_OK;
    });
    return rc;
}

// =====================================================================================================
extern "C" {

int ca_version(void) { return 200; }
int ca_shard_owner(int32_t i, int32_t world) { return world > 0 && i >= 0 ? owner_of(i, world) : -1; }
const char* ca_last_error(void) { return g_err; }

int ca_init(const int* devices, int ndev) {
    if (devices && ndev > 0) return init_devices(devices, ndev);
    const int zero = 0;
    return init_devices(&zero, 1);
}
// host-only: how a set of m partitions is dealt to ndev devices (the span -> device map of sharded label sets), for CPU tests
int ca_plan_spans(int32_t m, int32_t ndev, int32_t* owner_out, int32_t cap) {
    if (m < 1 || ndev < 1 || ndev > CA_MAX_DEVICES || !owner_out) {
        set_error("ca_plan_spans: bad argument");
        return CA_E_ARG;
    }
    const std::vector<int32_t> o = assign_spans(m, ndev);
    if ((int32_t)o.size() > cap) {
        set_error("ca_plan_spans: %d spans exceed cap=%d", (int)o.size(), cap);
        return CA_E_ARG;
    }
    for (size_t i = 0; i < o.size(); i++) owner_out[i] = o[i];
    return (int)o.size();
}
int ca_device_count(void) {
    auto& D = devs();
    return (!D.empty() && D[0]->c.inited) ? (int)D.size() : 0;
}

void ca_shutdown(void) {
    std::lock_guard<std::mutex> g(g_init_mu);
    shutdown_all();
}

int ca_set_stream(void* cuda_stream) {
    CA_TRY(ensure_init());
    Ctx& c = ctx();
    c.stream = cuda_stream ? (cudaStream_t)cuda_stream : c.own_stream;
    return CA_OK;
}
int ca_use_stream(void* cuda_stream, int which) {
    CA_TRY(ensure_init());
    Ctx& c = ctx();
    switch (which) {
        case CA_STREAM_LIBRARY: c.stream = c.own_stream; break;
        case CA_STREAM_HANDLE: c.stream = (cudaStream_t)cuda_stream; break;  // NULL = the legacy default stream
        case CA_STREAM_LEGACY: c.stream = cudaStreamLegacy; break;
        case CA_STREAM_PER_THREAD: c.stream = cudaStreamPerThread; break;
        default: set_error("ca_use_stream: unknown selector %d", which); return CA_E_ARG;
    }
    return CA_OK;
}

int ca_get_profile(double* ms_out, int cap) {
    if (!ms_out || cap <= 0) return 0;
    int k = cap < 10 ? cap : 10;
    for (int i = 0; i < k; i++) ms_out[i] = ctx().prof[i];
    return k;
}
int ca_get_profile_dev(int index, double* ms_out, int cap) {
    auto& D = devs();
    if (!ms_out || cap <= 0 || index < 0 || index >= (int)D.size()) return 0;
    int k = cap < 10 ? cap : 10;
    for (int i = 0; i < k; i++) ms_out[i] = D[index]->c.prof[i];
    return k;
}

int ca_label_range(const int32_t* x, int64_t n, int32_t* min_out, int32_t* max_out) {
    if (!x || n <= 0 || !min_out || !max_out) {
        set_error("ca_label_range: NULL argument or n <= 0");
        return CA_E_ARG;
    }
    int32_t lo = x[0], hi = x[0];
    for (int64_t i = 1; i < n; i++) {
        lo = x[i] < lo ? x[i] : lo;
        hi = x[i] > hi ? x[i] : hi;
    }
    if (lo == INT32_MIN) {
        set_error("label vector contains NA_integer_ (INT_MIN)");
        return CA_E_NA_LABEL;
    }
    *min_out = lo;
    *max_out = hi;
    return CA_OK;
}

// ---- device-resident label sets ----------------------------------------------------------------------
static int create_single(int64_t n, int32_t m, ca_labels_t** out) {
    ca_labels* h = new (std::nothrow) ca_labels();
    if (!h) return CA_E_NOMEM;
    h->n = h->L.n = n;
    h->m = h->L.m = m;
    h->L.dev_index = cur().index;
    size_t bytes = (size_t)std::max<int64_t>(n, 1) * (size_t)m * 4;
    cudaError_t e = dev_alloc((void**)&h->L.d_raw, bytes);
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu bytes) for the label matrix failed: %s", bytes, cudaGetErrorString(e));
        delete h;
        return CA_E_NOMEM;
    }
    g_live_handles++;
    *out = h;
    return CA_OK;
}

static void release_labels(Labels& L) {
    if (L.d_raw && L.owns_raw) dev_free(L.d_raw, (size_t)std::max<int64_t>(L.n, 1) * (size_t)std::max(L.m, 1) * 4);
    L.d_raw = nullptr;
    L.d_min.release(); L.d_max.release(); L.d_cnt.release(); L.d_rank.release(); L.d_kcount.release();
    L.d_roff.release(); L.d_sizes.release(); L.d_lab16.release();
    L.d_spack.release(); L.d_kinfo.release(); L.d_tmpl.release();
    L.d_span.release(); L.d_sizes_g.release(); L.d_kcount_g.release();
}

static int create_sharded(int64_t n, int32_t m, ca_labels_t** out) {
    const int G = ndevices();
    ca_labels* h = new (std::nothrow) ca_labels();
    if (!h) return CA_E_NOMEM;
    h->sharded = true;
    h->n = n;
    h->m = m;
    h->owner_of_span = assign_spans(m, G);
    for (int g = 0; g < G; g++) {
        h->sh.emplace_back(new Labels());
        Labels& L = *h->sh[g];
        L.n = n;
        L.m_glob = m;
        L.dev_index = g;
        for (int32_t sp = 0; sp < (int32_t)h->owner_of_span.size(); sp++) {
            if (h->owner_of_span[sp] != g) continue;
            L.spans.push_back(sp);
            for (int32_t i = sp * 32; i < std::min(m, sp * 32 + 32); i++) L.gid.push_back(i);
        }
        L.m = (int32_t)L.gid.size();
    }
    int rc = for_each_device([&](int g) -> int {
        Labels& L = *h->sh[g];
        size_t bytes = (size_t)std::max<int64_t>(n, 1) * (size_t)std::max(L.m, 1) * 4;
        cudaError_t e = dev_alloc((void**)&L.d_raw, bytes);
        if (e != cudaSuccess) {
            set_error("cudaMalloc(%zu bytes) for device %d's share of the label matrix failed: %s", bytes, ctx().device, cudaGetErrorString(e));
            return CA_E_NOMEM;
        }
        CA_TRY(L.d_span.ensure(std::max<size_t>(L.spans.size(), 1) * 4));
        if (!L.spans.empty()) {
            CA_CUDA(cudaMemcpyAsync(L.d_span.p, L.spans.data(), L.spans.size() * 4, cudaMemcpyHostToDevice, ctx().stream));
            CA_CUDA(cudaStreamSynchronize(ctx().stream));
        }
        return CA_OK;
    });
    if (rc != CA_OK) {
        std::string msg = g_err;
        for_each_device([&](int g) -> int { release_labels(*h->sh[g]); return CA_OK; });
        delete h;
        set_error("%s", msg.c_str());
        return rc;
    }
    g_live_handles++;
    *out = h;
    return CA_OK;
}

// More than one device selected and more than one span of partitions: the set is sharded (a job of 32 partitions or fewer has
// nothing to share out).
static bool want_sharded(int32_t m) { return ndevices() > 1 && m > 32; }

int ca_labels_create(int64_t n, int32_t m, ca_labels_t** out) {
    if (!out || n < 0 || m < 1 || n > 0x7ffffff0LL) {
        set_error("ca_labels_create: bad arguments (n=%lld m=%d)", (long long)n, m);
        return CA_E_ARG;
    }
    CA_TRY(ensure_init());
    return want_sharded(m) ? create_sharded(n, m, out) : create_single(n, m, out);
}
int ca_labels_create_on_primary(int64_t n, int32_t m, ca_labels_t** out) {
    if (!out || n < 0 || m < 1 || n > 0x7ffffff0LL) {
        set_error("ca_labels_create_on_primary: bad arguments (n=%lld m=%d)", (long long)n, m);
        return CA_E_ARG;
    }
    CA_TRY(ensure_init());
    return create_single(n, m, out);
}

int ca_labels_wrap(int64_t n, int32_t m, void* device_matrix, ca_labels_t** out) {
    if (!out || !device_matrix || n < 0 || m < 1 || n > 0x7ffffff0LL) {
        set_error("ca_labels_wrap: bad arguments (n=%lld m=%d)", (long long)n, m);
        return CA_E_ARG;
    }
    CA_TRY(ensure_init());
    ca_labels* h = new (std::nothrow) ca_labels();
    if (!h) return CA_E_NOMEM;
    h->n = h->L.n = n;
    h->m = h->L.m = m;
    h->L.d_raw = static_cast<int32_t*>(device_matrix);
    h->L.owns_raw = false;
    h->L.dev_index = cur().index;
    g_live_handles++;
    *out = h;
    return CA_OK;
}

int ca_labels_shape(ca_labels_t* h, int64_t* n_out, int32_t* m_out) {
    if (!h) { set_error("ca_labels_shape: NULL label set"); return CA_E_ARG; }
    if (n_out) *n_out = h->n;
    if (m_out) *m_out = h->m;
    return CA_OK;
}
int ca_labels_is_sharded(ca_labels_t* h) { return h && h->sharded ? 1 : 0; }

static void invalidate(ca_labels* h) {
    h->L.prepared = h->L.ranked = false;
    for (auto& sh : h->sh) sh->prepared = sh->ranked = false;
}

// rows [first, first + count) of the set, given either as a matrix or as column pointers, to device g's share
static int upload_share(ca_labels* h, int g, const int32_t* labels, const int32_t* const* cols) {
    Labels& L = *h->sh[g];
    const int64_t n = h->n;
    std::vector<UploadSeg> segs;
    for (int32_t k = 0; k < L.m;) {
        if (labels) {  // consecutive global ids are contiguous in the host matrix: one segment per span
            int32_t e = k + 1;
            while (e < L.m && L.gid[e] == L.gid[e - 1] + 1) e++;
            segs.push_back({labels + (size_t)L.gid[k] * n, L.d_raw + (size_t)k * n, (size_t)(e - k) * n * 4});
            k = e;
        } else {
            if (!cols[L.gid[k]]) { set_error("NULL column pointer %d", L.gid[k]); return CA_E_ARG; }
            segs.push_back({cols[L.gid[k]], L.d_raw + (size_t)k * n, (size_t)n * 4});
            k++;
        }
    }
    const int staging_threads = std::max(1, (int)std::thread::hardware_concurrency() / 2 / ndevices());
    CA_TRY(upload_segments(segs.data(), segs.size(), ctx().stream, staging_threads));
    CA_CUDA(cudaStreamSynchronize(ctx().stream));
    return CA_OK;
}

int ca_labels_upload(ca_labels_t* h, const int32_t* labels) {
    if (!h || !labels) { set_error("ca_labels_upload: NULL argument"); return CA_E_ARG; }
    CA_TRY(ensure_init());
    invalidate(h);
    if (h->sharded) return for_each_device([&](int g) { return upload_share(h, g, labels, nullptr); });
    DevScope scope(devs()[h->L.dev_index].get());
    cudaStream_t s = ctx().stream;
    const UploadSeg seg = {labels, h->L.d_raw, (size_t)h->L.n * h->L.m * 4};
    CA_TRY(upload_segments(&seg, 1, s));
    CA_CUDA(cudaStreamSynchronize(s));
    return CA_OK;
}

int ca_labels_upload_col(ca_labels_t* h, int32_t p, const int32_t* col) {
    if (!h || !col || p < 0 || p >= h->m) { set_error("ca_labels_upload_col: bad argument"); return CA_E_ARG; }
    CA_TRY(ensure_init());
    invalidate(h);
    if (h->sharded) {
        const int g = h->owner_of_span[p / 32];
        Labels& L = *h->sh[g];
        const int32_t k = (int32_t)(std::lower_bound(L.gid.begin(), L.gid.end(), p) - L.gid.begin());
        DevScope scope(devs()[g].get());
        const UploadSeg seg = {col, L.d_raw + (size_t)k * h->n, (size_t)h->n * 4};
        CA_TRY(upload_segments(&seg, 1, ctx().stream));
        CA_CUDA(cudaStreamSynchronize(ctx().stream));
        return CA_OK;
    }
    DevScope scope(devs()[h->L.dev_index].get());
    cudaStream_t s = ctx().stream;
    const UploadSeg seg = {col, h->L.d_raw + (size_t)p * h->L.n, (size_t)h->L.n * 4};
    CA_TRY(upload_segments(&seg, 1, s));
    CA_CUDA(cudaStreamSynchronize(s));  // the source may be freed by the caller right after this returns
    return CA_OK;
}

int ca_labels_upload_cols(ca_labels_t* h, const int32_t* const* cols) {
    if (!h || !cols) { set_error("ca_labels_upload_cols: NULL argument"); return CA_E_ARG; }
    CA_TRY(ensure_init());
    invalidate(h);
    if (h->sharded) return for_each_device([&](int g) { return upload_share(h, g, nullptr, cols); });
    DevScope scope(devs()[h->L.dev_index].get());
    cudaStream_t s = ctx().stream;
    std::vector<UploadSeg> segs((size_t)h->L.m);
    for (int32_t p = 0; p < h->L.m; p++) {
        if (!cols[p]) { set_error("NULL column pointer %d", p); return CA_E_ARG; }
        segs[p] = {cols[p], h->L.d_raw + (size_t)p * h->L.n, (size_t)h->L.n * 4};
    }
    CA_TRY(upload_segments(segs.data(), segs.size(), s));
    CA_CUDA(cudaStreamSynchronize(s));
    return CA_OK;
}

void* ca_labels_device_ptr(ca_labels_t* h) { return (h && !h->sharded) ? (void*)h->L.d_raw : nullptr; }

int ca_labels_invalidate(ca_labels_t* h) {
    if (!h) return CA_E_ARG;
    invalidate(h);
    return CA_OK;
}

int ca_labels_destroy(ca_labels_t* h) {
    if (!h) return CA_OK;
    auto& D = devs();
    if (h->sharded) {
        if ((int)D.size() == (int)h->sh.size() && !D.empty() && D[0]->c.inited)
            for_each_device([&](int g) -> int { release_labels(*h->sh[g]); return CA_OK; });
    } else if (h->L.dev_index < (int)D.size() && D[h->L.dev_index]->c.inited) {
        DevScope scope(D[h->L.dev_index].get());
        release_labels(h->L);
    }
    delete h;
    g_live_handles--;
    return CA_OK;
}

int ca_consistency_partial_dev(ca_labels_t* h, const double* weights_host, int32_t rank, int32_t world, double* ecc_dev,
                               double* sim_dev) {
    if (!h || world < 1 || rank < 0 || rank >= world) { set_error("ca_consistency_partial_dev: bad argument"); return CA_E_ARG; }
    if (h->sharded) { set_error("ca_consistency_partial_dev works on single-device label sets (ranks shard over processes)"); return CA_E_UNSUPPORTED; }
    if (h->L.m < 2) { set_error("At least two clusterings are required to calculate the consistency."); return CA_E_ARG; }
    CA_TRY(ensure_init());
    if (h->L.n == 0) return CA_OK;
    return run_allpairs(h->L, weights_host, rank, world, ecc_dev, sim_dev, -1, 0);
}

int ca_consistency_finish_dev(int64_t n, int32_t m, const double* weights_host, double* ecc_dev, double* sim_dev) {
    CA_TRY(ensure_init());
    Ctx& c = ctx();
    CA_CUDA(cudaEventRecord(c.ev[6], c.stream));
    CA_TRY(launch_finish(ecc_dev, n, ident_constant(weights_host, m), sim_dev, m, c.stream));
    CA_CUDA(cudaEventRecord(c.ev[7], c.stream));
    CA_CUDA(cudaStreamSynchronize(c.stream));
    float ms;
    CA_CUDA(cudaEventElapsedTime(&ms, c.ev[6], c.ev[7]));
    c.prof[5] = ms;
    c.prof[7] = (double)c.launches;
    return CA_OK;
}

static void fill_empty_sim(double* sim_out, int32_t m) {  // nothing to compare: mean() of an empty ECS vector is NaN in R
    for (int32_t i = 0; i < m; i++)
        for (int32_t j = 0; j < m; j++) sim_out[(size_t)i * m + j] = i == j ? 1.0 : 0.0 / 0.0;
}

// the all-pairs job of a label set that is already on the device(s): results to host buffers
static int set_allpairs(ca_labels_t* h, const double* weights, double* ecc_out, double* sim_out) {
    const int64_t n = h->n;
    const int32_t m = h->m;
    if (n == 0) {
        if (sim_out) fill_empty_sim(sim_out, m);
        return CA_OK;
    }
    if (h->sharded) return multi_allpairs(h, weights, ecc_out, sim_out);
    DevScope scope(devs()[h->L.dev_index].get());
    cudaStream_t s = ctx().stream;
    DevBuf &d_ecc = scratch(SCR_ECC), &d_sim = scratch(SCR_SIM);
    if (ecc_out) CA_TRY(d_ecc.ensure((size_t)n * 8));
    if (sim_out) CA_TRY(d_sim.ensure((size_t)m * m * 8));
    if (ecc_out) CA_CUDA(cudaMemsetAsync(d_ecc.p, 0, (size_t)n * 8, s));
    if (sim_out) CA_CUDA(cudaMemsetAsync(d_sim.p, 0, (size_t)m * m * 8, s));
    if (m >= 2) CA_TRY(run_allpairs(h->L, weights, 0, 1, ecc_out ? d_ecc.as<double>() : nullptr, sim_out ? d_sim.as<double>() : nullptr, -1, 0));
    CA_TRY(ca_consistency_finish_dev(n, m, weights, ecc_out ? d_ecc.as<double>() : nullptr, sim_out ? d_sim.as<double>() : nullptr));
    if (ecc_out) CA_CUDA(cudaMemcpyAsync(ecc_out, d_ecc.p, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
    if (sim_out) CA_CUDA(cudaMemcpyAsync(sim_out, d_sim.p, (size_t)m * m * 8, cudaMemcpyDeviceToHost, s));
    CA_CUDA(cudaStreamSynchronize(s));
    return CA_OK;
}

// ---- the pipelined host path (one device): the upload of later partitions overlaps the wave kernel of earlier ones ----------
// A job from HOST buffers used to be upload (2 GB over PCIe: 36 ms at config 5) THEN compute.  Owner partition i only meets
// partners j > i, so once the LAST partitions are on the device their pairs can run while the rest is still crossing PCIe.  The
// partitions are cut into span-aligned groups from the top -- 1, 2, 4 spans, then the rest -- uploaded in that order on a second
// stream (its own thread: pageable sources are staged through pinned slots); each group is one share of a sharded label set whose
// shares all live on this device: as soon as its upload event fires it is packed (ranks, sizes, templates, its slabs of the compact
// label array -- all into the whole-set arrays of the container) and the pairs of ITS owners against every partition above them run
// as one launch set of the wave kernel.  The work available after a fraction f of the upload is f^2 of the job, so the engine stops
// waiting after the first few milliseconds (pageable sources, which upload at half the rate, get six finer groups).  Partner rows
// of the sizes / template arrays have one stride for the whole job; it is
// fixed from the first group with a 2x margin, and a later group that needs more makes the function bail out (handled = false)
// so that the caller takes the plain path.  Results are bit-identical to the plain path (integer sums).
// A host -> device copy that leaves room for others: the DMA engine serves copies in the order they were SUBMITTED, whatever
// their streams, so two gigabytes queued at once would make every small copy the packing and planning stages issue meanwhile
// (range offsets, cluster starts, items) wait for all of it -- measured: the second group's packing finished 37 ms after the
// start instead of 7.  And a copy queue that never runs empty starves the OTHER streams' copies on the same engine even when
// it is short (four 8 MB pieces in flight: still 16.7 ms, the engine only switches channels when the active one idles).
// Pinned sources therefore go out in 16 MB pieces, ONE in flight: the few microseconds between two pieces are where the small
// copies get in (measured at config 5: 247 -> 220 ms end to end; CA_PIPE_DEPTH / CA_PIPE_PIECE_MB to experiment).  Pageable
// sources are staged through the pinned slots of upload_segments in its paced mode, which lets the engine run empty after every
// four 4 MB chunks for the same reason (eight staging threads keep the queue full otherwise, and the groups' wave kernels -- whose
// item lists are small copies -- did not start before the whole upload had finished: 256 -> 225 ms; CA_PIPE_BATCH, CA_PIPE_UNPACED).
static int upload_throttled(const UploadSeg* segs, size_t nseg, cudaStream_t s, int max_threads) {
    bool pinned = true;
    for (size_t i = 0; i < nseg; i++)
        if (segs[i].bytes && is_pageable(segs[i].src)) pinned = false;
    if (!pinned) return upload_segments(segs, nseg, s, max_threads, /*paced=*/getenv("CA_PIPE_UNPACED") == nullptr);
    static const size_t PIECE = (size_t)(getenv("CA_PIPE_PIECE_MB") ? atoi(getenv("CA_PIPE_PIECE_MB")) : 16) << 20;
    static const int DEPTH_ENV = getenv("CA_PIPE_DEPTH") ? atoi(getenv("CA_PIPE_DEPTH")) : 1;
    constexpr int MAXDEPTH = 8;
    const int DEPTH = std::max(1, std::min(DEPTH_ENV, MAXDEPTH));
    cudaEvent_t ev[MAXDEPTH] = {};
    for (int k = 0; k < DEPTH; k++)
        if (cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming) != cudaSuccess) {
            (void)cudaGetLastError();
            for (int q = 0; q < k; q++) cudaEventDestroy(ev[q]);
            return upload_segments(segs, nseg, s, max_threads);
        }
    cudaError_t e = cudaSuccess;
    size_t issued = 0;
    for (size_t i = 0; i < nseg && e == cudaSuccess; i++)
        for (size_t off = 0; off < segs[i].bytes && e == cudaSuccess; off += PIECE, issued++) {
            const int k = (int)(issued % DEPTH);
            if (issued >= (size_t)DEPTH) {  // wait until the piece submitted DEPTH pieces ago has landed -- polling with short sleeps: a
                // blocking cudaEventSynchronize in this thread kept the driver busy enough to hold up the main thread's launches
                while ((e = cudaEventQuery(ev[k])) == cudaErrorNotReady) std::this_thread::sleep_for(std::chrono::microseconds(40));
            }
            if (e == cudaSuccess)
                e = cudaMemcpyAsync((char*)segs[i].dst + off, (const char*)segs[i].src + off, std::min(PIECE, segs[i].bytes - off), cudaMemcpyHostToDevice, s);
            if (e == cudaSuccess) e = cudaEventRecord(ev[k], s);
        }
    for (int k = 0; k < DEPTH; k++) cudaEventDestroy(ev[k]);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        set_error("host-to-device copy failed: %s", cudaGetErrorString(e));
        return CA_E_CUDA;
    }
    return CA_OK;
}

static int pipelined_allpairs(const int32_t* labels, const int32_t* const* cols, int64_t n, int32_t m, const double* weights, double* ecc_out,
                              double* sim_out, bool* handled) {
    *handled = false;
    const int32_t nspan = (m + 31) / 32;
    // Worth it from about half a gigabyte: every group costs 2-3 ms of host round trips and planning with the engine idle, which a
    // short upload cannot pay back (config 4, 200 MB: 25.1 ms pipelined against 24.4 ms plain from pinned memory, 45 against 35 ms
    // from a thousand separate pageable columns; config 5, 2 GB: 217 against 243 ms).  CA_PIPE_MIN_MB moves the threshold (tests).
    const int64_t min_mb = getenv("CA_PIPE_MIN_MB") ? atoll(getenv("CA_PIPE_MIN_MB")) : 512;
    if (ndevices() != 1 || nspan < 8 || (int64_t)n * m * 4 < (min_mb << 20) || getenv("CA_NO_PIPELINE")) return CA_OK;
    Dev& D = cur();
    Ctx& c = D.c;
    cudaStream_t s = c.stream;
    // Groups of spans from the top.  Pinned sources cross PCIe at the link rate (36 ms for config 5's 2 GB): 1, 2, 4 spans, then the rest --
    // the engine stops waiting after the third group.  Pageable sources (R's vectors) are staged through pinned slots by host threads
    // at about half that rate, so the LAST group must not hold most of the work: six groups of nspan x (1, 2, 3, 3, 3, 4) / 16, which leaves
    // 44 % of the pairs instead of 82 % for after the upload.
    const bool pageable_src = is_pageable(labels ? (const void*)labels : (const void*)cols[m - 1]);
    std::vector<std::pair<int32_t, int32_t>> groups;  // [first span, last span + 1)
    const bool fine = getenv("CA_PIPE_FINE") ? atoi(getenv("CA_PIPE_FINE")) != 0 : pageable_src;
    if (fine && nspan >= 16) {
        const int parts16[5] = {1, 2, 3, 3, 3};
        int32_t hi = nspan;
        for (int k = 0; k < 5; k++) {
            const int32_t sz = std::max(1, nspan * parts16[k] / 16);
            groups.push_back({hi - sz, hi});
            hi -= sz;
        }
        groups.push_back({0, hi});
    } else {
        int32_t hi = nspan, sz = 1;
        while (hi > 0 && (int)groups.size() < 3 && hi - sz >= nspan / 2) {
            groups.push_back({hi - sz, hi});
            hi -= sz;
            sz *= 2;
        }
        groups.push_back({0, hi});
    }
    const int G = (int)groups.size();
    ca_labels H;  // the container: whole-set arrays in H.L, one share per group
    H.sharded = true;
    H.n = n;
    H.m = m;
    Labels& WL = H.L;
    WL.n = n;
    WL.m = 0;
    WL.m_glob = m;
    WL.mpad = (m + 31) & ~31;
    for (int g = 0; g < G; g++) {
        H.sh.emplace_back(new Labels());
        Labels& L = *H.sh[g];
        L.n = n;
        L.m_glob = m;
        L.whole = &WL;
        for (int32_t sp = groups[g].first; sp < groups[g].second; sp++) {
            L.spans.push_back(sp);
            for (int32_t i = sp * 32; i < std::min(m, sp * 32 + 32); i++) L.gid.push_back(i);
        }
        L.m = (int32_t)L.gid.size();
    }
    int rc = CA_OK;
    std::atomic<int> enqueued{0}, up_rc{CA_OK};
    std::string up_msg;
    std::thread uploader;
    auto cleanup = [&]() {
        if (uploader.joinable()) uploader.join();
        cudaStreamSynchronize(c.up_stream);
        cudaStreamSynchronize(s);
        for (auto& sh : H.sh) release_labels(*sh);
        release_labels(WL);
    };
    // ---- allocations, then the uploads (all of them queued at once, in group order) ----
    for (int g = 0; g < G && rc == CA_OK; g++) {
        Labels& L = *H.sh[g];
        cudaError_t e = dev_alloc((void**)&L.d_raw, (size_t)n * (size_t)L.m * 4);
        if (e != cudaSuccess) {
            L.d_raw = nullptr;
            set_error("cudaMalloc for the label matrix failed: %s", cudaGetErrorString(e));
            rc = CA_E_NOMEM;
            break;
        }
        rc = L.d_span.ensure(L.spans.size() * 4);
        if (rc == CA_OK && cudaMemcpyAsync(L.d_span.p, L.spans.data(), L.spans.size() * 4, cudaMemcpyHostToDevice, s) != cudaSuccess) rc = CA_E_CUDA;
    }
    if (rc == CA_OK) rc = WL.d_lab16.ensure((size_t)(n + 1) * WL.mpad * 2);
    if (rc == CA_OK) rc = WL.d_kinfo.ensure((size_t)WL.mpad * 4);
    if (rc == CA_OK) rc = WL.d_kcount_g.ensure((size_t)m * 4);
    DevBuf &d_ecc = scratch(SCR_ECC), &d_sim = scratch(SCR_SIM);
    if (rc == CA_OK && ecc_out) rc = d_ecc.ensure((size_t)n * 8);
    if (rc == CA_OK && sim_out) rc = d_sim.ensure((size_t)m * m * 8);
    if (rc == CA_OK) {
        cudaError_t e = cudaMemsetAsync(WL.d_kinfo.p, 0, (size_t)WL.mpad * 4, s);
        if (e == cudaSuccess && ecc_out) e = cudaMemsetAsync(d_ecc.p, 0, (size_t)n * 8, s);
        if (e == cudaSuccess && sim_out) e = cudaMemsetAsync(d_sim.p, 0, (size_t)m * m * 8, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);  // the span ids are on the device before the uploader's stream touches anything
        if (e != cudaSuccess) {
            set_error("pipelined path: set-up failed: %s", cudaGetErrorString(e));
            rc = CA_E_CUDA;
        }
    }
    if (rc != CA_OK) {
        std::string msg = g_err;
        cleanup();
        set_error("%s", msg.c_str());
        return rc;
    }
    Dev* dev_ptr = &D;
    uploader = std::thread([&, dev_ptr]() {
        DevScope scope(dev_ptr);
        for (int g = 0; g < G; g++) {
            Labels& L = *H.sh[g];
            int urc = CA_OK;
            if (labels) {  // a group is a contiguous block of rows
                const UploadSeg seg = {labels + (size_t)L.gid[0] * n, L.d_raw, (size_t)L.m * n * 4};
                urc = upload_throttled(&seg, 1, c.up_stream, Stager::MAXT);
            } else {
                std::vector<UploadSeg> segs((size_t)L.m);
                for (int32_t k = 0; k < L.m && urc == CA_OK; k++) {
                    if (!cols[L.gid[k]]) {
                        set_error("NULL column pointer %d", L.gid[k]);
                        urc = CA_E_ARG;
                    }
                    segs[k] = {cols[L.gid[k]], L.d_raw + (size_t)k * n, (size_t)n * 4};
                }
                if (urc == CA_OK) urc = upload_throttled(segs.data(), segs.size(), c.up_stream, Stager::MAXT);
            }
            if (urc == CA_OK && cudaEventRecord(c.up_ev[g], c.up_stream) != cudaSuccess) {
                set_error("cudaEventRecord failed");
                urc = CA_E_CUDA;
            }
            if (urc != CA_OK) {
                up_msg = g_err;
                up_rc.store(urc);
            }
            enqueued.store(g + 1, std::memory_order_release);  // the main thread may now wait for this group's event
        }
    });
    // ---- the groups, in upload order ----
    int32_t kmax_sofar = 0, kp_fixed = 0;
    bool bail = false;
    const bool dbg = getenv("CA_PIPE_DEBUG") != nullptr;
    const auto t_pipe0 = std::chrono::steady_clock::now();
    auto lap = [&](int g, const char* what) {
        if (dbg) fprintf(stderr, "[pipe] group %d %-22s %.3f ms\n", g, what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_pipe0).count());
    };
    for (int g = 0; g < G && rc == CA_OK && !bail; g++) {
        Labels& L = *H.sh[g];
        lap(g, "start");
        while (enqueued.load(std::memory_order_acquire) <= g) std::this_thread::yield();
        lap(g, "upload enqueued");
        if (up_rc.load() != CA_OK) {
            set_error("%s", up_msg.c_str());
            rc = up_rc.load();
            break;
        }
        if (getenv("CA_PIPE_HOSTWAIT")) {
            cudaEventSynchronize(c.up_ev[g]);
            lap(g, "upload event (host wait)");
        }
        if (cudaStreamWaitEvent(s, c.up_ev[g], 0) != cudaSuccess) {
            set_error("cudaStreamWaitEvent failed");
            rc = CA_E_CUDA;
            break;
        }
        rc = ranks_stage_a(L);  // min / max, histograms, ranks of this group (its host syncs also end the previous group's kernels)
        if (rc != CA_OK) break;
        lap(g, "stage A done");
        kmax_sofar = std::max(kmax_sofar, L.kmax);
        if (!engine_supports(kmax_sofar)) {
            bail = true;  // the plain path decides (pair-loop fallback)
            break;
        }
        if (g == 0) {
            int32_t lim = kmax_sofar;
            while (engine_supports(lim + 1024) && lim < 2 * kmax_sofar + 64) lim += 1024;  // up to 2x the first group's cluster counts
            kp_fixed = std::max(8, (std::max(lim, kmax_sofar) + 1 + 7) & ~7);
            rc = WL.d_sizes_g.ensure((size_t)m * kp_fixed * 4);
            if (rc == CA_OK) rc = WL.d_spack.ensure((size_t)m * kp_fixed * 4);
            int32_t half = 0;
            tile_smem_budget(kmax_sofar, nullptr, &half);
            WL.tmpl_bytes = std::max(half, 16);
            if (rc == CA_OK) rc = WL.d_tmpl.ensure((size_t)m * (size_t)WL.tmpl_bytes);
            if (rc != CA_OK) break;
            WL.kp = kp_fixed;
        } else if (kmax_sofar + 1 > kp_fixed) {
            bail = true;  // this group has more clusters than the row stride fixed from the first group allows
            break;
        }
        rc = ranks_stage_b(L, kmax_sofar, kp_fixed);
        if (rc != CA_OK) break;
        WL.kmax = kmax_sofar;
        const int32_t p0 = L.gid[0];
        int32_t max_rows = 0;
        tile_smem_budget(kmax_sofar, &max_rows);
        cudaError_t e = cudaMemcpyAsync(static_cast<int32_t*>(WL.d_sizes_g.p) + (size_t)p0 * kp_fixed, L.d_sizes.p, (size_t)L.m * kp_fixed * 4, cudaMemcpyDeviceToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(static_cast<int32_t*>(WL.d_kcount_g.p) + p0, L.d_kcount.p, (size_t)L.m * 4, cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) {
            set_error("pipelined path: device copy failed: %s", cudaGetErrorString(e));
            rc = CA_E_CUDA;
            break;
        }
        // this group's rows of the partner arrays: initial table entries, K | flags, table templates; its slabs of the label array
        rc = launch_spack(static_cast<int32_t*>(WL.d_sizes_g.p) + (size_t)p0 * kp_fixed, static_cast<int32_t*>(WL.d_kcount_g.p) + p0, L.m, kp_fixed,
                          static_cast<uint32_t*>(WL.d_spack.p) + (size_t)p0 * kp_fixed, static_cast<int32_t*>(WL.d_kinfo.p) + p0, s);
        if (rc == CA_OK)
            rc = launch_table_template(static_cast<uint32_t*>(WL.d_spack.p) + (size_t)p0 * kp_fixed, static_cast<int32_t*>(WL.d_kcount_g.p) + p0, L.m, kp_fixed,
                                       WL.tmpl_bytes, std::max(max_rows, 1),
                                       reinterpret_cast<uint32_t*>(static_cast<char*>(WL.d_tmpl.p) + (size_t)p0 * (size_t)WL.tmpl_bytes), s);
        if (rc != CA_OK) break;
        LabDst dst;
        dst.n = 1;
        dst.p[0] = static_cast<uint16_t*>(WL.d_lab16.p);
        rc = launch_relabel_transpose(L.d_raw, n, L.m, (int32_t)L.spans.size(), L.d_span.as<int32_t>(), L.d_min.as<int32_t>(),
                                      L.identity ? nullptr : L.d_rank.as<uint32_t>(), L.d_roff.as<int64_t>(), L.d_kcount.as<int32_t>(), dst, s);
        if (rc != CA_OK) break;
        L.mpad = WL.mpad;
        L.ranked = L.prepared = true;
        lap(g, "packed");
        rc = run_allpairs(L, weights, 0, 1, ecc_out ? d_ecc.as<double>() : nullptr, sim_out ? d_sim.as<double>() : nullptr, -1, 0, 0, /*finish_sync=*/false);
        lap(g, "wave launched");
    }
    if (rc == CA_OK && !bail) {
        rc = launch_finish(ecc_out ? d_ecc.as<double>() : nullptr, n, ident_constant(weights, m), sim_out ? d_sim.as<double>() : nullptr, m, s);
        cudaError_t e = cudaSuccess;
        if (rc == CA_OK && ecc_out) e = cudaMemcpyAsync(ecc_out, d_ecc.p, (size_t)n * 8, cudaMemcpyDeviceToHost, s);
        if (rc == CA_OK && e == cudaSuccess && sim_out) e = cudaMemcpyAsync(sim_out, d_sim.p, (size_t)m * m * 8, cudaMemcpyDeviceToHost, s);
        if (rc == CA_OK && e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (rc == CA_OK && e != cudaSuccess) {
            set_error("pipelined path: device-to-host copy failed: %s", cudaGetErrorString(e));
            rc = CA_E_CUDA;
        }
        lap(G, "results on the host");
    }
    std::string msg = g_err;
    cleanup();
    lap(G, "released");
    if (rc != CA_OK) set_error("%s", msg.c_str());
    *handled = rc == CA_OK && !bail;
    return rc;
}

// ---- host-pointer batched entry points -----------------------------------------------------------------
static int host_allpairs(const int32_t* labels, const int32_t* const* cols, int64_t n, int32_t m, const double* weights,
                         double* ecc_out, double* sim_out) {
    if ((!labels && !cols) || n < 0 || m < 1) { set_error("bad arguments (NULL labels, n < 0 or m < 1)"); return CA_E_ARG; }
    CA_TRY(ensure_init());
    if (n == 0) {
        if (sim_out) fill_empty_sim(sim_out, m);
        return CA_OK;
    }
    if (m >= 2) {  // large single-device jobs: the upload overlaps the computation
        bool handled = false;
        CA_TRY(pipelined_allpairs(labels, cols, n, m, weights, ecc_out, sim_out, &handled));
        if (handled) return CA_OK;
    }
    ca_labels_t* h = nullptr;
    CA_TRY(ca_labels_create(n, m, &h));
    int rc = labels ? ca_labels_upload(h, labels) : ca_labels_upload_cols(h, cols);  // all columns in one staged upload
    if (rc == CA_OK) rc = set_allpairs(h, weights, ecc_out, sim_out);
    ca_labels_destroy(h);
    return rc;
}

// ---- resident label sets: the same job without the upload ----------------------------------------------------
int ca_labels_select(ca_labels_t* src, const int32_t* idx, int32_t k, ca_labels_t** out) {
    if (!src || !idx || !out || k < 1) { set_error("ca_labels_select: bad argument"); return CA_E_ARG; }
    for (int32_t t = 0; t < k; t++)
        if (idx[t] < 0 || idx[t] >= src->m) { set_error("ca_labels_select: index %d out of range [0, %d)", idx[t], src->m); return CA_E_ARG; }
    CA_TRY(ensure_init());
    const int64_t n = src->n;
    const size_t row = (size_t)n * 4;
    // where partition p of src lives: device and local row
    auto locate = [&](int32_t p, int& g, const int32_t*& base) {
        if (!src->sharded) {
            g = src->L.dev_index;
            base = src->L.d_raw + (size_t)p * n;
            return;
        }
        g = src->owner_of_span[p / 32];
        Labels& L = *src->sh[g];
        const int32_t kk = (int32_t)(std::lower_bound(L.gid.begin(), L.gid.end(), p) - L.gid.begin());
        base = L.d_raw + (size_t)kk * n;
    };
    ca_labels_t* h = nullptr;
    CA_TRY(ca_labels_create(n, k, &h));
    int rc = CA_OK;
    // every destination row is one device-to-device copy (peer copies when source and destination differ)
    for (int32_t t = 0; t < k && row > 0 && rc == CA_OK; t++) {
        int gs;
        const int32_t* sp;
        locate(idx[t], gs, sp);
        int gd;
        int32_t* dp;
        if (h->sharded) {
            gd = h->owner_of_span[t / 32];
            Labels& L = *h->sh[gd];
            const int32_t kk = (int32_t)(std::lower_bound(L.gid.begin(), L.gid.end(), t) - L.gid.begin());
            dp = L.d_raw + (size_t)kk * n;
        } else {
            gd = h->L.dev_index;
            dp = h->L.d_raw + (size_t)t * n;
        }
        DevScope scope(devs()[gd].get());
        cudaError_t err = gs == gd ? cudaMemcpyAsync(dp, sp, row, cudaMemcpyDeviceToDevice, ctx().stream)
                                   : cudaMemcpyPeerAsync(dp, devs()[gd]->c.device, sp, devs()[gs]->c.device, row, ctx().stream);
        if (err != cudaSuccess) {
            set_error("device-to-device copy failed: %s", cudaGetErrorString(err));
            rc = CA_E_CUDA;
        }
    }
    for (auto& d : devs()) {  // the copies are complete before the caller may drop or rewrite the source
        DevScope scope(d.get());
        if (rc == CA_OK && cudaStreamSynchronize(ctx().stream) != cudaSuccess) {
            set_error("device-to-device copy failed");
            rc = CA_E_CUDA;
        }
    }
    if (rc != CA_OK) {
        std::string msg = g_err;
        ca_labels_destroy(h);
        set_error("%s", msg.c_str());
        return rc;
    }
    *out = h;
    return CA_OK;
}

static int resident_allpairs(ca_labels_t* h, const int32_t* idx, int32_t k, const double* weights, double* ecc_out, double* sim_out) {
    if (!h || (idx && k < 1)) { set_error("bad arguments (NULL label set or empty index list)"); return CA_E_ARG; }
    CA_TRY(ensure_init());
    ca_labels_t* sub = nullptr;
    if (idx) CA_TRY(ca_labels_select(h, idx, k, &sub));
    ca_labels_t* use = sub ? sub : h;
    int rc = CA_OK;
    if (ecc_out && use->m < 2) {
        set_error("At least two clusterings are required to calculate the consistency.");
        rc = CA_E_ARG;
    } else {
        rc = set_allpairs(use, weights, ecc_out, sim_out);
    }
    if (sub) {
        std::string msg = g_err;
        ca_labels_destroy(sub);
        if (rc != CA_OK) set_error("%s", msg.c_str());
    }
    return rc;
}

int ca_consistency_resident(ca_labels_t* h, const int32_t* idx, int32_t k, const double* weights, double* ecc_out, double* sim_out) {
    if (!ecc_out) { set_error("ca_consistency_resident: ecc_out is NULL"); return CA_E_ARG; }
    return resident_allpairs(h, idx, k, weights, ecc_out, sim_out);
}
int ca_sim_matrix_resident(ca_labels_t* h, const int32_t* idx, int32_t k, double* sim_out) {
    if (!sim_out) { set_error("ca_sim_matrix_resident: sim_out is NULL"); return CA_E_ARG; }
    return resident_allpairs(h, idx, k, nullptr, nullptr, sim_out);
}

int ca_consistency(const int32_t* labels, int64_t n, int32_t m, const double* weights, double* ecc_out, double* sim_out) {
    if (m < 2) { set_error("At least two clusterings are required to calculate the consistency."); return CA_E_ARG; }
    if (!ecc_out) { set_error("ca_consistency: ecc_out is NULL"); return CA_E_ARG; }
    return host_allpairs(labels, nullptr, n, m, weights, ecc_out, sim_out);
}
int ca_consistency_cols(const int32_t* const* cols, int64_t n, int32_t m, const double* weights, double* ecc_out, double* sim_out) {
    if (m < 2) { set_error("At least two clusterings are required to calculate the consistency."); return CA_E_ARG; }
    if (!ecc_out) { set_error("ca_consistency_cols: ecc_out is NULL"); return CA_E_ARG; }
    return host_allpairs(nullptr, cols, n, m, weights, ecc_out, sim_out);
}
int ca_sim_matrix(const int32_t* labels, int64_t n, int32_t m, double* sim_out) {
    if (!sim_out) { set_error("ca_sim_matrix: sim_out is NULL"); return CA_E_ARG; }
    return host_allpairs(labels, nullptr, n, m, nullptr, nullptr, sim_out);
}
int ca_sim_matrix_cols(const int32_t* const* cols, int64_t n, int32_t m, double* sim_out) {
    if (!sim_out) { set_error("ca_sim_matrix_cols: sim_out is NULL"); return CA_E_ARG; }
    return host_allpairs(nullptr, cols, n, m, nullptr, nullptr, sim_out);
}

static int agreement_impl(const int32_t* reference, const int32_t* labels, const int32_t* const* cols, int64_t n, int32_t m,
                          double* agr_out) {
    if (!reference || (!labels && !cols) || !agr_out || n < 0 || m < 1) { set_error("ca_agreement: bad argument"); return CA_E_ARG; }
    CA_TRY(ensure_init());
    if (n == 0) return CA_OK;
    DevScope scope(devs()[0].get());  // one reference against m partitions: a single device's job
    cudaStream_t s = ctx().stream;
    ca_labels_t* h = nullptr;
    CA_TRY(create_single(n, m + 1, &h));  // partitions 0..m-1 = the list, partition m = the reference
    int rc = CA_OK;
    DevBuf& d_out = scratch(SCR_OUT);
    do {
        if (labels) {
            const UploadSeg segs[2] = {{labels, h->L.d_raw, (size_t)n * m * 4}, {reference, h->L.d_raw + (size_t)n * m, (size_t)n * 4}};
            rc = upload_segments(segs, 2, s);
        } else {
            std::vector<const int32_t*> all(cols, cols + m);
            all.push_back(reference);
            rc = ca_labels_upload_cols(h, all.data());
        }
        if (rc) break;
        cudaError_t e = cudaSuccess;
        rc = d_out.ensure((size_t)std::max<int64_t>(n, 1) * 8);
        if (rc) break;
        e = cudaMemsetAsync(d_out.p, 0, (size_t)std::max<int64_t>(n, 1) * 8, s);
        if (e != cudaSuccess) { set_error("cudaMemsetAsync failed: %s", cudaGetErrorString(e)); rc = CA_E_CUDA; break; }
        rc = run_allpairs(h->L, nullptr, 0, 1, d_out.as<double>(), nullptr, m, m);
        if (rc) break;
        rc = launch_finish(d_out.as<double>(), n, 0.0, nullptr, 0, s);  // fixed point -> double
        if (rc) break;
        if (n > 0) e = cudaMemcpyAsync(agr_out, d_out.p, (size_t)n * 8, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { set_error("device-to-host copy failed: %s", cudaGetErrorString(e)); rc = CA_E_CUDA; }
    } while (0);
    std::string msg = g_err;
    ca_labels_destroy(h);
    if (rc != CA_OK) set_error("%s", msg.c_str());
    return rc;
}

int ca_agreement(const int32_t* reference, const int32_t* labels, int64_t n, int32_t m, double* agr_out) {
    return agreement_impl(reference, labels, nullptr, n, m, agr_out);
}
int ca_agreement_cols(const int32_t* reference, const int32_t* const* cols, int64_t n, int32_t m, double* agr_out) {
    return agreement_impl(reference, nullptr, cols, n, m, agr_out);
}

// ---- legacy per-pair entry points ------------------------------------------------------------------------
struct PairBufs {
    DevBuf &a, &b, &table, &sa, &sb, &err, &ecs, &partial, &out;
};
static PairBufs pb() {
    return PairBufs{scratch(SCR_PAIR_A),   scratch(SCR_PAIR_B),   scratch(SCR_PAIR_TABLE),   scratch(SCR_PAIR_SA), scratch(SCR_PAIR_SB),
                    scratch(SCR_PAIR_ERR), scratch(SCR_PAIR_ECS), scratch(SCR_PAIR_PARTIAL), scratch(SCR_PAIR_OUT)};
}

static int pair_upload(const int32_t* a, const int32_t* b, int64_t n, cudaStream_t s) {
    PairBufs B = pb();
    CA_TRY(B.a.ensure((size_t)std::max<int64_t>(n, 1) * 4));
    CA_TRY(B.b.ensure((size_t)std::max<int64_t>(n, 1) * 4));
    if (n > 0) {
        const UploadSeg segs[2] = {{a, B.a.p, (size_t)n * 4}, {b, B.b.p, (size_t)n * 4}};
        CA_TRY(upload_segments(segs, 2, s));
    }
    return CA_OK;
}

static int pair_tables(int64_t n, int32_t min_a, int32_t min_b, int32_t ka, int32_t kb, cudaStream_t s) {
    PairBufs B = pb();
    CA_TRY(B.table.ensure((size_t)ka * kb * 4));
    CA_TRY(B.sa.ensure((size_t)ka * 4));
    CA_TRY(B.sb.ensure((size_t)kb * 4));
    CA_TRY(B.err.ensure(4));
    return pair_contingency(B.a.as<int32_t>(), B.b.as<int32_t>(), n, min_a, min_b, ka, kb, B.table.as<uint32_t>(),
                            B.sa.as<uint32_t>(), B.sb.as<uint32_t>(), B.err.as<int32_t>(), s);
}

// dense table dimensions of a pair: both ranges can reach 2^32 - 1 (labels near +-2^31), so each is checked on its own before
// the product is formed
static int check_dims(int32_t min_a, int32_t max_a, int32_t min_b, int32_t max_b, int32_t* ka, int32_t* kb) {
    const int64_t ra = (int64_t)max_a - min_a + 1, rb = (int64_t)max_b - min_b + 1;
    if (ra > INT32_MAX || rb > INT32_MAX || ra > (1ll << 31) / rb) {
        set_error("contingency table %lld x %lld is too large", (long long)ra, (long long)rb);
        return CA_E_UNSUPPORTED;
    }
    *ka = (int32_t)ra;
    *kb = (int32_t)rb;
    return CA_OK;
}

static int pair_range_error(cudaStream_t s, const char* who) {  // the table kernel flags labels outside [min, min + k)
    int32_t err = 0;
    CA_CUDA(cudaMemcpyAsync(&err, pb().err.p, 4, cudaMemcpyDeviceToHost, s));
    CA_CUDA(cudaStreamSynchronize(s));
    if (err) {
        set_error("%s: a label lies outside [min, min+k)", who);
        return CA_E_RANGE;
    }
    return CA_OK;
}

int ca_contingency(const int32_t* a, const int32_t* b, int64_t n, int32_t min_a, int32_t min_b, int32_t ka, int32_t kb,
                   int32_t* table_out, int32_t* sizes_a, int32_t* sizes_b) {
    if (!a || !b || !table_out || n < 0 || ka < 1 || kb < 1) { set_error("ca_contingency: bad argument"); return CA_E_ARG; }
    if ((int64_t)ka * kb > (1ll << 31)) { set_error("contingency table %d x %d is too large", ka, kb); return CA_E_UNSUPPORTED; }
    CA_TRY(ensure_init());
    cudaStream_t s = ctx().stream;
    PairBufs B = pb();
    CA_TRY(pair_upload(a, b, n, s));
    CA_TRY(pair_tables(n, min_a, min_b, ka, kb, s));
    CA_CUDA(cudaMemcpyAsync(table_out, B.table.p, (size_t)ka * kb * 4, cudaMemcpyDeviceToHost, s));
    if (sizes_a) CA_CUDA(cudaMemcpyAsync(sizes_a, B.sa.p, (size_t)ka * 4, cudaMemcpyDeviceToHost, s));
    if (sizes_b) CA_CUDA(cudaMemcpyAsync(sizes_b, B.sb.p, (size_t)kb * 4, cudaMemcpyDeviceToHost, s));
    return pair_range_error(s, "ca_contingency");
}

int ca_ecs_pair(const int32_t* a, const int32_t* b, int64_t n, double* ecs_out, double* mean_out) {
    if (!a || !b || (!ecs_out && !mean_out) || n < 0) { set_error("ca_ecs_pair: bad argument"); return CA_E_ARG; }
    if (n == 0) { if (mean_out) *mean_out = 0.0 / 0.0; return CA_OK; }
    CA_TRY(ensure_init());
    cudaStream_t s = ctx().stream;
    PairBufs B = pb();
    int32_t min_a, max_a, min_b, max_b, ka, kb;
    CA_TRY(ca_label_range(a, n, &min_a, &max_a));
    CA_TRY(ca_label_range(b, n, &min_b, &max_b));
    CA_TRY(check_dims(min_a, max_a, min_b, max_b, &ka, &kb));
    CA_TRY(pair_upload(a, b, n, s));
    CA_TRY(pair_tables(n, min_a, min_b, ka, kb, s));
    CA_TRY(pair_range_error(s, "ca_ecs_pair"));
    CA_TRY(B.ecs.ensure((size_t)n * 8));
    CA_TRY(B.partial.ensure(4096 * 8));
    CA_TRY(B.out.ensure(8));
    int32_t nblocks = 0;
    CA_TRY(pair_ecs_gather(B.a.as<int32_t>(), B.b.as<int32_t>(), n, min_a, min_b, ka, B.table.as<uint32_t>(), B.sa.as<uint32_t>(),
                           B.sb.as<uint32_t>(), ecs_out ? B.ecs.as<double>() : nullptr, B.partial.as<double>(), &nblocks, s));
    if (mean_out) {
        CA_TRY(reduce_partials(B.partial.as<double>(), nblocks, 1.0 / (double)n, B.out.as<double>(), s));
        CA_CUDA(cudaMemcpyAsync(mean_out, B.out.p, 8, cudaMemcpyDeviceToHost, s));
    }
    if (ecs_out) CA_CUDA(cudaMemcpyAsync(ecs_out, B.ecs.p, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
    CA_CUDA(cudaStreamSynchronize(s));
    return CA_OK;
}

int ca_ecs_pair_mean(const int32_t* a, const int32_t* b, int64_t n, double* mean_out) {
    if (!a || !b || !mean_out || n < 0) { set_error("ca_ecs_pair_mean: bad argument"); return CA_E_ARG; }
    if (n == 0) { *mean_out = 0.0 / 0.0; return CA_OK; }
    CA_TRY(ensure_init());
    cudaStream_t s = ctx().stream;
    PairBufs B = pb();
    int32_t min_a, max_a, min_b, max_b, ka, kb;
    CA_TRY(ca_label_range(a, n, &min_a, &max_a));
    CA_TRY(ca_label_range(b, n, &min_b, &max_b));
    CA_TRY(check_dims(min_a, max_a, min_b, max_b, &ka, &kb));
    CA_TRY(pair_upload(a, b, n, s));
    CA_TRY(pair_tables(n, min_a, min_b, ka, kb, s));
    CA_TRY(pair_range_error(s, "ca_ecs_pair_mean"));
    CA_TRY(B.partial.ensure(4096 * 8));
    CA_TRY(B.out.ensure(8));
    int32_t nblocks = 0;
    CA_TRY(pair_mean_from_table(B.table.as<uint32_t>(), B.sa.as<uint32_t>(), B.sb.as<uint32_t>(), ka, kb, n, B.partial.as<double>(),
                                &nblocks, s));
    CA_TRY(reduce_partials(B.partial.as<double>(), nblocks, 1.0, B.out.as<double>(), s));
    CA_CUDA(cudaMemcpyAsync(mean_out, B.out.p, 8, cudaMemcpyDeviceToHost, s));
    CA_CUDA(cudaStreamSynchronize(s));
    return CA_OK;
}

int ca_merge_identical(const int32_t* labels, int64_t n, int32_t m, const double* freq_in, int32_t* keep_out, double* freq_out,
                       int32_t* u_out);  // defined in ca_dedup.cu
int ca_merge_identical_cols(const int32_t* const* cols, int64_t n, int32_t m, const double* freq_in, int32_t* keep_out,
                            double* freq_out, int32_t* u_out);  // defined in ca_dedup.cu

}  // extern "C"
