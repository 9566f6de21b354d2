// ca_api.cu -- C ABI entry points, context, device-resident label sets and the batched pipeline.
// See include/clustassess_b200.h for the contract and the reference lines every entry replaces.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <numeric>
#include <thread>

#include "ca_internal.h"

namespace ca {

// ---- errors / context -------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

Ctx& ctx() {
    static Ctx c;
    return c;
}

// ---- device memory: a small caching allocator ----------------------------------------------------------
// The host-pointer entry points (the ones the R shim binds) create and drop a label set per call; cudaMalloc /
// cudaFree of multi-gigabyte buffers cost more than the kernels of a small job and cudaFree synchronises the
// device.  Freed blocks are parked here and handed out again (best fit, at most 2x oversized); ca_shutdown()
// returns everything to the driver.
namespace {
struct CacheBlock {
    void* p;
    size_t bytes;
};
std::vector<CacheBlock>& cache() {
    static std::vector<CacheBlock> c;
    return c;
}
std::mutex& cache_mutex() {
    static std::mutex m;
    return m;
}
constexpr size_t CACHE_LIMIT = 96ull << 30;  // parked bytes beyond this go back to the driver
}  // namespace

cudaError_t dev_alloc(void** out, size_t bytes) {
    {
        std::lock_guard<std::mutex> g(cache_mutex());
        auto& c = cache();
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
    std::lock_guard<std::mutex> g(cache_mutex());
    auto& c = cache();
    size_t parked = bytes;
    for (auto& b : c) parked += b.bytes;
    if (parked > CACHE_LIMIT) {
        cudaFree(p);
        return;
    }
    c.push_back({p, bytes});
}
void dev_cache_release() {
    std::lock_guard<std::mutex> g(cache_mutex());
    for (auto& b : cache()) cudaFree(b.p);
    cache().clear();
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

static int do_init(int device) {
    Ctx& c = ctx();
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        set_error("no CUDA device available (%s); this library has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return CA_E_NO_DEVICE;
    }
    if (device < 0 || device >= count) {
        set_error("ca_init: device %d out of range (have %d)", device, count);
        return CA_E_ARG;
    }
    if (c.inited && c.device == device) return CA_OK;
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
    if (!c.stream) c.stream = c.own_stream;
    for (int i = 0; i < 8; i++)
        if (!c.ev[i]) CA_CUDA(cudaEventCreate(&c.ev[i]));
    c.inited = true;
    return CA_OK;
}

int ensure_init() {
    Ctx& c = ctx();
    if (c.inited) {
        CA_CUDA(cudaSetDevice(c.device));
        return CA_OK;
    }
    return do_init(0);
}

// ---- label preparation ------------------------------------------------------------------------------
static constexpr int64_t RANGE_LIMIT = 1ll << 24;        // per-partition label range (max-min+1)
static constexpr int64_t TOTAL_RANGE_LIMIT = 1ll << 28;  // summed over partitions
static constexpr int32_t K_LIMIT = 65535;                // compact labels are uint16

int prepare_ranks(Labels& L) {
    if (L.ranked) return CA_OK;
    Ctx& c = ctx();
    cudaStream_t s = c.stream;
    const int32_t m = L.m;
    const int64_t n = L.n;
    L.h_min.assign(m, 0);
    L.h_max.assign(m, 0);
    L.h_k.assign(m, 0);
    L.h_roff.assign(m + 1, 0);
    CA_TRY(L.d_min.ensure((size_t)m * 4));
    CA_TRY(L.d_max.ensure((size_t)m * 4));
    CA_TRY(launch_minmax(L.d_raw, n, m, L.d_min.as<int32_t>(), L.d_max.as<int32_t>(), s));
    CA_CUDA(cudaMemcpyAsync(L.h_min.data(), L.d_min.p, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
    CA_CUDA(cudaMemcpyAsync(L.h_max.data(), L.d_max.p, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
    CA_CUDA(cudaStreamSynchronize(s));
    int64_t max_range = 0;
    for (int32_t p = 0; p < m; p++) {
        if (L.h_min[p] == INT32_MIN) {
            set_error("partition %d contains NA_integer_ (INT_MIN); NA labels are not supported", p);
            return CA_E_NA_LABEL;
        }
        int64_t range = (int64_t)L.h_max[p] - (int64_t)L.h_min[p] + 1;
        if (range > RANGE_LIMIT) {
            set_error("partition %d: label range %lld exceeds the supported %lld", p, (long long)range, (long long)RANGE_LIMIT);
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
    CA_CUDA(cudaStreamSynchronize(s));
    L.kmax = 0;
    L.identity = true;
    for (int32_t p = 0; p < m; p++) {
        L.kmax = std::max(L.kmax, L.h_k[p]);
        if ((int64_t)L.h_k[p] != L.h_roff[p + 1] - L.h_roff[p]) L.identity = false;
    }
    if (L.kmax > K_LIMIT) {
        set_error("a partition has %d clusters; the engine supports at most %d", L.kmax, K_LIMIT);
        return CA_E_UNSUPPORTED;
    }
    L.kp = std::max(8, (L.kmax + 1 + 7) & ~7);  // + the pad column
    L.mpad = (m + 31) & ~31;
    CA_TRY(L.d_sizes.ensure((size_t)m * L.kp * 4));
    CA_CUDA(cudaMemsetAsync(L.d_sizes.p, 0, (size_t)m * L.kp * 4, s));
    CA_TRY(launch_sizes(L.d_cnt.as<uint32_t>(), L.d_rank.as<uint32_t>(), L.d_roff.as<int64_t>(), nullptr, nullptr, m,
                        (int32_t)max_range, L.kp, L.d_sizes.as<int32_t>(), s));
    // the wave kernel's initial table entries and K | flags per partition
    CA_TRY(L.d_spack.ensure((size_t)m * L.kp * 4));
    CA_TRY(L.d_kinfo.ensure((size_t)L.mpad * 4));
    CA_CUDA(cudaMemsetAsync(L.d_kinfo.p, 0, (size_t)L.mpad * 4, s));
    CA_TRY(launch_spack(L.d_sizes.as<int32_t>(), L.d_kcount.as<int32_t>(), m, L.kp, L.d_spack.as<uint32_t>(), L.d_kinfo.as<int32_t>(), s));
    {   // table templates: one table buffer of the wave kernel per partition
        int32_t max_rows = 0, half = 0;
        tile_smem_budget(L.kmax, &max_rows, &half);
        L.tmpl_bytes = std::max(half, 16);
        CA_TRY(L.d_tmpl.ensure((size_t)m * (size_t)L.tmpl_bytes));
        CA_TRY(launch_table_template(L.d_spack.as<uint32_t>(), L.d_kcount.as<int32_t>(), m, L.kp, L.tmpl_bytes, std::max(max_rows, 1),
                                     L.d_tmpl.as<uint32_t>(), s));
    }
    L.h_sizes.resize((size_t)m * L.kp);
    CA_CUDA(cudaMemcpyAsync(L.h_sizes.data(), L.d_sizes.p, (size_t)m * L.kp * 4, cudaMemcpyDeviceToHost, s));
    CA_CUDA(cudaStreamSynchronize(s));
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
    CA_CUDA(cudaEventRecord(c.ev[1], s));
    CA_TRY(L.d_lab16.ensure((size_t)(n + 1) * L.mpad * 2));  // n cells + the pad cell
    CA_TRY(launch_relabel_transpose(L.d_raw, n, m, L.mpad, L.d_min.as<int32_t>(),
                                    L.identity ? nullptr : L.d_rank.as<uint32_t>(), L.d_roff.as<int64_t>(),
                                    L.d_kcount.as<int32_t>(), L.d_lab16.as<uint16_t>(), s));
    CA_CUDA(cudaEventRecord(c.ev[2], s));  // no synchronisation: the host goes on to plan while the transposition runs
    L.prepared = true;
    return CA_OK;
}

// ---- host worker pool -----------------------------------------------------------------------------------------
// The item planner runs between two GPU stages of every batched call; creating its threads (and their thread-local
// scratch) per call cost as much as the planning.  The pool's threads live until ca_shutdown / process exit.
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
WorkerPool& pool() {
    static WorkerPool p;
    return p;
}
}  // namespace

// zigzag owner of partition i among `world` ranks: 0..w-1, w-1..0, ... (balances sum of M-1-i)
static inline int owner_of(int32_t i, int32_t world) {
    int32_t r = i % (2 * world);
    return r < world ? r : 2 * world - 1 - r;
}

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
};
static Workspace& ws() {
    static Workspace w;
    return w;
}

// ---- host -> device upload ----------------------------------------------------------------------------
// The R shim hands over R's own vectors, i.e. PAGEABLE memory.  cudaMemcpyAsync from pageable memory is staged by
// the driver through one small pinned buffer on one thread and reaches a fraction of the PCIe rate; at config 5
// (2 GB of labels) that copy would cost more than the whole computation.  Large pageable uploads are therefore
// staged here: up to 8 host threads copy 4 MB chunks into their own pair of pinned slots and enqueue the
// slot -> device DMA on the caller's stream, so the host memcpys of all threads overlap each other and the DMA.
// Pinned / registered sources (bench.py's e2e buffers, torch pinned tensors) and small uploads go straight to
// cudaMemcpyAsync.  CA_UPLOAD_DIRECT=1 forces the direct path (for A/B measurements).
struct UploadSeg {
    const void* src;
    void* dst;
    size_t bytes;
};
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
Stager& stager() {
    static Stager s;
    return s;
}
bool is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}
}  // namespace

static int upload_segments(const UploadSeg* segs, size_t nseg, cudaStream_t s) {
    size_t total = 0;
    for (size_t i = 0; i < nseg; i++) total += segs[i].bytes;
    if (total == 0) return CA_OK;
    bool staged = total >= 2 * Stager::CHUNK && getenv("CA_UPLOAD_DIRECT") == nullptr;
    for (size_t i = 0; staged && i < nseg; i++)
        if (segs[i].bytes && !is_pageable(segs[i].src)) staged = false;
    if (staged && stager().ensure() != CA_OK) staged = false;
    if (!staged) {
        for (size_t i = 0; i < nseg; i++)
            if (segs[i].bytes) CA_CUDA(cudaMemcpyAsync(segs[i].dst, segs[i].src, segs[i].bytes, cudaMemcpyHostToDevice, s));
        return CA_OK;
    }
    Stager& S = stager();
    std::vector<UploadSeg> chunks;
    chunks.reserve(total / Stager::CHUNK + nseg);
    for (size_t i = 0; i < nseg; i++)
        for (size_t off = 0; off < segs[i].bytes; off += Stager::CHUNK)
            chunks.push_back({(const char*)segs[i].src + off, (char*)segs[i].dst + off, std::min(Stager::CHUNK, segs[i].bytes - off)});
    const int nthreads = (int)std::max<size_t>(1, std::min<size_t>({(size_t)Stager::MAXT, (size_t)std::thread::hardware_concurrency() / 2, chunks.size()}));
    const int device = ctx().device;
    std::atomic<size_t> next{0};
    std::atomic<int> failed{(int)cudaSuccess};
    auto work = [&](int t) {
        if (cudaError_t e0 = cudaSetDevice(device)) { failed = (int)e0; return; }
        int k = 0;
        for (size_t c; (c = next.fetch_add(1)) < chunks.size() && failed == (int)cudaSuccess; k ^= 1) {
            cudaError_t e = S.used[t][k] ? cudaEventSynchronize(S.ev[t][k]) : cudaSuccess;  // the slot's previous DMA has read it
            if (e == cudaSuccess) {
                memcpy(S.slot(t, k), chunks[c].src, chunks[c].bytes);
                e = cudaMemcpyAsync(chunks[c].dst, S.slot(t, k), chunks[c].bytes, cudaMemcpyHostToDevice, s);
            }
            if (e == cudaSuccess) e = cudaEventRecord(S.ev[t][k], s);
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

// The all-pairs job (or one rank's share of it).  pairs: for every owned partition i, partners j in
// (i, m) -- or, in agreement mode, partition `ref_only` against all j in [0, m_partners).
static int run_allpairs(Labels& L, const double* weights_host, int32_t rank, int32_t world, double* d_ecc, double* d_sim,
                        int32_t ref_only, int32_t m_partners) {
    Ctx& c = ctx();
    cudaStream_t s = c.stream;
    const int32_t m = L.m;
    const int64_t n = L.n;
    auto t_call0 = std::chrono::steady_clock::now();
    for (double& v : c.prof) v = 0.0;
    c.launches = 0;
    CA_CUDA(cudaEventRecord(c.ev[0], s));
    const bool was_prepared = L.prepared;
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
        if (plan_dbg) fprintf(stderr, "[plan] %-28s %.3f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_plan0).count());
    };
    int32_t max_rows = 0;
    const int smem_bytes = tile_smem_budget(L.kmax, &max_rows);
    if (max_rows < 1) {
        set_error("%d clusters per partition need more shared memory per table row than one SM has", L.kmax);
        return CA_E_UNSUPPORTED;
    }
    std::vector<int32_t> parts;  // owned partitions, in slot order
    if (ref_only >= 0) {
        parts.push_back(ref_only);
    } else {
        for (int32_t i = 0; i < m - 1; i++)
            if (owner_of(i, world) == rank) parts.push_back(i);
    }
    const int32_t nparts = (int32_t)parts.size();
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
        const int nthreads = std::max(1, std::min<int>({(int)std::thread::hardware_concurrency(), 32, (nparts + 7) / 8}));
        const int nbatch = nthreads > 1 ? (nparts >= 256 ? 8 : nparts >= 64 ? 4 : 1) : 1;
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
                    const int32_t p = parts[sl];
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
                            it.m = p;
                            it.pos0 = po.pos0[t];
                            it.npos = po.npos[t];
                            it.nrows = po.nrows[t] | (po.wide[t] ? (1 << 30) : 0);  // bit 30: 32-bit counters
                            it.jbeg = ref_only >= 0 ? 0 : p + 1;
                            it.jend = ref_only >= 0 ? m_partners : m;
                            it.sa = po.sa[t];
                            it.pad_ = sl;  // perm slot
                            v.push_back(it);
                        }
                        std::stable_sort(v.begin(), v.end(), [](const Item& a, const Item& b) { return a.npos > b.npos; });
                        int32_t ns = 0;
                        for (const Item& it : v) ns += it.npos > TL_CAP ? 1 : 0;
                        nstream[sl] = ns;
                    }
                }
                left[b].fetch_sub(1, std::memory_order_release);
            }
        };
        if (nthreads > 1)
            pool().start(nthreads, work);
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
        if (nthreads > 1) pool().wait();
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
        // Two kinds of items, two launches of the wave kernel (each kind gets its own register allocation): streaming
        // items (one cluster with more than TL_CAP cells) first, resident items second.  Both lists keep the order of
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
            pool().start(nthreads, fill);
            pool().wait();
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
            while (t < nparts && (ref_only >= 0 ? 0 : parts[t] + 1) < (sp + 1) * TL_WU_SPAN) t++;
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
    TileParams T;
    T.lab16 = L.d_lab16.as<uint16_t>();
    T.n = n;
    T.perm = W.perm.as<int32_t>();
    T.sizes = L.d_sizes.as<int32_t>();
    T.tmpl = L.d_tmpl.as<uint32_t>();
    T.tmpl_bytes = L.tmpl_bytes;
    T.kinfo = L.d_kinfo.as<int32_t>();
    {
        static DevBuf zeros;
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
    T.ecc = d_ecc;
    T.sim = d_sim;
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
    if (!qinfo_done) CA_TRY(launch_qinfo(T, W.parts.as<int32_t>(), nparts, s));
    for (int kind = 1; kind >= 0; kind--) {  // the long streaming items first
        T.items = W.items.as<Item>() + (kind == 1 ? n_kind[0] : 0);
        T.wu_prefix = W.wu_prefix.as<int32_t>() + (kind == 1 ? npre0 : 0);
        T.nsuper = (int32_t)h_wu_prefix_kind[kind].size() - 1;
        T.nwu = h_wu_prefix_kind[kind].back();
        T.streaming = kind;
        CA_TRY(launch_tile(T, s));
    }
    CA_CUDA(cudaEventRecord(c.ev[5], s));
    CA_CUDA(cudaStreamSynchronize(s));
    float ms;
    CA_CUDA(cudaEventElapsedTime(&ms, c.ev[0], c.ev[1]));
    c.prof[0] = ms;
    CA_CUDA(cudaEventElapsedTime(&ms, c.ev[1], c.ev[2]));
    c.prof[1] = ms;
    CA_CUDA(cudaEventElapsedTime(&ms, c.ev[3], c.ev[4]));
    c.prof[3] = ms;
    CA_CUDA(cudaEventElapsedTime(&ms, c.ev[4], c.ev[5]));
    c.prof[4] = ms;
    c.prof[7] = (double)c.launches;  // every kernel of this call (pack, transpose, sort, qinfo, wave)
    c.prof[6] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call0).count();
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

}  // namespace ca

using namespace ca;

struct ca_labels {
    Labels L;
};
namespace ca {
Labels& labels_of(ca_labels_t* h) { return h->L; }
}

// =====================================================================================================
extern "C" {

int ca_version(void) { return 100; }
int ca_shard_owner(int32_t i, int32_t world) { return world > 0 && i >= 0 ? owner_of(i, world) : -1; }
const char* ca_last_error(void) { return g_err; }

int ca_init(const int* devices, int ndev) {
    int dev = (devices && ndev > 0) ? devices[0] : 0;
    return do_init(dev);
}

void ca_shutdown(void) {
    Ctx& c = ctx();
    if (!c.inited) return;
    cudaSetDevice(c.device);
    cudaDeviceSynchronize();
    Workspace& W = ws();
    W.perm.release(); W.cstart.release(); W.crow.release(); W.items.release(); W.parts.release(); W.w.release(); W.wu_prefix.release(); W.qinfo.release();
    W.h_cstart.release(); W.h_crow.release(); W.h_items.release(); W.h_small.release();
    stager().release();
    dev_cache_release();
    pool().shutdown();
    for (int i = 0; i < 8; i++)
        if (c.ev[i]) { cudaEventDestroy(c.ev[i]); c.ev[i] = nullptr; }
    if (c.own_stream) { cudaStreamDestroy(c.own_stream); c.own_stream = nullptr; }
    c.stream = nullptr;
    c.inited = false;
}

int ca_set_stream(void* cuda_stream) {
    CA_TRY(ensure_init());
    Ctx& c = ctx();
    c.stream = cuda_stream ? (cudaStream_t)cuda_stream : c.own_stream;
    return CA_OK;
}

int ca_get_profile(double* ms_out, int cap) {
    if (!ms_out || cap <= 0) return 0;
    int k = cap < 8 ? cap : 8;
    for (int i = 0; i < k; i++) ms_out[i] = ctx().prof[i];
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
int ca_labels_create(int64_t n, int32_t m, ca_labels_t** out) {
    if (!out || n < 0 || m < 1 || n > 0x7ffffff0LL) {
        set_error("ca_labels_create: bad arguments (n=%lld m=%d)", (long long)n, m);
        return CA_E_ARG;
    }
    CA_TRY(ensure_init());
    ca_labels* h = new (std::nothrow) ca_labels();
    if (!h) return CA_E_NOMEM;
    h->L.n = n;
    h->L.m = m;
    size_t bytes = (size_t)std::max<int64_t>(n, 1) * (size_t)m * 4;
    cudaError_t e = dev_alloc((void**)&h->L.d_raw, bytes);
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu bytes) for the label matrix failed: %s", bytes, cudaGetErrorString(e));
        delete h;
        return CA_E_NOMEM;
    }
    *out = h;
    return CA_OK;
}

int ca_labels_wrap(int64_t n, int32_t m, void* device_matrix, ca_labels_t** out) {
    if (!out || !device_matrix || n < 0 || m < 1 || n > 0x7ffffff0LL) {
        set_error("ca_labels_wrap: bad arguments (n=%lld m=%d)", (long long)n, m);
        return CA_E_ARG;
    }
    CA_TRY(ensure_init());
    ca_labels* h = new (std::nothrow) ca_labels();
    if (!h) return CA_E_NOMEM;
    h->L.n = n;
    h->L.m = m;
    h->L.d_raw = static_cast<int32_t*>(device_matrix);
    h->L.owns_raw = false;
    *out = h;
    return CA_OK;
}

int ca_labels_upload(ca_labels_t* h, const int32_t* labels) {
    if (!h || !labels) { set_error("ca_labels_upload: NULL argument"); return CA_E_ARG; }
    CA_TRY(ensure_init());
    cudaStream_t s = ctx().stream;
    const UploadSeg seg = {labels, h->L.d_raw, (size_t)h->L.n * h->L.m * 4};
    CA_TRY(upload_segments(&seg, 1, s));
    CA_CUDA(cudaStreamSynchronize(s));
    h->L.prepared = h->L.ranked = false;
    return CA_OK;
}

int ca_labels_upload_col(ca_labels_t* h, int32_t p, const int32_t* col) {
    if (!h || !col || p < 0 || p >= h->L.m) { set_error("ca_labels_upload_col: bad argument"); return CA_E_ARG; }
    CA_TRY(ensure_init());
    cudaStream_t s = ctx().stream;
    const UploadSeg seg = {col, h->L.d_raw + (size_t)p * h->L.n, (size_t)h->L.n * 4};
    CA_TRY(upload_segments(&seg, 1, s));
    h->L.prepared = h->L.ranked = false;
    return CA_OK;
}

int ca_labels_upload_cols(ca_labels_t* h, const int32_t* const* cols) {
    if (!h || !cols) { set_error("ca_labels_upload_cols: NULL argument"); return CA_E_ARG; }
    CA_TRY(ensure_init());
    cudaStream_t s = ctx().stream;
    std::vector<UploadSeg> segs((size_t)h->L.m);
    for (int32_t p = 0; p < h->L.m; p++) {
        if (!cols[p]) { set_error("NULL column pointer %d", p); return CA_E_ARG; }
        segs[p] = {cols[p], h->L.d_raw + (size_t)p * h->L.n, (size_t)h->L.n * 4};
    }
    CA_TRY(upload_segments(segs.data(), segs.size(), s));
    CA_CUDA(cudaStreamSynchronize(s));
    h->L.prepared = h->L.ranked = false;
    return CA_OK;
}

void* ca_labels_device_ptr(ca_labels_t* h) { return h ? (void*)h->L.d_raw : nullptr; }

int ca_labels_invalidate(ca_labels_t* h) {
    if (!h) return CA_E_ARG;
    h->L.prepared = h->L.ranked = false;
    return CA_OK;
}

int ca_labels_destroy(ca_labels_t* h) {
    if (!h) return CA_OK;
    if (ctx().inited) cudaSetDevice(ctx().device);
    Labels& L = h->L;
    if (L.d_raw && L.owns_raw) dev_free(L.d_raw, (size_t)std::max<int64_t>(L.n, 1) * (size_t)L.m * 4);
    L.d_min.release(); L.d_max.release(); L.d_cnt.release(); L.d_rank.release(); L.d_kcount.release();
    L.d_roff.release(); L.d_sizes.release(); L.d_lab16.release();
    L.d_spack.release(); L.d_kinfo.release(); L.d_tmpl.release();
    delete h;
    return CA_OK;
}

int ca_consistency_partial_dev(ca_labels_t* h, const double* weights_host, int32_t rank, int32_t world, double* ecc_dev,
                               double* sim_dev) {
    if (!h || world < 1 || rank < 0 || rank >= world) { set_error("ca_consistency_partial_dev: bad argument"); return CA_E_ARG; }
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

// ---- host-pointer batched entry points -----------------------------------------------------------------
static int host_allpairs(const int32_t* labels, const int32_t* const* cols, int64_t n, int32_t m, const double* weights,
                         double* ecc_out, double* sim_out) {
    if ((!labels && !cols) || n < 0 || m < 1) { set_error("bad arguments (NULL labels, n < 0 or m < 1)"); return CA_E_ARG; }
    CA_TRY(ensure_init());
    if (n == 0) {  // nothing to compare: ecc is empty, mean() of an empty ECS vector is NaN in R
        if (sim_out)
            for (int32_t i = 0; i < m; i++)
                for (int32_t j = 0; j < m; j++) sim_out[(size_t)i * m + j] = i == j ? 1.0 : 0.0 / 0.0;
        return CA_OK;
    }
    cudaStream_t s = ctx().stream;
    ca_labels_t* h = nullptr;
    CA_TRY(ca_labels_create(n, m, &h));
    int rc = CA_OK;
    static DevBuf d_ecc, d_sim;
    do {
        if (labels) {
            rc = ca_labels_upload(h, labels);
        } else {
            rc = ca_labels_upload_cols(h, cols);  // all columns in one staged upload
        }
        if (rc != CA_OK) break;
        if (ecc_out) { rc = d_ecc.ensure((size_t)std::max<int64_t>(n, 1) * 8); if (rc) break; }
        if (sim_out) { rc = d_sim.ensure((size_t)m * m * 8); if (rc) break; }
        cudaError_t e = cudaSuccess;
        if (ecc_out) e = cudaMemsetAsync(d_ecc.p, 0, (size_t)std::max<int64_t>(n, 1) * 8, s);
        if (e == cudaSuccess && sim_out) e = cudaMemsetAsync(d_sim.p, 0, (size_t)m * m * 8, s);
        if (e != cudaSuccess) { set_error("cudaMemsetAsync failed: %s", cudaGetErrorString(e)); rc = CA_E_CUDA; break; }
        if (m >= 2) {
            rc = run_allpairs(h->L, weights, 0, 1, ecc_out ? d_ecc.as<double>() : nullptr, sim_out ? d_sim.as<double>() : nullptr, -1, 0);
            if (rc) break;
        }
        rc = ca_consistency_finish_dev(n, m, weights, ecc_out ? d_ecc.as<double>() : nullptr, sim_out ? d_sim.as<double>() : nullptr);
        if (rc) break;
        if (ecc_out && n > 0) e = cudaMemcpyAsync(ecc_out, d_ecc.p, (size_t)n * 8, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess && sim_out) e = cudaMemcpyAsync(sim_out, d_sim.p, (size_t)m * m * 8, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { set_error("device-to-host copy failed: %s", cudaGetErrorString(e)); rc = CA_E_CUDA; }
    } while (0);
    ca_labels_destroy(h);
    return rc;
}

// ---- resident label sets: the same job without the upload ----------------------------------------------------
int ca_labels_select(ca_labels_t* src, const int32_t* idx, int32_t k, ca_labels_t** out) {
    if (!src || !idx || !out || k < 1) { set_error("ca_labels_select: bad argument"); return CA_E_ARG; }
    for (int32_t t = 0; t < k; t++)
        if (idx[t] < 0 || idx[t] >= src->L.m) { set_error("ca_labels_select: index %d out of range [0, %d)", idx[t], src->L.m); return CA_E_ARG; }
    CA_TRY(ensure_init());
    ca_labels_t* h = nullptr;
    CA_TRY(ca_labels_create(src->L.n, k, &h));
    cudaStream_t s = ctx().stream;
    const size_t row = (size_t)src->L.n * 4;
    for (int32_t t = 0; t < k && row > 0;) {  // runs of consecutive indices go out as one copy
        int32_t e = t + 1;
        while (e < k && idx[e] == idx[e - 1] + 1) e++;
        cudaError_t err = cudaMemcpyAsync(h->L.d_raw + (size_t)t * src->L.n, src->L.d_raw + (size_t)idx[t] * src->L.n, row * (size_t)(e - t),
                                          cudaMemcpyDeviceToDevice, s);
        if (err != cudaSuccess) {
            set_error("device-to-device copy failed: %s", cudaGetErrorString(err));
            ca_labels_destroy(h);
            return CA_E_CUDA;
        }
        t = e;
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
    const int64_t n = use->L.n;
    const int32_t m = use->L.m;
    int rc = CA_OK;
    static DevBuf d_ecc, d_sim;
    cudaStream_t s = ctx().stream;
    do {
        if (ecc_out && m < 2) { set_error("At least two clusterings are required to calculate the consistency."); rc = CA_E_ARG; break; }
        if (n == 0) {
            if (sim_out)
                for (int32_t i = 0; i < m; i++)
                    for (int32_t j = 0; j < m; j++) sim_out[(size_t)i * m + j] = i == j ? 1.0 : 0.0 / 0.0;
            break;
        }
        if (ecc_out) { rc = d_ecc.ensure((size_t)n * 8); if (rc) break; }
        if (sim_out) { rc = d_sim.ensure((size_t)m * m * 8); if (rc) break; }
        cudaError_t e = cudaSuccess;
        if (ecc_out) e = cudaMemsetAsync(d_ecc.p, 0, (size_t)n * 8, s);
        if (e == cudaSuccess && sim_out) e = cudaMemsetAsync(d_sim.p, 0, (size_t)m * m * 8, s);
        if (e != cudaSuccess) { set_error("cudaMemsetAsync failed: %s", cudaGetErrorString(e)); rc = CA_E_CUDA; break; }
        if (m >= 2) {
            rc = run_allpairs(use->L, weights, 0, 1, ecc_out ? d_ecc.as<double>() : nullptr, sim_out ? d_sim.as<double>() : nullptr, -1, 0);
            if (rc) break;
        }
        rc = ca_consistency_finish_dev(n, m, weights, ecc_out ? d_ecc.as<double>() : nullptr, sim_out ? d_sim.as<double>() : nullptr);
        if (rc) break;
        if (ecc_out) e = cudaMemcpyAsync(ecc_out, d_ecc.p, (size_t)n * 8, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess && sim_out) e = cudaMemcpyAsync(sim_out, d_sim.p, (size_t)m * m * 8, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { set_error("device-to-host copy failed: %s", cudaGetErrorString(e)); rc = CA_E_CUDA; }
    } while (0);
    if (sub) ca_labels_destroy(sub);
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
    cudaStream_t s = ctx().stream;
    ca_labels_t* h = nullptr;
    CA_TRY(ca_labels_create(n, m + 1, &h));  // partitions 0..m-1 = the list, partition m = the reference
    int rc = CA_OK;
    static DevBuf d_out;
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
        if (n > 0) e = cudaMemcpyAsync(agr_out, d_out.p, (size_t)n * 8, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { set_error("device-to-host copy failed: %s", cudaGetErrorString(e)); rc = CA_E_CUDA; }
    } while (0);
    ca_labels_destroy(h);
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
    DevBuf a, b, table, sa, sb, err, ecs, partial, out;
};
static PairBufs& pb() {
    static PairBufs p;
    return p;
}

static int pair_upload(const int32_t* a, const int32_t* b, int64_t n, cudaStream_t s) {
    PairBufs& B = pb();
    CA_TRY(B.a.ensure((size_t)std::max<int64_t>(n, 1) * 4));
    CA_TRY(B.b.ensure((size_t)std::max<int64_t>(n, 1) * 4));
    if (n > 0) {
        const UploadSeg segs[2] = {{a, B.a.p, (size_t)n * 4}, {b, B.b.p, (size_t)n * 4}};
        CA_TRY(upload_segments(segs, 2, s));
    }
    return CA_OK;
}

static int pair_tables(int64_t n, int32_t min_a, int32_t min_b, int32_t ka, int32_t kb, cudaStream_t s) {
    PairBufs& B = pb();
    CA_TRY(B.table.ensure((size_t)ka * kb * 4));
    CA_TRY(B.sa.ensure((size_t)ka * 4));
    CA_TRY(B.sb.ensure((size_t)kb * 4));
    CA_TRY(B.err.ensure(4));
    return pair_contingency(B.a.as<int32_t>(), B.b.as<int32_t>(), n, min_a, min_b, ka, kb, B.table.as<uint32_t>(),
                            B.sa.as<uint32_t>(), B.sb.as<uint32_t>(), B.err.as<int32_t>(), s);
}

static int check_dims(int32_t min_a, int32_t max_a, int32_t min_b, int32_t max_b, int32_t* ka, int32_t* kb) {
    int64_t ra = (int64_t)max_a - min_a + 1, rb = (int64_t)max_b - min_b + 1;
    if (ra * rb > (1ll << 31)) {
        set_error("contingency table %lld x %lld is too large", (long long)ra, (long long)rb);
        return CA_E_UNSUPPORTED;
    }
    *ka = (int32_t)ra;
    *kb = (int32_t)rb;
    return CA_OK;
}

int ca_contingency(const int32_t* a, const int32_t* b, int64_t n, int32_t min_a, int32_t min_b, int32_t ka, int32_t kb,
                   int32_t* table_out, int32_t* sizes_a, int32_t* sizes_b) {
    if (!a || !b || !table_out || n < 0 || ka < 1 || kb < 1) { set_error("ca_contingency: bad argument"); return CA_E_ARG; }
    if ((int64_t)ka * kb > (1ll << 31)) { set_error("contingency table %d x %d is too large", ka, kb); return CA_E_UNSUPPORTED; }
    CA_TRY(ensure_init());
    cudaStream_t s = ctx().stream;
    PairBufs& B = pb();
    CA_TRY(pair_upload(a, b, n, s));
    CA_TRY(pair_tables(n, min_a, min_b, ka, kb, s));
    int32_t err = 0;
    CA_CUDA(cudaMemcpyAsync(&err, B.err.p, 4, cudaMemcpyDeviceToHost, s));
    CA_CUDA(cudaMemcpyAsync(table_out, B.table.p, (size_t)ka * kb * 4, cudaMemcpyDeviceToHost, s));
    if (sizes_a) CA_CUDA(cudaMemcpyAsync(sizes_a, B.sa.p, (size_t)ka * 4, cudaMemcpyDeviceToHost, s));
    if (sizes_b) CA_CUDA(cudaMemcpyAsync(sizes_b, B.sb.p, (size_t)kb * 4, cudaMemcpyDeviceToHost, s));
    CA_CUDA(cudaStreamSynchronize(s));
    if (err) { set_error("ca_contingency: a label lies outside [min, min+k)"); return CA_E_RANGE; }
    return CA_OK;
}

int ca_ecs_pair(const int32_t* a, const int32_t* b, int64_t n, double* ecs_out, double* mean_out) {
    if (!a || !b || (!ecs_out && !mean_out) || n < 0) { set_error("ca_ecs_pair: bad argument"); return CA_E_ARG; }
    if (n == 0) { if (mean_out) *mean_out = 0.0 / 0.0; return CA_OK; }
    CA_TRY(ensure_init());
    cudaStream_t s = ctx().stream;
    PairBufs& B = pb();
    int32_t min_a, max_a, min_b, max_b, ka, kb;
    CA_TRY(ca_label_range(a, n, &min_a, &max_a));
    CA_TRY(ca_label_range(b, n, &min_b, &max_b));
    CA_TRY(check_dims(min_a, max_a, min_b, max_b, &ka, &kb));
    CA_TRY(pair_upload(a, b, n, s));
    CA_TRY(pair_tables(n, min_a, min_b, ka, kb, s));
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
    PairBufs& B = pb();
    int32_t min_a, max_a, min_b, max_b, ka, kb;
    CA_TRY(ca_label_range(a, n, &min_a, &max_a));
    CA_TRY(ca_label_range(b, n, &min_b, &max_b));
    CA_TRY(check_dims(min_a, max_a, min_b, max_b, &ka, &kb));
    CA_TRY(pair_upload(a, b, n, s));
    CA_TRY(pair_tables(n, min_a, min_b, ka, kb, s));
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
