// ca_dedup.cu -- identical-partition merge on the device.
//
// Replaces merge_identical_partitions / are_identical_memberships (R/ECS.R:1021-1054, 984-1016), which
// build up to M*U dense contingency tables through myContTable before any ECS is computed.
//
// Two partitions are the same partition (up to label names) iff their CANONICAL labelings agree, where
// the canonical label of a cell is the smallest cell index in its cluster.  Per partition:
//   first_cell_kernel   first[p][k] = min{ n : label_p(n) == k }            (atomicMin, O(N))
//   signature_kernel    sig[p] = sum_n mix(n, first[p][label_p(n)])  (2 x 64 bit, order-free, O(N))
// The host then replays the reference's greedy first-match loop on the signatures, confirming every
// signature match exactly with compare_kernel (element-wise equality of the two canonical labelings),
// and applies the reference's own decision rule, quirks included:
//   identical(j, i) = range_j == range_i  &&  ( range_j == n                        [R/ECS.R:994,999]
//                       || (no unused label inside either range && same partition) ) [R/ECS.R:1003-1015]
// where range = max - min + 1 (calculate_ecs.cpp:8-10: unused labels inside the range are table rows).
#include <algorithm>

#include <string>

#include "ca_internal.h"

namespace ca {

static constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ uint64_t mix64(uint64_t x) {  // splitmix64 finaliser
    x ^= x >> 30;
    x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27;
    x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}

__device__ __forceinline__ uint32_t compact_label(const int32_t* x, int64_t i, int mn, const uint32_t* rk) {
    uint32_t v = (uint32_t)(__ldg(x + i) - mn);
    return rk ? __ldg(rk + v) : v;
}

__global__ void __launch_bounds__(256) first_cell_kernel(const int32_t* __restrict__ raw, int64_t n, const int32_t* __restrict__ d_min,
                                                         const uint32_t* __restrict__ d_rank, const int64_t* __restrict__ d_roff,
                                                         int32_t kp, int32_t* __restrict__ first) {
    const int p = blockIdx.y;
    const int32_t* x = raw + (int64_t)p * n;
    const int mn = d_min[p];
    const uint32_t* rk = d_rank ? d_rank + d_roff[p] : nullptr;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t k = compact_label(x, i, mn, rk);
        int32_t* f = first + (int64_t)p * kp + k;
        if (*f > (int32_t)i) atomicMin(f, (int32_t)i);  // racy pre-check only skips hopeless atomics
    }
}

__global__ void __launch_bounds__(256) signature_kernel(const int32_t* __restrict__ raw, int64_t n, const int32_t* __restrict__ d_min,
                                                        const uint32_t* __restrict__ d_rank, const int64_t* __restrict__ d_roff,
                                                        int32_t kp, const int32_t* __restrict__ first,
                                                        unsigned long long* __restrict__ sig) {
    const int p = blockIdx.y;
    const int32_t* x = raw + (int64_t)p * n;
    const int mn = d_min[p];
    const uint32_t* rk = d_rank ? d_rank + d_roff[p] : nullptr;
    uint64_t lo = 0, hi = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t k = compact_label(x, i, mn, rk);
        uint64_t c = (uint64_t)(uint32_t)first[(int64_t)p * kp + k];
        uint64_t key = ((uint64_t)i << 32) | c;
        lo += mix64(key);
        hi += mix64(key ^ 0x9e3779b97f4a7c15ull);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo += __shfl_xor_sync(FULL, lo, o);
        hi += __shfl_xor_sync(FULL, hi, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(sig + 2 * p, (unsigned long long)lo);  // wrap-around integer sums: order-independent, deterministic
        atomicAdd(sig + 2 * p + 1, (unsigned long long)hi);
    }
}

__global__ void __launch_bounds__(256) compare_kernel(const int32_t* __restrict__ raw, int64_t n, const int32_t* __restrict__ d_min,
                                                      const uint32_t* __restrict__ d_rank, const int64_t* __restrict__ d_roff,
                                                      int32_t kp, const int32_t* __restrict__ first, int32_t p, int32_t q,
                                                      int32_t* __restrict__ differ) {
    const int32_t* xp = raw + (int64_t)p * n;
    const int32_t* xq = raw + (int64_t)q * n;
    const int mp = d_min[p], mq = d_min[q];
    const uint32_t* rp = d_rank ? d_rank + d_roff[p] : nullptr;
    const uint32_t* rq = d_rank ? d_rank + d_roff[q] : nullptr;
    bool bad = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int32_t cp = first[(int64_t)p * kp + compact_label(xp, i, mp, rp)];
        int32_t cq = first[(int64_t)q * kp + compact_label(xq, i, mq, rq)];
        bad |= (cp != cq);
    }
    if (__any_sync(FULL, bad) && (threadIdx.x & 31) == 0) *differ = 1;
}


int dedup_signatures(Labels& L, std::vector<uint64_t>& sig_lo, std::vector<uint64_t>& sig_hi, DevBuf& d_first, cudaStream_t s) {
    const int32_t m = L.m;
    const int64_t n = L.n;
    CA_TRY(prepare_ranks(L));
    CA_TRY(d_first.ensure((size_t)m * L.kp * 4));
    DevBuf& g_sig = scratch(SCR_SIG);
    CA_TRY(g_sig.ensure((size_t)m * 16));
    CA_TRY(launch_fill_i32(d_first.as<int32_t>(), (int64_t)m * L.kp, INT32_MAX, s));
    CA_CUDA(cudaMemsetAsync(g_sig.p, 0, (size_t)m * 16, s));
    int64_t nb = (n + 256 * 8 - 1) / (256 * 8);
    nb = std::max<int64_t>(1, std::min<int64_t>(nb, 2048));
    dim3 grid((unsigned)nb, (unsigned)m);
    const uint32_t* rk = L.identity ? nullptr : L.d_rank.as<uint32_t>();
    first_cell_kernel<<<grid, 256, 0, s>>>(L.d_raw, n, L.d_min.as<int32_t>(), rk, L.d_roff.as<int64_t>(), L.kp, d_first.as<int32_t>());
    signature_kernel<<<grid, 256, 0, s>>>(L.d_raw, n, L.d_min.as<int32_t>(), rk, L.d_roff.as<int64_t>(), L.kp, d_first.as<int32_t>(),
                                          g_sig.as<unsigned long long>());
    CA_CUDA(cudaGetLastError()); ca::ctx().launches += 2;
    std::vector<uint64_t> h((size_t)m * 2);
    CA_CUDA(cudaMemcpyAsync(h.data(), g_sig.p, (size_t)m * 16, cudaMemcpyDeviceToHost, s));
    CA_CUDA(cudaStreamSynchronize(s));
    sig_lo.resize(m);
    sig_hi.resize(m);
    for (int32_t p = 0; p < m; p++) {
        sig_lo[p] = h[2 * p];
        sig_hi[p] = h[2 * p + 1];
    }
    return CA_OK;
}

int dedup_same_partition(Labels& L, const DevBuf& d_first, int32_t p, int32_t q, bool* same, cudaStream_t s) {
    DevBuf& g_flag = scratch(SCR_FLAG);
    CA_TRY(g_flag.ensure(4));
    CA_CUDA(cudaMemsetAsync(g_flag.p, 0, 4, s));
    int64_t nb = (L.n + 256 * 8 - 1) / (256 * 8);
    nb = std::max<int64_t>(1, std::min<int64_t>(nb, 1184));
    compare_kernel<<<(unsigned)nb, 256, 0, s>>>(L.d_raw, L.n, L.d_min.as<int32_t>(), L.identity ? nullptr : L.d_rank.as<uint32_t>(),
                                                L.d_roff.as<int64_t>(), L.kp, static_cast<const int32_t*>(d_first.p), p, q,
                                                g_flag.as<int32_t>());
    CA_CUDA(cudaGetLastError()); ca::ctx().launches++;
    int32_t differ = 0;
    CA_CUDA(cudaMemcpyAsync(&differ, g_flag.p, 4, cudaMemcpyDeviceToHost, s));
    CA_CUDA(cudaStreamSynchronize(s));
    *same = differ == 0;
    return CA_OK;
}

}  // namespace ca

using namespace ca;

extern "C" int ca_labels_create_on_primary(int64_t n, int32_t m, ca_labels_t** out);
extern "C" int ca_labels_upload(ca_labels_t* h, const int32_t* labels);
extern "C" int ca_labels_upload_cols(ca_labels_t* h, const int32_t* const* cols);
extern "C" int ca_labels_destroy(ca_labels_t* h);
extern "C" int ca_labels_is_sharded(ca_labels_t* h);
namespace ca {
Labels& labels_of(ca_labels_t* h);
}

// greedy first-match dedup of a label set that is already on the device: exactly the loop of R/ECS.R:1030-1050, with the
// contingency-table test of are_identical_memberships (R/ECS.R:984-1016) replaced by canonical labelling + signature + exact compare
static int merge_identical_on(Labels& L, const double* freq_in, int32_t* keep_out, double* freq_out, int32_t* u_out) {
    cudaStream_t s = ctx().stream;
    const int32_t m = L.m;
    const int64_t n = L.n;
    DevBuf& d_first = scratch(SCR_FIRST);
    std::vector<uint64_t> lo, hi;
    int rc = dedup_signatures(L, lo, hi, d_first, s);
    int32_t u = 0;
    if (rc == CA_OK) {
        std::vector<int64_t> range(m);
        std::vector<char> gapless(m);
        for (int32_t p = 0; p < m; p++) {
            range[p] = L.h_roff[p + 1] - L.h_roff[p];
            gapless[p] = (int64_t)L.h_k[p] == range[p];
        }
        for (int32_t i = 0; i < m && rc == CA_OK; i++) {
            int32_t assigned = -1;
            for (int32_t t = 0; t < u && rc == CA_OK; t++) {
                const int32_t j = keep_out[t];
                if (range[j] != range[i]) continue;                   // R/ECS.R:994
                bool ident = range[j] == n;                           // R/ECS.R:999
                if (!ident && gapless[j] && gapless[i] && lo[j] == lo[i] && hi[j] == hi[i])
                    rc = dedup_same_partition(L, d_first, j, i, &ident, s);  // exact confirmation
                if (ident) {
                    assigned = t;
                    break;
                }
            }
            if (rc != CA_OK) break;
            const double f = freq_in ? freq_in[i] : 1.0;
            if (assigned >= 0) {
                freq_out[assigned] += f;
            } else {
                keep_out[u] = i;
                freq_out[u] = f;
                u++;
            }
        }
    }
    if (rc == CA_OK) *u_out = u;
    return rc;
}

static int merge_identical_impl(const int32_t* labels, const int32_t* const* cols, int64_t n, int32_t m, const double* freq_in,
                               int32_t* keep_out, double* freq_out, int32_t* u_out) {
    if ((!labels && !cols) || !keep_out || !freq_out || !u_out || n < 1 || m < 1) {
        set_error("ca_merge_identical: bad argument");
        return CA_E_ARG;
    }
    CA_TRY(ensure_init());
    ca_labels_t* h = nullptr;
    CA_TRY(ca_labels_create_on_primary(n, m, &h));  // the pre-pass is O(M N) streaming work: one device
    int rc = labels ? ca_labels_upload(h, labels) : ca_labels_upload_cols(h, cols);
    if (rc == CA_OK) rc = merge_identical_on(labels_of(h), freq_in, keep_out, freq_out, u_out);
    std::string msg = ca_last_error();
    ca_labels_destroy(h);
    if (rc != CA_OK) set_error("%s", msg.c_str());
    return rc;
}

extern "C" int ca_merge_identical(const int32_t* labels, int64_t n, int32_t m, const double* freq_in, int32_t* keep_out,
                                  double* freq_out, int32_t* u_out) {
    return merge_identical_impl(labels, nullptr, n, m, freq_in, keep_out, freq_out, u_out);
}
extern "C" int ca_merge_identical_cols(const int32_t* const* cols, int64_t n, int32_t m, const double* freq_in, int32_t* keep_out,
                                       double* freq_out, int32_t* u_out) {
    return merge_identical_impl(nullptr, cols, n, m, freq_in, keep_out, freq_out, u_out);
}
// the same on a resident single-device label set: merge_partitions / element_consistency upload their partitions ONCE, dedup
// them here and run the consistency of the kept ones (ca_consistency_resident with their indices) on the same device memory
extern "C" int ca_merge_identical_resident(ca_labels_t* h, const double* freq_in, int32_t* keep_out, double* freq_out, int32_t* u_out) {
    if (!h || !keep_out || !freq_out || !u_out) {
        set_error("ca_merge_identical_resident: NULL argument");
        return CA_E_ARG;
    }
    if (ca_labels_is_sharded(h)) {
        set_error("ca_merge_identical_resident needs a single-device label set (ca_labels_create_on_primary)");
        return CA_E_UNSUPPORTED;
    }
    CA_TRY(ensure_init());
    if (labels_of(h).n < 1) {
        set_error("ca_merge_identical_resident: empty label set");
        return CA_E_ARG;
    }
    return merge_identical_on(labels_of(h), freq_in, keep_out, freq_out, u_out);
}
