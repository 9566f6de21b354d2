// ca_pack.cu -- label packing for the all-pairs ECS engine (sm_100a).
//
// Replaces the R-side label handling of the reference (labels stay in R vectors and are re-scanned
// for min/max on EVERY pair: src/calculate_ecs.cpp:8,23; R/ECS.R:989) by one device pass per call:
//   minmax   -> per-partition min / max                        (calculate_ecs.cpp:8,23 semantics)
//   count    -> per-partition label histogram over [min, max]  (= cluster sizes, calculate_ecs.cpp:31)
//   rank     -> dense rank of every used label (unused labels inside the range only produce all-zero
//               table rows in the reference and never influence ECS, so the engine drops them)
//   sizes    -> compact cluster sizes S[p][k]
//   relabel_transpose -> uint16 label matrix lab16[mpad/16][n + 1][16]: 32 contiguous bytes per (cell, 16 partitions)
//   sort     -> per-partition counting sort: perm[p] lists the cells grouped by cluster, every
//               cluster at the padded start position chosen by the host planner (ca_plan.cpp)
// All kernels are HBM-bound streaming passes over the M x N int32 input (coalesced loads), a few
// percent of the wave kernel's time.
#include <algorithm>

#include <stdlib.h>

#include "ca_internal.h"

namespace ca {

static constexpr unsigned FULL = 0xffffffffu;

// ------------------------------------------------------------------------------------------------
// per-partition min / max
// ------------------------------------------------------------------------------------------------
// A partition's label vector starts at raw + p n: 16-byte aligned only if p n is a multiple of 4.  vec_bounds() cuts it
// into an unaligned head (< 4 elements), a body of int4 and a tail, so that the streaming kernels below read 16 bytes
// per thread and load whatever n is.
__device__ __forceinline__ void vec_bounds(const int32_t* x, int64_t n, int64_t& head, int64_t& nvec) {
    head = (int64_t)((4 - ((reinterpret_cast<uintptr_t>(x) >> 2) & 3)) & 3);
    if (head > n) head = n;
    nvec = (n - head) >> 2;
}

__global__ void __launch_bounds__(256) minmax_kernel(const int32_t* __restrict__ raw, int64_t n, int32_t* __restrict__ d_min,
                                                     int32_t* __restrict__ d_max) {
    const int p = blockIdx.y;
    const int32_t* x = raw + (int64_t)p * n;
    int lo = INT32_MAX, hi = INT32_MIN;
    int64_t head, nvec;
    vec_bounds(x, n, head, nvec);
    const int4* xv = reinterpret_cast<const int4*>(x + head);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < nvec; i += 2 * stride) {  // two independent 16-byte loads in flight per thread
        const int4 a = __ldg(xv + i), b = __ldg(xv + i + stride);
        lo = min(min(min(lo, a.x), min(a.y, a.z)), min(min(a.w, b.x), min(min(b.y, b.z), b.w)));
        hi = max(max(max(hi, a.x), max(a.y, a.z)), max(max(a.w, b.x), max(max(b.y, b.z), b.w)));
    }
    if (i < nvec) {
        const int4 a = __ldg(xv + i);
        lo = min(min(lo, a.x), min(min(a.y, a.z), a.w));
        hi = max(max(hi, a.x), max(max(a.y, a.z), a.w));
    }
    if (blockIdx.x == 0 && threadIdx.x < 8) {  // head and tail: at most 3 elements each
        const int64_t t = threadIdx.x < 4 ? (int64_t)threadIdx.x : head + 4 * nvec + (threadIdx.x - 4);
        const bool ok = threadIdx.x < 4 ? t < head : t < n;
        if (ok) {
            const int v = __ldg(x + t);
            lo = min(lo, v);
            hi = max(hi, v);
        }
    }
    lo = __reduce_min_sync(FULL, lo);
    hi = __reduce_max_sync(FULL, hi);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(d_min + p, lo);
        atomicMax(d_max + p, hi);
    }
}

__global__ void fill_i32_kernel(int32_t* x, int64_t n, int32_t v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = v;
}

int launch_fill_i32(int32_t* x, int64_t n, int32_t v, cudaStream_t s) {
    if (n <= 0) return CA_OK;
    fill_i32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x, n, v);
    CA_CUDA(cudaGetLastError()); ca::ctx().launches++;
    return CA_OK;
}

int launch_minmax(const int32_t* raw, int64_t n, int32_t m, int32_t* d_min, int32_t* d_max, cudaStream_t s) {
    fill_i32_kernel<<<(m + 255) / 256, 256, 0, s>>>(d_min, m, INT32_MAX);
    fill_i32_kernel<<<(m + 255) / 256, 256, 0, s>>>(d_max, m, INT32_MIN);
    int64_t nb = (n + 256 * 32 - 1) / (256 * 32);  // 32 elements (two 16-byte loads, four trips) per thread
    if (nb < 1) nb = 1;
    if (nb > 4096) nb = 4096;
    dim3 grid((unsigned)nb, (unsigned)m);
    minmax_kernel<<<grid, 256, 0, s>>>(raw, n, d_min, d_max);
    CA_CUDA(cudaGetLastError()); ca::ctx().launches += 3;
    return CA_OK;
}

// ------------------------------------------------------------------------------------------------
// per-partition label histogram over [min, max]; shared-memory privatised when the range fits
// ------------------------------------------------------------------------------------------------
template <bool SMEM>
__global__ void __launch_bounds__(512) count_kernel(const int32_t* __restrict__ raw, int64_t n, const int32_t* __restrict__ d_min,
                                                    const int64_t* __restrict__ d_roff, int32_t range_cap,
                                                    uint32_t* __restrict__ d_cnt) {
    extern __shared__ uint32_t sh[];
    const int p = blockIdx.y;
    const int32_t* x = raw + (int64_t)p * n;
    const int mn = d_min[p];
    uint32_t* g = d_cnt + d_roff[p];
    const int range = (int)(d_roff[p + 1] - d_roff[p]);
    const bool use_sh = SMEM && range <= range_cap;
    if (use_sh) {
        for (int i = threadIdx.x; i < range; i += blockDim.x) sh[i] = 0;
        __syncthreads();
    }
    int64_t head, nvec;
    vec_bounds(x, n, head, nvec);
    const int4* xv = reinterpret_cast<const int4*>(x + head);
    auto body = [&](auto add) {  // instantiated once per address space, so that the shared-memory case stays ATOMS
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
            const int4 a = __ldg(xv + i);
            add(a.x - mn);
            add(a.y - mn);
            add(a.z - mn);
            add(a.w - mn);
        }
        if (blockIdx.x == 0 && threadIdx.x < 8) {  // head and tail: at most 3 elements each
            const int64_t t = threadIdx.x < 4 ? (int64_t)threadIdx.x : head + 4 * nvec + (threadIdx.x - 4);
            const bool ok = threadIdx.x < 4 ? t < head : t < n;
            if (ok) add(__ldg(x + t) - mn);
        }
    };
    if (use_sh)
        body([&](int v) { atomicAdd(sh + v, 1u); });
    else
        body([&](int v) { atomicAdd(g + v, 1u); });
    if (use_sh) {
        __syncthreads();
        for (int i = threadIdx.x; i < range; i += blockDim.x) {
            uint32_t c = sh[i];
            if (c) atomicAdd(g + i, c);
        }
    }
}

int launch_count(const int32_t* raw, int64_t n, int32_t m, const int32_t* d_min, const int64_t* d_roff,
                 const int32_t* /*h_range*/, int32_t max_range, uint32_t* d_cnt, cudaStream_t s) {
    int64_t nb = (n + 512 * 64 - 1) / (512 * 64);  // 64 elements (sixteen 16-byte loads) per thread
    if (nb < 1) nb = 1;
    if (nb > 1024) nb = 1024;
    dim3 grid((unsigned)nb, (unsigned)m);
    const int32_t cap = 12288;  // 48 KB of counters: no opt-in needed
    if (max_range <= cap) {
        count_kernel<true><<<grid, 512, (size_t)max_range * 4, s>>>(raw, n, d_min, d_roff, cap, d_cnt);
    } else {
        // mixed: partitions whose range fits use shared memory, the rest go straight to global atomics
        count_kernel<true><<<grid, 512, (size_t)cap * 4, s>>>(raw, n, d_min, d_roff, cap, d_cnt);
    }
    CA_CUDA(cudaGetLastError()); ca::ctx().launches++;
    return CA_OK;
}

// ------------------------------------------------------------------------------------------------
// dense rank of used labels: rank[v] = #{u < v : cnt[u] > 0}; one block per partition
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) rank_kernel(const uint32_t* __restrict__ d_cnt, const int64_t* __restrict__ d_roff,
                                                    uint32_t* __restrict__ d_rank, int32_t* __restrict__ d_kcount) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry;
    const int p = blockIdx.x;
    const int64_t off = d_roff[p];
    const int range = (int)(d_roff[p + 1] - off);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < range; base += blockDim.x) {
        int v = base + threadIdx.x;
        uint32_t f = (v < range && d_cnt[off + v] != 0) ? 1u : 0u;
        unsigned bal = __ballot_sync(FULL, f);
        uint32_t in_warp = __popc(bal & ((1u << lane) - 1));
        if (lane == 0) warp_tot[wid] = __popc(bal);
        __syncthreads();
        uint32_t before = carry;
        for (int w = 0; w < wid; w++) before += warp_tot[w];
        if (v < range) d_rank[off + v] = before + in_warp;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += warp_tot[w];
            carry += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) d_kcount[p] = (int32_t)carry;
}

int launch_rank(const uint32_t* d_cnt, const int64_t* d_roff, const int32_t*, const int32_t*, int32_t m, uint32_t* d_rank,
                int32_t* d_kcount, cudaStream_t s) {
    rank_kernel<<<m, 1024, 0, s>>>(d_cnt, d_roff, d_rank, d_kcount);
    CA_CUDA(cudaGetLastError()); ca::ctx().launches++;
    return CA_OK;
}

__global__ void sizes_kernel(const uint32_t* __restrict__ d_cnt, const uint32_t* __restrict__ d_rank,
                             const int64_t* __restrict__ d_roff, int32_t kp, int32_t* __restrict__ d_sizes) {
    const int p = blockIdx.y;
    const int64_t off = d_roff[p];
    const int range = (int)(d_roff[p + 1] - off);
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < range; v += gridDim.x * blockDim.x) {
        uint32_t c = d_cnt[off + v];
        if (c) d_sizes[(int64_t)p * kp + d_rank[off + v]] = (int32_t)c;
    }
}

int launch_sizes(const uint32_t* d_cnt, const uint32_t* d_rank, const int64_t* d_roff, const int32_t*, const int32_t*,
                 int32_t m, int32_t max_range, int32_t kp, int32_t* d_sizes, cudaStream_t s) {
    int nb = (max_range + 255) / 256;
    if (nb > 256) nb = 256;
    if (nb < 1) nb = 1;
    dim3 grid((unsigned)nb, (unsigned)m);
    sizes_kernel<<<grid, 256, 0, s>>>(d_cnt, d_rank, d_roff, kp, d_sizes);
    CA_CUDA(cudaGetLastError()); ca::ctx().launches++;
    return CA_OK;
}

// initial table entries of the wave kernel, spack[p][k] = min(size, 65535) << 16 (0xFFFF = look the size up),
// and kinfo[p] = K_p | (1 << 30 if partition p has a cluster with >= 65535 cells); one block per partition
__global__ void __launch_bounds__(256) spack_kernel(const int32_t* __restrict__ sizes, const int32_t* __restrict__ kcount, int32_t kp,
                                                    uint32_t* __restrict__ spack, int32_t* __restrict__ kinfo) {
    const int p = blockIdx.x;
    int big = 0;
    for (int k = threadIdx.x; k < kp; k += blockDim.x) {
        const int32_t v = sizes[(int64_t)p * kp + k];
        big |= v >= 65535;
        // beyond K_p (the pad column and the row's alignment slack): size 1, so that a gather of it never divides by 0
        spack[(int64_t)p * kp + k] = k < kcount[p] ? (uint32_t)min(v, 65535) << 16 : 1u << 16;
    }
    big = __syncthreads_or(big);
    if (threadIdx.x == 0) kinfo[p] = kcount[p] | (big ? (1 << 30) : 0);
}
int launch_spack(const int32_t* d_sizes, const int32_t* d_kcount, int32_t m, int32_t kp, uint32_t* d_spack, int32_t* d_kinfo, cudaStream_t s) {
    spack_kernel<<<(unsigned)m, 256, 0, s>>>(d_sizes, d_kcount, kp, d_spack, d_kinfo);
    CA_CUDA(cudaGetLastError()); ca::ctx().launches++;
    return CA_OK;
}

// Sharded label sets: a device's rows of the cluster-size array (and its cluster counts) go into the whole-set arrays of EVERY device
// (peer stores).  Local partition k is global partition span_of_block[k / 32] * 32 + k % 32.
__global__ void __launch_bounds__(256) scatter_sizes_kernel(const int32_t* __restrict__ sizes_loc, const int32_t* __restrict__ kcount_loc,
                                                            const int32_t* __restrict__ span_of_block, int32_t kp, const SizeDst dst) {
    const int k = blockIdx.x;
    const int gid = span_of_block[k >> 5] * 32 + (k & 31);
    const int32_t* src = sizes_loc + (int64_t)k * kp;
    for (int d = 0; d < dst.n; d++) {
        int32_t* o = dst.sizes[d] + (int64_t)gid * kp;
        for (int w = threadIdx.x; w < kp; w += blockDim.x) o[w] = src[w];
        if (threadIdx.x == 0) dst.kcount[d][gid] = kcount_loc[k];
    }
}
int launch_scatter_sizes(const int32_t* sizes_loc, const int32_t* kcount_loc, const int32_t* d_span_of_block, int32_t m_loc, int32_t kp, const SizeDst& dst,
                         cudaStream_t s) {
    if (m_loc <= 0) return CA_OK;
    scatter_sizes_kernel<<<(unsigned)m_loc, 256, 0, s>>>(sizes_loc, kcount_loc, d_span_of_block, kp, dst);
    CA_CUDA(cudaGetLastError()); ca::ctx().launches++;
    return CA_OK;
}

// Table templates for the wave kernel: partition p's template is its table row (K_p entries min(size, 65535) << 16,
// then 1 << 16 up to the row stride round4(K_p + 1)) repeated for as many rows as fit into `tmpl_bytes` (at most
// max_rows), so that ONE bulk copy of R * row bytes initialises a whole R-row table (size : count = 0 in every entry).
__global__ void __launch_bounds__(256) table_template_kernel(const uint32_t* __restrict__ spack, const int32_t* __restrict__ kcount, int32_t kp,
                                                             int64_t tmpl_words, int32_t max_rows, uint32_t* __restrict__ tmpl) {
    const int p = blockIdx.y;
    const int roww = (kcount[p] + 1 + 3) & ~3;
    const int64_t rows = min((int64_t)max_rows, tmpl_words / roww);
    const uint32_t* src = spack + (int64_t)p * kp;
    uint32_t* dst = tmpl + (int64_t)p * tmpl_words;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * roww; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i % roww];
}
int launch_table_template(const uint32_t* d_spack, const int32_t* d_kcount, int32_t m, int32_t kp, int64_t tmpl_bytes, int32_t max_rows,
                          uint32_t* d_tmpl, cudaStream_t s) {
    table_template_kernel<<<dim3(16, (unsigned)m), 256, 0, s>>>(d_spack, d_kcount, kp, tmpl_bytes / 4, max_rows, d_tmpl);
    CA_CUDA(cudaGetLastError()); ca::ctx().launches++;
    return CA_OK;
}

// ------------------------------------------------------------------------------------------------
// relabel + transpose: raw int32 [m][n] (partition-major)  ->  lab16 uint16 [mpad/16][n + 1][16] (super-tile-major, lab16_index; cell n = pad cell with label K_p)
// A block moves a 32-partition x 128-cell tile through shared memory: coalesced 4-byte reads along n,
// 32-byte writes along the partition axis.
// ------------------------------------------------------------------------------------------------
// Sharded label sets (one process, several GPUs): a device relabels the 32-partition spans it owns -- local block y is
// the global span span_of_block[y] -- and stores every 32-byte sector into the lab16 arrays dst.p[0..n).  The kernel CAN store
// into peer memory over NVLink (n > 1: it is then its own all-gather), but 16-byte peer stores reached only 290 GB/s of egress
// at 8 GPUs, so multi_allpairs (ca_api.cu) passes the local array alone and lets the copy engines push the slabs to the peers.
constexpr int TP = 32, TC = 128;
__global__ void __launch_bounds__(256) relabel_transpose_kernel(const int32_t* __restrict__ raw, int64_t n, int32_t m,
                                                                const int32_t* __restrict__ span_of_block,
                                                                const int32_t* __restrict__ d_min,
                                                                const uint32_t* __restrict__ d_rank,
                                                                const int64_t* __restrict__ d_roff,
                                                                const int32_t* __restrict__ d_kcount,
                                                                const LabDst dst) {
    __shared__ __align__(16) uint16_t tile[TP][TC + 2];
    const int64_t c0 = (int64_t)blockIdx.x * TC;
    const int p0 = blockIdx.y * TP;                                                    // local partitions of this block
    const int g0 = (span_of_block ? span_of_block[blockIdx.y] : (int)blockIdx.y) * TP;  // their global ids
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // every partition's row is 16-byte aligned and the tile is full
    const bool vec = (n & 3) == 0 && c0 + TC <= n && (reinterpret_cast<uintptr_t>(raw) & 15) == 0;
    for (int pp = wid; pp < TP; pp += 8) {
        const int p = p0 + pp;
        if (p < m) {
            const int32_t* x = raw + (int64_t)p * n;
            const int mn = d_min[p];
            const uint32_t* rk = d_rank ? d_rank + d_roff[p] : nullptr;
            if (vec) {  // one 16-byte load per lane: cells c0 + 4 lane .. + 3
                const int4 a = __ldg(reinterpret_cast<const int4*>(x + c0) + lane);
                uint32_t v0 = (uint32_t)(a.x - mn), v1 = (uint32_t)(a.y - mn), v2 = (uint32_t)(a.z - mn), v3 = (uint32_t)(a.w - mn);
                if (rk) {
                    v0 = __ldg(rk + v0);
                    v1 = __ldg(rk + v1);
                    v2 = __ldg(rk + v2);
                    v3 = __ldg(rk + v3);
                }
                uint32_t* t32 = reinterpret_cast<uint32_t*>(&tile[pp][4 * lane]);  // row stride 130 uint16 = 260 bytes: 4-byte aligned
                t32[0] = (v0 & 0xffffu) | (v1 << 16);
                t32[1] = (v2 & 0xffffu) | (v3 << 16);
            } else {
#pragma unroll
                for (int q = 0; q < TC / 32; q++) {
                    int64_t c = c0 + lane + 32 * q;
                    uint32_t v = 0;
                    if (c < n) {
                        v = (uint32_t)(__ldg(x + c) - mn);
                        if (rk) v = __ldg(rk + v);
                    } else if (c == n) {
                        v = (uint32_t)d_kcount[p];  // the pad cell
                    }
                    tile[pp][lane + 32 * q] = (uint16_t)v;
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < TC / 32; q++) tile[pp][lane + 32 * q] = 0;
        }
    }
    __syncthreads();
    const int cell = threadIdx.x >> 1, half = threadIdx.x & 1;
    const int64_t c = c0 + cell;
    if (c <= n) {
        uint32_t wds[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            uint32_t lo = tile[half * 16 + 2 * q][cell], hi = tile[half * 16 + 2 * q + 1][cell];
            wds[q] = lo | (hi << 16);
        }
        const int64_t at = lab16_index(n, c, g0 + half * 16);
        const uint4 w0 = make_uint4(wds[0], wds[1], wds[2], wds[3]), w1 = make_uint4(wds[4], wds[5], wds[6], wds[7]);
        for (int d = 0; d < dst.n; d++) {  // own memory first, then the peers'
            uint4* o = reinterpret_cast<uint4*>(dst.p[d] + at);
            o[0] = w0;
            o[1] = w1;
        }
    }
}

int launch_relabel_transpose(const int32_t* raw, int64_t n, int32_t m, int32_t nblocks, const int32_t* d_span_of_block, const int32_t* d_min,
                             const uint32_t* d_rank, const int64_t* d_roff, const int32_t* d_kcount, const LabDst& dst, cudaStream_t s) {
    if (n == 0 || nblocks <= 0) return CA_OK;
    dim3 grid((unsigned)((n + 1 + TC - 1) / TC), (unsigned)nblocks);  // n cells + the pad cell
    relabel_transpose_kernel<<<grid, 256, 0, s>>>(raw, n, m, d_span_of_block, d_min, d_rank, d_roff, d_kcount, dst);
    CA_CUDA(cudaGetLastError()); ca::ctx().launches++;
    return CA_OK;
}

// ------------------------------------------------------------------------------------------------
// counting sort of the cells of every owned partition by cluster.
// A partition's cells are cut into `nseg` contiguous segments; one BLOCK sorts one (partition, segment):
//   seg_count_kernel : per (partition, segment) histogram of the labels
//   seg_scan_kernel  : cursor[p][seg][k] = cstart[p][k] + #cells of cluster k in earlier segments
//   sort_kernel      : every thread takes a cell, claims the next free position of its cluster with one
//                      shared-memory atomic on the segment's cursor and writes its cell id there.
// The order of the cells INSIDE a cluster is therefore arbitrary (it varies from run to run); nothing downstream
// depends on it: a cell's ECS is a function of its (row, label) table entry only, and the quad a cell shares a
// thread with is irrelevant to its value.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t compact_of(const int32_t* x, int64_t i, int mn, const uint32_t* rk) {
    uint32_t v = (uint32_t)(__ldg(x + i) - mn);
    return rk ? __ldg(rk + v) : v;
}

__global__ void __launch_bounds__(256) seg_count_kernel(const int32_t* __restrict__ raw, int64_t n, const int32_t* __restrict__ d_min,
                                                        const uint32_t* __restrict__ d_rank, const int64_t* __restrict__ d_roff,
                                                        int32_t kp, const int32_t* __restrict__ d_parts, int32_t nseg, int64_t seglen,
                                                        int32_t* __restrict__ d_segcur) {
    extern __shared__ int32_t cnt_sh[];
    const int slot = blockIdx.y, seg = blockIdx.x;
    const int p = d_parts[slot];
    const int32_t* x = raw + (int64_t)p * n;
    const int mn = d_min[p];
    const uint32_t* rk = d_rank ? d_rank + d_roff[p] : nullptr;
    int32_t* out = d_segcur + ((int64_t)slot * nseg + seg) * kp;
    for (int k = threadIdx.x; k < kp; k += blockDim.x) cnt_sh[k] = 0;
    __syncthreads();
    const int64_t lo = (int64_t)seg * seglen, hi = min(n, lo + seglen);
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) atomicAdd(cnt_sh + compact_of(x, i, mn, rk), 1);
    __syncthreads();
    for (int k = threadIdx.x; k < kp; k += blockDim.x) out[k] = cnt_sh[k];
}

__global__ void seg_scan_kernel(const int32_t* __restrict__ d_cstart, int32_t kp, int32_t nparts, int32_t nseg,
                                int32_t* __restrict__ d_segcur) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)nparts * kp) return;
    const int slot = (int)(idx / kp), k = (int)(idx % kp);
    int32_t run = d_cstart[idx];
    for (int seg = 0; seg < nseg; seg++) {
        int32_t* c = d_segcur + ((int64_t)slot * nseg + seg) * kp + k;
        const int32_t t = *c;
        *c = run;
        run += t;
    }
}

template <bool SMEM_CURSOR>
__global__ void __launch_bounds__(512) sort_kernel(const int32_t* __restrict__ raw, int64_t n, const int32_t* __restrict__ d_min,
                                                   const uint32_t* __restrict__ d_rank, const int64_t* __restrict__ d_roff,
                                                   int32_t kp, const int32_t* __restrict__ d_parts, int32_t nseg, int64_t seglen,
                                                   int32_t* __restrict__ d_perm, int64_t perm_stride, int32_t* __restrict__ d_segcur) {
    extern __shared__ int32_t cur_sh[];
    const int slot = blockIdx.y, seg = blockIdx.x;
    const int p = d_parts[slot];
    const int32_t* x = raw + (int64_t)p * n;
    const int mn = d_min[p];
    const uint32_t* rk = d_rank ? d_rank + d_roff[p] : nullptr;
    int32_t* gcur = d_segcur + ((int64_t)slot * nseg + seg) * kp;
    int32_t* cur = SMEM_CURSOR ? cur_sh : gcur;
    if (SMEM_CURSOR) {
        for (int k = threadIdx.x; k < kp; k += blockDim.x) cur[k] = gcur[k];
        __syncthreads();
    }
    int32_t* out = d_perm + (int64_t)slot * perm_stride;
    const int64_t lo = (int64_t)seg * seglen, hi = min(n, lo + seglen);
    constexpr int U = 4;
    for (int64_t base = lo + threadIdx.x; base < hi; base += (int64_t)blockDim.x * U) {
        uint32_t lab[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int64_t i = base + (int64_t)u * blockDim.x;
            lab[u] = i < hi ? compact_of(x, i, mn, rk) : 0u;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int64_t i = base + (int64_t)u * blockDim.x;
            if (i < hi) out[atomicAdd(cur + lab[u], 1)] = (int32_t)i;
        }
    }
}

// The same sort, one chunk of SC_CELLS cells per block, with the chunk sorted in shared memory first:
//   1. every cell's label goes to shared memory and claims its rank inside (chunk, cluster) with a shared-memory atomic;
//   2. an exclusive scan of the per-cluster counts gives the chunk-local start of every cluster, and ONE global atomic per
//      non-empty (chunk, cluster) claims that many consecutive positions of the cluster's segment in perm;
//   3. the cells are written to shared memory in cluster order, then out: adjacent threads write adjacent positions of a
//      cluster's run, so a warp's store touches a handful of 32-byte sectors instead of 32 (the scattered 4-byte stores of
//      sort_kernel are bound by the LSU's sector rate, like the label gather of the wave kernel).
// No per-segment histogram / scan passes: the cursors start at cstart and move by atomics.
constexpr int SC_CELLS = 16384, SC_THREADS = 1024;
__global__ void __launch_bounds__(SC_THREADS, 2) sort_chunk_kernel(const int32_t* __restrict__ raw, int64_t n, const int32_t* __restrict__ d_min,
                                                                const uint32_t* __restrict__ d_rank, const int64_t* __restrict__ d_roff,
                                                                const int32_t* __restrict__ d_kcount, int32_t kp,
                                                                const int32_t* __restrict__ d_parts, int32_t* __restrict__ d_perm,
                                                                int64_t perm_stride, int32_t* __restrict__ d_cursor,
                                                                const uint16_t* __restrict__ d_crow, const int32_t* __restrict__ d_sizes,
                                                                uint2* __restrict__ d_qinfo) {
    extern __shared__ __align__(16) unsigned char sc_smem[];
    uint32_t* cnt = reinterpret_cast<uint32_t*>(sc_smem);                 // [kp] counts, then chunk-local starts
    uint32_t* gbase = cnt + kp;                                           // [kp] first claimed position of the cluster's run
    uint16_t* lab = reinterpret_cast<uint16_t*>(gbase + kp);              // [SC_CELLS] compact label of cell lo + i
    uint16_t* rnk = lab + SC_CELLS;                                       // [SC_CELLS] rank inside (chunk, cluster)
    uint16_t* sidx = rnk + SC_CELLS;                                      // [SC_CELLS] cell index at sorted position t
    __shared__ uint32_t warp_tot[SC_THREADS / 32];
    const int slot = blockIdx.y;
    const int p = d_parts[slot];
    const int32_t* x = raw + (int64_t)p * n;
    const int mn = d_min[p];
    const uint32_t* rk = d_rank ? d_rank + d_roff[p] : nullptr;
    const int K = d_kcount[p];
    const int64_t lo = (int64_t)blockIdx.x * SC_CELLS;
    const int nloc = (int)((n - lo) < (int64_t)SC_CELLS ? (n - lo) : (int64_t)SC_CELLS);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int k = tid; k < K; k += SC_THREADS) cnt[k] = 0;
    __syncthreads();
    for (int i = tid; i < nloc; i += SC_THREADS) {
        const uint32_t l = compact_of(x, lo + i, mn, rk);
        lab[i] = (uint16_t)l;
        rnk[i] = (uint16_t)atomicAdd(cnt + l, 1u);
    }
    __syncthreads();
    // exclusive scan of cnt[0..K): thread t owns the entries [t per, (t + 1) per)
    const int per = (K + SC_THREADS - 1) / SC_THREADS;
    const int k0 = min(K, tid * per), k1 = min(K, k0 + per);
    uint32_t mine = 0;
    for (int k = k0; k < k1; k++) mine += cnt[k];
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    uint32_t run = incl - mine;
    for (int w = 0; w < wid; w++) run += warp_tot[w];
    int32_t* cursor = d_cursor + (int64_t)slot * kp;
    for (int k = k0; k < k1; k++) {
        const uint32_t c = cnt[k];
        cnt[k] = run;
        gbase[k] = c ? (uint32_t)atomicAdd(cursor + k, (int32_t)c) : 0u;
        run += c;
    }
    __syncthreads();
    for (int i = tid; i < nloc; i += SC_THREADS) sidx[cnt[lab[i]] + rnk[i]] = (uint16_t)i;
    __syncthreads();
    // The thread that writes the first position of a quad also writes the quad's info for the wave kernel: {row of the
    // cluster inside its item, cluster size}.  A cluster's segment starts at a multiple of 4 and its padding sits at the
    // tail, so the first position of every quad of the segment holds a real cell: exactly one writer per quad.
    int32_t* out = d_perm + (int64_t)slot * perm_stride;
    uint2* qout = d_qinfo + (int64_t)slot * (perm_stride >> 2);
    const uint16_t* crow = d_crow + (int64_t)slot * kp;
    const int32_t* sizes = d_sizes + (int64_t)p * kp;
    for (int t = tid; t < nloc; t += SC_THREADS) {
        const uint32_t i = sidx[t], l = lab[i];
        const uint32_t pos = gbase[l] + ((uint32_t)t - cnt[l]);
        out[pos] = (int32_t)(lo + i);
        if ((pos & 3u) == 0u) qout[pos >> 2] = make_uint2((uint32_t)__ldg(crow + l), (uint32_t)__ldg(sizes + l));
    }
}

int launch_sort(const int32_t* raw, int64_t n, const int32_t* d_min, const uint32_t* d_rank, const int64_t* d_roff,
                const int32_t* d_kcount, const int32_t* d_cstart, int32_t kp, const int32_t* d_parts, int32_t nparts, int32_t* d_perm,
                int64_t perm_stride, const uint16_t* d_crow, const int32_t* d_sizes, uint32_t* d_qinfo, bool* qinfo_done, cudaStream_t s) {
    *qinfo_done = false;
    if (nparts == 0) return CA_OK;
    CA_CUDA(cudaMemsetAsync(d_perm, 0xff, (size_t)nparts * (size_t)perm_stride * sizeof(int32_t), s));
    DevBuf& segcur = scratch(SCR_SEGCUR);  // per device; batches of one call reuse it in stream order
    {   // chunk sort: counts + run bases ([kp] each) and three uint16 arrays of SC_CELLS entries in shared memory
        const size_t smem = (size_t)kp * 8 + (size_t)SC_CELLS * 6;
        if (smem <= 200 * 1024 && getenv("CA_SORT_LEGACY") == nullptr) {  // two blocks per SM up to ~2100 clusters, one beyond
            CA_TRY(segcur.ensure((size_t)nparts * kp * 4));
            CA_CUDA(cudaMemcpyAsync(segcur.p, d_cstart, (size_t)nparts * kp * 4, cudaMemcpyDeviceToDevice, s));
            CA_CUDA(cudaFuncSetAttribute(sort_chunk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const dim3 grid((unsigned)((n + SC_CELLS - 1) / SC_CELLS), (unsigned)nparts);
            sort_chunk_kernel<<<grid, SC_THREADS, smem, s>>>(raw, n, d_min, d_rank, d_roff, d_kcount, kp, d_parts, d_perm, perm_stride,
                                                             segcur.as<int32_t>(), d_crow, d_sizes, reinterpret_cast<uint2*>(d_qinfo));
            CA_CUDA(cudaGetLastError()); ca::ctx().launches++;
            *qinfo_done = true;
            return CA_OK;
        }
    }
    // very large cluster counts: per-segment cursors, scattered stores
    const size_t per_warp = (size_t)kp * 4;
    const bool smem_ok = per_warp <= 200 * 1024;
    // segments: enough blocks to fill the machine, each at least 16k cells long; the shared-memory count
    // kernel needs the label table in shared memory too, so very large cluster counts stay at one segment
    int32_t nseg = 1;
    if (smem_ok) {
        // at least 32 segments: short segments keep the co-resident blocks' output windows (one short run per cluster and
        // segment) inside L2, so the 4-byte stores of a sector merge there (measured on C5: 5 segments 6.9 ms, 32 5.6 ms)
        int64_t want = std::max<int64_t>((int64_t)ctx().sm_count * 16 / nparts + 1, 32);
        int64_t cap = n / 16384 + 1;
        nseg = (int32_t)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(want, cap), 64));
        if (const char* e = getenv("CA_SORT_NSEG")) nseg = (int32_t)std::max<int64_t>(1, std::min<int64_t>(atoi(e), cap));  // tuning
    }
    const int64_t seglen = ((n + nseg - 1) / nseg + 127) & ~127LL;
    CA_TRY(segcur.ensure((size_t)nparts * nseg * kp * 4));
    if (nseg > 1) {
        if (per_warp > 48 * 1024)
            CA_CUDA(cudaFuncSetAttribute(seg_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)per_warp));
        seg_count_kernel<<<dim3((unsigned)nseg, (unsigned)nparts), 256, per_warp, s>>>(raw, n, d_min, d_rank, d_roff, kp, d_parts, nseg,
                                                                                      seglen, segcur.as<int32_t>());
        const int64_t tot = (int64_t)nparts * kp;
        seg_scan_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(d_cstart, kp, nparts, nseg, segcur.as<int32_t>());
        ca::ctx().launches += 2;
    } else {
        CA_CUDA(cudaMemcpyAsync(segcur.p, d_cstart, (size_t)nparts * kp * 4, cudaMemcpyDeviceToDevice, s));
    }
    const dim3 grid((unsigned)nseg, (unsigned)nparts);
    if (smem_ok) {
        if (per_warp > 48 * 1024)
            CA_CUDA(cudaFuncSetAttribute(sort_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)per_warp));
        sort_kernel<true><<<grid, 512, per_warp, s>>>(raw, n, d_min, d_rank, d_roff, kp, d_parts, nseg, seglen, d_perm, perm_stride,
                                                      segcur.as<int32_t>());
    } else {  // cursors stay in global memory for very large cluster counts
        sort_kernel<false><<<grid, 512, 0, s>>>(raw, n, d_min, d_rank, d_roff, kp, d_parts, nseg, seglen, d_perm, perm_stride,
                                                segcur.as<int32_t>());
    }
    CA_CUDA(cudaGetLastError()); ca::ctx().launches++;
    return CA_OK;
}

}  // namespace ca
